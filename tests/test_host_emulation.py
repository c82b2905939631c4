"""The __host__ __device__ cores of the CUDA kernels (psb_fft_core.cuh, psb_fcomb_core.cuh) compiled with g++ and checked on
the CPU: small DFT codelets incl. the composite 16/18/20/32-point ones, the runtime radix plans, the two-stage production plans
(360 = 20 x 18, 256 = 16 x 16) in the exact form the fused kernel runs them, and the closed form of fcomb against the oracle."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'host_emu', 'emu.cpp')
SO = os.path.join(HERE, 'host_emu', 'libpsb_emu.so')


@pytest.fixture(scope='module')
def emu():
    deps = [SRC] + [os.path.join(HERE, '..', 'pyspectrum_b200', 'csrc', f) for f in ('psb_fft_core.cuh', 'psb_fcomb_core.cuh', 'psb_common.cuh')]
    if not os.path.isfile(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(['g++', '-O1', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off', '-o', SO, SRC])
    return ctypes.CDLL(SO)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize('R', [2, 3, 4, 5, 8, 9, 16, 18, 20, 32])
def test_dft_codelets(emu, R):
    rng = np.random.default_rng(R)
    x = rng.normal(size=R) + 1j * rng.normal(size=R)
    for d in (+1, -1):
        v = np.ascontiguousarray(x.astype(np.complex128))
        assert emu.emu_dft(_ptr(v), R, d) == 0
        ref = np.fft.ifft(x) * R if d > 0 else np.fft.fft(x)
        assert np.abs(v - ref).max() < 1e-13 * R


@pytest.mark.parametrize('N', [24, 32, 36, 48, 64, 100, 128, 256, 360, 512, 1024])
def test_runtime_plans(emu, N):
    rng = np.random.default_rng(N)
    x = (rng.normal(size=N) + 1j * rng.normal(size=N)).astype(np.complex128)
    rad = (ctypes.c_int * 16)()
    ns = emu.emu_plan(N, rad)
    assert ns > 0 and np.prod(list(rad)[:ns]) == N
    for d in (+1, -1):
        v = x.copy()
        assert emu.emu_fft_f64(_ptr(v), N, d, 1) == 0
        ref = np.fft.ifft(x) * N if d > 0 else np.fft.fft(x)
        assert np.abs(v - ref).max() < 1e-12 * N
    v32 = x.astype(np.complex64)
    assert emu.emu_fft_f32(_ptr(v32), N, -1, 1) == 0
    assert np.abs(v32 - np.fft.fft(x)).max() < 2e-6 * np.abs(np.fft.fft(x)).max() * np.log2(N)


@pytest.mark.parametrize('N', [256, 360])
def test_two_stage_production_plans(emu, N):
    rng = np.random.default_rng(N + 1)
    x = (rng.normal(size=N) + 1j * rng.normal(size=N)).astype(np.complex128)
    for d in (+1, -1):
        out = np.empty_like(x)
        assert emu.emu_two_stage_f64(_ptr(x), _ptr(out), N, d) == 0
        ref = np.fft.ifft(x) * N if d > 0 else np.fft.fft(x)
        assert np.abs(out - ref).max() < 1e-12 * N
        x32 = x.astype(np.complex64)
        o32 = np.empty_like(x32)
        assert emu.emu_two_stage_f32(_ptr(x32), _ptr(o32), N, d) == 0
        assert np.abs(o32 - ref).max() < 3e-6 * np.abs(ref).max() * np.log2(N)


@pytest.mark.parametrize('N,periodic', [(8, 1), (12, 1), (16, 0)])
def test_fcomb_closed_form_matches_sequential_oracle(emu, N, periodic):
    from oracle import pyspec_oracle as O
    rng = np.random.default_rng(N)
    full = (rng.normal(size=(N, N, N)) + 1j * rng.normal(size=(N, N, N))).astype(np.complex64)      # [iz,iy,ix] == Fortran (ix,iy,iz)
    sumw = 37.0
    half = np.zeros((N, N, N // 2 + 1), np.complex64)
    emu.emu_fcomb.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_int]
    emu.emu_fcomb(_ptr(full), _ptr(half), N, ctypes.c_float(sumw), periodic)
    ref = np.asfortranarray(full.transpose(2, 1, 0).copy())                                      # (ix,iy,iz) Fortran order
    if periodic:
        O.fcomb_periodic(ref, sumw, N)
    else:
        O.fcomb_survey(ref, N)
    ref_half = ref[:N // 2 + 1, :, :].transpose(2, 1, 0)                                         # -> [iz,iy,ix]
    assert np.array_equal(half.view(np.uint32), np.ascontiguousarray(ref_half).view(np.uint32))


@pytest.mark.parametrize('N,world', [(12, 1), (12, 2), (24, 4), (36, 3)])
@pytest.mark.parametrize('periodic', [1, 0])
def test_slab_decomposed_fcomb_matches_oracle(emu, N, world, periodic):
    """The sharded path's slab pipeline (multigpu.slab_mesh_to_delta: x,y passes per z-slab, separation of the two grids' spectra,
    z-slabs -> ky-slabs, z pass, fcomb rebuilt from P and Q) with numpy FFTs standing in for the line-FFT kernels and the two
    element-wise kernels emulated: equals the oracle's sequential fcomb on the full transform to float32 rounding on EVERY mode,
    the self-conjugate planes (last write wins) included."""
    from pyspectrum_b200.multigpu import slab_geometry, z_to_y_chunks
    from oracle import pyspec_oracle as O
    import torch
    rng = np.random.default_rng(N + world)
    L, Np = 50., 3000
    xyz = rng.uniform(0, L, (3, Np))
    xyz[:, :Np // 2] = (xyz[:, :Np // 2] * 0.3 + 10.) % L
    w = rng.uniform(0.5, 2., Np)
    sumw = float(np.sum(w))
    mesh = O.assign_mesh(xyz, w, L, N)                                            # (2N,N,N) Fortran order
    ref = O._FFT(mesh, N)
    if periodic:
        O.fcomb_periodic(ref, sumw, N)
    else:
        O.fcomb_survey(ref, N)
    ref_half = np.ascontiguousarray(ref[:N // 2 + 1].transpose(2, 1, 0))          # [kz][ky][kx]
    d = (np.ascontiguousarray(mesh[0::2].transpose(2, 1, 0)) + 1j * np.ascontiguousarray(mesh[1::2].transpose(2, 1, 0))).astype(np.complex64)
    nz, hp = slab_geometry(N, world)
    emu.emu_slab_split.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 3
    emu.emu_slab_fcomb.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 4 + [ctypes.c_float, ctypes.c_int]
    pq = []
    for g in range(world):                                                        # phase 1 on every emulated rank
        s = d[g * nz:(g + 1) * nz]
        D = np.ascontiguousarray(np.fft.ifft2(s, axes=(1, 2), norm='forward').astype(np.complex64))
        p = np.zeros((nz, N, hp), np.complex64)
        q = np.zeros((nz, N, hp), np.complex64)
        emu.emu_slab_split(_ptr(D), _ptr(p), _ptr(q), N, nz, hp)
        pq.append((p, q))
    half = np.zeros((N, N, N // 2 + 1), np.complex64)
    for r in range(world):                                                        # exchange (the product's own chunking) + phase 2
        def gather(i):
            parts = [z_to_y_chunks(torch.from_numpy(pq[g][i].view(np.float32).reshape(nz, N, hp, 2)), world)[r].numpy() for g in range(world)]
            a = np.concatenate(parts, axis=0)                                     # [N z][ny][hp][2]
            return np.ascontiguousarray(a).view(np.complex64)[..., 0]
        py = np.ascontiguousarray(np.fft.ifft(gather(0), axis=0, norm='forward').astype(np.complex64))
        qy = np.ascontiguousarray(np.fft.ifft(gather(1), axis=0, norm='forward').astype(np.complex64))
        out = np.zeros((N, nz, N // 2 + 1), np.complex64)
        emu.emu_slab_fcomb(_ptr(py), _ptr(qy), _ptr(out), N, r * nz, nz, hp, ctypes.c_float(sumw), periodic)
        half[:, r * nz:(r + 1) * nz] = out
    scale = np.abs(ref_half).max()
    assert np.abs(half - ref_half).max() <= 2e-6 * scale
    h = N // 2
    for plane in (half[:, :, 0] - ref_half[:, :, 0], half[:, :, h] - ref_half[:, :, h], half[:, h] - ref_half[:, h], half[h] - ref_half[h]):
        assert np.abs(plane).max() <= 2e-6 * scale
