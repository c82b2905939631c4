"""GPU tests of the slab-decomposed mesh -> delta(k) building blocks of the sharded path (multigpu.slab_*) on ONE
device: the ranks are emulated and the exchanges are slices, so every kernel and the index bookkeeping of the distributed
version run exactly as they would under NCCL (whose collectives are covered on gloo in tests/test_multigpu_host.py).
Reference = the single-GPU K2+K3 (psb_fft_mesh_to_delta), itself pinned to the oracle in tests/test_gpu_parity.py.
Tolerance: 2e-6 of max|delta| (the slab path rebuilds F(k) from the separated spectra: one more float32 rounding),
including the self-conjugate planes where the Fortran's last write wins."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def mods():
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pyspectrum_b200 import pyspectrum as pySpec, multigpu
    return pySpec, multigpu


def _mesh(pySpec, N, Np, L, seed):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(0, L, (3, Np))
    xyz[:, :Np // 2] = (xyz[:, :Np // 2] * 0.25 + 0.3 * L) % L
    w = rng.uniform(0.5, 2., Np)
    pipe = pySpec.PeriodicPipeline.get(N)
    pos, aos, wt = pipe.to_device(xyz, w)
    mesh, sumw = pipe.assign(pos, aos, wt, L)
    return pipe, mesh, sumw


@pytest.mark.parametrize('N', [24, 36, 40, 64])
def test_slab_passes_compose_to_the_3d_transform(mods, N):
    import torch
    pySpec, M = mods
    pipe = pySpec.PeriodicPipeline.get(N)
    x = torch.randn((N, N, N, 2), device='cuda', dtype=torch.float32)
    ref = x.clone()
    pySpec.check(pipe.L.psb_fft_c2c_3d(pySpec._ptr(ref), N, 1, pySpec._ptr(pipe.tw32), pySpec._stream()), 'psb_fft_c2c_3d')
    for nz in (N, N // 2, N // 4):                                   # slabs of different thickness
        a = x.clone()
        for z0 in range(0, N, nz):
            s = a[z0:z0 + nz]
            pySpec.check(pipe.L.psb_fft_slab_xy(pySpec._ptr(s), N, nz, 1, pySpec._ptr(pipe.tw32), pySpec._stream()), 'psb_fft_slab_xy')
        # z pass on ky-slabs [N][ny][N] cut out of the cube
        for y0 in range(0, N, nz):
            t = a[:, y0:y0 + nz].contiguous()
            pySpec.check(pipe.L.psb_fft_slab_z(pySpec._ptr(t), N, nz, N, 1, pySpec._ptr(pipe.tw32), pySpec._stream()), 'psb_fft_slab_z')
            a[:, y0:y0 + nz] = t
        assert (a - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()


@pytest.mark.parametrize('N,world', [(24, 1), (24, 2), (24, 4), (36, 3), (40, 2), (64, 8), (360, 2)])
@pytest.mark.parametrize('periodic', [1, 0])
def test_slab_pipeline_matches_single_gpu_delta(mods, N, world, periodic):
    pySpec, M = mods
    if N == 360 and not periodic:
        pytest.skip('one large case is enough')
    pipe, mesh, sumw = _mesh(pySpec, N, 20000 if N < 360 else 2000000, 100., N + world)
    ref = pipe.mesh_to_delta(mesh.clone(), sumw, periodic=periodic)
    got = M.slab_mesh_to_delta_emulated(pipe, mesh, sumw, world, periodic=periodic)
    scale = ref.abs().max().item()
    err = (got - ref).abs()
    assert err.max().item() <= 2e-6 * scale
    h = N // 2
    for sl in (err[:, :, 0], err[:, :, h], err[:, h], err[h]):      # the planes where images coincide (last write wins)
        assert sl.max().item() <= 2e-6 * scale


def test_slab_single_rank_distributed_entry_point(mods):
    pySpec, M = mods
    pipe, mesh, sumw = _mesh(pySpec, 32, 20000, 100., 5)
    ref = pipe.mesh_to_delta(mesh.clone(), sumw)
    got = M.slab_mesh_to_delta(pipe, mesh.clone(), sumw)             # no process group: world = 1
    assert (got - ref).abs().max().item() <= 2e-6 * ref.abs().max().item()
    with pytest.raises(ValueError):
        M.slab_geometry(32, 5)
