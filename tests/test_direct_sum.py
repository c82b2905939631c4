"""Restatement-independent pin of a1+a2+a3: the closed-form interlaced PCS delta(k) (tests/direct_sum.py) evaluated by direct
summation over particles and aliases, against (CPU) the oracle's C restatement of estimator.f and (GPU) the CUDA path."""
import numpy as np
import pytest

import direct_sum as DS


def _cat(seed, Np, L):
    rng = np.random.default_rng(seed)
    xyz = rng.uniform(0, L, (3, Np))
    xyz[:, :Np // 3] = (xyz[:, :Np // 3] * 0.15 + 0.4 * L) % L          # a clump
    xyz[:, 0] = [0., L * (1 - 1e-6), 0.5 * L]                           # faces: the stencil wraps
    return xyz, rng.uniform(0.5, 2., Np)


@pytest.mark.parametrize('N,Np,weighted', [(12, 1, False), (12, 300, True), (16, 500, False), (20, 400, True)])
def test_oracle_delta_equals_direct_sum(N, Np, weighted):
    from oracle import pyspec_oracle as O
    L = 100.
    xyz, w = _cat(N + Np, Np, L)
    w = w if weighted else None
    kv = DS.test_wavevectors(N, seed=N)
    got = DS.pick(O.FFT_periodic(xyz, w, L, N), N, kv)
    ref = DS.direct_delta(xyz, w, L, N, kv)
    assert abs(got[0] - 1.) < 1e-6                                       # delta(k=0) = 1
    assert np.abs(got - ref).max() <= 1e-6 * np.abs(ref).max(), np.abs(got - ref).max()    # measured 1e-7: float32 mesh + FFT rounding


@pytest.mark.gpu
@pytest.mark.parametrize('N,Np,weighted', [(12, 300, True), (16, 500, False), (24, 2000, True)])
def test_cuda_delta_equals_direct_sum(N, Np, weighted):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from pyspectrum_b200 import pyspectrum as pySpec
    L = 100.
    xyz, w = _cat(N + Np, Np, L)
    w = w if weighted else None
    kv = DS.test_wavevectors(N, seed=N)
    got = DS.pick(pySpec.FFT_periodic(xyz, w=w, Lbox=L, Ngrid=N), N, kv)
    ref = DS.direct_delta(xyz, w, L, N, kv)
    assert np.abs(got - ref).max() <= 2e-6 * np.abs(ref).max(), np.abs(got - ref).max()
