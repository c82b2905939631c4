#!/usr/bin/env python
"""Run the other BASELINE configs on the GPU box and check them against the CPU oracle.

    python tools/run_configs.py c1 c3 c4 [--no-oracle]

  c1  Pk_periodic, 1e6 uniform-random particles (seed 1), Lbox=1000, Ngrid=256          -> full oracle comparison
  c3  Pk_periodic_rsd (rsd=z), ~1e8-particle lognormal catalogue with a z displacement, Ngrid=512 -> full oracle comparison
  c4  Bk_periodic, Ngrid=512, step=2, Ncut=3, Nmax=80 (46 700 triangles, 80 shells)     -> oracle on a subset of triangles
      (the reference algorithm needs 87 GB of float64 shell fields here; the oracle streams a few shells)
  survey  B0_survey (survey geometry): 1e6 data + 5e6 randoms in a cone, Lbox=3600, Ngrid=360, step=3,
      Ncut=3, Nmax=40 -> timing; parity against the oracle at Ngrid=96 on a 1/20 subsample (the oracle's float64 shell
      store does not fit the time budget at 360)
Prints one JSON line per config (timings by CUDA events / wall clock, parity figures)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench                                             # noqa: E402  (catalogue generators)
from pyspectrum_b200 import pyspectrum as pySpec         # noqa: E402

ORACLE = '--no-oracle' not in sys.argv
dev = torch.device('cuda', 0)


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return out, (time.perf_counter() - t0) / reps


def stage_times(pipe, xyz_dev, L, fn_after):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    ev[0].record()
    mesh, sumw = pipe.assign(xyz_dev, 0, None, L)
    ev[1].record()
    half = pipe.mesh_to_delta(mesh, sumw)
    ev[2].record()
    fn_after(half)
    ev[3].record()
    torch.cuda.synchronize()
    return [ev[i].elapsed_time(ev[i + 1]) for i in range(3)]


def c1():
    from oracle import pyspec_oracle as O
    rng = np.random.default_rng(1)
    xyz = rng.uniform(0, 1000, (3, 10 ** 6))
    pk, t = timed(lambda: pySpec.Pk_periodic(xyz, Lbox=1000, Ngrid=256))
    pipe = pySpec.PeriodicPipeline.get(256)
    xd = torch.from_numpy(xyz).to(dev)
    st = stage_times(pipe, xd, 1000., lambda h: pipe.pk_monopole(h, 1000.))
    out = {'config': 'C1 Pk_periodic 1e6 uniform, L=1000, N=256', 'e2e_s': t, 'stage_ms': dict(zip(['assign', 'fft_fcomb', 'binning'], st)),
           'assign_mpart_s': 1e6 / st[0] / 1e3}
    if ORACLE:
        t0 = time.perf_counter()
        ref = O.Pk_periodic(xyz, None, 1000, 256, workers=os.cpu_count())
        out['oracle_s'] = time.perf_counter() - t0
        out['counts_exact'] = bool(np.array_equal(pk['counts'], ref['counts']))
        out['k_max_rel'] = float(np.abs(pk['k'] / ref['k'] - 1).max())
        out['p0k_raw_max_rel'] = float(np.abs((pk['p0k'] + pk['p0k_sn']) / (ref['p0k'] + ref['p0k_sn']) - 1).max())
    print(json.dumps(out), flush=True)


def c3():
    from oracle import pyspec_oracle as O
    L, N = 2600., 512
    xyz_dev = bench.lognormal_catalogue_torch(3, dev, 10 ** 8, L, N)
    Np = xyz_dev.shape[1]
    # anisotropy: sinusoidal z displacement (any anisotropy suffices)
    xyz_dev[2] = (xyz_dev[2] + 12.0 * torch.sin(2 * np.pi * xyz_dev[2] / 130.0) + 7.0 * torch.sin(2 * np.pi * xyz_dev[0] / 90.0)) % L
    pipe = pySpec.PeriodicPipeline.get(N)
    pr, t_dev = timed(lambda: pySpec.Pk_periodic_rsd(xyz_dev, Lbox=L, Ngrid=N, rsd=2, Nmubin=10))
    st = stage_times(pipe, xyz_dev, L, lambda h: pipe.pk_multipoles(h, int(L), 2, 10))
    xyz = xyz_dev.cpu().numpy()
    pr2, t_host = timed(lambda: pySpec.Pk_periodic_rsd(xyz, Lbox=L, Ngrid=N, rsd=2, Nmubin=10), reps=1)
    out = {'config': 'C3 Pk_periodic_rsd rsd=z, %d particles, L=2600, N=512, Nmubin=10' % Np, 'device_resident_s': t_dev,
           'e2e_pageable_host_s': t_host, 'stage_ms': dict(zip(['assign', 'fft_fcomb', 'multipoles'], st)), 'assign_mpart_s': Np / st[0] / 1e3,
           'assign_alg_gbs': (16.0 * Np + 8.0 * N ** 3) / st[0] / 1e6, 'fft_alg_gbs': 44.0 * N ** 3 / st[1] / 1e6}
    del xyz_dev
    if ORACLE:
        t0 = time.perf_counter()
        ref = O.Pk_periodic_rsd(xyz, None, L, N, rsd=2, Nmubin=10, workers=os.cpu_count())
        out['oracle_s'] = time.perf_counter() - t0
        out['counts_exact'] = bool(np.array_equal(pr['counts'], ref['counts']))
        out['counts_kmu_exact'] = bool(np.array_equal(pr['counts_kmu'], ref['counts_kmu']))
        sn = ref['p_sn'][0]
        out['p0k_raw_max_rel'] = float(np.abs((pr['p0k'] + sn) / (ref['p0k'] + sn) - 1).max())
        scale = np.abs(ref['p0k'] + sn)
        out['p2k_max_err_over_p0'] = float((np.abs(pr['p2k'] - ref['p2k']) / scale).max())
        out['p4k_max_err_over_p0'] = float((np.abs(pr['p4k'] - ref['p4k']) / scale).max())
        m = ref['counts_kmu'] > 0
        out['p_kmu_raw_max_rel'] = float(np.abs((pr['p_kmu'] + sn)[m] / (ref['p_kmu'] + sn)[m] - 1).max())
    print(json.dumps(out), flush=True)


def c4():
    from oracle import pyspec_oracle as O
    import scipy.fft as sfft
    L, N, step, Ncut, Nmax = 2600., 512, 2, 3, 80
    xyz_dev = bench.lognormal_catalogue_torch(3, dev, 10 ** 8, L, N)
    Np = xyz_dev.shape[1]
    pipe = pySpec.PeriodicPipeline.get(N)
    t0 = time.perf_counter()
    counts = pipe.counts(Nmax, Ncut, step)
    torch.cuda.synchronize()
    t_counts = time.perf_counter() - t0
    bk, t_dev = timed(lambda: pySpec.Bk_periodic(xyz_dev, Lbox=L, Ngrid=N, step=step, Ncut=Ncut, Nmax=Nmax), reps=2)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record(); mesh, sumw = pipe.assign(xyz_dev, 0, None, L)
    ev[1].record(); half = pipe.mesh_to_delta(mesh, sumw)
    ev[2].record(); fields, sumsq, scales, maxabs = pipe.shell_fields(half, step, 1, Nmax, scaled=True)
    ev[3].record(); sums = pipe.triangle_sums(fields, Nmax, Ncut, step)
    ev[4].record(); torch.cuda.synchronize()
    st = [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
    ntri = len(bk['b123'])
    S = Nmax
    out = {'config': 'C4 Bk_periodic N=512 step=2 Ncut=3 Nmax=80, %d particles' % Np, 'triangles': ntri, 'device_resident_s': t_dev,
           'counts_float64_once_s': t_counts, 'stage_ms': dict(zip(['assign', 'fft_fcomb', 'shell_fields', 'triangles'], st)),
           'shell_alg_gbs_dense_model': 20.0 * N ** 3 * S / st[2] / 1e6,
           'tri_alg_tflops': (S * (S + 1) / 2 + 2.0 * ntri) * N ** 3 / st[3] / 1e9}
    del fields, mesh
    if ORACLE:
        t0 = time.perf_counter()
        xyz = xyz_dev.cpu().numpy()
        delta = O.FFT_periodic(xyz, None, L, N, workers=os.cpu_count())
        dfull = O.reflect_delta(delta, N)
        irk = O.shell_index(N, step)
        shells = [2, 15, 30, 31, 45, 61, 80]
        tris = [(i, j, l) for i in shells for j in shells for l in shells if i >= j >= l and l >= max(i - j, 1)]
        f32 = O.shell_fields(dfull, irk, shells, workers=os.cpu_count())
        ones = np.ones((N,) * 3, dtype=np.complex128)
        f64 = O.shell_fields(ones, irk, shells, workers=os.cpu_count(), dtype=np.float64)
        idx = {tuple(t): k for k, t in enumerate(pySpec.triangle_list(Nmax, Ncut, step).tolist())}
        nbar = Np / L ** 3
        kf = 2 * np.pi / L
        worst, cnt_ok = 0.0, True
        Nk = pipe.shell_mode_counts(step, Nmax)
        for (i, j, l) in tris:
            c_ref = np.rint(np.einsum('i,i,i', f64[i], f64[j], f64[l]) / N ** 3) * float(N) ** 3
            cnt_ok &= bool(c_ref == counts[i - 1, j - 1, l - 1])
            s_ref = O._triple(f32[i], f32[j], f32[l])
            p = [np.einsum('i,i', f32[s].astype(np.float64), f32[s].astype(np.float64)) / N ** 3 / Nk[s] * (2 * np.pi) ** 3 / kf ** 3 for s in (i, j, l)]
            b_ref_raw = s_ref / c_ref * (2 * np.pi) ** 6 / kf ** 6
            k = idx[(i, j, l)]
            b_raw = bk['b123'][k] + bk['b123_sn'][k]
            worst = max(worst, abs(b_raw - b_ref_raw) / abs(b_ref_raw))
            worst_p = abs((bk['p0k1'][k] + bk['p0k_sn']) / p[0] - 1)
        out.update({'oracle_s': time.perf_counter() - t0, 'oracle_triangles_checked': len(tris), 'counts_exact_on_subset': cnt_ok,
                    'b123_raw_max_rel_on_subset': float(worst), 'p0k1_raw_rel_last': float(worst_p)})
    print(json.dumps(out), flush=True)


def survey():
    from oracle import pyspec_oracle as O
    rng = np.random.default_rng(7)

    def cone(n):
        return np.array([rng.uniform(100., 260., n), np.degrees(np.arcsin(rng.uniform(-0.1, 0.9, n))),
                         rng.uniform(0.15 ** 3, 0.6 ** 3, n) ** (1. / 3.)])
    nz = lambda z: 3e-4 * np.exp(-((z - 0.4) / 0.25) ** 2)
    data, rand = cone(10 ** 6), cone(5 * 10 ** 6)
    kw = dict(P0_fkp=1e4, Lbox=3600., step=3, Ncut=3, Nmax=40)
    pySpec.PeriodicPipeline.get(360).counts(40, 3, 3)
    t0 = time.perf_counter()
    from pyspectrum_b200 import util as UT
    UT.radecz_to_cartesian(rand)
    t_host = time.perf_counter() - t0
    bk, t = timed(lambda: pySpec.B0_survey(data, nz(data[2]), radecz_r=rand, nbar_r=nz(rand[2]), Ngrid=360, **kw), reps=3)
    out = {'config': 'B0_survey 1e6 data + 5e6 randoms, L=3600, N=360, step=3, Ncut=3, Nmax=40', 'triangles': len(bk['b123']),
           'e2e_pageable_host_s': t, 'host_radecz_to_cartesian_randoms_s': t_host,
           'finite': bool(np.all(np.isfinite(bk['b123'])) and np.all(np.isfinite(bk['q123'])))}
    if ORACLE:
        d, r = data[:, ::20], rand[:, ::20]
        kw2 = dict(P0_fkp=1e4, Lbox=3600., Ngrid=96, step=2, Ncut=3, Nmax=12)
        t0 = time.perf_counter()
        ref = O.B0_survey(d, nz(d[2]), radecz_r=r, nbar_r=nz(r[2]), workers=os.cpu_count(), **kw2)
        out['oracle_s'] = time.perf_counter() - t0
        got = pySpec.B0_survey(d, nz(d[2]), radecz_r=r, nbar_r=nz(r[2]), **kw2)
        fr = O.FFT_survey_mono(r, nz(r[2]), P0_fkp=1e4, Lbox=3600., Ngrid=96)
        alpha = d.shape[1] / fr[1]
        I12, I13, I22, I23, I33 = [alpha * x for x in fr[2:]]
        sn_b = ((ref['p0k1'] + ref['p0k2'] + ref['p0k3']) * I23 + (1. - alpha ** 2) * I13) / I33
        scale = np.abs(ref['b123'] + sn_b)
        out['idx_exact'] = bool(np.array_equal(got['i_k1'], ref['i_k1']) and np.array_equal(got['i_k3'], ref['i_k3']))
        out['p0k1_raw_max_rel'] = float(np.abs((got['p0k1'] - ref['p0k1']) / (ref['p0k1'] + (1 + alpha) * I12 / I22)).max())
        # delta_d - alpha delta_r cancels the survey window: the float32 mesh rounding is amplified per shell by
        # rms|delta_d| / rms|delta| (tests/test_gpu_survey.py), the tolerance scales with it
        od = O.FFT_survey_mono(d, nz(d[2]), P0_fkp=1e4, Lbox=3600., Ngrid=96)
        hd, hh = np.asarray(od[0]), np.asarray(od[0]) - alpha * np.asarray(fr[0])
        irk = O.shell_index(96, 2)[:49]
        amp = np.ones(13)
        for j in range(1, 13):
            m = irk == j
            amp[j] = max(1., np.sqrt(np.sum(np.abs(hd[m]) ** 2) / np.sum(np.abs(hh[m]) ** 2)))
        tri = O.triangle_list(12, 3, 2)
        amp3 = np.maximum(np.maximum(amp[tri[:, 0]], amp[tri[:, 1]]), amp[tri[:, 2]])
        out['window_amplification_max'] = float(amp.max())
        out['b123_max_err_over_plain_tol'] = float((np.abs(got['b123'] - ref['b123']) / (1e-5 * scale + 1e-6 * scale.max())).max())
        out['b123_max_err_over_amplified_tol'] = float((np.abs(got['b123'] - ref['b123']) / (amp3 * (1e-5 * scale + 1e-6 * scale.max()))).max())
    print(json.dumps(out), flush=True)


if __name__ == '__main__':
    for name in sys.argv[1:]:
        if name in ('c1', 'c3', 'c4', 'survey'):
            {'c1': c1, 'c3': c3, 'c4': c4, 'survey': survey}[name]()
            torch.cuda.empty_cache()
