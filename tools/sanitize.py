"""Small end-to-end run of every hand-written kernel family for compute-sanitizer (memcheck / racecheck / synccheck):
    compute-sanitizer --tool memcheck  python tools/sanitize.py
    compute-sanitizer --tool racecheck python tools/sanitize.py
CI-size grids: K1 (tile scatter and per-particle scatter, full grid and slab), route kernels, K2+K3, slab FFT, K4 (all modes),
K5 (one grid, a coarser transform grid, routed output), the peer-store particle scatter, K6 (tcgen05 kernel, both lane layouts, and the FFMA kernel), routed slab-FFT
separation, survey pre-step.  PSB_ASSIGN_TWOPASS=5 in the environment also runs K1's two-pass sort kernels."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyspectrum_b200 import pyspectrum as P, multigpu as M

NO_TC = bool(os.environ.get('SANITIZE_NO_TC'))          # racecheck stops recording after the tcgen05 kernel's (false) hazards: run the rest alone
rng = np.random.default_rng(0)
L = 200.
for N, Np in [(32, 30000), (48, 3000)]:                     # dense -> tile scatter, sparse -> per-particle scatter
    xyz = rng.uniform(0, L, (3, Np))
    xyz[:, :Np // 2] = (xyz[:, :Np // 2] * 0.2 + 50.) % L
    w = rng.uniform(0.5, 2., Np)
    pk = P.Pk_periodic(xyz, w=w, Lbox=L, Ngrid=N)
    pr = P.Pk_periodic_rsd(xyz, Lbox=L, Ngrid=N, rsd=1, Nmubin=5)
    pp = P.Pk_periodic_rsd(xyz, Lbox=L, Ngrid=N, rsd=2, Nmubin=5, code='python')
    if NO_TC:
        continue
    bk = P.Bk_periodic(xyz, w=w, Lbox=L, Ngrid=N, step=1, Ncut=1, Nmax=10)            # tcgen05, <= 128 pair rows: plain rows
    print('N=%d: Pk %.4e  Pk_rsd %.4e  Bk %.4e (%d triangles)' % (N, pk['p0k'][2], pr['p2k'][2], bk['b123'][5], len(bk['b123'])))
pipe = P.PeriodicPipeline.get(64)
xyz = rng.uniform(0, L, (3, 100000))
half, _ = pipe.fft_periodic(xyz, None, L)
fields, sumsq, scales, maxabs = pipe.shell_fields(half, 1, 1, 30, scaled=True)
for lay in (() if NO_TC else ('1', '0')):
    os.environ['PSB_TC_LAYOUT'] = lay
    s = pipe.triangle_sums(fields, 30, 1, 1, engine='tc')                              # 2x2 blocks, then one-i-per-lane
    print('K6 tc layout', lay, float(s[100]))
print('K6 fma', float(pipe.triangle_sums(fields, 30, 1, 1, engine='fma')[100]))
os.environ['PSB_BK_LEVELS'] = '32'                                                      # a coarser transform grid for the low triangles
if not NO_TC:
    print('levels', [pc.N for pc, _, _, _ in pipe.bk_levels(1, 1, 14)[1]], float(pipe.bispectrum_sums(half, 1, 1, 14)[0][50]))
for world in (2, 4):                                                                    # slab path, emulated ranks
    pos, aos, wt = pipe.to_device(xyz, None)
    counts, sw = M.route_counts(pipe, pos, aos, wt, L, world)
    send = M.route_scatter(pipe, pos, aos, wt, L, world, counts)
    c = np.concatenate([[0], np.cumsum(counts.cpu().numpy())])
    nz = 64 // world
    slab = M.assign_slab(pipe, send[c[1]:c[2]].contiguous(), nz, nz, L)
    mesh, sumw = pipe.assign(pos, aos, wt, L)
    d = M.slab_mesh_to_delta_emulated(pipe, mesh, sumw, world)
    ny = 64 // world
    out = M.slab_pk_monopole(pipe, d[:, :ny].contiguous(), 0, L)
    car = M.low_k_carrier(pipe, d[:, :ny].contiguous(), 0, 32)
    print('slab world', world, float(slab.sum()), float(out[1]), float(car.abs().sum()))
# peer-store kernels with emulated ranks (the "peers" are buffers of this device): routed K5 z pass and routed slab-FFT separation
world = 2
half32, _ = pipe.fft_periodic(xyz, None, L)
scales = pipe.shell_scales(half32, 1, 1, 12)
deal, per = M.pair_assignment(6, world)
ranks = M.SlabBuffers.emulated(pipe.dev, world, world * 2 * per, 64 ** 3 // world)
for r in range(world):
    pipe.shell_fields(half32, 1, 1, 12, scaled=True, pairs=deal[r], scales=scales, routed=(ranks[r].route(per, deal[r], 12, pipe.dev), 64 // world, world))
nz, hp = M.slab_geometry(64, world)
fr = M.SlabBuffers.emulated(pipe.dev, world, 2, 64 * nz * hp * 2)
for r in range(world):
    M.slab_phase1_routed(pipe, mesh[r * nz:(r + 1) * nz].clone(), fr[r], r * nz)
# peer-store particle scatter: this device's buffers stand in for the receive buffers of two ranks
cnts, _ = M.route_counts(pipe, pos, aos, wt, L, world)
rb = [torch.zeros((int(cnts[d].item()) + 8, 4), dtype=torch.float32, device=pipe.dev) for d in range(world)]
dest = torch.tensor([b.data_ptr() for b in rb], dtype=torch.int64, device=pipe.dev)
cur = torch.zeros(world, dtype=torch.int64, device=pipe.dev)
P.check(pipe.L.psb_slab_route_scatter_peer(P._ptr(pos), int(pos.dtype == torch.float64), aos, None, 0, pos.shape[1], 64, float(L),
                                           np.float32(64 / L), np.float32(0.), 64 // world, world, P._ptr(dest), P._ptr(cur), P._stream()),
        'psb_slab_route_scatter_peer')
print('routed', float(ranks[0].local.abs().sum()), float(fr[1].local.abs().sum()), cur.cpu().tolist() == cnts.cpu().tolist())
radecz = np.stack([rng.uniform(100, 140, 20000), rng.uniform(-5, 30, 20000), rng.uniform(0.2, 0.5, 20000)])
d, Ntot, I12, I13, I22, I23, I33 = P.FFT_survey_mono(radecz, np.full(20000, 3e-4), Lbox=3000., Ngrid=48)
print('survey', Ntot, I22)
torch.cuda.synchronize()
print('sanitize run complete')
