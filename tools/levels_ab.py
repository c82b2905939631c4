"""A/B of the coarse-grid level sets for the shell/triangle stage on one GPU (CUDA events, device-resident catalogue).
usage: python tools/levels_ab.py c2|c4|c5like  spec1 spec2 ...   (spec = PSB_BK_LEVELS value, e.g. off 256 256,320)"""
import json
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from pyspectrum_b200 import pyspectrum as P

cfg = {'c2': dict(N=360, step=3, Ncut=3, Nmax=40, Np=10 ** 7), 'c4': dict(N=512, step=2, Ncut=3, Nmax=80, Np=10 ** 7),
       'c5like': dict(N=1024, step=3, Ncut=3, Nmax=40, Np=2 * 10 ** 7)}[sys.argv[1]]
N, step, Ncut, Nmax = cfg['N'], cfg['step'], cfg['Ncut'], cfg['Nmax']
dev = torch.device('cuda', 0)
xyz = bench.lognormal_catalogue_torch(2, dev, cfg['Np'], 2600., min(N, 360))
pipe = P.PeriodicPipeline.get(N)
half, sumw = pipe.fft_periodic(xyz, None, 2600.)
ref = None
for spec in sys.argv[2:]:
    os.environ['PSB_BK_LEVELS'] = spec
    tri, lev = pipe.bk_levels(step, Ncut, Nmax)
    for _ in range(2):
        pipe.bispectrum_finish(pipe.bispectrum_launch(half, step, Ncut, Nmax))
    torch.cuda.synchronize()
    acc = {'shell_fields': 0., 'triangles': 0.}
    per = {}
    reps = 5 if N < 500 else 3
    t0 = time.perf_counter()
    for _ in range(reps):
        tm = []
        out = pipe.bispectrum_finish(pipe.bispectrum_launch(half, step, Ncut, Nmax, timers=tm))
        for k, (name, a, b) in enumerate(tm):
            acc[name] += a.elapsed_time(b)
            g = lev[k // 2][0].N
            per[(g, name)] = per.get((g, name), 0.) + a.elapsed_time(b)
    wall = (time.perf_counter() - t0) / reps * 1e3
    s = out[0]
    if ref is None:
        ref = s
    print(json.dumps({'config': sys.argv[1], 'levels': spec, 'grids': [(pc.N, len(idx), int(smax)) for pc, idx, _, smax in lev],
                      'shell_fields_ms': acc['shell_fields'] / reps, 'triangles_ms': acc['triangles'] / reps, 'wall_ms': wall,
                      'per_level_ms': {'%d/%s' % k: v / reps for k, v in per.items()},
                      'max_rel_vs_first': float(np.abs((s - ref) / (np.abs(ref) + 1e-3 * np.abs(ref).max())).max())}), flush=True)
