"""K6 timing for row subsets and tile counts on synthetic packed fields (CUDA events).
usage: python tools/k6_rows.py N step Nmax levels  (env PSB_TC_LAYOUT / PSB_TC_MT select the plan; one process per setting)
levels: 'off' = all triangles, or a comma list of coarse sizes: times the triangles LEFT on the top grid N."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pyspectrum_b200 import pyspectrum as P

N, step, Nmax, spec = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
Ncut = 3
s0 = Ncut // step
S = Nmax - s0 + 1
pipe = P.PeriodicPipeline.get(N)
ncell = N ** 3
g = torch.Generator(device='cuda').manual_seed(1)
fields = torch.empty((S, ncell), dtype=torch.float32, device='cuda')
for r in range(S):
    v = torch.randn(ncell // 2, 2, device='cuda', generator=g) * 2.0
    hi = v.half()
    lo = (v - hi.float()).half()
    fields[r] = torch.stack([hi, lo], dim=1).reshape(-1).view(torch.float32)
    del v, hi, lo
fields.psb_packed = True
tri = P.triangle_list(Nmax, Ncut, step)
sub = None
if spec != 'off':
    lev = P.coarse_levels(N, step, tri, [int(x) for x in spec.split(',')])
    assert lev[-1][0] == N
    sub = tri[lev[-1][1]]
for _ in range(2):
    pipe.triangle_sums(fields, Nmax, Ncut, step, engine='tc', tri=sub)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
reps = 5 if ncell < 10 ** 8 else 2
for _ in range(reps):
    pipe.triangle_sums(fields, Nmax, Ncut, step, engine='tc', tri=sub)
e1.record()
torch.cuda.synchronize()
t, NT, passes, layout = pipe.tc_passes(Nmax, Ncut, step, tri=sub)
print('N=%d shells=%d levels=%s triangles=%d pairs=%d layout=%d MT=%d NT=%d passes=%d  %.3f ms  FLUSH=%s' % (
    N, S, spec, len(t), len({(a, b) for a, b, _ in t.tolist()}), layout, passes[0][1], NT, len(passes), e0.elapsed_time(e1) / reps,
    os.environ.get('PSB_TC_FLUSH', '4')), flush=True)
