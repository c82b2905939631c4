#!/usr/bin/env python
"""K6 ablation timing at the C2 shape (N=360, 40 shells) on synthetic packed fields: PSB_TC_DEBUG bits switch pieces of the
kernel off (results are then wrong by construction; this only measures where the time goes).
    python tools/ablate_k6.py [debug values...]"""
import os
import sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspectrum_b200 import pyspectrum as pySpec

N, Nmax, Ncut, step = 360, 40, 3, 3
pipe = pySpec.PeriodicPipeline.get(N)
ncell = N ** 3
g = torch.Generator(device='cuda').manual_seed(1)
fields = torch.empty((40, ncell), dtype=torch.float32, device='cuda')
for r in range(40):
    v = torch.randn(ncell // 2, 2, device='cuda', generator=g) * 2.0
    hi = v.half()
    lo = (v - hi.float()).half()
    fields[r] = torch.stack([hi, lo], dim=1).reshape(-1).view(torch.float32)        # [pair][hi2|lo2]
    del v, hi, lo
fields.psb_packed = True
vals = [int(a) for a in sys.argv[1:]] or [0]
for dbg in vals:
    os.environ['PSB_TC_DEBUG'] = str(dbg)
    for _ in range(2):
        pipe.triangle_sums(fields, Nmax, Ncut, step, engine='tc')
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        pipe.triangle_sums(fields, Nmax, Ncut, step, engine='tc')
    e1.record()
    torch.cuda.synchronize()
    print('debug=%4d  %.3f ms' % (dbg, e0.elapsed_time(e1) / 5), flush=True)
    if os.environ.get('ABLATE_TRACE'):
        os.environ['PSB_TC_TRACE'] = 'gpurun_out/trace_dbg%d.bin' % dbg
        pipe.triangle_sums(fields, Nmax, Ncut, step, engine='tc')
        torch.cuda.synchronize()
        del os.environ['PSB_TC_TRACE']
