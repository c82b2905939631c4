"""One K5 pass over all shell pairs at the C2 shape on a random half field (profiling driver)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrum_b200 import pyspectrum as pySpec
N = int(os.environ.get('K5_N', 360)); Nmax = int(os.environ.get('K5_NMAX', 40)); step = int(os.environ.get('K5_STEP', 3))
pipe = pySpec.PeriodicPipeline.get(N)
half = torch.randn((N, N, N // 2 + 1, 2), device='cuda', dtype=torch.float32)
for it in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pipe.shell_fields(half, step, 1, Nmax, scaled=True)
    e1.record(); torch.cuda.synchronize()
    print('K5 all pairs: %.3f ms' % e0.elapsed_time(e1), flush=True)
    del out
