"""K5 timing at C2 (20 packed shell pairs at 360^3) for A/B runs of two builds (PSB200_LIB selects the library)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyspectrum_b200 import pyspectrum as P
N = int(os.environ.get('K5_N', 360))
dev = torch.device('cuda', 0)
xyz = bench.lognormal_catalogue_torch(2, dev, 10 ** 7, 2600., min(N, 360))
pipe = P.PeriodicPipeline.get(N)
half, _ = pipe.fft_periodic(xyz, None, 2600.)
step, Nmax = (3, 40) if N != 512 else (2, 80)
sc = pipe.shell_scales(half, step, 1, Nmax)
for _ in range(2):
    pipe.shell_fields(half, step, 1, Nmax, scaled=True, scales=sc)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    f = pipe.shell_fields(half, step, 1, Nmax, scaled=True, scales=sc)
e1.record(); torch.cuda.synchronize()
print('lib=%s N=%d: K5 %.3f ms  (checksum %.6e)' % (os.path.basename(os.environ.get('PSB200_LIB', 'libpsb200.so')), N, e0.elapsed_time(e1) / 5, f[1].sum().item()), flush=True)
