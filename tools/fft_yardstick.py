#!/usr/bin/env python
"""cuFFT (torch.fft) as the speed yardstick for the hand-written 3-D c2c transform: same in-place complex64 cube."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrum_b200 import pyspectrum as P
for N in [int(a) for a in sys.argv[1:]] or [360, 512]:
    pipe = P.PeriodicPipeline.get(N)
    x = torch.randn((N, N, N, 2), device='cuda', dtype=torch.float32)
    xc = torch.view_as_complex(x.clone())
    def ours():
        P.check(pipe.L.psb_fft_c2c_3d(P._ptr(x), N, 1, P._ptr(pipe.tw32), P._stream()), 'fft')
    def cufft():
        return torch.fft.ifftn(xc, norm='forward')
    res = {}
    for name, fn in (('psb_fft_c2c_3d', ours), ('cufft_c2c_out_of_place', cufft)):
        for _ in range(3): fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20): fn()
        e1.record(); torch.cuda.synchronize()
        res[name] = e0.elapsed_time(e1) / 20
    print('N=%d  ours %.3f ms (%.0f GB/s on the 48 N^3 three-pass model)  cuFFT %.3f ms' % (N, res['psb_fft_c2c_3d'], 48.0 * N ** 3 / res['psb_fft_c2c_3d'] / 1e6, res['cufft_c2c_out_of_place']), flush=True)
