"""Where does the time go in Bk_periodic_many at the C2 size?  (host timings around each phase)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pyspectrum_b200 import pyspectrum as pySpec
L, N = 2600., 360
dev = torch.device('cuda', 0)
xyz_dev = bench.lognormal_catalogue_torch(1, dev, 10 ** 7, L, N)
Np = xyz_dev.shape[1]
xyz_host = torch.empty((3, Np), dtype=torch.float64, pin_memory=True)
xyz_host.copy_(xyz_dev)
kw = dict(Lbox=L, Ngrid=N, step=3, Ncut=3, Nmax=40)
for _ in range(3):
    pySpec.Bk_periodic(xyz_host, **kw)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(8):
    pySpec.Bk_periodic(xyz_host, **kw)
torch.cuda.synchronize()
print('single calls: %.2f ms/catalogue' % ((time.perf_counter() - t0) / 8 * 1e3))
t0 = time.perf_counter()
for _ in range(8):
    pySpec.Bk_periodic(xyz_dev, **kw)
torch.cuda.synchronize()
print('device-resident calls: %.2f ms/catalogue' % ((time.perf_counter() - t0) / 8 * 1e3))
for rep in range(2):
    t0 = time.perf_counter()
    marks = []
    for out in pySpec.Bk_periodic_many([xyz_host] * 8, **kw):
        marks.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    print('many: %.2f ms/catalogue; yields at' % ((time.perf_counter() - t0) / 8 * 1e3), ['%.1f' % (m * 1e3) for m in marks])
# raw copy timing on a side stream while a kernel-heavy loop runs
cs = torch.cuda.Stream()
t0 = time.perf_counter()
with torch.cuda.stream(cs):
    d = xyz_host.to(dev, non_blocking=True)
t1 = time.perf_counter()
cs.synchronize()
t2 = time.perf_counter()
print('H2D enqueue %.2f ms, complete %.2f ms (%.1f GB/s)' % ((t1 - t0) * 1e3, (t2 - t0) * 1e3, xyz_host.numel() * 8 / (t2 - t0) / 1e9))
