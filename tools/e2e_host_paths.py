"""Drop-in call from pageable numpy memory vs pinned memory (C2 catalogue; Pk_periodic_rsd at C3 size with E2E_C3=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from pyspectrum_b200 import pyspectrum as pySpec
dev = torch.device('cuda', 0)
def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
xyz = bench.lognormal_catalogue_torch(2, dev, 10 ** 7, 2600., 360).cpu().numpy()
pin = torch.from_numpy(xyz).pin_memory()
kw = dict(Lbox=2600., Ngrid=360, step=3, Ncut=3, Nmax=40)
pySpec.Bk_periodic(xyz, **kw)
print('C2 Bk_periodic  pinned first: %.2f ms' % timed(lambda: pySpec.Bk_periodic(pin, **kw)), flush=True)
print('C2 Bk_periodic  pageable numpy float64: %.2f ms   pinned float64: %.2f ms   pageable float32: %.2f ms' % (
    timed(lambda: pySpec.Bk_periodic(xyz, **kw)), timed(lambda: pySpec.Bk_periodic(pin, **kw)),
    timed(lambda: pySpec.Bk_periodic(xyz.astype(np.float32), **kw))), flush=True)
if os.environ.get('E2E_C3'):
    x3 = bench.lognormal_catalogue_torch(3, dev, 10 ** 8, 2600., 512).cpu().numpy()
    p3 = torch.from_numpy(x3).pin_memory()
    k3 = dict(Lbox=2600., Ngrid=512, rsd=2, Nmubin=10)
    print('C3 Pk_periodic_rsd  pageable numpy float64: %.1f ms   pinned: %.1f ms' % (
        timed(lambda: pySpec.Pk_periodic_rsd(x3, **k3), 3), timed(lambda: pySpec.Pk_periodic_rsd(p3, **k3), 3)), flush=True)
# one-piece uploads (to_device -> upload): a 1e8-particle shard as the sharded path / survey path would pass it
pipe = pySpec.PeriodicPipeline.get(360)
big = np.random.default_rng(0).uniform(0, 2600., (3, 5 * 10 ** 7))
for flag in ('1', '0'):
    os.environ['PSB_HOST_STAGING'] = flag
    pipe.to_device(big); torch.cuda.synchronize()
    t0 = time.perf_counter(); pos, aos, wt = pipe.to_device(big); torch.cuda.synchronize()
    print('to_device 1.2 GB pageable, staging %s: %.1f ms (%.1f GB/s)  equal %s' % (flag, (time.perf_counter() - t0) * 1e3, 1.2 / (time.perf_counter() - t0),
          bool(torch.equal(pos.cpu(), torch.from_numpy(big)))), flush=True)
