"""GPU diagnostic at BASELINE config 2 size: run-to-run / weight-rescaling consistency of Bk_periodic and engine agreement."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyspectrum_b200 import pyspectrum as pySpec
N, L, Np = 360, 2600., 10 ** 7
import bench
which = sys.argv[1] if len(sys.argv) > 1 else 'lognormal'
if which == 'lognormal':
    xyz = bench.lognormal_catalogue_torch(2, torch.device('cuda', 0), Np, L, N).cpu().numpy()
    Np = xyz.shape[1]
else:
    rng = np.random.default_rng(2)
    npar = Np // 40
    par = rng.uniform(0, L, (3, npar))
    kids = par[:, rng.integers(0, npar, Np // 2)] + rng.normal(0, 0.004 * L, (3, Np // 2))
    xyz = np.ascontiguousarray(np.concatenate([kids, rng.uniform(0, L, (3, Np - Np // 2))], axis=1) % L)
print(which, Np)
def raw(bk): return bk['b123'] + bk['b123_sn']
bk1 = pySpec.Bk_periodic(xyz, Lbox=L, Ngrid=N)
bk1b = pySpec.Bk_periodic(xyz, Lbox=L, Ngrid=N)
bk2 = pySpec.Bk_periodic(xyz, w=np.full(Np, 3.0), Lbox=L, Ngrid=N)
p1, p2, p3 = [bk1[k] + bk1['p0k_sn'] for k in ('p0k1', 'p0k2', 'p0k3')]
sigma = np.sqrt(L ** 3 * p1 * p2 * p3 / bk1['counts'])
for name, other in (('same input again', bk1b), ('weights x3', bk2)):
    d = np.abs(raw(bk1) - raw(other))
    r = d / (np.abs(raw(bk1)) + sigma)
    k = np.argsort(r)[-5:]
    print(name, 'max', r.max(), 'median', np.median(r), ' max d/|B|', (d / np.abs(raw(bk1))).max())
    for t in k:
        print('   tri', bk1['i_k1'][t], bk1['i_k2'][t], bk1['i_k3'][t], 'B', raw(bk1)[t], 'other', raw(other)[t], 'sigma', sigma[t], 'cnt', bk1['counts'][t])
# engines on identical fields
pipe = pySpec.PeriodicPipeline.get(N)
half, _ = pipe.fft_periodic(xyz, None, L)
fields, sumsq, scales, maxabs = pipe.shell_fields(half, 3, 1, 40, scaled=True)
a = pipe.triangle_sums(fields, 40, 3, 3, engine='tc').cpu().numpy()
b = pipe.triangle_sums(fields, 40, 3, 3, engine='fma').cpu().numpy()
a2 = pipe.triangle_sums(fields, 40, 3, 3, engine='tc').cpu().numpy()
print('tc vs fma: max rel', (np.abs(a - b) / np.abs(b)).max(), 'median', np.median(np.abs(a - b) / np.abs(b)), ' tc vs tc', (np.abs(a - a2) / np.abs(a)).max())
print('maxabs', maxabs.view(torch.float32).max().item(), 'scales', scales[:4].tolist())
