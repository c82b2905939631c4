"""K1 (assignment) timing at the C2 and C3 shapes; PSB_ASSIGN_VARIANT=0/1/2 selects the scatter kernel (default 2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyspectrum_b200 import pyspectrum as pySpec
dev = torch.device('cuda', 0)
shapes = ((360, 10 ** 7), (512, 10 ** 8)) if os.environ.get('K1_SHAPES', 'all') == 'all' else ((360, 10 ** 7),)
for (N, Np) in shapes:
    xyz = bench.lognormal_catalogue_torch(3, dev, Np, 2600., min(N, 512))
    pipe = pySpec.PeriodicPipeline.get(N)
    for _ in range(2):
        mesh, sumw = pipe.assign(xyz, 0, None, 2600.)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        mesh, sumw = pipe.assign(xyz, 0, None, 2600.)
    e1.record(); torch.cuda.synchronize()
    print('N=%d Np=%d: K1 %.3f ms  (mesh sum %.6e, sumw %.1f)' % (N, xyz.shape[1], e0.elapsed_time(e1) / 5, mesh.double().sum().item(), sumw.item()), flush=True)
    del xyz, mesh
