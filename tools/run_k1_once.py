"""K1 (assignment) timing; PSB_ASSIGN_VARIANT selects the scatter kernel (3 = tile scatter in shared memory, default;
2 = per-particle vector reductions).  K1_SHAPES = comma list of N:Np (default: C1, C2, C3 shapes).
With K1_CHECK=1 the mesh is compared with the variant-2 mesh of a fresh process (written to / read from gpurun_out/)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from pyspectrum_b200 import pyspectrum as pySpec
dev = torch.device('cuda', 0)
spec = os.environ.get('K1_SHAPES', '256:1000000,360:10000000,512:100000000')
var = os.environ.get('PSB_ASSIGN_VARIANT', '3')
for item in spec.split(','):
    N, Np = [int(v) for v in item.split(':')]
    if Np <= 2 * 10 ** 8:
        xyz = bench.lognormal_catalogue_torch(3, dev, Np, 2600., min(N, 512))
    else:
        xyz = bench.c5_shard(dev, 0, 1, Np, 2600.)
    if os.environ.get('K1_SHUFFLE'):                       # the generators emit cell-ordered particles: destroy the order
        xyz = xyz[:, torch.randperm(xyz.shape[1], device=dev)].contiguous()
    pipe = pySpec.PeriodicPipeline.get(N)
    for _ in range(2):
        mesh, sumw = pipe.assign(xyz, 0, None, 2600.)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    reps = 5 if Np <= 10 ** 8 else 2
    for _ in range(reps):
        mesh, sumw = pipe.assign(xyz, 0, None, 2600.)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    n = xyz.shape[1]
    line = 'variant=%s N=%d Np=%d: K1 %.3f ms = %.2f Gpart/s, %.0f GB/s on 16 Np + 8 N^3  (mesh sum/216 %.3f, sumw %.1f)' % (
        var, N, n, ms, n / ms / 1e6, (16. * n + 8. * N ** 3) / ms / 1e6, mesh.double().sum().item() / 216. / 2, sumw.item())
    if os.environ.get('K1_CHECK') and N <= 512:
        f = 'gpurun_out/k1_mesh_%d_%d.pt' % (N, Np)
        if var == '2':
            torch.save(mesh.cpu(), f)
        elif os.path.isfile(f):
            ref = torch.load(f).to(dev)
            line += '  max|d|/max = %.2e' % ((mesh - ref).abs().max().item() / ref.abs().max().item())
            del ref
    print(line, flush=True)
    del xyz, mesh
