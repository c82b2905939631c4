"""K1 in slab mode at the shape one rank of the C5 / 8-GPU run sees: 1.29e8 routed particles (float4) onto planes [128, 256) of a
1024^3 grid (profiling driver: CUDA-event time + mesh mass; wrap in ncu for the launch list)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pyspectrum_b200 import pyspectrum as pySpec, multigpu as M
N, L, G = 1024, 4000., 8
n = int(os.environ.get('K1_NP', 129000000))
dev = torch.device('cuda', 0)
g = torch.Generator(device=dev); g.manual_seed(7)
nz = N // G
cell = L / N
xyzw = torch.empty((n, 4), dtype=torch.float32, device=dev)
xyzw[:, 0] = torch.rand(n, generator=g, device=dev) * L
xyzw[:, 1] = torch.rand(n, generator=g, device=dev) * L
xyzw[:, 2] = (nz - 3 + torch.rand(n, generator=g, device=dev) * (nz + 4)) * cell      # cells that reach planes [nz, 2 nz)
xyzw[:, 3] = 1.0
pipe = pySpec.PeriodicPipeline.get(N)
for it in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    mesh = M.assign_slab(pipe, xyzw, nz, nz, L)
    e1.record(); torch.cuda.synchronize()
    print('slab K1: %.3f ms = %.2f Gpart/s (mesh sum/216/2 = %.1f of %d)' % (e0.elapsed_time(e1), n / e0.elapsed_time(e1) / 1e6,
                                                                            mesh.double().sum().item() / 432., n), flush=True)
    del mesh
