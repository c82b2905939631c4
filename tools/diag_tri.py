"""GPU diagnostic: accuracy of the K6 engines against a float64 torch evaluation on the same stored fields."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pyspectrum_b200 import pyspectrum as pySpec

def cat(seed, Np, L):
    rng = np.random.default_rng(seed)
    npar = max(Np // 40, 1)
    par = rng.uniform(0, L, (3, npar))
    kids = par[:, rng.integers(0, npar, Np // 2)] + rng.normal(0, 0.03 * L, (3, Np // 2))
    return np.ascontiguousarray(np.concatenate([kids, rng.uniform(0, L, (3, Np - Np // 2))], axis=1) % L)

CASES = [(32, 20000, 1, 1, 12), (64, 100000, 1, 3, 30), (128, 2000000, 3, 3, 20)]
if len(sys.argv) > 1 and sys.argv[1] == 'big':
    CASES = [(192, 400000, 1, 3, 72), (192, 4000000, 1, 3, 72)]
for (N, Np, step, Ncut, Nmax) in CASES:
    L = 300.
    pipe = pySpec.PeriodicPipeline.get(N)
    half, _ = pipe.fft_periodic(cat(N + Nmax, Np, L), None, L)
    s0 = Ncut // step
    fields, sumsq, scales, maxabs = pipe.shell_fields(half, step, s0, Nmax, scaled=True)
    tri = pySpec.triangle_list(Nmax, Ncut, step)
    f64 = pipe.unpack_fields(fields).double()                # the stored fields are fp16 hi/lo packed
    ti = torch.from_numpy(tri.astype(np.int64) - s0).to(fields.device)
    ref = torch.empty(len(tri), dtype=torch.float64, device=fields.device); nrm = torch.empty_like(ref)
    B = max(1, (1 << 26) // fields.shape[1])
    for a in range(0, len(tri), B):
        t = ti[a:a + B]
        prod = f64[t[:, 0]] * f64[t[:, 1]] * f64[t[:, 2]]
        ref[a:a + B] = prod.sum(dim=1); nrm[a:a + B] = prod.pow(2).sum(dim=1).sqrt()
    print('N=%d ntri=%d  median |ref|/nrm = %.2f max = %.1f  maxabs=%.1f' % (N, len(tri), (ref.abs() / nrm).median().item(), (ref.abs() / nrm).max().item(),
          maxabs.view(torch.float32).max().item()))
    for engine in ('fma', 'tc'):
        got = pipe.triangle_sums(fields, Nmax, Ncut, step, engine=engine)
        err = (got - ref).abs()
        print('  %-4s max err/nrm %.2e   max err/|ref| %.2e   max err/(nrm+|ref|) %.2e   median err/(nrm+|ref|) %.2e' % (
            engine, (err / nrm).max().item(), (err / ref.abs()).max().item(), (err / (nrm + ref.abs())).max().item(),
            (err / (nrm + ref.abs())).median().item()))
        bias = ((got - ref) / ref)[ref.abs() > 5 * nrm]
        if bias.numel():
            print('       signed rel err on signal-dominated triangles: mean %.2e  min %.2e  max %.2e (n=%d)' % (bias.mean().item(), bias.min().item(), bias.max().item(), bias.numel()))
