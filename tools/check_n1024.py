"""GPU check of the Ngrid=1024 code paths (FFT plan 8.8.4.4, pruned shell transforms) against torch.fft (cuFFT is a
yardstick here, never on the product path)."""
import ctypes, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyspectrum_b200 import pyspectrum as P, _lib
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device('cuda', 0)
pipe = P.PeriodicPipeline.get(N)
L = _lib.lib()
g = torch.Generator(device=dev); g.manual_seed(1)
x = torch.randn((N, N, N, 2), generator=g, device=dev, dtype=torch.float32)
y = x.clone()
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
rc = L.psb_fft_c2c_3d(ctypes.c_void_p(y.data_ptr()), N, 1, ctypes.c_void_p(pipe.tw32.data_ptr()), st)
torch.cuda.synchronize()
t0 = time.perf_counter(); L.psb_fft_c2c_3d(ctypes.c_void_p(y.data_ptr()), N, -1, ctypes.c_void_p(pipe.tw32.data_ptr()), st); torch.cuda.synchronize(); t_fft = time.perf_counter() - t0
# forward(backward(x)) = N^3 x
err_rt = ((torch.view_as_complex(y) / N ** 3 - torch.view_as_complex(x)).abs().max() / torch.view_as_complex(x).abs().max()).item()
y = x.clone()
L.psb_fft_c2c_3d(ctypes.c_void_p(y.data_ptr()), N, 1, ctypes.c_void_p(pipe.tw32.data_ptr()), st)
ref = torch.fft.ifftn(torch.view_as_complex(x)) * N ** 3
err = ((torch.view_as_complex(y) - ref).abs().max() / ref.abs().max()).item()
print('N=%d c2c rc=%d  max err vs cuFFT %.2e  round trip %.2e  one 3-D transform %.2f ms (%.0f GB/s of 48 N^3)' % (N, rc, err, err_rt, t_fft * 1e3, 48.0 * N ** 3 / t_fft / 1e9))
del x, y, ref
torch.cuda.empty_cache()
# shell fields vs torch
rng = np.random.default_rng(0)
xyz = rng.uniform(0, 1000., (3, 2 * 10 ** 6))
half, sumw = pipe.fft_periodic(xyz, None, 1000.)
hc = torch.view_as_complex(half)                       # [kz,ky,kx<=h]
h = N // 2
full = torch.zeros((N, N, N), dtype=torch.complex64, device=dev)   # [kz,ky,kx]
full[:, :, :h + 1] = hc
idx = (-torch.arange(N, device=dev)) % N
full[:, :, h + 1:] = torch.conj(hc[idx][:, idx][:, :, 1:h].flip(2))
k = torch.arange(N, device=dev); k = torch.where(k <= h, k, k - N)
m = (k[:, None, None] ** 2 + k[None, :, None] ** 2 + k[None, None, :] ** 2)
step, Nmax = 3, 40
irk = pipe.irk_table(step).long()[m]
del m
for pairs in ([0], [19]):
    fields, sumsq = pipe.shell_fields(half, step, 1, Nmax, pairs=pairs)
    for e in (0, 1):
        s = 1 + 2 * pairs[0] + e
        ref = torch.fft.fftn(torch.where(irk == s, full, torch.zeros_like(full))).real      # [z,y,x]
        got = fields[e].view(N, N, N)
        print('shell %2d: max |diff| / rms = %.2e   sumsq rel %.2e' % (s, ((got - ref).abs().max() / ref.pow(2).mean().sqrt()).item(),
              abs(sumsq[e].item() / ref.double().pow(2).sum().item() - 1)))
        del ref
