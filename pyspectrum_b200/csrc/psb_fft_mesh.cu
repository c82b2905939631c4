// psb_fft_mesh.cu -- mesh (A + iB) -> delta(k) half field: x pass, y pass, z pass fused with fcomb.
// Replaces pyspectrum.py:1060-1080 (_FFT / pyfftw ifftn, unnormalised, sign +) followed by
// estimator.f:605-675 (fcomb_periodic) and the [:N/2+1] slice of pyspectrum.py:959.
#include "psb_fft_lines.cuh"

namespace psb {

int fft_mesh_to_delta(Cx<float>* mesh, Cx<float>* half, int N, const Cx<float>* tw,
                      const Cx<double>* rec, const float* Wk, const double* sumw, int periodic, cudaStream_t st)
{
    FftPlan p;
    if (N % 2 || !make_plan(N, &p)) return PSB_ERR_UNSUPPORTED_N;
    return dispatch_plan(N, [&](auto cfg) -> int {
        const int LPC = cfg_lpc(cfg, p);
        if (LPC < 2) return (int)PSB_ERR_UNSUPPORTED_N;
        const long long nrows = (long long)N * N;
        IoRows<float> io1{ mesh, nrows };
        int rc = launch_any<float, +1>(cfg, p, LPC, dim3((unsigned)((nrows + LPC - 1) / LPC)), tw, io1, st);
        if (rc) return rc;
        IoCols<float, false, false> io2{ mesh, mesh, nullptr, nullptr, nullptr, nullptr, nullptr, 0, N, (long long)N * N, N, (long long)N * N, N, 0, 0 };
        rc = launch_any<float, +1>(cfg, p, LPC, dim3((N + LPC - 1) / LPC, N), tw, io2, st);
        if (rc) return rc;
        IoZFcomb io3{ mesh, half, rec, Wk, sumw, periodic };
        const int HW = LPC / 2;
        return launch_any<float, +1>(cfg, p, LPC, dim3((N / 2 + 1 + HW - 1) / HW, N), tw, io3, st);
    });
}

// plain in-place 3-D c2c transform of [z][y][x] data, unnormalised; dir=+1 is FFTW_BACKWARD.
// Drop-in for estimator.ffting (estimator.f:266-282).
int fft_c2c_3d(Cx<float>* data, int N, int dir, const Cx<float>* tw, cudaStream_t st)
{
    FftPlan p;
    if (!make_plan(N, &p)) return PSB_ERR_UNSUPPORTED_N;
    return dispatch_plan(N, [&](auto cfg) -> int {
        const int LPC = cfg_lpc(cfg, p);
        if (LPC < 1) return (int)PSB_ERR_UNSUPPORTED_N;
        const long long nrows = (long long)N * N;
        IoRows<float> io1{ data, nrows };
        IoCols<float, false, false> io2{ data, data, nullptr, nullptr, nullptr, nullptr, nullptr, 0, N, (long long)N * N, N, (long long)N * N, N, 0, 0 };
        IoCols<float, false, false> io3{ data, data, nullptr, nullptr, nullptr, nullptr, nullptr, 0, N, N, (long long)N * N, N, (long long)N * N, 0, 0 };
        dim3 g1((unsigned)((nrows + LPC - 1) / LPC)), g2((N + LPC - 1) / LPC, N);
        int rc;
        if (dir > 0) {
            if ((rc = launch_any<float, +1>(cfg, p, LPC, g1, tw, io1, st))) return rc;
            if ((rc = launch_any<float, +1>(cfg, p, LPC, g2, tw, io2, st))) return rc;
            return launch_any<float, +1>(cfg, p, LPC, g2, tw, io3, st);
        }
        if ((rc = launch_any<float, -1>(cfg, p, LPC, g1, tw, io1, st))) return rc;
        if ((rc = launch_any<float, -1>(cfg, p, LPC, g2, tw, io2, st))) return rc;
        return launch_any<float, -1>(cfg, p, LPC, g2, tw, io3, st);
    });
}

// fcomb on an already transformed full grid (drop-in for estimator.fcomb_periodic / fcomb_survey called
// on its own): full is Fortran (N,N,N) complex64 = [z][y][x]; writes the half field.
__global__ void k_fcomb(const Cx<float>* __restrict__ full, Cx<float>* __restrict__ half, int N,
                        const Cx<double>* __restrict__ rec, const float* __restrict__ Wk, const double* sumw, int periodic)
{
    const int h = N / 2;
    const long long n = (long long)(h + 1) * N * N;
    const float cf = periodic ? 1.f / (864.f * (float)(*sumw)) : 1.f / 864.f;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(e % (h + 1));
        const long long r = e / (h + 1);
        const int iy = (int)(r % N), iz = (int)(r / N);
        const Cx<float> Fk = full[((long long)iz * N + iy) * N + ix];
        const Cx<float> Fm = full[((long long)kneg(iz, N) * N + kneg(iy, N)) * N + kneg(ix, N)];
        half[e] = fcomb_value(N, ix, iy, iz, Fk, Fm, rec, Wk, cf);
    }
}

int fcomb_standalone(const Cx<float>* full, Cx<float>* half, int N, const Cx<double>* rec, const float* Wk,
                     const double* sumw, int periodic, cudaStream_t st)
{
    if (N < 2 || N % 2) return PSB_ERR_ARG;
    k_fcomb<<<148 * 8, 256, 0, st>>>(full, half, N, rec, Wk, sumw, periodic);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

}  // namespace psb
