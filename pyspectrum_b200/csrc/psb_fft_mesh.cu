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
    k_fcomb<<<sm_count() * 8, 256, 0, st>>>(full, half, N, rec, Wk, sumw, periodic);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

// ---------------------------------------------------------------------------------------------------------------------
// Slab-decomposed mesh -> delta(k) for the sharded path: a rank owns nz = N/G z-planes of the reduced mesh.
//   1. fft_slab_xy      x and y passes of d = A + iB on the z-slab (the same kernels as the single-GPU transform)
//   2. slab_split_ab    per z-plane P = A^xy, Q = B^xy on kx = 0..N/2 (the partner (-kx,-ky) is in the same plane: A and B are
//                       real, so A^xy(k) = (D(k) + conj D(-k))/2, B^xy(k) = (D(k) - conj D(-k))/(2i)); rows padded to an even length hp
//   3. (host) all-to-all: z-slabs [nz][N][hp] -> y-slabs [N][ny][hp]
//   4. fft_slab_z       z pass of P and Q on the y-slab
//   5. slab_fcomb       F(k) = P + iQ and F(-k) = conj P + i conj Q rebuilt locally -> the closed form of fcomb (fcomb_value,
//                       including the Fortran's last-write-wins values on the self-conjugate planes) -> half field rows of the slab
// No conjugate-partner traffic between ranks; one all-to-all of 8 N^3 (1 + 2/N) bytes in total.
// ---------------------------------------------------------------------------------------------------------------------
int fft_slab_xy(Cx<float>* data, int N, int nz, int dir, const Cx<float>* tw, cudaStream_t st)
{
    FftPlan p;
    if (nz < 1 || !make_plan(N, &p)) return PSB_ERR_UNSUPPORTED_N;
    return dispatch_plan(N, [&](auto cfg) -> int {
        const int LPC = cfg_lpc(cfg, p);
        if (LPC < 1) return (int)PSB_ERR_UNSUPPORTED_N;
        const long long nrows = (long long)nz * N;
        IoRows<float> io1{ data, nrows };
        IoCols<float, false, false> io2{ data, data, nullptr, nullptr, nullptr, nullptr, nullptr, 0, N, (long long)N * N, N, (long long)N * N, N, 0, 0 };
        dim3 g1((unsigned)((nrows + LPC - 1) / LPC)), g2((N + LPC - 1) / LPC, nz);
        int rc;
        if (dir > 0) {
            if ((rc = launch_any<float, +1>(cfg, p, LPC, g1, tw, io1, st))) return rc;
            return launch_any<float, +1>(cfg, p, LPC, g2, tw, io2, st);
        }
        if ((rc = launch_any<float, -1>(cfg, p, LPC, g1, tw, io1, st))) return rc;
        return launch_any<float, -1>(cfg, p, LPC, g2, tw, io2, st);
    });
}

// z pass (in place) of an array [N z][ny][nx] (nx even): lines = x, batch = y row, element stride ny*nx
int fft_slab_z(Cx<float>* data, int N, int ny, int nx, int dir, const Cx<float>* tw, cudaStream_t st)
{
    FftPlan p;
    if (ny < 1 || nx < 2 || (nx & 1) || !make_plan(N, &p)) return PSB_ERR_UNSUPPORTED_N;
    return dispatch_plan(N, [&](auto cfg) -> int {
        const int LPC = cfg_lpc(cfg, p);
        if (LPC < 1) return (int)PSB_ERR_UNSUPPORTED_N;
        const long long zs = (long long)ny * nx;
        IoCols<float, false, false> io{ data, data, nullptr, nullptr, nullptr, nullptr, nullptr, 0, nx, nx, zs, nx, zs, 0, 0 };
        dim3 g((nx + LPC - 1) / LPC, ny);
        if (dir > 0) return launch_any<float, +1>(cfg, p, LPC, g, tw, io, st);
        return launch_any<float, -1>(cfg, p, LPC, g, tw, io, st);
    });
}

// route != null (multi-GPU): P and Q leave as the z-slab -> ky-slab exchange of the slab FFT itself.  Row ky belongs to rank ky / ny;
// route[q] / route[nranks + q] = address of rank q's P / Q array [N][ny][hp] (peer pointer over NVLink), plane z of this slab is plane
// zbase + z there.  A (z, ky) row is hp contiguous complex numbers (4 KB at 1024^3): full-width NVLink runs.
__global__ void k_slab_split_ab(const Cx<float>* __restrict__ d, Cx<float>* __restrict__ P, Cx<float>* __restrict__ Q, int N, int nz, int hp,
                                const long long* __restrict__ route, int zbase, int ny, int nranks)
{
    const int h = N / 2;
    const long long n = (long long)nz * N * hp;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int kx = (int)(e % hp);
        const long long r = e / hp;
        const int ky = (int)(r % N);
        const long long z = r / N;
        Cx<float> p = mk<float>(0.f, 0.f), q = p;
        if (kx <= h) {
            const Cx<float> Dk = d[(z * N + ky) * N + kx];
            const Cx<float> Dm = d[(z * N + kneg(ky, N)) * N + kneg(kx, N)];
            p = mk<float>(0.5f * (Dk.x + Dm.x), 0.5f * (Dk.y - Dm.y));              // (Dk + conj Dm) / 2
            q = mk<float>(0.5f * (Dk.y + Dm.y), -0.5f * (Dk.x - Dm.x));             // (Dk - conj Dm) / (2i)
        }
        if (route) {
            const int q_ = ky / ny;
            const long long o = (((zbase + z) * ny + (ky - q_ * ny)) * hp + kx) * (long long)sizeof(Cx<float>);
            *reinterpret_cast<Cx<float>*>(route[q_] + o) = p;
            *reinterpret_cast<Cx<float>*>(route[nranks + q_] + o) = q;
        } else {
            P[e] = p;
            Q[e] = q;
        }
    }
}

int slab_split_ab(const Cx<float>* d, Cx<float>* P, Cx<float>* Q, int N, int nz, int hp, cudaStream_t st, const long long* route, int zbase,
                  int nranks)
{
    if (N < 2 || (N & 1) || nz < 1 || hp < N / 2 + 1) return PSB_ERR_ARG;
    if (route ? (nranks < 1 || N % nranks || zbase < 0 || zbase + nz > N) : (!P || !Q)) return PSB_ERR_ARG;
    k_slab_split_ab<<<sm_count() * 8, 256, 0, st>>>(d, P, Q, N, nz, hp, route, zbase, route ? N / nranks : N, nranks);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

__global__ void k_slab_fcomb(const Cx<float>* __restrict__ P, const Cx<float>* __restrict__ Q, Cx<float>* __restrict__ half, int N, int ky0,
                             int ny, int hp, const Cx<double>* __restrict__ rec, const float* __restrict__ Wk, const double* sumw, int periodic)
{
    const int h = N / 2;
    const long long n = (long long)N * ny * (h + 1);
    const float cf = periodic ? 1.f / (864.f * (float)(*sumw)) : 1.f / 864.f;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int ix = (int)(e % (h + 1));
        const long long r = e / (h + 1);
        const int yl = (int)(r % ny), iz = (int)(r / ny);
        const long long s = ((long long)iz * ny + yl) * hp + ix;
        const Cx<float> p = P[s], q = Q[s];
        const Cx<float> Fk = mk<float>(p.x - q.y, p.y + q.x);                      // A^ + i B^
        const Cx<float> Fm = mk<float>(p.x + q.y, q.x - p.y);                      // conj A^ + i conj B^ = F(-k)
        half[e] = fcomb_value(N, ix, ky0 + yl, iz, Fk, Fm, rec, Wk, cf);
    }
}

int slab_fcomb(const Cx<float>* P, const Cx<float>* Q, Cx<float>* half, int N, int ky0, int ny, int hp, const Cx<double>* rec,
               const float* Wk, const double* sumw, int periodic, cudaStream_t st)
{
    if (N < 2 || (N & 1) || ny < 1 || ky0 < 0 || ky0 + ny > N || hp < N / 2 + 1) return PSB_ERR_ARG;
    k_slab_fcomb<<<sm_count() * 8, 256, 0, st>>>(P, Q, half, N, ky0, ny, hp, rec, Wk, sumw, periodic);
    return cudaGetLastError() == cudaSuccess ? PSB_OK : PSB_ERR_CUDA;
}

}  // namespace psb
