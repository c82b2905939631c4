// Host emulation harness (TEST ONLY): compiles the __host__ __device__ cores of the CUDA
// kernels with g++ so their index math / butterflies / combine rules are checked on CPU.
#include "../../pyspectrum_b200/csrc/psb_fft_core.cuh"
#include <vector>
#include <cmath>
using namespace psb;

template <typename T, int DIR>
static int run_fft(T* data, int N, int estride)
{
    FftPlan p;
    if (!make_plan(N, &p)) return -2;
    std::vector<Cx<T>> tw(N);
    for (int i = 0; i < N; ++i) { tw[i].x = (T)std::cos(2.0 * M_PI * i / N); tw[i].y = (T)std::sin(2.0 * M_PI * i / N); }
    Cx<T>* s = reinterpret_cast<Cx<T>*>(data);
    int Ns = 1;
    for (int st = 0; st < p.nstages; ++st) {
        const int R = p.radix[st];
        const int M = N / R;
        std::vector<Cx<T>> regs((size_t)M * R);
        // phase 1: every butterfly reads; phase 2: every butterfly writes (as the kernel does around a barrier)
        for (int j = 0; j < M; ++j) {
            Cx<T>* v = &regs[(size_t)j * R];
            switch (R) {
                case 2: stage_read<2, T>(s, estride, N, j, v); break;
                case 3: stage_read<3, T>(s, estride, N, j, v); break;
                case 4: stage_read<4, T>(s, estride, N, j, v); break;
                case 5: stage_read<5, T>(s, estride, N, j, v); break;
                case 8: stage_read<8, T>(s, estride, N, j, v); break;
                case 9: stage_read<9, T>(s, estride, N, j, v); break;
                default: return -9;
            }
        }
        for (int j = 0; j < M; ++j) {
            Cx<T>* v = &regs[(size_t)j * R];
            switch (R) {
                case 2: stage_write<2, DIR, T>(s, estride, N, Ns, j, tw.data(), v); break;
                case 3: stage_write<3, DIR, T>(s, estride, N, Ns, j, tw.data(), v); break;
                case 4: stage_write<4, DIR, T>(s, estride, N, Ns, j, tw.data(), v); break;
                case 5: stage_write<5, DIR, T>(s, estride, N, Ns, j, tw.data(), v); break;
                case 8: stage_write<8, DIR, T>(s, estride, N, Ns, j, tw.data(), v); break;
                case 9: stage_write<9, DIR, T>(s, estride, N, Ns, j, tw.data(), v); break;
            }
        }
        Ns *= R;
    }
    return 0;
}

extern "C" int emu_fft_f32(float* data, int N, int dir, int estride) { return dir > 0 ? run_fft<float, 1>(data, N, estride) : run_fft<float, -1>(data, N, estride); }
extern "C" int emu_fft_f64(double* data, int N, int dir, int estride) { return dir > 0 ? run_fft<double, 1>(data, N, estride) : run_fft<double, -1>(data, N, estride); }
extern "C" int emu_plan(int N, int* radices) { FftPlan p; if (!make_plan(N, &p)) return -1; for (int i = 0; i < p.nstages; ++i) radices[i] = p.radix[i]; return p.nstages; }


// ---- composite in-register DFTs (Dft<16|18|20|32>) and the two-stage plans built from them (360 = 20 x 18, 256 = 16 x 16)
template <int R, int DIR> static void run_dft(double* data) { Dft<R, DIR, double>::run(reinterpret_cast<Cx<double>*>(data)); }
extern "C" int emu_dft(double* data, int R, int dir)
{
    switch (R) {
        case 2: dir > 0 ? run_dft<2, 1>(data) : run_dft<2, -1>(data); return 0;
        case 3: dir > 0 ? run_dft<3, 1>(data) : run_dft<3, -1>(data); return 0;
        case 4: dir > 0 ? run_dft<4, 1>(data) : run_dft<4, -1>(data); return 0;
        case 5: dir > 0 ? run_dft<5, 1>(data) : run_dft<5, -1>(data); return 0;
        case 8: dir > 0 ? run_dft<8, 1>(data) : run_dft<8, -1>(data); return 0;
        case 9: dir > 0 ? run_dft<9, 1>(data) : run_dft<9, -1>(data); return 0;
        case 16: dir > 0 ? run_dft<16, 1>(data) : run_dft<16, -1>(data); return 0;
        case 18: dir > 0 ? run_dft<18, 1>(data) : run_dft<18, -1>(data); return 0;
        case 20: dir > 0 ? run_dft<20, 1>(data) : run_dft<20, -1>(data); return 0;
        case 32: dir > 0 ? run_dft<32, 1>(data) : run_dft<32, -1>(data); return 0;
        default: return -1;
    }
}

// the fused-edge two-stage line FFT exactly as fft_lines_fused_kernel runs it: stage 1 from "global" (in), Ns = 1, no twiddles;
// stage 2 with twiddles tw[q*j], outputs at j + q*M2
template <typename T, int R1, int R2, int DIR> static void two_stage(const T* in_, T* out_)
{
    constexpr int N = R1 * R2, M1 = N / R1, M2 = N / R2;
    const Cx<T>* in = reinterpret_cast<const Cx<T>*>(in_);
    Cx<T>* out = reinterpret_cast<Cx<T>*>(out_);
    std::vector<Cx<T>> s(N), tw(N);
    for (int i = 0; i < N; ++i) { tw[i].x = (T)std::cos(2.0 * M_PI * i / N); tw[i].y = (T)std::sin(2.0 * M_PI * i / N); }
    for (int j = 0; j < M1; ++j) {
        Cx<T> v[R1];
        for (int q = 0; q < R1; ++q) v[q] = in[j + q * M1];
        Dft<R1, DIR, T>::run(v);
        for (int q = 0; q < R1; ++q) s[j * R1 + q] = v[q];
    }
    for (int j = 0; j < M2; ++j) {
        Cx<T> v[R2];
        stage_read<R2, T>(s.data(), 1, N, j, v);
        for (int q = 1; q < R2; ++q) { Cx<T> w = tw[q * j]; if (DIR < 0) w.y = -w.y; v[q] = v[q] * w; }
        Dft<R2, DIR, T>::run(v);
        for (int q = 0; q < R2; ++q) out[j + q * M2] = v[q];
    }
}
extern "C" int emu_two_stage_f32(const float* in, float* out, int N, int dir)
{
    if (N == 360) { dir > 0 ? two_stage<float, 20, 18, 1>(in, out) : two_stage<float, 20, 18, -1>(in, out); return 0; }
    if (N == 256) { dir > 0 ? two_stage<float, 16, 16, 1>(in, out) : two_stage<float, 16, 16, -1>(in, out); return 0; }
    return -1;
}
extern "C" int emu_two_stage_f64(const double* in, double* out, int N, int dir)
{
    if (N == 360) { dir > 0 ? two_stage<double, 20, 18, 1>(in, out) : two_stage<double, 20, 18, -1>(in, out); return 0; }
    if (N == 256) { dir > 0 ? two_stage<double, 16, 16, 1>(in, out) : two_stage<double, 16, 16, -1>(in, out); return 0; }
    return -1;
}

#include "../../pyspectrum_b200/csrc/psb_fcomb_core.cuh"
// full: complex64 (N,N,N) Fortran order [ix + N*(iy + N*iz)] BEFORE fcomb; half: output (N/2+1,N,N) Fortran order
extern "C" void emu_fcomb(const float* full_, float* half_, int N, float sumw, int periodic)
{
    const Cx<float>* F = reinterpret_cast<const Cx<float>*>(full_);
    Cx<float>* H = reinterpret_cast<Cx<float>*>(half_);
    const int h = N / 2;
    std::vector<Cx<double>> rec(h + 1);
    std::vector<float> Wk(h + 1);
    fcomb_build_tables(N, rec.data(), Wk.data());
    const float cf = periodic ? 1.f / (864.f * sumw) : 1.f / 864.f;
    for (int iz = 0; iz < N; ++iz) for (int iy = 0; iy < N; ++iy) for (int ix = 0; ix <= h; ++ix) {
        Cx<float> Fk = F[ix + (size_t)N * (iy + (size_t)N * iz)];
        Cx<float> Fm = F[kneg(ix, N) + (size_t)N * (kneg(iy, N) + (size_t)N * kneg(iz, N))];
        H[ix + (size_t)(h + 1) * (iy + (size_t)N * iz)] = fcomb_value(N, ix, iy, iz, Fk, Fm, rec.data(), Wk.data(), cf);
    }
}

// Slab-decomposed fcomb (psb_fft_mesh.cu: k_slab_split_ab / k_slab_fcomb), element for element the arithmetic of the two kernels.
// d: [nz][N][N] complex64 after the x and y passes -> p, q: [nz][N][hp]
extern "C" void emu_slab_split(const float* d_, float* p_, float* q_, int N, int nz, int hp)
{
    const Cx<float>* d = reinterpret_cast<const Cx<float>*>(d_);
    Cx<float>* P = reinterpret_cast<Cx<float>*>(p_);
    Cx<float>* Q = reinterpret_cast<Cx<float>*>(q_);
    const int h = N / 2;
    for (long long z = 0; z < nz; ++z) for (int ky = 0; ky < N; ++ky) for (int kx = 0; kx < hp; ++kx) {
        Cx<float> p = mk<float>(0.f, 0.f), q = p;
        if (kx <= h) {
            const Cx<float> Dk = d[(z * N + ky) * N + kx];
            const Cx<float> Dm = d[(z * N + kneg(ky, N)) * N + kneg(kx, N)];
            p = mk<float>(0.5f * (Dk.x + Dm.x), 0.5f * (Dk.y - Dm.y));
            q = mk<float>(0.5f * (Dk.y + Dm.y), -0.5f * (Dk.x - Dm.x));
        }
        P[(z * N + ky) * hp + kx] = p;
        Q[(z * N + ky) * hp + kx] = q;
    }
}
// p, q: [N][ny][hp] after the z pass -> half: [N][ny][N/2+1] = rows ky0.. of the half field
extern "C" void emu_slab_fcomb(const float* p_, const float* q_, float* half_, int N, int ky0, int ny, int hp, float sumw, int periodic)
{
    const Cx<float>* P = reinterpret_cast<const Cx<float>*>(p_);
    const Cx<float>* Q = reinterpret_cast<const Cx<float>*>(q_);
    Cx<float>* H = reinterpret_cast<Cx<float>*>(half_);
    const int h = N / 2;
    std::vector<Cx<double>> rec(h + 1);
    std::vector<float> Wk(h + 1);
    fcomb_build_tables(N, rec.data(), Wk.data());
    const float cf = periodic ? 1.f / (864.f * sumw) : 1.f / 864.f;
    for (int iz = 0; iz < N; ++iz) for (int yl = 0; yl < ny; ++yl) for (int ix = 0; ix <= h; ++ix) {
        const size_t s = ((size_t)iz * ny + yl) * hp + ix;
        const Cx<float> p = P[s], q = Q[s];
        const Cx<float> Fk = mk<float>(p.x - q.y, p.y + q.x);
        const Cx<float> Fm = mk<float>(p.x + q.y, q.x - p.y);
        H[((size_t)iz * ny + yl) * (h + 1) + ix] = fcomb_value(N, ix, ky0 + yl, iz, Fk, Fm, rec.data(), Wk.data(), cf);
    }
}
