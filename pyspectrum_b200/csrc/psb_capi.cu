// psb_capi.cu -- extern "C" surface declared in include/psb200.h.
#include <cuda_runtime.h>
#include <cmath>
#include <cstring>
#include <map>
#include <tuple>
#include <vector>
#include "../../include/psb200.h"
#include "psb_kernels.h"
#include "psb_fcomb_core.cuh"

using namespace psb;

namespace psb {
// one process drives one GPU (torchrun, one rank per device): the count of the first device seen is kept
int sm_count()
{
    static const int n = [] {
        int dev = 0, v = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1) v = 148;
        return v;
    }();
    return n;
}
}  // namespace psb

namespace {

struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { return cudaMalloc(&p, n ? n : 1) == cudaSuccess ? PSB_OK : PSB_ERR_CUDA; }
    template <typename T> T* as() { return static_cast<T*>(p); }
};

#define PSB_TRY(x) do { int rc_ = (x); if (rc_ != PSB_OK) return rc_; } while (0)
#define PSB_CUDA(x) do { if ((x) != cudaSuccess) return PSB_ERR_CUDA; } while (0)

inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

// tiles of 4x4x4 in (i,j,l) slot space; tri = [ntri][3] shell indices, slot = shell - s0
int build_tiles(const int32_t* tri, int ntri, int s0, std::vector<int32_t>& out)
{
    std::map<std::tuple<int, int, int>, int> idx;
    out.clear();
    for (int t = 0; t < ntri; ++t) {
        const int a = tri[3 * t] - s0, b = tri[3 * t + 1] - s0, c = tri[3 * t + 2] - s0;
        if (a < 0 || b < 0 || c < 0) return PSB_ERR_ARG;
        auto key = std::make_tuple(a / 4, b / 4, c / 4);
        auto it = idx.find(key);
        int ti;
        if (it == idx.end()) {
            ti = (int)(out.size() / 68);
            idx[key] = ti;
            out.resize(out.size() + 68, -1);
            out[(size_t)ti * 68 + 0] = (a / 4) * 4; out[(size_t)ti * 68 + 1] = (b / 4) * 4;
            out[(size_t)ti * 68 + 2] = (c / 4) * 4; out[(size_t)ti * 68 + 3] = 0;
        } else ti = it->second;
        out[(size_t)ti * 68 + 4 + ((a % 4) * 4 + (b % 4)) * 4 + (c % 4)] = t;
    }
    return PSB_OK;
}

}  // namespace

extern "C" {

int psb_version(void) { return 100; }

const char* psb_error_string(int code)
{
    switch (code) {
        case PSB_OK: return "ok";
        case PSB_ERR_ARG: return "invalid argument";
        case PSB_ERR_UNSUPPORTED_N: return "unsupported grid size (need even N = 2^a 3^b 5^c within shared-memory limits)";
        case PSB_ERR_CUDA: return "CUDA error (no device / launch failure / out of memory)";
        case PSB_ERR_WORKSPACE: return "workspace too small";
        default: return "unknown error";
    }
}

// ---------------------------------------------------------------- host table builders
int psb_twiddles_f32(int N, float* tw)
{
    if (N < 1 || !tw) return PSB_ERR_ARG;
    for (int k = 0; k < N; ++k) { const double a = 2.0 * M_PI * (double)k / (double)N; tw[2 * k] = (float)cos(a); tw[2 * k + 1] = (float)sin(a); }
    return PSB_OK;
}
int psb_twiddles_f64(int N, double* tw)
{
    if (N < 1 || !tw) return PSB_ERR_ARG;
    for (int k = 0; k < N; ++k) { const double a = 2.0 * M_PI * (double)k / (double)N; tw[2 * k] = cos(a); tw[2 * k + 1] = sin(a); }
    return PSB_OK;
}
int psb_fcomb_tables(int N, double* rec, float* wk)
{
    if (N < 2 || N % 2 || !rec || !wk) return PSB_ERR_ARG;
    fcomb_build_tables(N, reinterpret_cast<Cx<double>*>(rec), wk);
    return PSB_OK;
}
int psb_rsd_trig(int irsd, float* t)
{
    // estimator.f:164, 172-181: pi is the single-precision parameter 3.141592654
    const float pi = 3.141592654f;
    float th = 0.f, ph = 0.f;
    if (irsd == 0) { th = 0.5f * pi; ph = 0.f; }
    else if (irsd == 1) { th = 0.5f * pi; ph = 0.5f * pi; }
    else if (irsd == 2) { th = 0.f; ph = 0.f; }
    else return PSB_ERR_ARG;
    t[0] = cosf(th); t[1] = sinf(th); t[2] = cosf(ph); t[3] = sinf(ph);
    return PSB_OK;
}
int psb_rsd_bin_table(int N, int nbin, int mmax, uint16_t* out)
{
    if (N < 2 || nbin < 1 || mmax < 0 || !out) return PSB_ERR_ARG;
    const float half = (float)(N / 2);
    for (int m = 0; m <= mmax; ++m) {
        const float rk = sqrtf((float)m);                       // exact-input single sqrt, f:206
        const long b = lroundf((float)nbin * rk / half);        // nint(), f:207
        out[m] = (uint16_t)(b > 65535 ? 65535 : b);
    }
    return PSB_OK;
}
int psb_irk_table_f32(float step, int mmax, uint16_t* out)
{
    if (!(step > 0.f) || mmax < 0 || !out) return PSB_ERR_ARG;
    for (int m = 0; m <= mmax; ++m) {
        const float dist = sqrtf((float)m);
        const int b = (int)(dist / step + 0.5f);                // estimator.f:35-36
        out[m] = (uint16_t)(b > 65535 ? 65535 : b);
    }
    return PSB_OK;
}

// ---------------------------------------------------------------- device pipeline
size_t psb_assign_workspace_bytes(int64_t np, int ngrid) { return assign_workspace_bytes(np, ngrid); }

int psb_assign_pcs_interlaced(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64,
                              int64_t np, int ngrid, double lbox_clip, float kf_ks, float offset,
                              float* mesh, int zero_mesh, void* ws, size_t ws_bytes, double* sumw, void* stream)
{
    if ((!pos && np > 0) || !mesh || !ws || !sumw) return PSB_ERR_ARG;
    AssignIn in;
    in.pos = pos; in.pos_f64 = pos_f64; in.pos_aos = pos_aos; in.w = w; in.w_f64 = w_f64; in.Np = np; in.N = ngrid;
    in.do_clip = lbox_clip > 0.0; in.clip_hi = lbox_clip * (1. - 1e-6); in.kf_ks = kf_ks; in.offset = offset;
    in.zbase = 0; in.nzs = ngrid; in.tiles = 0;
    return assign_pcs_interlaced(in, mesh, zero_mesh, ws, ws_bytes, sumw, S(stream));
}

static AssignIn route_in(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64, int64_t np, int ngrid, double lbox_clip,
                         float kf_ks, float offset)
{
    AssignIn in;
    in.pos = pos; in.pos_f64 = pos_f64; in.pos_aos = pos_aos; in.w = w; in.w_f64 = w_f64; in.Np = np; in.N = ngrid;
    in.do_clip = lbox_clip > 0.0; in.clip_hi = lbox_clip * (1. - 1e-6); in.kf_ks = kf_ks; in.offset = offset;
    in.zbase = 0; in.nzs = ngrid; in.tiles = 0;
    return in;
}

int psb_slab_route_count(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64, int64_t np, int ngrid, double lbox_clip,
                         float kf_ks, float offset, int nz_per_rank, int nranks, uint64_t* counts, double* sumw, void* stream)
{
    if ((!pos && np > 0) || !counts || !sumw) return PSB_ERR_ARG;
    return slab_route_count(route_in(pos, pos_f64, pos_aos, w, w_f64, np, ngrid, lbox_clip, kf_ks, offset), nz_per_rank, nranks,
                            reinterpret_cast<unsigned long long*>(counts), sumw, S(stream));
}

int psb_slab_route_scatter(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64, int64_t np, int ngrid, double lbox_clip,
                           float kf_ks, float offset, int nz_per_rank, int nranks, const uint64_t* base, uint64_t* cursor, float* send_xyzw,
                           void* stream)
{
    if ((!pos && np > 0) || !base || !cursor || (!send_xyzw && np > 0)) return PSB_ERR_ARG;
    return slab_route_scatter(route_in(pos, pos_f64, pos_aos, w, w_f64, np, ngrid, lbox_clip, kf_ks, offset), nz_per_rank, nranks,
                              reinterpret_cast<const unsigned long long*>(base), reinterpret_cast<unsigned long long*>(cursor),
                              reinterpret_cast<float4*>(send_xyzw), S(stream));
}

int psb_slab_route_scatter_peer(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64, int64_t np, int ngrid, double lbox_clip,
                                float kf_ks, float offset, int nz_per_rank, int nranks, const int64_t* dest_addr, uint64_t* cursor, void* stream)
{
    if ((!pos && np > 0) || !dest_addr || !cursor) return PSB_ERR_ARG;
    return slab_route_scatter_peer(route_in(pos, pos_f64, pos_aos, w, w_f64, np, ngrid, lbox_clip, kf_ks, offset), nz_per_rank, nranks,
                                   reinterpret_cast<const long long*>(dest_addr), reinterpret_cast<unsigned long long*>(cursor), S(stream));
}

int psb_assign_slab(const float* xyzw, int64_t np, int ngrid, float kf_ks, float offset, int zbase, int nzs, float* mesh_slab, int zero_mesh,
                    void* ws, size_t ws_bytes, double* sumw_scratch, void* stream)
{
    if ((!xyzw && np > 0) || !mesh_slab || !ws || !sumw_scratch) return PSB_ERR_ARG;
    AssignIn in = route_in(xyzw, 0, 2, nullptr, 0, np, ngrid, 0.0, kf_ks, offset);
    in.zbase = zbase; in.nzs = nzs;
    return assign_pcs_interlaced(in, mesh_slab, zero_mesh, ws, ws_bytes, sumw_scratch, S(stream));
}

int psb_fft_mesh_to_delta(float* mesh, float* half, int N, const float* tw, const double* rec, const float* wk,
                          const double* sumw, int periodic, void* stream)
{
    if (!mesh || !half || !tw || !rec || !wk || (periodic && !sumw)) return PSB_ERR_ARG;
    return fft_mesh_to_delta(reinterpret_cast<Cx<float>*>(mesh), reinterpret_cast<Cx<float>*>(half), N,
                             reinterpret_cast<const Cx<float>*>(tw), reinterpret_cast<const Cx<double>*>(rec), wk, sumw, periodic, S(stream));
}

int psb_fft_c2c_3d(float* data, int N, int dir, const float* tw, void* stream)
{
    if (!data || !tw || (dir != 1 && dir != -1)) return PSB_ERR_ARG;
    return fft_c2c_3d(reinterpret_cast<Cx<float>*>(data), N, dir, reinterpret_cast<const Cx<float>*>(tw), S(stream));
}

int psb_fcomb(const float* full, float* half, int N, const double* rec, const float* wk, const double* sumw, int periodic, void* stream)
{
    if (!full || !half || !rec || !wk || (periodic && !sumw)) return PSB_ERR_ARG;
    return fcomb_standalone(reinterpret_cast<const Cx<float>*>(full), reinterpret_cast<Cx<float>*>(half), N,
                            reinterpret_cast<const Cx<double>*>(rec), wk, sumw, periodic, S(stream));
}

int psb_fft_slab_xy(float* data, int N, int nz, int dir, const float* tw, void* stream)
{
    if (!data || !tw || (dir != 1 && dir != -1)) return PSB_ERR_ARG;
    return fft_slab_xy(reinterpret_cast<Cx<float>*>(data), N, nz, dir, reinterpret_cast<const Cx<float>*>(tw), S(stream));
}

int psb_fft_slab_z(float* data, int N, int ny, int nx, int dir, const float* tw, void* stream)
{
    if (!data || !tw || (dir != 1 && dir != -1)) return PSB_ERR_ARG;
    return fft_slab_z(reinterpret_cast<Cx<float>*>(data), N, ny, nx, dir, reinterpret_cast<const Cx<float>*>(tw), S(stream));
}

int psb_slab_split_ab(const float* d, float* p, float* q, int N, int nz, int hp, void* stream)
{
    if (!d || !p || !q) return PSB_ERR_ARG;
    return slab_split_ab(reinterpret_cast<const Cx<float>*>(d), reinterpret_cast<Cx<float>*>(p), reinterpret_cast<Cx<float>*>(q), N, nz, hp,
                         S(stream));
}

int psb_slab_split_ab_routed(const float* d, int N, int nz, int hp, int zbase, int nranks, const int64_t* route, void* stream)
{
    if (!d || !route) return PSB_ERR_ARG;
    return slab_split_ab(reinterpret_cast<const Cx<float>*>(d), nullptr, nullptr, N, nz, hp, S(stream), reinterpret_cast<const long long*>(route),
                         zbase, nranks);
}

int psb_slab_fcomb(const float* p, const float* q, float* half, int N, int ky0, int ny, int hp, const double* rec, const float* wk,
                   const double* sumw, int periodic, void* stream)
{
    if (!p || !q || !half || !rec || !wk || (periodic && !sumw)) return PSB_ERR_ARG;
    return slab_fcomb(reinterpret_cast<const Cx<float>*>(p), reinterpret_cast<const Cx<float>*>(q), reinterpret_cast<Cx<float>*>(half), N, ky0,
                      ny, hp, reinterpret_cast<const Cx<double>*>(rec), wk, sumw, periodic, S(stream));
}

int psb_pk_monopole_slab(const float* half, int N, int ky0, int ny, const uint16_t* bin, int nbin, double kf, double* out, void* stream)
{
    if (!half || !bin || !out) return PSB_ERR_ARG;
    SpectraIn in{};
    in.half = reinterpret_cast<const Cx<float>*>(half); in.N = N; in.bin = bin; in.Nbin = nbin; in.mode = 0; in.kf = kf; in.Nmu = 1;
    in.ky0 = ky0; in.ny = ny;
    return binned_spectra(in, out, S(stream));
}

int psb_pk_monopole(const float* half, int N, const uint16_t* bin, int nbin, double kf, double* out, void* stream)
{
    return psb_pk_monopole_slab(half, N, 0, N, bin, nbin, kf, out, stream);
}

int psb_pk_multipoles_slab(const float* half, int N, int ky0, int ny, const uint16_t* bin, int nbin, int nmu, float kf32, const float* trig4,
                           double* out, void* stream)
{
    if (!half || !bin || !out || !trig4 || nmu < 1) return PSB_ERR_ARG;
    SpectraIn in{};
    in.half = reinterpret_cast<const Cx<float>*>(half); in.N = N; in.bin = bin; in.Nbin = nbin; in.mode = 1; in.kf32 = kf32; in.Nmu = nmu;
    in.costh = trig4[0]; in.sinth = trig4[1]; in.cosph = trig4[2]; in.sinph = trig4[3];
    in.ky0 = ky0; in.ny = ny;
    return binned_spectra(in, out, S(stream));
}

int psb_pk_multipoles(const float* half, int N, const uint16_t* bin, int nbin, int nmu, float kf32, const float* trig4, double* out, void* stream)
{
    return psb_pk_multipoles_slab(half, N, 0, N, bin, nbin, nmu, kf32, trig4, out, stream);
}

int psb_half_extract(const float* half_slab, int N, int ky0, int ny, float* carrier, int Ng, void* stream)
{
    return half_extract(reinterpret_cast<const Cx<float>*>(half_slab), N, ky0, ny, reinterpret_cast<Cx<float>*>(carrier), Ng, S(stream));
}

int psb_pk_kmu_python(const float* full, int N, const uint16_t* bin, int nbin, int nmu, double kf, const double* trig4, double* out, void* stream)
{
    return kmu_python(reinterpret_cast<const Cx<float>*>(full), N, bin, nbin, nmu, kf, trig4, out, S(stream));
}

int psb_shell_mode_counts(int N, const uint16_t* irk, int nshell, uint64_t* nk, void* stream)
{
    if (!irk || !nk) return PSB_ERR_ARG;
    return shell_mode_counts(N, irk, nshell, reinterpret_cast<unsigned long long*>(nk), S(stream));
}

int psb_bk_shell_pair_f32(const float* half, const uint16_t* irk, int N, int Ns, int sa, int sb, int R, float* t1, float* t2,
                          float* fa, float* fb, double* sumsq, const float* scale2, uint32_t* maxabs2, int pack_half, const float* tw, void* stream)
{
    if (!irk || !t1 || !t2 || !fa || !sumsq || !tw || R < 0 || (sb >= 0 && !fb)) return PSB_ERR_ARG;
    return fft_shell_pair<float>(reinterpret_cast<const Cx<float>*>(half), irk, N, Ns, sa, sb, R, reinterpret_cast<Cx<float>*>(t1),
                                 reinterpret_cast<Cx<float>*>(t2), fa, fb, sumsq, scale2, maxabs2, pack_half, reinterpret_cast<const Cx<float>*>(tw), S(stream));
}

int psb_bk_shell_pair_f32_routed(const float* half, const uint16_t* irk, int N, int Ns, int sa, int sb, int R, float* t1, float* t2,
                                 const int64_t* route, int planes_per_rank, int nranks, double* sumsq, const float* scale2,
                                 uint32_t* maxabs2, int pack_half, const float* tw, void* stream)
{
    if (!irk || !t1 || !t2 || !route || !sumsq || !tw || R < 0) return PSB_ERR_ARG;
    return fft_shell_pair<float>(reinterpret_cast<const Cx<float>*>(half), irk, N, Ns, sa, sb, R, reinterpret_cast<Cx<float>*>(t1),
                                 reinterpret_cast<Cx<float>*>(t2), nullptr, nullptr, sumsq, scale2, maxabs2, pack_half,
                                 reinterpret_cast<const Cx<float>*>(tw), S(stream), reinterpret_cast<const long long*>(route),
                                 planes_per_rank, nranks);
}

int psb_bk_shell_pair_f64(const float* half, const uint16_t* irk, int N, int Ns, int sa, int sb, int R, double* t1, double* t2,
                          double* fa, double* fb, double* sumsq, const double* tw, void* stream)
{
    if (!irk || !t1 || !t2 || !fa || !sumsq || !tw || R < 0 || (sb >= 0 && !fb)) return PSB_ERR_ARG;
    return fft_shell_pair<double>(reinterpret_cast<const Cx<float>*>(half), irk, N, Ns, sa, sb, R, reinterpret_cast<Cx<double>*>(t1),
                                  reinterpret_cast<Cx<double>*>(t2), fa, fb, sumsq, nullptr, nullptr, 0, reinterpret_cast<const Cx<double>*>(tw), S(stream));
}

int psb_bk_shell_power(const float* half, int N, const uint16_t* irk, int nshell, double* psum, void* stream)
{
    return shell_power(reinterpret_cast<const Cx<float>*>(half), N, irk, nshell, psum, S(stream));
}

int psb_bk_shell_scales(const double* psum, int nshell, float target_rms, float* scales, void* stream)
{
    return shell_scales(psum, nshell, target_rms, scales, S(stream));
}

size_t psb_bk_triangle_workspace_bytes(int ntiles) { return triangle_workspace_bytes(ntiles); }

int psb_bk_triangle_sums_f32(const float* const* fields, int nfields, int64_t ncell, const int32_t* tiles, int ntiles,
                             double* sums, void* ws, size_t ws_bytes, int packed_half, void* stream)
{
    if (!fields || !tiles || !sums || !ws) return PSB_ERR_ARG;
    return triangle_sums_tiles<float>(fields, nfields, ncell, tiles, ntiles, sums, ws, ws_bytes, packed_half, S(stream));
}
int psb_bk_triangle_sums_f64(const double* const* fields, int nfields, int64_t ncell, const int32_t* tiles, int ntiles,
                             double* sums, void* ws, size_t ws_bytes, void* stream)
{
    if (!fields || !tiles || !sums || !ws) return PSB_ERR_ARG;
    return triangle_sums_tiles<double>(fields, nfields, ncell, tiles, ntiles, sums, ws, ws_bytes, 0, S(stream));
}

size_t psb_bk_triangle_tc_workspace_bytes(int mt, int nt) { return triangle_tc_workspace_bytes(mt, nt); }

int psb_bk_triangle_sums_tc(const float* const* fields, int nshell, int64_t ncell, const int32_t* lane_ij, int lane_layout, int mt, int nt,
                            const int32_t* tri_rc, int ntri, double* sums, void* ws, size_t ws_bytes, void* stream)
{
    if (!fields || !lane_ij || !tri_rc || !sums || !ws) return PSB_ERR_ARG;
    return triangle_sums_tc_pass(fields, nshell, ncell, lane_ij, lane_layout, mt, nt, tri_rc, ntri, sums, ws, ws_bytes, S(stream));
}

/* host helper shared with the Python layer: tri [ntri][3] shell indices -> tiles; call with tiles == NULL to size */
int psb_bk_build_tiles(const int32_t* tri, int ntri, int s0, int32_t* tiles, int* ntiles)
{
    if (!tri || !ntiles || ntri < 0) return PSB_ERR_ARG;
    std::vector<int32_t> v;
    PSB_TRY(build_tiles(tri, ntri, s0, v));
    *ntiles = (int)(v.size() / 68);
    if (tiles) std::memcpy(tiles, v.data(), v.size() * sizeof(int32_t));
    return PSB_OK;
}

int psb_apply_rsd(const double* xyz, const double* v_los, int64_t np, int i_los, double rsd_factor, double lbox, double* out, void* stream)
{
    return apply_rsd(xyz, v_los, np, i_los, rsd_factor, lbox, out, S(stream));
}

int psb_survey_prepare(const double* radecz, const double* nbar, const double* w, int64_t np, const double* dist_table, int nnodes,
                       double zmax, double p0_fkp, float* xyz_f32, float* w_f32, double* out12, void* stream)
{
    if (!radecz || !nbar || !dist_table || !xyz_f32 || !w_f32 || !out12 || nnodes < 2) return PSB_ERR_ARG;
    return psb::survey_prepare(radecz, nbar, w, (long long)np, dist_table, nnodes - 1, zmax, p0_fkp, xyz_f32, w_f32, out12,
                               (cudaStream_t)stream);
}

// ---------------------------------------------------------------- f2py-shaped host drop-ins
int psb_host_assign_quad(const float* r, const float* w, float* dtl, int64_t np, int N, float kf_ks, float offset,
                         int ia, int ib, int ic, int id)
{
    if (!r || !w || !dtl || np < 0) return PSB_ERR_ARG;
    const bool quad = ia || ib || ic || id;
    const size_t nmesh = 2 * (size_t)N * N * N;
    DevBuf dr, dw, dm, dws, ds, dq;
    const size_t wsb = assign_workspace_bytes(np, N);
    PSB_TRY(dr.alloc(sizeof(float) * 3 * np)); PSB_TRY(dw.alloc(sizeof(float) * np)); PSB_TRY(dm.alloc(sizeof(float) * nmesh));
    PSB_TRY(dws.alloc(wsb)); PSB_TRY(ds.alloc(sizeof(double)));
    PSB_CUDA(cudaMemcpy(dr.p, r, sizeof(float) * 3 * np, cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(dw.p, w, sizeof(float) * np, cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(dm.p, dtl, sizeof(float) * nmesh, cudaMemcpyHostToDevice));
    const void* wdev = dw.p;
    if (quad) {                                          // f:294-300: the weight becomes w r_ia r_ib / r^2 (or the four-index form)
        PSB_TRY(dq.alloc(sizeof(float) * np));
        PSB_TRY(quad_weights(dr.as<float>(), dw.as<float>(), np, ia, ib, ic, id, dq.as<float>(), nullptr));
        wdev = dq.p;
    }
    PSB_TRY(psb_assign_pcs_interlaced(dr.p, 0, 1, wdev, 0, np, N, 0.0, kf_ks, offset, dm.as<float>(), 0, dws.p, wsb, ds.as<double>(), nullptr));
    PSB_CUDA(cudaMemcpy(dtl, dm.p, sizeof(float) * nmesh, cudaMemcpyDeviceToHost));
    return PSB_OK;
}

int psb_quad_weights(const float* r, const float* w, int64_t np, int ia, int ib, int ic, int id, float* we, void* stream)
{
    return quad_weights(r, w, np, ia, ib, ic, id, we, S(stream));
}

int psb_quad_fields(int mode, const float* a, const float* b, const float* c, const float* d, float* out, int N, int irsd, void* stream)
{
    return quad_fields(mode, reinterpret_cast<const Cx<float>*>(a), reinterpret_cast<const Cx<float>*>(b), reinterpret_cast<const Cx<float>*>(c),
                       reinterpret_cast<const Cx<float>*>(d), reinterpret_cast<Cx<float>*>(out), N, irsd, S(stream));
}

// host arrays (ngrid/2+1,ngrid,ngrid) complex64 F-order: in[k] may be null; `out` is uploaded, combined in place and read back
static int host_quad_fields(int mode, const float* const in[4], float* out, int N, int irsd)
{
    if (!out || N < 2 || N % 2) return PSB_ERR_ARG;
    const size_t nb = 8 * (size_t)(N / 2 + 1) * N * N;
    DevBuf din[4], dout;
    PSB_TRY(dout.alloc(nb));
    PSB_CUDA(cudaMemcpy(dout.p, out, nb, cudaMemcpyHostToDevice));
    for (int k = 0; k < 4; ++k) {
        if (!in[k]) continue;
        PSB_TRY(din[k].alloc(nb));
        PSB_CUDA(cudaMemcpy(din[k].p, in[k], nb, cudaMemcpyHostToDevice));
    }
    PSB_TRY(psb_quad_fields(mode, din[0].as<float>(), din[1].as<float>(), din[2].as<float>(), din[3].as<float>(), dout.as<float>(), N, irsd, nullptr));
    PSB_CUDA(cudaMemcpy(out, dout.p, nb, cudaMemcpyDeviceToHost));
    return PSB_OK;
}

int psb_host_fivedelta2g_1(float* xx, const float* yy, const float* zz, int N)
{
    if (!yy || !zz) return PSB_ERR_ARG;
    const float* in[4] = { yy, zz, nullptr, nullptr };
    return host_quad_fields(1, in, xx, N, 0);
}

int psb_host_fivedelta2g_2(const float* dcg, float* xx, const float* xy, const float* yz, const float* zx, int N)
{
    if (!dcg || !xy || !yz || !zx) return PSB_ERR_ARG;
    const float* in[4] = { dcg, xy, yz, zx };
    return host_quad_fields(2, in, xx, N, 0);
}

int psb_host_build_quad(const float* d1, float* d2, int irsd, int N)
{
    if (!d1) return PSB_ERR_ARG;
    const float* in[4] = { d1, nullptr, nullptr, nullptr };
    return host_quad_fields(3, in, d2, N, irsd);
}

static int host_fcomb(float* dcl, float n, int N, int periodic)
{
    if (!dcl || N < 2 || N % 2) return PSB_ERR_ARG;
    const int h = N / 2;
    const size_t nfull = (size_t)N * N * N, nhalf = (size_t)(h + 1) * N * N;
    std::vector<double> rec(2 * (h + 1)); std::vector<float> wk(h + 1);
    PSB_TRY(psb_fcomb_tables(N, rec.data(), wk.data()));
    DevBuf df, dh, drec, dwk, ds;
    PSB_TRY(df.alloc(8 * nfull)); PSB_TRY(dh.alloc(8 * nhalf)); PSB_TRY(drec.alloc(16 * (h + 1))); PSB_TRY(dwk.alloc(4 * (h + 1))); PSB_TRY(ds.alloc(8));
    const double nd = (double)n;
    PSB_CUDA(cudaMemcpy(df.p, dcl, 8 * nfull, cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(drec.p, rec.data(), 16 * (h + 1), cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(dwk.p, wk.data(), 4 * (h + 1), cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(ds.p, &nd, 8, cudaMemcpyHostToDevice));
    PSB_TRY(psb_fcomb(df.as<float>(), dh.as<float>(), N, drec.as<double>(), dwk.as<float>(), ds.as<double>(), periodic, nullptr));
    std::vector<float> half(2 * nhalf);
    PSB_CUDA(cudaMemcpy(half.data(), dh.p, 8 * nhalf, cudaMemcpyDeviceToHost));
    // write the half field and its mirror images, as f:658-665 leave them
    Cx<float>* D = reinterpret_cast<Cx<float>*>(dcl);
    const Cx<float>* H = reinterpret_cast<const Cx<float>*>(half.data());
    for (int iz = 0; iz < N; ++iz) for (int iy = 0; iy < N; ++iy) for (int ix = 0; ix <= h; ++ix) {
        const Cx<float> v = H[((size_t)iz * N + iy) * (h + 1) + ix];
        D[((size_t)iz * N + iy) * N + ix] = v;
        if (ix > 0 && ix < h) D[((size_t)kneg(iz, N) * N + kneg(iy, N)) * N + (N - ix)] = conj(v);
    }
    return PSB_OK;
}
int psb_host_fcomb_periodic(float* dcl, float n, int N) { return host_fcomb(dcl, n, N, 1); }
int psb_host_fcomb_survey(float* dcl, int N) { return host_fcomb(dcl, 1.f, N, 0); }

int psb_host_ffting(float* dtl, int N)
{
    if (!dtl || N < 2) return PSB_ERR_ARG;
    const size_t n = (size_t)N * N * N;
    std::vector<float> tw(2 * N);
    PSB_TRY(psb_twiddles_f32(N, tw.data()));
    DevBuf d, dt;
    PSB_TRY(d.alloc(8 * n)); PSB_TRY(dt.alloc(8 * N));
    PSB_CUDA(cudaMemcpy(d.p, dtl, 8 * n, cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(dt.p, tw.data(), 8 * N, cudaMemcpyHostToDevice));
    PSB_TRY(psb_fft_c2c_3d(d.as<float>(), N, +1, dt.as<float>(), nullptr));
    PSB_CUDA(cudaMemcpy(dtl, d.p, 8 * n, cudaMemcpyDeviceToHost));
    return PSB_OK;
}

int psb_host_pk_pbox_rsd(const float* dtl, double* k, double* p0, double* p2, double* p4, double* nk,
                         double* km, double* mk, double* pkm, double* nkm, int irsd, int lbox, int nbin, int nmu, int N)
{
    if (!dtl || !k || !p0 || !p2 || !p4 || !nk || !km || !mk || !pkm || !nkm || nbin < 1 || nmu < 1 || N < 2 || N % 2 || lbox == 0) return PSB_ERR_ARG;
    const int h = N / 2;
    const size_t nhalf = (size_t)(h + 1) * N * N;
    const int mmax = 3 * h * h;
    std::vector<uint16_t> bin(mmax + 1);
    PSB_TRY(psb_rsd_bin_table(N, nbin, mmax, bin.data()));
    float trig[4];
    PSB_TRY(psb_rsd_trig(irsd, trig));
    const float pi = 3.141592654f, tpi = 2.f * pi;
    const float kf = tpi / (float)lbox;                                   // f:169
    const size_t nout = (5 + 4 * (size_t)nmu) * nbin;
    DevBuf dh, db, dout;
    PSB_TRY(dh.alloc(8 * nhalf)); PSB_TRY(db.alloc(2 * (mmax + 1))); PSB_TRY(dout.alloc(8 * nout));
    PSB_CUDA(cudaMemcpy(dh.p, dtl, 8 * nhalf, cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(db.p, bin.data(), 2 * (mmax + 1), cudaMemcpyHostToDevice));
    PSB_TRY(psb_pk_multipoles(dh.as<float>(), N, db.as<uint16_t>(), nbin, nmu, kf, trig, dout.as<double>(), nullptr));
    std::vector<double> o(nout);
    PSB_CUDA(cudaMemcpy(o.data(), dout.p, 8 * nout, cudaMemcpyDeviceToHost));
    // normalisation, f:246-262
    const double kf3 = (double)(kf * kf * kf);
    for (int i = 0; i < nbin; ++i) {
        nk[i] = o[i]; k[i] = o[nbin + i]; p0[i] = o[2 * nbin + i]; p2[i] = o[3 * nbin + i]; p4[i] = o[4 * nbin + i];
        if (nk[i] > 0) { k[i] = k[i] / nk[i]; p0[i] = p0[i] / nk[i] / kf3; p2[i] = p2[i] / nk[i] / kf3; p4[i] = p4[i] / nk[i] / kf3; }
    }
    const size_t tb = (size_t)nbin * nmu;
    for (size_t e = 0; e < tb; ++e) {
        nkm[e] = o[5 * nbin + e]; km[e] = o[5 * nbin + tb + e]; mk[e] = o[5 * nbin + 2 * tb + e]; pkm[e] = o[5 * nbin + 3 * tb + e];
        if (nkm[e] > 0) { km[e] = km[e] / nkm[e]; mk[e] = mk[e] / nkm[e]; pkm[e] = pkm[e] / nkm[e] / kf3; }
    }
    return PSB_OK;
}

int psb_host_bk_counts(double* coun, int N, float step, int ncut, int nmax)
{
    if (!coun || N < 2 || N % 2 || !(step >= 1.f) || nmax < 1) return PSB_ERR_ARG;
    const int ncuts = ncut / (int)step;                                   // f:20
    if (ncuts < 1 || ncuts > nmax) return PSB_ERR_ARG;
    const int h = N / 2, mmax = 3 * h * h;
    const size_t ncell = (size_t)N * N * N;
    std::vector<uint16_t> irk(mmax + 1);
    PSB_TRY(psb_irk_table_f32(step, mmax, irk.data()));
    std::vector<double> tw(2 * N);
    PSB_TRY(psb_twiddles_f64(N, tw.data()));
    const int nsh = nmax - ncuts + 1;
    const int nfield = (nsh + 3) / 4 * 4;
    // triangle list of f:90-92: l outer, j, i inner, i <= j <= l
    std::vector<int32_t> tri;
    for (int l = ncuts; l <= nmax; ++l) for (int j = ncuts; j <= l; ++j) for (int i = std::max(ncuts, l - j); i <= j; ++i) { tri.push_back(i); tri.push_back(j); tri.push_back(l); }
    const int ntri = (int)(tri.size() / 3);
    std::vector<int32_t> tiles;
    PSB_TRY(build_tiles(tri.data(), ntri, ncuts, tiles));
    const int ntiles = (int)(tiles.size() / 68);
    const int R = (int)std::floor(((double)nmax + 0.5) * (double)step) + 1;
    const int Wc = std::min(2 * R + 1, N);
    DevBuf dirk, dtw, dt1, dt2, dfields, dptr, dsq, dtiles, dsums, dws;
    PSB_TRY(dirk.alloc(2 * (mmax + 1))); PSB_TRY(dtw.alloc(16 * N));
    PSB_TRY(dt1.alloc(16 * (size_t)Wc * Wc * N)); PSB_TRY(dt2.alloc(16 * (size_t)Wc * N * N));
    PSB_TRY(dfields.alloc(8 * ncell * (nsh + 1))); PSB_TRY(dptr.alloc(8 * nfield)); PSB_TRY(dsq.alloc(16));
    PSB_TRY(dtiles.alloc(4 * tiles.size())); PSB_TRY(dsums.alloc(8 * (size_t)ntri));
    const size_t wsb = triangle_workspace_bytes(ntiles);
    PSB_TRY(dws.alloc(wsb));
    PSB_CUDA(cudaMemcpy(dirk.p, irk.data(), 2 * (mmax + 1), cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(dtw.p, tw.data(), 16 * N, cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemcpy(dtiles.p, tiles.data(), 4 * tiles.size(), cudaMemcpyHostToDevice));
    PSB_CUDA(cudaMemset(dsq.p, 0, 16));
    std::vector<const double*> ptrs(nfield);
    for (int f = 0; f < nfield; ++f) ptrs[f] = dfields.as<double>() + ncell * (size_t)std::min(f, nsh - 1);
    PSB_CUDA(cudaMemcpy(dptr.p, ptrs.data(), 8 * nfield, cudaMemcpyHostToDevice));
    for (int s = 0; s < nsh; s += 2) {
        const int sa = ncuts + s, sb = (s + 1 < nsh) ? sa + 1 : -1;
        const int Rp = (int)std::floor(((double)(sb >= 0 ? sb : sa) + 0.5) * (double)step) + 1;
        double* fa = dfields.as<double>() + ncell * (size_t)s;
        double* fb = dfields.as<double>() + ncell * (size_t)(s + 1);       // spare plane exists (nsh+1 allocated)
        PSB_TRY(psb_bk_shell_pair_f64(nullptr, dirk.as<uint16_t>(), N, N, sa, sb, Rp, dt1.as<double>(), dt2.as<double>(), fa, fb,
                                      dsq.as<double>(), dtw.as<double>(), nullptr));
    }
    PSB_TRY(psb_bk_triangle_sums_f64(dptr.as<const double*>(), nfield, (int64_t)ncell, dtiles.as<int32_t>(), ntiles, dsums.as<double>(), dws.p, wsb, nullptr));
    std::vector<double> sums(ntri);
    PSB_CUDA(cudaMemcpy(sums.data(), dsums.p, 8 * (size_t)ntri, cudaMemcpyDeviceToHost));
    const double n3 = (double)N * (double)N * (double)N;
    for (int t = 0; t < ntri; ++t) {
        const int i = tri[3 * t], j = tri[3 * t + 1], l = tri[3 * t + 2];
        // exact integer number of closed triangles, scaled back by N^3 (the reference stores the raw sum)
        coun[(size_t)(i - 1) + (size_t)nmax * ((size_t)(j - 1) + (size_t)nmax * (size_t)(l - 1))] = std::llround(sums[t] / n3) * n3;
    }
    return PSB_OK;
}

}  // extern "C"
