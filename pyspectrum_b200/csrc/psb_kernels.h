// psb_kernels.h -- internal host-side entry points of the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "psb_common.cuh"

namespace psb {

// SMs of the current device (148 on B200), cached: grid-stride and persistent kernels size their grids as multiples of it
int sm_count();

struct AssignIn {
    const void* pos;      // positions
    int pos_f64;          // 1: float64, 0: float32
    int pos_aos;          // 1: [Np][3] (Fortran (3,Np)), 0: [3][Np] (numpy 3xN C order), 2: float4 {x,y,z,w} (routed slab particles)
    const void* w;        // weights or null
    int w_f64;            // 1: float64, 0: float32
    long long Np;
    int N;
    int do_clip;          // pyspectrum.py:938-941
    double clip_hi;       // Lbox*(1-1e-6)
    float kf_ks, offset;
    int zbase, nzs;       // planes [zbase, zbase+nzs) held by `mesh` (full grid: 0, N); see k_assign_tri
    int tiles;            // set by assign_pcs_interlaced: sort by tile (tile scatter) instead of by (z,y) row
};

size_t assign_workspace_bytes(long long Np, int N);
int assign_pcs_interlaced(const AssignIn& in, float* mesh, int zero_mesh, void* ws, size_t ws_bytes, double* sumw, cudaStream_t st);

// slab routing of the multi-GPU path (psb_assign.cu): per-destination counts, then the destination-major float4 send buffer
int slab_route_count(const AssignIn& in, int nz_per_rank, int nranks, unsigned long long* counts, double* sumw, cudaStream_t st);
int slab_route_scatter(const AssignIn& in, int nz_per_rank, int nranks, const unsigned long long* base, unsigned long long* cursor,
                       float4* send, cudaStream_t st);
int slab_route_scatter_peer(const AssignIn& in, int nz_per_rank, int nranks, const long long* dest_addr, unsigned long long* cursor,
                            cudaStream_t st);

int fft_mesh_to_delta(Cx<float>* mesh, Cx<float>* half, int N, const Cx<float>* tw,
                      const Cx<double>* rec, const float* Wk, const double* sumw, int periodic, cudaStream_t st);
int fft_c2c_3d(Cx<float>* data, int N, int dir, const Cx<float>* tw, cudaStream_t st);
int fcomb_standalone(const Cx<float>* full, Cx<float>* half, int N, const Cx<double>* rec, const float* Wk,
                     const double* sumw, int periodic, cudaStream_t st);
template <typename T>
int fft_shell_pair(const Cx<float>* half, const unsigned short* irk, int N, int Ns, int sa, int sb, int R,
                   Cx<T>* t1, Cx<T>* t2, T* fa, T* fb, double* sumsq, const float* scale2, unsigned int* maxabs2, int halfpack,
                   const Cx<T>* tw, cudaStream_t st, const long long* route = nullptr, int rplanes = 0, int rranks = 0);

struct SpectraIn {
    const Cx<float>* half;       // [kz][ky][kx], ky = ky0 .. ky0+ny-1 (the whole half field: ky0 = 0, ny = N)
    int N;
    int ky0, ny;
    const unsigned short* bin;   // bin index (1-based, 0 = none) of m = |k|^2, host table
    int Nbin;
    int mode;                    // 0: Pk_periodic (float64 |k|, realified self-conjugate points)
                                 // 1: pk_pbox_rsd (float32 per-mode math of estimator.f:196-244)
    double kf;                   // mode 0: 2*pi/Lbox (float64)
    float kf32;                  // mode 1: tpi/Lbox in single (f:169)
    int Nmu;
    float costh, sinth, cosph, sinph;   // f:172-181 evaluated with the host libm
};
// out layout (all float64): mode 0: nk[Nbin], ksum[Nbin], psum[Nbin]
//                           mode 1: nk,k,p0,p2,p4 [Nbin] then nkm,km,mk,pkm [Nmu][Nbin] (Fortran (Nbin,Nmu))
int binned_spectra(const SpectraIn& in, double* out, cudaStream_t st);
// code='python' (k,mu) estimator on a full field [kx][ky][kz] (pyspectrum.py:545-626); trig4 = cos/sin theta_obs, cos/sin phi_obs (float64)
int kmu_python(const Cx<float>* full, int N, const unsigned short* bin, int Nbin, int Nmu, double kf, const double* trig4, double* out,
               cudaStream_t st);
int shell_power(const Cx<float>* half, int N, const unsigned short* irk, int nshell, double* psum, cudaStream_t st);
int half_extract(const Cx<float>* src, int N, int ky0, int ny, Cx<float>* dst, int Ng, cudaStream_t st);
int shell_scales(const double* psum, int nshell, float target_rms, float* scales, cudaStream_t st);
int shell_mode_counts(int N, const unsigned short* irk, int nshell_max, unsigned long long* nk, cudaStream_t st);

// tiles: int32 [ntiles][68] = {i0,j0,l0,pad, slot[64]}; fields[] holds S device pointers (S padded so i0+3 < S)
template <typename T>
int triangle_sums_tiles(const T* const* fields, int S, long long ncell, const int* tiles, int ntiles,
                        double* sums, void* ws, size_t ws_bytes, int packed_half, cudaStream_t st);
size_t triangle_workspace_bytes(int ntiles);

// tensor-core K6 (psb_triangles_tc.cu): one pass over MT*128 pair rows x NT shell columns
size_t triangle_tc_workspace_bytes(int MT, int NT);
int triangle_sums_tc_pass(const float* const* fields, int S, long long ncell, const int* lane_ij, int lane_layout, int MT, int NT,
                          const int* tri_rc, int ntri, double* sums, void* ws, size_t ws_bytes, cudaStream_t st);

// slab-decomposed mesh -> delta(k) building blocks of the sharded path (psb_fft_mesh.cu)
int fft_slab_xy(Cx<float>* data, int N, int nz, int dir, const Cx<float>* tw, cudaStream_t st);
int fft_slab_z(Cx<float>* data, int N, int ny, int nx, int dir, const Cx<float>* tw, cudaStream_t st);
int slab_split_ab(const Cx<float>* d, Cx<float>* P, Cx<float>* Q, int N, int nz, int hp, cudaStream_t st, const long long* route = nullptr,
                  int zbase = 0, int nranks = 0);
int slab_fcomb(const Cx<float>* P, const Cx<float>* Q, Cx<float>* half, int N, int ky0, int ny, int hp, const Cx<double>* rec,
               const float* Wk, const double* sumw, int periodic, cudaStream_t st);

// quadrupole-field pieces (psb_quad.cu): Q_ij / Q_ijkl particle weights (f:294-300), FiveDelta2g_1 / _2 / build_quad (f:514-603)
int quad_weights(const float* r, const float* w, long long np, int ia, int ib, int ic, int id, float* we, cudaStream_t st);
int quad_fields(int mode, const Cx<float>* a, const Cx<float>* b, const Cx<float>* c, const Cx<float>* d, Cx<float>* out, int N, int irsd,
                cudaStream_t st);

// survey-geometry catalogue pre-step (psb_survey.cu)
int survey_prepare(const double* radecz, const double* nb, const double* w, long long np, const double* tab, int nn, double zmax,
                   double p0_fkp, float* xyz, float* wout, double* out12, cudaStream_t st);

int apply_rsd(const double* xyz, const double* v_los, long long np, int i_los, double fac, double L, double* out, cudaStream_t st);

}  // namespace psb
