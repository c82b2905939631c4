/*
 * psb200.h -- C ABI of libpsb200.so, the B200 (sm_100a) replacement for pySpectrum's native layer.
 *
 * What it replaces (reference = changhoonhahn/pySpectrum):
 *   - the f2py extension module `estimator` built from pyspectrum/estimator.f (setup.py:11-52,
 *     imported at pyspectrum/pyspectrum.py:8), i.e. assign_quad, fcomb_periodic, fcomb_survey,
 *     pk_pbox_rsd, bk_counts, ffting;
 *   - the FFTW calls made through pyfftw (pyspectrum/pyspectrum.py:204-214, 389-399, 1000-1010,
 *     1064-1075) and through sfftw_* (estimator.f:69-87, 275-280).
 *
 * Two families of entry points:
 *   psb_host_*   same argument meaning as the f2py signatures (SURVEY 8b level 2): HOST pointers,
 *                Fortran-ordered arrays, in/out semantics preserved; host<->device copies inside.
 *                These are what a maintainer binds in place of `import estimator`.
 *   psb_*        DEVICE pointers + a CUDA stream: the resident pipeline the Python API
 *                (pyspectrum_b200.pyspectrum) drives; nothing is allocated behind the caller.
 *
 * Conventions: every function returns 0 on success or a negative PSB_ERR_* code and never throws;
 * `stream` is a cudaStream_t passed as void*; complex arrays are interleaved (re,im) float pairs;
 * "c64" = complex64.  There is no CPU fallback: without a CUDA device every compute call fails
 * with PSB_ERR_CUDA.
 */
#ifndef PSB200_H
#define PSB200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSB_OK 0
#define PSB_ERR_ARG (-1)            /* bad argument (odd grid, unsupported option, null pointer)        */
#define PSB_ERR_UNSUPPORTED_N (-2)  /* grid size has a prime factor other than 2,3,5 or is too large    */
#define PSB_ERR_CUDA (-3)           /* CUDA runtime error (no device, launch failure, out of memory)    */
#define PSB_ERR_WORKSPACE (-4)      /* workspace smaller than the corresponding *_workspace_bytes()     */

int psb_version(void);
const char* psb_error_string(int code);

/* ------------------------------------------------------------------------------------------------
 * Host-side table builders (pure host code, no GPU needed).
 * ---------------------------------------------------------------------------------------------- */
/* tw[k] = exp(+2 pi i k/N), k = 0..N-1, evaluated in double; f32 or f64 interleaved output. */
int psb_twiddles_f32(int ngrid, float* tw_c64);
int psb_twiddles_f64(int ngrid, double* tw_c128);
/* fcomb tables (estimator.f:615-645): rec[j] double-complex phase recurrence with single-rounded base,
 * wk[j] single-precision sinc^4 window, j = 0..N/2. */
int psb_fcomb_tables(int ngrid, double* rec_c128, float* wk);
/* line-of-sight trig of estimator.f:172-181 with the host libm: trig4 = {cos th, sin th, cos ph, sin ph} */
int psb_rsd_trig(int irsd, float* trig4);
/* bin index imk = nint(Nbin*sqrt(m)/(N/2)) in single precision (estimator.f:206-207) for m = 0..mmax */
int psb_rsd_bin_table(int ngrid, int nbin, int mmax, uint16_t* bin_of_m);
/* shell index int(sqrt(m)/step + 0.5) in single precision (estimator.f:32-36) for m = 0..mmax */
int psb_irk_table_f32(float step, int mmax, uint16_t* irk_of_m);

/* ------------------------------------------------------------------------------------------------
 * Device-pointer pipeline.
 * ---------------------------------------------------------------------------------------------- */
/* K1  estimator.f:284-512 (+ clip/cast of pyspectrum.py:938-941 when lbox_clip > 0).
 *   pos      positions: float64 or float32; layout [3][np] (pos_aos=0) or [np][3] (pos_aos=1)
 *   w        weights (float64/float32) or NULL (all ones)
 *   mesh     float32 [N][N][N][2]: (A,B) interlaced = the reference's dtl(2*Ngrid,Ngrid,Ngrid)
 *   zero_mesh  1: mesh is cleared first (FFT_periodic);  0: accumulate (f2py intent(inout))
 *   sumw     device double: receives sum(w) in float64 (pyspectrum.py:957 np.sum(w)) */
size_t psb_assign_workspace_bytes(int64_t np, int ngrid);
int psb_assign_pcs_interlaced(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64,
                              int64_t np, int ngrid, double lbox_clip, float kf_ks, float offset,
                              float* mesh, int zero_mesh, void* ws, size_t ws_bytes, double* sumw, void* stream);

/* Slab-owned assignment for ONE catalogue sharded over G GPUs: rank q owns the mesh planes [q*nz, (q+1)*nz), nz = N/G.
 * A particle in cell c touches planes c-1 .. c+3, so it goes to the owner of plane c-1 and, if different, to the owner of plane
 * c+3 (ghost copy); no mesh collective is needed afterwards.
 *   psb_slab_route_count    counts [G] (device uint64): particles (ghost copies included) this rank sends to every rank;
 *                           sumw: sum of the weights of the rank's own particles (no copies) in float64
 *   psb_slab_route_scatter  send_xyzw: destination-major buffer of float4 {x,y,z,w} (float32 after the float64 clip of
 *                           py:938-941); base [G] = exclusive scan of counts (device), cursor [G] scratch
 *   (caller: all-to-all of the counts, then of the float4 rows)
 *   psb_assign_slab         K1 on the received particles onto the rank's planes: mesh_slab float32 [nzs][N][N][2]; contributions
 *                           to planes outside [zbase, zbase+nzs) are dropped (the owner has its own copy of the particle);
 *                           workspace psb_assign_workspace_bytes(np, ngrid); sumw_scratch receives the sum over the received copies */
int psb_slab_route_count(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64, int64_t np, int ngrid, double lbox_clip,
                         float kf_ks, float offset, int nz_per_rank, int nranks, uint64_t* counts, double* sumw, void* stream);
int psb_slab_route_scatter(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64, int64_t np, int ngrid, double lbox_clip,
                           float kf_ks, float offset, int nz_per_rank, int nranks, const uint64_t* base, uint64_t* cursor,
                           float* send_xyzw, void* stream);
/* The same with the exchange fused in: every copy is stored straight into the receive buffer of its destination rank (NVLink peer
 * stores; the copies of a 2048-particle tile leave destination by destination in 16-byte-contiguous runs).  dest_addr = device
 * int64 [G]: the address at which THIS rank's segment starts in rank d's receive buffer (peer pointer + 16 B x the particles the
 * ranks in front of this one send to d: needs the all-gathered counts); cursor [G] scratch.  The caller brackets the call with a
 * barrier over the ranks (the buffers must be free before, complete after). */
int psb_slab_route_scatter_peer(const void* pos, int pos_f64, int pos_aos, const void* w, int w_f64, int64_t np, int ngrid, double lbox_clip,
                                float kf_ks, float offset, int nz_per_rank, int nranks, const int64_t* dest_addr, uint64_t* cursor,
                                void* stream);
int psb_assign_slab(const float* xyzw, int64_t np, int ngrid, float kf_ks, float offset, int zbase, int nzs, float* mesh_slab,
                    int zero_mesh, void* ws, size_t ws_bytes, double* sumw_scratch, void* stream);

/* Survey-geometry catalogue pre-step, one pass on the device (pyspectrum.py:776-806 + util.py:27-51):
 * (RA, Dec, z) -> comoving Cartesian float32 positions, FKP weights, normalisation sums, bounding box.
 *   radecz      device float64 [3][np]: RA [deg], Dec [deg], redshift
 *   nbar, w     device float64 [np]; w may be NULL (all ones)
 *   dist_table  device float64 [nnodes][2]: cubic Hermite nodes over z in [0, zmax], {D(z_k), D'(z_k) dz} with D the
 *               line-of-sight comoving distance in Mpc/h of the caller's cosmology and dz = zmax/(nnodes-1)
 *   xyz_f32     out, device float32 [3][np] (the layout psb_assign_pcs_interlaced takes with pos_f64=0, pos_aos=0)
 *   w_f32       out, device float32 [np]: w / (1 + nbar p0_fkp)
 *   out12       out, device float64 [12]: Ntot = sum w, I12, I13, I22, I23, I33 (py:794, 802-806), min x,y,z, max x,y,z */
int psb_survey_prepare(const double* radecz, const double* nbar, const double* w, int64_t np, const double* dist_table, int nnodes,
                       double zmax, double p0_fkp, float* xyz_f32, float* w_f32, double* out12, void* stream);
/* util.applyRSD (util.py:54-75) on device arrays: out [3][np] float64 = xyz [3][np] with row i_los replaced by
 * (x + (rsd_factor * v_los + lbox)) mod lbox, numpy's float64 arithmetic and np.remainder bit for bit; v_los [np] = the velocity row
 * along the line of sight (km/s), rsd_factor = (1+z)/(100 E(z)) from the caller's cosmology. */
int psb_apply_rsd(const double* xyz, const double* v_los, int64_t np, int i_los, double rsd_factor, double lbox, double* out, void* stream);

/* K2+K3  pyspectrum.py:1060-1080 + estimator.f:605-675 + the [:N/2+1] slice (py:959).
 *   mesh_c64  in: (A + iB) on [z][y][x]; destroyed (x and y passes run in place)
 *   half_c64  out: delta(k) on [kz][ky][kx], kx = 0..N/2 (the reference's Fortran (N/2+1,N,N) array)
 *   periodic  1: fcomb_periodic (divide by *sumw), 0: fcomb_survey */
int psb_fft_mesh_to_delta(float* mesh_c64, float* half_c64, int ngrid, const float* tw_c64,
                          const double* rec_c128, const float* wk, const double* sumw, int periodic, void* stream);
/* plain in-place unnormalised 3-D c2c (dir=+1: FFTW_BACKWARD); estimator.f:266-282 */
int psb_fft_c2c_3d(float* data_c64, int ngrid, int dir, const float* tw_c64, void* stream);
/* fcomb alone on a transformed full grid -> half field */
int psb_fcomb(const float* full_c64, float* half_c64, int ngrid, const double* rec_c128, const float* wk,
              const double* sumw, int periodic, void* stream);

/* Slab-decomposed K2+K3 for ONE catalogue sharded over G GPUs (pyspectrum_b200/multigpu.py).  A rank owns nz = N/G
 * z-planes of the reduced mesh (A + iB) and, after the exchange, ny = N/G ky-rows of k-space; hp = N/2+1 rounded up to even.
 *   psb_fft_slab_xy    in-place x and y passes of data [nz][N][N] (dir=+1: FFTW_BACKWARD as py:1064-1075)
 *   psb_slab_split_ab  d [nz][N][N] -> p = A^xy, q = B^xy on [nz][N][hp] (kx = 0..N/2; the conjugate partner is in the same plane)
 *   (caller: all-to-all  [nz][N][hp] -> [N][ny][hp])
 *   psb_fft_slab_z     in-place z pass of an array [N][ny][nx] (nx even)
 *   psb_slab_fcomb     p, q [N][ny][hp] -> rows ky0..ky0+ny-1 of the half field, half [N][ny][N/2+1], through the same closed
 *                      form of estimator.f:605-745 as psb_fft_mesh_to_delta (F(k) = p + iq, F(-k) = conj p + i conj q) */
int psb_fft_slab_xy(float* data_c64, int ngrid, int nz, int dir, const float* tw_c64, void* stream);
int psb_fft_slab_z(float* data_c64, int ngrid, int ny, int nx, int dir, const float* tw_c64, void* stream);
int psb_slab_split_ab(const float* d_c64, float* p_c64, float* q_c64, int ngrid, int nz, int hp, void* stream);
/* The same with the exchange fused in (z-slabs -> ky-slabs through peer memory instead of an all-to-all): row ky of plane z is stored
 * straight into rank ky/(N/nranks)'s arrays p, q [N][N/nranks][hp] at plane zbase + z.  route = device int64 [2][nranks]: the address of
 * every rank's p array, then of every rank's q array (peer pointers, e.g. from torch.distributed._symmetric_memory). */
int psb_slab_split_ab_routed(const float* d_c64, int ngrid, int nz, int hp, int zbase, int nranks, const int64_t* route, void* stream);
int psb_slab_fcomb(const float* p_c64, const float* q_c64, float* half_c64, int ngrid, int ky0, int ny, int hp,
                   const double* rec_c128, const float* wk, const double* sumw, int periodic, void* stream);

/* K4  power spectra from the half field.
 * psb_pk_monopole   : pyspectrum.py:690-716.  out (float64) = nk[nbin], sum|k|[nbin], sum|delta|^2[nbin]
 * psb_pk_multipoles : estimator.f:196-244 (raw sums, before f:246-262).  out (float64) =
 *                     nk,k,p0,p2,p4 [nbin] then nkm,km,mk,pkm [nmu][nbin] (= Fortran (nbin,nmu)) */
int psb_pk_monopole(const float* half_c64, int ngrid, const uint16_t* bin_of_m, int nbin, double kf,
                    double* out, void* stream);
int psb_pk_multipoles(const float* half_c64, int ngrid, const uint16_t* bin_of_m, int nbin, int nmu,
                      float kf32, const float* trig4_host, double* out, void* stream);

/* The same on a rank's ky-slab of the half field, [kz][ky0 .. ky0+ny-1][kx] (multi-GPU: partial sums, all-reduced by the caller). */
int psb_pk_monopole_slab(const float* half_slab_c64, int ngrid, int ky0, int ny, const uint16_t* bin_of_m, int nbin, double kf,
                         double* out, void* stream);
int psb_pk_multipoles_slab(const float* half_slab_c64, int ngrid, int ky0, int ny, const uint16_t* bin_of_m, int nbin, int nmu,
                           float kf32, const float* trig4_host, double* out, void* stream);
/* Low-|k| modes of a ky-slab -> half field [ngrid_carrier][ngrid_carrier][ngrid_carrier/2+1] of a coarser (or equal) grid, zero
 * elsewhere; |k_a| < ngrid_carrier/2 are kept (everything if the grids are equal).  The sum over the ranks is the carrier field
 * the shell stage transforms from (see psb_bk_shell_pair_f32, ngrid_src). */
int psb_half_extract(const float* half_slab_c64, int ngrid, int ky0, int ny, float* carrier_c64, int ngrid_carrier, void* stream);

/* code='python' branch of _Pk_periodic_rsd (pyspectrum.py:545-626): float64 (k,mu) estimator over ALL modes of a FULL field
 * full_c64 [kx][ky][kz] (what reflect_delta returns, C order) with |k_a| = min(i, N-i), mu bin ceil(mu * nmu);
 * bin_of_m = int(nbin*rk/phys_nyq + 0.5) host table (py:560), trig4_host = {cos theta_obs, sin theta_obs, cos phi_obs, sin phi_obs}
 * in float64 (py:563-564, 577-578).  out (float64) = nk, sum rk, sum|d|^2, sum|d|^2 L2, sum|d|^2 L4 [nbin] (the last two over the
 * mu-binned modes only, as py:606-607), then N_kmu, sum rk, sum mu, sum|d|^2 as [nbin][nmu]. */
int psb_pk_kmu_python(const float* full_c64, int ngrid, const uint16_t* bin_of_m, int nbin, int nmu, double kf,
                      const double* trig4_host, double* out, void* stream);

/* K5  pyspectrum.py:373-404.  nk[j] = number of modes of the FULL grid in shell j (py:380). */
int psb_shell_mode_counts(int ngrid, const uint16_t* irk_of_m, int nshell, uint64_t* nk, void* stream);
/* One packed pair of shells (sa -> real part, sb -> imaginary part, sb<0: none), pruned to |k_i| <= R:
 *   fa,fb   out: I_sa(x), I_sb(x) on [z][y][x] (fb may be NULL when sb < 0)
 *   sumsq   device double[2]: += sum_x I_sa^2, sum_x I_sb^2   (py:404)
 *   t1,t2   scratch of (2R+1)^2*N and (2R+1)*N^2 complex elements (clamped to N per axis)
 *   half_c64 == NULL -> delta == 1 (triangle counts, py:977 / estimator.f:74-80)
 *   ngrid_src  grid of half_c64 (0 or ngrid: the same).  ngrid_src > ngrid transforms the shells on a COARSER grid than delta(k)
 *              lives on: a shell field is band limited to |k_i| <= R, so for 2R < ngrid the coarse transform is the same function
 *              sampled at fewer points, and sum_x I_i I_j I_l / ngrid^3 is unchanged as long as R_i + R_j + R_l < ngrid (no
 *              triangle can close through a wrap on either grid).  irk_of_m is the table of the TRANSFORM grid (it depends on m only). */
int psb_bk_shell_pair_f32(const float* half_c64, const uint16_t* irk_of_m, int ngrid, int ngrid_src, int sa, int sb, int R,
                          float* t1_c64, float* t2_c64, float* fa, float* fb, double* sumsq,
                          const float* scale2, uint32_t* maxabs2, int pack_half, const float* tw_c64, void* stream);
/* The same with ROUTED output (multi-GPU, SURVEY 8e: "z-pass epilogue -> peer-memory slab write"): the output planes z are cut
 * into slabs of planes_per_rank planes, slab q belongs to rank q, and the z pass stores every plane straight into rank q's memory
 * (a peer pointer mapped over NVLink, e.g. from torch.distributed._symmetric_memory) instead of a local field + an all-to-all.
 *   route  device int64 [2][nranks]: for plane a (shell sa) and plane b (shell sb), the byte address the field WOULD start at on rank
 *          q if that rank's buffer held all planes, i.e. (rank q's slab row of this shell) - q * planes_per_rank * N^2 * 4; 0 = not stored.
 * Needs a compiled multi-stage plan for ngrid (256, 320, 360, 400, 512, 1024, ...); otherwise PSB_ERR_UNSUPPORTED_N. */
int psb_bk_shell_pair_f32_routed(const float* half_c64, const uint16_t* irk_of_m, int ngrid, int ngrid_src, int sa, int sb, int R,
                                 float* t1_c64, float* t2_c64, const int64_t* route, int planes_per_rank, int nranks, double* sumsq,
                                 const float* scale2, uint32_t* maxabs2, int pack_half, const float* tw_c64, void* stream);
/* pack_half != 0: each aligned pair of cells (x, x+1) of fa/fb is stored as the two 32-bit words
 * {half2 hi(x,x+1), half2 lo(x,x+1)} with hi = fp16(v), lo = fp16(v - hi) (same 4 bytes per cell; the layout the
 * tensor-core triangle kernel consumes; requires scale2 so that the values sit in fp16 range).
 * Optional exact power-of-two normalisation of the stored shell fields (needed by the fp16-split tensor-core
 * triangle kernel): scale2 = device float[2] multiplying (I_sa, I_sb) before the store and before sumsq;
 * maxabs2 = device uint32[2] receiving max|stored value| as float bits (atomicMax; zero it first).
 * psb_bk_shell_scales turns per-shell power sums (psb_pk_monopole with the shell table as bin table:
 * psum[j] = sum_{k in shell j+1} |delta|^2) into scales[j] = 2^round(log2(target_rms / sqrt(psum[j]))). */
/* psum[j] = sum_{k in shell j+1, full grid} |delta(k)|^2, j = 0..nshell-1 (Parseval: = sum_x I_{j+1}^2 / N^3), float64 atomics
 * (last-bit order dependent) -- it only feeds the power-of-two scale, the shell power that is reported comes from K5; nshell <= 1024 */
int psb_bk_shell_power(const float* half_c64, int ngrid, const uint16_t* irk_of_m, int nshell, double* psum, void* stream);
int psb_bk_shell_scales(const double* psum, int nshell, float target_rms, float* scales, void* stream);
int psb_bk_shell_pair_f64(const float* half_c64, const uint16_t* irk_of_m, int ngrid, int ngrid_src, int sa, int sb, int R,
                          double* t1_c128, double* t2_c128, double* fa, double* fb, double* sumsq,
                          const double* tw_c128, void* stream);

/* K6  pyspectrum.py:415-430.  fields: device array of nfields device pointers (field slot = shell - s0),
 * tiles: int32 [ntiles][68] = {i0,j0,l0,0, slot[64]} with slot[(a*4+b)*4+c] = index into sums[] of
 * triangle (i0+a, j0+b, l0+c) or -1; i0+3, j0+3, l0+3 must be < nfields (pad the pointer array).
 * sums (float64, device) receives sum_x I_i I_j I_l.  packed_half != 0 (f32 only): the fields are in the packed hi/lo layout
 * of psb_bk_shell_pair_f32(pack_half=1) and are decoded on load. */
size_t psb_bk_triangle_workspace_bytes(int ntiles);
int psb_bk_triangle_sums_f32(const float* const* fields, int nfields, int64_t ncell, const int32_t* tiles,
                             int ntiles, double* sums, void* ws, size_t ws_bytes, int packed_half, void* stream);
int psb_bk_triangle_sums_f64(const double* const* fields, int nfields, int64_t ncell, const int32_t* tiles,
                             int ntiles, double* sums, void* ws, size_t ws_bytes, void* stream);

/* K6 on the tensor cores (tcgen05: fp16 hi/lo split operands, A operand in TMEM, fp32/float64 drains).
 * One pass covers 128 "lanes" x mt tiles of pair rows against all `nshell` fields (columns padded to nt, a multiple
 * of 16, <= 128):  lane_ij = int32 [128][5] = {field slot i of the lane, field slot j of its row in tile 0..3}
 * (-1 = padding).  Row (m*128 + lane) is the pair (i, j_m).  lane_layout = 1 (needs mt == 4): 2x2 row blocks,
 * lane_ij = {i0, i1, j0, j1, -}, rows (i0,j0), (i0,j1), (i1,j0), (i1,j1) in tiles 0..3.  tri_rc[t] = (row, column) of triangle t in this pass or
 * (-1,-1); sums[t] is written for the triangles of this pass.  Needs ncell % 64 == 0, mt <= 4 (nt <= 64) or 2.
 * Fields: the packed hi/lo halves written by psb_bk_shell_pair_f32(pack_half=1), pre-scaled so that |I_i I_j| stays inside
 * the fp16 range (psb_bk_shell_scales). */
size_t psb_bk_triangle_tc_workspace_bytes(int mt, int nt);
int psb_bk_triangle_sums_tc(const float* const* fields, int nshell, int64_t ncell, const int32_t* lane_ij, int lane_layout,
                            int mt, int nt, const int32_t* tri_rc, int ntri, double* sums, void* ws, size_t ws_bytes,
                            void* stream);

/* host helper: triangle list [ntri][3] (shell indices i,j,l) -> tile descriptors; tiles == NULL only sizes */
int psb_bk_build_tiles(const int32_t* tri_ijl, int ntri, int s0, int32_t* tiles, int* ntiles);

/* Quadrupole-field pieces on DEVICE arrays.
 *   psb_quad_weights: we[i] of estimator.f:294-300 from r = Fortran (3,np) float32 and w; feed `we` to psb_assign_pcs_interlaced.
 *   psb_quad_fields:  mode 1 FiveDelta2g_1 (out=dcgxx, a=dcgyy, b=dcgzz), mode 2 FiveDelta2g_2 (out=dcgxx, a=dcg, b=dcgxy, c=dcgyz,
 *                     d=dcgzx), mode 3 build_quad (out=dclr2, a=dclr1, irsd 1..3); half arrays [kz][ky][kx<=N/2] complex64. */
int psb_quad_weights(const float* r_aos, const float* w, int64_t np, int ia, int ib, int ic, int id, float* we, void* stream);
int psb_quad_fields(int mode, const float* a_c64, const float* b_c64, const float* c_c64, const float* d_c64, float* out_c64,
                    int ngrid, int irsd, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Host-buffer drop-ins: the f2py signatures of `estimator` (f2py -h, SURVEY 8b level 2).
 * All arrays are HOST memory in Fortran order exactly as f2py would hand them to the Fortran.
 * ---------------------------------------------------------------------------------------------- */
/* assign_quad(r,w,dtl,kf_ks,offset,ia,ib,ic,id,[np,ngrid])   estimator.f:284
 *   r (3,np) float32, w (np) float32, dtl (2*ngrid,ngrid,ngrid) float32 intent(inout).
 *   ia=ib=ic=id=0: the delta branch (f:292-293); ia,ib in 1..3 with ic=id=0: Q_ij weights w r_ia r_ib / r^2 (f:294-296);
 *   all four in 1..3: Q_ijkl weights w r_ia r_ib r_ic r_id / r^4 (f:297-299).  Any other combination indexes r(0,i) in the
 *   Fortran (undefined) and returns PSB_ERR_ARG here. */
int psb_host_assign_quad(const float* r, const float* w, float* dtl, int64_t np, int ngrid,
                         float kf_ks, float offset, int ia, int ib, int ic, int id);
/* fcomb_periodic(dcl,n,[ngrid]) estimator.f:605; fcomb_survey(dcl,[ngrid]) estimator.f:677; in place,
 * all 8 mirror images written as the Fortran does. */
int psb_host_fcomb_periodic(float* dcl_c64, float n, int ngrid);
int psb_host_fcomb_survey(float* dcl_c64, int ngrid);
/* ffting(dtl,n,[ngrid]) estimator.f:266: in-place FFTW_BACKWARD c2c */
int psb_host_ffting(float* dtl_c64, int ngrid);
/* pk_pbox_rsd(dtl,irsd,lbox,nbin,nmu,[ngrid]) -> k,p0,p2,p4,nk,km,mk,pkm,nkm   estimator.f:155
 *   dtl (ngrid/2+1,ngrid,ngrid) complex64; lbox INTEGER; outputs float64, (nbin) and (nbin,nmu) F-order */
int psb_host_pk_pbox_rsd(const float* dtl_c64, double* k, double* p0, double* p2, double* p4, double* nk,
                         double* km, double* mk, double* pkm, double* nkm,
                         int irsd, int lbox, int nbin, int nmu, int ngrid);
/* FiveDelta2g_1(dcgxx,dcgyy,dcgzz,ngrid) estimator.f:514, FiveDelta2g_2(dcg,dcgxx,dcgxy,dcgyz,dcgzx,ngrid) estimator.f:541,
 * build_quad(dclr1,dclr2,irsd,ngrid) estimator.f:574: (ngrid/2+1,ngrid,ngrid) complex64 arrays; dcgxx / dclr2 are intent(inout).
 * Reproduced as written, including the implicitly INTEGER unit-vector components of FiveDelta2g_* (see psb_quad.cu);
 * irsd outside 1..3 (the Fortran stops) returns PSB_ERR_ARG. */
int psb_host_fivedelta2g_1(float* dcgxx_c64, const float* dcgyy_c64, const float* dcgzz_c64, int ngrid);
int psb_host_fivedelta2g_2(const float* dcg_c64, float* dcgxx_c64, const float* dcgxy_c64, const float* dcgyz_c64,
                           const float* dcgzx_c64, int ngrid);
int psb_host_build_quad(const float* dclr1_c64, float* dclr2_c64, int irsd, int ngrid);
/* bk_counts(coun,nside,step,ncut,[nmax])  estimator.f:2: coun (nmax,nmax,nmax) float64 F-order, filled for
 * ncut/step <= i <= j <= l with sum_x N_i N_j N_l = nside^3 * (exact integer; computed in float64). */
int psb_host_bk_counts(double* coun, int nside, float step, int ncut, int nmax);

#ifdef __cplusplus
}
#endif
#endif /* PSB200_H */
