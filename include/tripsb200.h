/* libtripsb200 - C ABI of the B200 (sm_100a) Krylov hot path for TRIPs-Py.
 *
 * This is the drop-in boundary: plain C, device pointers + sizes + a cudaStream_t (passed as void*), no
 * torch / numpy / C++ types.  The reference (mpasha3/trips-py) is pure Python; the "FFI" a maintainer binds is
 * ctypes.  Each entry point names the reference call site(s) it replaces; paths are
 * relative to the reference repository root.
 *
 * Conventions
 *   - return value: 0 = ok; 1..999 = cudaError_t; >= 1000 = TB200_E*; message via tb200_last_error().
 *   - every pointer is a DEVICE pointer unless the parameter name ends in _host.
 *   - all launches are asynchronous on `stream`; the library allocates nothing, keeps no pointer after the
 *     call returns and never frees caller memory.  Workspaces are caller-owned (tb200_*_workspace_len()).
 *   - a "basis" is kmax columns, column j contiguous at V + j*ld (ld >= column length).
 *   - scalars that may live on the device come as a (x_host, x_dev) pair: x_dev wins when it is not NULL.
 *   - norm outputs are two doubles: out[0] = sum of squares, out[1] = its square root.
 *   - norms and dot products are accumulated in double-double and return the CORRECTLY ROUNDED exact sum (no
 *     atomics): bitwise reproducible run to run, independent of grid size, and reproducible by a CPU oracle.
 */
#ifndef TRIPSB200_H
#define TRIPSB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TB200_EINVAL 1001
#define TB200_ENOTSM100 1002

/* ---- library ------------------------------------------------------------------------------------------ */
const char* tb200_last_error(void);
int tb200_version(void);
int tb200_device_info(int* sm_count, int64_t* l2_bytes, int64_t* mem_bytes, int* cc);
int tb200_require_sm100(void);
/* Measured fp64 instruction-issue peak (the roofline denominator of the matrix-free CT projectors, which read almost
 * no DRAM): one launch of a DFMA-chain microbenchmark (8 independent chains per thread, 8 CTAs x 256 threads per SM, no
 * memory traffic) executing tb200_fp64_peak_instructions(iters) thread-level fp64 instructions.  The caller times it
 * with CUDA events (bench.py).  No reference counterpart: measurement support. */
int64_t tb200_fp64_peak_instructions(int iters);
int tb200_fp64_peak_run(int iters, double* sink, void* stream);

/* ---- CSR SpMV: y = A x - coef * z, optional fused ||y||^2 ---------------------------------------------
 * Replaces `A @ v` and (on the explicitly stored transpose) `A.T @ u`:
 *   trips/utilities/decompositions.py:177,181 (golub_kahan), :235-240 (golub_kahan_update), :212 (arnoldi_update)
 *   trips/solvers/CGLS.py:45-46,60,68; GKS.py:37,82,92; MMGKS.py:43-44,56,115,124
 * i.e. scipy sparsetools csr_matvec / csc_matvec behind scipy.sparse `__matmul__`.
 * The epilogue `- coef*z` is the three-term recurrence of decompositions.py:237,240 (z = NULL: plain product).
 * rowptr: int64[m+1]; colidx: int32[nnz] (16-byte aligned); vals: fp64 (or fp32 for _f32s; 32-byte aligned).
 * norm_out (nullable): 2 doubles; ws: tb200_spmv_workspace_len(m) doubles, required iff norm_out != NULL.
 * order 0: products of a row are added one after the other in index order with separately rounded multiply and
 *          add - the arithmetic of scipy's csr_matvec (and of csc_matvec when applied to the stored transpose):
 *          results are BIT-IDENTICAL to the reference's.  order 1: per-row tree reduction with FMA (fastest). */
int64_t tb200_spmv_workspace_len(int64_t m);
int tb200_spmv_launches(int with_norm);
int tb200_spmv_set_variant(int variant);
int tb200_spmv_csr_f64(int order, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx,
                       const double* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                       const double* z, double* norm_out, double* ws, void* stream);
/* fp32-storage / fp64-accumulate variant (reported separately from the fp64 parity build). */
int tb200_spmv_csr_f32s(int order, int64_t m, int64_t n, int64_t nnz, const int64_t* rowptr, const int32_t* colidx,
                        const float* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                        const double* z, double* norm_out, double* ws, void* stream);
/* Same product on the SELL-32-4 layout ("row-interleaved CSR"): rows in slices of 32, entry j of row r at
 * sliceptr[r/32] + (j/4)*128 + (r%32)*4 + j%4, rows of a slice zero-padded to the longest (rounded up to 4).
 * Every row is still summed in index order with separately rounded multiply/add: bit-identical to
 * tb200_spmv_csr_f64(order 0) and to scipy.  This is the production layout for the CT matrices: the 32 lanes of
 * a warp (lane = row) read contiguous 512 B / 1 KB per load.  sliceptr: int64[ceil(m/32)+1], rowlen: int32[m]. */
int tb200_spmv_sell_f64(int64_t m, int64_t n, const int64_t* sliceptr, const int32_t* rowlen, const int32_t* colidx,
                        const double* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                        const double* z, double* norm_out, double* ws, void* stream);
int tb200_spmv_sell_f32s(int64_t m, int64_t n, const int64_t* sliceptr, const int32_t* rowlen, const int32_t* colidx,
                         const float* vals, const double* x, double* y, double coef_host, const double* coef_dev,
                         const double* z, double* norm_out, double* ws, void* stream);
/* One complete Golub-Kahan step = the reference's golub_kahan_update (trips/utilities/decompositions.py:230-255) on
 * SELL-32-4 matrices A (m x n) and A^T (n x m): 6 kernels enqueued on `stream`, all scalars stay on the device.
 * v_prev / beta_prev_dev are NULL on the first step.  alpha_pair, beta_pair: 2 doubles each (sum of squares, norm).
 * events_host (nullable): four cudaEvent_t recorded around the A^T and the A launch (live per-launch timing). */
int tb200_gk_step_sell_f64(int64_t m, int64_t n, const int64_t* a_sliceptr, const int32_t* a_rowlen, const int32_t* a_col,
                           const double* a_val, const int64_t* at_sliceptr, const int32_t* at_rowlen,
                           const int32_t* at_col, const double* at_val, const double* u_k, const double* v_prev,
                           const double* beta_prev_dev, double* v_out, double* u_out, double* alpha_pair,
                           double* beta_pair, double* ws, void* const* events_host, void* stream);
int tb200_reduce_finalize(const double* partials, int64_t n, double* out, void* stream);

/* ---- BLAS-1 between operator applies -------------------------------------------------------------------
 * v/alpha, u/beta: decompositions.py:238-242 ; x+beta*p, r-beta*w, t+c*p, norms: CGLS.py:49-76 ;
 * smoothed Holder weights: weights.py:66-68, MMGKS.py:57,93 ; wf*(AVy-b), wr*(LVy): MMGKS.py:111-113 ;
 * la.norm(x - x_true): Hybrid_LSQR.py:110, GKS.py:101.  Element-wise ops round like NumPy (no FMA contraction). */
int64_t tb200_reduce_workspace_len(void);
int tb200_vec_div(int64_t n, const double* x, double d_host, const double* d_dev, double* out, void* stream);
int tb200_vec_axpy(int64_t n, double a_host, const double* a_dev, double sign, const double* x, const double* y,
                   double* out, double* norm_out, double* ws, void* stream); /* out = y + sign*(a*x), sign = +-1 */
/* The (hi, lo) partials of x.y (y == NULL: sum x^2) without the final reduction, for tb200_comm_allreduce_dd. */
int tb200_vec_dot_partials(int64_t n, const double* x, const double* y, double* ws, int64_t* n_partials_out, void* stream);
int tb200_vec_norm2(int64_t n, const double* x, double* out, double* ws, void* stream);
int tb200_vec_dot(int64_t n, const double* x, const double* y, double* out, double* ws, void* stream);
int tb200_vec_diffnorm2(int64_t n, const double* x, const double* y, double* out, double* ws, void* stream);
/* mode 0: out = x*y ; 1: out = x-y ; 2: out = w*(x-y) ; 3: out = x+y */
int tb200_vec_binary(int mode, int64_t n, const double* x, const double* y, const double* w, double* out, void* stream);
int tb200_irls_weights(int64_t n, const double* v, double eps, double expo, double* out, void* stream);

/* ---- tall-skinny basis kernels ---------------------------------------------------------------------------
 * h = V^T w and w -= V h: MMGKS.py:119-120 (CGS2), GKS.py:86-88 (x3), decompositions.py:216-218 (MGS columns);
 * lifts x = V y, AV y, LV y: Hybrid_LSQR.py:105, Hybrid_GMRES.py:77, GKS.py:76-83, MMGKS.py:107-116;
 * U^T b: reg_param/discrepancy_principle.py:34. */
int64_t tb200_basis_workspace_len(int64_t k);
int tb200_basis_dots(int64_t n, int64_t k, const double* V, int64_t ld, const double* w, double* h, double* ws,
                     void* stream);
/* out = w + sign * V h (w NULL => out = sign * V h); optional fused norm of out */
int tb200_basis_combine(int64_t n, int64_t k, const double* V, int64_t ld, const double* h, const double* w, double sign,
                        double* out, double* norm_out, double* ws, void* stream);
/* A/B switch: 1 (default) = two rows per 16-byte access where n, ld and the pointers allow it; same results. */
void tb200_basis_set_vec2(int on);

/* ---- weighted Gram + k x k factorisation (replaces the per-iteration host QRs) ----------------------------
 * la.qr(AV*wf), la.qr(LV*wr), Q_A.T@b: MMGKS.py:57-59,94-106 ; la.qr(AV), la.qr(LV): GKS.py:54-58.
 * Double-double accumulation; tb200_gram_factor_dd runs on the HOST (all pointers host). */
int64_t tb200_gram_workspace_len(int64_t K);
/* A/B switch: 1 (default) = tiles arrive by double-buffered cp.async.bulk, 0 = ordinary loads, single buffer. */
void tb200_gram_set_bulk(int on);
/* Block shape of the full pass: 0 (default) = chosen per K, 2 = 2 x 2 blocks (<= 512 threads), 4 = 4 x 4 (<= 256). */
void tb200_gram_set_block(int bs);
int tb200_weighted_gram(int64_t m, int64_t k, const double* B, int64_t ld, const double* w, int n_extra,
                        const double* const* extras_host, const int* extra_weighted_host, double* Ghi, double* Glo,
                        double* ws, void* stream);
/* Only the columns c >= c0 of G (and their mirror rows) are computed and written: for a basis whose first c0 columns have
 * not changed since their Gram matrix was formed (GKS.py:54-58 re-factors the whole of AV, LV every iteration). */
int tb200_weighted_gram_panel(int64_t m, int64_t k, const double* B, int64_t ld, const double* w, int n_extra,
                              const double* const* extras_host, const int* extra_weighted_host, int64_t c0, double* Ghi,
                              double* Glo, double* ws, void* stream);
int tb200_gram_factor_dd(int k, int ne, const double* Ghi_host, const double* Glo_host, double* R_host, double* C_host,
                         double* resid2_host);

/* ---- parallel-beam CT system matrix (A in CSR, A^T in CSR) -----------------------------------------------
 * Stands where ASTRA does in the reference (trips/test_problems/Tomography.py:49-88, utilities/io.py:392-400,
 * utilities/cil_io.py:271-275): theta = linspace(0, pi, views, endpoint=False), n_det = int(sqrt(2)*nx),
 * unit detector spacing, entry = chord length of the ray through the unit pixel.
 * Row of A = angle*n_det + det over the n_ang angles given by the cos/sin tables; column = iy*nx + ix.
 * Fill: sell = 0 writes CSR (ptr = rowptr), sell = 1 writes SELL-32-4 (ptr = slice pointers; arrays pre-zeroed). */
int tb200_ct_count_rows(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv, int32_t* counts,
                        void* stream);
int tb200_ct_fill_rows(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                       const int64_t* ptr, int sell, int32_t* colidx, double* vals, void* stream);
int tb200_ct_count_cols(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv, int32_t* counts,
                        void* stream);
int tb200_ct_fill_cols(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                       const int64_t* ptr, int sell, int32_t* colidx, double* vals, void* stream);

/* Flat-detector fan beam, the geometry the reference requests from ASTRA ('fanflat' projection geometry with the
 * 'line_fanflat' projector, trips/test_problems/Tomography.py:57-67): source at so*(sin, -cos), detector centre at
 * dd*(-sin, cos), detector axis (cos, sin), bins of width dps.  Same passes, layouts and entry function (chord of the
 * source->bin ray through the unit pixel) as the parallel-beam builder above. */
int tb200_ctfan_count_rows(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                           const double* sinv, int32_t* counts, void* stream);
int tb200_ctfan_fill_rows(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                          const double* sinv, const int64_t* ptr, int sell, int32_t* colidx, double* vals, void* stream);
int tb200_ctfan_count_cols(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                           const double* sinv, int32_t* counts, void* stream);
int tb200_ctfan_fill_cols(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                          const double* sinv, const int64_t* ptr, int sell, int32_t* colidx, double* vals, void* stream);

/* ---- parallel-beam CT projectors with the matrix values re-evaluated on the fly ---------------------------
 * The reference's tomography operator is matrix-free (astra.OpTomo behind pylops.FunctionOperator,
 * trips/test_problems/Tomography.py:73-83); SURVEY.md section 8(f) item 1.  Same matrix as the builder above, same
 * bits as the sequential-order SpMV on the stored matrices, but an entry costs ~10 fp64 instructions instead of
 * 12 streamed bytes: the forward projector reads only A's SELL-32-4 column indices (tb200_ct_fill_rows with
 * vals = NULL), the back-projector reads no matrix at all.
 * geom: 6 doubles per angle (c, s, d2, 1/hi, 1/(hi*lo), 0) written by tb200_ct_geometry, 16-byte aligned.
 * coef / z / norm_out / ws as for tb200_spmv_sell_f64; back-projector ws: tb200_ct_backproject_workspace_len.
 * cta_order (nullable): int32 permutation of the groups of four 32-row slices, heaviest first (schedules the long
 * central rays before the short peripheral ones; the result does not depend on it).
 * Row-aligned index layout (optional, rowskip nullable): tb200_ct_count_rows_first also reports the first image row each
 * ray crosses; the caller derives rowskip[r] (leading padding of row r inside its lane) so that the 32 rays of a slice
 * walk through the same image rows at the same positions - their x-gathers then share 32-byte sectors - and
 * tb200_ct_fill_rows_aligned writes the column indices at [rowskip[r], rowskip[r] + rowlen[r]).
 * Shallow rays (|sin| > |cos|) run along the image rows: neighbouring rays then sit in DIFFERENT rows and their gathers
 * share nothing.  With transpose_shallow != 0 their indices address the transposed image (ix*ny + iy), which
 * tb200_ct_forward_f64 forms in xT_scratch (nx*ny doubles) before the product: they then behave like steep rays.
 * first_run (nullable output of the count pass) is their alignment key: distance along x, from the side the ray comes
 * from, at which it enters its first row.  The order in which a row's entries are added never changes. */
int tb200_ct_geometry(int n_ang, const double* cosv, const double* sinv, double* geom, void* stream);
int tb200_ct_count_rows_first(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                              int32_t* counts, int32_t* first_row, int32_t* first_run, void* stream);
int tb200_ct_fill_rows_aligned(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                               const int64_t* sliceptr, const int32_t* rowskip, int transpose_shallow, int32_t* colidx,
                               void* stream);
int tb200_ct_forward_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* sliceptr,
                         const int32_t* rowlen, const int32_t* rowskip, const int32_t* colidx, const int32_t* cta_order,
                         double* xT_scratch, const double* x, double* y, double coef_host, const double* coef_dev,
                         const double* z, double* norm_out, double* ws, void* stream);
/* The STORED parallel-beam CT matrix over the same row-aligned index layout: tb200_ct_fill_rows_aligned_vals writes the
 * values next to the indices, tb200_ct_spmv_sell_f64 is the SELL-32-4 product y = A x - coef*z on it (vals double, or
 * float with vals_f32 != 0), tb200_gk_step_sell_ct_f64 one golub_kahan_update (decompositions.py:230-255) with A in this
 * layout and A^T plain SELL.  Same bits as tb200_spmv_sell_f64 / scipy on the plain layout; the x-gathers of the A launch
 * share sectors (the north-star's stored CSR SpMV, trips/utilities/decompositions.py:240). */
int tb200_ct_fill_rows_aligned_vals(int nx, int ny, int n_det, int n_ang, const double* cosv, const double* sinv,
                                    const int64_t* sliceptr, const int32_t* rowskip, int transpose_shallow, int32_t* colidx,
                                    double* vals, void* stream);
int tb200_ct_spmv_sell_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* sliceptr,
                           const int32_t* rowlen, const int32_t* rowskip, const int32_t* colidx, const void* vals,
                           int vals_f32, const int32_t* cta_order, double* xT_scratch, const double* x, double* y,
                           double coef_host, const double* coef_dev, const double* z, double* norm_out, double* ws,
                           void* stream);
int tb200_gk_step_sell_ct_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* a_sliceptr,
                              const int32_t* a_rowlen, const int32_t* a_rowskip, const int32_t* a_col, const double* a_val,
                              const int32_t* a_cta_order, double* xT_scratch, const int64_t* at_sliceptr,
                              const int32_t* at_rowlen, const int32_t* at_col, const double* at_val, const double* u_k,
                              const double* v_prev, const double* beta_prev_dev, double* v_out, double* u_out,
                              double* alpha_pair, double* beta_pair, double* ws, void* const* events_host, void* stream);
/* Ray-driven forward projection with NO index array (csrc/ct_forward.cu): the pixels of a ray are enumerated row by row
 * from the ray equation (candidate bracket + the builder's exact predicate), summed in ascending column index: same
 * bits as tb200_ct_forward_f64 / tb200_spmv_sell_f64 on the stored matrix and as scipy's A @ x.  Replaces `A @ v`
 * (trips/utilities/decompositions.py:240; astra.OpTomo's forward projection, trips/test_problems/Tomography.py:73-83).
 * ws: tb200_ct_forward_rays_workspace_len(n_det, n_ang) doubles iff norm_out != NULL. */
int64_t tb200_ct_forward_rays_workspace_len(int n_det, int n_ang);
int tb200_ct_forward_set_tuning(double run_tan, int min_ctas); /* tuning knobs; results never depend on them */
int tb200_ct_forward_rays_plan(int n_det, int n_ang, int* rays_per_cta, int* blocks_per_angle);
/* cta_order (nullable): n_ang * blocks_per_angle int32 (tb200_ct_forward_rays_plan), entry b = angle * blocks_per_angle +
 * block of the CTA to run b-th - heaviest first keeps the SMs evenly loaded to the end; the result does not depend on it. */
int tb200_ct_forward_rays_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const double* x, double* y,
                              double coef_host, const double* coef_dev, const double* z, double* norm_out, double* ws,
                              const int32_t* cta_order, void* stream);
/* The same ray-driven forward projection for the flat-detector fan beam of tb200_ctfan_* (per-ray geometry): replaces the
 * forward product of astra's 'line_fanflat' projector (trips/test_problems/Tomography.py:57-67, 73-83). */
int tb200_ctfan_forward_rays_f64(double so, double dd, double dps, int nx, int ny, int n_det, int n_ang, const double* cosv,
                                 const double* sinv, const double* x, double* y, double coef_host, const double* coef_dev,
                                 const double* z, double* norm_out, double* ws, void* stream);
int64_t tb200_ct_backproject_workspace_len(int nx, int ny);
int tb200_ct_backproject_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const double* u, double* y,
                             double coef_host, const double* coef_dev, const double* z, double* norm_out, double* ws,
                             void* stream);
/* Image rows [iy_begin, iy_end) only (y, z indexed by the global pixel number): a row-sharded caller back-projects the
 * image in bands and all-reduces each band while the next one is computed (trips-py_b200/dist.py). */
int tb200_ct_backproject_rows_f64(int nx, int ny, int iy_begin, int iy_end, int n_det, int n_ang, const double* geom,
                                  const double* u, double* y, double coef_host, const double* coef_dev, const double* z,
                                  double* norm_out, double* ws, void* stream);
/* Sharded forms of the two projectors (one process per GPU, trips-py_b200/dist.py; the reference has no parallelism:
 * SURVEY.md 2.2, 8e).  The sinogram is split by projection angle, the image by row band; each projector reads a WHOLE
 * vector (this rank's copy) and produces this rank's part of the other one, which its epilogue stores into every rank's
 * copy over NVLink peer memory (peers_host: n_peers <= 15 device pointers into the other ranks' arenas, see
 * tb200_comm_*), then fences at system scope.  Norm partials are left in `partials` for tb200_comm_allreduce_dd.
 * Back-projection: image rows [iy_begin, iy_end) from all n_ang angles; geom[6a+5] holds, as an int64 bit pattern, the
 * row of u at which global angle a starts (angles grouped by owner); z (nullable) = this rank's slice of the previous
 * basis vector, z[pix - iy_begin*nx].  Forward: this rank's n_ang angles (its own geom table) from the whole image.
 * Both sum in the single-GPU order: results are bit-identical to one GPU. */
int tb200_ct_backproject_sharded_f64(int nx, int ny, int iy_begin, int iy_end, int n_det, int n_ang, const double* geom,
                                     const double* u, double* y, double* const* peers_host, int n_peers, double coef_host,
                                     const double* coef_dev, const double* z, double* partials, int64_t* n_partials_out,
                                     void* stream);
int tb200_ct_forward_rays_sharded_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const double* x, double* y,
                                      double* const* peers_host, int n_peers, double coef_host, const double* coef_dev,
                                      const double* z, double* partials, int64_t* n_partials_out, const int32_t* cta_order,
                                      void* stream);
/* One Golub-Kahan step (trips/utilities/decompositions.py:230-255) on the matrix-free operator, as
 * tb200_gk_step_sell_f64; ws: max(tb200_spmv_workspace_len(m), tb200_ct_backproject_workspace_len(nx, ny),
 * tb200_ct_forward_rays_workspace_len(n_det, n_ang)).  colidx == NULL: fully matrix-free (ray-driven forward
 * projector; sliceptr .. xT_scratch are ignored and may be NULL). */
int tb200_gk_step_ct_f64(int nx, int ny, int n_det, int n_ang, const double* geom, const int64_t* sliceptr,
                         const int32_t* rowlen, const int32_t* rowskip, const int32_t* colidx, const int32_t* cta_order,
                         double* xT_scratch, const double* u_k, const double* v_prev, const double* beta_prev_dev,
                         double* v_out, double* u_out, double* alpha_pair, double* beta_pair, double* ws,
                         void* const* events_host, void* stream);

/* ---- stencils ---------------------------------------------------------------------------------------------
 * PSF blur and its reference "adjoint": trips/test_problems/Deblurring2D.py:66-73 (scipy.ndimage.convolve,
 * mode='reflect'); zero-padded variant for gen_data :121-133 (mode 1).  Bit-identical to ndimage (tap order and
 * rounding).  Pass W = flipped PSF for A, W = PSF for A^T; ch/cw = ph/2, pw/2 minus one for even sizes. */
int tb200_correlate2d_f64(int nrow, int ncol, const double* x, const double* W, int ph, int pw, int ch, int cw, int mode,
                          double* out, void* stream);
/* Forward-difference regularisation operators: trips/utilities/operators.py:24-28 (1-D), :30-36 (2-D, nt = 1),
 * :39-45 (space-time).  Fused IRLS weights wout = (u^2+eps^2)^expo: MMGKS.py:60,93; weighted adjoint
 * L^T (w . r): MMGKS.py:113-117.  x_next / rt_prev / wt_prev: one-frame halos for frame-sharded dynamic CT. */
int64_t tb200_fd_rows(int nt, int nrow, int ncol, int has_next);
int tb200_fd_apply(int nt, int nrow, int ncol, const double* x, const double* x_next, double* u, double* wout, double eps,
                   double expo, void* stream);
int tb200_fd_adjoint(int nt, int nrow, int ncol, int has_next, const double* r, const double* w, const double* rt_prev,
                     const double* wt_prev, double* out, void* stream);
/* Centred-difference gradient [I (x) D ; D (x) I], D = 3-point centred first derivative with zero end rows: the fp64
 * statement of first_derivative_operator_2d (trips/utilities/operators_old.py:35-45, float32 pylops there), with the
 * isotropic-TV weights (u1^2+u2^2+eps^2)^expo of MMGKS.py:64-78 fused into the apply pass. */
int tb200_cd2d_apply(int nrow, int ncol, const double* x, double* u, double* wout, double eps, double expo, void* stream);
int tb200_cd2d_adjoint(int nrow, int ncol, const double* r, const double* w, double* out, void* stream);
int tb200_fd1d_apply(int64_t n, const double* x, double* u, void* stream);
int tb200_fd1d_adjoint(int64_t n, const double* r, double* out, void* stream);

/* ---- intra-node communication over NVLink peer memory (csrc/comm.cu) -----------------------
 * No reference counterpart (the reference is single-process).  Every rank allocates one arena and publishes its CUDA IPC
 * handle; tb200_comm_connect maps all peers' arenas.  The first tb200_comm_mailbox_bytes() bytes of an arena are the
 * mailboxes of tb200_comm_allreduce_dd; the caller lays out the rest.
 *   tb200_comm_init       arena_bytes of zero-filled device memory + its handle (tb200_comm_handle_bytes() bytes)
 *   tb200_comm_connect    all_handles = nranks handles in rank order (exchanged by the caller out of band)
 *   tb200_comm_arena      device address, valid in this process, of rank `peer`'s arena
 *   tb200_comm_allreduce_dd  sum over the ranks of nval (<= 64) double-double values, each given as npart local (hi, lo)
 *                         partials; out[2j] = total, out[2j+1] = sqrt(total); summed in rank order => bit-identical on
 *                         every rank and equal to the single-GPU value; also the barrier that makes earlier pushes /
 *                         projector peer stores visible.  box in [0, 8), epoch strictly increasing per box.
 *   tb200_comm_push       src[0:n) -> offset_bytes of the arena of every rank in rank_mask
 *   tb200_halo_exchange   one-frame halos with rank-1 / rank+1 for the temporal difference operator of frame-sharded
 *                         dynamic CT (replaces the halo of trips-py_b200/dist.py FrameComm; operators.py:39-45)
 *   tb200_comm_scale      out = x / d over a gathered vector, own slice copied into the basis column in the same pass
 *                         (v / alpha, u / beta of decompositions.py:239,242) */
int64_t tb200_comm_mailbox_bytes(void);
int tb200_comm_handle_bytes(void);
int tb200_comm_max_ranks(void);
int tb200_comm_init(int rank, int nranks, int64_t arena_bytes, void** comm_out, unsigned char* handle_out);
int tb200_comm_connect(void* comm, const unsigned char* all_handles);
void* tb200_comm_arena(void* comm, int peer);
int tb200_comm_destroy(void* comm);
int tb200_comm_allreduce_dd(void* comm, int box, int64_t epoch, const double* partials, int64_t npart, int nval, double* out,
                            void* stream);
int tb200_comm_allreduce_scale(void* comm, int box, int64_t epoch, const double* partials, int64_t npart, int64_t n,
                               const double* x, double* out, int64_t keep_begin, int64_t keep_n, double* keep, double* pair_out,
                               void* stream); /* allreduce_dd of one value + scale by its square root, one launch */
int tb200_comm_push(void* comm, unsigned rank_mask, int64_t offset_bytes, const double* src, int64_t n, void* stream);
int tb200_halo_exchange(void* comm, int box, int64_t epoch, int64_t offset_bytes, const double* send_prev, const double* send_next,
                        int64_t n, double* recv_prev, double* recv_next, double* scratch_pair, void* stream);
int tb200_comm_scale(int64_t n, const double* x, const double* d_dev, double* out, int64_t keep_begin, int64_t keep_n,
                     double* keep, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TRIPSB200_H */
