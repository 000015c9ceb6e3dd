/*
 * wifi_b200.h -- C-ABI of the B200-native 802.11 channel-estimation hot path.
 *
 * Drop-in boundary for the hot path of usmandroid/80211ParallelEstimation: the five
 * estimators of main.c (prototypes main.c:4-8, bodies main.c:66-212), the complex matrix
 * routines of utils.c (prototypes utils.h:38-60) and the MATLAB equalizer
 * (WiFi_Equalization.m:1-9).  Plain pointers and sizes only; every batched entry point
 * returns a wifi_status.  There is NO CPU fallback: without a CUDA device wifi_create()
 * fails with WIFI_ERR_NO_DEVICE and nothing else can be called.
 *
 * Layouts (kept from the reference, inputs.h:20,75,130,928 and utils.h:10-19):
 *   - a complex element is interleaved (re, im): float2 (WIFI_F32) or double2 (WIFI_F64);
 *   - a block vector is 53 consecutive elements, sub-carrier k at [k]; DC bin = 26,
 *     pilots at 5, 19, 33, 47;
 *   - a whole frame is 15 block vectors: element [53*b + k] (inputs.h tx_symb/rx_symb);
 *   - a batch is n_frames consecutive block vectors / frames; `frame_stride` (in complex
 *     elements) lets the pilot estimators read block b of whole frames in place
 *     (pass ptr + 53*b and frame_stride = 795) or stacked block vectors (stride 53);
 *   - matrices are dense row-major (what the reference's row-pointer tables point into,
 *     utils.c:817-835), batches are consecutive matrices.
 *
 * Pointers named *_dev / unqualified in the `_batch` functions are DEVICE pointers on the
 * context's GPU; the `_host` functions take HOST pointers (pinned or pageable) and run the
 * H2D copy, the kernels and the D2H copy themselves.
 */
#ifndef WIFI_B200_H
#define WIFI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WIFI_NSC 53        /* SAMPUTIL, utils.h:13 */
#define WIFI_NBLK 15       /* OFDMBLK,  utils.h:15 */
#define WIFI_FRAME (WIFI_NSC * WIFI_NBLK) /* SIZESYMBOL, utils.h:12 */
#define WIFI_DC 26
#define WIFI_PACKET 1200   /* time samples per frame: 15 x (16 CP + 64), WiFi_RX.m:11-14 */
#define WIFI_LPTOT 160     /* long-training field samples (32 GI + 2 x 64) */
#define WIFI_P0 5          /* utils.h:16-19 */
#define WIFI_P1 19
#define WIFI_P2 33
#define WIFI_P3 47
#define WIFI_MAX_ORDER 64  /* largest matrix order of the batched utils */

typedef enum { WIFI_F32 = 0, WIFI_F64 = 1 } wifi_dtype;

typedef enum {
    WIFI_OK = 0,
    WIFI_ERR_INVALID = 1,   /* bad argument / "Matrices dimension missmatch" (utils.c:18-19): nothing is written */
    WIFI_ERR_CUDA = 2,      /* a CUDA call failed; see wifi_last_error() */
    WIFI_ERR_NOMEM = 3,
    WIFI_ERR_SINGULAR = 4,  /* a pivot was exactly zero (the reference silently yields NaN/Inf, utils.c:543-569) */
    WIFI_ERR_NO_DEVICE = 5,
    WIFI_ERR_STATE = 6      /* e.g. shared-filter apply before wifi_mmse_filter_* */
} wifi_status;

/* which-estimator bit mask for wifi_ps_batch */
#define WIFI_PS_LINEAR 1
#define WIFI_PS_CUBIC 2
#define WIFI_PS_SINC 4
/* OR-ed in: MATLAB semantics (WiFi_channel_estimation_PS_*.m) instead of main.c's -- the estimate is averaged over OFDM
 * blocks 1..4 of a whole frame (frame_stride >= 212) and Cubic uses the true divided-difference spans 14/28/42 */
#define WIFI_PS_MATLAB 8

/* flags of wifi_mmse_perframe_batch.  Every default mode meets the accuracy bound of its storage type (1e-10 for WIFI_F64,
 * 1e-4 for WIFI_F32, relative per sub-carrier): with WIFI_F32 the arrays are FP32 and the solve runs in FP64 arithmetic --
 * sigma2/|x|^2 (1e-10..1e-7) lies below the FP32 resolution of R (1e-4), so an FP32 elimination of R + D cannot. */
#define WIFI_SOLVE_PIVOT 0      /* LU with partial pivoting + back-substitution, H = R z (any non-singular R + D) */
#define WIFI_SOLVE_HPD 1        /* R Hermitian PSD: un-pivoted blocked L D L^H with the trailing updates on the FP64 tensor path
                                 * (DMMA), H = y - D z (growth factor 1) */
#define WIFI_SOLVE_WIDE 2       /* accepted and ignored: FP32 storage + FP64 arithmetic is what WIFI_F32 does by default */
#define WIFI_SOLVE_REFINE WIFI_SOLVE_WIDE   /* former name */
#define WIFI_SOLVE_FAST32 4     /* explicit opt-in, WIFI_F32 only: FP32 ARITHMETIC inside the solve (register-resident L D L^H on the
                                 * FP32 cores, 2x the throughput).  Documented accuracy: ~4e-3 with WIFI_SOLVE_HPD, ~1e-1 with
                                 * WIFI_SOLVE_PIVOT at sigma2 = 1e-8 -- it does NOT meet the 1e-4 bound.  For frames that share
                                 * |tx_k|^2 prefer wifi_mmse_eig_* (FP32 arithmetic, 1e-5, 20x faster). */

/* wifi_chermitian_batch / wifi_cadd_batch semantics */
#define WIFI_AS_WRITTEN 0       /* bit-compatible with utils.c:3-7 (Re-Im, real-valued) / utils.c:111-121 (M1+M1) */
#define WIFI_INTENDED 1         /* true conjugate transpose / M1+M2 */

typedef struct wifi_ctx wifi_ctx;   /* opaque; one per GPU; calls on one ctx must not race */

/* ---- context ------------------------------------------------------------------- */
int wifi_create(int device, wifi_ctx **out);
int wifi_destroy(wifi_ctx *ctx);
int wifi_set_stream(wifi_ctx *ctx, void *cuda_stream);      /* cudaStream_t; NULL = default stream */
int wifi_synchronize(wifi_ctx *ctx);
const char *wifi_last_error(wifi_ctx *ctx);
const char *wifi_version(void);
/* kernels launched through this ctx since creation (bench.py's gpu_launches) */
int64_t wifi_launch_count(wifi_ctx *ctx);
/* elapsed device time in ms of the LAST kernel launched with timing enabled (CUDA events on the ctx stream) */
int wifi_enable_kernel_timing(wifi_ctx *ctx, int on);
int wifi_last_kernel_ms(wifi_ctx *ctx, float *ms);

/* ---- LS / interpolation estimators (HBM-bound) ---------------------------------- */
/* main.c:66-75  WiFi_channel_estimation_LT_LS over n_frames preambles [n][53] -> H [n][53] */
int wifi_lt_ls_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, void *H, int64_t n_frames);
/* main.c:77-146 PS_Linear / PS_Cubic / PS_Sinc fused: one pilot-LS pass feeds every requested
 * interpolator.  Outputs [n][53]; pointers of estimators not in `which` are ignored. */
int wifi_ps_batch(wifi_ctx *ctx, wifi_dtype dt, int which, const void *tx_symbols, const void *rx_symbols,
                  int64_t frame_stride, void *H_linear, void *H_cubic, void *H_sinc, int64_t n_frames);
/* WiFi_Equalization.m:1-9: rx [n][15][53], H_lt/H_ps [n][53] -> eq [n][15][53] */
int wifi_equalize_batch(wifi_ctx *ctx, wifi_dtype dt, const void *rx_frames, const void *H_lt, const void *H_ps,
                        void *eq, int64_t n_frames);

/* BASELINE configs[4] in one call: all five estimators of main.c:4-8 (+ the equalizer when eq != NULL) for n frames.
 * tx_pre/rx_pre [n][53]; tx_symbols/rx_symbols: block 0 at frame_stride (53: block vectors, 795: whole frames in place);
 * PS_MMSE is the shared-filter form (wifi_mmse_filter_form/_set first); eq [n][15][53] needs whole frames (frame_stride 795)
 * and blends H_lt with H_linear like WiFi_RX.m:60.  Four launches (LT_LS, pilot-LS + interpolators, PS_MMSE GEMM, equalizer). */
int wifi_estimate_all_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, const void *tx_symbols,
                            const void *rx_symbols, int64_t frame_stride, void *H_lt, void *H_linear, void *H_cubic, void *H_sinc,
                            void *H_mmse, void *eq, int64_t n_frames);

/* ---- receiver front-end: time samples -> the estimators' inputs ------------------------------------
 * WiFi_blocks_extraction.m:5-10 and WiFi_RX.m:19-31, for one side (tx or rx) of n frames:
 *   packet [n][1200] = 15 OFDM blocks of 16 cyclic-prefix + 64 samples;  lptot [n][160] = long-training field
 *   symb [n][15][53] = keep53(circshift(fft64(block without CP), 26));   pre_fft [n][53] = the same of (p1 + p2)/2,
 *   p1 = lptot[96..159], p2 = lptot[32..95];   ow2 [n] (real, optional) = sum |p2 - p1|^2 / 128.
 * packet and lptot must be 16-byte aligned. */
int wifi_frontend_batch(wifi_ctx *ctx, wifi_dtype dt, const void *packet, const void *lptot, void *symb, void *pre_fft,
                        void *ow2, int64_t n_frames);

/* ---- fused receiver chain: time samples -> every estimate (+ equalized symbols) in ONE launch ----
 * WiFi_RX.m:17-60 with the C estimators of main.c:66-146 (OFDM block 0, C cubic): the front-end's symbols never round-trip
 * through HBM.  tx_packet / rx_packet [n][1200], tx_lptot / rx_lptot [n][160], 16-byte aligned; of tx_packet only OFDM block 0
 * (samples 16..79 of every frame) is read.  Every output may be NULL:
 *   H_lt, H_linear, H_cubic, H_sinc [n][53]   LT_LS and the three pilot interpolators
 *   H_mmse_cconv [n][53]   PS_MMSE in the main.c:148 calling convention (R_f = H_lt H_lt^H, ow2 = the frame's noise estimate)
 *   H_ls0 [n][53]          rx/tx of OFDM block 0 (the LS input of the shared-filter / per-frame PS_MMSE)
 *   H_mmse_shared [n][53]  the installed shared filter (wifi_mmse_filter_form/_set) applied to H_ls0 (a second launch)
 *   eq [n][15][53]         WiFi_Equalization.m with H_lt and H_linear;  rx_symb [n][15][53] the rx symbols themselves
 *   ow2 [n] real           WiFi_RX.m:31 */
typedef struct {
    void *H_lt, *H_linear, *H_cubic, *H_sinc, *H_mmse_cconv, *H_ls0, *H_mmse_shared, *eq, *rx_symb, *ow2;
} wifi_rx_chain_out;
int wifi_rx_chain_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_packet, const void *tx_lptot, const void *rx_packet,
                        const void *rx_lptot, const wifi_rx_chain_out *out, int64_t n_frames);

/* ---- PS_MMSE, intended formula  H = R (R + s2 (X X^H)^-1)^-1 (rx/tx) ------------- */
/* Shared-filter case.  Form W = R (R + diag(d))^-1 once in FP64 on the device
 * (R: 53x53 double2 row-major, d: 53 doubles = s2/|x_k|^2, W_out: optional 53x53 double2)
 * and install it in the context as the operand of wifi_mmse_shared_*. */
int wifi_mmse_filter_form(wifi_ctx *ctx, const void *R_f64, const double *d_f64, void *W_out_f64);
/* install an externally formed filter (53x53 double2, device) */
int wifi_mmse_filter_set(wifi_ctx *ctx, const void *W_f64);
/* H[n][53] = H_ls[n][53] W^T   (multiply utils.c:16-31 over all frames as one GEMM) */
int wifi_mmse_shared_apply_batch(wifi_ctx *ctx, wifi_dtype dt, const void *H_ls, void *H, int64_t n_frames);
/* fused: per-block LS divide rx/tx (main.c:83 arithmetic on all 53 bins) + the GEMM */
int wifi_mmse_shared_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_symbols, const void *rx_symbols,
                           int64_t frame_stride, void *H, int64_t n_frames);
/* Shared, known tx block vector (training symbols; 53 double2, device): fold the LS divide into the installed filter once,
 * W' = W diag(1/tx), then H[n][53] = rx[n][53] W'^T reads only rx -- 848 instead of 1 272 bytes per frame in FP32.  The
 * filter of wifi_mmse_filter_form/_set stays installed for the per-frame-tx calls. */
int wifi_mmse_filter_fold_tx(wifi_ctx *ctx, const void *tx_block_f64);
int wifi_mmse_shared_rx_batch(wifi_ctx *ctx, wifi_dtype dt, const void *rx_symbols, int64_t frame_stride, void *H, int64_t n_frames);
/* Per-frame case: A_f = R + diag(sigma2[f]/|tx_k|^2); solve A_f z = rx/tx; H = R z (= rx/tx - D_f z).
 * R in the storage dtype (53x53), sigma2 real [n] in the storage dtype; flags: WIFI_SOLVE_*. */
int wifi_mmse_perframe_batch(wifi_ctx *ctx, wifi_dtype dt, const void *R, const void *tx_symbols,
                             const void *rx_symbols, int64_t frame_stride, const void *sigma2, void *H,
                             int64_t n_frames, int flags);
/* Per-frame case in the eigen domain, for frames that share their modulus pattern |tx_k|^2 (any constant-modulus
 * constellation; signs/phases are free): with S = |x| R |x| = V L V^H computed once (FP64 Jacobi on the device),
 * H = y - G2 (s_f (.) (G y)), s_fi = sigma2_f/(l_i + sigma2_f) -- two shared-matrix products on the tensor cores instead of
 * a 53x53 solve per frame, and accurate to ~1e-5 in FP32.  At most one null bin (|x_k|^2 < 1e-6 max, e.g. DC) is carried
 * exactly as a border.  R: 53x53 double2 (device), absx2: 53 doubles (device).  The frames passed to the apply call MUST
 * have |tx_k|^2 == absx2[k]; this is not checked. */
int wifi_mmse_eig_prepare(wifi_ctx *ctx, const void *R_f64, const double *absx2_f64);
int wifi_mmse_perframe_eig_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_symbols, const void *rx_symbols,
                                 int64_t frame_stride, const void *sigma2, void *H, int64_t n_frames);
/* Per-frame case for a LOW-RANK covariance (a channel of L taps: rank L; numerical rank <= WIFI_LOWRANK_MAX), sigma2 AND the
 * modulus pattern |tx_k|^2 free per frame (QAM): with R = U L U^H (FP64 Jacobi on the device, eigenvalues below 64 eps l_max
 * dropped)  H = U (sigma2 L^-1 + U^H diag(|x|^2) U)^-1 U^H (conj(x) (.) rx)  -- the same H = R (R + sigma2 diag(1/|x|^2))^-1 (rx/tx)
 * as wifi_mmse_perframe_batch (WiFi_channel_estimation_PS_MMSE.m:16-33) by the push-through identity, as an r x r solve per frame
 * in ONE HBM-bound launch.  The r x r system is well conditioned, so WIFI_F32 runs in FP32 arithmetic and meets the 1e-4 bound.
 * A bin with tx = 0 contributes nothing (the 53 x 53 solve yields NaN there).  R: 53x53 double2 (device); *rank_out (may be NULL)
 * receives the numerical rank; rank 0 or > WIFI_LOWRANK_MAX -> WIFI_ERR_INVALID and no operands are installed. */
#define WIFI_LOWRANK_MAX 8
int wifi_mmse_lowrank_prepare(wifi_ctx *ctx, const void *R_f64, int *rank_out);
int wifi_mmse_perframe_lowrank_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_symbols, const void *rx_symbols,
                                     int64_t frame_stride, const void *sigma2, void *H, int64_t n_frames);
/* C calling convention of main.c:148 batched: R_f = H_ls,f H_ls,f^H (main.c:186-189 intent),
 * tx/rx block vectors [n][53], ow2 [n] real, H_ls [n][53] -> H [n][53] */
int wifi_mmse_cconv_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_symbols, const void *rx_symbols,
                          const void *ow2, const void *H_ls, void *H, int64_t n_frames);
/* WiFi_channel_estimation_PS_MMSE.m as written (Rhh = ifft(H_ls) ifft(H_ls)', X unconjugated in Rhy), averaged over OFDM
 * blocks 1..4: tx/rx whole frames [n][15][53], ow2 [n] real, H_ls [n][53] -> H [n][53].  Both calls use the closed form of
 * the rank-one covariance, H = H_ls (v^H rx)/(ow2 + v^H v), v = tx (.) H_ls (212 complex values of traffic per frame). */
int wifi_mmse_matlab_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_frames, const void *rx_frames,
                           const void *ow2, const void *H_ls, void *H, int64_t n_frames);

/* ---- batched complex matrix utils (utils.h:38-60), order <= WIFI_MAX_ORDER --------- */
/* multiply utils.c:16-31: C[b] = A[b] (r1 x c1) * B[b] (r2 x c2); c1 != r2 -> WIFI_ERR_INVALID, nothing written */
int wifi_cmatmul_batch(wifi_ctx *ctx, wifi_dtype dt, const void *A, int r1, int c1, const void *B, int r2, int c2,
                       void *C, int64_t batch);
/* hermitian utils.c:3-7 (mode WIFI_AS_WRITTEN: res[c][r] = Re - Im) or conjugate transpose (WIFI_INTENDED) */
int wifi_chermitian_batch(wifi_ctx *ctx, wifi_dtype dt, int mode, const void *M, int row, int col, void *res, int64_t batch);
/* addition utils.c:111-121 (WIFI_AS_WRITTEN: M1+M1) or M1+M2 (WIFI_INTENDED) */
int wifi_cadd_batch(wifi_ctx *ctx, wifi_dtype dt, int mode, const void *M1, int r1, int c1, const void *M2, int r2, int c2,
                    void *res, int64_t batch);
/* multiplyVxVeqM utils.c:55-65: res[r][c] = M1[r][0] * M2[0][c] */
int wifi_couter_batch(wifi_ctx *ctx, wifi_dtype dt, const void *M1, int r1, int c1, const void *M2, int r2, int c2,
                      void *res, int64_t batch);
/* identity utils.c:84-93 */
int wifi_cidentity_batch(wifi_ctx *ctx, wifi_dtype dt, void *Id, int size, double scalar, int64_t batch);
/* inverse utils.c:141-170 replaced by elimination with partial pivoting (in-place Gauss-Jordan in registers for orders 33..64,
   LU + back-substitution in shared memory below); info[b] (device int, may be NULL) = 1 if singular */
int wifi_cinverse_batch(wifi_ctx *ctx, wifi_dtype dt, const void *A, int order, void *Y, int64_t batch, int *info);

/* ---- synthetic frames of the inputs.h shape, generated on the device ---- */
/* per_frame_sigma: 0 -> sigma2 = 9.6172e-08 for every frame, 1 -> log-uniform [1e-8, 1e-5].
 * Any output pointer may be NULL.  tx_pre/rx_pre/H_true [n][53], tx_symb/rx_symb [n][15][53], sigma2 [n] real. */
int wifi_synth_frames(wifi_ctx *ctx, wifi_dtype dt, uint64_t seed, int64_t first_frame, int64_t n_frames, int per_frame_sigma,
                      void *tx_pre, void *rx_pre, void *tx_symb, void *rx_symb, void *H_true, void *sigma2);
/* theoretical channel covariance of the generator: 53x53 double2 (device) */
int wifi_synth_covariance(wifi_ctx *ctx, void *R_f64);
/* per-shard error statistics, stats[4] (device doubles): sum|H-Href|^2, sum|Href|^2, count, max|H-Href| */
int wifi_error_stats(wifi_ctx *ctx, wifi_dtype dt, const void *H, const void *H_ref, int64_t n_elems, double *stats);

/* measured ceilings of this GPU, for the roofline fractions bench.py reports (blocking; a few ms each):
 * which = 0 FP32 FMA TFLOP/s, 1 FP64 FMA TFLOP/s, 2 FP64 DMMA (mma.sync m8n8k4) TFLOP/s, 3 streaming copy GB/s */
int wifi_measure_peak(wifi_ctx *ctx, int which, double *value);

/* ---- host-pointer variants: H2D + kernels + D2H inside, chunked and double-buffered ---- */
int wifi_lt_ls_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, void *H, int64_t n_frames);
int wifi_ps_host(wifi_ctx *ctx, wifi_dtype dt, int which, const void *tx_symbols, const void *rx_symbols,
                 int64_t frame_stride, void *H_linear, void *H_cubic, void *H_sinc, int64_t n_frames);
int wifi_equalize_host(wifi_ctx *ctx, wifi_dtype dt, const void *rx_frames, const void *H_lt, const void *H_ps,
                       void *eq, int64_t n_frames);
/* whole frames [n][15][53] on the host; eq may be NULL (then only block 0 of rx crosses the bus) */
int wifi_estimate_all_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, const void *tx_frames, const void *rx_frames,
                           void *H_lt, void *H_linear, void *H_cubic, void *H_sinc, void *H_mmse, void *eq, int64_t n_frames);
int wifi_frontend_host(wifi_ctx *ctx, wifi_dtype dt, const void *packet, const void *lptot, void *symb, void *pre_fft,
                       void *ow2, int64_t n_frames);
/* host arrays: of tx_packet only the 80 samples of OFDM block 0 cross the bus */
int wifi_rx_chain_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_packet, const void *tx_lptot, const void *rx_packet,
                       const void *rx_lptot, const wifi_rx_chain_out *out, int64_t n_frames);
int wifi_mmse_filter_form_host(wifi_ctx *ctx, const void *R_f64, const double *d_f64, void *W_out_f64);
int wifi_mmse_shared_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_symbols, const void *rx_symbols,
                          int64_t frame_stride, void *H, int64_t n_frames);
int wifi_mmse_filter_fold_tx_host(wifi_ctx *ctx, const void *tx_block_f64);
int wifi_mmse_shared_rx_host(wifi_ctx *ctx, wifi_dtype dt, const void *rx_symbols, int64_t frame_stride, void *H, int64_t n_frames);
int wifi_mmse_perframe_host(wifi_ctx *ctx, wifi_dtype dt, const void *R, const void *tx_symbols, const void *rx_symbols,
                            int64_t frame_stride, const void *sigma2, void *H, int64_t n_frames, int flags);
int wifi_mmse_perframe_eig_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_symbols, const void *rx_symbols,
                                int64_t frame_stride, const void *sigma2, void *H, int64_t n_frames);
int wifi_mmse_perframe_lowrank_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_symbols, const void *rx_symbols,
                                    int64_t frame_stride, const void *sigma2, void *H, int64_t n_frames);
int wifi_mmse_cconv_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_symbols, const void *rx_symbols,
                         const void *ow2, const void *H_ls, void *H, int64_t n_frames);
int wifi_mmse_matlab_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_frames, const void *rx_frames,
                          const void *ow2, const void *H_ls, void *H, int64_t n_frames);
int wifi_cmatmul_host(wifi_ctx *ctx, wifi_dtype dt, const void *A, int r1, int c1, const void *B, int r2, int c2, void *C, int64_t batch);
int wifi_chermitian_host(wifi_ctx *ctx, wifi_dtype dt, int mode, const void *M, int row, int col, void *res, int64_t batch);
int wifi_cadd_host(wifi_ctx *ctx, wifi_dtype dt, int mode, const void *M1, int r1, int c1, const void *M2, int r2, int c2, void *res, int64_t batch);
int wifi_couter_host(wifi_ctx *ctx, wifi_dtype dt, const void *M1, int r1, int c1, const void *M2, int r2, int c2, void *res, int64_t batch);
int wifi_cidentity_host(wifi_ctx *ctx, wifi_dtype dt, void *Id, int size, double scalar, int64_t batch);
int wifi_cinverse_host(wifi_ctx *ctx, wifi_dtype dt, const void *A, int order, void *Y, int64_t batch, int *info_host);
/* per-array staging target of one chunk of the `_host` pipeline (default 48 MiB; two chunks are in flight) */
int wifi_set_host_chunk_bytes(wifi_ctx *ctx, size_t bytes);
/* PCIe ceiling under the conditions of the `_host` pipeline: ONE pinned H2D copy of h2d_bytes and ONE D2H copy of d2h_bytes
 * issued together on the pipeline's two streams; *ms = wall time until both are done (either size may be 0) */
int wifi_pcie_probe(wifi_ctx *ctx, const void *h_src, void *h_dst, size_t h2d_bytes, size_t d2h_bytes, double *ms);
/* pinned host memory for the `_host` calls (cudaHostAlloc / cudaFreeHost) */
int wifi_host_alloc(void **p, size_t bytes);
int wifi_host_free(void *p);

/* ---- process-global default context used by the single-frame drop-ins (wifi_dropin.h) ---- */
wifi_ctx *wifi_default_ctx(void);   /* created on first use on device $WIFI_B200_DEVICE (default 0); aborts loudly if no GPU */

#ifdef __cplusplus
}
#endif
#endif /* WIFI_B200_H */
