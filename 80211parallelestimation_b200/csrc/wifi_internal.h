// wifi_internal.h -- launcher prototypes shared by the .cu translation units and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>
#include "../../include/wifi_b200.h"

#define WIFI_DMMA_BS 116   /* row stride of the FP64 filter image in doubles (4 mod 16: conflict-free fragment loads) */

namespace wifi {

// interpolation weight tables: H_k = sum_i w[est][k][i] * Hp_i  (est: 0 linear, 1 cubic, 2 sinc)
struct InterpTables {
    float *w32;   // [4][53][4]  (table 3: MATLAB cubic, true divided differences)
    double *w64;  // [4][53][4]
};

// shared MMSE filter operand images (built by filter_install from W, 53x53 double2 row-major)
struct FilterImages {
    double *W64;     // [53][53] double2, row-major (the FP64 truth)
    float *Bhi;      // tf32 "hi" image of the real embedding, UMMA canonical K-major no-swizzle layout [112 x 112]
    float *Blo;      // tf32 "lo" image (W - hi)
    double *B64;     // real embedding for the FP64 DMMA path, Bt[112][WIFI_DMMA_BS] (n-major, k contiguous, rows padded)
    int valid;
};

cudaError_t launch_lt_ls(wifi_dtype dt, const void *tx, const void *rx, void *H, int64_t n_frames, cudaStream_t s);
// hp_in != NULL: the four pilot LS values of every frame, [n][4], as written by the PS_MMSE GEMM kernels (hp_out) -- tx / rx are not read
cudaError_t launch_ps(wifi_dtype dt, int which, const void *tx, const void *rx, int64_t frame_stride, void *Hl, void *Hc,
                      void *Hs, int64_t n_frames, const InterpTables &tab, cudaStream_t s, const void *hp_in = nullptr);
cudaError_t launch_equalize(wifi_dtype dt, const void *rx, const void *Hlt, const void *Hps, void *eq, int64_t n_frames,
                            cudaStream_t s);

// PS_MMSE with R_f = H_ls H_ls^H (main.c:148 convention) in closed form; matlab = 1: the .m text, averaged over blocks 0..3
cudaError_t launch_mmse_rank1(wifi_dtype dt, int matlab, const void *tx, const void *rx, int64_t frame_stride, const void *ow2, const void *Hls,
                              void *H, int64_t n_frames, cudaStream_t s);

// receiver front-end (wifi_frontend.cu): packet [n][1200], lptot [n][160] -> symb [n][15][53], pre_fft [n][53], ow2 [n] (may be NULL)
cudaError_t launch_frontend(wifi_dtype dt, const void *packet, const void *lptot, void *symb, void *pre_fft, void *ow2,
                            int64_t n_frames, cudaStream_t s);

// fused receiver chain (wifi_frontend.cu): time samples -> estimates (+ equalized symbols) in one launch; every output may be NULL.
// tx_packet is read at tx_pkt_stride complex values per frame (only its OFDM block 0 -- samples 16..79 -- is used)
cudaError_t launch_rx_chain(wifi_dtype dt, const void *tx_packet, int64_t tx_pkt_stride, const void *tx_lptot, const void *rx_packet,
                            const void *rx_lptot, void *H_lt, void *H_lin, void *H_cub, void *H_sinc, void *H_cconv, void *H_ls0, void *eq,
                            void *rx_symb, void *ow2, int64_t n_frames, const InterpTables &tab, cudaStream_t s);

// dense solves (wifi_solve.cu)
cudaError_t launch_filter_form(const void *R64, const double *d64, void *W64, int *info, cudaStream_t s);
cudaError_t launch_cinverse(wifi_dtype dt, const void *A, int order, void *Y, int64_t batch, int *info, cudaStream_t s);
cudaError_t launch_cinverse_tc(wifi_dtype dt, const void *A, int order, void *Y, int64_t batch, int *info, cudaStream_t s);   // orders 33..64 (wifi_inverse_tc.cu)
cudaError_t launch_mmse_perframe_pivot(wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                                       const void *sigma2, const void *Hls_for_R, void *H, int64_t n_frames, int *info, int fast32,
                                       cudaStream_t s);
cudaError_t launch_mmse_pivot_tc(wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, void *H,
                                 int64_t n_frames, int *info, cudaStream_t s);   // shared R, FP64 arithmetic: warp-pair LU, DMMA updates (wifi_inverse_tc.cu)
cudaError_t launch_mmse_perframe_hpd(wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                                     const void *sigma2, void *H, int64_t n_frames, int fast32, cudaStream_t s);

// shared-filter GEMM (wifi_gemm_tc.cu, wifi_gemm_dmma.cu) + small matrix utils (wifi_gemm.cu)
cudaError_t launch_filter_fold(const void *W64, const void *tx64, void *Wout64, cudaStream_t s);   // W diag(1/tx)
cudaError_t launch_filter_install_tc(FilterImages &img, cudaStream_t s);     // W64 -> Bhi/Blo (UMMA canonical layout)
cudaError_t launch_mmse_shared_tc(const FilterImages &img, const void *tx_or_hls, const void *rx, int64_t frame_stride, void *H,
                                  int64_t n_frames, cudaStream_t s, void *hp_out = nullptr);   // FP32 I/O, 3xTF32 on tcgen05; hp_out [n][4]: pilot LS
cudaError_t launch_mmse_shared_tc_eig_h(const FilterImages &img, const void *u, const void *tx, const void *rx, int64_t frame_stride, int dc,
                                        const void *sigma2, const double *lam, const void *p, double Rdd, double md, void *H,
                                        int64_t n_frames, cudaStream_t s);   // eigen-domain MMSE, 2nd product: H = rx/tx - (s (.) (u - p z_d)) G2^T
cudaError_t launch_filter_install_dmma(FilterImages &img, cudaStream_t s);   // W64 -> B64
cudaError_t launch_mmse_shared_dmma(const FilterImages &img, const void *tx_or_hls, const void *rx, int64_t frame_stride, void *H,
                                    int64_t n_frames, cudaStream_t s, void *hp_out = nullptr);   // FP64 I/O, DMMA m8n8k4; hp_out [n][4]: pilot LS
cudaError_t launch_mmse_shared_dmma_eig(const FilterImages &img, const void *u, const void *tx, const void *rx, int64_t frame_stride,
                                        int dc, const void *sigma2, const double *lam, const void *p, double Rdd, double md, void *H,
                                        int64_t n_frames, cudaStream_t s);   // FP64: H = rx/tx - v G2^T, v formed from u in the kernel
cudaError_t launch_cmatmul(wifi_dtype dt, const void *A, int r1, int c1, const void *B, int c2, void *C, int64_t batch,
                           cudaStream_t s);
cudaError_t launch_chermitian(wifi_dtype dt, int mode, const void *M, int row, int col, void *res, int64_t batch, cudaStream_t s);
cudaError_t launch_cadd(wifi_dtype dt, int mode, const void *M1, const void *M2, void *res, int64_t n_elems, cudaStream_t s);
cudaError_t launch_couter(wifi_dtype dt, const void *M1, int r1, int c1, const void *M2, int c2, void *res, int64_t batch,
                          cudaStream_t s);
cudaError_t launch_cidentity(wifi_dtype dt, void *Id, int size, double scalar, int64_t batch, cudaStream_t s);

// synthetic frames + statistics (wifi_synth.cu)
cudaError_t launch_synth(wifi_dtype dt, uint64_t seed, int64_t first, int64_t n, int per_frame_sigma, void *tx_pre, void *rx_pre,
                         void *tx_symb, void *rx_symb, void *H_true, void *sigma2, cudaStream_t s);
cudaError_t launch_synth_cov(void *R64, cudaStream_t s);
cudaError_t launch_error_stats(wifi_dtype dt, const void *H, const void *Href, int64_t n_elems, double *stats, cudaStream_t s);

// eigen-domain per-frame MMSE (wifi_eig.cu)
cudaError_t launch_eig_prepare(const void *R64, const double *absx2, void *W1, void *W2, double *lam, void *p, double *scal, int *info,
                               cudaStream_t s);

// low-rank per-frame MMSE (wifi_lowrank.cu): tables from the eigen pairs of R (host), one launch per batch
cudaError_t launch_mmse_lowrank(wifi_dtype dt, int rank_padded, const void *tab, const void *tx, const void *rx, int64_t frame_stride,
                                const void *sigma2, void *H, int64_t n_frames, cudaStream_t s);

int lowrank_build_tables(const double *V_re_im, const double *lam, int *rank_padded, std::vector<float> &t32, std::vector<double> &t64);

// measured ceilings (wifi_peaks.cu): which = 0 FP32 FMA, 1 FP64 FMA, 2 FP64 DMMA (TFLOP/s), 3 streaming copy (GB/s)
cudaError_t measure_peak(int which, double *value, cudaStream_t s);

// number of kernel launches the last launcher call issued (for gpu_launches accounting)
extern thread_local int g_last_launches;

}  // namespace wifi
