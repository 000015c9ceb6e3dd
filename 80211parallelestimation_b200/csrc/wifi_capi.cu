// wifi_capi.cu -- the C-ABI of include/wifi_b200.h: context, argument checking, kernel selection by dtype,
// and the host-pointer variants (chunked H2D -> kernels -> D2H on two streams).  No CPU compute path exists
// here: every entry point either launches the sm_100a kernels or returns an error code.
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <mutex>
#include <vector>

#include "wifi_internal.h"

using namespace wifi;


struct wifi_ctx {
    int device;
    cudaStream_t stream;
    InterpTables tab;
    FilterImages img;
    FilterImages img_rx;     // W diag(1/tx) for a shared, known tx block vector (wifi_mmse_filter_fold_tx): applied to rx directly
    FilterImages eig[2];     // eigen-domain per-frame MMSE: G = V^H M^-1/2 and G2 = M^1/2 V as shared-filter operands
    double *eig_lam; void *eig_p; double *eig_scal; int eig_valid; int eig_dc; double eig_Rdd, eig_md;
    void *eig_u[2]; size_t eig_u_bytes[2];   // scratch, one per pipeline stream (the two chunks of a *_host call are in flight at the
                                             // same time): [n][53] between the two eigen-domain products, [n][4] pilot records of estimate_all
    void *lr_tab32, *lr_tab64; int lr_rank, lr_rank_padded;    // low-rank per-frame MMSE (wifi_lowrank.cu): tables of U, conj(U_i) U_j, 1/l
    int *d_info;             // device scratch: singularity flags
    int *h_info;             // pinned mirror
    char err[512];
    int64_t launches;
    int timing;
    cudaEvent_t ev0, ev1;
    int ev_valid;
    // host-pointer pipeline
    cudaStream_t hstream[2];
    void *stage[2];
    size_t stage_bytes[2];
    cudaEvent_t hev[2];
    size_t chunk_bytes;      // per-array staging target per chunk of the *_host pipeline (wifi_set_host_chunk_bytes)
};

static size_t esize(wifi_dtype dt) { return dt == WIFI_F32 ? sizeof(float2) : sizeof(double2); }
static size_t rsize(wifi_dtype dt) { return dt == WIFI_F32 ? sizeof(float) : sizeof(double); }

static int fail(wifi_ctx *c, int code, const char *fmt, ...)
{
    if (c) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(c->err, sizeof(c->err), fmt, ap);
        va_end(ap);
    }
    return code;
}
#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e__ = (call);                                                                                  \
        if (e__ != cudaSuccess) return fail(ctx, WIFI_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define NEED(cond)                                                                     \
    do {                                                                               \
        if (!(cond)) return fail(ctx, WIFI_ERR_INVALID, "invalid argument: %s", #cond); \
    } while (0)
#define ENTER()                                     \
    if (!ctx) return WIFI_ERR_INVALID;              \
    CK(cudaSetDevice(ctx->device))

// bracket a launcher with the optional timing events and the launch counter
struct Timed {
    wifi_ctx *c; cudaStream_t s;
    Timed(wifi_ctx *c_, cudaStream_t s_) : c(c_), s(s_) { if (c->timing) cudaEventRecord(c->ev0, s); }
    ~Timed() { if (c->timing) { cudaEventRecord(c->ev1, s); c->ev_valid = 1; } c->launches += g_last_launches; }
};

// ---- interpolation weights (host, long double, once per context) -------------------------------
// H_k = sum_i w[k][i] Hp_i.  Linear: main.c:86-100; Cubic: main.c:112-120 expanded (every divided
// difference / 14, sic); Sinc: main.c:135-145 with sinc of utils.c:727-733 evaluated in double like the reference.
static void build_tables(double *w /* [4][53][4] */)
{
    const int P[4] = {WIFI_P0, WIFI_P1, WIFI_P2, WIFI_P3};
    const long double delta = WIFI_P1 - WIFI_P0;
    for (int k = 0; k < WIFI_NSC; ++k) {
        long double wl[4] = {0, 0, 0, 0};
        int s = k < WIFI_P1 ? 0 : (k < WIFI_P2 ? 1 : 2);
        long double alpha = (k - P[s]) / delta;
        wl[s] = 1.0L - alpha; wl[s + 1] = alpha;
        long double u = (k - WIFI_P0) / delta, v = (k - WIFI_P1) / delta, x = (k - WIFI_P2) / delta;
        // f0 + f01 (k-P0) + f012 (k-P0)(k-P1) + f0123 (k-P0)(k-P1)(k-P2), differences all / delta
        long double wc[4] = {1.0L - u + u * v - u * v * x, u - 2.0L * u * v + 3.0L * u * v * x, u * v - 3.0L * u * v * x, u * v * x};
        for (int i = 0; i < 4; ++i) {
            double a = (double)((k - P[i]) / delta);
            double sc = a != 0 ? sin(M_PI * a) / (M_PI * a) : 1.0;
            w[(0 * WIFI_NSC + k) * 4 + i] = (double)wl[i];
            w[(1 * WIFI_NSC + k) * 4 + i] = (double)wc[i];
            w[(2 * WIFI_NSC + k) * 4 + i] = sc;
            // table 3 = WiFi_channel_estimation_PS_Cubic.m:7-15: Newton form with the TRUE spans 14 / 28 / 42, evaluated
            // on the unit vector e_i
            long double h[4] = {0, 0, 0, 0};
            h[i] = 1;
            long double f01 = (h[1] - h[0]) / 14, f12 = (h[2] - h[1]) / 14, f23 = (h[3] - h[2]) / 14;
            long double f012 = (f12 - f01) / 28, f123 = (f23 - f12) / 28, f0123 = (f123 - f012) / 42;
            w[(3 * WIFI_NSC + k) * 4 + i] = (double)(h[0] + f01 * (k - P[0]) + f012 * (k - P[0]) * (k - P[1]) +
                                                      f0123 * (k - P[0]) * (k - P[1]) * (k - P[2]));
        }
    }
}

static bool alloc_images(FilterImages &im)
{
    const size_t nW = (size_t)WIFI_NSC * WIFI_NSC;
    im.valid = 0;
    return cudaMalloc(&im.W64, nW * sizeof(double2)) == cudaSuccess &&
           cudaMalloc(&im.Bhi, 112 * 112 * sizeof(float)) == cudaSuccess && cudaMalloc(&im.Blo, 112 * 112 * sizeof(float)) == cudaSuccess &&
           cudaMalloc(&im.B64, 112 * WIFI_DMMA_BS * sizeof(double)) == cudaSuccess;
}
static void free_images(FilterImages &im) { cudaFree(im.W64); cudaFree(im.Bhi); cudaFree(im.Blo); cudaFree(im.B64); }

extern "C" {

const char *wifi_version(void) { return "wifi_b200 0.1 (sm_100a)"; }

int wifi_create(int device, wifi_ctx **out)
{
    if (!out) return WIFI_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return WIFI_ERR_NO_DEVICE;
    if (device < 0 || device >= count) return WIFI_ERR_INVALID;
    wifi_ctx *ctx = (wifi_ctx *)calloc(1, sizeof(wifi_ctx));
    if (!ctx) return WIFI_ERR_NOMEM;
    ctx->device = device;
    ctx->chunk_bytes = (size_t)48 << 20;   // 1 Mi frames e2e: 18.2 ms at 48 MB, 18.9 at 16, 22.4 at 4 (the pass is H2D-bound)
    if (cudaSetDevice(device) != cudaSuccess) { free(ctx); return WIFI_ERR_CUDA; }
    double w[4 * WIFI_NSC * 4];
    float wf[4 * WIFI_NSC * 4];
    build_tables(w);
    for (int i = 0; i < 4 * WIFI_NSC * 4; ++i) wf[i] = (float)w[i];
    bool ok = cudaMalloc(&ctx->tab.w64, sizeof(w)) == cudaSuccess && cudaMalloc(&ctx->tab.w32, sizeof(wf)) == cudaSuccess &&
              cudaMemcpy(ctx->tab.w64, w, sizeof(w), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(ctx->tab.w32, wf, sizeof(wf), cudaMemcpyHostToDevice) == cudaSuccess &&
              alloc_images(ctx->img) && alloc_images(ctx->img_rx) && alloc_images(ctx->eig[0]) && alloc_images(ctx->eig[1]) &&
              cudaMalloc(&ctx->eig_lam, 64 * sizeof(double)) == cudaSuccess && cudaMalloc(&ctx->eig_p, 64 * sizeof(double2)) == cudaSuccess &&
              cudaMalloc(&ctx->eig_scal, 4 * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&ctx->d_info, 4096 * sizeof(int)) == cudaSuccess &&
              cudaHostAlloc(&ctx->h_info, 4096 * sizeof(int), cudaHostAllocDefault) == cudaSuccess &&
              cudaEventCreate(&ctx->ev0) == cudaSuccess && cudaEventCreate(&ctx->ev1) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->hstream[0], cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->hstream[1], cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->hev[0], cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->hev[1], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { wifi_destroy(ctx); return WIFI_ERR_CUDA; }
    *out = ctx;
    return WIFI_OK;
}

int wifi_destroy(wifi_ctx *ctx)
{
    if (!ctx) return WIFI_OK;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    cudaFree(ctx->tab.w64); cudaFree(ctx->tab.w32);
    free_images(ctx->img); free_images(ctx->img_rx); free_images(ctx->eig[0]); free_images(ctx->eig[1]);
    cudaFree(ctx->eig_lam); cudaFree(ctx->eig_p); cudaFree(ctx->eig_scal); cudaFree(ctx->eig_u[0]); cudaFree(ctx->eig_u[1]);
    cudaFree(ctx->lr_tab32); cudaFree(ctx->lr_tab64);
    cudaFree(ctx->d_info);
    if (ctx->h_info) cudaFreeHost(ctx->h_info);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (int i = 0; i < 2; ++i) {
        if (ctx->hstream[i]) cudaStreamDestroy(ctx->hstream[i]);
        if (ctx->hev[i]) cudaEventDestroy(ctx->hev[i]);
        cudaFree(ctx->stage[i]);
    }
    free(ctx);
    return WIFI_OK;
}

int wifi_set_stream(wifi_ctx *ctx, void *s) { if (!ctx) return WIFI_ERR_INVALID; ctx->stream = (cudaStream_t)s; return WIFI_OK; }
int wifi_synchronize(wifi_ctx *ctx) { ENTER(); CK(cudaStreamSynchronize(ctx->stream)); return WIFI_OK; }
const char *wifi_last_error(wifi_ctx *ctx) { return ctx ? ctx->err : "null context"; }
int64_t wifi_launch_count(wifi_ctx *ctx) { return ctx ? ctx->launches : 0; }
int wifi_enable_kernel_timing(wifi_ctx *ctx, int on) { if (!ctx) return WIFI_ERR_INVALID; ctx->timing = on; ctx->ev_valid = 0; return WIFI_OK; }
int wifi_last_kernel_ms(wifi_ctx *ctx, float *ms)
{
    ENTER();
    NEED(ms && ctx->ev_valid);
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaEventElapsedTime(ms, ctx->ev0, ctx->ev1));
    return WIFI_OK;
}

// ---- estimators -----------------------------------------------------------------------------
int wifi_lt_ls_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx_pre && rx_pre && H)) && (dt == WIFI_F32 || dt == WIFI_F64));
    Timed t(ctx, ctx->stream);
    CK(launch_lt_ls(dt, tx_pre, rx_pre, H, n, ctx->stream));
    return WIFI_OK;
}

int wifi_ps_batch(wifi_ctx *ctx, wifi_dtype dt, int which, const void *tx, const void *rx, int64_t frame_stride, void *Hl,
                  void *Hc, void *Hs, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (dt == WIFI_F32 || dt == WIFI_F64) && (which & 7) != 0 && which < 16 && frame_stride >= WIFI_NSC);
    NEED(!(which & WIFI_PS_MATLAB) || frame_stride >= 4 * WIFI_NSC);      // MATLAB mode reads OFDM blocks 0..3 of whole frames
    NEED(n == 0 || (tx && rx));
    NEED(n == 0 || ((!(which & WIFI_PS_LINEAR) || Hl) && (!(which & WIFI_PS_CUBIC) || Hc) && (!(which & WIFI_PS_SINC) || Hs)));
    Timed t(ctx, ctx->stream);
    CK(launch_ps(dt, which, tx, rx, frame_stride, Hl, Hc, Hs, n, ctx->tab, ctx->stream));
    return WIFI_OK;
}

int wifi_equalize_batch(wifi_ctx *ctx, wifi_dtype dt, const void *rx, const void *Hlt, const void *Hps, void *eq, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (rx && Hlt && Hps && eq)) && (dt == WIFI_F32 || dt == WIFI_F64));
    Timed t(ctx, ctx->stream);
    CK(launch_equalize(dt, rx, Hlt, Hps, eq, n, ctx->stream));
    return WIFI_OK;
}

int wifi_frontend_batch(wifi_ctx *ctx, wifi_dtype dt, const void *packet, const void *lptot, void *symb, void *pre_fft, void *ow2, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (packet && lptot && symb && pre_fft)) && (dt == WIFI_F32 || dt == WIFI_F64));
    NEED(((((uintptr_t)packet) | ((uintptr_t)lptot)) & 15) == 0);      // samples are read as 16-byte vectors
    Timed t(ctx, ctx->stream);
    CK(launch_frontend(dt, packet, lptot, symb, pre_fft, ow2, n, ctx->stream));
    return WIFI_OK;
}

// ---- MMSE ---------------------------------------------------------------------------------------
static int install_images(wifi_ctx *ctx, FilterImages &im, cudaStream_t s)
{
    CK(launch_filter_install_tc(im, s));
    CK(launch_filter_install_dmma(im, s));
    ctx->launches += 2;
    im.valid = 1;
    return WIFI_OK;
}
static int install_filter(wifi_ctx *ctx, cudaStream_t s) { return install_images(ctx, ctx->img, s); }

int wifi_mmse_filter_form(wifi_ctx *ctx, const void *R, const double *d, void *W_out)
{
    ENTER();
    NEED(R && d);
    CK(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), ctx->stream));
    {
        Timed t(ctx, ctx->stream);
        CK(launch_filter_form(R, d, ctx->img.W64, ctx->d_info, ctx->stream));
    }
    int rc = install_filter(ctx, ctx->stream);
    if (rc) return rc;
    if (W_out) CK(cudaMemcpyAsync(W_out, ctx->img.W64, sizeof(double2) * WIFI_NSC * WIFI_NSC, cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_info[0]) { ctx->img.valid = 0; return fail(ctx, WIFI_ERR_SINGULAR, "R + diag(d) is singular"); }
    return WIFI_OK;
}

int wifi_mmse_filter_set(wifi_ctx *ctx, const void *W)
{
    ENTER();
    NEED(W);
    CK(cudaMemcpyAsync(ctx->img.W64, W, sizeof(double2) * WIFI_NSC * WIFI_NSC, cudaMemcpyDeviceToDevice, ctx->stream));
    return install_filter(ctx, ctx->stream);
}

static int gemm_with(wifi_ctx *ctx, const FilterImages &im, wifi_dtype dt, const void *a, const void *rx, int64_t frame_stride, void *H,
                     int64_t n, cudaStream_t s, void *hp_out = nullptr)
{
    Timed t(ctx, s);
    // FP32: 3xTF32 on the tcgen05 tensor cores; FP64: DMMA (tcgen05 has no FP64 kind)
    if (dt == WIFI_F32) CK(launch_mmse_shared_tc(im, a, rx, frame_stride, H, n, s, hp_out));
    else CK(launch_mmse_shared_dmma(im, a, rx, frame_stride, H, n, s, hp_out));
    return WIFI_OK;
}

// eigen-domain per-frame MMSE on device pointers (wifi_eig.cu): two shared-matrix products with a per-frame scaling between
static int mmse_eig(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, void *H, int64_t n,
                    cudaStream_t s, int slot = 0)
{
    if (!ctx->eig_valid) return fail(ctx, WIFI_ERR_STATE, "no eigen-domain operands installed: call wifi_mmse_eig_prepare first");
    if (n == 0) return WIFI_OK;
    // Two launches, u = (rx/tx) G^T through a scratch array.  Measured and rejected (round 2): ONE tcgen05 kernel that keeps u in
    // tensor memory and reads the same filter images transposed for the second product (G2 = M G^H, MN-major operand descriptor):
    // kind::tf32 with an MN-major, un-swizzled B operand returned zeros on this hardware, a second image pair does not fit next to
    // the staging buffers (308 KB), and with the A region of TMEM shared by both products the tiles serialise: 0.516 ms per 1 Mi
    // frames (2.0 G frames/s) against 0.65 ms for the two launches -- not worth a cta_group::2 rewrite.
    const size_t need = (size_t)n * WIFI_NSC * esize(dt);
    if (ctx->eig_u_bytes[slot] < need) {
        CK(cudaStreamSynchronize(s));
        cudaFree(ctx->eig_u[slot]); ctx->eig_u[slot] = nullptr; ctx->eig_u_bytes[slot] = 0;
        if (cudaMalloc(&ctx->eig_u[slot], need) != cudaSuccess) return fail(ctx, WIFI_ERR_NOMEM, "eigen-domain scratch cudaMalloc(%zu) failed", need);
        ctx->eig_u_bytes[slot] = need;
    }
    void *U = ctx->eig_u[slot];
    int rc = gemm_with(ctx, ctx->eig[0], dt, tx, rx, frame_stride, U, n, s);            // u = (rx/tx) G^T
    if (rc) return rc;
    // v = s (.) (u - p z_d) in the converter / producer stage, c = v G2^T on the tensor cores, H = rx/tx - c in the epilogue: y stays
    // exact and only the correction passes through the product (the alternative H = (u - v) G2^T, without the second read of tx / rx,
    // was measured in round 2: faster, but ~2e-6 of the frame's peak everywhere instead of ~2e-6 of the correction -- rejected)
    Timed t(ctx, s);
    if (dt == WIFI_F32)
        CK(launch_mmse_shared_tc_eig_h(ctx->eig[1], U, tx, rx, frame_stride, ctx->eig_dc, sigma2, ctx->eig_lam, ctx->eig_p, ctx->eig_Rdd,
                                       ctx->eig_md, H, n, s));
    else
        CK(launch_mmse_shared_dmma_eig(ctx->eig[1], U, tx, rx, frame_stride, ctx->eig_dc, sigma2, ctx->eig_lam, ctx->eig_p, ctx->eig_Rdd,
                                       ctx->eig_md, H, n, s));
    return WIFI_OK;
}

static int mmse_shared(wifi_ctx *ctx, wifi_dtype dt, const void *a, const void *rx, int64_t frame_stride, void *H, int64_t n,
                       cudaStream_t s)
{
    if (!ctx->img.valid) return fail(ctx, WIFI_ERR_STATE, "no shared filter installed: call wifi_mmse_filter_form/_set first");
    return gemm_with(ctx, ctx->img, dt, a, rx, frame_stride, H, n, s);
}

// Shared, known tx block vector (training symbols): rx/tx = diag(1/tx) rx, so the divide folds into the filter once,
// W' = W diag(1/tx), and the per-frame work is the plain product H = rx W'^T -- 848 instead of 1 272 bytes per frame (FP32).
int wifi_mmse_filter_fold_tx(wifi_ctx *ctx, const void *tx_block_f64)
{
    ENTER();
    NEED(tx_block_f64);
    if (!ctx->img.valid) return fail(ctx, WIFI_ERR_STATE, "no shared filter installed: call wifi_mmse_filter_form/_set first");
    {
        Timed t(ctx, ctx->stream);
        CK(launch_filter_fold(ctx->img.W64, tx_block_f64, ctx->img_rx.W64, ctx->stream));
    }
    return install_images(ctx, ctx->img_rx, ctx->stream);
}

int wifi_mmse_shared_rx_batch(wifi_ctx *ctx, wifi_dtype dt, const void *rx, int64_t frame_stride, void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (rx && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    if (!ctx->img_rx.valid) return fail(ctx, WIFI_ERR_STATE, "no tx-folded filter installed: call wifi_mmse_filter_fold_tx first");
    return gemm_with(ctx, ctx->img_rx, dt, rx, nullptr, frame_stride, H, n, ctx->stream);
}

int wifi_mmse_shared_apply_batch(wifi_ctx *ctx, wifi_dtype dt, const void *H_ls, void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (H_ls && H)) && (dt == WIFI_F32 || dt == WIFI_F64));
    return mmse_shared(ctx, dt, H_ls, nullptr, WIFI_NSC, H, n, ctx->stream);
}

int wifi_mmse_shared_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, int64_t frame_stride, void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx && rx && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    return mmse_shared(ctx, dt, tx, rx, frame_stride, H, n, ctx->stream);
}

// ---- low-rank covariance: H = U (sigma2 L^-1 + U^H diag(|x|^2) U)^-1 U^H (conj(x) (.) rx), one launch (wifi_lowrank.cu) ----
int wifi_mmse_lowrank_prepare(wifi_ctx *ctx, const void *R, int *rank_out)
{
    ENTER();
    NEED(R);
    ctx->lr_rank = 0;
    if (rank_out) *rank_out = 0;
    // eigen-decomposition of R itself: the eigen-domain set-up kernel with |x| = 1 (no null bin, S = R), into scratch of its own
    const size_t nW = sizeof(double2) * WIFI_NSC * WIFI_NSC;
    char *scr = nullptr;
    const size_t off_w2 = nW, off_lam = 2 * nW, off_p = off_lam + 64 * sizeof(double), off_scal = off_p + 64 * sizeof(double2),
                 off_ones = off_scal + 4 * sizeof(double), total = off_ones + 64 * sizeof(double);
    if (cudaMalloc(&scr, total) != cudaSuccess) return fail(ctx, WIFI_ERR_NOMEM, "low-rank set-up scratch cudaMalloc(%zu) failed", total);
    std::vector<double> hbuf(2 * WIFI_NSC * WIFI_NSC + 64, 1.0);
    double *hV = hbuf.data(), *hlam = hbuf.data() + 2 * WIFI_NSC * WIFI_NSC;
    int rc = WIFI_OK;
    cudaError_t e = cudaMemcpyAsync(scr + off_ones, hlam, 64 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream);      // 64 ones
    if (e == cudaSuccess) e = cudaMemsetAsync(ctx->d_info, 0, sizeof(int), ctx->stream);
    if (e == cudaSuccess) {
        Timed t(ctx, ctx->stream);
        e = launch_eig_prepare(R, (const double *)(scr + off_ones), scr, scr + off_w2, (double *)(scr + off_lam), scr + off_p,
                               (double *)(scr + off_scal), ctx->d_info, ctx->stream);
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(hV, scr + off_w2, nW, cudaMemcpyDeviceToHost, ctx->stream);                // W2 = V (columns)
    if (e == cudaSuccess) e = cudaMemcpyAsync(hlam, scr + off_lam, WIFI_NSC * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(scr);
    if (e != cudaSuccess) return fail(ctx, WIFI_ERR_CUDA, "low-rank set-up: %s", cudaGetErrorString(e));
    std::vector<float> t32;
    std::vector<double> t64;
    int padded = 0;
    const int r = lowrank_build_tables(hV, hlam, &padded, t32, t64);
    if (rank_out) *rank_out = r;
    if (r == 0) return fail(ctx, WIFI_ERR_INVALID, "low-rank MMSE: the covariance has no positive eigenvalue");
    if (r > WIFI_LOWRANK_MAX)
        return fail(ctx, WIFI_ERR_INVALID, "low-rank MMSE: numerical rank %d of the covariance exceeds %d -- use wifi_mmse_perframe_* or wifi_mmse_eig_*", r,
                    WIFI_LOWRANK_MAX);
    cudaFree(ctx->lr_tab32); cudaFree(ctx->lr_tab64); ctx->lr_tab32 = ctx->lr_tab64 = nullptr;
    CK(cudaMalloc(&ctx->lr_tab32, t32.size() * sizeof(float)));
    CK(cudaMalloc(&ctx->lr_tab64, t64.size() * sizeof(double)));
    CK(cudaMemcpy(ctx->lr_tab32, t32.data(), t32.size() * sizeof(float), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->lr_tab64, t64.data(), t64.size() * sizeof(double), cudaMemcpyHostToDevice));
    ctx->lr_rank = r; ctx->lr_rank_padded = padded;
    return rc;
}

static int mmse_lowrank(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, void *H, int64_t n,
                        cudaStream_t s)
{
    if (!ctx->lr_rank) return fail(ctx, WIFI_ERR_STATE, "no low-rank operands installed: call wifi_mmse_lowrank_prepare first");
    Timed t(ctx, s);
    CK(launch_mmse_lowrank(dt, ctx->lr_rank_padded, dt == WIFI_F32 ? ctx->lr_tab32 : ctx->lr_tab64, tx, rx, frame_stride, sigma2, H, n, s));
    return WIFI_OK;
}

int wifi_mmse_perframe_lowrank_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2,
                                     void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx && rx && sigma2 && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    return mmse_lowrank(ctx, dt, tx, rx, frame_stride, sigma2, H, n, ctx->stream);
}

static int mmse_perframe(wifi_ctx *ctx, wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                         const void *sigma2, const void *hls, void *H, int64_t n, int flags, cudaStream_t s, bool check);

// fused receiver chain on device pointers (wifi_frontend.cu rx_chain_kernel); the shared-filter PS_MMSE, when asked for, is the
// tcgen05 / DMMA GEMM on the H_ls0 plane the chain kernel wrote (scratch if the caller does not want it)
static int rx_chain(wifi_ctx *ctx, wifi_dtype dt, const void *tx_packet, int64_t tx_pkt_stride, const void *tx_lptot, const void *rx_packet,
                    const void *rx_lptot, const wifi_rx_chain_out &o, int64_t n, cudaStream_t s, int slot = 0)
{
    if (n == 0) return WIFI_OK;
    void *hls0 = o.H_ls0;
    if (o.H_mmse_shared) {
        if (!ctx->img.valid) return fail(ctx, WIFI_ERR_STATE, "no shared filter installed: call wifi_mmse_filter_form/_set first");
        if (!hls0) {
            const size_t need = (size_t)n * WIFI_NSC * esize(dt);
            if (ctx->eig_u_bytes[slot] < need) {
                CK(cudaStreamSynchronize(s));
                cudaFree(ctx->eig_u[slot]); ctx->eig_u[slot] = nullptr; ctx->eig_u_bytes[slot] = 0;
                if (cudaMalloc(&ctx->eig_u[slot], need) != cudaSuccess) return fail(ctx, WIFI_ERR_NOMEM, "H_ls scratch cudaMalloc(%zu) failed", need);
                ctx->eig_u_bytes[slot] = need;
            }
            hls0 = ctx->eig_u[slot];
        }
    }
    {
        Timed t(ctx, s);
        CK(launch_rx_chain(dt, tx_packet, tx_pkt_stride, tx_lptot, rx_packet, rx_lptot, o.H_lt, o.H_linear, o.H_cubic, o.H_sinc, o.H_mmse_cconv,
                           hls0, o.eq, o.rx_symb, o.ow2, n, ctx->tab, s));
    }
    if (o.H_mmse_shared) return gemm_with(ctx, ctx->img, dt, hls0, nullptr, WIFI_NSC, o.H_mmse_shared, n, s);
    return WIFI_OK;
}

int wifi_rx_chain_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_packet, const void *tx_lptot, const void *rx_packet, const void *rx_lptot,
                        const wifi_rx_chain_out *out, int64_t n)
{
    ENTER();
    NEED(out && n >= 0 && (n == 0 || (tx_packet && tx_lptot && rx_packet && rx_lptot)) && (dt == WIFI_F32 || dt == WIFI_F64));
    NEED(((((uintptr_t)tx_packet) | ((uintptr_t)tx_lptot) | ((uintptr_t)rx_packet) | ((uintptr_t)rx_lptot)) & 15) == 0);   // 16-byte vector loads
    return rx_chain(ctx, dt, tx_packet, WIFI_PACKET, tx_lptot, rx_packet, rx_lptot, *out, n, ctx->stream);
}

// BASELINE configs[4]: all five estimators (+ the equalizer) of n frames behind one call, four launches: LT_LS; the shared-filter
// PS_MMSE GEMM (reading block 0 of tx / rx in place), whose LS-divide stage also hands the four pilot LS values of every frame
// to the interpolators as a 32-byte record; the three interpolators from those records (no pilot gather: 8 isolated values per
// frame cost 512 B of DRAM traffic, the records 32 B); the equalizer.  Two deeper fusions were built, measured on B200 and
// rejected (round 2, 1 Mi frames, FP32; four stand-alone launches take 2.93 ms):
//   * LT_LS and the interpolators emitted by the GEMM kernel's converter warps from the block LS values they stage (one launch
//     for the five estimates): those 8 warps are latency-bound, the launch took 1.00 ms against 0.83 ms for the three kernels;
//   * a "frame finish" pass (LT_LS + interpolators + equalizer, no estimate re-read: 1 912 instead of 2 022 complex values per
//     frame), as a shared-memory tile kernel with flat 16-byte vectors (3.8 ms) and as one thread per (frame, sub-carrier) with
//     the estimates in registers (2.73 ms at 6.0 TB/s of DRAM traffic; 8-byte accesses, 27 warps per SM): 3.05 ms with the GEMM.
static int estimate_all(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, const void *tx, const void *rx, int64_t frame_stride,
                        void *H_lt, void *H_lin, void *H_cub, void *H_sinc, void *H_mmse, void *eq, int64_t n, cudaStream_t s, int slot = 0)
{
    if (!ctx->img.valid) return fail(ctx, WIFI_ERR_STATE, "no shared filter installed: call wifi_mmse_filter_form/_set first");
    if (n == 0) return WIFI_OK;
    const size_t need = (size_t)n * 4 * esize(dt);                    // pilot records [n][4]; the scratch is shared with the eigen-domain path
    if (ctx->eig_u_bytes[slot] < need) {
        CK(cudaStreamSynchronize(s));
        cudaFree(ctx->eig_u[slot]); ctx->eig_u[slot] = nullptr; ctx->eig_u_bytes[slot] = 0;
        if (cudaMalloc(&ctx->eig_u[slot], need) != cudaSuccess) return fail(ctx, WIFI_ERR_NOMEM, "pilot scratch cudaMalloc(%zu) failed", need);
        ctx->eig_u_bytes[slot] = need;
    }
    void *hp = ctx->eig_u[slot];
    { Timed t(ctx, s); CK(launch_lt_ls(dt, tx_pre, rx_pre, H_lt, n, s)); }
    int rc = gemm_with(ctx, ctx->img, dt, tx, rx, frame_stride, H_mmse, n, s, hp);
    if (rc) return rc;
    { Timed t(ctx, s); CK(launch_ps(dt, WIFI_PS_LINEAR | WIFI_PS_CUBIC | WIFI_PS_SINC, tx, rx, frame_stride, H_lin, H_cub, H_sinc, n, ctx->tab, s, hp)); }
    if (eq) { Timed t(ctx, s); CK(launch_equalize(dt, rx, H_lt, H_lin, eq, n, s)); }      // WiFi_RX.m:60: H_EST_LT_LS and H_EST_PS_Linear
    return WIFI_OK;
}

int wifi_estimate_all_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, const void *tx_symbols, const void *rx_symbols,
                            int64_t frame_stride, void *H_lt, void *H_linear, void *H_cubic, void *H_sinc, void *H_mmse, void *eq, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    NEED(n == 0 || (tx_pre && rx_pre && tx_symbols && rx_symbols && H_lt && H_linear && H_cubic && H_sinc && H_mmse));
    NEED(!eq || frame_stride == WIFI_FRAME);                  // the equalizer works on whole frames [n][15][53]
    return estimate_all(ctx, dt, tx_pre, rx_pre, tx_symbols, rx_symbols, frame_stride, H_lt, H_linear, H_cubic, H_sinc, H_mmse, eq, n, ctx->stream);
}

// R_f = H_ls H_ls^H: closed form (wifi_ls.cu mmse_rank1_kernel)
static int mmse_cconv(wifi_ctx *ctx, wifi_dtype dt, int matlab, const void *tx, const void *rx, int64_t frame_stride, const void *ow2,
                      const void *hls, void *H, int64_t n, cudaStream_t s)
{
    Timed t(ctx, s);
    CK(launch_mmse_rank1(dt, matlab, tx, rx, frame_stride, ow2, hls, H, n, s));
    return WIFI_OK;
}

int wifi_mmse_eig_prepare(wifi_ctx *ctx, const void *R, const double *absx2)
{
    ENTER();
    NEED(R && absx2);
    ctx->eig_valid = 0;
    CK(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), ctx->stream));
    {
        Timed t(ctx, ctx->stream);
        CK(launch_eig_prepare(R, absx2, ctx->eig[0].W64, ctx->eig[1].W64, ctx->eig_lam, ctx->eig_p, ctx->eig_scal, ctx->d_info, ctx->stream));
    }
    int rc = install_images(ctx, ctx->eig[0], ctx->stream);
    if (!rc) rc = install_images(ctx, ctx->eig[1], ctx->stream);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->h_info, ctx->d_info, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (*ctx->h_info) return fail(ctx, WIFI_ERR_INVALID, "eigen-domain MMSE supports at most one null bin (|x_k|^2 < 1e-6 max) and needs max |x|^2 > 0");
    double scal[4];
    CK(cudaMemcpy(scal, ctx->eig_scal, sizeof scal, cudaMemcpyDeviceToHost));
    ctx->eig_dc = (int)scal[2]; ctx->eig_Rdd = scal[0]; ctx->eig_md = scal[1];
    ctx->eig_valid = 1;
    return WIFI_OK;
}

int wifi_mmse_perframe_eig_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2,
                                 void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx && rx && sigma2 && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    return mmse_eig(ctx, dt, tx, rx, frame_stride, sigma2, H, n, ctx->stream);
}

static int mmse_perframe(wifi_ctx *ctx, wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                         const void *sigma2, const void *hls, void *H, int64_t n, int flags, cudaStream_t s, bool check)
{
    if (check) CK(cudaMemsetAsync(ctx->d_info, 0, sizeof(int), s));
    {
        Timed t(ctx, s);
        const int fast32 = (flags & WIFI_SOLVE_FAST32) ? 1 : 0;      // FP32 arithmetic is an explicit opt-in (4e-3 / 1e-1 accuracy)
        if ((flags & WIFI_SOLVE_HPD) && !hls)
            CK(launch_mmse_perframe_hpd(dt, R, tx, rx, frame_stride, sigma2, H, n, fast32, s));
        else
            CK(launch_mmse_perframe_pivot(dt, R, tx, rx, frame_stride, sigma2, hls, H, n, check ? ctx->d_info : nullptr, fast32, s));
    }
    return WIFI_OK;
}

int wifi_mmse_perframe_batch(wifi_ctx *ctx, wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                             const void *sigma2, void *H, int64_t n, int flags)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (R && tx && rx && sigma2 && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    return mmse_perframe(ctx, dt, R, tx, rx, frame_stride, sigma2, nullptr, H, n, flags, ctx->stream, false);
}

int wifi_mmse_cconv_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, const void *ow2, const void *H_ls, void *H,
                          int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx && rx && ow2 && H_ls && H)) && (dt == WIFI_F32 || dt == WIFI_F64));
    return mmse_cconv(ctx, dt, 0, tx, rx, WIFI_NSC, ow2, H_ls, H, n, ctx->stream);
}

int wifi_mmse_matlab_batch(wifi_ctx *ctx, wifi_dtype dt, const void *tx_frames, const void *rx_frames, const void *ow2, const void *H_ls,
                           void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx_frames && rx_frames && ow2 && H_ls && H)) && (dt == WIFI_F32 || dt == WIFI_F64));
    return mmse_cconv(ctx, dt, 1, tx_frames, rx_frames, WIFI_FRAME, ow2, H_ls, H, n, ctx->stream);
}

// ---- utils ------------------------------------------------------------------------------------------
static bool order_ok(int a, int b) { return a >= 0 && b >= 0 && a <= WIFI_MAX_ORDER && b <= WIFI_MAX_ORDER; }

int wifi_cmatmul_batch(wifi_ctx *ctx, wifi_dtype dt, const void *A, int r1, int c1, const void *B, int r2, int c2, void *C, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(r1, c1) && order_ok(r2, c2) && (dt == WIFI_F32 || dt == WIFI_F64));
    if (c1 != r2) return fail(ctx, WIFI_ERR_INVALID, "Matrices dimension missmatch");   // utils.c:18-19: nothing written
    NEED(batch == 0 || (A && B && C));
    Timed t(ctx, ctx->stream);
    CK(launch_cmatmul(dt, A, r1, c1, B, c2, C, batch, ctx->stream));
    return WIFI_OK;
}

int wifi_chermitian_batch(wifi_ctx *ctx, wifi_dtype dt, int mode, const void *M, int row, int col, void *res, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(row, col) && (mode == WIFI_AS_WRITTEN || mode == WIFI_INTENDED) && (batch == 0 || (M && res)));
    Timed t(ctx, ctx->stream);
    CK(launch_chermitian(dt, mode, M, row, col, res, batch, ctx->stream));
    return WIFI_OK;
}

int wifi_cadd_batch(wifi_ctx *ctx, wifi_dtype dt, int mode, const void *M1, int r1, int c1, const void *M2, int r2, int c2, void *res,
                    int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(r1, c1) && (mode == WIFI_AS_WRITTEN || mode == WIFI_INTENDED));
    if (r1 != r2 || c1 != c2) return fail(ctx, WIFI_ERR_INVALID, "Matrices dimension missmatch");   // utils.c:112-113
    NEED(batch == 0 || (M1 && res && (mode == WIFI_AS_WRITTEN || M2)));
    Timed t(ctx, ctx->stream);
    CK(launch_cadd(dt, mode, M1, M2, res, batch * r1 * c1, ctx->stream));
    return WIFI_OK;
}

int wifi_couter_batch(wifi_ctx *ctx, wifi_dtype dt, const void *M1, int r1, int c1, const void *M2, int r2, int c2, void *res, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(r1, c1) && order_ok(r2, c2));
    if (c1 != r2) return fail(ctx, WIFI_ERR_INVALID, "Matrices dimension missmatch");   // utils.c:56-57
    NEED(batch == 0 || (M1 && M2 && res));
    Timed t(ctx, ctx->stream);
    CK(launch_couter(dt, M1, r1, c1, M2, c2, res, batch, ctx->stream));
    return WIFI_OK;
}

int wifi_cidentity_batch(wifi_ctx *ctx, wifi_dtype dt, void *Id, int size, double scalar, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(size, size) && (batch == 0 || Id));
    Timed t(ctx, ctx->stream);
    CK(launch_cidentity(dt, Id, size, scalar, batch, ctx->stream));
    return WIFI_OK;
}

int wifi_cinverse_batch(wifi_ctx *ctx, wifi_dtype dt, const void *A, int order, void *Y, int64_t batch, int *info)
{
    ENTER();
    NEED(batch >= 0 && order >= 1 && order <= WIFI_MAX_ORDER && (batch == 0 || (A && Y)));
    Timed t(ctx, ctx->stream);
    CK(launch_cinverse(dt, A, order, Y, batch, info, ctx->stream));
    return WIFI_OK;
}

// ---- synthetic data / statistics ---------------------------------------------------------------------
int wifi_synth_frames(wifi_ctx *ctx, wifi_dtype dt, uint64_t seed, int64_t first, int64_t n, int per_frame_sigma, void *tx_pre,
                      void *rx_pre, void *tx_symb, void *rx_symb, void *H_true, void *sigma2)
{
    ENTER();
    NEED(n >= 0 && first >= 0);
    Timed t(ctx, ctx->stream);
    CK(launch_synth(dt, seed, first, n, per_frame_sigma, tx_pre, rx_pre, tx_symb, rx_symb, H_true, sigma2, ctx->stream));
    return WIFI_OK;
}

int wifi_synth_covariance(wifi_ctx *ctx, void *R)
{
    ENTER();
    NEED(R);
    Timed t(ctx, ctx->stream);
    CK(launch_synth_cov(R, ctx->stream));
    return WIFI_OK;
}

int wifi_error_stats(wifi_ctx *ctx, wifi_dtype dt, const void *H, const void *Href, int64_t n_elems, double *stats)
{
    ENTER();
    NEED(n_elems >= 0 && stats && (n_elems == 0 || (H && Href)));
    Timed t(ctx, ctx->stream);
    CK(launch_error_stats(dt, H, Href, n_elems, stats, ctx->stream));
    return WIFI_OK;
}

int wifi_measure_peak(wifi_ctx *ctx, int which, double *value)
{
    ENTER();
    NEED(which >= 0 && which <= 3 && value);
    CK(measure_peak(which, value, ctx->stream));
    return WIFI_OK;
}

// ---- host-pointer pipeline ------------------------------------------------------------------------------
int wifi_set_host_chunk_bytes(wifi_ctx *ctx, size_t bytes)
{
    if (!ctx || bytes < 4096) return WIFI_ERR_INVALID;
    ctx->chunk_bytes = bytes;
    return WIFI_OK;
}
int wifi_host_alloc(void **p, size_t bytes) { return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? WIFI_OK : WIFI_ERR_NOMEM; }
int wifi_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? WIFI_OK : WIFI_ERR_CUDA; }

}  // extern "C"

namespace {

// one per-frame array crossing the PCIe boundary
struct Arr {
    const void *h_in;    // host source (inputs) or nullptr
    void *h_out;         // host destination (outputs) or nullptr
    size_t row_bytes;    // bytes per frame actually needed
    size_t pitch_bytes;  // host distance between consecutive frames
};


// Runs `body(dev_ptrs, n_chunk, stream)` over chunks of frames; arrays are staged compactly (pitch = row_bytes).
template <typename Body>
int host_pipeline(wifi_ctx *ctx, int64_t n_frames, const std::vector<Arr> &arrs, Body body)
{
    if (n_frames == 0) return WIFI_OK;
    size_t max_row = 1, sum_row = 0;
    for (auto &a : arrs) { max_row = std::max(max_row, a.row_bytes); sum_row += (a.row_bytes + 255) / 256 * 256; }
    int64_t chunk = std::max<int64_t>(1, (int64_t)(ctx->chunk_bytes / max_row));
    chunk = std::min(chunk, n_frames);
    if (chunk > 1 && (chunk & 1)) --chunk;                      // keep 16-byte alignment of [n][53] float2 chunks
    std::vector<size_t> off(arrs.size());
    size_t need = 0;
    for (size_t i = 0; i < arrs.size(); ++i) { off[i] = need; need += (arrs[i].row_bytes * (size_t)chunk + 255) / 256 * 256; }
    for (int b = 0; b < 2; ++b)
        if (ctx->stage_bytes[b] < need) {
            cudaFree(ctx->stage[b]);
            ctx->stage[b] = nullptr; ctx->stage_bytes[b] = 0;
            if (cudaMalloc(&ctx->stage[b], need) != cudaSuccess) return fail(ctx, WIFI_ERR_NOMEM, "staging cudaMalloc(%zu) failed", need);
            ctx->stage_bytes[b] = need;
        }
    std::vector<void *> dev(arrs.size());
    int it = 0;
    for (int64_t f0 = 0; f0 < n_frames; f0 += chunk, ++it) {
        const int b = it & 1;
        const int64_t nc = std::min(chunk, n_frames - f0);
        cudaStream_t s = ctx->hstream[b];
        for (size_t i = 0; i < arrs.size(); ++i) {
            dev[i] = (char *)ctx->stage[b] + off[i];
            const Arr &a = arrs[i];
            if (a.h_in) {
                const char *src = (const char *)a.h_in + (size_t)f0 * a.pitch_bytes;
                cudaError_t e = a.pitch_bytes == a.row_bytes
                                    ? cudaMemcpyAsync(dev[i], src, a.row_bytes * (size_t)nc, cudaMemcpyHostToDevice, s)
                                    : cudaMemcpy2DAsync(dev[i], a.row_bytes, src, a.pitch_bytes, a.row_bytes, (size_t)nc, cudaMemcpyHostToDevice, s);
                if (e != cudaSuccess) return fail(ctx, WIFI_ERR_CUDA, "H2D: %s", cudaGetErrorString(e));
            }
        }
        int rc = body(dev, nc, f0, s);
        if (rc) return rc;
        for (size_t i = 0; i < arrs.size(); ++i) {
            const Arr &a = arrs[i];
            if (a.h_out) {
                char *dst = (char *)a.h_out + (size_t)f0 * a.pitch_bytes;
                cudaError_t e = a.pitch_bytes == a.row_bytes
                                    ? cudaMemcpyAsync(dst, dev[i], a.row_bytes * (size_t)nc, cudaMemcpyDeviceToHost, s)
                                    : cudaMemcpy2DAsync(dst, a.pitch_bytes, dev[i], a.row_bytes, a.row_bytes, (size_t)nc, cudaMemcpyDeviceToHost, s);
                if (e != cudaSuccess) return fail(ctx, WIFI_ERR_CUDA, "D2H: %s", cudaGetErrorString(e));
            }
        }
    }
    for (int b = 0; b < 2; ++b) {
        cudaError_t e = cudaStreamSynchronize(ctx->hstream[b]);
        if (e != cudaSuccess) return fail(ctx, WIFI_ERR_CUDA, "pipeline sync: %s", cudaGetErrorString(e));
    }
    return WIFI_OK;
}

Arr in_arr(const void *p, size_t row, size_t pitch) { return Arr{p, nullptr, row, pitch}; }
Arr out_arr(void *p, size_t row) { return Arr{nullptr, p, row, row}; }

}  // namespace

extern "C" {

// small utils: one shot through the ctx stream
#define UTIL_STAGE(total_bytes)                                                   \
    if (ctx->stage_bytes[0] < (total_bytes)) {                                    \
        cudaFree(ctx->stage[0]); ctx->stage[0] = nullptr; ctx->stage_bytes[0] = 0; \
        CK(cudaMalloc(&ctx->stage[0], (total_bytes)));                            \
        ctx->stage_bytes[0] = (total_bytes);                                      \
    }

int wifi_lt_ls_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx_pre && rx_pre && H)) && (dt == WIFI_F32 || dt == WIFI_F64));
    const size_t row = WIFI_NSC * esize(dt);
    return host_pipeline(ctx, n, {in_arr(tx_pre, row, row), in_arr(rx_pre, row, row), out_arr(H, row)},
                         [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                             Timed t(ctx, s);
                             CK(launch_lt_ls(dt, d[0], d[1], d[2], nc, s));
                             return (int)WIFI_OK;
                         });
}

int wifi_ps_host(wifi_ctx *ctx, wifi_dtype dt, int which, const void *tx, const void *rx, int64_t frame_stride, void *Hl, void *Hc,
                 void *Hs, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (dt == WIFI_F32 || dt == WIFI_F64) && (which & 7) != 0 && which < 16 && frame_stride >= WIFI_NSC && (n == 0 || (tx && rx)));
    NEED(!(which & WIFI_PS_MATLAB) || frame_stride >= 4 * WIFI_NSC);
    NEED(n == 0 || ((!(which & WIFI_PS_LINEAR) || Hl) && (!(which & WIFI_PS_CUBIC) || Hc) && (!(which & WIFI_PS_SINC) || Hs)));
    const int nb = (which & WIFI_PS_MATLAB) ? 4 : 1;                       // OFDM blocks that cross the bus per frame
    const size_t row = WIFI_NSC * esize(dt), pitch = (size_t)frame_stride * esize(dt);
    std::vector<Arr> arrs = {in_arr(tx, nb * row, pitch), in_arr(rx, nb * row, pitch)};
    int il = -1, ic = -1, is = -1;
    if (which & WIFI_PS_LINEAR) { il = (int)arrs.size(); arrs.push_back(out_arr(Hl, row)); }
    if (which & WIFI_PS_CUBIC) { ic = (int)arrs.size(); arrs.push_back(out_arr(Hc, row)); }
    if (which & WIFI_PS_SINC) { is = (int)arrs.size(); arrs.push_back(out_arr(Hs, row)); }
    return host_pipeline(ctx, n, arrs, [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
        Timed t(ctx, s);
        CK(launch_ps(dt, which, d[0], d[1], nb * WIFI_NSC, il >= 0 ? d[il] : nullptr, ic >= 0 ? d[ic] : nullptr, is >= 0 ? d[is] : nullptr,
                     nc, ctx->tab, s));
        return (int)WIFI_OK;
    });
}

int wifi_equalize_host(wifi_ctx *ctx, wifi_dtype dt, const void *rx, const void *Hlt, const void *Hps, void *eq, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (rx && Hlt && Hps && eq)) && (dt == WIFI_F32 || dt == WIFI_F64));
    const size_t row = WIFI_NSC * esize(dt), frow = WIFI_FRAME * esize(dt);
    return host_pipeline(ctx, n, {in_arr(rx, frow, frow), in_arr(Hlt, row, row), in_arr(Hps, row, row), out_arr(eq, frow)},
                         [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                             Timed t(ctx, s);
                             CK(launch_equalize(dt, d[0], d[1], d[2], d[3], nc, s));
                             return (int)WIFI_OK;
                         });
}

int wifi_frontend_host(wifi_ctx *ctx, wifi_dtype dt, const void *packet, const void *lptot, void *symb, void *pre_fft, void *ow2, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (packet && lptot && symb && pre_fft)) && (dt == WIFI_F32 || dt == WIFI_F64));
    const size_t es = esize(dt);
    std::vector<Arr> arrs = {in_arr(packet, WIFI_PACKET * es, WIFI_PACKET * es), in_arr(lptot, WIFI_LPTOT * es, WIFI_LPTOT * es),
                             out_arr(symb, WIFI_FRAME * es), out_arr(pre_fft, WIFI_NSC * es)};
    if (ow2) arrs.push_back(out_arr(ow2, es / 2));
    return host_pipeline(ctx, n, arrs, [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
        Timed t(ctx, s);
        CK(launch_frontend(dt, d[0], d[1], d[2], d[3], ow2 ? d[4] : nullptr, nc, s));
        return (int)WIFI_OK;
    });
}

int wifi_rx_chain_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_packet, const void *tx_lptot, const void *rx_packet, const void *rx_lptot,
                       const wifi_rx_chain_out *out, int64_t n)
{
    ENTER();
    NEED(out && n >= 0 && (n == 0 || (tx_packet && tx_lptot && rx_packet && rx_lptot)) && (dt == WIFI_F32 || dt == WIFI_F64));
    const size_t es = esize(dt), row = WIFI_NSC * es, frow = WIFI_FRAME * es;
    // of tx_packet only OFDM block 0 with its cyclic prefix (80 samples) is staged; the kernel then reads it at a stride of 80
    std::vector<Arr> arrs = {in_arr(tx_packet, 80 * es, WIFI_PACKET * es), in_arr(tx_lptot, WIFI_LPTOT * es, WIFI_LPTOT * es),
                             in_arr(rx_packet, WIFI_PACKET * es, WIFI_PACKET * es), in_arr(rx_lptot, WIFI_LPTOT * es, WIFI_LPTOT * es)};
    void *const outs[10] = {out->H_lt, out->H_linear, out->H_cubic, out->H_sinc, out->H_mmse_cconv, out->H_ls0, out->H_mmse_shared, out->eq,
                            out->rx_symb, out->ow2};
    const size_t rows[10] = {row, row, row, row, row, row, row, frow, frow, es / 2};
    int at[10];
    for (int i = 0; i < 10; ++i) { at[i] = -1; if (outs[i]) { at[i] = (int)arrs.size(); arrs.push_back(out_arr(outs[i], rows[i])); } }
    return host_pipeline(ctx, n, arrs, [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
        wifi_rx_chain_out o;
        void **po[10] = {&o.H_lt, &o.H_linear, &o.H_cubic, &o.H_sinc, &o.H_mmse_cconv, &o.H_ls0, &o.H_mmse_shared, &o.eq, &o.rx_symb, &o.ow2};
        for (int i = 0; i < 10; ++i) *po[i] = at[i] >= 0 ? d[at[i]] : nullptr;
        return rx_chain(ctx, dt, d[0], 80, d[1], d[2], d[3], o, nc, s, s == ctx->hstream[1] ? 1 : 0);
    });
}

int wifi_mmse_filter_form_host(wifi_ctx *ctx, const void *R, const double *d, void *W_out)
{
    ENTER();
    NEED(R && d);
    const size_t nW = sizeof(double2) * WIFI_NSC * WIFI_NSC;
    const size_t need = 2 * nW + 256 * sizeof(double);
    if (ctx->stage_bytes[0] < need) {
        cudaFree(ctx->stage[0]); ctx->stage[0] = nullptr; ctx->stage_bytes[0] = 0;
        CK(cudaMalloc(&ctx->stage[0], need));
        ctx->stage_bytes[0] = need;
    }
    char *base = (char *)ctx->stage[0];
    CK(cudaMemcpyAsync(base, R, nW, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(base + 2 * nW, d, sizeof(double) * WIFI_NSC, cudaMemcpyHostToDevice, ctx->stream));
    int rc = wifi_mmse_filter_form(ctx, base, (const double *)(base + 2 * nW), W_out ? base + nW : nullptr);
    if (rc) return rc;
    if (W_out) {
        CK(cudaMemcpyAsync(W_out, base + nW, nW, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return WIFI_OK;
}

int wifi_mmse_shared_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, int64_t frame_stride, void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx && rx && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    CK(cudaStreamSynchronize(ctx->stream));   // the filter images were built on the ctx stream
    const size_t row = WIFI_NSC * esize(dt), pitch = (size_t)frame_stride * esize(dt);
    return host_pipeline(ctx, n, {in_arr(tx, row, pitch), in_arr(rx, row, pitch), out_arr(H, row)},
                         [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                             return mmse_shared(ctx, dt, d[0], d[1], WIFI_NSC, d[2], nc, s);
                         });
}

int wifi_estimate_all_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_pre, const void *rx_pre, const void *tx_frames, const void *rx_frames,
                           void *H_lt, void *H_linear, void *H_cubic, void *H_sinc, void *H_mmse, void *eq, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (dt == WIFI_F32 || dt == WIFI_F64));
    NEED(n == 0 || (tx_pre && rx_pre && tx_frames && rx_frames && H_lt && H_linear && H_cubic && H_sinc && H_mmse));
    if (!ctx->img.valid) return fail(ctx, WIFI_ERR_STATE, "no shared filter installed: call wifi_mmse_filter_form/_set first");
    CK(cudaStreamSynchronize(ctx->stream));   // the filter images were built on the ctx stream
    // whole frames [n][15][53] on the host; only block 0 of tx crosses the bus, rx crosses whole when the equalizer is asked for
    const size_t row = WIFI_NSC * esize(dt), frow = WIFI_FRAME * esize(dt);
    std::vector<Arr> arrs = {in_arr(tx_pre, row, row), in_arr(rx_pre, row, row), in_arr(tx_frames, row, frow), in_arr(rx_frames, eq ? frow : row, frow),
                             out_arr(H_lt, row), out_arr(H_linear, row), out_arr(H_cubic, row), out_arr(H_sinc, row), out_arr(H_mmse, row)};
    if (eq) arrs.push_back(out_arr(eq, frow));
    return host_pipeline(ctx, n, arrs, [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
        // tx is staged as block vectors (stride 53); rx too without the equalizer, as whole frames (stride 795) with it
        const int slot = s == ctx->hstream[1] ? 1 : 0;
        if (!eq) return estimate_all(ctx, dt, d[0], d[1], d[2], d[3], WIFI_NSC, d[4], d[5], d[6], d[7], d[8], nullptr, nc, s, slot);
        // with the equalizer the estimators want tx and rx at ONE stride: block 0 of rx is gathered on the device into the eq
        // staging buffer, which is free until the equalizer (the last kernel of the chunk) overwrites it
        cudaError_t e = cudaMemcpy2DAsync(d[9], row, d[3], frow, row, (size_t)nc, cudaMemcpyDeviceToDevice, s);
        if (e != cudaSuccess) return fail(ctx, WIFI_ERR_CUDA, "gather: %s", cudaGetErrorString(e));
        int rc = estimate_all(ctx, dt, d[0], d[1], d[2], d[9], WIFI_NSC, d[4], d[5], d[6], d[7], d[8], nullptr, nc, s, slot);
        if (rc) return rc;
        Timed t(ctx, s);
        CK(launch_equalize(dt, d[3], d[4], d[5], d[9], nc, s));
        return (int)WIFI_OK;
    });
}

int wifi_mmse_filter_fold_tx_host(wifi_ctx *ctx, const void *tx_block_f64)
{
    ENTER();
    NEED(tx_block_f64);
    const size_t nb = sizeof(double2) * WIFI_NSC;
    UTIL_STAGE(nb);
    CK(cudaMemcpyAsync(ctx->stage[0], tx_block_f64, nb, cudaMemcpyHostToDevice, ctx->stream));
    return wifi_mmse_filter_fold_tx(ctx, ctx->stage[0]);
}

int wifi_mmse_shared_rx_host(wifi_ctx *ctx, wifi_dtype dt, const void *rx, int64_t frame_stride, void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (rx && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    if (!ctx->img_rx.valid) return fail(ctx, WIFI_ERR_STATE, "no tx-folded filter installed: call wifi_mmse_filter_fold_tx first");
    CK(cudaStreamSynchronize(ctx->stream));   // the filter images were built on the ctx stream
    const size_t row = WIFI_NSC * esize(dt), pitch = (size_t)frame_stride * esize(dt);
    return host_pipeline(ctx, n, {in_arr(rx, row, pitch), out_arr(H, row)},
                         [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                             return gemm_with(ctx, ctx->img_rx, dt, d[0], nullptr, WIFI_NSC, d[1], nc, s);
                         });
}

// PCIe ceiling of this GPU under the same conditions as the *_host pipeline: one H2D copy of h2d_bytes and one D2H copy of
// d2h_bytes, pinned host buffers, issued together on the pipeline's two streams; *ms = host wall time until both are done.
int wifi_pcie_probe(wifi_ctx *ctx, const void *h_src, void *h_dst, size_t h2d_bytes, size_t d2h_bytes, double *ms)
{
    ENTER();
    NEED(ms && (h2d_bytes == 0 || h_src) && (d2h_bytes == 0 || h_dst));
    const size_t need[2] = {h2d_bytes, d2h_bytes};
    for (int b = 0; b < 2; ++b)
        if (ctx->stage_bytes[b] < need[b]) {
            cudaFree(ctx->stage[b]); ctx->stage[b] = nullptr; ctx->stage_bytes[b] = 0;
            if (cudaMalloc(&ctx->stage[b], need[b]) != cudaSuccess) return fail(ctx, WIFI_ERR_NOMEM, "probe cudaMalloc(%zu) failed", need[b]);
            ctx->stage_bytes[b] = need[b];
        }
    CK(cudaStreamSynchronize(ctx->hstream[0])); CK(cudaStreamSynchronize(ctx->hstream[1]));
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    if (h2d_bytes) CK(cudaMemcpyAsync(ctx->stage[0], h_src, h2d_bytes, cudaMemcpyHostToDevice, ctx->hstream[0]));
    if (d2h_bytes) CK(cudaMemcpyAsync(h_dst, ctx->stage[1], d2h_bytes, cudaMemcpyDeviceToHost, ctx->hstream[1]));
    CK(cudaStreamSynchronize(ctx->hstream[0])); CK(cudaStreamSynchronize(ctx->hstream[1]));
    clock_gettime(CLOCK_MONOTONIC, &t1);
    *ms = 1e3 * (double)(t1.tv_sec - t0.tv_sec) + 1e-6 * (double)(t1.tv_nsec - t0.tv_nsec);
    return WIFI_OK;
}

int wifi_mmse_perframe_host(wifi_ctx *ctx, wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                            const void *sigma2, void *H, int64_t n, int flags)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (R && tx && rx && sigma2 && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    void *dR = nullptr;
    const size_t nR = esize(dt) * WIFI_NSC * WIFI_NSC;
    CK(cudaMalloc(&dR, nR));
    cudaError_t e = cudaMemcpy(dR, R, nR, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(dR); return fail(ctx, WIFI_ERR_CUDA, "H2D R: %s", cudaGetErrorString(e)); }
    const size_t row = WIFI_NSC * esize(dt), pitch = (size_t)frame_stride * esize(dt);
    int rc = host_pipeline(ctx, n, {in_arr(tx, row, pitch), in_arr(rx, row, pitch), in_arr(sigma2, rsize(dt), rsize(dt)), out_arr(H, row)},
                           [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                               return mmse_perframe(ctx, dt, dR, d[0], d[1], WIFI_NSC, d[2], nullptr, d[3], nc, flags, s, false);
                           });
    cudaFree(dR);
    return rc;
}

int wifi_mmse_perframe_eig_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2,
                                void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx && rx && sigma2 && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    const size_t row = WIFI_NSC * esize(dt), pitch = (size_t)frame_stride * esize(dt);
    return host_pipeline(ctx, n, {in_arr(tx, row, pitch), in_arr(rx, row, pitch), in_arr(sigma2, rsize(dt), rsize(dt)), out_arr(H, row)},
                         [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                             return mmse_eig(ctx, dt, d[0], d[1], WIFI_NSC, d[2], d[3], nc, s, s == ctx->hstream[1] ? 1 : 0);
                         });
}

int wifi_mmse_perframe_lowrank_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2,
                                    void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx && rx && sigma2 && H)) && (dt == WIFI_F32 || dt == WIFI_F64) && frame_stride >= WIFI_NSC);
    if (!ctx->lr_rank) return fail(ctx, WIFI_ERR_STATE, "no low-rank operands installed: call wifi_mmse_lowrank_prepare first");
    const size_t row = WIFI_NSC * esize(dt), pitch = (size_t)frame_stride * esize(dt);
    return host_pipeline(ctx, n, {in_arr(tx, row, pitch), in_arr(rx, row, pitch), in_arr(sigma2, rsize(dt), rsize(dt)), out_arr(H, row)},
                         [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                             return mmse_lowrank(ctx, dt, d[0], d[1], WIFI_NSC, d[2], d[3], nc, s);
                         });
}

int wifi_mmse_cconv_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx, const void *rx, const void *ow2, const void *H_ls, void *H,
                         int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx && rx && ow2 && H_ls && H)) && (dt == WIFI_F32 || dt == WIFI_F64));
    const size_t row = WIFI_NSC * esize(dt);
    return host_pipeline(ctx, n, {in_arr(tx, row, row), in_arr(rx, row, row), in_arr(ow2, rsize(dt), rsize(dt)), in_arr(H_ls, row, row), out_arr(H, row)},
                         [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                             return mmse_cconv(ctx, dt, 0, d[0], d[1], WIFI_NSC, d[2], d[3], d[4], nc, s);
                         });
}

int wifi_mmse_matlab_host(wifi_ctx *ctx, wifi_dtype dt, const void *tx_frames, const void *rx_frames, const void *ow2, const void *H_ls,
                          void *H, int64_t n)
{
    ENTER();
    NEED(n >= 0 && (n == 0 || (tx_frames && rx_frames && ow2 && H_ls && H)) && (dt == WIFI_F32 || dt == WIFI_F64));
    const size_t row = WIFI_NSC * esize(dt), frow = WIFI_FRAME * esize(dt);
    return host_pipeline(ctx, n, {in_arr(tx_frames, 4 * row, frow), in_arr(rx_frames, 4 * row, frow), in_arr(ow2, rsize(dt), rsize(dt)), in_arr(H_ls, row, row), out_arr(H, row)},
                         [&](std::vector<void *> &d, int64_t nc, int64_t, cudaStream_t s) {
                             return mmse_cconv(ctx, dt, 1, d[0], d[1], 4 * WIFI_NSC, d[2], d[3], d[4], nc, s);
                         });
}

int wifi_cmatmul_host(wifi_ctx *ctx, wifi_dtype dt, const void *A, int r1, int c1, const void *B, int r2, int c2, void *C, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(r1, c1) && order_ok(r2, c2));
    if (c1 != r2) return fail(ctx, WIFI_ERR_INVALID, "Matrices dimension missmatch");
    if (batch == 0) return WIFI_OK;
    NEED(A && B && C);
    size_t na = esize(dt) * r1 * c1 * batch, nb = esize(dt) * r2 * c2 * batch, nc = esize(dt) * r1 * c2 * batch;
    size_t oa = 0, ob = (na + 255) / 256 * 256, oc = ob + (nb + 255) / 256 * 256;
    UTIL_STAGE(oc + nc);
    char *base = (char *)ctx->stage[0];
    CK(cudaMemcpyAsync(base + oa, A, na, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(base + ob, B, nb, cudaMemcpyHostToDevice, ctx->stream));
    int rc = wifi_cmatmul_batch(ctx, dt, base + oa, r1, c1, base + ob, r2, c2, base + oc, batch);
    if (rc) return rc;
    CK(cudaMemcpyAsync(C, base + oc, nc, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return WIFI_OK;
}

int wifi_chermitian_host(wifi_ctx *ctx, wifi_dtype dt, int mode, const void *M, int row, int col, void *res, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(row, col) && (mode == WIFI_AS_WRITTEN || mode == WIFI_INTENDED));
    if (batch == 0 || row * col == 0) return WIFI_OK;
    NEED(M && res);
    size_t n = esize(dt) * row * col * batch, o = (n + 255) / 256 * 256;
    UTIL_STAGE(o + n);
    char *base = (char *)ctx->stage[0];
    CK(cudaMemcpyAsync(base, M, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = wifi_chermitian_batch(ctx, dt, mode, base, row, col, base + o, batch);
    if (rc) return rc;
    CK(cudaMemcpyAsync(res, base + o, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return WIFI_OK;
}

int wifi_cadd_host(wifi_ctx *ctx, wifi_dtype dt, int mode, const void *M1, int r1, int c1, const void *M2, int r2, int c2, void *res, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(r1, c1) && (mode == WIFI_AS_WRITTEN || mode == WIFI_INTENDED));
    if (r1 != r2 || c1 != c2) return fail(ctx, WIFI_ERR_INVALID, "Matrices dimension missmatch");
    if (batch == 0 || r1 * c1 == 0) return WIFI_OK;
    NEED(M1 && res && (mode == WIFI_AS_WRITTEN || M2));
    size_t n = esize(dt) * r1 * c1 * batch, o = (n + 255) / 256 * 256;
    UTIL_STAGE(3 * o);
    char *base = (char *)ctx->stage[0];
    CK(cudaMemcpyAsync(base, M1, n, cudaMemcpyHostToDevice, ctx->stream));
    if (M2) CK(cudaMemcpyAsync(base + o, M2, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = wifi_cadd_batch(ctx, dt, mode, base, r1, c1, base + o, r2, c2, base + 2 * o, batch);
    if (rc) return rc;
    CK(cudaMemcpyAsync(res, base + 2 * o, n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return WIFI_OK;
}

int wifi_couter_host(wifi_ctx *ctx, wifi_dtype dt, const void *M1, int r1, int c1, const void *M2, int r2, int c2, void *res, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(r1, c1) && order_ok(r2, c2));
    if (c1 != r2) return fail(ctx, WIFI_ERR_INVALID, "Matrices dimension missmatch");
    if (batch == 0 || r1 * c2 == 0) return WIFI_OK;
    NEED(M1 && M2 && res);
    size_t na = esize(dt) * r1 * c1 * batch, nb = esize(dt) * r2 * c2 * batch, nc = esize(dt) * r1 * c2 * batch;
    size_t ob = (na + 255) / 256 * 256, oc = ob + (nb + 255) / 256 * 256;
    UTIL_STAGE(oc + nc);
    char *base = (char *)ctx->stage[0];
    CK(cudaMemcpyAsync(base, M1, na, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(base + ob, M2, nb, cudaMemcpyHostToDevice, ctx->stream));
    int rc = wifi_couter_batch(ctx, dt, base, r1, c1, base + ob, r2, c2, base + oc, batch);
    if (rc) return rc;
    CK(cudaMemcpyAsync(res, base + oc, nc, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return WIFI_OK;
}

int wifi_cidentity_host(wifi_ctx *ctx, wifi_dtype dt, void *Id, int size, double scalar, int64_t batch)
{
    ENTER();
    NEED(batch >= 0 && order_ok(size, size));
    if (batch == 0 || size == 0) return WIFI_OK;
    NEED(Id);
    size_t n = esize(dt) * size * size * batch;
    UTIL_STAGE(n);
    int rc = wifi_cidentity_batch(ctx, dt, ctx->stage[0], size, scalar, batch);
    if (rc) return rc;
    CK(cudaMemcpyAsync(Id, ctx->stage[0], n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return WIFI_OK;
}

int wifi_cinverse_host(wifi_ctx *ctx, wifi_dtype dt, const void *A, int order, void *Y, int64_t batch, int *info_host)
{
    ENTER();
    NEED(batch >= 0 && order >= 1 && order <= WIFI_MAX_ORDER);
    if (batch == 0) return WIFI_OK;
    NEED(A && Y);
    size_t n = esize(dt) * order * order * batch, o = (n + 255) / 256 * 256;
    size_t oi = 2 * o;
    UTIL_STAGE(oi + sizeof(int) * batch);
    char *base = (char *)ctx->stage[0];
    CK(cudaMemcpyAsync(base, A, n, cudaMemcpyHostToDevice, ctx->stream));
    int rc = wifi_cinverse_batch(ctx, dt, base, order, base + o, batch, (int *)(base + oi));
    if (rc) return rc;
    CK(cudaMemcpyAsync(Y, base + o, n, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<int> info(batch);
    CK(cudaMemcpyAsync(info.data(), base + oi, sizeof(int) * batch, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    int any = 0;
    for (int64_t b = 0; b < batch; ++b) { if (info_host) info_host[b] = info[b]; any |= info[b]; }
    return any ? fail(ctx, WIFI_ERR_SINGULAR, "singular matrix in batch") : WIFI_OK;
}

// ---- default context for the single-frame drop-ins ------------------------------------------------------
wifi_ctx *wifi_default_ctx(void)
{
    static std::mutex mu;
    static wifi_ctx *g = nullptr;
    std::lock_guard<std::mutex> lk(mu);
    if (!g) {
        const char *env = getenv("WIFI_B200_DEVICE");
        int dev = env ? atoi(env) : 0;
        int rc = wifi_create(dev, &g);
        if (rc != WIFI_OK) {
            fprintf(stderr, "wifi_b200: no usable CUDA device (wifi_create(%d) -> %d); there is no CPU fallback\n", dev, rc);
            abort();
        }
    }
    return g;
}

}  // extern "C"
