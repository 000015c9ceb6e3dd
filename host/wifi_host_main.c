/*
 * wifi_host_main.c -- C host driver over the C-ABI (replaces the reference's main.c / main_openmp.c / main_mpi.c drivers).
 *
 *   part 1  mirrors main.c:10-64 on the inputs.h frame (tests/golden/inputs_h_frame.f64: the inputs.h globals as raw
 *           doubles): block 0 is extracted (main.c:30-33) and all five estimators are called through the reference's own
 *           entry points (include/wifi_dropin.h), printing H_EST like main.c:42-44.
 *   part 2  frame-sharded batch over every visible GPU: one pthread + one wifi_ctx per device, contiguous shard
 *           [N g/G, N (g+1)/G) of the global synthetic sequence, no exchange on the estimation path; the per-shard error
 *           statistics (4 doubles per device) are summed on the host.
 *
 *   part 3  (--file in out) frame-file pipeline: a WIFI_FILE_FREQ or WIFI_FILE_TIME file (include/wifi_frame_file.h) is
 *           read in chunks into pinned buffers, copied to the device, run through the front-end (TIME files), all five
 *           estimators (PS_MMSE in main.c:148's calling convention, R_f = H_lt H_lt^H) and the equalizer with the
 *           intermediates staying in HBM, and the results are written as a WIFI_FILE_EST file.
 *
 *   usage: wifi_host_main [frames_total=4194304] [fixture=tests/golden/inputs_h_frame.f64]
 *          wifi_host_main --file frames.bin estimates.bin
 */
#include <complex.h>
#include <cuda_runtime.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "wifi_b200.h"
#include "wifi_dropin.h"
#include "wifi_frame_file.h"

#define NSC WIFI_NSC

static double now_s(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }

static int part1(const char *fixture)
{
    FILE *f = fopen(fixture, "rb");
    if (!f) { fprintf(stderr, "cannot open %s\n", fixture); return 1; }
    double ow2, buf[2 * (2 * NSC + 2 * WIFI_FRAME)];
    if (fread(&ow2, sizeof ow2, 1, f) != 1 || fread(buf, sizeof(double), sizeof buf / sizeof(double), f) != sizeof buf / sizeof(double)) {
        fprintf(stderr, "short read on %s\n", fixture); fclose(f); return 1;
    }
    fclose(f);
    long double complex tx_pre[NSC], rx_pre[NSC], tx_vec[NSC], rx_vec[NSC];
    long double complex H_lt[NSC], H_lin[NSC], H_cub[NSC], H_sinc[NSC], H_mmse[NSC];
    const double *p = buf;
    for (int k = 0; k < NSC; ++k) tx_pre[k] = p[2 * k] + p[2 * k + 1] * I;
    p += 2 * NSC;
    for (int k = 0; k < NSC; ++k) rx_pre[k] = p[2 * k] + p[2 * k + 1] * I;
    p += 2 * NSC;
    const int OFDM_block = 0;                                               /* main.c:16 */
    for (int r = 0; r < NSC; ++r) {                                         /* main.c:30-33 */
        tx_vec[r] = p[2 * (NSC * OFDM_block + r)] + p[2 * (NSC * OFDM_block + r) + 1] * I;
        rx_vec[r] = p[2 * WIFI_FRAME + 2 * (NSC * OFDM_block + r)] + p[2 * WIFI_FRAME + 2 * (NSC * OFDM_block + r) + 1] * I;
    }
    printf("**** Processing Block %d\n", OFDM_block);
    WiFi_channel_estimation_LT_LS(tx_pre, rx_pre, H_lt);
    WiFi_channel_estimation_PS_Linear(tx_vec, rx_vec, H_lin);
    WiFi_channel_estimation_PS_Cubic(tx_vec, rx_vec, H_cub);
    WiFi_channel_estimation_PS_Sinc(tx_vec, rx_vec, H_sinc);
    WiFi_channel_estimation_PS_MMSE(tx_vec, rx_vec, NULL, ow2, H_lt, H_mmse);
    for (int i = 0; i < NSC; i += 13)
        printf("H_EST[%2d] LT_LS % .12f%+.12fi  Linear % .12f%+.12fi  Cubic % .12f%+.12fi  Sinc % .12f%+.12fi  MMSE % .12f%+.12fi\n", i,
               (double)creall(H_lt[i]), (double)cimagl(H_lt[i]), (double)creall(H_lin[i]), (double)cimagl(H_lin[i]),
               (double)creall(H_cub[i]), (double)cimagl(H_cub[i]), (double)creall(H_sinc[i]), (double)cimagl(H_sinc[i]),
               (double)creall(H_mmse[i]), (double)cimagl(H_mmse[i]));
    /* SURVEY App. C known answers */
    double e1 = cabsl(H_lin[52] - (0.00051507313233677404L - 0.012534787556476966L * I));
    double e2 = cabsl(H_mmse[0] - (0.0090896514477585877939L + 0.00092809776449416434917L * I));
    printf("check vs reference known answers: |dLinear[52]| = %.2e, |dMMSE[0]| = %.2e -> %s\n", e1, e2, (e1 < 1e-12 && e2 < 1e-12) ? "OK" : "MISMATCH");
    return !(e1 < 1e-12 && e2 < 1e-12);
}

typedef struct { int dev, ndev; long n_total; double stats[4]; double seconds; int rc; } shard_t;

static void *shard_main(void *arg)
{
    shard_t *s = (shard_t *)arg;
    const long lo = s->n_total * s->dev / s->ndev, hi = s->n_total * (s->dev + 1) / s->ndev, n = hi - lo;
    wifi_ctx *ctx = NULL;
    s->rc = wifi_create(s->dev, &ctx);
    if (s->rc) return NULL;
    cudaSetDevice(s->dev);
    float2 *tx, *rx, *Ht, *H; double2 *R; double *d, *stats;
    size_t vb = (size_t)n * NSC * sizeof(float2), fb = (size_t)n * WIFI_FRAME * sizeof(float2);
    cudaMalloc((void **)&tx, fb); cudaMalloc((void **)&rx, fb); cudaMalloc((void **)&Ht, vb); cudaMalloc((void **)&H, vb);
    cudaMalloc((void **)&R, NSC * NSC * sizeof(double2)); cudaMalloc((void **)&d, NSC * sizeof(double)); cudaMalloc((void **)&stats, 4 * sizeof(double));
    cudaMemset(stats, 0, 4 * sizeof(double));
    double hd[NSC];
    for (int k = 0; k < NSC; ++k) hd[k] = 9.6172e-08 / (k == WIFI_DC ? 1e-8 : 8.875 * 8.875);
    cudaMemcpy(d, hd, sizeof hd, cudaMemcpyHostToDevice);
    /* this shard of the global sequence, generated on the device: no data exchange */
    s->rc = wifi_synth_frames(ctx, WIFI_F32, 0x80211, lo, n, 0, NULL, NULL, tx, rx, Ht, NULL);
    if (!s->rc) s->rc = wifi_synth_covariance(ctx, R);
    if (!s->rc) s->rc = wifi_mmse_filter_form(ctx, R, d, NULL);
    if (!s->rc) s->rc = wifi_mmse_shared_batch(ctx, WIFI_F32, tx, rx, WIFI_FRAME, H, n);      /* warm-up */
    wifi_synchronize(ctx);
    double t0 = now_s();
    for (int it = 0; it < 10 && !s->rc; ++it) s->rc = wifi_mmse_shared_batch(ctx, WIFI_F32, tx, rx, WIFI_FRAME, H, n);
    wifi_synchronize(ctx);
    s->seconds = (now_s() - t0) / 10;
    if (!s->rc) s->rc = wifi_error_stats(ctx, WIFI_F32, H, Ht, n * NSC, stats);
    wifi_synchronize(ctx);
    cudaMemcpy(s->stats, stats, sizeof s->stats, cudaMemcpyDeviceToHost);
    if (s->rc) fprintf(stderr, "device %d: %s\n", s->dev, wifi_last_error(ctx));
    cudaFree(tx); cudaFree(rx); cudaFree(Ht); cudaFree(H); cudaFree(R); cudaFree(d); cudaFree(stats);
    wifi_destroy(ctx);
    return NULL;
}

/* ---- part 3: frame-file pipeline ------------------------------------------------------------------------ */
#define FCK(x) do { int rc_ = (x); if (rc_) { fprintf(stderr, "%s:%d: %s -> %d (%s)\n", __FILE__, __LINE__, #x, rc_, ctx ? wifi_last_error(ctx) : ""); return 1; } } while (0)
#define CCK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d: %s -> %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return 1; } } while (0)

static int file_pipeline(const char *in_path, const char *out_path)
{
    wifi_ctx *ctx = NULL;
    FILE *fi = fopen(in_path, "rb");
    if (!fi) { fprintf(stderr, "cannot open %s\n", in_path); return 1; }
    wifi_file_header h;
    if (fread(&h, sizeof h, 1, fi) != 1 || memcmp(h.magic, WIFI_FILE_MAGIC, 8) || h.dtype > 1 || h.kind > WIFI_FILE_TIME) {
        fprintf(stderr, "%s: not a FREQ/TIME frame file\n", in_path); fclose(fi); return 1;
    }
    const wifi_dtype dt = (wifi_dtype)h.dtype;
    const size_t es = dt == WIFI_F32 ? 8 : 16;                       /* bytes per complex value */
    const long n = (long)h.n_frames;
    /* input planes (values per frame) and where each starts in the file */
    const size_t in_w[4] = {h.kind == WIFI_FILE_FREQ ? NSC : WIFI_PACKET, h.kind == WIFI_FILE_FREQ ? NSC : WIFI_PACKET,
                            h.kind == WIFI_FILE_FREQ ? WIFI_FRAME : WIFI_LPTOT, h.kind == WIFI_FILE_FREQ ? WIFI_FRAME : WIFI_LPTOT};
    /* FREQ: tx_pre rx_pre tx_symb rx_symb;  TIME: tx_packet rx_packet tx_lptot rx_lptot */
    size_t in_off[4], o = sizeof h;
    for (int i = 0; i < 4; ++i) { in_off[i] = o; o += (size_t)n * in_w[i] * es; }
    /* output planes */
    const size_t out_w[6] = {NSC, NSC, NSC, NSC, NSC, WIFI_FRAME};
    size_t out_off[7]; o = sizeof h;
    for (int i = 0; i < 6; ++i) { out_off[i] = o; o += (size_t)n * out_w[i] * es; }
    out_off[6] = o;                                                 /* ow2 [n] real */
    FILE *fo = fopen(out_path, "wb");
    if (!fo) { fprintf(stderr, "cannot create %s\n", out_path); fclose(fi); return 1; }
    wifi_file_header ho = h; ho.kind = WIFI_FILE_EST;
    fwrite(&ho, sizeof ho, 1, fo);

    FCK(wifi_create(0, &ctx));
    const long chunk = n < 32768 ? (n > 0 ? n : 1) : 32768;
    void *hin[4], *hout[7], *din[4], *dsymb[2], *dpre[2], *dow2, *dH[5], *deq;
    for (int i = 0; i < 4; ++i) { FCK(wifi_host_alloc(&hin[i], chunk * in_w[i] * es)); CCK(cudaMalloc(&din[i], chunk * in_w[i] * es)); }
    for (int i = 0; i < 6; ++i) FCK(wifi_host_alloc(&hout[i], chunk * out_w[i] * es));
    FCK(wifi_host_alloc(&hout[6], chunk * es / 2));
    for (int i = 0; i < 2; ++i) { CCK(cudaMalloc(&dsymb[i], chunk * WIFI_FRAME * es)); CCK(cudaMalloc(&dpre[i], chunk * NSC * es)); }
    for (int i = 0; i < 5; ++i) CCK(cudaMalloc(&dH[i], chunk * NSC * es));
    CCK(cudaMalloc(&deq, chunk * WIFI_FRAME * es)); CCK(cudaMalloc(&dow2, chunk * es / 2));
    double t0 = now_s();
    for (long f0 = 0; f0 < n; f0 += chunk) {
        const long nc = n - f0 < chunk ? n - f0 : chunk;
        for (int i = 0; i < 4; ++i) {
            fseek(fi, (long)(in_off[i] + (size_t)f0 * in_w[i] * es), SEEK_SET);
            if (fread(hin[i], in_w[i] * es, nc, fi) != (size_t)nc) { fprintf(stderr, "short read\n"); return 1; }
            CCK(cudaMemcpyAsync(din[i], hin[i], nc * in_w[i] * es, cudaMemcpyHostToDevice, 0));
        }
        const void *tx_pre, *rx_pre, *tx_symb, *rx_symb;
        if (h.kind == WIFI_FILE_TIME) {                              /* time samples -> symbols, preamble spectra, noise estimate */
            FCK(wifi_frontend_batch(ctx, dt, din[0], din[2], dsymb[0], dpre[0], NULL, nc));
            FCK(wifi_frontend_batch(ctx, dt, din[1], din[3], dsymb[1], dpre[1], dow2, nc));
            tx_pre = dpre[0]; rx_pre = dpre[1]; tx_symb = dsymb[0]; rx_symb = dsymb[1];
        } else {
            tx_pre = din[0]; rx_pre = din[1]; tx_symb = din[2]; rx_symb = din[3];
            /* inputs.h carries ow2 as a constant (inputs.h:18); FREQ files use it for every frame */
            if (dt == WIFI_F32) { float *w = (float *)hout[6]; for (long i = 0; i < nc; ++i) w[i] = 9.6172e-08f; }
            else { double *w = (double *)hout[6]; for (long i = 0; i < nc; ++i) w[i] = 9.6172e-08; }
            CCK(cudaMemcpyAsync(dow2, hout[6], nc * es / 2, cudaMemcpyHostToDevice, 0));
        }
        FCK(wifi_lt_ls_batch(ctx, dt, tx_pre, rx_pre, dH[0], nc));
        FCK(wifi_ps_batch(ctx, dt, WIFI_PS_LINEAR | WIFI_PS_CUBIC | WIFI_PS_SINC, tx_symb, rx_symb, WIFI_FRAME, dH[1], dH[2], dH[3], nc));
        /* main.c:148 calling convention on block 0 (main.c:30-33): frame_stride is fixed at 53 there, so gather block 0 */
        CCK(cudaMemcpy2DAsync(dsymb[0] == tx_symb ? deq : dsymb[0], NSC * es, tx_symb, WIFI_FRAME * es, NSC * es, nc, cudaMemcpyDeviceToDevice, 0));
        void *tx0 = dsymb[0] == tx_symb ? deq : dsymb[0];
        void *rx0 = (char *)tx0 + (size_t)nc * NSC * es;
        CCK(cudaMemcpy2DAsync(rx0, NSC * es, rx_symb, WIFI_FRAME * es, NSC * es, nc, cudaMemcpyDeviceToDevice, 0));
        FCK(wifi_mmse_cconv_batch(ctx, dt, tx0, rx0, dow2, dH[0], dH[4], nc));
        FCK(wifi_synchronize(ctx));
        FCK(wifi_equalize_batch(ctx, dt, rx_symb, dH[0], dH[1], deq, nc));
        for (int i = 0; i < 5; ++i) CCK(cudaMemcpyAsync(hout[i], dH[i], nc * NSC * es, cudaMemcpyDeviceToHost, 0));
        CCK(cudaMemcpyAsync(hout[5], deq, nc * WIFI_FRAME * es, cudaMemcpyDeviceToHost, 0));
        CCK(cudaMemcpyAsync(hout[6], dow2, nc * es / 2, cudaMemcpyDeviceToHost, 0));
        CCK(cudaDeviceSynchronize());
        for (int i = 0; i < 6; ++i) { fseek(fo, (long)(out_off[i] + (size_t)f0 * out_w[i] * es), SEEK_SET); fwrite(hout[i], out_w[i] * es, nc, fo); }
        fseek(fo, (long)(out_off[6] + (size_t)f0 * es / 2), SEEK_SET); fwrite(hout[6], es / 2, nc, fo);
    }
    double dt_s = now_s() - t0;
    fclose(fi); fclose(fo);
    printf("frame file %s (%s, %s, %ld frames) -> %s: %.3f s = %.3e frames/s incl. file I/O\n", in_path, h.kind == WIFI_FILE_TIME ? "time samples" : "frequency domain",
           dt == WIFI_F32 ? "complex64" : "complex128", n, out_path, dt_s, n / dt_s);
    wifi_destroy(ctx);
    return 0;
}

int main(int argc, char **argv)
{
    if (argc == 4 && !strcmp(argv[1], "--file")) return file_pipeline(argv[2], argv[3]);
    long n_total = argc > 1 ? atol(argv[1]) : 4194304L;
    const char *fixture = argc > 2 ? argv[2] : "tests/golden/inputs_h_frame.f64";
    printf("%s\n", wifi_version());
    int bad = part1(fixture);
    int ndev = 0;
    cudaGetDeviceCount(&ndev);
    if (ndev > 8) ndev = 8;
    shard_t sh[8];
    pthread_t th[8];
    for (int g = 0; g < ndev; ++g) { memset(&sh[g], 0, sizeof sh[g]); sh[g].dev = g; sh[g].ndev = ndev; sh[g].n_total = n_total; pthread_create(&th[g], NULL, shard_main, &sh[g]); }
    double tmax = 0, st[4] = {0, 0, 0, 0};
    for (int g = 0; g < ndev; ++g) {
        pthread_join(th[g], NULL);
        bad |= sh[g].rc;
        if (sh[g].seconds > tmax) tmax = sh[g].seconds;
        st[0] += sh[g].stats[0]; st[1] += sh[g].stats[1]; st[2] += sh[g].stats[2];
        if (sh[g].stats[3] > st[3]) st[3] = sh[g].stats[3];
    }
    if (ndev)
        printf("shared-filter PS_MMSE, %ld frames sharded over %d GPU(s): %.3f ms per pass (max over devices) = %.3e frames/s; NMSE vs true channel %.3e over %.0f values\n",
               n_total, ndev, 1e3 * tmax, n_total / tmax, st[0] / st[1], st[2]);
    return bad ? 1 : 0;
}
