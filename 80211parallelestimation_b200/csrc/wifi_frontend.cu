// wifi_frontend.cu -- the receiver front-end that produces the estimators' inputs:
//
//   symb[f][b][i]  = X_b[(i - 26) mod 64], i < 53, X_b = 64-point DFT of OFDM block b of the packet without its 16-sample
//                    cyclic prefix                                     (WiFi_blocks_extraction.m:5-10)
//   pre_fft[f][i]  = the same transform of the averaged long-training symbols (p1 + p2)/2,
//                    p1 = lptot[96..159], p2 = lptot[32..95]           (WiFi_RX.m:19-23 / 25-29)
//   ow2[f]         = sum |p2 - p1|^2 / (2*64)                          (WiFi_RX.m:31)
//
// HBM-bound: 1 088 complex values in (the cyclic prefixes and the first 32 samples of lptot are never fetched), 848 out
// per frame and side.  16 transforms per frame; a warp takes 4 at a time (8 lanes per transform) and computes the
// 64-point DFT as 8 x 8 Cooley-Tukey with both radix-8 passes in registers:
//   n = j + 8 m, k = p + 8 c:   X[p + 8c] = sum_j w8^(jc) ( w64^(jp) sum_m x[j + 8m] w8^(mp) )
//   A. coalesced 16-byte loads of the 4 x 64 samples into the warp's shared-memory tile (preamble: averaged on the fly);
//   B. lane (t, j) reads x[j + 8m], 8-point DFT over m, multiplies by its 7 twiddles w64^(jp) (registers);
//   C. transpose through the tile; lane (t, p) runs the 8-point DFT over j -> X[p + 8c];
//   D. circshift 26 / keep 53 while scattering into the tile, then the 4 x 53 outputs -- contiguous in symb -- leave as
//      one coalesced run.
// ~60 instructions per transform and warp against a 42-cycle HBM budget per transform and SM (FP32).
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

constexpr int FE_THREADS = 256;
constexpr int FE_WARPS = FE_THREADS / 32;
constexpr int FE_TS = 72;                   // tile stride per transform in complex values (64 + 8: the 4 transforms of a
                                            // warp start in different bank groups)
constexpr int FE_PKT = 1200, FE_BLK = 80, FE_CP = 16, FE_LP = 160;

template <typename T> __device__ __forceinline__ cx<T> cmul_mi(cx<T> a) { return mk<T>(a.y, -a.x); }            // a * (-i)

// in-place forward 8-point DFT, natural order in and out:  v[p] <- sum_m v[m] exp(-2 pi i m p / 8)
template <typename T> __device__ __forceinline__ void dft8(cx<T> (&v)[8])
{
    const T h = (T)0.70710678118654752440;
    // stage 1 (span 4), twiddles w8^i
    cx<T> a0 = cadd(v[0], v[4]), b0 = csub(v[0], v[4]);
    cx<T> a1 = cadd(v[1], v[5]), b1 = csub(v[1], v[5]);
    cx<T> a2 = cadd(v[2], v[6]), b2 = csub(v[2], v[6]);
    cx<T> a3 = cadd(v[3], v[7]), b3 = csub(v[3], v[7]);
    b1 = mk<T>(h * (b1.x + b1.y), h * (b1.y - b1.x));           // * (1 - i)/sqrt2
    b2 = cmul_mi<T>(b2);                                        // * -i
    b3 = mk<T>(h * (b3.y - b3.x), -h * (b3.x + b3.y));          // * (-1 - i)/sqrt2
    // stage 2 (span 2), twiddles 1, -i
    cx<T> c0 = cadd(a0, a2), c2 = csub(a0, a2), c1 = cadd(a1, a3), c3 = cmul_mi<T>(csub(a1, a3));
    cx<T> d0 = cadd(b0, b2), d2 = csub(b0, b2), d1 = cadd(b1, b3), d3 = cmul_mi<T>(csub(b1, b3));
    // stage 3 (span 1) + bit reversal
    v[0] = cadd(c0, c1); v[4] = csub(c0, c1);
    v[2] = cadd(c2, c3); v[6] = csub(c2, c3);
    v[1] = cadd(d0, d1); v[5] = csub(d0, d1);
    v[3] = cadd(d2, d3); v[7] = csub(d2, d3);
}

// two consecutive complex values from a 16-byte aligned address
__device__ __forceinline__ void ld_pair(const float2 *p, float2 &a, float2 &b)
{
    const float4 v = ld_stream(reinterpret_cast<const float4 *>(p));
    a = make_float2(v.x, v.y); b = make_float2(v.z, v.w);
}
__device__ __forceinline__ void ld_pair(const double2 *p, double2 &a, double2 &b) { a = ld_stream(p); b = ld_stream(p + 1); }

template <typename T> __device__ __forceinline__ void sincospi_t(T x, T *s, T *c);
template <> __device__ __forceinline__ void sincospi_t<float>(float x, float *s, float *c) { sincospif(x, s, c); }
template <> __device__ __forceinline__ void sincospi_t<double>(double x, double *s, double *c) { sincospi(x, s, c); }

// Steps B-D for the 4 transforms a warp holds in its tile (natural-order samples at tile[t * FE_TS + n]); on return the 4 x 53
// kept bins are packed back to back, tile[t * 53 + i] = X_t[(i - 26) mod 64], and visible to the whole warp.
template <typename T> __device__ __forceinline__ void fft4_rows(cx<T> *tile, const int t, const int j, const cx<T> (&tw)[8])
{
    // ---- B. first radix-8 pass over m, twiddle ----
    cx<T> v[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) v[m] = tile[t * FE_TS + j + 8 * m];
    dft8<T>(v);
#pragma unroll
    for (int p = 1; p < 8; ++p) v[p] = cmul(v[p], tw[p]);
    __syncwarp();
    // ---- C. transpose: tile[t][p][j] with rows of 9 (a lane then reads 8 consecutive values at a 72-byte lane
    //         stride: conflict-free; rows of 8 would be a 4-way conflict), second pass over j on lane (t, p = j) ----
#pragma unroll
    for (int p = 0; p < 8; ++p) tile[t * FE_TS + p * 9 + j] = v[p];
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) v[jj] = tile[t * FE_TS + j * 9 + jj];      // this lane's p is its j
    dft8<T>(v);                                                                // v[c] = X[p + 8 c]
    __syncwarp();
    // ---- D. circshift 26, keep 53: X[k] -> y[(k + 26) mod 64]; rows of 53 packed back to back in the tile ----
#pragma unroll
    for (int c = 0; c < 8; ++c) {
        const int i = (j + 8 * c + 26) & 63;
        if (i < NSC) tile[t * NSC + i] = v[c];
    }
    __syncwarp();
}

// PREF (FP32): the samples of the warp's NEXT group are requested into registers before the transforms of the current one (and
// the group after that is pulled into L2) -- the kernel is bound by load latency (ncu: long_scoreboard 9 warps per issue); in
// FP64 the 40 extra registers cost a CTA per SM, so there only the L2 prefetch is used.
template <typename T, bool PREF>
__global__ void __launch_bounds__(FE_THREADS) frontend_kernel(const cx<T> *__restrict__ packet, const cx<T> *__restrict__ lptot,
                                                              cx<T> *__restrict__ symb, cx<T> *__restrict__ pre_fft, T *__restrict__ ow2,
                                                              int64_t n_frames)
{
    __shared__ __align__(16) cx<T> tile_all[FE_WARPS][4 * FE_TS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    cx<T> *tile = tile_all[warp];
    const int t = lane >> 3, j = lane & 7;                      // transform within the group / radix-8 index
    // twiddles w64^(j p), p = 1..7
    cx<T> tw[8];
#pragma unroll
    for (int p = 1; p < 8; ++p) {
        T s, c;
        sincospi_t<T>((T)(j * p) * (T)(-1.0 / 32.0), &s, &c);
        tw[p] = mk<T>(c, s);
    }
    const int64_t n_groups = n_frames * 4;                       // 16 transforms per frame, 4 per group
    const int64_t gstride = (int64_t)gridDim.x * FE_WARPS;
    const int n2 = 2 * lane;                                     // this lane's two consecutive samples of every transform
    // the raw samples of one group: transform i of the group = OFDM block 4 q + i, or (q = 3, i = 3) the two long-training symbols
    auto fetch = [&](int64_t g, cx<T> (&raw)[5][2]) {
        const int64_t f = g >> 2;
        const int q = (int)(g & 3);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if (4 * q + i < NBLK) ld_pair(packet + f * FE_PKT + (4 * q + i) * FE_BLK + FE_CP + n2, raw[i][0], raw[i][1]);
            else {
                ld_pair(lptot + f * FE_LP + 96 + n2, raw[3][0], raw[3][1]);
                ld_pair(lptot + f * FE_LP + 32 + n2, raw[4][0], raw[4][1]);
            }
        }
    };
    cx<T> raw[5][2];
    int64_t g = (int64_t)blockIdx.x * FE_WARPS + warp;
    if (PREF && g < n_groups) fetch(g, raw);
    for (; g < n_groups; g += gstride) {
        const int64_t f = g >> 2;
        const int q = (int)(g & 3);                              // transforms 4q .. 4q+3 of frame f; transform 15 = preamble
        // HBM -> L2 for a LATER group of this warp (no registers, no shared memory): its loads then cost an L2 latency.
        // lanes 0..3: one 64-sample block each; the preamble of a frame's last group: p1 and p2 (lanes 3, 4)
        {
            const int64_t gn = g + (PREF ? 2 : 1) * gstride;
            if (gn < n_groups && lane < 5) {
                const int64_t fn = gn >> 2;
                const int bn = 4 * (int)(gn & 3) + lane;
                const cx<T> *src = nullptr;
                if (lane < 4 && bn < NBLK) src = packet + fn * FE_PKT + bn * FE_BLK + FE_CP;
                else if ((gn & 3) == 3 && lane >= 3) src = lptot + fn * FE_LP + (lane == 3 ? 96 : 32);
                if (src) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(64 * sizeof(cx<T>))) : "memory");
            }
        }
        // ---- A. samples -> tile (natural order), two consecutive samples per lane and transform ----
        if (!PREF) fetch(g, raw);
        T nv = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            cx<T> x0 = raw[i][0], x1 = raw[i][1];
            if (4 * q + i >= NBLK) {                             // (p1 + p2) / 2 and the noise estimate, WiFi_RX.m:19-31
                const cx<T> p10 = raw[3][0], p11 = raw[3][1], p20 = raw[4][0], p21 = raw[4][1];
                x0 = mk<T>((p10.x + p20.x) * (T)0.5, (p10.y + p20.y) * (T)0.5);
                x1 = mk<T>((p11.x + p21.x) * (T)0.5, (p11.y + p21.y) * (T)0.5);
                const T d0x = p20.x - p10.x, d0y = p20.y - p10.y, d1x = p21.x - p11.x, d1y = p21.y - p11.y;
                nv += d0x * d0x + d0y * d0y + d1x * d1x + d1y * d1y;
            }
            tile[i * FE_TS + n2] = x0;
            tile[i * FE_TS + n2 + 1] = x1;
        }
        if (q == 3 && ow2) {                                     // warp-uniform
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
            if (lane == 0) ow2[f] = nv * (T)(1.0 / 128.0);
        }
        __syncwarp();
        if (PREF && g + gstride < n_groups) fetch(g + gstride, raw);
        fft4_rows<T>(tile, t, j, tw);                             // steps B-D
        const int n_data = q < 3 ? 4 : 3;                        // data rows in this group (group 3 ends with the preamble)
        cx<T> *dst = symb + (f * NBLK + 4 * q) * NSC;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            const int o = i * 32 + lane;
            if (o < n_data * NSC) st_stream(dst + o, tile[o]);
            else if (o < 4 * NSC) st_stream(pre_fft + f * NSC + (o - 3 * NSC), tile[o]);
        }
        __syncwarp();
    }
}

cudaError_t launch_frontend(wifi_dtype dt, const void *packet, const void *lptot, void *symb, void *pre_fft, void *ow2,
                            int64_t n_frames, cudaStream_t s)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    const int64_t need = (n_frames * 4 + FE_WARPS - 1) / FE_WARPS;
    const unsigned grid = (unsigned)std::min<int64_t>(need, 148 * 8);
    if (dt == WIFI_F32)
        frontend_kernel<float, true><<<grid, FE_THREADS, 0, s>>>((const float2 *)packet, (const float2 *)lptot, (float2 *)symb, (float2 *)pre_fft,
                                                           (float *)ow2, n_frames);
    else
        frontend_kernel<double, false><<<grid, FE_THREADS, 0, s>>>((const double2 *)packet, (const double2 *)lptot, (double2 *)symb,
                                                            (double2 *)pre_fft, (double *)ow2, n_frames);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Fused receiver chain: time samples -> all estimates (+ equalized symbols) in ONE pass (WiFi_RX.m:17-60 with
// the C estimators of main.c:66-146).  The symbols the front-end produces never touch HBM unless the caller asks for them:
//     in   rx_packet (15 x 64 samples, cyclic prefixes skipped), rx_lptot, tx_lptot, OFDM block 0 of tx_packet   1 280 c
//     out  H_lt, H_linear, H_cubic, H_sinc (53 each) [+ H_mmse (main.c:148 convention) + H_ls of block 0 + eq (795) + ow2]
// against front-end x 2 (3 872 c) + estimators + equalizer (2 014 c) through HBM.  One warp per frame, five groups of four
// transforms:   group 0 = { tx preamble, rx preamble (+ noise estimate), tx block 0, rx block 0 }  ->  LT_LS, the four pilot LS
// values, the three interpolators, the closed-form rank-one PS_MMSE and the equalized block 0;   groups 1-4 = rx blocks 1-14
// -> equalized with the blend of the LT_LS and PS_Linear estimates the warp keeps in shared memory (WiFi_Equalization.m:1-9).
// ------------------------------------------------------------------------------------------
template <typename T> struct ChainOut {
    cx<T> *H_lt, *H_lin, *H_cub, *H_sinc, *H_cconv, *H_ls0, *eq, *rx_symb;
    T *ow2;
};

constexpr int RC_WT = 3 * NSC * 4;           // interpolation weights staged per CTA: [est][k][4]
constexpr int RC_WARP_C = 4 * FE_TS + 2 * 56;   // complex values of shared memory per warp: tile + H_lt + H_linear

// complex divides of the chain kernel: FP32 with one hardware reciprocal (eq_div: 1-2 ulp, 0/0 -> NaN like the IEEE form, a
// quarter of its instructions -- the FP32 chain is issue-bound, ncu 66 % issue-active); FP64 the textbook cdiv
__device__ __forceinline__ float2 rc_div(float2 a, float2 b) { return eq_div(a, b); }
__device__ __forceinline__ double2 rc_div(double2 a, double2 b) { return cdiv(a, b); }
// main.c:69-72 as written: c = Re(tx) - Im(tx), H = (c rx) / (c tx); Re(tx) == Im(tx) gives 0/0 = NaN
template <typename T> __device__ __forceinline__ cx<T> rc_lt_ls(cx<T> tx, cx<T> rx)
{
    const T c = tx.x - tx.y;
    return rc_div(mk<T>(c * rx.x, c * rx.y), mk<T>(c * tx.x, c * tx.y));
}

template <typename T> __device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, int MINB>
__global__ void __launch_bounds__(FE_THREADS, MINB) rx_chain_kernel(const cx<T> *__restrict__ tx_packet, const cx<T> *__restrict__ tx_lptot,
                                                              const cx<T> *__restrict__ rx_packet, const cx<T> *__restrict__ rx_lptot,
                                                              int64_t tx_pkt_stride, ChainOut<T> out, const T *__restrict__ wtab, int64_t n_frames)
{
    extern __shared__ __align__(16) unsigned char rc_smem[];
    T *wt = reinterpret_cast<T *>(rc_smem);                                           // [3][53][4]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    cx<T> *tile = reinterpret_cast<cx<T> *>(rc_smem + ((RC_WT * sizeof(T) + 15) & ~(size_t)15)) + warp * RC_WARP_C;
    cx<T> *s_hl = tile + 4 * FE_TS, *s_hp = s_hl + 56;
    for (int i = threadIdx.x; i < RC_WT; i += FE_THREADS) wt[i] = wtab[i];
    __syncthreads();
    const int t = lane >> 3, j = lane & 7;
    cx<T> tw[8];
#pragma unroll
    for (int p = 1; p < 8; ++p) {
        T s, c;
        sincospi_t<T>((T)(j * p) * (T)(-1.0 / 32.0), &s, &c);
        tw[p] = mk<T>(c, s);
    }
    const int n2 = 2 * lane;                                      // this lane's two consecutive samples of every transform
    const int64_t wstride = (int64_t)gridDim.x * FE_WARPS;
    for (int64_t f = (int64_t)blockIdx.x * FE_WARPS + warp; f < n_frames; f += wstride) {
        // HBM -> L2 for this warp's NEXT frame: 15 rx blocks (cyclic prefixes skipped), the two long-training symbols of both
        // sides, tx block 0
        {
            const int64_t fn = f + wstride;
            if (fn < n_frames && lane < 18) {
                const cx<T> *src; uint32_t cnt = 64;
                if (lane < NBLK) src = rx_packet + fn * FE_PKT + lane * FE_BLK + FE_CP;
                else if (lane == 15) { src = rx_lptot + fn * FE_LP + 32; cnt = 128; }
                else if (lane == 16) { src = tx_lptot + fn * FE_LP + 32; cnt = 128; }
                else src = tx_packet + fn * tx_pkt_stride + FE_CP;
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(cnt * sizeof(cx<T>))) : "memory");
            }
        }
        // ---- group 0: tx preamble, rx preamble (averaged on the fly, WiFi_RX.m:19-29), tx block 0, rx block 0 ----
        T nv;
        {
            cx<T> a0, a1, b0, b1, c0, c1, d0, d1, x0, x1, y0, y1;
            ld_pair(tx_lptot + f * FE_LP + 96 + n2, a0, a1);
            ld_pair(tx_lptot + f * FE_LP + 32 + n2, b0, b1);
            ld_pair(rx_lptot + f * FE_LP + 96 + n2, c0, c1);
            ld_pair(rx_lptot + f * FE_LP + 32 + n2, d0, d1);
            ld_pair(tx_packet + f * tx_pkt_stride + FE_CP + n2, x0, x1);
            ld_pair(rx_packet + f * FE_PKT + FE_CP + n2, y0, y1);
            tile[0 * FE_TS + n2] = mk<T>((a0.x + b0.x) * (T)0.5, (a0.y + b0.y) * (T)0.5);
            tile[0 * FE_TS + n2 + 1] = mk<T>((a1.x + b1.x) * (T)0.5, (a1.y + b1.y) * (T)0.5);
            tile[1 * FE_TS + n2] = mk<T>((c0.x + d0.x) * (T)0.5, (c0.y + d0.y) * (T)0.5);
            tile[1 * FE_TS + n2 + 1] = mk<T>((c1.x + d1.x) * (T)0.5, (c1.y + d1.y) * (T)0.5);
            tile[2 * FE_TS + n2] = x0; tile[2 * FE_TS + n2 + 1] = x1;
            tile[3 * FE_TS + n2] = y0; tile[3 * FE_TS + n2 + 1] = y1;
            const T e0x = d0.x - c0.x, e0y = d0.y - c0.y, e1x = d1.x - c1.x, e1y = d1.y - c1.y;     // WiFi_RX.m:31
            nv = warp_sum<T>(e0x * e0x + e0y * e0y + e1x * e1x + e1y * e1y) * (T)(1.0 / 128.0);
        }
        if (out.ow2 && lane == 0) out.ow2[f] = nv;
        __syncwarp();
        fft4_rows<T>(tile, t, j, tw);
        // rows: 0 = tx_pre, 1 = rx_pre, 2 = tx block 0, 3 = rx block 0
        {
            // pilot LS (main.c:82-84): lane i & 3 divides pilot i, lanes 0..3 broadcast
            const int pk = WIFI_P0 + (WIFI_P1 - WIFI_P0) * (lane & 3);
            const cx<T> mine = rc_div(tile[3 * NSC + pk], tile[2 * NSC + pk]);
            cx<T> hp[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) hp[i] = mk<T>(__shfl_sync(0xffffffffu, mine.x, i), __shfl_sync(0xffffffffu, mine.y, i));
            cx<T> hl[2], v[2], r0[2];
            T a3x = 0, a3y = 0, vv = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = lane + 32 * h;
                hl[h] = v[h] = r0[h] = mk<T>(0, 0);
                if (k < NSC) {
                    const cx<T> x0 = tile[2 * NSC + k];
                    r0[h] = tile[3 * NSC + k];
                    hl[h] = k == DCBIN ? mk<T>(0, 0) : rc_lt_ls<T>(tile[k], tile[NSC + k]);          // main.c:66-75
                    if (out.H_lt) st_stream(out.H_lt + f * NSC + k, hl[h]);
                    if (out.H_ls0) st_stream(out.H_ls0 + f * NSC + k, rc_div(r0[h], x0));
                    const T *w = wt + k * 4;
                    const cx<T> lin = mk<T>(w[0] * hp[0].x + w[1] * hp[1].x + w[2] * hp[2].x + w[3] * hp[3].x,
                                            w[0] * hp[0].y + w[1] * hp[1].y + w[2] * hp[2].y + w[3] * hp[3].y);
                    if (out.H_lin) st_stream(out.H_lin + f * NSC + k, lin);
                    if (out.H_cub) {
                        const T *wc = wt + (NSC + k) * 4;
                        st_stream(out.H_cub + f * NSC + k, mk<T>(wc[0] * hp[0].x + wc[1] * hp[1].x + wc[2] * hp[2].x + wc[3] * hp[3].x,
                                                                  wc[0] * hp[0].y + wc[1] * hp[1].y + wc[2] * hp[2].y + wc[3] * hp[3].y));
                    }
                    if (out.H_sinc) {
                        const T *ws = wt + (2 * NSC + k) * 4;
                        st_stream(out.H_sinc + f * NSC + k, mk<T>(ws[0] * hp[0].x + ws[1] * hp[1].x + ws[2] * hp[2].x + ws[3] * hp[3].x,
                                                                   ws[0] * hp[0].y + ws[1] * hp[1].y + ws[2] * hp[2].y + ws[3] * hp[3].y));
                    }
                    s_hl[k] = hl[h]; s_hp[k] = lin;
                    // rank-one PS_MMSE (main.c:148 convention, mmse_rank1_kernel): v = tx (.) H_lt, a3 = v^H rx, vv = v^H v
                    v[h] = cmul(x0, hl[h]);
                    a3x += v[h].x * r0[h].x + v[h].y * r0[h].y;
                    a3y += v[h].x * r0[h].y - v[h].y * r0[h].x;
                    vv += v[h].x * v[h].x + v[h].y * v[h].y;
                    if (out.rx_symb) st_stream(out.rx_symb + f * FRAME + k, r0[h]);
                    if (out.eq) {                                                                       // block 0: i = 1
                        const T wp = (T)(1.0 / NBLK), wl = (T)1 - wp;
                        st_stream(out.eq + f * FRAME + k,
                                  k == DCBIN ? mk<T>(0, 0) : eq_div(r0[h], mk<T>(wl * hl[h].x + wp * lin.x, wl * hl[h].y + wp * lin.y)));
                    }
                }
            }
            if (out.H_cconv) {                                     // warp-uniform
                a3x = warp_sum<T>(a3x); a3y = warp_sum<T>(a3y); vv = warp_sum<T>(vv);
                const T den = nv + vv;
                const cx<T> c = mk<T>(a3x / den, a3y / den);
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (lane + 32 * h < NSC) st_stream(out.H_cconv + f * NSC + lane + 32 * h, cmul(hl[h], c));
            }
        }
        __syncwarp();
        if (!out.eq && !out.rx_symb) continue;                     // warp-uniform: the estimates alone need block 0 only
        // ---- groups 1-4: rx blocks 1..14 -> equalized symbols (and the symbols themselves on request); the samples of the next
        //      group are requested before the transforms of the current one ----
        cx<T> nx[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) ld_pair(rx_packet + f * FE_PKT + (1 + i) * FE_BLK + FE_CP + n2, nx[i][0], nx[i][1]);
#pragma unroll 1
        for (int g = 0; g < 4; ++g) {
            const int b0 = 1 + 4 * g;
            const int rows = g < 3 ? 4 : 2;
#pragma unroll
            for (int i = 0; i < 4; ++i) { tile[i * FE_TS + n2] = nx[i][0]; tile[i * FE_TS + n2 + 1] = nx[i][1]; }
            __syncwarp();
            if (g < 3) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    nx[i][0] = nx[i][1] = mk<T>(0, 0);
                    if (b0 + 4 + i < NBLK) ld_pair(rx_packet + f * FE_PKT + (b0 + 4 + i) * FE_BLK + FE_CP + n2, nx[i][0], nx[i][1]);
                }
            }
            fft4_rows<T>(tile, t, j, tw);
            const int64_t base = f * FRAME + b0 * NSC;
#pragma unroll
            for (int i = 0; i < 7; ++i) {
                const int o = i * 32 + lane;
                if (o < rows * NSC) {
                    const cx<T> r = tile[o];
                    if (out.rx_symb) st_stream(out.rx_symb + base + o, r);
                    if (out.eq) {
                        const int tt = o / NSC, k = o - tt * NSC;
                        const T wp = (T)(b0 + tt + 1) * (T)(1.0 / NBLK), wl = (T)1 - wp;               // WiFi_Equalization.m:4-5, i = b + 1
                        const cx<T> hl = s_hl[k], hq = s_hp[k];
                        st_stream(out.eq + base + o, k == DCBIN ? mk<T>(0, 0) : eq_div(r, mk<T>(wl * hl.x + wp * hq.x, wl * hl.y + wp * hq.y)));
                    }
                }
            }
            __syncwarp();
        }
    }
}

cudaError_t launch_rx_chain(wifi_dtype dt, const void *tx_packet, int64_t tx_pkt_stride, const void *tx_lptot, const void *rx_packet, const void *rx_lptot,
                            void *H_lt, void *H_lin, void *H_cub, void *H_sinc, void *H_cconv, void *H_ls0, void *eq, void *rx_symb, void *ow2,
                            int64_t n_frames, const InterpTables &tab, cudaStream_t s)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    const int64_t need = (n_frames + FE_WARPS - 1) / FE_WARPS;
    cudaError_t e;
    // measured on B200, 256 Ki frames, all planes: FP32 1.035 / 0.976 / 1.191 ms at 2 / 3 / 4 CTAs per SM (116 / 80 / 64 registers);
    // FP64 2.151 / 3.085 / 3.588 ms (128 registers, then spills)
#define RC_LAUNCH(T, T2, MB, tabw)                                                                                                         \
    do {                                                                                                                                   \
        const size_t smem = ((RC_WT * sizeof(T) + 15) & ~(size_t)15) + (size_t)FE_WARPS * RC_WARP_C * sizeof(T2);                          \
        if ((e = cudaFuncSetAttribute(rx_chain_kernel<T, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e; \
        const ChainOut<T> o = {(T2 *)H_lt, (T2 *)H_lin, (T2 *)H_cub, (T2 *)H_sinc, (T2 *)H_cconv, (T2 *)H_ls0, (T2 *)eq, (T2 *)rx_symb, (T *)ow2}; \
        rx_chain_kernel<T, MB><<<std::min<int64_t>(need, 148 * MB * 2), FE_THREADS, smem, s>>>((const T2 *)tx_packet, (const T2 *)tx_lptot, (const T2 *)rx_packet, \
                                                                  (const T2 *)rx_lptot, tx_pkt_stride, o, tabw, n_frames);                 \
    } while (0)
    if (dt == WIFI_F32) RC_LAUNCH(float, float2, 3, tab.w32);
    else RC_LAUNCH(double, double2, 2, tab.w64);
#undef RC_LAUNCH
    return cudaGetLastError();
}

}  // namespace wifi
