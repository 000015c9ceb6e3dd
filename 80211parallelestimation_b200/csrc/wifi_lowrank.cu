// wifi_lowrank.cu -- per-frame PS_MMSE for a LOW-RANK channel covariance, in one HBM-bound launch.
//
//   H_f = R (R + sigma2_f diag(1/|x_fk|^2))^-1 (rx_f / x_f)            (WiFi_channel_estimation_PS_MMSE.m:16-33, north-star form)
//
// A channel of L taps has a covariance of rank L: R = U L U^H with U 53 x r (r <= 8 here; the synthetic channel of SURVEY 8(d)
// has r = 4).  The push-through identity  R (R + D)^-1 = U (L^-1 + U^H D^-1 U)^-1 U^H D^-1  with D^-1 = diag(|x_k|^2) / sigma2 gives
//     t = U^H (conj(x) (.) rx)                       r complex values            (|x|^2 rx/x = conj(x) rx: no divide, no 1/x)
//     G = U^H diag(|x|^2) U                          r x r Hermitian
//     S = sigma2 L^-1 + G,   S w = t                 r x r Hermitian positive definite solve in registers
//     H = U w
// -- 6 kflop per frame at r = 4 instead of the 442 kflop of the 53 x 53 solve (wifi_solve_hpd.cu), on the frame's own 159 complex
// values of traffic.  Both sigma2 AND the modulus pattern |x_k|^2 may differ per frame (QAM), which the eigen-domain path
// (wifi_eig.cu: shared |x_k|^2 only) cannot do; a bin without transmit energy (the DC bin) simply contributes nothing.
// S is well conditioned (G ~ |x|^2 I, sigma2/l tiny), so FP32 ARITHMETIC meets the 1e-4 bound here, where FP32 elimination of
// the 53 x 53 matrix R + D (condition 1e7) cannot.
//
// Kernels: one persistent CTA per SM; a warp owns a chunk of 16 frames at a time (a 10 / 20 KB private shared-memory tile per warp),
// lane = (frame, half of the bins); the two halves of a frame are summed by one shuffle step.
//   a. coalesced element-wise pass over the chunk (a contiguous run of 16-byte vectors: tile index = element index):
//      a = conj(x) rx and m = |x|^2 into the tile;
//   b. t and G accumulated over the bins from the tile (conflict-free row strides) and the shared tables U_k, P_k = conj(U_ki) U_kj
//      (broadcast 16-byte loads);   c. S w = t by Hermitian elimination with static register indices;
//   d. H = U w written back through the tile, coalesced streaming store.
// FP32 accumulations are packed FFMA2 pairs.  The next chunk of the warp is pulled into L2 (cp.async.bulk.prefetch.L2) while the current one is processed.
#include <algorithm>
#include <cmath>
#include <type_traits>
#include <vector>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

// tables (in T): U[53][2 R] (re of U_kj, j < R; then im: planar, so that FP32 pairs over j are adjacent) |
//                P[53][R R] (|U_ki|^2, i < R; then re, im of conj(U_ki) U_kj, i < j) | 1 / l_j [R]
template <int R> struct LrTab {
    static constexpr int U = 0, P = NSC * 2 * R, L = P + NSC * R * R, SIZE = (L + R + 3) & ~3;
};
__host__ __device__ constexpr int lr_pair(int R, int i, int j) { return i * R - i * (i + 1) / 2 + (j - i - 1); }     // i < j

template <int N> __device__ __forceinline__ void lr_load2(const float *p, f32x2 (&v)[N])          // N pairs = N / 2 16-byte vectors
{
    static_assert(N % 2 == 0, "whole 16-byte vectors");
#pragma unroll
    for (int q = 0; q < N / 2; ++q) { const ulonglong2 w = reinterpret_cast<const ulonglong2 *>(p)[q]; v[2 * q] = w.x; v[2 * q + 1] = w.y; }
}
template <int N> __device__ __forceinline__ void lr_load(const double *p, double (&v)[N])
{
    static_assert(N % 2 == 0, "whole 16-byte vectors");
#pragma unroll
    for (int q = 0; q < N / 2; ++q) { const double2 w = reinterpret_cast<const double2 *>(p)[q]; v[2 * q] = w.x; v[2 * q + 1] = w.y; }
}
__device__ __forceinline__ void lr_l2_prefetch(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

// S = G + sigma2 / l on the diagonal;  S w = t by Hermitian elimination (upper triangle stored), back-substitution; w overwrites t
template <typename T, int R>
__device__ __forceinline__ void lr_solve(T (&tr)[R], T (&ti)[R], T (&gd)[R], T (&gor)[R * (R - 1) / 2], T (&goi)[R * (R - 1) / 2], T s2, const T *linv)
{
    T dinv[R];
#pragma unroll
    for (int j = 0; j < R; ++j) gd[j] = fma(s2, linv[j], gd[j]);
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const T inv = (T)1 / gd[k];
        dinv[k] = inv;
#pragma unroll
        for (int i = k + 1; i < R; ++i) {
            const T sr = gor[lr_pair(R, k, i)], si = goi[lr_pair(R, k, i)];
            const T lr = sr * inv, li = -si * inv;                         // l = conj(S_ki) / S_kk
            const T tkr = tr[k], tki = ti[k];
            tr[i] -= lr * tkr - li * tki;
            ti[i] -= lr * tki + li * tkr;
            gd[i] -= (sr * sr + si * si) * inv;
#pragma unroll
            for (int j = i + 1; j < R; ++j) {
                const T kr = gor[lr_pair(R, k, j)], ki = goi[lr_pair(R, k, j)];
                gor[lr_pair(R, i, j)] -= lr * kr - li * ki;
                goi[lr_pair(R, i, j)] -= lr * ki + li * kr;
            }
        }
    }
#pragma unroll
    for (int k = R - 1; k >= 0; --k) {
        T ar = tr[k], ai = ti[k];
#pragma unroll
        for (int j = k + 1; j < R; ++j) {
            const T sr = gor[lr_pair(R, k, j)], si = goi[lr_pair(R, k, j)];
            ar -= sr * tr[j] - si * ti[j];
            ai -= sr * ti[j] + si * tr[j];
        }
        tr[k] = ar * dinv[k];
        ti[k] = ai * dinv[k];
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Both kernels: chunk = 16 frames, lane = (frame f = lane & 15, half h = lane >> 4 of the bins: h = 0 -> bins 0..26, h = 1 -> bins
// 27..52).  The tile holds 16 x 53 values (a: complex, m: real), the chunk is ONE contiguous run of 16-byte vectors in HBM and
// staging is index-free; the two halves of a frame meet in one __shfl_xor(16) step per accumulator.  A 10 KB (FP32) / 20 KB (FP64)
// tile per warp puts 16 / 10 warps on an SM.
constexpr int LR_FR = 16, LR_TILE = LR_FR * NSC, LR_HB = 27;

// FP32: 16 warps at rank <= 4 (128 registers), 12 beyond (168)
template <int R> struct Lr32Warps { static constexpr int N = R <= 4 ? 16 : 12; };

template <int R>
__global__ void __launch_bounds__(Lr32Warps<R>::N * 32, 1)
    mmse_lowrank_f32_kernel(const float *__restrict__ tab_g, const float2 *__restrict__ tx, const float2 *__restrict__ rx, int64_t stride,
                            const float *__restrict__ sigma2, float2 *__restrict__ H, int64_t n, int aligned16)
{
    constexpr int NP = R * (R - 1) / 2, UNR = 2, LR32_WARPS = Lr32Warps<R>::N;
    using TB = LrTab<R>;
    extern __shared__ __align__(16) unsigned char lr_smem[];
    float *tab = reinterpret_cast<float *>(lr_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float2 *A = reinterpret_cast<float2 *>(tab + TB::SIZE) + warp * LR_TILE;                           // this warp's tile: a, later H
    float *M = reinterpret_cast<float *>(reinterpret_cast<float2 *>(tab + TB::SIZE) + LR32_WARPS * LR_TILE) + warp * LR_TILE;
    for (int e = threadIdx.x; e < TB::SIZE; e += LR32_WARPS * 32) tab[e] = tab_g[e];
    __syncthreads();

    const int64_t n_chunks = (n + LR_FR - 1) / LR_FR, cstep = (int64_t)gridDim.x * LR32_WARPS;
    const bool dense = stride == NSC && (aligned16 & 1), out16 = (aligned16 & 2) != 0;
    constexpr uint32_t CHUNK_BYTES = LR_TILE * sizeof(float2);
    const int f = lane & 15, hb = lane >> 4, k0 = hb * LR_HB, nb = hb ? NSC - LR_HB : LR_HB;
    // chunks are dealt so that the whole grid sweeps HBM as one moving window: (round * gridDim + block) * warps + warp
    for (int64_t chunk = (int64_t)blockIdx.x * LR32_WARPS + warp; chunk < n_chunks; chunk += cstep) {
        const int64_t f0 = chunk * LR_FR;
        const int valid = (int)min((int64_t)LR_FR, n - f0);
        const bool fast = dense && valid == LR_FR, fast_out = out16 && valid == LR_FR;
        {
            const int64_t fn = (chunk + cstep) * LR_FR;
            if (dense && lane == 0 && fn + LR_FR <= n) { lr_l2_prefetch(tx + fn * NSC, CHUNK_BYTES); lr_l2_prefetch(rx + fn * NSC, CHUNK_BYTES); }
        }
        // ---- a. element-wise, coalesced: a = conj(x) rx, m = |x|^2 -> tile ----
        if (fast) {
            // dense 16-byte aligned rows, whole chunk: one contiguous run of 424 float4 (two bins each); tile index = element index,
            // no address arithmetic.  Four batches of 4 vectors per array, two in flight.
            const float4 *px = reinterpret_cast<const float4 *>(tx + f0 * NSC) + lane, *pr = reinterpret_cast<const float4 *>(rx + f0 * NSC) + lane;
            float4 *A4 = reinterpret_cast<float4 *>(A) + lane;
            float2 *M2 = reinterpret_cast<float2 *>(M) + lane;
            constexpr int VB = 4;
            float4 xv[2][VB], rv[2][VB];
#pragma unroll
            for (int j = 0; j < VB; ++j) { xv[0][j] = ld_stream(px + 32 * j); rv[0][j] = ld_stream(pr + 32 * j); }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (b < 3) {
#pragma unroll
                    for (int j = 0; j < VB; ++j) {
                        const int i = (b + 1) * VB + j;
                        if (i < 13 || (i == 13 && lane < 8)) { xv[(b + 1) & 1][j] = ld_stream(px + 32 * i); rv[(b + 1) & 1][j] = ld_stream(pr + 32 * i); }
                    }
                }
#pragma unroll
                for (int j = 0; j < VB; ++j) {
                    const int i = b * VB + j;
                    if (i < 13 || (i == 13 && lane < 8)) {
                        const float4 x = xv[b & 1][j], r = rv[b & 1][j];
                        A4[32 * i] = make_float4(fmaf(x.x, r.x, x.y * r.y), fmaf(x.x, r.y, -(x.y * r.x)), fmaf(x.z, r.z, x.w * r.w), fmaf(x.z, r.w, -(x.w * r.z)));
                        M2[32 * i] = make_float2(fmaf(x.x, x.x, x.y * x.y), fmaf(x.z, x.z, x.w * x.w));
                    }
                }
            }
        } else {
            // row by row (strided, 8-byte aligned or ragged): lane kk takes bins kk and kk + 32 of the row, 4 rows in flight
            constexpr int RB = 4;
#pragma unroll 1
            for (int fr0 = 0; fr0 < LR_FR; fr0 += RB) {
#pragma unroll 1
                for (int kk = lane; kk < NSC; kk += 32) {
                    float2 xv[RB], rv[RB];
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        xv[j] = rv[j] = make_float2(0.f, 0.f);
                        if (fr0 + j < valid) {
                            const int64_t off = (f0 + fr0 + j) * stride + kk;
                            xv[j] = ld_stream(tx + off);
                            rv[j] = ld_stream(rx + off);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        const float2 x = xv[j], r = rv[j];
                        A[(fr0 + j) * NSC + kk] = make_float2(fmaf(x.x, r.x, x.y * r.y), fmaf(x.x, r.y, -(x.y * r.x)));
                        M[(fr0 + j) * NSC + kk] = fmaf(x.x, x.x, x.y * x.y);
                    }
                }
            }
        }
        __syncwarp();
        // ---- b. lane = (frame, half): t += conj(U_k) a_k, G += m_k P_k over this half's bins as FFMA2 -- pairs over j (t) and
        //         (re, im) pairs (G): 16 instead of 32 FMA instructions per bin ----
        f32x2 trp[R / 2], tip[R / 2], gdp[R / 2], gop[NP];
#pragma unroll
        for (int h = 0; h < R / 2; ++h) trp[h] = tip[h] = gdp[h] = 0ull;
#pragma unroll
        for (int q = 0; q < NP; ++q) gop[q] = 0ull;
        float2 *Ar = A + f * NSC + k0;
        const float *Mr = M + f * NSC + k0;
        const float *Uh = tab + TB::U + k0 * 2 * R, *Ph = tab + TB::P + k0 * R * R;
#pragma unroll UNR
        for (int kk = 0; kk < LR_HB; ++kk) {
            const bool on = kk < nb;                                           // the upper half has 26 bins
            const int kc = on ? kk : 0;
            float2 a = Ar[kc];
            float m = Mr[kc];
            if (!on) { a = make_float2(0.f, 0.f); m = 0.f; }
            f32x2 u[R], p[R * R / 2];
            lr_load2<R>(Uh + kc * 2 * R, u);                                   // u[0 .. R/2) = re pairs, u[R/2 .. R) = im pairs
            lr_load2<R * R / 2>(Ph + kc * R * R, p);
            const f32x2 axx = pack2(a.x, a.x), ayy = pack2(a.y, a.y), nax = pack2(-a.x, -a.x), mm = pack2(m, m);
#pragma unroll
            for (int h = 0; h < R / 2; ++h) {
                trp[h] = ffma2(u[h], axx, trp[h]); trp[h] = ffma2(u[R / 2 + h], ayy, trp[h]);
                tip[h] = ffma2(u[h], ayy, tip[h]); tip[h] = ffma2(u[R / 2 + h], nax, tip[h]);
                gdp[h] = ffma2(mm, p[h], gdp[h]);
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) gop[q] = ffma2(mm, p[R / 2 + q], gop[q]);
        }
        // ---- c. the other half of the frame lives 16 lanes away; the r x r solve (both halves, redundantly) ----
        float tr[R], ti[R], gd[R], gor[NP], goi[NP];
#pragma unroll
        for (int h = 0; h < R / 2; ++h) { unpack2(trp[h], tr[2 * h], tr[2 * h + 1]); unpack2(tip[h], ti[2 * h], ti[2 * h + 1]); unpack2(gdp[h], gd[2 * h], gd[2 * h + 1]); }
#pragma unroll
        for (int q = 0; q < NP; ++q) unpack2(gop[q], gor[q], goi[q]);
#pragma unroll
        for (int j = 0; j < R; ++j) {
            tr[j] += __shfl_xor_sync(0xffffffffu, tr[j], 16); ti[j] += __shfl_xor_sync(0xffffffffu, ti[j], 16); gd[j] += __shfl_xor_sync(0xffffffffu, gd[j], 16);
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) { gor[q] += __shfl_xor_sync(0xffffffffu, gor[q], 16); goi[q] += __shfl_xor_sync(0xffffffffu, goi[q], 16); }
        __syncwarp();
        lr_solve<float, R>(tr, ti, gd, gor, goi, f < valid ? sigma2[f0 + f] : 1.f, tab + TB::L);
        // ---- d. H = U w for this half's bins through the tile (pairs over j, one final add), coalesced streaming store ----
        {
            f32x2 wr2[R / 2], wi2[R / 2], nwi2[R / 2];
#pragma unroll
            for (int h = 0; h < R / 2; ++h) {
                wr2[h] = pack2(tr[2 * h], tr[2 * h + 1]); wi2[h] = pack2(ti[2 * h], ti[2 * h + 1]); nwi2[h] = pack2(-ti[2 * h], -ti[2 * h + 1]);
            }
#pragma unroll 3
            for (int kk = 0; kk < LR_HB; ++kk) {
                if (kk < nb) {
                    f32x2 u[R];
                    lr_load2<R>(Uh + kk * 2 * R, u);
                    f32x2 hr2 = 0ull, hi2 = 0ull;
#pragma unroll
                    for (int h = 0; h < R / 2; ++h) {
                        hr2 = ffma2(u[h], wr2[h], hr2); hr2 = ffma2(u[R / 2 + h], nwi2[h], hr2);
                        hi2 = ffma2(u[h], wi2[h], hi2); hi2 = ffma2(u[R / 2 + h], wr2[h], hi2);
                    }
                    float r0, r1, i0, i1;
                    unpack2(hr2, r0, r1); unpack2(hi2, i0, i1);
                    Ar[kk] = make_float2(r0 + r1, i0 + i1);
                }
            }
        }
        __syncwarp();
        if (fast_out) {
            const float4 *A4 = reinterpret_cast<const float4 *>(A) + lane;
            float4 *po = reinterpret_cast<float4 *>(H + f0 * NSC) + lane;
#pragma unroll 7
            for (int i = 0; i < 14; ++i)
                if (i < 13 || lane < 8) st_stream(po + 32 * i, A4[32 * i]);
        } else {
            for (int fr = 0; fr < valid; ++fr)
                for (int kk = lane; kk < NSC; kk += 32) st_stream(H + (f0 + fr) * NSC + kk, A[fr * NSC + kk]);
        }
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// FP64: 10 warps (8 at rank > 4), one complex128 value per 16-byte vector.
// (First version: 32 frames per warp in two slabs of 27 / 26 bins, staged row by row with 27 of 32 lanes: 0.732 ms per 1 Mi frames,
// ncu long_scoreboard 28 % of the samples in the staging loads; this one: 0.614 ms.)
template <int R> struct Lr64Warps { static constexpr int N = R <= 4 ? 10 : 8; };
constexpr int LR64_FR = LR_FR, LR64_TILE = LR_TILE, LR64_HB = LR_HB;

template <int R>
__global__ void __launch_bounds__(Lr64Warps<R>::N * 32, 1)
    mmse_lowrank_f64_kernel(const double *__restrict__ tab_g, const double2 *__restrict__ tx, const double2 *__restrict__ rx, int64_t stride,
                            const double *__restrict__ sigma2, double2 *__restrict__ H, int64_t n, int aligned16)
{
    constexpr int NP = R * (R - 1) / 2, WARPS = Lr64Warps<R>::N, UNR = R <= 4 ? 2 : 1;
    using TB = LrTab<R>;
    extern __shared__ __align__(16) unsigned char lr_smem[];
    double *tab = reinterpret_cast<double *>(lr_smem);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double2 *A = reinterpret_cast<double2 *>(tab + TB::SIZE) + warp * LR64_TILE;
    double *M = reinterpret_cast<double *>(reinterpret_cast<double2 *>(tab + TB::SIZE) + WARPS * LR64_TILE) + warp * LR64_TILE;
    for (int e = threadIdx.x; e < TB::SIZE; e += WARPS * 32) tab[e] = tab_g[e];
    __syncthreads();

    const int64_t n_chunks = (n + LR64_FR - 1) / LR64_FR, cstep = (int64_t)gridDim.x * WARPS;
    const bool dense = stride == NSC && (aligned16 & 1);                   // (complex128 arrays are 16-byte aligned unless the caller offsets them oddly)
    const bool out16 = (aligned16 & 2) != 0;
    constexpr uint32_t CHUNK_BYTES = LR64_TILE * sizeof(double2);
    const int f = lane & 15, hb = lane >> 4, k0 = hb * LR64_HB, nb = hb ? NSC - LR64_HB : LR64_HB;
    for (int64_t chunk = (int64_t)blockIdx.x * WARPS + warp; chunk < n_chunks; chunk += cstep) {
        const int64_t f0 = chunk * LR64_FR;
        const int valid = (int)min((int64_t)LR64_FR, n - f0);
        const bool fast = dense && valid == LR64_FR, fast_out = out16 && valid == LR64_FR;
        {
            const int64_t fn = (chunk + cstep) * LR64_FR;
            if (dense && lane == 0 && fn + LR64_FR <= n) { lr_l2_prefetch(tx + fn * NSC, CHUNK_BYTES); lr_l2_prefetch(rx + fn * NSC, CHUNK_BYTES); }
        }
        // ---- a. element-wise, coalesced ----
        if (fast) {
            // 848 complex128 values = 848 16-byte vectors, contiguous: seven batches of 4 vectors per array, two in flight
            const double2 *px = tx + f0 * NSC + lane, *pr = rx + f0 * NSC + lane;
            constexpr int VB = 4;
            double2 xv[2][VB], rv[2][VB];
#pragma unroll
            for (int j = 0; j < VB; ++j) { xv[0][j] = ld_stream(px + 32 * j); rv[0][j] = ld_stream(pr + 32 * j); }
#pragma unroll
            for (int b = 0; b < 7; ++b) {
                if (b < 6) {
#pragma unroll
                    for (int j = 0; j < VB; ++j) {
                        const int i = (b + 1) * VB + j;
                        if (i < 26 || (i == 26 && lane < 16)) { xv[(b + 1) & 1][j] = ld_stream(px + 32 * i); rv[(b + 1) & 1][j] = ld_stream(pr + 32 * i); }
                    }
                }
#pragma unroll
                for (int j = 0; j < VB; ++j) {
                    const int i = b * VB + j;
                    if (i < 26 || (i == 26 && lane < 16)) {
                        const double2 x = xv[b & 1][j], r = rv[b & 1][j];
                        A[32 * i + lane] = make_double2(fma(x.x, r.x, x.y * r.y), fma(x.x, r.y, -(x.y * r.x)));
                        M[32 * i + lane] = fma(x.x, x.x, x.y * x.y);
                    }
                }
            }
        } else {
            constexpr int RB = 4;
#pragma unroll 1
            for (int fr0 = 0; fr0 < LR64_FR; fr0 += RB) {
#pragma unroll 1
                for (int kk = lane; kk < NSC; kk += 32) {
                    double2 xv[RB], rv[RB];
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        xv[j] = rv[j] = make_double2(0.0, 0.0);
                        if (fr0 + j < valid) {
                            const int64_t off = (f0 + fr0 + j) * stride + kk;
                            xv[j] = ld_stream(tx + off);
                            rv[j] = ld_stream(rx + off);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        const double2 x = xv[j], r = rv[j];
                        A[(fr0 + j) * NSC + kk] = make_double2(fma(x.x, r.x, x.y * r.y), fma(x.x, r.y, -(x.y * r.x)));
                        M[(fr0 + j) * NSC + kk] = fma(x.x, x.x, x.y * x.y);
                    }
                }
            }
        }
        __syncwarp();
        // ---- b. lane = (frame, half): partial t and G over this half's bins ----
        double tr[R], ti[R], gd[R], gor[NP], goi[NP];
#pragma unroll
        for (int j = 0; j < R; ++j) tr[j] = ti[j] = gd[j] = 0.0;
#pragma unroll
        for (int q = 0; q < NP; ++q) gor[q] = goi[q] = 0.0;
        double2 *Ar = A + f * NSC + k0;
        const double *Mr = M + f * NSC + k0;
        const double *Uh = tab + TB::U + k0 * 2 * R, *Ph = tab + TB::P + k0 * R * R;
#pragma unroll UNR
        for (int kk = 0; kk < LR64_HB; ++kk) {
            const bool on = kk < nb;                                           // the upper half has 26 bins
            const int kc = on ? kk : 0;
            double2 a = Ar[kc];
            double m = Mr[kc];
            if (!on) { a = make_double2(0.0, 0.0); m = 0.0; }
            double u[2 * R], p[R * R];
            lr_load<2 * R>(Uh + kc * 2 * R, u);
            lr_load<R * R>(Ph + kc * R * R, p);
#pragma unroll
            for (int j = 0; j < R; ++j) {
                tr[j] = fma(u[j], a.x, tr[j]); tr[j] = fma(u[R + j], a.y, tr[j]);
                ti[j] = fma(u[j], a.y, ti[j]); ti[j] = fma(-u[R + j], a.x, ti[j]);
                gd[j] = fma(m, p[j], gd[j]);
            }
#pragma unroll
            for (int q = 0; q < NP; ++q) { gor[q] = fma(m, p[R + 2 * q], gor[q]); goi[q] = fma(m, p[R + 2 * q + 1], goi[q]); }
        }
        // the other half of the frame lives 16 lanes away
#pragma unroll
        for (int j = 0; j < R; ++j) {
            tr[j] += __shfl_xor_sync(0xffffffffu, tr[j], 16); ti[j] += __shfl_xor_sync(0xffffffffu, ti[j], 16); gd[j] += __shfl_xor_sync(0xffffffffu, gd[j], 16);
        }
#pragma unroll
        for (int q = 0; q < NP; ++q) { gor[q] += __shfl_xor_sync(0xffffffffu, gor[q], 16); goi[q] += __shfl_xor_sync(0xffffffffu, goi[q], 16); }
        __syncwarp();
        // ---- c. the r x r solve (both halves, redundantly) ----
        lr_solve<double, R>(tr, ti, gd, gor, goi, f < valid ? sigma2[f0 + f] : 1.0, tab + TB::L);
        // ---- d. H = U w for this half's bins (four independent chains per bin), coalesced streaming store ----
#pragma unroll 2
        for (int kk = 0; kk < LR64_HB; ++kk) {
            if (kk < nb) {
                double u[2 * R];
                lr_load<2 * R>(Uh + kk * 2 * R, u);
                double hra = 0.0, hrb = 0.0, hia = 0.0, hib = 0.0;
#pragma unroll
                for (int j = 0; j < R; ++j) {
                    hra = fma(u[j], tr[j], hra); hrb = fma(-u[R + j], ti[j], hrb);
                    hia = fma(u[j], ti[j], hia); hib = fma(u[R + j], tr[j], hib);
                }
                Ar[kk] = make_double2(hra + hrb, hia + hib);
            }
        }
        __syncwarp();
        if (fast_out) {
            double2 *po = H + f0 * NSC + lane;
#pragma unroll 9
            for (int i = 0; i < 27; ++i)
                if (i < 26 || lane < 16) st_stream(po + 32 * i, A[32 * i + lane]);
        } else {
            for (int fr = 0; fr < valid; ++fr)
                for (int kk = lane; kk < NSC; kk += 32) st_stream(H + (f0 + fr) * NSC + kk, A[fr * NSC + kk]);
        }
        __syncwarp();
    }
}

template <int R>
static cudaError_t launch_lr32(const void *tab, const void *tx, const void *rx, int64_t stride, const void *sigma2, void *H, int64_t n, int aligned16, cudaStream_t s)
{
    constexpr int LR32_WARPS = Lr32Warps<R>::N;
    const size_t smem = sizeof(float) * LrTab<R>::SIZE + (size_t)LR32_WARPS * LR_TILE * (sizeof(float2) + sizeof(float));
    auto kern = mmse_lowrank_f32_kernel<R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t n_chunks = (n + LR_FR - 1) / LR_FR;
    const unsigned grid = (unsigned)std::min<int64_t>((n_chunks + LR32_WARPS - 1) / LR32_WARPS, 148);
    kern<<<grid, LR32_WARPS * 32, smem, s>>>((const float *)tab, (const float2 *)tx, (const float2 *)rx, stride, (const float *)sigma2, (float2 *)H, n, aligned16);
    return cudaGetLastError();
}

template <int R>
static cudaError_t launch_lr64(const void *tab, const void *tx, const void *rx, int64_t stride, const void *sigma2, void *H, int64_t n, int aligned16, cudaStream_t s)
{
    constexpr int WARPS = Lr64Warps<R>::N;
    const size_t smem = sizeof(double) * LrTab<R>::SIZE + (size_t)WARPS * LR64_TILE * (sizeof(double2) + sizeof(double));
    auto kern = mmse_lowrank_f64_kernel<R>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t n_chunks = (n + LR64_FR - 1) / LR64_FR;
    const unsigned grid = (unsigned)std::min<int64_t>((n_chunks + WARPS - 1) / WARPS, 148);
    kern<<<grid, WARPS * 32, smem, s>>>((const double *)tab, (const double2 *)tx, (const double2 *)rx, stride, (const double *)sigma2, (double2 *)H, n, aligned16);
    return cudaGetLastError();
}

cudaError_t launch_mmse_lowrank(wifi_dtype dt, int rank_padded, const void *tab, const void *tx, const void *rx, int64_t frame_stride,
                                const void *sigma2, void *H, int64_t n_frames, cudaStream_t s)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    const int aligned16 = (((((uintptr_t)tx) | ((uintptr_t)rx)) & 15) == 0 ? 1 : 0) | ((((uintptr_t)H) & 15) == 0 ? 2 : 0);
    if (dt == WIFI_F32)
        return rank_padded == 4 ? launch_lr32<4>(tab, tx, rx, frame_stride, sigma2, H, n_frames, aligned16, s)
                                : launch_lr32<8>(tab, tx, rx, frame_stride, sigma2, H, n_frames, aligned16, s);
    return rank_padded == 4 ? launch_lr64<4>(tab, tx, rx, frame_stride, sigma2, H, n_frames, aligned16, s)
                            : launch_lr64<8>(tab, tx, rx, frame_stride, sigma2, H, n_frames, aligned16, s);
}

// Host side of wifi_mmse_lowrank_prepare: eigenvectors V [53][53] (columns; row-major double2) and eigenvalues lam [53] of R
// (wifi_eig.cu with |x| = 1; eigenvalues below 64 eps l_max arrive as exact zeros) -> the kernel's tables in both precisions.
// Returns the rank (0: R = 0; > WIFI_LOWRANK_MAX: unsupported, tables untouched).
int lowrank_build_tables(const double *V /* re, im interleaved */, const double *lam, int *rank_padded, std::vector<float> &t32, std::vector<double> &t64)
{
    int idx[NSC], r = 0;
    for (int i = 0; i < NSC; ++i) if (lam[i] > 0.0) idx[r++] = i;
    std::sort(idx, idx + r, [&](int a, int b) { return lam[a] > lam[b]; });
    if (r == 0 || r > WIFI_LOWRANK_MAX) return r;
    const int R = r <= 4 ? 4 : 8;
    *rank_padded = R;
    const int offU = 0, offP = NSC * 2 * R, offL = offP + NSC * R * R, size = (offL + R + 3) & ~3;
    t64.assign(size, 0.0);
    for (int k = 0; k < NSC; ++k) {
        double ur[8] = {0}, ui[8] = {0};
        for (int j = 0; j < r; ++j) { ur[j] = V[2 * (k * NSC + idx[j])]; ui[j] = V[2 * (k * NSC + idx[j]) + 1]; }
        double *U = &t64[offU + k * 2 * R], *P = &t64[offP + k * R * R];
        for (int j = 0; j < R; ++j) { U[j] = ur[j]; U[R + j] = ui[j]; P[j] = ur[j] * ur[j] + ui[j] * ui[j]; }
        for (int i = 0; i < R; ++i)
            for (int j = i + 1; j < R; ++j) {                       // conj(U_ki) U_kj
                const int q = lr_pair(R, i, j);
                P[R + 2 * q] = ur[i] * ur[j] + ui[i] * ui[j];
                P[R + 2 * q + 1] = ur[i] * ui[j] - ui[i] * ur[j];
            }
    }
    for (int j = 0; j < R; ++j) t64[offL + j] = j < r ? 1.0 / lam[idx[j]] : 1.0;      // padded directions: U = 0, S_jj = sigma2, t_j = 0 -> w_j = 0
    t32.resize(size);
    for (int e = 0; e < size; ++e) t32[e] = (float)t64[e];
    return r;
}

}  // namespace wifi
