// wifi_solve_hpd.cu -- per-frame PS_MMSE for Hermitian PSD R (WIFI_SOLVE_HPD):
//
//   A_f = R + diag(d_f),  d_fk = sigma2_f / |tx_k|^2   (Hermitian positive definite when R is Hermitian PSD)
//   A_f z = y = rx/tx,    H = R z = (A_f - D_f) z = y - D_f z
//
// The last identity replaces the 53 x 53 product R z of the formula by 53 multiplies, and it is the better-conditioned
// evaluation: an error dz enters H as D dz (D ~ 1e-10 .. 1e-7) instead of R dz (measured in FP64: 4.6e-12 vs 9.9e-12).
//
// A frame is owned by a PR x 8 grid of lanes (PR = 8: two warps, PR = 4: one warp).  Lane (pr, pc) keeps the
// 2-D-cyclic slice  a[li][lj] = M[PR li + pr][8 lj + pc]  of the bordered Hermitian matrix
//        M = [ A   . ]      (row 53 = y^H carries the right-hand side through the elimination)
//            [ y^H . ]
// in REGISTERS, and only the local positions that can lie on or below the diagonal (28 of 49 for PR = 8, 56 of 98 for
// PR = 4).  Elimination is the symmetric (L D L^H) form without pivoting -- for a Hermitian positive-definite matrix
// the growth factor is 1 -- and is unrolled over groups of PR steps so every register index is static.  Step k:
//   * the lanes of lane-column k%8 publish the raw column  c_i = a_ik  to a double-buffered 64-entry scratch, permuted
//     so that the entries one lane needs afterwards (its rows, its columns) are contiguous 64-byte runs;
//   * one sync (named barrier for two warps, __syncwarp for one);
//   * every lane reads the runs of its rows and its columns and the pivot, forms t_j = conj(c_j)/a_kk for its columns
//     and updates  a_ij -= c_i t_j  on its lower local positions (4 FMAs each).  Positions that are already finished
//     (row or column <= k) are NOT masked: they receive garbage that is never read again;
//   * row k of U' = D^-1 L^H, U'_ki = conj(c_i)/a_kk, goes to shared memory (one contiguous run per step, rows k and
//     52-k folded into one 54-entry line); its entry 53 is the forward-substituted right-hand side.
// The cyclic distribution keeps all lanes busy as the active sub-matrix shrinks.  Back-substitution with the
// unit-diagonal U' runs on one warp (column axpys, z_j broadcast by shuffle) and finishes with H = y - D z from
// registers.  tx/rx of the group's next frame are prefetched during the elimination (registers in FP32, L2 in FP64).
//
// TIO is the storage type of R/tx/rx/sigma2/H and T the arithmetic type: <float,float>, <double,double>, and
// <double,float> = FP32 I/O with FP64 arithmetic (the default for WIFI_F32: d_f is below the FP32 resolution of R).
//
// Two kernels: mmse_hpd_kernel (the layout above; the default for FP32 arithmetic) and mmse_hpd_dmma_kernel further down
// (FP64 arithmetic: the same elimination blocked by two columns with the trailing updates on the FP64 tensor path, DESIGN.md
// 4.2a).  Both share the back-substitution stage.
//
// Replaces the two inverse() calls of main.c:186,201 (utils.c:141-170, O(n^5)) for the intended formula.
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

constexpr int H_N1 = NSC + 1;                   // 54 rows of the bordered matrix
constexpr int H_NLC = 7;                        // local columns (53 columns / 8)
constexpr int H_US = 27 * 54 + 2;               // folded U' store: line m = row m (53-m entries) ++ row 52-m (m+1 entries)
// The covariance image Rt[j * RtStride + i] = R[i][j].  FP32 arithmetic: zero-padded to 56 x 56 so that the per-frame set-up
// of the register slice needs no bounds tests, with a row stride of 58 that makes the slice gather (lanes = 8 consecutive
// j x 4 consecutive i) conflict-free (53, the dense stride, is a 2-way conflict): 77.4 -> 80.7 M frames/s.  FP64 arithmetic
// keeps the dense image and the guarded set-up: the branch-free form has more loads in flight than the 168-register budget
// holds (32 bytes of spills) and measured 3 % slower.
template <typename T> struct RtStride { static constexpr int v = sizeof(T) == 4 ? 58 : NSC; };
template <typename T> struct RtSize { static constexpr int v = ((sizeof(T) == 4 ? 56 : NSC) * RtStride<T>::v + 7) & ~7; };   // keeps the group runs 64-byte aligned

// The published column is stored as 8 runs (run r = rows/columns i with i%8 == r, position i/8).  Runs are RS entries
// apart with RS * sizeof(cx<T>) / 16 odd, so the 8 runs start in 8 different 16-byte bank groups and a 16-byte shared
// load that touches the same position of all 8 runs is conflict-free (RS = 8 would be an 8-way conflict).
template <typename T> struct RunStride { static constexpr int v = sizeof(T) == 4 ? 10 : 9; };

// per-group shared-memory layout, in units of cx<T>
struct HpdSmem {
    static constexpr int US = 0;
    static constexpr int LB = US + H_US;        // [2][80]
    static constexpr int YB = LB + 160;         // [56]
    static constexpr int DB = YB + 56;          // T[56] <= 28 cx
    static constexpr int GROUP = DB + 28;       // 1704 (multiple of 8: every run stays 16-byte aligned)
};
// offset of U'_ki (i > k) in the folded store
__host__ __device__ constexpr int us_off(int k, int i) { return k <= 26 ? k * 54 + i - k - 1 : (52 - k) * 54 + i; }

template <int LANES> __device__ __forceinline__ void group_sync(int id)
{
    if (LANES == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(LANES) : "memory");
}

// 1/x for a positive pivot without the IEEE slow path: hardware approximation + Newton steps
__device__ __forceinline__ float pivot_rcp(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return fmaf(r, fmaf(-x, r, 1.0f), r);
}
__device__ __forceinline__ double pivot_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));      // MUFU.RCP64H: ~20 bits
    r = fma(r, fma(-x, r, 1.0), r);
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

// sign flip on the integer pipe (the FP64 pipe is the DMMA pipe)
__device__ __forceinline__ double dneg(double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }
__device__ __forceinline__ void dmma_m8n8k4(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// two Newton steps (2^-20 -> 2^-40 -> 2^-80): enough for FP64 once the seed has more than 14 good bits
__device__ __forceinline__ double pivot_rcp2(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

// entries [lo, 8) of a 64-byte-aligned run of 8 complex values, with 16-byte shared loads
template <int LO> __device__ __forceinline__ void load_run(const float2 *p, float2 (&v)[8])
{
#pragma unroll
    for (int m = LO / 2; m < 4; ++m) {
        float4 q = *reinterpret_cast<const float4 *>(p + 2 * m);
        v[2 * m] = make_float2(q.x, q.y);
        v[2 * m + 1] = make_float2(q.z, q.w);
    }
}
template <int LO> __device__ __forceinline__ void load_run(const double2 *p, double2 (&v)[8])
{
#pragma unroll
    for (int m = LO; m < 7; ++m) v[m] = p[m];
}

// Per-lane addresses into the published-column scratch (buffer 0; the live buffer is `+ bo`, bo in {0, 80})
template <typename T> struct HpdLane {
    cx<T> *lb;      // scratch base
    cx<T> *rowp;    // run of my rows (run pr; for one-warp groups the odd local rows are in run pr + 4)
    cx<T> *colp;    // run of my columns (run pc)
    cx<T> *usl;     // entry of row i = lane (row lane + 32 is 4 positions further)
    cx<T> *Us;      // folded U' store
    int lane, pr, pc, bar_id;
};

// Elimination steps are grouped by the local row / column block they start in: the PR steps K = PR*G .. PR*G + PR-1
// touch the same register positions (local rows >= G, local columns >= PR*G/8), so one compiled body per group runs as a
// rolled loop over the steps of the group -- the step index only enters through the publisher predicate, the scratch
// buffer parity and the U' row offset.  (A fully unrolled 53-step elimination is 75-120 KB of straight-line code per
// frame and misses in the instruction cache: ncu no_instruction stalls.)
// (Forming the reciprocal of the next pivot on its owner lane at the end of a step and publishing it with the column
// was measured slower: f64 31.2 -> 28.1 M frames/s; every lane recomputes it after the sync instead.)
template <typename T, int PR, int G>
struct HpdGroup {
    static constexpr int LANES = PR * 8;
    static constexpr int NLR = (H_N1 + PR - 1) / PR;
    static constexpr int K0 = PR * G, NQ = (NSC - K0 < PR) ? NSC - K0 : PR;
    static constexpr int klr = G, klc = K0 >> 3;       // first local row / column that can still be live
    static constexpr int RS = RunStride<T>::v;
    static __device__ __forceinline__ bool live(int li, int lj) { return PR * li + PR - 1 >= 8 * lj; }
    static __device__ __forceinline__ int rowpos(int li) { return PR == 8 ? li : (li & 1) * 4 * RS + (li >> 1); }

    static __device__ __forceinline__ void run(cx<T> (&a)[NLR][H_NLC], const HpdLane<T> &L, int &bo)
    {
#pragma unroll 1
        for (int q = 0; q < NQ; ++q) {
            const int K = K0 + q, kc = K & 7;              // kc: owner lane-column of column K
            if (L.pc == kc) {
                if (PR == 4 && sizeof(T) == 4) {
                    // local rows li and li + 2 are neighbours in their run: one 16-byte store for two rows (a finished row
                    // that shares a store with a live one is published too and never read)
#pragma unroll
                    for (int p = 0; p < 2; ++p)
#pragma unroll
                        for (int m = 0; m < 7; m += 2) {
                            const int l0 = 2 * m + p, l1 = l0 + 2;
                            if (l1 < NLR && l1 >= klr && live(l0, klc)) {
                                *reinterpret_cast<float4 *>(&L.rowp[bo + rowpos(l0)]) =
                                    make_float4((float)a[l0][klc].x, (float)a[l0][klc].y, (float)a[l1][klc].x, (float)a[l1][klc].y);
                            } else {
                                if (l0 >= klr && live(l0, klc)) L.rowp[bo + rowpos(l0)] = a[l0][klc];
                                if (l1 < NLR && l1 >= klr && live(l1, klc)) L.rowp[bo + rowpos(l1)] = a[l1][klc];
                            }
                        }
                } else {
#pragma unroll
                    for (int li = klr; li < NLR; ++li)
                        if (live(li, klc)) L.rowp[bo + rowpos(li)] = a[li][klc];          // row PR li + pr
                }
            }
            group_sync<LANES>(L.bar_id);
            const T inv = pivot_rcp(L.lb[bo + kc * RS + klc].x);   // the pivot a_KK of a Hermitian matrix is real
            // my columns: run pc, local columns klc..6
            cx<T> cc[8], t[H_NLC];
            load_run<klc>(L.colp + bo, cc);
#pragma unroll
            for (int lj = klc; lj < H_NLC; ++lj) t[lj] = mk<T>(cc[lj].x * inv, -cc[lj].y * inv);
            // row K of U' (entries i = K+1 .. 53), one contiguous run of the folded store.  (U'_Kj is t_j, which the eight lanes
            // of lane-row 0 already hold: storing it from there saves the loads and multiplies below but its seven strided
            // 8-lane stores were slower, 81.5 -> 76.0 M frames/s in FP32.)
            cx<T> *urow = L.Us + (K <= 26 ? K * 53 - 1 : (52 - K) * 54) + L.lane;
#pragma unroll
            for (int i0 = 0; i0 < H_N1; i0 += LANES) {
                if (i0 + LANES - 1 > K0) {
                    const cx<T> c = L.usl[bo + (i0 >> 3)];
                    if (i0 + L.lane > K && i0 + L.lane < H_N1) urow[i0] = mk<T>(c.x * inv, -c.y * inv);
                }
            }
            // my rows
            if (PR == 8) {
                cx<T> cr[8];
                load_run<klr>(L.rowp + bo, cr);
#pragma unroll
                for (int li = klr; li < NLR; ++li)
#pragma unroll
                    for (int lj = klc; lj < H_NLC; ++lj)
                        if (live(li, lj)) cfms(a[li][lj], cr[li], t[lj]);
            } else {
                // PR == 4: rows 4 li + pr -> run pr (li even) and run pr + 4 (li odd), position li / 2
                cx<T> ce[8], co[8];
                load_run<(klr + 1) / 2>(L.rowp + bo, ce);
                load_run<klr / 2>(L.rowp + bo + 4 * RS, co);
#pragma unroll
                for (int li = klr; li < NLR; ++li) {
                    const cx<T> c = (li & 1) ? co[li >> 1] : ce[li >> 1];
#pragma unroll
                    for (int lj = klc; lj < H_NLC; ++lj)
                        if (live(li, lj)) cfms(a[li][lj], c, t[lj]);
                }
            }
            bo ^= 80;
        }
        HpdGroup<T, PR, G + 1>::run(a, L, bo);
    }
};
template <typename T> struct HpdGroup<T, 8, 7> {
    static __device__ __forceinline__ void run(cx<T> (&)[7][H_NLC], const HpdLane<T> &, int &) {}
};
template <typename T> struct HpdGroup<T, 4, 14> {
    static __device__ __forceinline__ void run(cx<T> (&)[14][H_NLC], const HpdLane<T> &, int &) {}
};

template <typename C> __device__ __forceinline__ C shfl_cx(C v, int src)
{
    v.x = __shfl_sync(0xffffffffu, v.x, src); v.y = __shfl_sync(0xffffffffu, v.y, src);
    return v;
}
template <typename T, typename TIO> __device__ __forceinline__ cx<T> widen(cx<TIO> v) { return mk<T>((T)v.x, (T)v.y); }

// R accessors of the back-substitution stage: the dense transposed image of mmse_hpd_kernel; the caller's global array for
// mmse_hpd_dmma_kernel (whose shared-memory image is in fragment order)
template <typename T> struct RtDense {
    const cx<T> *Rt;
    __device__ __forceinline__ T diag(int k) const { return Rt[k * RtStride<T>::v + k].x; }
    __device__ __forceinline__ cx<T> at(int k, int l) const { return Rt[l * RtStride<T>::v + k]; }    // R[k][l]
};
template <typename TIO> struct RGlobal {                       // the caller's row-major R in global memory (rare path only)
    const cx<TIO> *R;
    __device__ __forceinline__ double diag(int k) const { return (double)R[k * NSC + k].x; }
    __device__ __forceinline__ double2 at(int k, int l) const { const cx<TIO> v = R[k * NSC + l]; return mk<double>((double)v.x, (double)v.y); }
};

// Back-substitution with the unit-diagonal U' of the folded store on ONE warp (lane owns rows lane and lane + 32; entry 53 of
// a row is the forward-substituted right-hand side), then H = y - D z, and H_k = sum_j R_kj z_j for the bins whose noise
// term dominates the diagonal.
template <typename T, typename TIO, typename RA_t>
__device__ __forceinline__ void hpd_backsub_store(const cx<T> *Us, const cx<T> *yb, const T *db, const RA_t RA, cx<TIO> *Hf, int lane)
{
    // lane owns rows i0 = lane and i1 = lane + 32; U'_ij lives at ub + j
    const cx<T> *ub0 = Us + (lane <= 26 ? lane * 53 - 1 : (52 - lane) * 54);
    const cx<T> *ub1 = Us + (lane <= 20 ? (20 - lane) * 54 : 0);
    cx<T> y0 = ub0[NSC], y1 = (lane + 32 < NSC) ? ub1[NSC] : mk<T>(0, 0);
    // (Solving four columns per round -- four shuffles issued together, the 4 x 4 diagonal block solved redundantly by every
    // lane -- halves the serial chain but adds FP64 instructions: measured 5.6 % slower in FP32, 3 % in FP64 on the CUDA-core
    // kernels (issue-bound) and 5 % slower on the DMMA kernel, whose FP64 pipe is shared with the tile updates.)
#pragma unroll 4
    for (int j = NSC - 1; j >= 32; --j) {
        cx<T> zj = shfl_cx(y1, j - 32);
        cfms(y0, ub0[j], zj);
        if (lane + 32 < j) cfms(y1, ub1[j], zj);
    }
#pragma unroll 4
    for (int j = 31; j >= 1; --j) {
        cx<T> zj = shfl_cx(y0, j);
        if (lane < j) cfms(y0, ub0[j], zj);
    }
    const bool second = lane + 32 < NSC;
    const T d0 = db[lane], d1 = second ? db[lane + 32] : (T)0;
    cx<T> h0, h1 = mk<T>(0, 0);
    { const cx<T> y = yb[lane]; h0 = mk<T>(y.x - d0 * y0.x, y.y - d0 * y0.y); }
    if (second) { const cx<T> y = yb[lane + 32]; h1 = mk<T>(y.x - d1 * y1.x, y.y - d1 * y1.y); }
    // Bins whose noise term dominates the diagonal (d_k > R_kk: a bin with almost no transmit energy, the DC bin of the
    // inputs.h frame) have y_k ~ d_k z_k >> H_k, so y - d z cancels (measured 5.5e-10 at |y_dc| = 50 in FP64): they take
    // H_k = sum_j R_kj z_j instead, as a warp reduction (z_j lives in lanes j and j - 32).
    unsigned m0 = __ballot_sync(0xffffffffu, d0 > RA.diag(lane));
    unsigned m1 = __ballot_sync(0xffffffffu, second && d1 > RA.diag(lane + 32));
    while (m0 | m1) {
        const int k = m0 ? __ffs(m0) - 1 : 32 + __ffs(m1) - 1;
        if (m0) m0 &= m0 - 1; else m1 &= m1 - 1;
        cx<T> acc = cmul(RA.at(k, lane), y0);
        if (second) cfma(acc, RA.at(k, lane + 32), y1);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
            acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
        }
        if (k == lane) h0 = acc;
        if (k == lane + 32) h1 = acc;
    }
    st_stream(Hf + lane, mk<TIO>((TIO)h0.x, (TIO)h0.y));
    if (second) st_stream(Hf + lane + 32, mk<TIO>((TIO)h1.x, (TIO)h1.y));
}

template <typename T, typename TIO, int PR, int FPC, int MINB, int VAR>
__global__ void __launch_bounds__(PR * 8 * FPC, MINB)
    mmse_hpd_kernel(const cx<TIO> *__restrict__ R, const cx<TIO> *__restrict__ tx, const cx<TIO> *__restrict__ rx, int64_t frame_stride,
                    const TIO *__restrict__ sigma2, cx<TIO> *__restrict__ H, int64_t n_frames)
{
    constexpr int LANES = PR * 8;
    constexpr int NLR = (H_N1 + PR - 1) / PR;              // local rows: 7 (PR = 8) or 14 (PR = 4)
    extern __shared__ __align__(16) unsigned char hpd_smem[];
    using S = HpdSmem;
    cx<T> *Rt = (cx<T> *)hpd_smem;                         // Rt[j * RTS + i] = R[i][j] (FP32: zero for i or j >= 53)
    constexpr int RTS = RtStride<T>::v, H_RT = RtSize<T>::v;
    const int grp = threadIdx.x / LANES, lane = threadIdx.x % LANES;
    const int pr = lane >> 3, pc = lane & 7;
    cx<T> *gs = Rt + H_RT + grp * S::GROUP;
    cx<T> *Us = gs + S::US, *lb = gs + S::LB, *yb = gs + S::YB;
    T *db = (T *)(gs + S::DB);
    const int bar_id = grp + 1;
    constexpr int RS = RunStride<T>::v;
    const HpdLane<T> L = {lb, lb + pr * RS, lb + pc * RS, lb + (lane & 7) * RS + (lane >> 3), Us, lane, pr, pc, bar_id};

    if (VAR) {
        for (int e = threadIdx.x; e < H_RT; e += LANES * FPC) Rt[e] = mk<T>(0, 0);
        if (lane < 3) { yb[NSC + lane] = mk<T>(0, 0); db[NSC + lane] = (T)0; }     // padding read by the branch-free set-up
        __syncthreads();
    }
    for (int e = threadIdx.x; e < NSC * NSC; e += LANES * FPC) {
        int i = e / NSC, j = e - i * NSC;
        Rt[j * RTS + i] = widen<T, TIO>(R[e]);
    }
    __syncthreads();

    const int64_t fstep = (int64_t)gridDim.x * FPC;
    int64_t f = (int64_t)blockIdx.x * FPC + grp;
    // inputs: lane k (and k + 32 for one-warp groups) holds sub-carrier k.  FP32 arithmetic: tx/rx of the group's next frame
    // are prefetched into registers during the elimination.  FP64 arithmetic: the next frame's rows are only pulled into L2
    // -- holding them in registers (10 of 168) cost more in spills and lost load/FMA overlap inside the elimination than the
    // exposed L2 latency at the top of a frame (~1 % of a frame, covered by the CTA's other groups).  Measured, register
    // prefetch -> L2 prefetch: f64 32.9 -> 34.4 M frames/s, FP32 storage + FP64 arithmetic 33.8 -> 34.7 M, f32 77.4 -> 75.8 M.
    constexpr bool REGPF = sizeof(T) == 4;
    constexpr int NIN = (NSC + LANES - 1) / LANES;
    // y_k = rx/tx and d_k = sigma2/|tx|^2 of sub-carrier k with ONE reciprocal of |tx|^2 (hardware seed + Newton steps, ~1 ulp)
    auto set_inputs = [&](int k, cx<TIO> tv, cx<TIO> rv, TIO sg) {
        const cx<T> t = widen<T, TIO>(tv), r = widen<T, TIO>(rv);
        if (VAR) {
            const T it = pivot_rcp(cabs2(t));
            yb[k] = mk<T>((r.x * t.x + r.y * t.y) * it, (r.y * t.x - r.x * t.y) * it);
            db[k] = (T)sg * it;
        } else {
            yb[k] = cdiv(r, t);
            db[k] = (T)sg / cabs2(t);
        }
    };
    cx<TIO> tin[NIN], rin[NIN];
    TIO sin = (TIO)0;
    if constexpr (REGPF) {
        if (f < n_frames) {
#pragma unroll
            for (int q = 0; q < NIN; ++q) {
                const int k = lane + q * LANES;
                if (k < NSC) { tin[q] = ld_stream(tx + f * frame_stride + k); rin[q] = ld_stream(rx + f * frame_stride + k); }
            }
            sin = sigma2[f];
        }
    }
    for (; f < n_frames; f += fstep) {
        // ---- per-frame inputs ----
        const int64_t fn = f + fstep;
        if constexpr (REGPF) {
#pragma unroll
            for (int q = 0; q < NIN; ++q) {
                const int k = lane + q * LANES;
                if (k < NSC) set_inputs(k, tin[q], rin[q], sin);
            }
            if (fn < n_frames) {
#pragma unroll
                for (int q = 0; q < NIN; ++q) {
                    const int k = lane + q * LANES;
                    if (k < NSC) { tin[q] = ld_stream(tx + fn * frame_stride + k); rin[q] = ld_stream(rx + fn * frame_stride + k); }
                }
                sin = sigma2[fn];
            }
        } else {
            const TIO sg = sigma2[f];
#pragma unroll
            for (int q = 0; q < NIN; ++q) {
                const int k = lane + q * LANES;
                if (k < NSC) set_inputs(k, ld_stream(tx + f * frame_stride + k), ld_stream(rx + f * frame_stride + k), sg);
            }
            if (fn < n_frames && lane < 8) {
                const cx<TIO> *p = (lane < 4 ? tx : rx) + fn * frame_stride;
                asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)p + (lane & 3) * 128));
            }
        }
        group_sync<LANES>(bar_id);
        // local slice of the bordered matrix; only positions that can be on/below the diagonal are ever touched
        // (branch-free: Rt, yb and db are zero-padded; only the positions of the tiles that straddle the diagonal test
        // i == j, only the local row that holds row 53 takes the right-hand side)
        cx<T> a[NLR][H_NLC];
#pragma unroll
        for (int li = 0; li < NLR; ++li) {
            const int i = PR * li + pr;
#pragma unroll
            for (int lj = 0; lj < H_NLC; ++lj) {
                if (PR * li + PR - 1 >= 8 * lj) {
                    const int j = 8 * lj + pc;
                    cx<T> v;
                    if (VAR) {
                        v = Rt[j * RTS + i];
                        if (PR * li <= 8 * lj + 7) v.x += (i == j) ? db[j] : (T)0;       // tile touches the diagonal
                        if (li == NSC / PR) { const cx<T> yv = yb[j]; if (i == NSC) v = cconj(yv); }
                    } else {
                        v = mk<T>(0, 0);
                        if (j < NSC) {
                            if (i < NSC) { v = Rt[j * RTS + i]; if (i == j) v.x += db[i]; }
                            else if (i == NSC) v = cconj(yb[j]);
                        }
                    }
                    a[li][lj] = v;
                }
            }
        }

        // ---- symmetric elimination: compile-time recursion over the step groups (every register index is static) ----
        {
            int bo = 0;
            HpdGroup<T, PR, 0>::run(a, L, bo);
        }
        group_sync<LANES>(bar_id);

        // ---- back-substitution on the first warp of the group, then H = y - D z from registers ----
        if (lane < 32) hpd_backsub_store<T, TIO>(Us, yb, db, RtDense<T>{Rt}, H + f * NSC, lane);
        group_sync<LANES>(bar_id);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// FP64 arithmetic on the FP64 tensor path: the same un-pivoted L D L^H elimination of the bordered matrix, BLOCKED by two
// columns, with the trailing update  A22 -= L (D L^H)  issued as DMMA.8x8x4 (mma.sync.m8n8k4.f64).
//
// A complex rank-2 update is a real rank-4 update -- exactly the k of the instruction -- in the half-embedded form
//      C~ (2m x n) -= L~ (2m x 4) U~ (4 x n),   C~[2i + part][j] = Re/Im a_ij,
//      L~ rows 2i, 2i+1 = [l0.re, -l0.im, l1.re, -l1.im], [l0.im, l0.re, l1.im, l1.re],   U~ rows = u0.re, u0.im, u1.re, u1.im,
// with l_ip = a_ip / d_p and u_pj = conj(a_jp): no redundant flops (a full 2m x 2n embedding would double them).
// A frame is held by two warps as 8 x 8 real tiles = 4 complex rows x 8 complex columns, lower triangle only: warp w owns the
// tile rows I = 2t + w (t = 0..6), tile columns J <= t -- 28 accumulator tiles = 112 registers per lane, the same footprint
// as the CUDA-core layout, and the identical static register structure for both warps.  One block step:
//   (1) the 8 lanes per tile that hold columns K, K+1 write them to the panel buffer P; barrier;
//   (2) lane i factors row i of the panel (both pivots' reciprocals are formed redundantly by every lane from broadcast
//       loads), writes its rows of -L~ and its row of U~ already in fragment order, and its entries of rows K, K+1 of the
//       U' store for the back-substitution; barrier;
//   (3) each warp loads one A fragment per live tile row and one B fragment per live tile column (one conflict-free 8-byte
//       load each: fragment (I) is the 256 contiguous bytes L~[32 I + lane]) and issues one DMMA per live tile.
// One warp instruction carries 256 FMAs with two operand registers per lane, so the elimination needs ~3.5 k instructions
// per warp and frame instead of ~6.7 k, and the FP64 pipe is fed by 28 instructions per block step while the other warps'
// panel work proceeds.  Tiles are grouped by h = K / 8 (four block steps touch the same tiles: t, J >= h) so that every
// register index is static; positions above the diagonal inside a straddling tile, finished rows and the padding rows /
// columns (54, 55 / 53..55) receive garbage that never reaches a live element (an accumulator element only ever feeds
// itself; panel rows <= K+1 and columns >= 53 are written as zeros).
constexpr int DM_LANES = 64;
constexpr int DM_PL = 116;          // plane stride of -L~ in doubles: 112 rows + 4, so the four planes start 8 banks apart
struct DmSmem {                                  // per group, in units of double2
    static constexpr int US = 0;                 // folded U' store, as above
    static constexpr int YB = US + H_US;         // [56]
    static constexpr int DB = YB + 56;           // double[56]
    static constexpr int PB = DB + 28;           // panel columns: P[p][i], p = 0, 1, i < 56
    static constexpr int LT = PB + 112;          // -L~ : double[4][DM_PL], plane k = column k of the A operand
    static constexpr int UT = LT + 232;          // U~  : double[56][4]
    static constexpr int GROUP = UT + 112;       // 2000
};
constexpr int DM_RF = 2 * 28 * 32;          // R in accumulator-fragment order: [warp of the pair][tile t (t + 1) / 2 + J][lane] -> (c0, c1)

struct DmLane {
    double *Pd;         // panel buffer as doubles: Pd[p * 112 + 2 i + part]
    double2 *Pc;        // the same as complex: Pc[p * 56 + i]
    double *Ld, *Ud;    // -L~ and U~ as doubles
    double2 *Us;
    int w, l32, gl, bar_id;     // warp of the pair, lane in the warp, lane in the pair (= panel row)
};

// (1) columns K, K+1 live in tile column HH = K / 8, column pair q = (K & 7) / 2 of the accumulator fragment
template <int HH>
__device__ __forceinline__ void dm_extract(const double (&acc)[7][7][2], const DmLane &L, int q)
{
    if ((L.l32 & 3) == q) {
        double *p0 = L.Pd + 8 * L.w + (L.l32 >> 2);           // 2 i + part = 8 I + (lane >> 2), I = 2 t + w
#pragma unroll
        for (int t = HH; t < 7; ++t) {
            p0[16 * t] = acc[t][HH][0];
            p0[16 * t + 112] = acc[t][HH][1];
        }
    }
}

// (2) panel row i = lane of the pair.  Finished rows / columns and the padding are NOT masked: whatever they put into -L~ and
// U~ only reaches accumulator rows / columns that are never read again.
__device__ __forceinline__ void dm_panel(const DmLane &L, int K)
{
    {
        // lane g of the pair takes row K + 1 + g (rows <= K are finished); lanes past row 55 repeat it -- no branch, so that the
        // compiler can interleave this chain with the DMMAs around it.  From K = 22 on, the first warp covers every live row.
        const int i = K + 1 + L.gl < 56 ? K + 1 + L.gl : 55;
        const double2 pk = L.Pc[K], c1 = L.Pc[K + 1], e1 = L.Pc[56 + K + 1];
        const double2 a0 = L.Pc[i];
        double2 a1 = L.Pc[56 + i];
        // both pivots' reciprocals from ONE reciprocal chain: with det = d_K e - |c|^2 (the 2 x 2 leading minor),
        // 1/d_K = det / (d_K det) and 1/d'_(K+1) = d_K / det = d_K^2 / (d_K det)
        const double d0 = pk.x;
        const double det = K == NSC - 1 ? 1.0 : fma(e1.x, d0, -fma(c1.x, c1.x, c1.y * c1.y));   // (there is no column 53)
        const double r = pivot_rcp2(d0 * det);
        const double ninv0 = -(det * r), ninv1 = -(d0 * d0 * r);
        const double2 t1 = mk<double>(c1.x * ninv0, -c1.y * ninv0);         // -conj(a_(K+1)K) / d_K
        cfma(a1, a0, t1);                                                    // column K+1 after step K
        const double2 nl0 = mk<double>(a0.x * ninv0, a0.y * ninv0), nl1 = mk<double>(a1.x * ninv1, a1.y * ninv1);   // -l
        // rows K, K+1 of U' = D^-1 L^H (entry 53 = forward-substituted right-hand side)
        if (i > K && i < H_N1) L.Us[us_off(K, i)] = mk<double>(dneg(nl0.x), nl0.y);
        if (K + 1 < NSC && i > K + 1 && i < H_N1) L.Us[us_off(K + 1, i)] = mk<double>(dneg(nl1.x), nl1.y);
        // -L~ in four planes (k = 0..3) of rows 2i, 2i+1: a lane's stores and the fragment loads are conflict-free
        double2 *lp = reinterpret_cast<double2 *>(L.Ld + 2 * i);
        lp[0] = mk<double>(nl0.x, nl0.y);
        lp[DM_PL / 2] = mk<double>(dneg(nl0.y), nl0.x);
        lp[DM_PL] = mk<double>(nl1.x, nl1.y);
        lp[3 * DM_PL / 2] = mk<double>(dneg(nl1.y), nl1.x);
        *reinterpret_cast<double4 *>(L.Ud + 4 * i) = make_double4(a0.x, dneg(a0.y), a1.x, dneg(a1.y));
    }
}

// Block steps are grouped by the tile column HH of the NEXT panel: the trailing update of step K = 8 HH - 2 + 2 q touches rows
// and columns >= K + 2 = 8 HH + 2 q, i.e. the tiles t, J >= HH whatever q is, so one compiled body serves the group.  The update
// is split: the tiles of column HH -- they hold the next panel -- go first and are extracted at once; the next panel is then
// factored (barrier, FP64 chain, barrier) while the DMMAs of the remaining tiles, whose fragments are already in registers,
// drain through the same pipe.
template <int HH>
struct DmGroup {
    static __device__ __forceinline__ void run(double (&acc)[7][7][2], const DmLane &L)
    {
        constexpr int Q0 = HH == 0 ? 1 : 0, Q1 = HH == 6 ? 3 : 4;
#pragma unroll 1
        for (int q = Q0; q < Q1; ++q) {
            // fragments of step K = 8 HH - 2 + 2 q (panel already factored and published)
            double A[7], B[7];
#pragma unroll
            for (int t = HH; t < 7; ++t) A[t] = L.Ld[(L.l32 & 3) * DM_PL + 8 * (2 * t + L.w) + (L.l32 >> 2)];
#pragma unroll
            for (int J = HH; J < 7; ++J) B[J] = L.Ud[32 * J + L.l32];
#pragma unroll
            for (int t = HH; t < 7; ++t) dmma_m8n8k4(acc[t][HH][0], acc[t][HH][1], A[t], B[HH]);
            dm_extract<HH>(acc, L, q);
            group_sync<DM_LANES>(L.bar_id);
            const int Kn = 8 * HH + 2 * q;
            // the second warp's rows (K + 33 ..) are all finished from K = 22 on: it skips the panel's FP64 work (the FP64 pipe
            // is the DMMA pipe).  The tile updates are repeated in both branches so that the panel chain and the DMMAs stay in
            // one basic block.
            if (HH < 2 || L.w == 0 || Kn < 22) {
                dm_panel(L, Kn);
#pragma unroll
                for (int t = HH + 1; t < 7; ++t)
#pragma unroll
                    for (int J = HH + 1; J <= t; ++J) dmma_m8n8k4(acc[t][J][0], acc[t][J][1], A[t], B[J]);
            } else {
#pragma unroll
                for (int t = HH + 1; t < 7; ++t)
#pragma unroll
                    for (int J = HH + 1; J <= t; ++J) dmma_m8n8k4(acc[t][J][0], acc[t][J][1], A[t], B[J]);
            }
            group_sync<DM_LANES>(L.bar_id);
        }
        DmGroup<HH + 1>::run(acc, L);
    }
};
template <> struct DmGroup<7> {
    static __device__ __forceinline__ void run(double (&)[7][7][2], const DmLane &) {}
};

template <typename TIO, int FPC>
__global__ void __launch_bounds__(DM_LANES * FPC, 1)
    mmse_hpd_dmma_kernel(const cx<TIO> *__restrict__ R, const cx<TIO> *__restrict__ tx, const cx<TIO> *__restrict__ rx, int64_t frame_stride,
                         const TIO *__restrict__ sigma2, cx<TIO> *__restrict__ H, int64_t n_frames)
{
    using S = DmSmem;
    extern __shared__ __align__(16) unsigned char hpd_smem[];
    double2 *Rf = (double2 *)hpd_smem;
    const int grp = threadIdx.x / DM_LANES, gl = threadIdx.x % DM_LANES, w = gl >> 5, l32 = gl & 31;
    double2 *gs = Rf + DM_RF + grp * S::GROUP;
    double2 *Us = gs + S::US, *yb = gs + S::YB;
    double *db = (double *)(gs + S::DB);
    const DmLane L = {(double *)(gs + S::PB), gs + S::PB, (double *)(gs + S::LT), (double *)(gs + S::UT), Us, w, l32, gl, grp + 1};

    // R once per CTA, already in the order the accumulator fragments want it: the per-frame set-up is one conflict-free
    // 16-byte load per tile (elements above the diagonal carry the true R values and are never used; padding is zero)
    for (int e = threadIdx.x; e < DM_RF; e += DM_LANES * FPC) {
        const int el = e & 31, tl = (e >> 5) % 28, ew = e / (32 * 28);
        int t = 0;
        while ((t + 1) * (t + 2) / 2 <= tl) ++t;
        const int J = tl - t * (t + 1) / 2;
        const int i = 4 * (2 * t + ew) + (el >> 3), pt = (el >> 2) & 1, j0 = 8 * J + 2 * (el & 3);
        double2 v = mk<double>(0, 0);
        if (i < NSC) {
            if (j0 < NSC) { const cx<TIO> r0 = R[i * NSC + j0]; v.x = (double)(pt ? r0.y : r0.x); }
            if (j0 + 1 < NSC) { const cx<TIO> r1 = R[i * NSC + j0 + 1]; v.y = (double)(pt ? r1.y : r1.x); }
        }
        Rf[e] = v;
    }
    if (gl < 3) { yb[NSC + gl] = mk<double>(0, 0); db[NSC + gl] = 0.0; }
    for (int e = gl; e < S::GROUP - S::PB; e += DM_LANES) gs[S::PB + e] = mk<double>(0, 0);
    __syncthreads();

    // accumulator element (tile t, J) of this lane: complex row 4 (2t + w) + (l32 >> 3), part (l32 >> 2) & 1, columns 8 J + 2 (l32 & 3), +1
    const int part = (l32 >> 2) & 1, rsub = l32 >> 3, csub = 2 * (l32 & 3);
    const int64_t fstep = (int64_t)gridDim.x * FPC;
    for (int64_t f = (int64_t)blockIdx.x * FPC + grp; f < n_frames; f += fstep) {
        // ---- per-frame inputs: y = rx/tx, d = sigma2/|tx|^2 ----
        if (gl < NSC) {
            const cx<double> t = widen<double, TIO>(ld_stream(tx + f * frame_stride + gl)), r = widen<double, TIO>(ld_stream(rx + f * frame_stride + gl));
            const double it = pivot_rcp(cabs2(t));            // one reciprocal (~1 ulp) serves rx/tx and sigma2/|tx|^2
            yb[gl] = mk<double>((r.x * t.x + r.y * t.y) * it, (r.y * t.x - r.x * t.y) * it);
            db[gl] = (double)sigma2[f] * it;
        }
        {
            const int64_t fn = f + fstep;
            if (fn < n_frames && gl < 8) {
                const cx<TIO> *p = (gl < 4 ? tx : rx) + fn * frame_stride;
                asm volatile("prefetch.global.L2 [%0];" ::"l"((const char *)p + (gl & 3) * 128));
            }
        }
        group_sync<DM_LANES>(L.bar_id);
        // ---- accumulator tiles of the bordered matrix [A; y^H], lower triangle ----
        double acc[7][7][2];
        const double2 *rf = Rf + w * (28 * 32) + l32;
#pragma unroll
        for (int t = 0; t < 7; ++t) {
            const int i = 4 * (2 * t + w) + rsub;
#pragma unroll
            for (int J = 0; J <= t; ++J) {
                const int j0 = 8 * J + csub, j1 = j0 + 1;
                const double2 v = rf[(t * (t + 1) / 2 + J) * 32];
                double v0 = v.x, v1 = v.y;
                if (J == t) {                                            // the diagonal runs through tile column t only
                    const double dd = part ? 0.0 : db[i];                // (zero beyond row 52)
                    if (i == j0) v0 += dd;
                    if (i == j1) v1 += dd;
                }
                if (t == 6) {                                            // rows 52..55 (w = 1): row 53 is the right-hand side
                    const double2 y0 = yb[j0], y1 = yb[j1];              // zero beyond column 52
                    if (i == NSC) { v0 = part ? -y0.y : y0.x; v1 = part ? -y1.y : y1.x; }
                }
                acc[t][J][0] = v0;
                acc[t][J][1] = v1;
            }
        }
        // ---- blocked elimination: panel 0, then the look-ahead loop ----
        dm_extract<0>(acc, L, 0);
        group_sync<DM_LANES>(L.bar_id);
        dm_panel(L, 0);
        group_sync<DM_LANES>(L.bar_id);
        DmGroup<0>::run(acc, L);
        // ---- back-substitution on one warp of the pair, then H = y - D z ----
        // (on the SECOND warp: the first one carries every panel factorization from K = 22 on, this balances the FP64 work of
        // the pair's two schedulers)
        if (w == 1) hpd_backsub_store<double, TIO>(Us, yb, db, RGlobal<TIO>{R}, H + f * NSC, l32);
        group_sync<DM_LANES>(L.bar_id);
    }
}

template <typename TIO, int FPC>
static cudaError_t launch_hpd_dmma(const void *R, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, void *H,
                                   int64_t n_frames, cudaStream_t s)
{
    size_t smem = sizeof(double2) * (DM_RF + FPC * DmSmem::GROUP);
    auto kern = mmse_hpd_dmma_kernel<TIO, FPC>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t need = (n_frames + FPC - 1) / FPC;
    unsigned grid = (unsigned)std::min<int64_t>(need, 148);
    kern<<<grid, DM_LANES * FPC, smem, s>>>((const cx<TIO> *)R, (const cx<TIO> *)tx, (const cx<TIO> *)rx, frame_stride,
                                            (const TIO *)sigma2, (cx<TIO> *)H, n_frames);
    return cudaGetLastError();
}

template <typename T, typename TIO, int PR, int FPC, int MINB, int VAR = (sizeof(T) == 4)>
static cudaError_t launch_hpd(const void *R, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, void *H,
                              int64_t n_frames, cudaStream_t s)
{
    using S = HpdSmem;
    size_t smem = sizeof(cx<T>) * (RtSize<T>::v + FPC * S::GROUP);
    auto kern = mmse_hpd_kernel<T, TIO, PR, FPC, MINB, VAR>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t need = (n_frames + FPC - 1) / FPC;
    unsigned grid = (unsigned)std::min<int64_t>(need, 148 * MINB);
    kern<<<grid, PR * 8 * FPC, smem, s>>>((const cx<TIO> *)R, (const cx<TIO> *)tx, (const cx<TIO> *)rx, frame_stride,
                                          (const TIO *)sigma2, (cx<TIO> *)H, n_frames);
    return cudaGetLastError();
}

// Kernel choice (measured on B200, 256 Ki frames; the alternatives that lost are gone from the build):
//   FP64 arithmetic (WIFI_F64, and WIFI_F32 by default: FP32 storage, FP64 arithmetic -- the FP32-I/O mode that meets the 1e-4
//   bound): mmse_hpd_dmma_kernel, 6 frames per CTA (5: 31.7 M frames/s, 4: 31.9, 2: 22.6; the CUDA-core layout <8,6>: 34.5).
//   FP32 arithmetic (WIFI_SOLVE_FAST32, documented accuracy 4e-3): mmse_hpd_kernel<float, float, 4, 12> (<8,8> 60.2, <4,8> 68.9).
cudaError_t launch_mmse_perframe_hpd(wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                                     const void *sigma2, void *H, int64_t n_frames, int fast32, cudaStream_t s)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    if (dt == WIFI_F32 && fast32) return launch_hpd<float, float, 4, 12, 1>(R, tx, rx, frame_stride, sigma2, H, n_frames, s);
    if (dt == WIFI_F32) return launch_hpd_dmma<float, 6>(R, tx, rx, frame_stride, sigma2, H, n_frames, s);
    return launch_hpd_dmma<double, 6>(R, tx, rx, frame_stride, sigma2, H, n_frames, s);
}

}  // namespace wifi
