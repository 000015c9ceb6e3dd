// wifi_inverse_tc.cu -- batched inverse() (utils.c:141-170) for orders 33..64, and the general (pivoted) per-frame PS_MMSE solve at
// the end of the file, with the trailing updates on the tensor cores.
//
// TWO WARPS PER MATRIX, the matrix resident in shared memory, no CTA barrier (named 64-thread barriers per pair): the blocked
// in-place Gauss-Jordan with implicit partial pivoting of tests/test_inverse_blocked_model.py,
//        A  <-  A - C' R~      per NB pivot columns (the panel columns then take the factored panel),
// C' = the NB multiplier vectors (+ e_(r_s): stands in for zeroing the pivot rows), R~ = the NB pivot rows after a unit-lower-
// triangular transform.  A complex rank-NB update is a real rank-2NB product: with the matrix stored as real planes -- per row
// tile the real parts of TR complex rows followed by their imaginary parts --
//        [Re A; Im A]  -=  [Re C', -Im C'; Im C', Re C'] [Re R~; Im R~]
// has exactly the shape of one warp-level MMA per tile and block step, and a lane builds its whole A (B) fragment from ONE
// complex element of C' (R~):
//   FP32  mma.sync.m16n8k8.tf32, NB = 4, 3xTF32 (hi/lo split of both operands, FP32 accumulate: FP32-level accuracy); row tile =
//         8 complex rows.  A fragment of lane (g, t): a0 = a3 = Re c, a1 = -a2 = Im c, c = C'[8 rt + g][t]; B: b0, b1 = Re, Im of
//         R~[t][8 ct + g]; C: (c0, c1) = Re, (c2, c3) = Im of a[8 rt + g][8 ct + 2t, +1].
//   FP64  mma.sync.m8n8k4.f64 (DMMA), NB = 2; row tile = 4 complex rows.  Lane (g, t): A = +-Re/Im of C'[4 rt + (g & 3)][t & 1],
//         B = Re/Im of R~[t & 1][8 ct + g], C = a[4 rt + (g & 3)][8 ct + 2t, +1] (Re for g < 4, Im for g >= 4).
// Per block step: (1) warp 0 reads the panel columns (lane = rows lane, lane + 32) and (2) factors the panel in registers -- per
// column one redux.sync arg-max over the rows not used yet, the pivot row's NB values by shuffle, one reciprocal -- and publishes
// -C'; (3) both warps read and transform the NB pivot rows (warp w = columns lane + 32 w); (4) the tile updates, row tiles split
// between the warps, accumulators streamed through shared memory (conflict-free row stride) in a software pipeline; (5) warp 0
// writes the factored panel back.  FP64 groups two panels into an outer block so that all but one tile column make ONE pass
// through shared memory per four pivot columns (the kernel is bound by shared-memory wavefronts; see the loop comment).
// History (order 53, matrices/s through the API, FP32 / FP64): CTA per matrix, matrix in registers, FFMA / DFMA updates (rounds
// 1-2) 9.3 M / 4.8 M -> one warp per matrix, tensor-core updates 13.5 M / 6.1 M -> + software-pipelined accumulator streaming,
// batched global loads, fused pivot-row multiplier 18.2 M / 6.9 M -> + warp pairs (order 64: 9.1 -> 11.4 M) -> + outer blocks in
// FP64 17.8 M / 8.4 M.  Measured and rejected: outer blocks in FP32 (7 instead of 8 matrices per SM: 13.7 M), four panels per
// outer block in FP64 (3 instead of 4 matrices per SM: 6.6 M).
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

template <typename T> struct IwT;
template <> struct IwT<float> { static constexpr int NB = 4, TR = 8, NP = 1; };     // pivot columns per inner panel, complex rows per row tile,
template <> struct IwT<double> { static constexpr int NB = 2, TR = 4, NP = 2; };    // inner panels per outer block

template <typename T, int NT> struct IwLayout {
    static constexpr int NB = IwT<T>::NB, TR = IwT<T>::TR, NP = IwT<T>::NP;
    static constexpr int N8 = 8 * NT;                                     // padded order (column tiles of 8)
    static constexpr int LD = N8 + ((N8 % 16 == 0) ? 8 : 0);              // row stride of the planes: LD mod 16 == 8 -> conflict-free accumulator tiles
    static constexpr int ROWS = 2 * N8;
    static constexpr int RLD = N8 + 4;                                    // row stride of R~ in complex values (RLD mod 8 == 4: conflict-free B fragments)
    static constexpr size_t BYTES = sizeof(T) * ((size_t)ROWS * LD + 2 * NP * N8 * NB + 2 * NP * NB * RLD) + 2 * 64 + 16;     // + rowof, kof, rs (bytes)
    static constexpr int WPC = (227 * 1024) / BYTES < 8 ? (int)((227 * 1024) / BYTES) : 8;       // matrices (warp pairs) per CTA: what 227 KB hold
};

__device__ __forceinline__ float iw_rcp(float d)
{
    float x;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(x) : "f"(d));
    return fmaf(x, fmaf(-d, x, 1.0f), x);
}
__device__ __forceinline__ double iw_rcp(double d)
{
    double x;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
    x = fma(x, fma(-d, x, 1.0), x);
    return fma(x, fma(-d, x, 1.0), x);
}
__device__ __forceinline__ unsigned iw_bits(float v) { return __float_as_uint(v); }
__device__ __forceinline__ unsigned iw_bits(double v) { return (unsigned)__double2hiint(v); }

__device__ __forceinline__ void split_tf32(float x, uint32_t &hi, uint32_t &lo)
{
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(x));
    const float r = x - __uint_as_float(hi);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo) : "f"(r));
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f64(double (&c)[2], double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[0]), "+d"(c[1]) : "d"(a), "d"(b));
}

// Accumulator tiles go through ld/st.volatile.shared: ptxas keeps volatile accesses in program order, which pins the software
// pipeline of iw_update (with plain accesses it re-serialises every tile into LDS -> MMA -> STS on two register sets).
__device__ __forceinline__ float2 lds_acc(const float *p)
{
    float2 v;
    asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ void sts_acc(float *p, float x, float y)
{
    asm volatile("st.volatile.shared.v2.f32 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "f"(x), "f"(y) : "memory");
}
__device__ __forceinline__ double2 lds_acc(const double *p)
{
    double2 v;
    asm volatile("ld.volatile.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"((uint32_t)__cvta_generic_to_shared(p)));
    return v;
}
__device__ __forceinline__ void sts_acc(double *p, double x, double y)
{
    asm volatile("st.volatile.shared.v2.f64 [%0], {%1, %2};" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "d"(x), "d"(y) : "memory");
}

// NB consecutive plane values (one panel row, real or imaginary parts) as one 16-byte access
__device__ __forceinline__ void ld_panel(const float *p, float (&v)[4]) { const float4 q = *reinterpret_cast<const float4 *>(p); v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w; }
__device__ __forceinline__ void ld_panel(const double *p, double (&v)[2]) { const double2 q = *reinterpret_cast<const double2 *>(p); v[0] = q.x; v[1] = q.y; }
__device__ __forceinline__ void st_panel(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st_panel(double *p, const double (&v)[2]) { *reinterpret_cast<double2 *>(p) = make_double2(v[0], v[1]); }
// one row of -C' (NB complex values) as 16-byte stores
__device__ __forceinline__ void st_crow(float2 *p, const float2 (&c)[4])
{
    reinterpret_cast<float4 *>(p)[0] = make_float4(c[0].x, c[0].y, c[1].x, c[1].y);
    reinterpret_cast<float4 *>(p)[1] = make_float4(c[2].x, c[2].y, c[3].x, c[3].y);
}
__device__ __forceinline__ void st_crow(double2 *p, const double2 (&c)[2]) { p[0] = c[0]; p[1] = c[1]; }

// plane row of the real part of complex row i (the imaginary part sits TR rows below)
template <int TR> __device__ __forceinline__ int erow(int i) { return (i / TR) * (2 * TR) + (i % TR); }

// ---- tile updates, FP32: row tiles RT0 .. RT0+NR-1 (8 complex rows each) x NC column tiles, block steps KS0 .. KS0+NKS-1, 3xTF32 ----
// Column tile i of the call is cfirst + i, stepping over cskip.  Software-pipelined over the column tiles: the accumulators of the
// next tile column are loaded and their MMAs issued BEFORE the results of the current one are stored, so a store never waits for a
// tensor-core result (ncu, first version: a quarter of all stall samples sat on the STS behind the HMMA chain).
template <int NT, int RT0, int NR, int KS0, int NKS, int NC>
__device__ __forceinline__ void iw_update(float *M, const float2 *Cn, const float2 *Rho, int lane, int cfirst, int cskip)
{
    using L = IwLayout<float, NT>;
    const int g = lane >> 2, t = lane & 3;
    uint32_t arh[NKS][NR], arl[NKS][NR], aih[NKS][NR], ail[NKS][NR];
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks)
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const float2 c = Cn[(KS0 + ks) * L::N8 * 4 + (8 * (RT0 + q) + g) * 4 + t];        // -C'[8 rt + g][4 ks + t]
            split_tf32(c.x, arh[ks][q], arl[ks][q]);
            split_tf32(c.y, aih[ks][q], ail[ks][q]);
        }
    float *const p0 = M + ((16 * RT0 + g) * L::LD + 2 * t);
    const float2 *const r0 = Rho + (KS0 * 4 + t) * L::RLD + g;
    auto col = [&](int i) { const int c = cfirst + i; return c + (c >= cskip ? 1 : 0); };
    float acc[2][NR][4];
    auto issue = [&](int ct, float (&a)[NR][4]) {
        uint32_t brh[NKS], brl[NKS], bih[NKS], bil[NKS];
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            const float2 b = r0[ks * 4 * L::RLD + 8 * ct];
            split_tf32(b.x, brh[ks], brl[ks]);
            split_tf32(b.y, bih[ks], bil[ks]);
        }
        const float *p = p0 + 8 * ct;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const float2 re = lds_acc(p + q * 16 * L::LD), im = lds_acc(p + (q * 16 + 8) * L::LD);
            a[q][0] = re.x; a[q][1] = re.y; a[q][2] = im.x; a[q][3] = im.y;
        }
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
#pragma unroll
            for (int q = 0; q < NR; ++q) mma_tf32(a[q], arl[ks][q], ail[ks][q], ail[ks][q] ^ 0x80000000u, arl[ks][q], brh[ks], bih[ks]);       // A_lo B_hi
#pragma unroll
            for (int q = 0; q < NR; ++q) mma_tf32(a[q], arh[ks][q], aih[ks][q], aih[ks][q] ^ 0x80000000u, arh[ks][q], brl[ks], bil[ks]);       // A_hi B_lo
#pragma unroll
            for (int q = 0; q < NR; ++q) mma_tf32(a[q], arh[ks][q], aih[ks][q], aih[ks][q] ^ 0x80000000u, arh[ks][q], brh[ks], bih[ks]);       // A_hi B_hi
        }
    };
    issue(col(0), acc[0]);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        if (i + 1 < NC) issue(col(i + 1), acc[(i + 1) & 1]);
        float *p = p0 + 8 * col(i);
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            sts_acc(p + q * 16 * L::LD, acc[i & 1][q][0], acc[i & 1][q][1]);
            sts_acc(p + (q * 16 + 8) * L::LD, acc[i & 1][q][2], acc[i & 1][q][3]);
        }
    }
}

// ---- tile updates, FP64: row tiles RT0 .. RT0+NR-1 (4 complex rows each; 2 NT in all), one DMMA per tile and block step ----
template <int NT, int RT0, int NR, int KS0, int NKS, int NC>
__device__ __forceinline__ void iw_update(double *M, const double2 *Cn, const double2 *Rho, int lane, int cfirst, int cskip)
{
    using L = IwLayout<double, NT>;
    const int g = lane >> 2, t = lane & 3;
    // A[g][k]: Re rows (g < 4): k < 2 -> Re c, k >= 2 -> -Im c;  Im rows: k < 2 -> Im c, k >= 2 -> Re c;  c = -C'[4 rt + (g & 3)][2 ks + (k & 1)]
    const int apart = (g >> 2) ^ (t >> 1);
    const bool aneg = g < 4 && t >= 2;
    double af[NKS][NR];
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks)
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const double v = reinterpret_cast<const double *>(Cn + (KS0 + ks) * L::N8 * 2 + (4 * (RT0 + q) + (g & 3)) * 2 + (t & 1))[apart];
            af[ks][q] = aneg ? -v : v;
        }
    double *const p0 = M + ((8 * RT0 + g) * L::LD + 2 * t);
    const double *const r0 = reinterpret_cast<const double *>(Rho + (KS0 * 2 + (t & 1)) * L::RLD + g) + (t >> 1);
    auto col = [&](int i) { const int c = cfirst + i; return c + (c >= cskip ? 1 : 0); };
    double acc[2][NR][2];
    auto issue = [&](int ct, double (&a)[NR][2]) {
        double bf[NKS];
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) bf[ks] = r0[ks * 4 * L::RLD + 16 * ct];              // R~[2 ks + (t & 1)][8 ct + g], Re (t < 2) or Im
        const double *p = p0 + 8 * ct;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const double2 v = lds_acc(p + q * 8 * L::LD);
            a[q][0] = v.x; a[q][1] = v.y;
        }
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks)
#pragma unroll
            for (int q = 0; q < NR; ++q) mma_f64(a[q], af[ks][q], bf[ks]);
    };
    issue(col(0), acc[0]);
#pragma unroll
    for (int i = 0; i < NC; ++i) {
        if (i + 1 < NC) issue(col(i + 1), acc[(i + 1) & 1]);
        double *p = p0 + 8 * col(i);
#pragma unroll
        for (int q = 0; q < NR; ++q) sts_acc(p + q * 8 * L::LD, acc[i & 1][q][0], acc[i & 1][q][1]);
    }
}

// the tile updates of one matrix split between the two warps of its pair by row tiles: warp 0 (which also factors the panels) takes
// the smaller half.  KS0 / NKS: block steps applied; NC column tiles starting at cfirst, stepping over cskip.
template <int NT, int KS0, int NKS, int NC>
__device__ __forceinline__ void iw_update_half(float *M, const float2 *Cn, const float2 *Rho, int lane, int w, int cfirst, int cskip)
{
    if (w == 0) iw_update<NT, 0, NT / 2, KS0, NKS, NC>(M, Cn, Rho, lane, cfirst, cskip);
    else iw_update<NT, NT / 2, NT - NT / 2, KS0, NKS, NC>(M, Cn, Rho, lane, cfirst, cskip);
}
template <int NT, int KS0, int NKS, int NC>
__device__ __forceinline__ void iw_update_half(double *M, const double2 *Cn, const double2 *Rho, int lane, int w, int cfirst, int cskip)
{
    if (w == 0) iw_update<NT, 0, NT - 1, KS0, NKS, NC>(M, Cn, Rho, lane, cfirst, cskip);     // 2 NT row tiles of 4 complex rows: NT - 1 and NT + 1
    else iw_update<NT, NT - 1, NT + 1, KS0, NKS, NC>(M, Cn, Rho, lane, cfirst, cskip);
}

__device__ __forceinline__ void pair_sync(int pair) { asm volatile("bar.sync %0, 64;" ::"r"(pair + 1) : "memory"); }

// the panel columns K .. K+NB-1 take the factored panel (warp 0, rows lane and lane + 32; warp 0 is also the only reader of the next panel)
template <typename T, int NT> __device__ __forceinline__ void iw_store_panel(T *M, int K, int lane, const cx<T> (&pc)[2][IwT<T>::NB])
{
    using L = IwLayout<T, NT>;
    constexpr int NB = L::NB;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        if (i < L::N8) {
            T *pr = M + erow<L::TR>(i) * L::LD + K;
            T vr[NB], vi[NB];
#pragma unroll
            for (int u = 0; u < NB; ++u) { vr[u] = pc[h][u].x; vi[u] = pc[h][u].y; }
            st_panel(pr, vr); st_panel(pr + L::TR * L::LD, vi);
        }
    }
    __syncwarp();
}

template <typename T, int NT>
__global__ void __launch_bounds__(64 * IwLayout<T, NT>::WPC) cinverse_warp_kernel(const cx<T> *__restrict__ A, int n, cx<T> *__restrict__ Y, int *info, int64_t batch)
{
    using L = IwLayout<T, NT>;
    constexpr int NB = L::NB, NP = L::NP, OB = NB * NP, TR = L::TR, N8 = L::N8, LD = L::LD;
    extern __shared__ __align__(16) unsigned char iw_smem[];
    const int pair = threadIdx.x >> 6, w = (threadIdx.x >> 5) & 1, lane = threadIdx.x & 31;
    const int64_t mat = (int64_t)blockIdx.x * (blockDim.x >> 6) + pair;   // two warps = one matrix; pairs never meet
    if (mat >= batch) return;
    T *M = reinterpret_cast<T *>(iw_smem + pair * L::BYTES);              // [2 N8][LD] real planes
    cx<T> *Cn = reinterpret_cast<cx<T> *>(M + (size_t)L::ROWS * LD);     // [NP][N8][NB]: -C' of the inner panels of one outer block
    cx<T> *Rho = Cn + NP * N8 * NB;                                      // [OB][RLD]: transformed pivot rows of one outer block
    unsigned char *rowof = reinterpret_cast<unsigned char *>(Rho + OB * L::RLD), *kof = rowof + 64, *rs = kof + 64;
    const cx<T> *Ab = A + mat * n * n;
    cx<T> *Yb = Y + mat * n * n;
    const cx<T> zero = mk<T>((T)0, (T)0), one = mk<T>((T)1, (T)0);

    // ---- load: [A, 0; 0, I] into the planes (rows / columns n .. N8-1 carry a unit diagonal and pivot on themselves) ----
    // (lane = columns lane, lane + 32; eight rows = sixteen loads in flight before the first store; the warps alternate row groups)
    for (int i0 = 8 * w; i0 < N8; i0 += 16) {
        cx<T> v[8][2];
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = i0 + q, j = lane + 32 * h;
                v[q][h] = (i == j) ? one : zero;
                if (i < n && j < n) v[q][h] = Ab[i * n + j];
            }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int r = erow<TR>(i0 + q);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = lane + 32 * h;
                if (j < N8) { M[r * LD + j] = v[q][h].x; M[(r + TR) * LD + j] = v[q][h].y; }
            }
        }
    }
    if (w == 0) {
        rowof[lane] = (unsigned char)lane; rowof[lane + 32] = (unsigned char)(lane + 32);
        kof[lane] = (unsigned char)lane; kof[lane + 32] = (unsigned char)(lane + 32);
    }
    unsigned usedw = 0;                                                  // warp 0; bit h: row lane + 32 h has been a pivot row (or lies beyond N8)
    if (lane >= N8) usedw |= 1u;
    if (lane + 32 >= N8) usedw |= 2u;
    int bad = 0;
    pair_sync(pair);

    // An OUTER block = NP inner panels of NB pivot columns inside one tile column ct0.  Each inner panel is factored, its pivot
    // rows are transformed, and its rank-NB update is applied to tile column ct0 ONLY (the next inner panel lives there); the other
    // NT - 1 tile columns take the rank-(NP NB) update of the whole outer block in ONE pass over their accumulators -- half the
    // shared-memory traffic of one pass per inner panel (the limiter, ncu: 54 % / 73 % of the LSU wavefront budget in FP32 / FP64).
    // Pivot row s of the outer block, rho^(s) = a[r_s] - sum_{t<s} c^(t)[r_s] rho^(t): in the columns of ct0 the terms of EARLIER
    // inner panels are already in a[r_s] (narrow updates), elsewhere none is (tests/test_inverse_blocked_model.py, lookahead form).
#pragma unroll 1
    for (int K0 = 0; K0 < N8; K0 += OB) {
        const int ct0 = K0 >> 3;
        const int jmine = lane + 32 * w;
        const bool in0 = (jmine >> 3) == ct0;
        cx<T> rho[OB];
        cx<T> pc[2][NB];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int K = K0 + p * NB;
            if (w == 0) {
                // ---- 1. panel columns K .. K+NB-1 of rows lane, lane + 32 ----
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = lane + 32 * h;
#pragma unroll
                    for (int u = 0; u < NB; ++u) pc[h][u] = zero;
                    if (i < N8) {
                        const T *pr = M + erow<TR>(i) * LD + K;
                        T vr[NB], vi[NB];
                        ld_panel(pr, vr); ld_panel(pr + TR * LD, vi);
#pragma unroll
                        for (int u = 0; u < NB; ++u) pc[h][u] = mk<T>(vr[u], vi[u]);
                    }
                }
                // ---- 2. NB scalar Gauss-Jordan steps on the panel, in registers ----
                cx<T> cn[2][NB];
#pragma unroll
                for (int s_ = 0; s_ < NB; ++s_) {
                    unsigned key = 0;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const unsigned kh = 0x80000000u | (iw_bits(cabs2(pc[h][s_])) & 0x7fffffc0u) | (unsigned)(63 - (lane + 32 * h));
                        if (!((usedw >> h) & 1u) && kh > key) key = kh;
                    }
                    const int r = 63 - (int)(__reduce_max_sync(0xffffffffu, key) & 63u);
                    const int hr = r >> 5, ol = r & 31;
                    cx<T> prow[NB];
#pragma unroll
                    for (int u = 0; u < NB; ++u) {
                        const cx<T> v = hr ? pc[1][u] : pc[0][u];
                        prow[u].x = __shfl_sync(0xffffffffu, v.x, ol);
                        prow[u].y = __shfl_sync(0xffffffffu, v.y, ol);
                    }
                    const cx<T> piv = prow[s_];
                    const T den = cabs2(piv), rden = iw_rcp(den);
                    const cx<T> inv = mk<T>(piv.x * rden, -piv.y * rden);
                    bad |= !(den > (T)0);
                    // Column s of C' is c + e_(r_s) with c_i = a_is / p and c_(r_s) = -1 / p, i.e. C'_(r_s) = (p - 1) / p: with THAT multiplier
                    // the pivot row needs no zeroing -- rho_u - ((p - 1) / p) rho_u = rho_u / p -- and only the pivot column itself is
                    // special (a_is <- -c_i, a_(r_s)s <- 1 / p).
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const bool mine = (lane + 32 * h) == r;
                        cx<T> tq = pc[h][s_];
                        if (mine) tq.x -= (T)1;
                        const cx<T> c = cmul(tq, inv);
                        cn[h][s_] = mk<T>(-c.x, -c.y);
#pragma unroll
                        for (int u = 0; u < NB; ++u)
                            if (u != s_) cfms(pc[h][u], c, prow[u]);
                        pc[h][s_] = mine ? inv : cn[h][s_];
                        if (mine) usedw |= 1u << h;
                    }
                    if (lane == 0) { rowof[K + s_] = (unsigned char)r; kof[r] = (unsigned char)(K + s_); rs[p * NB + s_] = (unsigned char)r; }
                }
#pragma unroll
                for (int h = 0; h < 2; ++h)
                    if (lane + 32 * h < N8) st_crow(Cn + p * N8 * NB + (lane + 32 * h) * NB, cn[h]);
            }
            pair_sync(pair);
            // ---- 3. pivot rows of this inner panel, warp w = columns lane + 32 w ----
#pragma unroll
            for (int s_ = 0; s_ < NB; ++s_) {
                const int S = p * NB + s_;
                const int r = rs[S];
                const T *pr = M + erow<TR>(r) * LD;
                cx<T> v = jmine < N8 ? mk<T>(pr[jmine], pr[TR * LD + jmine]) : zero;
#pragma unroll
                for (int tt = 0; tt < S; ++tt) {
                    cx<T> cf = Cn[(tt / NB) * N8 * NB + r * NB + (tt % NB)];       // -c^(tt)[r_s]
                    if (tt < p * NB && in0) cf = zero;                             // already applied to tile column ct0
                    cfma(v, cf, rho[tt]);
                }
                rho[S] = v;
                if (jmine < N8) Rho[S * L::RLD + jmine] = v;
            }
            pair_sync(pair);
            if (NP == 1) break;                                    // one inner panel: a single pass over all tile columns below
            // ---- 4a. rank-NB update of tile column ct0 (half of the row tiles per warp) ----
            iw_update_half<NT, 0, 1, 1>(M, Cn + p * N8 * NB, Rho + p * NB * L::RLD, lane, w, ct0, 99);
            pair_sync(pair);
            // ---- 5. the panel columns take the factored panel (warp 0; it is also the only reader of the next panel) ----
            if (w == 0) iw_store_panel<T, NT>(M, K, lane, pc);
        }
        // ---- 4b. rank-(NP NB) update of the other NT - 1 tile columns (NP = 1: of all NT), one pass over their accumulators ----
        if (NP > 1) {
            iw_update_half<NT, 0, NP, NT - 1>(M, Cn, Rho, lane, w, 0, ct0);
            pair_sync(pair);
        } else {
            iw_update_half<NT, 0, 1, NT>(M, Cn, Rho, lane, w, 0, 99);
            pair_sync(pair);
            if (w == 0) iw_store_panel<T, NT>(M, K0, lane, pc);
        }
    }
    // ---- un-permute on the way out: Y[kof[i]][rowof[j]] = a_ij (the warps alternate rows) ----
    {
        const int yj0 = rowof[lane], yj1 = rowof[lane + 32];
#pragma unroll 4
        for (int i = w; i < n; i += 2) {
            const int r = erow<TR>(i), yi = kof[i];
            if (yi < n) {
                if (lane < n && yj0 < n) Yb[yi * n + yj0] = mk<T>(M[r * LD + lane], M[(r + TR) * LD + lane]);
                if (lane + 32 < n && yj1 < n) Yb[yi * n + yj1] = mk<T>(M[r * LD + lane + 32], M[(r + TR) * LD + lane + 32]);
            }
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (w == 0 && lane == 0 && info) info[mat] = bad;
}

template <typename T, int NT>
static cudaError_t launch_iw(const void *A, int n, void *Y, int64_t batch, int *info, cudaStream_t s)
{
    // as many warp pairs (= matrices) per CTA as 227 KB of shared memory hold, one CTA per SM
    const int wpc = IwLayout<T, NT>::WPC;       // order 53: 8 x 28 944 B (FP32), 4 x 57 744 B (FP64)
    const size_t smem = wpc * IwLayout<T, NT>::BYTES;
    cudaError_t e = cudaFuncSetAttribute(cinverse_warp_kernel<T, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cinverse_warp_kernel<T, NT><<<(unsigned)((batch + wpc - 1) / wpc), 64 * wpc, smem, s>>>((const cx<T> *)A, n, (cx<T> *)Y, info, batch);
    return cudaGetLastError();
}

// orders 33 .. 64
cudaError_t launch_cinverse_tc(wifi_dtype dt, const void *A, int order, void *Y, int64_t batch, int *info, cudaStream_t s)
{
    const int nt = (order + 7) / 8;
#define IW(T)                                                                            \
    switch (nt) {                                                                        \
    case 5: return launch_iw<T, 5>(A, order, Y, batch, info, s);                         \
    case 6: return launch_iw<T, 6>(A, order, Y, batch, info, s);                         \
    case 7: return launch_iw<T, 7>(A, order, Y, batch, info, s);                         \
    default: return launch_iw<T, 8>(A, order, Y, batch, info, s);                        \
    }
    if (dt == WIFI_F32) { IW(float) }
    IW(double)
#undef IW
}

// ------------------------------------------------------------------------------------------------------------------------
// General per-frame PS_MMSE solve (any non-singular R + D: the WIFI_SOLVE_PIVOT path) on the same machinery, FP64 arithmetic:
//   A = R + diag(sigma2 / |tx_k|^2),  y = rx / tx,  A z = y  by LU with partial pivoting + back-substitution,  H = y - D z
// (Gauss-Jordan is not usable here: its forward error in z is not the image of a small backward error).
// A warp pair per frame; [A | y] lives in shared memory as the real planes above (56 x 56: y is column 54, rows / columns 53..55
// carry a unit diagonal); implicit pivoting -- no row swaps: a row that has been a pivot row gets zero multipliers from then on and
// simply stops changing, so U stays in place for the back-substitution.  Per block of two columns: warp 0 factors the panel in
// registers (multipliers l_i = a_ik / p of the rows still to be eliminated, zero for the others), both warps form the two pivot
// rows, and the rank-2 update of the column tiles right of the panel is one DMMA per 8 x 8 real tile, accumulators streamed
// through shared memory in the software pipeline of iw_update.  Back-substitution: warp 0, the right-hand side in registers
// (lane = rows lane, lane + 32), one shuffle broadcast per unknown.  Replaces the CTA-per-frame shared-memory LU for a shared R
// (that kernel stays for the rank-one calling convention of main.c:148 and the FP32-arithmetic opt-in).
struct PvL {
    static constexpr int NT = 7, N8 = 56, TR = 4, LD = 56, RLD = 60, ROWS = 112, YC = 54, WPC = 4;
    static constexpr size_t BYTES = sizeof(double) * ((size_t)ROWS * LD + 4 * N8 + 4 * RLD + 2 * N8 + 2 * N8) + 2 * 64 + 16;
};

template <int RT0, int NR>
__device__ __forceinline__ void pv_update(double *M, const double2 *Cn, const double2 *Rho, int lane, int c0, int nc)
{
    const int g = lane >> 2, t = lane & 3;
    const int apart = (g >> 2) ^ (t >> 1);
    const bool aneg = g < 4 && t >= 2;
    double af[NR];
#pragma unroll
    for (int q = 0; q < NR; ++q) {
        const double v = reinterpret_cast<const double *>(Cn + (4 * (RT0 + q) + (g & 3)) * 2 + (t & 1))[apart];
        af[q] = aneg ? -v : v;
    }
    double *const p0 = M + ((8 * RT0 + g) * PvL::LD + 2 * t);
    const double *const r0 = reinterpret_cast<const double *>(Rho + (t & 1) * PvL::RLD + g) + (t >> 1);
    auto issue = [&](int ct, double (&a)[NR][2]) {
        const double bf = r0[16 * ct];
        const double *p = p0 + 8 * ct;
#pragma unroll
        for (int q = 0; q < NR; ++q) {
            const double2 v = lds_acc(p + q * 8 * PvL::LD);
            a[q][0] = v.x; a[q][1] = v.y;
        }
#pragma unroll
        for (int q = 0; q < NR; ++q) mma_f64(a[q], af[q], bf);
    };
    auto store = [&](int ct, const double (&a)[NR][2]) {
        double *p = p0 + 8 * ct;
#pragma unroll
        for (int q = 0; q < NR; ++q) sts_acc(p + q * 8 * PvL::LD, a[q][0], a[q][1]);
    };
    double accA[NR][2], accB[NR][2];
    issue(c0, accA);
#pragma unroll 1
    for (int i = 0; i < nc; i += 2) {
        if (i + 1 < nc) issue(c0 + i + 1, accB);
        store(c0 + i, accA);
        if (i + 1 < nc) {
            if (i + 2 < nc) issue(c0 + i + 2, accA);
            store(c0 + i + 1, accB);
        }
    }
}

template <typename TIO>
__global__ void __launch_bounds__(64 * PvL::WPC) mmse_pivot_tc_kernel(const cx<TIO> *__restrict__ R, const cx<TIO> *__restrict__ tx,
                                                                     const cx<TIO> *__restrict__ rx, int64_t frame_stride,
                                                                     const TIO *__restrict__ sigma2, cx<TIO> *__restrict__ H, int *info,
                                                                     int64_t n_frames)
{
    constexpr int N8 = PvL::N8, TR = PvL::TR, LD = PvL::LD, YC = PvL::YC;
    extern __shared__ __align__(16) unsigned char iw_smem[];
    const int pair = threadIdx.x >> 6, w = (threadIdx.x >> 5) & 1, lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * PvL::WPC + pair;
    if (f >= n_frames) return;
    double *M = reinterpret_cast<double *>(iw_smem + pair * PvL::BYTES);
    double2 *Cn = reinterpret_cast<double2 *>(M + (size_t)PvL::ROWS * LD);      // [N8][2]: -l
    double2 *Rho = Cn + N8 * 2;                                                 // [2][RLD]: the two pivot rows of the block
    double2 *zv = Rho + 2 * PvL::RLD;                                           // [N8] solution
    double2 *pinv = zv + N8;                                                    // [N8] 1 / pivot of column k
    unsigned char *rowof = reinterpret_cast<unsigned char *>(pinv + N8), *kof = rowof + 64, *rs = kof + 64;
    const cx<TIO> *txf = tx + f * frame_stride, *rxf = rx + f * frame_stride;
    const double s2 = (double)sigma2[f];
    auto wd = [](cx<TIO> v) { return make_double2((double)v.x, (double)v.y); };
    const double2 zero = make_double2(0.0, 0.0), one = make_double2(1.0, 0.0);

    // ---- load [R + D | . | y | .]: lane = columns lane, lane + 32; the warps alternate groups of eight rows ----
    for (int i0 = 8 * w; i0 < N8; i0 += 16) {
        double2 v[8][2], tq[8], rq[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q;
            tq[q] = one; rq[q] = zero;
            if (i < NSC) { tq[q] = wd(txf[i]); rq[q] = wd(rxf[i]); }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = lane + 32 * h;
                v[q][h] = (i == j) ? one : zero;
                if (i < NSC && j < NSC) v[q][h] = wd(R[i * NSC + j]);
            }
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int i = i0 + q, r = erow<TR>(i);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int j = lane + 32 * h;
                double2 a = v[q][h];
                if (i < NSC && j == i) a.x += s2 / cabs2(tq[q]);
                if (i < NSC && j == YC) a = cdiv(rq[q], tq[q]);
                if (j < N8) { M[r * LD + j] = a.x; M[(r + TR) * LD + j] = a.y; }
            }
        }
    }
    if (w == 0) {
        rowof[lane] = (unsigned char)lane; rowof[lane + 32] = (unsigned char)(lane + 32);
        kof[lane] = (unsigned char)lane; kof[lane + 32] = (unsigned char)(lane + 32);
    }
    unsigned usedw = 0;
    if (lane >= N8) usedw |= 1u;
    if (lane + 32 >= N8) usedw |= 2u;
    int bad = 0;
    pair_sync(pair);

#pragma unroll 1
    for (int K = 0; K < YC; K += 2) {
        if (w == 0) {
            double2 pc[2][2], cn[2][2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int i = lane + 32 * h;
                pc[h][0] = pc[h][1] = zero;
                if (i < N8) {
                    const double *pr = M + erow<TR>(i) * LD + K;
                    double vr[2], vi[2];
                    ld_panel(pr, vr); ld_panel(pr + TR * LD, vi);
                    pc[h][0] = make_double2(vr[0], vi[0]); pc[h][1] = make_double2(vr[1], vi[1]);
                }
            }
#pragma unroll
            for (int s_ = 0; s_ < 2; ++s_) {
                unsigned key = 0;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned kh = 0x80000000u | (iw_bits(cabs2(pc[h][s_])) & 0x7fffffc0u) | (unsigned)(63 - (lane + 32 * h));
                    if (!((usedw >> h) & 1u) && kh > key) key = kh;
                }
                const int r = 63 - (int)(__reduce_max_sync(0xffffffffu, key) & 63u);
                const int hr = r >> 5, ol = r & 31;
                double2 prow[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const double2 v = hr ? pc[1][u] : pc[0][u];
                    prow[u].x = __shfl_sync(0xffffffffu, v.x, ol);
                    prow[u].y = __shfl_sync(0xffffffffu, v.y, ol);
                }
                const double2 piv = prow[s_];
                const double den = cabs2(piv), rden = iw_rcp(den);
                const double2 inv = make_double2(piv.x * rden, -piv.y * rden);
                bad |= !(den > 0.0);
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int i = lane + 32 * h;
                    const bool elim = !((usedw >> h) & 1u) && i != r;            // rows still to be eliminated
                    const double2 l = elim ? cmul(pc[h][s_], inv) : zero;
                    cn[h][s_] = make_double2(-l.x, -l.y);
                    if (s_ == 0) cfms(pc[h][1], l, prow[1]);
                    if (i == r) usedw |= 1u << h;
                }
                if (lane == 0) { rowof[K + s_] = (unsigned char)r; kof[r] = (unsigned char)(K + s_); rs[s_] = (unsigned char)r; pinv[K + s_] = inv; }
            }
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (lane + 32 * h < N8) st_crow(Cn + (lane + 32 * h) * 2, cn[h]);
        }
        pair_sync(pair);
        {   // the two pivot rows, warp w = columns lane + 32 w: rho0 = a[r0], rho1 = a[r1] - l0[r1] rho0
            const int j = lane + 32 * w;
            const int r0 = rs[0], r1 = rs[1];
            if (j < N8) {
                const double *p0 = M + erow<TR>(r0) * LD, *p1 = M + erow<TR>(r1) * LD;
                const double2 v0 = make_double2(p0[j], p0[TR * LD + j]);
                double2 v1 = make_double2(p1[j], p1[TR * LD + j]);
                cfma(v1, Cn[r1 * 2], v0);
                Rho[j] = v0; Rho[PvL::RLD + j] = v1;
            }
        }
        pair_sync(pair);
        {   // rank-2 update of the column tiles that still hold live columns (K + 2 ..), row tiles split 6 + 8
            const int c0 = (K + 2) >> 3, nc = PvL::NT - c0;
            if (w == 0) pv_update<0, PvL::NT - 1>(M, Cn, Rho, lane, c0, nc);
            else pv_update<PvL::NT - 1, PvL::NT + 1>(M, Cn, Rho, lane, c0, nc);
        }
        pair_sync(pair);
    }
    if (w != 0) return;
    // ---- back-substitution (warp 0): z_k = y'[r_k] / p_k;  y'[i] -= U[i][k] z_k for the pivot rows of earlier columns ----
    double2 yv[2];
    int kofv[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        yv[h] = zero; kofv[h] = -1;
        if (i < N8) { const double *pr = M + erow<TR>(i) * LD + YC; yv[h] = make_double2(pr[0], pr[TR * LD]); kofv[h] = kof[i]; }
    }
#pragma unroll 1
    for (int k = NSC - 1; k >= 0; --k) {
        const int r = rowof[k], hr = r >> 5, ol = r & 31;
        const double2 ys = hr ? yv[1] : yv[0];
        const double2 yr = make_double2(__shfl_sync(0xffffffffu, ys.x, ol), __shfl_sync(0xffffffffu, ys.y, ol));
        const double2 zk = cmul(yr, pinv[k]);
        if (lane == 0) zv[k] = zk;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            if (kofv[h] >= 0 && kofv[h] < k) {
                const double *pr = M + erow<TR>(i) * LD + k;
                cfms(yv[h], make_double2(pr[0], pr[TR * LD]), zk);
            }
        }
    }
    __syncwarp();
    // ---- H = y - D z; a bin whose noise term dominates the diagonal (d_k > R_kk: the DC bin) takes H_k = sum_j R_kj z_j ----
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int i = lane + 32 * h;
        bool viaR = false;
        if (i < NSC) {
            const double2 t = wd(txf[i]), rr = wd(rxf[i]);
            const double d = s2 / cabs2(t);
            viaR = d > (double)R[i * NSC + i].x;
            if (!viaR) {
                const double2 y = cdiv(rr, t), z = zv[i];
                H[f * NSC + i] = mk<TIO>((TIO)(y.x - d * z.x), (TIO)(y.y - d * z.y));
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, viaR);
        while (m) {
            const int i2 = (__ffs(m) - 1) + 32 * h;
            m &= m - 1;
            double2 acc = zero;
            for (int j = lane; j < NSC; j += 32) cfma(acc, wd(R[i2 * NSC + j]), zv[j]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) { acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o); acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o); }
            if (lane == 0) H[f * NSC + i2] = mk<TIO>((TIO)acc.x, (TIO)acc.y);
        }
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0 && info && bad) atomicExch(info, 1);
}

// shared R, FP64 arithmetic (complex64 or complex128 storage)
cudaError_t launch_mmse_pivot_tc(wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, void *H,
                                 int64_t n_frames, int *info, cudaStream_t s)
{
    const size_t smem = PvL::WPC * PvL::BYTES;
    const unsigned grid = (unsigned)((n_frames + PvL::WPC - 1) / PvL::WPC);
    cudaError_t e;
    if (dt == WIFI_F32) {
        if ((e = cudaFuncSetAttribute(mmse_pivot_tc_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        mmse_pivot_tc_kernel<float><<<grid, 64 * PvL::WPC, smem, s>>>((const float2 *)R, (const float2 *)tx, (const float2 *)rx, frame_stride,
                                                                      (const float *)sigma2, (float2 *)H, info, n_frames);
    } else {
        if ((e = cudaFuncSetAttribute(mmse_pivot_tc_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return e;
        mmse_pivot_tc_kernel<double><<<grid, 64 * PvL::WPC, smem, s>>>((const double2 *)R, (const double2 *)tx, (const double2 *)rx, frame_stride,
                                                                       (const double *)sigma2, (double2 *)H, info, n_frames);
    }
    return cudaGetLastError();
}

}  // namespace wifi
