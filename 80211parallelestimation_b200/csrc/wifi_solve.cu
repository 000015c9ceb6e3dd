// wifi_solve.cu -- dense complex solves that replace inverse() (utils.c:141-170, the O(n^5) un-pivoted
// cofactor expansion) on the device:
//   gj_solve            CTA-cooperative LU with partial pivoting + back-substitution on [A | B] held in shared
//                       memory; the pivot arg-max is a warp-shuffle reduction.
//   filter_form_kernel  W = R (R + diag d)^-1 in double-double, once per batch (main.c:183-201 intent).
//   (batched inverse, orders 33..64: wifi_inverse_tc.cu)
//   cinverse_kernel     batched inverse, orders <= 32: the shared-memory LU, one CTA per matrix.
//   mmse_pivot_kernel   per-frame MMSE, general (any non-singular R + D), one CTA per frame: the rank-one calling convention of main.c:148
//                       and the FP32-arithmetic opt-in; a shared R in FP64 arithmetic goes to mmse_pivot_tc_kernel (wifi_inverse_tc.cu).
// The register-resident un-pivoted fast path for Hermitian-PSD R lives in wifi_solve_hpd.cu.
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

// LU with partial pivoting (pivot rows normalised, so U has a unit diagonal) on the n x ncols augmented matrix
// `a` (row stride ld) in shared memory, followed by back-substitution of the ncols-n right-hand-side columns; on
// return columns n..ncols-1 hold A^-1 B.  Forward elimination + back-substitution rather than Gauss-Jordan: on the
// ill-conditioned R + D systems of the MMSE path the Jordan variant's forward error in z is not the image of a small
// backward error, and H = R z then loses ~cond(A) more digits (measured: 1e-7 vs 1e-11 in FP64).
// lcol: n elements of scratch, ctl: 2 ints of scratch.  Every thread of the CTA must call.
// Returns 1 (to all threads) if a pivot column was exactly zero.
// IDENT (the right-hand side is the identity, ncols = 2n; inverse()): the right half is kept in PIVOT ORDER and starts as zero --
// storage column n + t belongs to the original row index of the t-th pivot row, whose 1 enters when that row is chosen --
// so that after k steps only columns n .. n+k can be non-zero and the elimination skips the rest (a quarter of the flops of
// eliminating [A | I] densely).  orig[] (n ints, initialised to 0..n-1 by the caller) returns that column order.
template <typename T, bool IDENT = false>
__device__ int gj_solve(cx<T> *a, int n, int ld, int ncols, cx<T> *lcol, int *ctl, int *orig = nullptr)
{
    const int tid = threadIdx.x, nt = blockDim.x;
    const int tx = tid & 31, ty = tid >> 5, ny = nt >> 5;      // (every caller launches a multiple of 32 threads)
    int singular = 0;
    for (int k = 0; k < n; ++k) {
        // 1. pivot search down column k (warp 0, shuffle arg-max on |a|^2)
        if (tid < 32) {
            T best = -1;
            int bi = k;
            for (int i = k + tid; i < n; i += 32) {
                T v = cabs2(a[i * ld + k]);
                if (v > best) { best = v; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                T ov = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { ctl[0] = bi; ctl[1] = (best > (T)0) ? 0 : 1; }
        }
        __syncthreads();
        const int p = ctl[0];
        if (ctl[1]) { singular = 1; }
        // 2. swap rows k <-> p and scale the pivot row (columns >= k)
        const cx<T> inv = crecip(a[p * ld + k]);
        const int jend = IDENT ? n + k + 1 : ncols;          // columns that can be non-zero after this step
        __syncthreads();   // everyone has read the pivot before it is overwritten
        if (IDENT && tid == 0) { const int t = orig[k]; orig[k] = orig[p]; orig[p] = t; }
        for (int j = k + tid; j < jend; j += nt) {
            cx<T> top = a[k * ld + j], piv = a[p * ld + j];
            if (IDENT && j == n + k) piv = mk<T>((T)1, (T)0);      // the pivot row's own identity entry enters here
            a[k * ld + j] = cmul(piv, inv);
            if (p != k) a[p * ld + j] = top;
        }
        __syncthreads();
        // 3. multipliers of the rows below
        for (int i = k + 1 + tid; i < n; i += nt) lcol[i] = a[i * ld + k];
        __syncthreads();
        // 4. eliminate column k from the rows below.  A warp walks along a row (lane = column: conflict-free with the odd row
        // stride), warps take rows round-robin; the multiplier is a broadcast load and the pivot-row value stays in a register
        // for all the rows of a column (the flat e / w, e % w indexing this replaces spent more on integer division than on FMAs).
        for (int j = k + 1 + tx; j < jend; j += 32) {
            const cx<T> u = a[k * ld + j];
            int i = k + 1 + ty;
            for (; i + 3 * ny < n; i += 4 * ny) {          // four independent rows in flight
                cx<T> v0 = a[i * ld + j], v1 = a[(i + ny) * ld + j], v2 = a[(i + 2 * ny) * ld + j], v3 = a[(i + 3 * ny) * ld + j];
                cfms(v0, lcol[i], u); cfms(v1, lcol[i + ny], u); cfms(v2, lcol[i + 2 * ny], u); cfms(v3, lcol[i + 3 * ny], u);
                a[i * ld + j] = v0; a[(i + ny) * ld + j] = v1; a[(i + 2 * ny) * ld + j] = v2; a[(i + 3 * ny) * ld + j] = v3;
            }
            for (; i < n; i += ny) {
                cx<T> v = a[i * ld + j];
                cfms(v, lcol[i], u);
                a[i * ld + j] = v;
            }
        }
        __syncthreads();
    }
    // back-substitution (unit upper-triangular U): rows above k lose u_ik * x_k
    for (int k = n - 1; k >= 1; --k) {
        for (int j = n + tx; j < ncols; j += 32) {
            const cx<T> x = a[k * ld + j];
            int i = ty;
            for (; i + 3 * ny < k; i += 4 * ny) {
                cx<T> v0 = a[i * ld + j], v1 = a[(i + ny) * ld + j], v2 = a[(i + 2 * ny) * ld + j], v3 = a[(i + 3 * ny) * ld + j];
                cfms(v0, a[i * ld + k], x); cfms(v1, a[(i + ny) * ld + k], x); cfms(v2, a[(i + 2 * ny) * ld + k], x); cfms(v3, a[(i + 3 * ny) * ld + k], x);
                a[i * ld + j] = v0; a[(i + ny) * ld + j] = v1; a[(i + 2 * ny) * ld + j] = v2; a[(i + 3 * ny) * ld + j] = v3;
            }
            for (; i < k; i += ny) {
                cx<T> v = a[i * ld + j];
                cfms(v, a[i * ld + k], x);
                a[i * ld + j] = v;
            }
        }
        __syncthreads();
    }
    return singular;
}

// ------------------------------------------------------------------------------------------
// W = R (R + diag d)^-1, formed ONCE per batch in double-double arithmetic.
//
// cond(R + D) reaches 1e7..1e8 at the SNRs of the inputs.h frame, so a plain FP64 solve leaves ~1e-10 relative error
// in H = W y -- the whole FP64 parity budget.  The filter is a one-off 53 x 106 elimination, so it is done in
// double-double (two FP64 words per real, ~1e-32 unit roundoff, error-free transformations built on FMA) and rounded to
// FP64 at the end; after that W is exact to FP64 rounding and the per-frame error is the GEMM's alone.
// With S = diag(d^-1/2): R + D = S^-1 (B + I) S^-1, B = S R S, W = S^-1 [B (B+I)^-1] S; B + I is equilibrated.
// ------------------------------------------------------------------------------------------
struct dd { double hi, lo; };
__device__ __forceinline__ dd dd_make(double a) { dd r; r.hi = a; r.lo = 0.0; return r; }
// __dadd_rn / __dmul_rn are never contracted into FMAs, which the error-free transformations rely on
__device__ __forceinline__ dd dd_fix(double s, double e) { dd r; r.hi = __dadd_rn(s, e); r.lo = __dadd_rn(e, -__dadd_rn(r.hi, -s)); return r; }
__device__ __forceinline__ dd dd_add(dd a, dd b)
{
    double s = __dadd_rn(a.hi, b.hi), bb = __dadd_rn(s, -a.hi);
    double e = __dadd_rn(__dadd_rn(a.hi, -__dadd_rn(s, -bb)), __dadd_rn(b.hi, -bb));
    e = __dadd_rn(e, __dadd_rn(a.lo, b.lo));
    return dd_fix(s, e);
}
__device__ __forceinline__ dd dd_neg(dd a) { a.hi = -a.hi; a.lo = -a.lo; return a; }
__device__ __forceinline__ dd dd_mul(dd a, dd b)
{
    double p = __dmul_rn(a.hi, b.hi), e = __fma_rn(a.hi, b.hi, -p);
    e = __fma_rn(a.hi, b.lo, e);
    e = __fma_rn(a.lo, b.hi, e);
    return dd_fix(p, e);
}
__device__ __forceinline__ dd dd_div(dd a, dd b)
{
    double q1 = a.hi / b.hi;
    dd r = dd_add(a, dd_neg(dd_mul(b, dd_make(q1))));
    double q2 = r.hi / b.hi;
    r = dd_add(r, dd_neg(dd_mul(b, dd_make(q2))));
    double q3 = r.hi / b.hi;
    return dd_add(dd_fix(q1, q2), dd_make(q3));
}
struct cdd { dd x, y; };
__device__ __forceinline__ cdd cdd_mul(cdd a, cdd b)
{
    cdd r;
    r.x = dd_add(dd_mul(a.x, b.x), dd_neg(dd_mul(a.y, b.y)));
    r.y = dd_add(dd_mul(a.x, b.y), dd_mul(a.y, b.x));
    return r;
}
__device__ __forceinline__ cdd cdd_sub(cdd a, cdd b) { cdd r; r.x = dd_add(a.x, dd_neg(b.x)); r.y = dd_add(a.y, dd_neg(b.y)); return r; }
__device__ __forceinline__ cdd cdd_recip(cdd a)
{
    dd den = dd_add(dd_mul(a.x, a.x), dd_mul(a.y, a.y));
    cdd r; r.x = dd_div(a.x, den); r.y = dd_neg(dd_div(a.y, den)); return r;
}

constexpr int FF_THREADS = 1024;
constexpr int FF_LD = 2 * NSC + 1;

__global__ void __launch_bounds__(FF_THREADS) filter_form_kernel(const double2 *__restrict__ R, const double *__restrict__ d,
                                                                 double2 *__restrict__ W, int *info)
{
    extern __shared__ __align__(16) unsigned char ff_raw[];
    cdd *a = reinterpret_cast<cdd *>(ff_raw);        // [53][107]
    cdd *lcol = a + NSC * FF_LD;                      // [53]
    __shared__ int ctl[2];
    __shared__ double sc[NSC];                        // s_i = d_i^-1/2, or 1 if some d_i <= 0 (no equilibration then)
    __shared__ int all_pos;
    const int tid = threadIdx.x;
    if (tid == 0) {
        int ok = 1;
        for (int i = 0; i < NSC; ++i) ok &= (d[i] > 0.0);
        all_pos = ok;
    }
    __syncthreads();
    for (int i = tid; i < NSC; i += FF_THREADS) sc[i] = all_pos ? rsqrt(d[i]) : 1.0;
    __syncthreads();
    for (int e = tid; e < NSC * NSC; e += FF_THREADS) {
        int i = e / NSC, j = e - i * NSC;
        double2 r = R[e];
        dd sij = dd_mul(dd_make(sc[i]), dd_make(sc[j]));
        cdd b; b.x = dd_mul(dd_make(r.x), sij); b.y = dd_mul(dd_make(r.y), sij);
        cdd at = b;
        if (i == j) at.x = dd_add(at.x, all_pos ? dd_mul(dd_mul(dd_make(sc[i]), dd_make(sc[i])), dd_make(d[i])) : dd_make(d[i]));
        a[j * FF_LD + i] = at;             // (B + S D S)^T ; S D S = I up to the rounding of s_i, kept consistent here
        a[j * FF_LD + NSC + i] = b;        // B^T
    }
    __syncthreads();
    // Gauss-Jordan with partial pivoting on the 53 x 106 augmented matrix (X^T = (B+I)^-T B^T)
    int singular = 0;
    for (int k = 0; k < NSC; ++k) {
        if (tid < 32) {
            double best = -1.0;
            int bi = k;
            for (int i = k + tid; i < NSC; i += 32) {
                cdd v = a[i * FF_LD + k];
                double m = v.x.hi * v.x.hi + v.y.hi * v.y.hi;
                if (m > best) { best = m; bi = i; }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, o);
                int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (tid == 0) { ctl[0] = bi; ctl[1] = (best > 0.0) ? 0 : 1; }
        }
        __syncthreads();
        const int p = ctl[0];
        if (ctl[1]) singular = 1;
        const cdd inv = cdd_recip(a[p * FF_LD + k]);
        __syncthreads();
        for (int j = k + tid; j < 2 * NSC; j += FF_THREADS) {
            cdd top = a[k * FF_LD + j], piv = a[p * FF_LD + j];
            a[k * FF_LD + j] = cdd_mul(piv, inv);
            if (p != k) a[p * FF_LD + j] = top;
        }
        __syncthreads();
        for (int i = tid; i < NSC; i += FF_THREADS) lcol[i] = a[i * FF_LD + k];
        __syncthreads();
        const int w = 2 * NSC - (k + 1);
        for (int e = tid; e < NSC * w; e += FF_THREADS) {
            int i = e / w, j = k + 1 + (e - i * w);
            if (i != k) a[i * FF_LD + j] = cdd_sub(a[i * FF_LD + j], cdd_mul(lcol[i], a[k * FF_LD + j]));
        }
        __syncthreads();
    }
    for (int e = tid; e < NSC * NSC; e += FF_THREADS) {
        int i = e / NSC, j = e - i * NSC;
        cdd x = a[j * FF_LD + NSC + i];               // X[i][j] = (X^T)[j][i]
        dd f = dd_div(dd_make(sc[j]), dd_make(sc[i]));
        dd wr = dd_mul(x.x, f), wi = dd_mul(x.y, f);
        W[e] = make_double2(wr.hi + wr.lo, wi.hi + wi.lo);
    }
    if (tid == 0 && info) *info = singular;
}

cudaError_t launch_filter_form(const void *R64, const double *d64, void *W64, int *info, cudaStream_t s)
{
    g_last_launches = 1;
    size_t smem = sizeof(cdd) * (NSC * FF_LD + NSC);
    cudaError_t e = cudaFuncSetAttribute(filter_form_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    filter_form_kernel<<<1, FF_THREADS, smem, s>>>((const double2 *)R64, d64, (double2 *)W64, info);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// batched inverse
// ------------------------------------------------------------------------------------------
constexpr int INV_THREADS = 256;

template <typename T>
__global__ void __launch_bounds__(INV_THREADS) cinverse_kernel(const cx<T> *__restrict__ A, int n, cx<T> *__restrict__ Y, int *info)
{
    extern __shared__ __align__(16) unsigned char inv_smem[];
    cx<T> *a = (cx<T> *)inv_smem;
    const int ld = 2 * n + 1;
    cx<T> *lcol = a + n * ld;
    __shared__ int ctl[2];
    const cx<T> *Ab = A + (int64_t)blockIdx.x * n * n;
    cx<T> *Yb = Y + (int64_t)blockIdx.x * n * n;
    __shared__ int orig[WIFI_MAX_ORDER];
    for (int e = threadIdx.x; e < n * n; e += INV_THREADS) {
        int i = e / n, j = e - i * n;
        a[i * ld + j] = Ab[e];
        a[i * ld + n + j] = mk<T>(0, 0);          // the identity enters pivot by pivot (gj_solve<T, true>)
    }
    if (threadIdx.x < n) orig[threadIdx.x] = threadIdx.x;
    __syncthreads();
    int sing = gj_solve<T, true>(a, n, ld, 2 * n, lcol, ctl, orig);
    for (int e = threadIdx.x; e < n * n; e += INV_THREADS) {
        int i = e / n, t = e - i * n;
        Yb[i * n + orig[t]] = a[i * ld + n + t];
    }
    if (threadIdx.x == 0 && info) info[blockIdx.x] = sing;
}

cudaError_t launch_cinverse(wifi_dtype dt, const void *A, int order, void *Y, int64_t batch, int *info, cudaStream_t s)
{
    g_last_launches = 0;
    if (batch == 0 || order == 0) return cudaSuccess;
    g_last_launches = 1;
    const int ld = 2 * order + 1;
    cudaError_t e;
    // orders 33..64: warp pair per matrix, matrix in shared memory, trailing updates on the tensor cores (wifi_inverse_tc.cu)
    if (order > 32) return launch_cinverse_tc(dt, A, order, Y, batch, info, s);
    if (dt == WIFI_F32) {
        size_t smem = sizeof(float2) * ((size_t)order * ld + order);
        e = cudaFuncSetAttribute(cinverse_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cinverse_kernel<float><<<(unsigned)batch, INV_THREADS, smem, s>>>((const float2 *)A, order, (float2 *)Y, info);
    } else {
        size_t smem = sizeof(double2) * ((size_t)order * ld + order);
        e = cudaFuncSetAttribute(cinverse_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cinverse_kernel<double><<<(unsigned)batch, INV_THREADS, smem, s>>>((const double2 *)A, order, (double2 *)Y, info);
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// per-frame MMSE, general pivoted path: one CTA per frame
//   A = R + diag(sigma2/|tx_k|^2),  y = rx/tx,  A z = y,  H = R z
// R is shared (R != NULL) or the per-frame rank-1 H_ls H_ls^H of the C calling convention (main.c:186-189)
// ------------------------------------------------------------------------------------------
constexpr int PV_THREADS = 128;
constexpr int PV_LD = NSC + 2;

// TIO = storage type of R/tx/rx/sigma2/hls/H, T = arithmetic type.  <float, double> is the WIFI_F32 default: sigma2/|tx|^2
// (1e-10 .. 1e-7) is below the FP32 resolution of R, so the FP32-arithmetic elimination (<float, float>, WIFI_SOLVE_FAST32)
// is only accurate to ~1e-1 at the 60 dB end.
template <typename TIO, typename T>
__global__ void __launch_bounds__(PV_THREADS) mmse_pivot_kernel(const cx<TIO> *__restrict__ R, const cx<TIO> *__restrict__ tx,
                                                                const cx<TIO> *__restrict__ rx, int64_t frame_stride,
                                                                const TIO *__restrict__ sigma2, const cx<TIO> *__restrict__ hls,
                                                                cx<TIO> *__restrict__ H, int *info)
{
    extern __shared__ __align__(16) unsigned char pv_smem[];
    cx<T> *a = (cx<T> *)pv_smem;          // [53][55]: A | y
    cx<T> *lcol = a + NSC * PV_LD;        // [53]
    cx<T> *hv = lcol + NSC;               // [53] H_ls of this frame (rank-1 mode)
    __shared__ int ctl[2];
    const int64_t f = blockIdx.x;
    const cx<TIO> *txf = tx + f * frame_stride, *rxf = rx + f * frame_stride;
    auto wd = [](cx<TIO> v) { return mk<T>((T)v.x, (T)v.y); };
    const T s2 = (T)sigma2[f];
    if (hls) {
        for (int i = threadIdx.x; i < NSC; i += PV_THREADS) hv[i] = wd(hls[f * NSC + i]);
        __syncthreads();
    }
    for (int e = threadIdx.x; e < NSC * NSC; e += PV_THREADS) {
        int i = e / NSC, j = e - i * NSC;
        cx<T> r = hls ? cmul(hv[i], cconj(hv[j])) : wd(R[e]);
        if (i == j) r.x += s2 / cabs2(wd(txf[i]));
        a[i * PV_LD + j] = r;
    }
    for (int i = threadIdx.x; i < NSC; i += PV_THREADS) a[i * PV_LD + NSC] = cdiv(wd(rxf[i]), wd(txf[i]));
    __syncthreads();
    int sing = gj_solve<T>(a, NSC, PV_LD, NSC + 1, lcol, ctl);
    // z is column 53; H = R z
    if (hls) {
        // R z = h (h^H z)
        if (threadIdx.x < 32) {
            cx<T> acc = mk<T>(0, 0);
            for (int j = threadIdx.x; j < NSC; j += 32) cfma(acc, cconj(hv[j]), a[j * PV_LD + NSC]);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o);
                acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
            }
            if (threadIdx.x == 0) lcol[0] = acc;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < NSC; i += PV_THREADS) { const cx<T> h = cmul(hv[i], lcol[0]); H[f * NSC + i] = mk<TIO>((TIO)h.x, (TIO)h.y); }
    } else {
        for (int i = threadIdx.x; i < NSC; i += PV_THREADS) {
            cx<T> acc = mk<T>(0, 0);
            for (int j = 0; j < NSC; ++j) cfma(acc, wd(R[i * NSC + j]), a[j * PV_LD + NSC]);
            H[f * NSC + i] = mk<TIO>((TIO)acc.x, (TIO)acc.y);
        }
    }
    if (threadIdx.x == 0 && info && sing) atomicExch(info, 1);
}

template <typename TIO, typename T>
static cudaError_t launch_pivot(const void *R, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, const void *hls,
                                void *H, int64_t n_frames, int *info, cudaStream_t s)
{
    const size_t smem = sizeof(cx<T>) * (NSC * PV_LD + 2 * NSC);
    cudaError_t e = cudaFuncSetAttribute(mmse_pivot_kernel<TIO, T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mmse_pivot_kernel<TIO, T><<<(unsigned)n_frames, PV_THREADS, smem, s>>>((const cx<TIO> *)R, (const cx<TIO> *)tx, (const cx<TIO> *)rx, frame_stride,
                                                                          (const TIO *)sigma2, (const cx<TIO> *)hls, (cx<TIO> *)H, info);
    return cudaGetLastError();
}

cudaError_t launch_mmse_perframe_pivot(wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                                       const void *sigma2, const void *Hls_for_R, void *H, int64_t n_frames, int *info, int fast32,
                                       cudaStream_t s)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    if (dt == WIFI_F32 && fast32) return launch_pivot<float, float>(R, tx, rx, frame_stride, sigma2, Hls_for_R, H, n_frames, info, s);
    // shared R in FP64 arithmetic: the warp-pair LU with tensor-core updates (wifi_inverse_tc.cu)
    if (!Hls_for_R) return launch_mmse_pivot_tc(dt, R, tx, rx, frame_stride, sigma2, H, n_frames, info, s);
    if (dt == WIFI_F32) return launch_pivot<float, double>(R, tx, rx, frame_stride, sigma2, Hls_for_R, H, n_frames, info, s);
    return launch_pivot<double, double>(R, tx, rx, frame_stride, sigma2, Hls_for_R, H, n_frames, info, s);
}

}  // namespace wifi
