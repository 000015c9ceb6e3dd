// wifi_gemm.cu -- the small batched complex matrix utils of utils.h:38-60 (multiply, hermitian, addition, multiplyVxVeqM,
// identity).  The shared-filter PS_MMSE GEMM over all frames lives in wifi_gemm_tc.cu (FP32: 3xTF32 on tcgen05) and
// wifi_gemm_dmma.cu (FP64: DMMA); the CUDA-core version they replaced (0.998 ms per 1 Mi frames against 0.23) is gone.
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

// W' = W diag(1/tx): the LS divide by a shared, known tx block vector folded into the 53 x 53 filter (FP64, once per batch)
__global__ void filter_fold_kernel(const double2 *__restrict__ W, const double2 *__restrict__ tx, double2 *__restrict__ Wout)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < NSC * NSC) Wout[e] = cmul(W[e], crecip(tx[e % NSC]));
}

cudaError_t launch_filter_fold(const void *W64, const void *tx64, void *Wout64, cudaStream_t s)
{
    g_last_launches = 1;
    filter_fold_kernel<<<(NSC * NSC + 255) / 256, 256, 0, s>>>((const double2 *)W64, (const double2 *)tx64, (double2 *)Wout64);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// batched small-matrix utils (utils.h:38-60)
// ------------------------------------------------------------------------------------------
// one complex element global -> shared without a register round trip: the staging loops are pure latency otherwise (a dozen
// dependent load -> store iterations per thread: ~19 k cycles per 53 x 53 pair, more than the product itself)
template <typename C> __device__ __forceinline__ void stage_async(C *dst, const C *src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "n"((int)sizeof(C)) : "memory");
}
__device__ __forceinline__ void stage_wait()
{
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

// multiply utils.c:16-31: one CTA per matrix pair, k ascending like the reference's inner loop.  Both operands are staged in
// shared memory; a thread owns a 4 x 4 set of outputs -- rows tr + Sr i, columns tc + Sc j (Sr = ceil(r1 / 4), Sc = ceil(c2 / 4):
// strided, so that the B loads of neighbouring threads are neighbouring addresses and the A loads are broadcasts) -- and does
// 16 complex FMAs per 8 shared loads.  (One output per thread, 2 loads per complex FMA: 17 / 8 TFLOP/s in FP32 / FP64 at
// order 53.)
template <typename T>
__global__ void __launch_bounds__(256) cmatmul_kernel(const cx<T> *__restrict__ A, int r1, int c1, const cx<T> *__restrict__ B, int c2, cx<T> *__restrict__ C)
{
    extern __shared__ __align__(16) unsigned char mm_smem[];
    cx<T> *sa = (cx<T> *)mm_smem, *sb = sa + r1 * c1;
    const cx<T> *Ab = A + (int64_t)blockIdx.x * r1 * c1, *Bb = B + (int64_t)blockIdx.x * c1 * c2;
    cx<T> *Cb = C + (int64_t)blockIdx.x * r1 * c2;
    for (int e = threadIdx.x; e < r1 * c1; e += blockDim.x) stage_async(sa + e, Ab + e);
    for (int e = threadIdx.x; e < c1 * c2; e += blockDim.x) stage_async(sb + e, Bb + e);
    stage_wait();
    __syncthreads();
    const int Sr = (r1 + 3) >> 2, Sc = (c2 + 3) >> 2;
    if ((int)threadIdx.x >= Sr * Sc) return;
    const int tr = threadIdx.x / Sc, tc = threadIdx.x - tr * Sc;
    // rows / columns past the edge are clamped for the loads and dropped at the store
    const cx<T> *pa[4];
    int cb[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = tr + Sr * i, c = tc + Sc * i;
        pa[i] = sa + (r < r1 ? r : r1 - 1) * c1;
        cb[i] = c < c2 ? c : c2 - 1;
    }
    cx<T> acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = mk<T>(0, 0);
#pragma unroll 2
    for (int k = 0; k < c1; ++k) {
        cx<T> a[4], b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = pa[i][k];
#pragma unroll
        for (int j = 0; j < 4; ++j) b[j] = sb[k * c2 + cb[j]];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) cfma(acc[i][j], a[i], b[j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = tr + Sr * i;
        if (r >= r1) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = tc + Sc * j;
            if (c < c2) Cb[r * c2 + c] = acc[i][j];
        }
    }
}

// FP64 multiply on the FP64 tensor path.  C = A B as the real product C~ (2 r1 x c2) = A~ (2 r1 x 2 c1) B~ (2 c1 x c2) in the
// half-embedded form (rows 2i, 2i+1 of C~ = Re, Im of row i; A~ rows = [re, -im | im, re] per complex entry, B~ rows = Re, Im
// of row p): the flop count of the complex product, no redundancy.  Both operands are staged in shared memory as complex
// values, zero-padded to the tile grid (4 complex rows x 8 complex columns per 8 x 8 real tile, 2 complex k per DMMA); a warp
// takes whole tile rows, builds its A fragment from one 16-byte load (select + sign by lane) and reads the B fragments as
// plain 8-byte loads of the complex array.  k runs ascending like the reference's inner loop.
__device__ __forceinline__ void mm_dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(256) cmatmul_dmma_kernel(const double2 *__restrict__ A, int r1, int c1, const double2 *__restrict__ B, int c2,
                                                           double2 *__restrict__ C)
{
    extern __shared__ __align__(16) unsigned char mm_smem[];
    const int R1P = (r1 + 3) & ~3, C1P = (c1 + 1) & ~1, C2P = (c2 + 7) & ~7;
    double2 *sa = (double2 *)mm_smem, *sb = sa + R1P * C1P;
    const double2 *Ab = A + (int64_t)blockIdx.x * r1 * c1, *Bb = B + (int64_t)blockIdx.x * c1 * c2;
    double2 *Cb = C + (int64_t)blockIdx.x * r1 * c2;
    for (int e = threadIdx.x; e < R1P * C1P; e += blockDim.x) {
        const int i = e / C1P, p = e - i * C1P;
        if (i < r1 && p < c1) stage_async(sa + e, Ab + i * c1 + p);
        else sa[e] = make_double2(0.0, 0.0);
    }
    for (int e = threadIdx.x; e < C1P * C2P; e += blockDim.x) {
        const int p = e / C2P, j = e - p * C2P;
        if (p < c1 && j < c2) stage_async(sb + e, Bb + p * c2 + j);
        else sb[e] = make_double2(0.0, 0.0);
    }
    stage_wait();
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int r = lane >> 2, q = lane & 3, part = r & 1, kk = q & 1;
    const int ntc = C2P >> 3, nks = C1P >> 1;
    const double *sbd = (const double *)sb + (size_t)((q >> 1) * C2P + r) * 2 + kk;       // B~[k = q][col = r] of tile column 0, k-step 0
    for (int I = w; 4 * I < R1P; I += nw) {
        const int i = 4 * I + (r >> 1);
        const double2 *arow = sa + i * C1P + (q >> 1);
        double acc[8][2];
#pragma unroll
        for (int J = 0; J < 8; ++J) acc[J][0] = acc[J][1] = 0.0;
        for (int ks = 0; ks < nks; ++ks) {
            const double2 a = arow[2 * ks];
            const double av = part == kk ? a.x : (part ? a.y : -a.y);
            const double *bp = sbd + (size_t)ks * 4 * C2P;
#pragma unroll
            for (int J = 0; J < 8; ++J)
                if (J < ntc) mm_dmma(acc[J][0], acc[J][1], av, bp[16 * J]);
        }
        // a lane holds Re (part 0) or Im (part 1) of columns 8J + 2q, +1 of row i: pair up with the lane that holds the other part
#pragma unroll
        for (int J = 0; J < 8; ++J) {
            if (J < ntc) {
                const double send = part ? acc[J][0] : acc[J][1];
                const double recv = __shfl_xor_sync(0xffffffffu, send, 4);
                const double2 out = part ? make_double2(recv, acc[J][1]) : make_double2(acc[J][0], recv);
                const int col = 8 * J + 2 * q + part;
                if (i < r1 && col < c2) Cb[i * c2 + col] = out;
            }
        }
    }
}

cudaError_t launch_cmatmul(wifi_dtype dt, const void *A, int r1, int c1, const void *B, int c2, void *C, int64_t batch, cudaStream_t s)
{
    g_last_launches = 0;
    if (batch == 0 || r1 * c2 == 0) return cudaSuccess;
    g_last_launches = 1;
    cudaError_t e;
    if (dt == WIFI_F32) {
        size_t smem = sizeof(float2) * ((size_t)r1 * c1 + (size_t)c1 * c2);
        e = cudaFuncSetAttribute(cmatmul_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cmatmul_kernel<float><<<(unsigned)batch, 256, smem, s>>>((const float2 *)A, r1, c1, (const float2 *)B, c2, (float2 *)C);
    } else if ((int64_t)r1 * c1 * c2 >= 4096 && c1 > 0) {
        const int R1P = (r1 + 3) & ~3, C1P = (c1 + 1) & ~1, C2P = (c2 + 7) & ~7;
        size_t smem = sizeof(double2) * ((size_t)R1P * C1P + (size_t)C1P * C2P);
        e = cudaFuncSetAttribute(cmatmul_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        // whole tile rows per warp: as many warps as keep the rounds full (14 tile rows at order 53 -> 7 warps x 2 rounds)
        const int ntr = R1P / 4, rounds = (ntr + 7) / 8, warps = (ntr + rounds - 1) / rounds;
        cmatmul_dmma_kernel<<<(unsigned)batch, 32 * warps, smem, s>>>((const double2 *)A, r1, c1, (const double2 *)B, c2, (double2 *)C);
    } else {
        size_t smem = sizeof(double2) * ((size_t)r1 * c1 + (size_t)c1 * c2);
        e = cudaFuncSetAttribute(cmatmul_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        cmatmul_kernel<double><<<(unsigned)batch, 256, smem, s>>>((const double2 *)A, r1, c1, (const double2 *)B, c2, (double2 *)C);
    }
    return cudaGetLastError();
}

// hermitian utils.c:3-7.  mode AS_WRITTEN: res[c][r] = Re(M[r][c]) - Im(M[r][c]) (real-valued, sic);
// mode INTENDED: conjugate transpose.  Tiled through shared memory so both sides are coalesced.
template <typename T>
__global__ void chermitian_kernel(int mode, const cx<T> *__restrict__ M, int row, int col, cx<T> *__restrict__ res)
{
    __shared__ cx<T> tile[32][33];
    const cx<T> *Mb = M + (int64_t)blockIdx.z * row * col;
    cx<T> *Rb = res + (int64_t)blockIdx.z * row * col;
    int c = blockIdx.x * 32 + threadIdx.x, r = blockIdx.y * 32 + threadIdx.y;
    if (r < row && c < col) tile[threadIdx.y][threadIdx.x] = Mb[r * col + c];
    __syncthreads();
    int rr = blockIdx.y * 32 + threadIdx.x, cc = blockIdx.x * 32 + threadIdx.y;   // res[cc][rr]
    if (rr < row && cc < col) {
        cx<T> m = tile[threadIdx.x][threadIdx.y];
        Rb[cc * row + rr] = mode == WIFI_AS_WRITTEN ? mk<T>(m.x - m.y, (T)0) : mk<T>(m.x, -m.y);
    }
}

cudaError_t launch_chermitian(wifi_dtype dt, int mode, const void *M, int row, int col, void *res, int64_t batch, cudaStream_t s)
{
    g_last_launches = 0;
    if (batch == 0 || row * col == 0) return cudaSuccess;
    g_last_launches = 1;
    dim3 grid((col + 31) / 32, (row + 31) / 32, (unsigned)batch), blk(32, 32);
    if (dt == WIFI_F32) chermitian_kernel<float><<<grid, blk, 0, s>>>(mode, (const float2 *)M, row, col, (float2 *)res);
    else chermitian_kernel<double><<<grid, blk, 0, s>>>(mode, (const double2 *)M, row, col, (double2 *)res);
    return cudaGetLastError();
}

// addition utils.c:111-121.  AS_WRITTEN: res = M1 + M1 (M2 ignored, sic); INTENDED: M1 + M2
template <typename T>
__global__ void cadd_kernel(int mode, const cx<T> *__restrict__ M1, const cx<T> *__restrict__ M2, cx<T> *__restrict__ res, int64_t n)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) { cx<T> a = M1[e]; res[e] = cadd(a, mode == WIFI_AS_WRITTEN ? a : M2[e]); }
}

cudaError_t launch_cadd(wifi_dtype dt, int mode, const void *M1, const void *M2, void *res, int64_t n, cudaStream_t s)
{
    g_last_launches = 0;
    if (n == 0) return cudaSuccess;
    g_last_launches = 1;
    unsigned grid = (unsigned)((n + 255) / 256);
    if (dt == WIFI_F32) cadd_kernel<float><<<grid, 256, 0, s>>>(mode, (const float2 *)M1, (const float2 *)M2, (float2 *)res, n);
    else cadd_kernel<double><<<grid, 256, 0, s>>>(mode, (const double2 *)M1, (const double2 *)M2, (double2 *)res, n);
    return cudaGetLastError();
}

// multiplyVxVeqM utils.c:55-65: res[r][c] = M1[r][0] * M2[0][c]
template <typename T>
__global__ void couter_kernel(const cx<T> *__restrict__ M1, int r1, int c1, const cx<T> *__restrict__ M2, int c2, int r2,
                              cx<T> *__restrict__ res)
{
    const cx<T> *a = M1 + (int64_t)blockIdx.y * r1 * c1, *b = M2 + (int64_t)blockIdx.y * r2 * c2;
    cx<T> *o = res + (int64_t)blockIdx.y * r1 * c2;
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < r1 * c2) { int r = e / c2, c = e - r * c2; o[e] = cmul(a[r * c1], b[c]); }
}

cudaError_t launch_couter(wifi_dtype dt, const void *M1, int r1, int c1, const void *M2, int c2, void *res, int64_t batch, cudaStream_t s)
{
    g_last_launches = 0;
    if (batch == 0 || r1 * c2 == 0) return cudaSuccess;
    g_last_launches = 1;
    dim3 grid((r1 * c2 + 255) / 256, (unsigned)batch);
    if (dt == WIFI_F32) couter_kernel<float><<<grid, 256, 0, s>>>((const float2 *)M1, r1, c1, (const float2 *)M2, c2, c1, (float2 *)res);
    else couter_kernel<double><<<grid, 256, 0, s>>>((const double2 *)M1, r1, c1, (const double2 *)M2, c2, c1, (double2 *)res);
    return cudaGetLastError();
}

// identity utils.c:84-93
template <typename T> __global__ void cidentity_kernel(cx<T> *Id, int size, T scalar, int64_t n)
{
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) { int64_t w = e % ((int64_t)size * size); Id[e] = mk<T>((w / size == w % size) ? scalar : (T)0, (T)0); }
}

cudaError_t launch_cidentity(wifi_dtype dt, void *Id, int size, double scalar, int64_t batch, cudaStream_t s)
{
    g_last_launches = 0;
    int64_t n = batch * size * size;
    if (n == 0) return cudaSuccess;
    g_last_launches = 1;
    unsigned grid = (unsigned)((n + 255) / 256);
    if (dt == WIFI_F32) cidentity_kernel<float><<<grid, 256, 0, s>>>((float2 *)Id, size, (float)scalar, n);
    else cidentity_kernel<double><<<grid, 256, 0, s>>>((double2 *)Id, size, scalar, n);
    return cudaGetLastError();
}

}  // namespace wifi
