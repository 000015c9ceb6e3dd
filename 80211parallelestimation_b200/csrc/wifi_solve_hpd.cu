// wifi_solve_hpd.cu -- per-frame PS_MMSE solve, register-resident fast path (WIFI_SOLVE_HPD).
//
//   A_f = R + diag(sigma2_f / |tx_k|^2)   (Hermitian positive definite when R is Hermitian PSD)
//   A_f z = rx/tx,   H = R z
//
// A frame is owned by a PR x 8 grid of lanes (PR = 8: two warps, PR = 4: one warp).  Lane (pr, pc) keeps the
// 2-D-cyclic slice  a[li][lj] = M[PR li + pr][8 lj + pc]  of the bordered Hermitian matrix
//        M = [ A   . ]      (row 53 = y^H carries the right-hand side through the elimination)
//            [ y^H . ]
// in REGISTERS, and only the local positions that can lie on or below the diagonal (28 of 49 for PR = 8).
// Elimination is the symmetric (L D L^H) form without pivoting -- for a Hermitian positive-definite matrix the growth
// factor is 1 -- and is fully unrolled over the 53 steps so every register index is static.  Step k:
//   * the lanes of lane-column k%8 publish the raw column  c_i = a_ik  (i >= k) to a double-buffered 56-entry scratch;
//   * one sync (named 64-thread barrier or __syncwarp);
//   * every lane reads c for its rows and its columns and the pivot, forms t_j = conj(c_j)/a_kk and updates
//     a_ij -= c_i t_j  on its lower local positions (4 FMAs each, ~40 % fewer than the unsymmetric update);
//   * conj(c_i)/a_kk = U'_ki is stored (one entry per lane) into a packed triangle that later serves the
//     back-substitution; the entry of row 53 is the forward-substituted right-hand side.
// The cyclic distribution keeps all lanes busy as the active sub-matrix shrinks.  Back-substitution (unit-diagonal U',
// columns contiguous in shared memory) runs on one warp with shuffle broadcasts; H = R z reads R^T from shared memory.
//
// Replaces the two inverse() calls of main.c:186,201 (utils.c:141-170, O(n^5)) for the intended formula.
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

constexpr int H_N1 = NSC + 1;                   // 54 rows of the bordered matrix
constexpr int H_NLC = 7;                        // local columns (53 columns / 8)
constexpr int H_UT = 1432;                      // packed strict upper triangle incl. rhs: sum_{i=1..53} i = 1431
constexpr int H_RT = NSC * NSC + 7;             // 2816

__device__ __forceinline__ int ut_off(int j) { return (j * (j - 1)) >> 1; }   // column j of U' holds rows 0..j-1

// per-group shared-memory layout, in units of cx<T>
struct HpdSmem {
    static constexpr int UT = 0;
    static constexpr int LB = UT + H_UT;        // [2][56]
    static constexpr int YB = LB + 112;         // [56]
    static constexpr int ZB = YB + 56;          // [56]
    static constexpr int DB = ZB + 56;          // T[56] <= 28 cx
    static constexpr int GROUP = DB + 28;       // 1684
};

// frames per CTA / lane-grid height per precision (tuned on B200)
constexpr int HPD_F32_PR = 8, HPD_F32_FPC = 8, HPD_F64_FPC = 4;

template <int LANES> __device__ __forceinline__ void group_sync(int id)
{
    if (LANES == 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(LANES) : "memory");
}

template <typename T, int PR, int FPC>
__global__ void __launch_bounds__(PR * 8 * FPC, 1)
    mmse_hpd_kernel(const cx<T> *__restrict__ R, const cx<T> *__restrict__ tx, const cx<T> *__restrict__ rx, int64_t frame_stride,
                    const T *__restrict__ sigma2, cx<T> *__restrict__ H, int64_t n_frames)
{
    constexpr int LANES = PR * 8;
    constexpr int NLR = (H_N1 + PR - 1) / PR;              // local rows: 7 (PR = 8) or 14 (PR = 4)
    extern __shared__ __align__(16) unsigned char hpd_smem[];
    using S = HpdSmem;
    cx<T> *Rt = (cx<T> *)hpd_smem;                         // Rt[j*53 + i] = R[i][j]
    const int grp = threadIdx.x / LANES, lane = threadIdx.x % LANES;
    const int pr = lane >> 3, pc = lane & 7;
    cx<T> *gs = Rt + H_RT + grp * S::GROUP;
    cx<T> *Ut = gs + S::UT, *lb = gs + S::LB, *yb = gs + S::YB, *zb = gs + S::ZB;
    T *db = (T *)(gs + S::DB);
    const int bar_id = grp + 1;

    for (int e = threadIdx.x; e < NSC * NSC; e += LANES * FPC) {
        int i = e / NSC, j = e - i * NSC;
        Rt[j * NSC + i] = R[e];
    }
    __syncthreads();

    for (int64_t f = (int64_t)blockIdx.x * FPC + grp; f < n_frames; f += (int64_t)gridDim.x * FPC) {
        // ---- per-frame inputs: y = rx/tx, d = sigma2/|tx|^2 ----
        for (int k = lane; k < NSC; k += LANES) {
            cx<T> t = ld_stream(tx + f * frame_stride + k), r = ld_stream(rx + f * frame_stride + k);
            yb[k] = cdiv(r, t);
            db[k] = sigma2[f] / cabs2(t);
        }
        group_sync<LANES>(bar_id);
        // local slice of the bordered matrix; only positions that can be on/below the diagonal are ever touched
        cx<T> a[NLR][H_NLC];
#pragma unroll
        for (int li = 0; li < NLR; ++li) {
            const int i = PR * li + pr;
#pragma unroll
            for (int lj = 0; lj < H_NLC; ++lj) {
                if (PR * li + PR - 1 >= 8 * lj) {
                    const int j = 8 * lj + pc;
                    cx<T> v = mk<T>(0, 0);
                    if (j < NSC) {
                        if (i < NSC) { v = Rt[j * NSC + i]; if (i == j) v.x += db[i]; }
                        else if (i == NSC) v = cconj(yb[j]);
                    }
                    a[li][lj] = v;
                }
            }
        }

        // ---- symmetric elimination, fully unrolled ----
#pragma unroll
        for (int K = 0; K < NSC; ++K) {
            const int kc = K & 7, klc = K >> 3;            // owner lane-column and local column of column K
            const int klr = K / PR;                        // first local row that can hold a row >= K
            cx<T> *lbk = lb + (K & 1) * 56;
            if (pc == kc) {
#pragma unroll
                for (int li = klr; li < NLR; ++li)
                    if (PR * li + PR - 1 >= 8 * klc) lbk[PR * li + pr] = a[li][klc];
            }
            group_sync<LANES>(bar_id);
            const T inv = (T)1 / lbk[K].x;                 // pivot a_KK is real for a Hermitian matrix
            cx<T> t[H_NLC];
#pragma unroll
            for (int lj = klc; lj < H_NLC; ++lj) {
                const int j = 8 * lj + pc;
                cx<T> c = lbk[j];
                t[lj] = (j > K && j < NSC) ? mk<T>(c.x * inv, -c.y * inv) : mk<T>(0, 0);
            }
#pragma unroll
            for (int li = klr; li < NLR; ++li) {
                const int i = PR * li + pr;
                cx<T> c = lbk[i];
                if (!(i > K && i < H_N1)) c = mk<T>(0, 0);           // rows <= K are finished, rows >= 54 do not exist
                if (pc == (li & 7) && i > K && i < H_N1) Ut[ut_off(i) + K] = mk<T>(c.x * inv, -c.y * inv);   // U'_Ki
#pragma unroll
                for (int lj = klc; lj < H_NLC; ++lj)
                    if (PR * li + PR - 1 >= 8 * lj) cfms(a[li][lj], c, t[lj]);
            }
        }
        group_sync<LANES>(bar_id);

        // ---- back-substitution on the first warp of the group: U' has unit diagonal, column j = rows 0..j-1 contiguous ----
        if (lane < 32) {
            const cx<T> *ycol = Ut + ut_off(NSC);
            cx<T> y0 = ycol[lane], y1 = (lane + 32 < NSC) ? ycol[lane + 32] : mk<T>(0, 0);
            for (int j = NSC - 1; j >= 32; --j) {
                cx<T> zj = mk<T>(__shfl_sync(0xffffffffu, y1.x, j - 32), __shfl_sync(0xffffffffu, y1.y, j - 32));
                const cx<T> *col = Ut + ut_off(j);
                cfms(y0, col[lane], zj);
                if (lane + 32 < j) cfms(y1, col[lane + 32], zj);
            }
            for (int j = 31; j >= 1; --j) {
                cx<T> zj = mk<T>(__shfl_sync(0xffffffffu, y0.x, j), __shfl_sync(0xffffffffu, y0.y, j));
                if (lane < j) cfms(y0, (Ut + ut_off(j))[lane], zj);
            }
            zb[lane] = y0;
            if (lane + 32 < NSC) zb[lane + 32] = y1;
        }
        group_sync<LANES>(bar_id);
        // ---- H = R z ----
        for (int i = lane; i < NSC; i += LANES) {
            cx<T> acc = mk<T>(0, 0);
#pragma unroll 4
            for (int j = 0; j < NSC; ++j) cfma(acc, Rt[j * NSC + i], zb[j]);
            st_stream(H + f * NSC + i, acc);
        }
    }
}

template <typename T, int PR, int FPC>
static cudaError_t launch_hpd(const void *R, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, void *H,
                              int64_t n_frames, cudaStream_t s)
{
    using S = HpdSmem;
    size_t smem = sizeof(cx<T>) * (H_RT + FPC * S::GROUP);
    cudaError_t e = cudaFuncSetAttribute(mmse_hpd_kernel<T, PR, FPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t need = (n_frames + FPC - 1) / FPC;
    unsigned grid = (unsigned)std::min<int64_t>(need, 148);
    mmse_hpd_kernel<T, PR, FPC><<<grid, PR * 8 * FPC, smem, s>>>((const cx<T> *)R, (const cx<T> *)tx, (const cx<T> *)rx, frame_stride,
                                                               (const T *)sigma2, (cx<T> *)H, n_frames);
    return cudaGetLastError();
}

cudaError_t launch_mmse_perframe_hpd(wifi_dtype dt, const void *R, const void *tx, const void *rx, int64_t frame_stride,
                                     const void *sigma2, void *H, int64_t n_frames, int refine, const void *R64, cudaStream_t s)
{
    (void)R64;
    g_last_launches = 0;
    if (refine) return cudaErrorNotSupported;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    if (dt == WIFI_F32) return launch_hpd<float, HPD_F32_PR, HPD_F32_FPC>(R, tx, rx, frame_stride, sigma2, H, n_frames, s);
    return launch_hpd<double, 8, HPD_F64_FPC>(R, tx, rx, frame_stride, sigma2, H, n_frames, s);
}

}  // namespace wifi
