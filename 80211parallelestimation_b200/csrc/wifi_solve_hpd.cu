// wifi_solve_hpd.cu -- per-frame PS_MMSE solve, register-resident fast path (WIFI_SOLVE_HPD).
//
//   A_f = R + diag(sigma2_f / |tx_k|^2)   (Hermitian positive definite when R is Hermitian PSD)
//   A_f z = rx/tx,   H = R z
//
// One frame is owned by 64 lanes (two warps) arranged as an 8 x 8 grid; lane (pr, pc) keeps the
// 2-D-cyclic slice  a[li][lj] = [A | y][8 li + pr][8 lj + pc]  (7 x 7 complex values, 54 columns with the
// right-hand side) entirely in REGISTERS.  Elimination is un-pivoted (for a Hermitian positive-definite
// matrix the growth factor is 1, so pivoting buys nothing) and fully unrolled over the 53 steps so every
// register index is static.  Per step only the pivot row (normalised by the pivot, owners = one lane-row)
// and the pivot column (owners = one lane-column) go through shared memory: the row into a packed
// upper-triangular store that doubles as U for the back-substitution, the column into a double-buffered
// 64-entry scratch; one 64-thread named barrier per step.  The cyclic distribution keeps all 64 lanes busy
// as the active sub-matrix shrinks.  Back-substitution (unit-diagonal U, columns contiguous in shared
// memory) runs on one warp with shuffle broadcasts; H = R z reads R^T from shared memory.
//
// Replaces the two inverse() calls of main.c:186,201 (utils.c:141-170, O(n^5)) for the intended formula.
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

constexpr int HG = 64;                          // lanes per frame
constexpr int HL = 7;                           // local rows / cols per lane
constexpr int H_UT = 1432;                      // packed strict upper triangle incl. rhs column: sum_{j=1..53} j = 1431
constexpr int H_RT = NSC * NSC + 7;             // 2816

__device__ __forceinline__ int ut_off(int j) { return (j * (j - 1)) >> 1; }   // column j holds rows 0..j-1

template <typename T> struct HpdSmem {
    // per-group layout in units of cx<T>
    static constexpr int UT = 0;
    static constexpr int LB = UT + H_UT;        // [2][64]
    static constexpr int YB = LB + 128;         // [56]
    static constexpr int ZB = YB + 56;          // [56]
    static constexpr int DB = ZB + 56;          // T[56] = 28 cx
    static constexpr int GROUP = DB + 28;       // 1700
};

__device__ __forceinline__ void group_barrier(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(HG) : "memory"); }

template <typename T, int FPC>
__global__ void __launch_bounds__(HG *FPC, 1)
    mmse_hpd_kernel(const cx<T> *__restrict__ R, const cx<T> *__restrict__ tx, const cx<T> *__restrict__ rx, int64_t frame_stride,
                    const T *__restrict__ sigma2, cx<T> *__restrict__ H, int64_t n_frames)
{
    extern __shared__ __align__(16) unsigned char hpd_smem[];
    using S = HpdSmem<T>;
    cx<T> *Rt = (cx<T> *)hpd_smem;                         // Rt[j*53 + i] = R[i][j]
    const int grp = threadIdx.x / HG, lane = threadIdx.x % HG;
    const int pr = lane >> 3, pc = lane & 7;
    cx<T> *gs = Rt + H_RT + grp * S::GROUP;
    cx<T> *Ut = gs + S::UT, *lb = gs + S::LB, *yb = gs + S::YB, *zb = gs + S::ZB;
    T *db = (T *)(gs + S::DB);
    const int bar_id = grp + 1;

    for (int e = threadIdx.x; e < NSC * NSC; e += HG * FPC) {
        int i = e / NSC, j = e - i * NSC;
        Rt[j * NSC + i] = R[e];
    }
    __syncthreads();

    for (int64_t f = (int64_t)blockIdx.x * FPC + grp; f < n_frames; f += (int64_t)gridDim.x * FPC) {
        // ---- per-frame inputs: y = rx/tx, d = sigma2/|tx|^2 ----
        if (lane < NSC) {
            cx<T> t = ld_stream(tx + f * frame_stride + lane), r = ld_stream(rx + f * frame_stride + lane);
            yb[lane] = cdiv(r, t);
            db[lane] = sigma2[f] / cabs2(t);
        }
        group_barrier(bar_id);
        cx<T> a[HL][HL];
#pragma unroll
        for (int li = 0; li < HL; ++li) {
            const int i = 8 * li + pr;
#pragma unroll
            for (int lj = 0; lj < HL; ++lj) {
                const int j = 8 * lj + pc;
                cx<T> v = mk<T>(0, 0);
                if (i < NSC) {
                    if (j < NSC) { v = Rt[j * NSC + i]; if (i == j) v.x += db[i]; }
                    else if (j == NSC) v = yb[i];
                }
                a[li][lj] = v;
            }
        }

        // ---- elimination, fully unrolled ----
#pragma unroll
        for (int K = 0; K < NSC; ++K) {
            const int kr = K & 7, kl = K >> 3;
            cx<T> *lbk = lb + (K & 1) * 64;
            if (pr == kr) {
                // pivot-row owners (8 lanes of one warp): fetch the pivot from lane pc == kr, publish the normalised row
                const unsigned m8 = 0xFFu << (8 * (kr & 3));
                const int src = 8 * (kr & 3) + kr;
                T px = __shfl_sync(m8, a[kl][kl].x, src), py = __shfl_sync(m8, a[kl][kl].y, src);
                const cx<T> inv = crecip(mk<T>(px, py));
#pragma unroll
                for (int lj = kl; lj < HL; ++lj) {
                    const int j = 8 * lj + pc;
                    if (j > K && j <= NSC) Ut[ut_off(j) + K] = cmul(a[kl][lj], inv);
                }
            }
            if (pc == kr) {
                // pivot-column owners: raw column entries of the local rows
#pragma unroll
                for (int li = kl; li < HL; ++li) lbk[pr * 8 + li] = a[li][kl];
            }
            group_barrier(bar_id);
            cx<T> u[HL];
#pragma unroll
            for (int lj = kl; lj < HL; ++lj) {
                const int j = 8 * lj + pc;
                u[lj] = (j > K && j <= NSC) ? Ut[ut_off(j) + K] : mk<T>(0, 0);
            }
#pragma unroll
            for (int li = kl; li < HL; ++li) {
                cx<T> c = lbk[pr * 8 + li];
                if (li == kl && pr <= kr) c = mk<T>(0, 0);          // rows <= K are finished
#pragma unroll
                for (int lj = kl; lj < HL; ++lj) cfms(a[li][lj], c, u[lj]);
            }
        }

        // ---- back-substitution on warp 0 of the group: U has unit diagonal, column j = rows 0..j-1 contiguous ----
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
        group_barrier(bar_id);
        // ---- H = R z ----
        if (lane < NSC) {
            cx<T> acc = mk<T>(0, 0);
#pragma unroll 4
            for (int j = 0; j < NSC; ++j) cfma(acc, Rt[j * NSC + lane], zb[j]);
            st_stream(H + f * NSC + lane, acc);
        }
    }
}

template <typename T, int FPC>
static cudaError_t launch_hpd(const void *R, const void *tx, const void *rx, int64_t frame_stride, const void *sigma2, void *H,
                              int64_t n_frames, cudaStream_t s)
{
    using S = HpdSmem<T>;
    size_t smem = sizeof(cx<T>) * (H_RT + FPC * S::GROUP);
    cudaError_t e = cudaFuncSetAttribute(mmse_hpd_kernel<T, FPC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int64_t need = (n_frames + FPC - 1) / FPC;
    unsigned grid = (unsigned)std::min<int64_t>(need, 148);
    mmse_hpd_kernel<T, FPC><<<grid, HG * FPC, smem, s>>>((const cx<T> *)R, (const cx<T> *)tx, (const cx<T> *)rx, frame_stride,
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
    if (dt == WIFI_F32) return launch_hpd<float, 8>(R, tx, rx, frame_stride, sigma2, H, n_frames, s);
    return launch_hpd<double, 4>(R, tx, rx, frame_stride, sigma2, H, n_frames, s);
}

}  // namespace wifi
