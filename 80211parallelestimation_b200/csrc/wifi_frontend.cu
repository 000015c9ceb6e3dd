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

template <typename T>
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
    for (int64_t g = (int64_t)blockIdx.x * FE_WARPS + warp; g < n_groups; g += (int64_t)gridDim.x * FE_WARPS) {
        const int64_t f = g >> 2;
        const int q = (int)(g & 3);                              // transforms 4q .. 4q+3 of frame f; transform 15 = preamble
        // HBM -> L2 for the warp's NEXT group (no registers, no shared memory): its loads below then cost an L2 latency.
        // lanes 0..3: one 64-sample block each; the preamble of a frame's last group: p1 and p2 (lanes 3, 4)
        {
            const int64_t gn = g + (int64_t)gridDim.x * FE_WARPS;
            if (gn < n_groups && lane < 5) {
                const int64_t fn = gn >> 2;
                const int bn = 4 * (int)(gn & 3) + lane;
                const cx<T> *src = nullptr;
                if (lane < 4 && bn < NBLK) src = packet + fn * FE_PKT + bn * FE_BLK + FE_CP;
                else if ((gn & 3) == 3 && lane >= 3) src = lptot + fn * FE_LP + (lane == 3 ? 96 : 32);
                if (src) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"((uint32_t)(64 * sizeof(cx<T>))) : "memory");
            }
        }
        // ---- A. samples -> tile (natural order), two consecutive samples per lane and load ----
        T nv = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int e = 2 * (i * 32 + lane);                   // 0..254: transform e / 64, samples e % 64 and + 1
            const int tt = e >> 6, n = e & 63, b = 4 * q + tt;
            cx<T> x0, x1;
            if (b < NBLK) {
                const cx<T> *src = packet + f * FE_PKT + b * FE_BLK + FE_CP + n;
                ld_pair(src, x0, x1);
            } else {
                cx<T> p10, p11, p20, p21;
                ld_pair(lptot + f * FE_LP + 96 + n, p10, p11);
                ld_pair(lptot + f * FE_LP + 32 + n, p20, p21);
                x0 = mk<T>((p10.x + p20.x) * (T)0.5, (p10.y + p20.y) * (T)0.5);
                x1 = mk<T>((p11.x + p21.x) * (T)0.5, (p11.y + p21.y) * (T)0.5);
                const T d0x = p20.x - p10.x, d0y = p20.y - p10.y, d1x = p21.x - p11.x, d1y = p21.y - p11.y;
                nv += d0x * d0x + d0y * d0y + d1x * d1x + d1y * d1y;
            }
            tile[tt * FE_TS + n] = x0;
            tile[tt * FE_TS + n + 1] = x1;
        }
        if (q == 3 && ow2) {                                     // warp-uniform
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
            if (lane == 0) ow2[f] = nv * (T)(1.0 / 128.0);
        }
        __syncwarp();
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
        frontend_kernel<float><<<grid, FE_THREADS, 0, s>>>((const float2 *)packet, (const float2 *)lptot, (float2 *)symb, (float2 *)pre_fft,
                                                           (float *)ow2, n_frames);
    else
        frontend_kernel<double><<<grid, FE_THREADS, 0, s>>>((const double2 *)packet, (const double2 *)lptot, (double2 *)symb,
                                                            (double2 *)pre_fft, (double *)ow2, n_frames);
    return cudaGetLastError();
}

}  // namespace wifi
