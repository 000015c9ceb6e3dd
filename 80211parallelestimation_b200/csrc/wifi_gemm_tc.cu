// wifi_gemm_tc.cu -- shared-filter PS_MMSE on the 5th-generation tensor cores (FP32 mode).
//
//   H[n][53] = (rx/tx)[n][53] * W^T        complex, as the real product  [n x 106] * [106 x 106]   (K, N padded to 112)
//
// 3xTF32: every FP32 operand is split into a TF32 "hi" part and a residual "lo" part and the product is accumulated as
// A_hi*B_hi + A_lo*B_hi + A_hi*B_lo in FP32 (error ~2^-21 per term), so FP32-level accuracy comes out of kind::tf32 MMAs.
//
// One persistent CTA per SM, 128-frame tiles:
//   warps 0-7 = two groups of four that alternate tiles (one thread per frame = one TMEM lane):
//       a. coalesced LDG.128 of the warp's 32-frame chunk of tx and rx (contiguous 13.5 KB each), LS divide rx/tx
//          (main.c:83 arithmetic on all 53 bins), result staged row-major in the warp's private shared-memory buffer;
//       b. each thread re-reads ITS frame, splits hi/lo and writes both operand images straight into TENSOR MEMORY
//          (tcgen05.st) -- the A operand never exists in shared memory;
//       c. epilogue of the previous tile: tcgen05.ld of the FP32 accumulators (TMEM, double buffered), staged through the
//          same private buffer so the global store is one contiguous 13.5 KB run per warp.
//   warp 8, one elected thread: 14 K-steps x 3 tcgen05.mma.kind::tf32 (M=128, N=112, K=8; A from TMEM, B = the filter's
//       hi/lo images resident in shared memory in the canonical K-major no-swizzle UMMA layout), tcgen05.commit -> mbarrier.
// The MMA of tile t overlaps the epilogue of tile t-1 and the loads of tile t+1.  HBM traffic is exactly the algorithmic
// 159 complex values per frame.
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

constexpr int TC_M = 128;                 // frames per tile
constexpr int TC_K = 112;                 // 106 padded to a multiple of 8
constexpr int TC_N = 112;                 // 106 padded to a multiple of 16
constexpr int TC_CONV_WARPS = 8;          // two groups of 4 (one warp per TMEM lane quarter and group)
constexpr int TC_THREADS = (TC_CONV_WARPS + 1) * 32;
constexpr int TC_ROWF = 2 * NSC;          // 106 floats per frame
constexpr int TC_CHUNK_F = 32 * TC_ROWF;  // 3392 floats per warp chunk
constexpr int TC_B_BYTES = TC_N * TC_K * 4;       // 50176
constexpr int TC_B_SBO = (TC_K / 4) * 128;        // 3584: byte distance between 8-row groups
constexpr int TC_B_LBO = 128;                     // byte distance between adjacent 16-byte K chunks
// TMEM columns
constexpr int TC_COL_D0 = 0, TC_COL_D1 = 128, TC_COL_AHI = 256, TC_COL_ALO = 384;

// canonical K-major SWIZZLE_NONE layout: element (n, k) of B -> float index
__host__ __device__ __forceinline__ int b_canon_index(int n, int k)
{
    return ((n >> 3) * TC_B_SBO + (k >> 2) * TC_B_LBO + (n & 7) * 16 + (k & 3) * 4) >> 2;
}

// ---- filter images: W (53x53 double2) -> real embedding B[n][k] = Wembed[k][n], split hi/lo, canonical layout ----
__global__ void filter_install_tc_kernel(const double2 *__restrict__ W, float *__restrict__ Bhi, float *__restrict__ Blo)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= TC_N * TC_K) return;
    int n = e / TC_K, k = e - n * TC_K;
    double v = 0.0;
    if (n < TC_ROWF && k < TC_ROWF) {
        double2 w = W[(n >> 1) * NSC + (k >> 1)];        // W[j][i], j = n/2 (output), i = k/2 (input)
        // out[2j] = sum Wr in[2i] - Wi in[2i+1];  out[2j+1] = sum Wi in[2i] + Wr in[2i+1]
        v = (n & 1) ? ((k & 1) ? w.x : w.y) : ((k & 1) ? -w.y : w.x);
    }
    // hi = TF32 round-to-nearest of v, lo = TF32 round-to-nearest of the remainder: hi + lo = v (1 + 2^-22)
    uint32_t hb, lb;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hb) : "f"((float)v));
    float hi = __uint_as_float(hb);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lb) : "f"((float)(v - (double)hi)));
    float lo = __uint_as_float(lb);
    int idx = b_canon_index(n, k);
    Bhi[idx] = hi;
    Blo[idx] = lo;
}

cudaError_t launch_filter_install_tc(FilterImages &img, cudaStream_t s)
{
    filter_install_tc_kernel<<<(TC_N * TC_K + 255) / 256, 256, 0, s>>>((const double2 *)img.W64, img.Bhi, img.Blo);
    return cudaGetLastError();
}

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// bounded wait: a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t addr = smem_u32(bar), done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && spin > (1u << 22)) __trap();
    }
}
// asynchronous HBM -> L2 prefetch of a 16-byte aligned range (size a multiple of 16)
__device__ __forceinline__ void l2_prefetch(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32])
{
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
          "=r"(v[30]), "=r"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]^T, kind::tf32, single CTA
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// K-major, SWIZZLE_NONE shared-memory matrix descriptor (sm_100 "version 1")
__device__ __forceinline__ uint64_t make_b_desc(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);                 // start address
    d |= (uint64_t)((TC_B_LBO >> 4) & 0x3FFF) << 16;        // leading (K-chunk) byte offset
    d |= (uint64_t)((TC_B_SBO >> 4) & 0x3FFF) << 32;        // stride (8-row group) byte offset
    d |= (uint64_t)1 << 46;                                 // descriptor version
    return d;                                               // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

constexpr uint32_t TC_IDESC = (1u << 4)            // D format F32
                              | (2u << 7)          // A format TF32
                              | (2u << 10)         // B format TF32
                              | (0u << 15)         // A K-major
                              | (0u << 16)         // B K-major
                              | ((uint32_t)(TC_N >> 3) << 17) | ((uint32_t)(TC_M >> 4) << 24);

struct TcSmem {
    float bhi[TC_B_BYTES / 4];
    float blo[TC_B_BYTES / 4];
    float chunk[TC_CONV_WARPS][TC_CHUNK_F];        // per-warp private staging (input H_ls chunk, then output chunk)
    uint64_t bar_a_ready;                          // 128 converter arrivals per tile
    uint64_t bar_mma_done[2];                      // tcgen05.commit per accumulator buffer
    uint64_t bar_b_ready;                          // the two bulk copies of the filter images
    float lam[56];                                 // MID: eigenvalues and border vector
    float2 pv[56];
    uint32_t tmem_base;
};

// rx/tx with ONE reciprocal: MUFU.RCP (rcp.approx: 1 ulp; branch-free; 0/0 still yields NaN like the reference).  The IEEE '/'
// expands to ~12 instructions and a slow-path branch per real divide, which made the LS divide two thirds of this kernel's
// instruction count; the Newton step round 1 kept after the MUFU cost another 2 x 53 instructions per frame for nothing the
// 1e-4 bound can see (0.2322 -> 0.2248 ms per 1 Mi frames without it: the converter warps are issue- and latency-bound).
__device__ __forceinline__ float2 cdiv_fast(float2 a, float2 b)
{
    const float den = fmaf(b.x, b.x, b.y * b.y);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
    return make_float2(fmaf(a.x, b.x, a.y * b.y) * r, fmaf(a.y, b.x, -a.x * b.y) * r);
}
// LS divide of one float4 (two complex bins) when FUSED, identity otherwise
template <bool FUSED> __device__ __forceinline__ float4 ls_pair(float4 a, float4 r)
{
    if (!FUSED) return a;
    float2 h0 = cdiv_fast(make_float2(r.x, r.y), make_float2(a.x, a.y));
    float2 h1 = cdiv_fast(make_float2(r.z, r.w), make_float2(a.z, a.w));
    return make_float4(h0.x, h0.y, h1.x, h1.y);
}

// Eigen-domain per-frame MMSE (wifi_eig.cu), the second of its two launches of this kernel (the first is the plain fused product
// u = (rx/tx) G^T):
//   MID    the input rows are u; the converter accumulates beta, gamma over its frame's row, forms z_d and writes
//          v = s (.) (u - p z_d), s_i = sigma2 / (l_i + sigma2), straight into the tensor-memory operand images; the null bin's H_d
//          (border formula) travels in a register from converter to epilogue;
//   RESID  the epilogue writes  H = rx/tx - acc  (acc = v G2^T): y = rx/tx is recomputed from the frames' own block vectors.
// H = y - c keeps y EXACT and passes only the correction c through the 3xTF32 product, so the error is ~2e-6 of |c|, not of |y|.
// (Round 2 measured the alternative H = (u - v) G2^T -- y reconstructed as G2 u, no second read of tx / rx: 2 120 instead of 2 968 B
// of DRAM traffic per frame and 8 % faster, but ~2e-6 of the frame's PEAK everywhere: 3e-4 at the survey's floor 1e-3 on one frame
// in ~300 against 6e-6 here.  Parity first: rejected.)
struct TcResid {
    const float2 *tx, *rx;      // the frames' block vectors
    int64_t stride;
    int dc;                     // null bin index or -1
    const float *sigma2;        // [n]
    const double *lam;          // [53] eigenvalues
    const double2 *p;           // [53] border vector
    float Rdd, md;              // R_dd and 1/|x_d|^2 of the null bin
};

// 1/x with the hardware approximation (1 ulp)
__device__ __forceinline__ float rcp_fast(float x)
{
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}

template <bool FUSED, bool RESID, bool MID>
__global__ void __launch_bounds__(TC_THREADS, 1)
    mmse_shared_tc_kernel(const float *__restrict__ Bhi_g, const float *__restrict__ Blo_g, const float2 *__restrict__ a_in,
                          const float2 *__restrict__ rx, int64_t frame_stride, float2 *__restrict__ H, int64_t n_frames, int aligned16,
                          TcResid res, float2 *__restrict__ hp_out)
{
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    TcSmem &sm = *reinterpret_cast<TcSmem *>(tc_smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---- one-time setup: filter images -> smem, barriers, TMEM ----
    if (threadIdx.x == 0) {
        mbar_init(&sm.bar_a_ready, 128);
        mbar_init(&sm.bar_mma_done[0], 1);
        mbar_init(&sm.bar_mma_done[1], 1);
        mbar_init(&sm.bar_b_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        // The two 50 KB filter images come in as TMA bulk copies that complete on an mbarrier only the MMA issuer waits for:
        // the converter warps start on their first tile at once (a 288-thread LDG/STS copy loop here cost ~5 us per launch).
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&sm.bar_b_ready)), "r"(2u * TC_B_BYTES) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm.bhi)), "l"(Bhi_g), "r"((uint32_t)TC_B_BYTES), "r"(smem_u32(&sm.bar_b_ready)) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(sm.blo)), "l"(Blo_g), "r"((uint32_t)TC_B_BYTES), "r"(smem_u32(&sm.bar_b_ready)) : "memory");
    }
    if (MID && threadIdx.x < 56) {
        const bool in = threadIdx.x < NSC;
        sm.lam[threadIdx.x] = in ? (float)res.lam[threadIdx.x] : 0.f;
        sm.pv[threadIdx.x] = in ? make_float2((float)res.p[threadIdx.x].x, (float)res.p[threadIdx.x].y) : make_float2(0.f, 0.f);
    }
    if (warp == TC_CONV_WARPS) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&sm.tmem_base)), "r"(512u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = sm.tmem_base;

    const int64_t n_tiles = (n_frames + TC_M - 1) / TC_M;
    const int my_tiles = (int)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x);   // tiles blockIdx.x, +gridDim.x, ...

    if (warp == TC_CONV_WARPS) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint64_t dhi = make_b_desc(smem_u32(sm.bhi)), dlo = make_b_desc(smem_u32(sm.blo));
            mbar_wait(&sm.bar_b_ready, 0);                                      // filter images have landed (async proxy writes)
            for (int it = 0; it < my_tiles; ++it) {
                mbar_wait(&sm.bar_a_ready, it & 1);
                tc_fence_after();
                const uint32_t d = tmem + ((it & 1) ? TC_COL_D1 : TC_COL_D0);
                // The accumulator adds with truncation once per instruction, at the magnitude of the running sum: the 28 small
                // cross products (~2^-11 of the result) go FIRST, while the sum is small, and the 14 hi x hi products last -- 14
                // full-magnitude truncations per output instead of 42.  Measured on 18 949 frames against a long-double evaluation on the host: worst bin
                // 5.0e-5 at floor 1e-3 (interleaved order: 1.44e-4), 1.2e-5 at floor 1e-2 (3.2e-5), 1.2e-6 of the frame's peak (2.5e-6);
                // the launch takes the same time.
#pragma unroll
                for (int ks = 0; ks < TC_K / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 256 >> 4);             // two 128-byte core matrices per K-step
                    umma_tf32_ts(d, tmem + TC_COL_ALO + 8 * ks, dhi + adv, TC_IDESC, ks > 0);
                    umma_tf32_ts(d, tmem + TC_COL_AHI + 8 * ks, dlo + adv, TC_IDESC, 1);
                }
#pragma unroll
                for (int ks = 0; ks < TC_K / 8; ++ks) {
                    const uint64_t adv = (uint64_t)(ks * 256 >> 4);
                    umma_tf32_ts(d, tmem + TC_COL_AHI + 8 * ks, dhi + adv, TC_IDESC, 1);
                }
                umma_commit(&sm.bar_mma_done[it & 1]);
            }
        }
    } else {
        // ================= converter / epilogue warps: thread = frame = TMEM lane =================
        // Two groups of 4 warps alternate tiles (group g owns the CTA's tiles g, g+2, ...), so the latency-bound load /
        // divide / stage phase of one tile overlaps the TMEM write, MMA and epilogue of the other group's tile.
        const int group = warp >> 2, quarter = warp & 3;
        float *buf = sm.chunk[warp];
        const uint32_t lane_base = tmem + ((uint32_t)(quarter * 32) << 16);
        const bool vec_ok = (frame_stride == NSC) && aligned16;
        float2 hd_stash = make_float2(0.f, 0.f);             // MID: H of the null bin of this lane's frame, from converter to epilogue
        for (int it = group; it < my_tiles + 2; it += 2) {
            // ---- 0. pull this warp's chunk of THIS tile from HBM into L2 now (no registers, no shared memory): the transfer
            //         runs under the epilogue of the group's previous tile below, and the LDGs of step a. then hit L2.
            //         Measured at 1 Mi frames: no prefetch 0.329 ms, this 0.238 ms; prefetching 2 / 4 / 6 tiles ahead is
            //         slower (0.281 / 0.343 / 0.355 ms) -- the prefetched-but-unread footprint of 148 CTAs starts to thrash L2. ----
            if (it < my_tiles) {
                const int64_t fn = (blockIdx.x + (int64_t)it * gridDim.x) * TC_M + quarter * 32;
                if (vec_ok) {
                    if (lane == 0 && fn + 32 <= n_frames) {
                        l2_prefetch(a_in + fn * NSC, TC_CHUNK_F * 4);
                        if (FUSED) l2_prefetch(rx + fn * NSC, TC_CHUNK_F * 4);
                    }
                } else if (fn + 32 < n_frames) {
                    // strided rows (block vectors read in place from whole frames): lane r prefetches row r, widened to
                    // 16-byte boundaries (the neighbouring value on either side belongs to the same array)
                    const float2 *ra = a_in + (fn + lane) * frame_stride;
                    const uintptr_t lo = (uintptr_t)ra & ~(uintptr_t)15, hi = ((uintptr_t)(ra + NSC) + 15) & ~(uintptr_t)15;
                    if (lo >= (uintptr_t)a_in) l2_prefetch((const void *)lo, (uint32_t)(hi - lo));
                    if (FUSED) {
                        const float2 *rr = rx + (fn + lane) * frame_stride;
                        const uintptr_t lo2 = (uintptr_t)rr & ~(uintptr_t)15, hi2 = ((uintptr_t)(rr + NSC) + 15) & ~(uintptr_t)15;
                        if (lo2 >= (uintptr_t)rx) l2_prefetch((const void *)lo2, (uint32_t)(hi2 - lo2));
                    }
                }
            }
            // ---- c. epilogue of this group's previous tile (it-2); must precede this iteration's a_ready arrival ----
            if (it >= 2) {
                const int pt = it - 2;
                if (RESID && lane == 0 && res.stride == NSC) {           // the epilogue's own inputs: HBM -> L2 while the MMAs finish
                    const int64_t fe = (blockIdx.x + (int64_t)pt * gridDim.x) * TC_M + quarter * 32;
                    if (fe + 32 <= n_frames && ((((uintptr_t)res.tx) | ((uintptr_t)res.rx)) & 15) == 0) {
                        l2_prefetch(res.tx + fe * NSC, TC_CHUNK_F * 4);
                        l2_prefetch(res.rx + fe * NSC, TC_CHUNK_F * 4);
                    }
                }
                mbar_wait(&sm.bar_mma_done[pt & 1], (pt >> 1) & 1);
                tc_fence_after();
                const int64_t tile = blockIdx.x + (int64_t)pt * gridDim.x;
                const int64_t f0 = tile * TC_M + quarter * 32;
                const int valid = (int)max((int64_t)0, min((int64_t)32, n_frames - f0));
                const uint32_t dcol = lane_base + ((pt & 1) ? TC_COL_D1 : TC_COL_D0);
                float2 *rowo = reinterpret_cast<float2 *>(buf + lane * TC_ROWF);
                // RESID: the epilogue's own inputs (tx, rx of this chunk) in four batches of 7 + 7 vectors; the first batch is
                // issued before the TMEM -> shared staging below and each next one before the previous is consumed
                constexpr int RB = 7;
                float4 qt[2][RB], qr[2][RB];
                const bool rvec = RESID && valid == 32 && res.stride == NSC &&
                                  ((((uintptr_t)res.tx) | ((uintptr_t)res.rx) | ((uintptr_t)H)) & 15) == 0;
                const float4 *pt4 = RESID ? reinterpret_cast<const float4 *>(res.tx + f0 * NSC) + lane : nullptr;
                const float4 *pr4 = RESID ? reinterpret_cast<const float4 *>(res.rx + f0 * NSC) + lane : nullptr;
                if (rvec) {
#pragma unroll
                    for (int j = 0; j < RB; ++j) { qt[0][j] = ld_stream(pt4 + 32 * j); qr[0][j] = ld_stream(pr4 + 32 * j); }
                }
                if (RESID) {
                    // (16 accumulator columns per tcgen05.ld here: the 36 vectors in flight above leave no room for 32)
#pragma unroll
                    for (int g = 0; g < TC_N / 16; ++g) {
                        uint32_t v[16];
                        tmem_ld16(dcol + 16 * g, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            int cc = g * 8 + c;
                            if (cc < NSC) rowo[cc] = make_float2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
                        }
                    }
                } else {
                    // 32 accumulator columns per tcgen05.ld and wait (four round trips to tensor memory per tile instead of seven)
#pragma unroll
                    for (int g = 0; g < 3; ++g) {
                        uint32_t v[32];
                        tmem_ld32(dcol + 32 * g, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int c = 0; c < 16; ++c) rowo[g * 16 + c] = make_float2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
                    }
                    {
                        uint32_t v[16];
                        tmem_ld16(dcol + 96, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int c = 0; c < 5; ++c) rowo[48 + c] = make_float2(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (RESID) {
                    // H = rx/tx - acc, null bin from the register the converter left
                    if (rvec) {
                        const float4 *b4 = reinterpret_cast<const float4 *>(buf) + lane;
                        float4 *po = reinterpret_cast<float4 *>(H + f0 * NSC) + lane;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            if (b < 3) {
#pragma unroll
                                for (int j = 0; j < RB; ++j) {
                                    const int i = (b + 1) * RB + j;
                                    if (i < 26 || (i == 26 && lane < 16)) { qt[(b + 1) & 1][j] = ld_stream(pt4 + 32 * i); qr[(b + 1) & 1][j] = ld_stream(pr4 + 32 * i); }
                                }
                            }
#pragma unroll
                            for (int j = 0; j < RB; ++j) {
                                const int i = b * RB + j;
                                if (i < 26 || (i == 26 && lane < 16)) {
                                    const float4 y = ls_pair<true>(qt[b & 1][j], qr[b & 1][j]), c = b4[32 * i];
                                    st_stream(po + 32 * i, make_float4(y.x - c.x, y.y - c.y, y.z - c.z, y.w - c.w));
                                }
                            }
                        }
                    } else {
                        const float2 *b2 = reinterpret_cast<const float2 *>(buf);
                        for (int e = lane; e < valid * NSC; e += 32) {
                            const int r = e / NSC, k = e - r * NSC;
                            const int64_t off = (f0 + r) * res.stride + k;
                            const float2 y = cdiv_fast(ld_stream(res.rx + off), ld_stream(res.tx + off)), c = b2[e];
                            H[f0 * NSC + e] = make_float2(y.x - c.x, y.y - c.y);
                        }
                    }
                    if (MID && res.dc >= 0) {
                        __syncwarp();                    // orders this warp's stores above before the one below (same addresses)
                        if (lane < valid) H[(f0 + lane) * NSC + res.dc] = hd_stash;
                    }
                } else if (valid == 32 && aligned16) {
                    const float4 *b4 = reinterpret_cast<const float4 *>(buf);
                    float4 *po = reinterpret_cast<float4 *>(H + f0 * NSC);
#pragma unroll 9
                    for (int i = 0; i < 27; ++i) {
                        int q = i * 32 + lane;
                        if (q < TC_CHUNK_F / 4) st_stream(po + q, b4[q]);
                    }
                } else {
                    const float2 *b2 = reinterpret_cast<const float2 *>(buf);
                    for (int e = lane; e < valid * NSC; e += 32) H[f0 * NSC + e] = b2[e];
                }
                __syncwarp();      // buf is free for the next chunk
            }
            if (it >= my_tiles) continue;
            const int64_t tile = blockIdx.x + (int64_t)it * gridDim.x;
            const int64_t f0 = tile * TC_M + quarter * 32;                       // first frame of this warp's chunk
            const int valid = (int)max((int64_t)0, min((int64_t)32, n_frames - f0));
            // MID: this lane's per-frame scalars, requested now and used after the chunk has been staged
            float m_s2 = 1.f;
            float2 m_td = make_float2(1.f, 0.f), m_rd = make_float2(0.f, 0.f);
            if (MID && lane < valid) {
                m_s2 = res.sigma2[f0 + lane];
                if (res.dc >= 0) { m_td = res.tx[(f0 + lane) * res.stride + res.dc]; m_rd = res.rx[(f0 + lane) * res.stride + res.dc]; }
            }
            // ---- a. chunk -> private buffer (row-major [32][106] floats), LS divide fused; loads software-pipelined ----
            if (vec_ok && valid == 32) {
                const float4 *pa = reinterpret_cast<const float4 *>(a_in + f0 * NSC) + lane;
                const float4 *pr = FUSED ? reinterpret_cast<const float4 *>(rx + f0 * NSC) + lane : nullptr;
                float4 *b4 = reinterpret_cast<float4 *>(buf) + lane;
                constexpr int NB = 7;                                             // 4 batches of 7 vectors (27 used)
                float4 va[2][NB], vr[2][NB];
#pragma unroll
                for (int j = 0; j < NB; ++j) { va[0][j] = ld_stream(pa + 32 * j); if (FUSED) vr[0][j] = ld_stream(pr + 32 * j); }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    if (b < 3) {
#pragma unroll
                        for (int j = 0; j < NB; ++j) {
                            const int i = (b + 1) * NB + j;
                            if (i < 26 || (i == 26 && lane < 16)) {
                                va[(b + 1) & 1][j] = ld_stream(pa + 32 * i);
                                if (FUSED) vr[(b + 1) & 1][j] = ld_stream(pr + 32 * i);
                            }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < NB; ++j) {
                        const int i = b * NB + j;
                        if (i < 26 || (i == 26 && lane < 16)) b4[32 * i] = ls_pair<FUSED>(va[b & 1][j], vr[b & 1][j]);
                    }
                }
            } else if (valid == 32) {
                // whole chunk, rows at frame_stride (or only 8-byte aligned): lane l takes sub-carriers l and l + 32 of every
                // row -- a warp load is one contiguous 256-byte (168-byte) piece of a row -- 4 rows = 16 loads in flight
                float2 *b2 = reinterpret_cast<float2 *>(buf);
                const float2 *pa = a_in + f0 * frame_stride + lane;
                const float2 *pr = FUSED ? rx + f0 * frame_stride + lane : nullptr;
                const bool second = lane + 32 < NSC;
                constexpr int RB = 4;
                float2 va[2][RB][2], vr[2][RB][2];
#pragma unroll
                for (int j = 0; j < RB; ++j) {
                    va[0][j][0] = ld_stream(pa + j * frame_stride);
                    if (FUSED) vr[0][j][0] = ld_stream(pr + j * frame_stride);
                    if (second) { va[0][j][1] = ld_stream(pa + j * frame_stride + 32); if (FUSED) vr[0][j][1] = ld_stream(pr + j * frame_stride + 32); }
                }
#pragma unroll
                for (int b = 0; b < 32 / RB; ++b) {
                    if (b + 1 < 32 / RB) {
#pragma unroll
                        for (int j = 0; j < RB; ++j) {
                            const int64_t ro = (int64_t)((b + 1) * RB + j) * frame_stride;
                            va[(b + 1) & 1][j][0] = ld_stream(pa + ro);
                            if (FUSED) vr[(b + 1) & 1][j][0] = ld_stream(pr + ro);
                            if (second) { va[(b + 1) & 1][j][1] = ld_stream(pa + ro + 32); if (FUSED) vr[(b + 1) & 1][j][1] = ld_stream(pr + ro + 32); }
                        }
                    }
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        const int r = b * RB + j;
                        b2[r * NSC + lane] = FUSED ? cdiv_fast(vr[b & 1][j][0], va[b & 1][j][0]) : va[b & 1][j][0];
                        if (second) b2[r * NSC + lane + 32] = FUSED ? cdiv_fast(vr[b & 1][j][1], va[b & 1][j][1]) : va[b & 1][j][1];
                    }
                }
            } else {
                float2 *b2 = reinterpret_cast<float2 *>(buf);
                for (int e = lane; e < 32 * NSC; e += 32) {
                    int r = e / NSC, c = e - r * NSC;
                    float2 o = make_float2(0.f, 0.f);
                    if (r < valid) {
                        int64_t off = (f0 + r) * frame_stride + c;
                        o = ld_stream(a_in + off);
                        if (FUSED) o = cdiv_fast(ld_stream(rx + off), o);
                    }
                    b2[e] = o;
                }
            }
            __syncwarp();
            // FUSED, on request: the four pilot LS values of my frame (H_ls[5, 19, 33, 47], main.c:82-84) go out as one 32-byte
            // record, hp_out[f][4] -- the interpolating estimators then need no pilot gather of their own (8 isolated values per
            // frame cost a 64-byte DRAM atom each: 512 B against these 32)
            if (FUSED && hp_out != nullptr && lane < valid) {
                const float2 *rowp = reinterpret_cast<const float2 *>(buf + lane * TC_ROWF);
                const float2 p0 = rowp[WIFI_P0], p1 = rowp[WIFI_P1], p2 = rowp[WIFI_P2], p3 = rowp[WIFI_P3];
                float4 *o = reinterpret_cast<float4 *>(hp_out + (f0 + lane) * 4);
                st_stream(o, make_float4(p0.x, p0.y, p1.x, p1.y));
                st_stream(o + 1, make_float4(p2.x, p2.y, p3.x, p3.y));
            }
            // ---- b. my frame -> hi/lo -> TMEM (A operand), after the previous tile's MMAs have released it ----
            if (it > 0) { mbar_wait(&sm.bar_mma_done[(it - 1) & 1], ((it - 1) >> 1) & 1); tc_fence_after(); }
            const float2 *row = reinterpret_cast<const float2 *>(buf + lane * TC_ROWF);
            float2 m_zd = make_float2(0.f, 0.f);
            if (MID) {
                // beta = sum conj(p_i) u_i / (l_i + s2), gamma = sum |p_i|^2 / (l_i + s2) over this lane's row (wifi_eig.cu)
                float br = 0.f, bi = 0.f, ga = 0.f;
#pragma unroll 4
                for (int cc = 0; cc < NSC; ++cc) {
                    const float inv = rcp_fast(sm.lam[cc] + m_s2);
                    const float2 u = row[cc], pp = sm.pv[cc];
                    br = fmaf((pp.x * u.x + pp.y * u.y), inv, br);
                    bi = fmaf((pp.x * u.y - pp.y * u.x), inv, bi);
                    ga = fmaf((pp.x * pp.x + pp.y * pp.y), inv, ga);
                }
                if (res.dc >= 0) {
                    const float2 yd = cdiv_fast(m_rd, m_td);
                    const float q = res.Rdd - ga, iden = rcp_fast(m_s2 * res.md + q);
                    const float2 dlt = make_float2(yd.x - br, yd.y - bi);
                    m_zd = make_float2(dlt.x * iden, dlt.y * iden);
                    hd_stash = make_float2(br + dlt.x * (q * iden), bi + dlt.y * (q * iden));
                }
            }
#pragma unroll
            for (int g = 0; g < TC_K / 16; ++g) {
                uint32_t hi[16], lo[16];
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    int cc = g * 8 + c;
                    float2 v = (cc < NSC) ? row[cc] : make_float2(0.f, 0.f);
                    if (MID && cc < NSC) {                                    // v_i = s_i (u_i - p_i z_d), s_i = s2 / (l_i + s2)
                        const float sc = m_s2 * rcp_fast(sm.lam[cc] + m_s2);
                        const float2 pp = sm.pv[cc];
                        v = make_float2(sc * (v.x - (pp.x * m_zd.x - pp.y * m_zd.y)), sc * (v.y - (pp.x * m_zd.y + pp.y * m_zd.x)));
                    }
                    // hi = v rounded to the 11 significant bits of TF32 by Veltkamp's splitting, c = v (2^13 + 1), hi = c - (c - v): three
                    // FP32 instructions and NaN-safe (cvt.rna.tf32.f32 is FOUR SASS instructions here -- add, mask, and an FSETP + SEL pair
                    // that keeps the canonical NaN 0x7fffffff from rounding up into -0; ncu: 106 of each per warp and tile).  The residual
                    // goes to tensor memory as plain FP32 bits: kind::tf32 reads the upper 19 bits of the word.
                    // (__fmul_rn / __fsub_rn: no FMA contraction -- the rounding of each step IS the algorithm)
                    const float cx_ = __fmul_rn(v.x, 8193.0f), cy_ = __fmul_rn(v.y, 8193.0f);
                    const float hx = __fsub_rn(cx_, __fsub_rn(cx_, v.x)), hy = __fsub_rn(cy_, __fsub_rn(cy_, v.y));
                    hi[2 * c] = __float_as_uint(hx); hi[2 * c + 1] = __float_as_uint(hy);
                    lo[2 * c] = __float_as_uint(v.x - hx);
                    lo[2 * c + 1] = __float_as_uint(v.y - hy);

                }
                tmem_st16(lane_base + TC_COL_AHI + 16 * g, hi);
                tmem_st16(lane_base + TC_COL_ALO + 16 * g, lo);
            }
            tmem_wait_st();
            tc_fence_before();
            mbar_arrive(&sm.bar_a_ready);
            __syncwarp();          // everyone has read its row before the buffer is reused
        }
    }

    // ---- teardown ----
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CONV_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
    }
}

template <bool FUSED, bool RESID, bool MID>
static cudaError_t launch_tc(const FilterImages &img, const void *a, const void *rx, int64_t frame_stride, void *H, int64_t n_frames,
                             const TcResid &res, cudaStream_t s, void *hp_out = nullptr)
{
    const size_t smem = sizeof(TcSmem);
    const int64_t n_tiles = (n_frames + TC_M - 1) / TC_M;
    const unsigned grid = (unsigned)std::min<int64_t>(n_tiles, 148);
    const int aligned16 = ((((uintptr_t)a) | ((uintptr_t)rx) | ((uintptr_t)H)) & 15) == 0;
    cudaError_t e = cudaFuncSetAttribute(mmse_shared_tc_kernel<FUSED, RESID, MID>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mmse_shared_tc_kernel<FUSED, RESID, MID><<<grid, TC_THREADS, smem, s>>>(img.Bhi, img.Blo, (const float2 *)a, (const float2 *)rx, frame_stride,
                                                                       (float2 *)H, n_frames, aligned16, res, (float2 *)hp_out);
    return cudaGetLastError();
}

cudaError_t launch_mmse_shared_tc(const FilterImages &img, const void *a, const void *rx, int64_t frame_stride, void *H,
                                  int64_t n_frames, cudaStream_t s, void *hp_out)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    const TcResid none = {nullptr, nullptr, 0, -1, nullptr, nullptr, nullptr, 0.f, 0.f};
    return rx ? launch_tc<true, false, false>(img, a, rx, frame_stride, H, n_frames, none, s, hp_out)
              : launch_tc<false, false, false>(img, a, nullptr, frame_stride, H, n_frames, none, s);
}

// Eigen-domain per-frame MMSE, second product: H = rx/tx - (s (.) (u - p z_d)) G2^T (img = G2), null bin from the border formula.
cudaError_t launch_mmse_shared_tc_eig_h(const FilterImages &img, const void *u, const void *tx, const void *rx, int64_t frame_stride, int dc,
                                        const void *sigma2, const double *lam, const void *p, double Rdd, double md, void *H,
                                        int64_t n_frames, cudaStream_t s)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    const TcResid res = {(const float2 *)tx, (const float2 *)rx, frame_stride, dc, (const float *)sigma2, lam, (const double2 *)p, (float)Rdd, (float)md};
    return launch_tc<false, true, true>(img, u, nullptr, NSC, H, n_frames, res, s);
}

}  // namespace wifi
