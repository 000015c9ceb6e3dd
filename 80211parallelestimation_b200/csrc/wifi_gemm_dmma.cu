// wifi_gemm_dmma.cu -- FP64 shared-filter PS_MMSE on the FP64 tensor-core path (DMMA, mma.sync m8n8k4 f64).
//
//   H[n][53] = (rx/tx)[n][53] * W^T     as the real product  [n x 106] * [106 x 106],  K padded to 108, N to 112.
//
// tcgen05 has no FP64 kind, so the FP64 mode of the shared-filter estimator uses the warp-level DMMA.  Measured on B200
// (profiles/microbench/dmma_peak.cu): every f64 mma shape lowers to DMMA.8x8x4 and sustains 37.0 TFLOP/s = 64 FMA/clk/SM,
// the same ceiling as DFMA -- but one warp instruction carries 256 FMAs with two operand registers per lane, so the
// pipe is fed without the LDS/issue pressure that held the CUDA-core kernel at 9.8 TFLOP/s.
//
// One persistent CTA per SM.  The filter image Bt[n][k] (n-major, row stride 116 doubles so the 8 rows a fragment load
// touches fall in different bank groups) is copied to shared memory once; the kernel is warp-specialised (below).  The
// symmetric version it replaced (every warp load -> convert -> DMMA -> store, 0.90-1.06 G frames/s against 1.19) is gone.
// Replaces multiply() utils.c:16-31 applied per frame (main.c:201-207 intent) in the FP64 mode.
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

constexpr int DM_K = 108;               // 106 padded to a multiple of 4
constexpr int DM_N = 112;               // 106 padded to a multiple of 8
constexpr int DM_BS = WIFI_DMMA_BS;     // Bt row stride (doubles): 116 = 4 mod 16
constexpr int DM_AS = 108;              // A row stride (doubles): 12 mod 16
constexpr int DM_NT = DM_N / 8;         // 14
constexpr int DM_KT = DM_K / 4;         // 27

// W (53 x 53 double2, row-major) -> Bt[n][k] = B[k][n], B = real embedding of W^T acting on interleaved (re, im) rows:
//   out[2r] = sum_j Wre[r][j] y[2j] - Wim[r][j] y[2j+1],   out[2r+1] = sum_j Wim[r][j] y[2j] + Wre[r][j] y[2j+1]
__global__ void filter_install_dmma_kernel(const double2 *__restrict__ W, double *__restrict__ Bt)
{
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= DM_N * DM_BS) return;
    int n = e / DM_BS, k = e - n * DM_BS;
    double v = 0.0;
    if (n < 2 * NSC && k < 2 * NSC) {
        const double2 w = W[(n >> 1) * NSC + (k >> 1)];
        v = (n & 1) ? ((k & 1) ? w.x : w.y) : ((k & 1) ? -w.y : w.x);
    }
    Bt[e] = v;
}

cudaError_t launch_filter_install_dmma(FilterImages &img, cudaStream_t s)
{
    filter_install_dmma_kernel<<<(DM_N * DM_BS + 255) / 256, 256, 0, s>>>((const double2 *)img.W64, img.B64);
    return cudaGetLastError();
}

__device__ __forceinline__ void dmma884(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// Warp-specialised version: the DMMA pipe never waits for a conversion.
// ------------------------------------------------------------------------------------------
// 16 warps = 8 (consumer, producer) pairs, two pairs per scheduler.  A producer loads 8-frame tiles of tx / rx (L2-prefetched
// one tile ahead, software-pipelined 16-byte loads), does the LS divide and fills one of its pair's two A buffers; the
// consumer issues the tile's 27 x 14 DMMAs back to back and stores its accumulators while the producer is already refilling.
// Two DMMA warps per scheduler are needed: ptxas paces a warp's consecutive DMMAs with NOPs, and one consumer alone left
// the pipe at 73 % (ncu); one producer for two consumers could not keep up (56 %).  Hand-over by mbarriers (full / empty
// per buffer, one elected arrival after __syncwarp).
constexpr int DW_ROWS = 8;                                  // frames per tile (one m8 fragment)
constexpr int DW_PAIRS = 8;
constexpr int DW_NEL = DW_ROWS * NSC;                       // 848 complex values per tile and array
constexpr int DW_LB = 5;                                    // load batch (vectors per lane and array), 3 batches cover 424 values
constexpr int DW_THREADS = DW_PAIRS * 64;                   // 8 pairs x (consumer warp + producer warp)

// 1/x for the LS divide with as few FP64-pipe instructions as possible (the FP64 pipe is the DMMA pipe and the producers'
// arithmetic queues behind the consumers' DMMAs): FP32 hardware reciprocal of the rounded argument as the seed (23 bits),
// two Newton steps in FP64 (46, 92 bits).  |tx|^2 of a frame is far inside the FP32 range.
__device__ __forceinline__ double dm_rcp(double x)
{
    float rf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)x));
    double r = (double)rf;
    r = fma(r, fma(-x, r, 1.0), r);
    return fma(r, fma(-x, r, 1.0), r);
}

__device__ __forceinline__ uint32_t dm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dm_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dm_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void dm_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(dm_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void dm_mbar_wait(uint64_t *bar, uint32_t parity)
{
    uint32_t addr = dm_smem_u32(bar), done = 0;
    for (uint32_t spin = 0; !done; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, 0x989680;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (!done && spin > (1u << 22)) __trap();             // a protocol bug traps instead of hanging the GPU
    }
}

// EIG (the second product of the eigen-domain per-frame MMSE, wifi_eig.cu; input u = y G^T, FUSED = false): the producer
// turns the staged rows into v = s (.) (u - p z_d) -- four lanes per frame, beta and gamma by quad shuffles -- and the
// consumer's epilogue writes rx/tx - acc, the null bin from the value the producer left next to the buffer.
struct DmEig {
    const double2 *tx, *rx;     // the frames' block vectors
    int64_t stride;
    const double *sigma2;       // [n]
    const double *lam;          // [53]
    const double2 *p;           // [53]
    double Rdd, md;
    int dc;
};

template <bool FUSED, bool EIG>
__global__ void __launch_bounds__(DW_THREADS, 1)
    mmse_shared_dmma_ws_kernel(const double *__restrict__ Bt_g, const double2 *__restrict__ a_in, const double2 *__restrict__ rx,
                               int64_t frame_stride, double2 *__restrict__ H, int64_t n_frames, DmEig eg, double2 *__restrict__ hp_out)
{
    extern __shared__ __align__(16) unsigned char dm_smem[];
    double *Bt = (double *)dm_smem;                                   // [112][116]
    double *As_all = Bt + DM_N * DM_BS;                               // [pair][2][16][108]
    uint64_t *bars = (uint64_t *)(As_all + DW_PAIRS * 2 * DW_ROWS * DM_AS);   // [pair][2] full, [pair][2] empty
    double *slam = (double *)(bars + DW_PAIRS * 4);                   // EIG: [56] eigenvalues
    double2 *sp = (double2 *)(slam + 56);                             // EIG: [56] border vector
    double2 *shd = sp + 56;                                           // EIG: [pair][2][8] H of the null bin
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pair = warp & 7;                                        // warps p (consumer) and p + 8 (producer): same scheduler
    const bool producer = warp >= DW_PAIRS;
    double *As = As_all + pair * (2 * DW_ROWS * DM_AS);
    uint64_t *full = bars + pair * 4, *empty = full + 2;

    {
        const double2 *src = (const double2 *)Bt_g;
        double2 *dst = (double2 *)Bt;
        for (int e = threadIdx.x; e < DM_N * DM_BS / 2; e += DW_THREADS) dst[e] = src[e];
        if (producer && lane < 2 * DW_ROWS) { As[lane * DM_AS + 106] = 0.0; As[lane * DM_AS + 107] = 0.0; }   // K padding of both buffers
        if (threadIdx.x < DW_PAIRS * 4) dm_mbar_init(bars + threadIdx.x, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (EIG && threadIdx.x < 56) {
            slam[threadIdx.x] = threadIdx.x < NSC ? eg.lam[threadIdx.x] : 0.0;
            sp[threadIdx.x] = threadIdx.x < NSC ? eg.p[threadIdx.x] : make_double2(0.0, 0.0);
        }
    }
    __syncthreads();

    const int64_t n_tiles = (n_frames + DW_ROWS - 1) / DW_ROWS;
    const int64_t tstep = (int64_t)gridDim.x * DW_PAIRS;
    const int64_t tile0 = (int64_t)blockIdx.x * DW_PAIRS + pair;

    if (producer) {
        int it = 0;
        for (int64_t tile = tile0; tile < n_tiles; tile += tstep, ++it) {
            const int b = it & 1, k = it >> 1;
            const int64_t f0 = tile * DW_ROWS;
            const int nf = (int)min((int64_t)DW_ROWS, n_frames - f0);
            double *Ab = As + b * (DW_ROWS * DM_AS);
            // first batch of loads goes out before the wait for the buffer
            double2 tv[2][DW_LB], rv[2][DW_LB];
            auto issue = [&](int batch, int slot) {
#pragma unroll
                for (int u = 0; u < DW_LB; ++u) {
                    const int e = (batch * DW_LB + u) * 32 + lane;
                    const int f = e / NSC, kk = e - f * NSC;
                    tv[slot][u] = make_double2(FUSED ? 1.0 : 0.0, 0.0);
                    if (FUSED) rv[slot][u] = make_double2(0.0, 0.0);
                    if (e < DW_NEL && f < nf) {
                        const int64_t off = (f0 + f) * frame_stride + kk;
                        tv[slot][u] = ld_stream(a_in + off);
                        if (FUSED) rv[slot][u] = ld_stream(rx + off);
                    }
                }
            };
            issue(0, 0);
            // HBM -> L2 for the team's next tile (no registers, no shared memory): its loads then cost an L2 latency each
            if (tile + tstep < n_tiles) {
                const int64_t fn = (tile + tstep) * DW_ROWS;
                const int nfn = (int)min((int64_t)DW_ROWS, n_frames - fn);
                if (frame_stride == NSC) {
                    if (lane == 0) {
                        asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a_in + fn * NSC), "r"((uint32_t)(nfn * NSC * 16)) : "memory");
                        if (FUSED) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rx + fn * NSC), "r"((uint32_t)(nfn * NSC * 16)) : "memory");
                    }
                } else if (lane < nfn) {
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(a_in + (fn + lane) * frame_stride), "r"((uint32_t)(NSC * 16)) : "memory");
                    if (FUSED) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(rx + (fn + lane) * frame_stride), "r"((uint32_t)(NSC * 16)) : "memory");
                }
            }
            if (k > 0) dm_mbar_wait(&empty[b], (k - 1) & 1);          // the consumer has finished with this buffer
#pragma unroll
            for (int batch = 0; batch < 3; ++batch) {
                if (batch < 2) issue(batch + 1, (batch + 1) & 1);
#pragma unroll
                for (int u = 0; u < DW_LB; ++u) {
                    const int e = (batch * DW_LB + u) * 32 + lane;
                    const int f = e / NSC, kk = e - f * NSC;
                    if (e < DW_NEL) {
                        double2 y = tv[batch & 1][u];
                        if (FUSED) {                                  // per-block LS rx/tx (main.c:83); rows >= nf: 0/1 = 0
                            const double2 t = tv[batch & 1][u], r = rv[batch & 1][u];
                            const double inv = dm_rcp(t.x * t.x + t.y * t.y);
                            y = make_double2((r.x * t.x + r.y * t.y) * inv, (r.y * t.x - r.x * t.y) * inv);
                            // on request the four pilot LS values of a frame (main.c:82-84) go out as hp_out[f][4] for the interpolators
                            if (hp_out != nullptr && f < nf && (kk == WIFI_P0 || kk == WIFI_P1 || kk == WIFI_P2 || kk == WIFI_P3))
                                hp_out[(f0 + f) * 4 + (kk - WIFI_P0) / (WIFI_P1 - WIFI_P0)] = y;
                        }
                        *reinterpret_cast<double2 *>(Ab + f * DM_AS + 2 * kk) = y;
                    }
                }
            }
            __syncwarp();
            if (EIG) {
                // row g = lane / 4 of the tile, eigen indices q, q + 4, ... of this lane
                const int g = lane >> 2, q = lane & 3;
                const bool live = g < nf;
                const double s2 = live ? eg.sigma2[f0 + g] : 1.0;
                double2 td = make_double2(1.0, 0.0), rd = make_double2(0.0, 0.0);
                if (live && eg.dc >= 0) { td = eg.tx[(f0 + g) * eg.stride + eg.dc]; rd = eg.rx[(f0 + g) * eg.stride + eg.dc]; }
                if (lane < 2 * nf) {                                  // the consumer's epilogue re-reads tx / rx of these rows
                    const double2 *row = (lane < nf ? eg.tx : eg.rx) + (f0 + (lane < nf ? lane : lane - nf)) * eg.stride;
                    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(row), "r"((uint32_t)(NSC * 16)) : "memory");
                }
                double2 *urow = reinterpret_cast<double2 *>(Ab + g * DM_AS);
                double inv[14];
                double br = 0.0, bi = 0.0, ga = 0.0;
#pragma unroll
                for (int m = 0; m < 14; ++m) {
                    const int cc = q + 4 * m;
                    inv[m] = 0.0;
                    if (cc < NSC) {
                        inv[m] = dm_rcp(slam[cc] + s2);
                        const double2 u = urow[cc], pp = sp[cc];
                        br = fma(pp.x * u.x + pp.y * u.y, inv[m], br);
                        bi = fma(pp.x * u.y - pp.y * u.x, inv[m], bi);
                        ga = fma(pp.x * pp.x + pp.y * pp.y, inv[m], ga);
                    }
                }
#pragma unroll
                for (int o = 1; o <= 2; o <<= 1) {
                    br += __shfl_xor_sync(0xffffffffu, br, o);
                    bi += __shfl_xor_sync(0xffffffffu, bi, o);
                    ga += __shfl_xor_sync(0xffffffffu, ga, o);
                }
                double2 zd = make_double2(0.0, 0.0);
                if (eg.dc >= 0) {
                    const double itd = dm_rcp(td.x * td.x + td.y * td.y);
                    const double2 yd = make_double2((rd.x * td.x + rd.y * td.y) * itd, (rd.y * td.x - rd.x * td.y) * itd);
                    const double qq = eg.Rdd - ga, iden = 1.0 / (s2 * eg.md + qq);
                    const double2 dlt = make_double2(yd.x - br, yd.y - bi);
                    zd = make_double2(dlt.x * iden, dlt.y * iden);
                    if (q == 0) shd[(pair * 2 + b) * 8 + g] = make_double2(br + dlt.x * (qq * iden), bi + dlt.y * (qq * iden));
                }
#pragma unroll
                for (int m = 0; m < 14; ++m) {
                    const int cc = q + 4 * m;
                    if (cc < NSC) {
                        const double2 u = urow[cc], pp = sp[cc];
                        const double sc = s2 * inv[m];
                        urow[cc] = make_double2(sc * (u.x - (pp.x * zd.x - pp.y * zd.y)), sc * (u.y - (pp.x * zd.y + pp.y * zd.x)));
                    }
                }
                __syncwarp();
            }
            if (lane == 0) dm_mbar_arrive(&full[b]);
        }
    } else {
        const int g = lane >> 2, q = lane & 3;                        // fragment row / position in the group of 4
        const double *bp = Bt + g * DM_BS + q;
        int it = 0;
        for (int64_t tile = tile0; tile < n_tiles; tile += tstep, ++it) {
            const int b = it & 1, k = it >> 1;
            const int64_t f0 = tile * DW_ROWS;
            const int nf = (int)min((int64_t)DW_ROWS, n_frames - f0);
            const double *ap = As + b * (DW_ROWS * DM_AS) + g * DM_AS + q;
            double c[DM_NT][2];
#pragma unroll
            for (int j = 0; j < DM_NT; ++j) c[j][0] = c[j][1] = 0.0;
            dm_mbar_wait(&full[b], k & 1);
            double a0 = ap[0];
#pragma unroll 3
            for (int kt = 0; kt < DM_KT; ++kt) {
                const double na0 = ap[(kt + 1 < DM_KT ? kt + 1 : kt) * 4];          // next k-step's A fragment
#pragma unroll
                for (int j = 0; j < DM_NT; ++j) dmma884(c[j][0], c[j][1], a0, bp[j * 8 * DM_BS + kt * 4]);
                a0 = na0;
            }
            if (EIG) {
                const double2 hd = (eg.dc >= 0 && g < nf) ? shd[(pair * 2 + b) * 8 + g] : make_double2(0.0, 0.0);   // read before the release
                __syncwarp();
                if (lane == 0) dm_mbar_arrive(&empty[b]);
                if (g < nf) {
                    const double2 *ptx = eg.tx + (f0 + g) * eg.stride + q, *prx = eg.rx + (f0 + g) * eg.stride + q;
                    double2 *out = H + (f0 + g) * NSC + q;
#pragma unroll
                    for (int j = 0; j < DM_NT; ++j)
                        if (4 * j + q < NSC) {
                            const double2 t = ld_stream(ptx + 4 * j), r = ld_stream(prx + 4 * j);
                            const double inv = dm_rcp(t.x * t.x + t.y * t.y);
                            double2 h = make_double2((r.x * t.x + r.y * t.y) * inv - c[j][0], (r.y * t.x - r.x * t.y) * inv - c[j][1]);
                            if (4 * j + q == eg.dc) h = hd;
                            st_stream(out + 4 * j, h);
                        }
                }
                continue;
            }
            __syncwarp();
            if (lane == 0) dm_mbar_arrive(&empty[b]);                 // the producer may refill while the results are stored
            if (g < nf) {
                double2 *out = H + (f0 + g) * NSC + q;
#pragma unroll
                for (int j = 0; j < DM_NT; ++j)
                    if (4 * j + q < NSC) st_stream(out + 4 * j, make_double2(c[j][0], c[j][1]));
            }
        }
    }
}

template <bool FUSED, bool EIG>
static cudaError_t launch_ws(const FilterImages &img, const void *a, const void *rx, int64_t frame_stride, void *H, int64_t n_frames,
                             const DmEig &eg, cudaStream_t s, void *hp_out = nullptr)
{
    const size_t smem = sizeof(double) * (DM_N * DM_BS + DW_PAIRS * 2 * DW_ROWS * DM_AS) + sizeof(uint64_t) * DW_PAIRS * 4 +
                        sizeof(double) * 56 + sizeof(double2) * (56 + DW_PAIRS * 2 * 8);
    const int64_t n_tiles = (n_frames + DW_ROWS - 1) / DW_ROWS;
    const unsigned grid = (unsigned)std::min<int64_t>((n_tiles + DW_PAIRS - 1) / DW_PAIRS, 148);
    cudaError_t e = cudaFuncSetAttribute(mmse_shared_dmma_ws_kernel<FUSED, EIG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mmse_shared_dmma_ws_kernel<FUSED, EIG><<<grid, DW_THREADS, smem, s>>>(img.B64, (const double2 *)a, (const double2 *)rx, frame_stride,
                                                                         (double2 *)H, n_frames, eg, (double2 *)hp_out);
    return cudaGetLastError();
}

// H = rx/tx - v G2^T with v = s (.) (u - p z_d) formed from u inside the kernel: the whole second half of the eigen-domain MMSE
cudaError_t launch_mmse_shared_dmma_eig(const FilterImages &img, const void *u, const void *tx, const void *rx, int64_t frame_stride,
                                        int dc, const void *sigma2, const double *lam, const void *p, double Rdd, double md, void *H,
                                        int64_t n_frames, cudaStream_t s)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    const DmEig eg = {(const double2 *)tx, (const double2 *)rx, frame_stride, (const double *)sigma2, lam, (const double2 *)p, Rdd, md, dc};
    return launch_ws<false, true>(img, u, nullptr, NSC, H, n_frames, eg, s);
}

cudaError_t launch_mmse_shared_dmma(const FilterImages &img, const void *a, const void *rx, int64_t frame_stride, void *H,
                                    int64_t n_frames, cudaStream_t s, void *hp_out)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    const DmEig none = {nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0.0, 0.0, -1};
    return rx ? launch_ws<true, false>(img, a, rx, frame_stride, H, n_frames, none, s, hp_out)
              : launch_ws<false, false>(img, a, nullptr, frame_stride, H, n_frames, none, s);
}

}  // namespace wifi
