// wifi_ls.cu -- the HBM-bound estimators: LT_LS (main.c:66-75), pilot-LS + Linear/Cubic/Sinc
// interpolation (main.c:77-146, utils.c:727-733) and the equalizer (WiFi_Equalization.m:1-9).
//
// All three are streaming kernels: every input byte is read once, every output byte written once.
//   lt_ls      flat over the contiguous [n][53] arrays, 16-byte vectors per thread, 4 vectors in flight.
//   ps_interp  a CTA owns a tile of frames; the 8 pilot values of each frame are gathered (sector-granular,
//              the other 45 sub-carriers of the row are never fetched) and divided once into shared memory;
//              all requested interpolators are then produced from that staged tile.  Each thread owns ONE
//              sub-carrier k (blockDim = 6*53), so its 4 real weights per estimator live in registers and the
//              stores of a warp are one contiguous run of the [n][53] output.
//   equalize   flat over the contiguous [n][15][53] arrays like lt_ls (16-byte vectors, 4 in flight); the two channel
//              values of an element are gathered through L1 (each is reused by the 15 OFDM blocks of its frame).
#include <algorithm>
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

thread_local int g_last_launches = 0;

// ------------------------------------------------------------------------------------------
// LT_LS
// ------------------------------------------------------------------------------------------
// lt_ls_one() -- main.c:69-72 as written -- lives in wifi_common.cuh (shared with the fused receiver chain)
constexpr int LT_THREADS = 256;
constexpr int LT_UNROLL = 4;

__global__ void __launch_bounds__(LT_THREADS) lt_ls_f32_kernel(const float4 *__restrict__ tx, const float4 *__restrict__ rx,
                                                               float4 *__restrict__ H, int64_t n_vec)
{
    // one float4 = 2 complex elements; element index e = 2*v
    const int64_t base = (int64_t)blockIdx.x * (LT_THREADS * LT_UNROLL);
    const int kbase = (int)((2 * base) % NSC);
    float4 a[LT_UNROLL], b[LT_UNROLL];
#pragma unroll
    for (int j = 0; j < LT_UNROLL; ++j) {
        int64_t v = base + j * LT_THREADS + threadIdx.x;
        if (v < n_vec) { a[j] = ld_stream(tx + v); b[j] = ld_stream(rx + v); }
    }
#pragma unroll
    for (int j = 0; j < LT_UNROLL; ++j) {
        int local = j * LT_THREADS + threadIdx.x;
        int64_t v = base + local;
        if (v < n_vec) {
            int k0 = (kbase + 2 * local) % NSC;
            float2 h0 = lt_ls_one<float>(make_float2(a[j].x, a[j].y), make_float2(b[j].x, b[j].y));
            float2 h1 = lt_ls_one<float>(make_float2(a[j].z, a[j].w), make_float2(b[j].z, b[j].w));
            if (k0 == DCBIN) h0 = make_float2(0.f, 0.f);             // main.c:74
            if (k0 == DCBIN - 1) h1 = make_float2(0.f, 0.f);
            st_stream(H + v, make_float4(h0.x, h0.y, h1.x, h1.y));
        }
    }
}

__global__ void __launch_bounds__(LT_THREADS) lt_ls_f64_kernel(const double2 *__restrict__ tx, const double2 *__restrict__ rx,
                                                               double2 *__restrict__ H, int64_t n_elems)
{
    const int64_t base = (int64_t)blockIdx.x * (LT_THREADS * LT_UNROLL);
    const int kbase = (int)(base % NSC);
    double2 a[LT_UNROLL], b[LT_UNROLL];
#pragma unroll
    for (int j = 0; j < LT_UNROLL; ++j) {
        int64_t e = base + j * LT_THREADS + threadIdx.x;
        if (e < n_elems) { a[j] = ld_stream(tx + e); b[j] = ld_stream(rx + e); }
    }
#pragma unroll
    for (int j = 0; j < LT_UNROLL; ++j) {
        int local = j * LT_THREADS + threadIdx.x;
        int64_t e = base + local;
        if (e < n_elems) {
            double2 h = lt_ls_one<double>(a[j], b[j]);
            if ((kbase + local) % NSC == DCBIN) h = make_double2(0.0, 0.0);
            st_stream(H + e, h);
        }
    }
}

// element-wise float path: the odd tail element (53 n odd) and arrays that are only 8-byte aligned (e.g. a view that
// starts at an odd frame), elements [e0, e1)
__global__ void lt_ls_f32_scalar_kernel(const float2 *tx, const float2 *rx, float2 *H, int64_t e0, int64_t e1)
{
    const int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= e1) return;
    float2 h = lt_ls_one<float>(tx[e], rx[e]);
    if (e % NSC == DCBIN) h = make_float2(0.f, 0.f);
    H[e] = h;
}

cudaError_t launch_lt_ls(wifi_dtype dt, const void *tx, const void *rx, void *H, int64_t n_frames, cudaStream_t s)
{
    g_last_launches = 0;
    const int64_t n_elems = n_frames * NSC;
    if (n_elems == 0) return cudaSuccess;
    const int64_t per_block = LT_THREADS * LT_UNROLL;
    if (dt == WIFI_F32) {
        const bool vec = ((((uintptr_t)tx) | ((uintptr_t)rx) | ((uintptr_t)H)) & 15) == 0;
        int64_t n_vec = vec ? n_elems / 2 : 0;
        if (n_vec) {
            lt_ls_f32_kernel<<<(unsigned)((n_vec + per_block - 1) / per_block), LT_THREADS, 0, s>>>(
                (const float4 *)tx, (const float4 *)rx, (float4 *)H, n_vec);
            ++g_last_launches;
        }
        if (2 * n_vec < n_elems) {
            const int64_t rest = n_elems - 2 * n_vec;
            lt_ls_f32_scalar_kernel<<<(unsigned)((rest + 255) / 256), 256, 0, s>>>((const float2 *)tx, (const float2 *)rx, (float2 *)H,
                                                                                  2 * n_vec, n_elems);
            ++g_last_launches;
        }
    } else {
        lt_ls_f64_kernel<<<(unsigned)((n_elems + per_block - 1) / per_block), LT_THREADS, 0, s>>>(
            (const double2 *)tx, (const double2 *)rx, (double2 *)H, n_elems);
        ++g_last_launches;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// pilot LS + Linear / Cubic / Sinc
// ------------------------------------------------------------------------------------------
constexpr int PS_FPP = 6;                  // frames per pass
constexpr int PS_THREADS = PS_FPP * NSC;   // 318: thread t owns sub-carrier t % 53 (2 idle lanes in 10 warps)
constexpr int PS_PASSES = 8;
constexpr int PS_TILE = PS_FPP * PS_PASSES;  // 48 frames per CTA

template <typename T>
__global__ void __launch_bounds__(PS_THREADS) ps_interp_kernel(const cx<T> *__restrict__ tx, const cx<T> *__restrict__ rx,
                                                               int64_t frame_stride, cx<T> *__restrict__ Hl,
                                                               cx<T> *__restrict__ Hc, cx<T> *__restrict__ Hs, int which,
                                                               int64_t n_frames, const T *__restrict__ wtab, const cx<T> *__restrict__ hp_in)
{
    __shared__ cx<T> hp[PS_TILE][4];
    const int64_t f0 = (int64_t)blockIdx.x * PS_TILE;
    const int nf = (int)min((int64_t)PS_TILE, n_frames - f0);

    // phase 1: pilot LS  Hp_i = rx[P_i] / tx[P_i]  (main.c:82-84), one (frame, pilot) pair per thread
    // MATLAB mode (WiFi_channel_estimation_PS_*.m): the estimators are linear in Hp, so the average of the estimates of OFDM
    // blocks 1..4 is the estimate of the averaged pilot LS values
    const int navg = (which & WIFI_PS_MATLAB) ? 4 : 1;
    for (int idx = threadIdx.x; idx < nf * 4; idx += PS_THREADS) {
        int f = idx >> 2, p = idx & 3;
        if (hp_in != nullptr) {                 // pilot LS values handed over by the PS_MMSE GEMM kernel of the same call: hp_in[f][4]
            hp[f][p] = ld_stream(hp_in + (f0 + f) * 4 + p);
            continue;
        }
        int64_t off = (f0 + f) * frame_stride + (WIFI_P0 + (WIFI_P1 - WIFI_P0) * p);
        cx<T> h = cdiv(ld_gather(rx + off), ld_gather(tx + off));
        if (navg == 4) {
#pragma unroll
            for (int b = 1; b < 4; ++b) h = cadd(h, cdiv(ld_gather(rx + off + b * NSC), ld_gather(tx + off + b * NSC)));
            h = mk<T>(h.x * (T)0.25, h.y * (T)0.25);
        }
        hp[f][p] = h;
    }
    // this thread's sub-carrier and its weights (registers)
    const int k = threadIdx.x % NSC, fsub = threadIdx.x / NSC;
    T wl[4], wc[4], ws[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        wl[i] = __ldg(wtab + (0 * NSC + k) * 4 + i);
        wc[i] = __ldg(wtab + (((which & WIFI_PS_MATLAB) ? 3 : 1) * NSC + k) * 4 + i);   // table 3: true divided differences
        ws[i] = __ldg(wtab + (2 * NSC + k) * 4 + i);
    }
    __syncthreads();

    // phase 2: H_k = sum_i w[k][i] Hp_i ; a warp writes one contiguous run of the [n][53] output
#pragma unroll
    for (int pass = 0; pass < PS_PASSES; ++pass) {
        int f = pass * PS_FPP + fsub;
        if (f < nf) {
            cx<T> h0 = hp[f][0], h1 = hp[f][1], h2 = hp[f][2], h3 = hp[f][3];
            int64_t o = (f0 + f) * NSC + k;
            if (which & WIFI_PS_LINEAR) {
                cx<T> v = mk<T>(wl[0] * h0.x + wl[1] * h1.x + wl[2] * h2.x + wl[3] * h3.x,
                                wl[0] * h0.y + wl[1] * h1.y + wl[2] * h2.y + wl[3] * h3.y);
                st_stream(Hl + o, v);
            }
            if (which & WIFI_PS_CUBIC) {
                cx<T> v = mk<T>(wc[0] * h0.x + wc[1] * h1.x + wc[2] * h2.x + wc[3] * h3.x,
                                wc[0] * h0.y + wc[1] * h1.y + wc[2] * h2.y + wc[3] * h3.y);
                st_stream(Hc + o, v);
            }
            if (which & WIFI_PS_SINC) {
                cx<T> v = mk<T>(ws[0] * h0.x + ws[1] * h1.x + ws[2] * h2.x + ws[3] * h3.x,
                                ws[0] * h0.y + ws[1] * h1.y + ws[2] * h2.y + ws[3] * h3.y);
                st_stream(Hs + o, v);
            }
        }
    }
}

cudaError_t launch_ps(wifi_dtype dt, int which, const void *tx, const void *rx, int64_t frame_stride, void *Hl, void *Hc,
                      void *Hs, int64_t n_frames, const InterpTables &tab, cudaStream_t s, const void *hp_in)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    unsigned grid = (unsigned)((n_frames + PS_TILE - 1) / PS_TILE);
    if (dt == WIFI_F32)
        ps_interp_kernel<float><<<grid, PS_THREADS, 0, s>>>((const float2 *)tx, (const float2 *)rx, frame_stride, (float2 *)Hl,
                                                            (float2 *)Hc, (float2 *)Hs, which, n_frames, tab.w32, (const float2 *)hp_in);
    else
        ps_interp_kernel<double><<<grid, PS_THREADS, 0, s>>>((const double2 *)tx, (const double2 *)rx, frame_stride,
                                                             (double2 *)Hl, (double2 *)Hc, (double2 *)Hs, which, n_frames,
                                                             tab.w64, (const double2 *)hp_in);
    g_last_launches = 1;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// equalizer
// ------------------------------------------------------------------------------------------
constexpr int EQ_THREADS = 256;
constexpr int EQ_UNROLL = 4;

// eq_div() -- r / h with one reciprocal in FP32 -- lives in wifi_common.cuh (shared with the fused receiver chain)

// one equalized value: element 795 (fb + q) + rem' of the flat [n][15][53] array, given as frame base fb and a 32-bit
// offset rem >= 0 from that frame's first element (64-bit divisions per element cost more than the memory traffic)
template <typename T>
__device__ __forceinline__ cx<T> equalize_one(cx<T> r, int64_t fb, unsigned rem, const cx<T> *__restrict__ Hlt, const cx<T> *__restrict__ Hps)
{
    const unsigned q = rem / FRAME;
    rem -= q * FRAME;
    const unsigned b = rem / NSC, k = rem - b * NSC;
    if (k == DCBIN) return mk<T>(0, 0);                                 // WiFi_Equalization.m:6-7 leaves row 27 at zero
    const int64_t g = (fb + q) * NSC + k;
    const cx<T> hl = __ldg(Hlt + g), hp = __ldg(Hps + g);               // reused by the 15 blocks of the frame: L1 / L2 hits
    const T wp = (T)(b + 1) * (T)(1.0 / NBLK), wl = (T)1 - wp;          // .m:4-5, i = b+1: (15-i)/15, i/15
    return eq_div(r, mk<T>(wl * hl.x + wp * hp.x, wl * hl.y + wp * hp.y));
}

// Flat over the contiguous [n][15][53] arrays like lt_ls: 16-byte vectors, EQ_UNROLL in flight per thread.  (The first
// version -- one thread per (frame, k), 15 eight-byte loads each -- reached 82 % of the HBM peak in FP32.)
__global__ void __launch_bounds__(EQ_THREADS) equalize_f32_kernel(const float4 *__restrict__ rx, const float2 *__restrict__ Hlt,
                                                                  const float2 *__restrict__ Hps, float4 *__restrict__ eq, int64_t n_vec)
{
    const int64_t base = (int64_t)blockIdx.x * (EQ_THREADS * EQ_UNROLL);
    const int64_t fb = (2 * base) / FRAME;                              // one 64-bit division per thread
    const unsigned rb = (unsigned)(2 * base - fb * FRAME);
    float4 r[EQ_UNROLL];
#pragma unroll
    for (int j = 0; j < EQ_UNROLL; ++j) {
        int64_t v = base + j * EQ_THREADS + threadIdx.x;
        if (v < n_vec) r[j] = ld_stream(rx + v);
    }
#pragma unroll
    for (int j = 0; j < EQ_UNROLL; ++j) {
        const unsigned local = j * EQ_THREADS + threadIdx.x;
        int64_t v = base + local;
        if (v < n_vec) {
            float2 o0 = equalize_one<float>(make_float2(r[j].x, r[j].y), fb, rb + 2 * local, Hlt, Hps);
            float2 o1 = equalize_one<float>(make_float2(r[j].z, r[j].w), fb, rb + 2 * local + 1, Hlt, Hps);
            st_stream(eq + v, make_float4(o0.x, o0.y, o1.x, o1.y));
        }
    }
}

// FP64: one thread per (frame, sub-carrier), both channel values in registers, 15 independent 16-byte loads/stores down
// the OFDM blocks (consecutive threads = consecutive k = coalesced): 98 % of the HBM peak, better than the flat form (94 %).
__global__ void __launch_bounds__(EQ_THREADS) equalize_f64_kernel(const double2 *__restrict__ rx, const double2 *__restrict__ Hlt,
                                                                  const double2 *__restrict__ Hps, double2 *__restrict__ eq, int64_t n_fk)
{
    const int64_t g = (int64_t)blockIdx.x * EQ_THREADS + threadIdx.x;   // g = 53*f + k
    if (g >= n_fk) return;
    const int64_t f = g / NSC;
    const int k = (int)(g - f * NSC);
    const int64_t e0 = f * FRAME + k;                                   // element [f][0][k]
    if (k == DCBIN) {                                                   // WiFi_Equalization.m:6-7 leaves row 27 at zero
#pragma unroll
        for (int b = 0; b < NBLK; ++b) st_stream(eq + e0 + b * NSC, make_double2(0.0, 0.0));
        return;
    }
    const double2 hl = ld_stream(Hlt + g), hpv = ld_stream(Hps + g);
    double2 r[NBLK];
#pragma unroll
    for (int b = 0; b < NBLK; ++b) r[b] = ld_stream(rx + e0 + b * NSC);
#pragma unroll
    for (int b = 0; b < NBLK; ++b) {
        const double wp = (double)(b + 1) * (1.0 / NBLK), wl = 1.0 - wp;   // .m:4-5, i = b+1
        st_stream(eq + e0 + b * NSC, cdiv(r[b], make_double2(wl * hl.x + wp * hpv.x, wl * hl.y + wp * hpv.y)));
    }
}

// element-wise float path: odd tail element (795 n odd) and arrays that are only 8-byte aligned, elements [e0, e1)
__global__ void equalize_f32_scalar_kernel(const float2 *rx, const float2 *Hlt, const float2 *Hps, float2 *eq, int64_t e0, int64_t e1)
{
    const int64_t e = e0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < e1) eq[e] = equalize_one<float>(rx[e], e / FRAME, (unsigned)(e % FRAME), Hlt, Hps);
}

cudaError_t launch_equalize(wifi_dtype dt, const void *rx, const void *Hlt, const void *Hps, void *eq, int64_t n_frames,
                            cudaStream_t s)
{
    g_last_launches = 0;
    const int64_t n_elems = n_frames * FRAME;
    if (n_elems == 0) return cudaSuccess;
    const int64_t per_block = EQ_THREADS * EQ_UNROLL;
    if (dt == WIFI_F32) {
        const bool vec = ((((uintptr_t)rx) | ((uintptr_t)eq)) & 15) == 0;
        const int64_t n_vec = vec ? n_elems / 2 : 0;
        if (n_vec) {
            equalize_f32_kernel<<<(unsigned)((n_vec + per_block - 1) / per_block), EQ_THREADS, 0, s>>>(
                (const float4 *)rx, (const float2 *)Hlt, (const float2 *)Hps, (float4 *)eq, n_vec);
            ++g_last_launches;
        }
        if (2 * n_vec < n_elems) {
            const int64_t rest = n_elems - 2 * n_vec;
            equalize_f32_scalar_kernel<<<(unsigned)((rest + 255) / 256), 256, 0, s>>>((const float2 *)rx, (const float2 *)Hlt, (const float2 *)Hps,
                                                                                     (float2 *)eq, 2 * n_vec, n_elems);
            ++g_last_launches;
        }
    } else {
        const int64_t n_fk = n_frames * NSC;
        equalize_f64_kernel<<<(unsigned)((n_fk + EQ_THREADS - 1) / EQ_THREADS), EQ_THREADS, 0, s>>>(
            (const double2 *)rx, (const double2 *)Hlt, (const double2 *)Hps, (double2 *)eq, n_fk);
        ++g_last_launches;
    }
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// PS_MMSE in the reference's own calling convention (main.c:148 / WiFi_channel_estimation_PS_MMSE.m): R_f = H_ls H_ls^H
// ------------------------------------------------------------------------------------------
// With the rank-one covariance the intended estimator  R (R + ow2 (X X^H)^-1)^-1 (rx/tx)  has the closed form (Sherman-
// Morrison)
//     H = H_ls * (v^H rx) / (ow2 + v^H v),   v = tx (.) H_ls
// -- 212 complex values of traffic per frame instead of a 53 x 53 solve (the pivoted kernel: 11 M frames/s).
// MODE 1 is the MATLAB text as written (WiFi_channel_estimation_PS_MMSE.m:27-33: Rhy = Rhh F' X with X NOT conjugated,
// result averaged over OFDM blocks 1..4 of whole frames):  c_b = (v'^H w),  w = (v v^H + ow2 I)^-1 rx,  v' = conj(tx) (.) H_ls,
// evaluated with e = v' - v (exactly zero for real tx) so that the BPSK case carries no cancellation:
//     c_b = a3/(ow2 + vv) + [er (ow2 + vv) - ev a3] / (ow2 (ow2 + vv)),   a3 = v^H rx, vv = v^H v, er = e^H rx, ev = e^H v.
template <typename T, int MODE>
__global__ void __launch_bounds__(256) mmse_rank1_kernel(const cx<T> *__restrict__ tx, const cx<T> *__restrict__ rx, int64_t frame_stride,
                                                         const T *__restrict__ ow2, const cx<T> *__restrict__ Hls, cx<T> *__restrict__ H,
                                                         int64_t n_frames)
{
    const int lane = threadIdx.x & 31;
    const int64_t wstride = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); f < n_frames; f += wstride) {
        const bool second = lane + 32 < NSC;
        const cx<T> h0 = Hls[f * NSC + lane], h1 = second ? Hls[f * NSC + lane + 32] : mk<T>(0, 0);
        const T s2 = ow2[f];
        cx<T> c = mk<T>(0, 0);
#pragma unroll
        for (int b = 0; b < (MODE == 1 ? 4 : 1); ++b) {
            const cx<T> *tp = tx + f * frame_stride + b * NSC, *rp = rx + f * frame_stride + b * NSC;
            const cx<T> x0 = ld_stream(tp + lane), r0 = ld_stream(rp + lane);
            const cx<T> x1 = second ? ld_stream(tp + lane + 32) : mk<T>(0, 0), r1 = second ? ld_stream(rp + lane + 32) : mk<T>(0, 0);
            const cx<T> v0 = cmul(x0, h0), v1 = cmul(x1, h1);
            // a3 = v^H rx, vv = v^H v
            T a3x = v0.x * r0.x + v0.y * r0.y + v1.x * r1.x + v1.y * r1.y;
            T a3y = v0.x * r0.y - v0.y * r0.x + v1.x * r1.y - v1.y * r1.x;
            T vv = v0.x * v0.x + v0.y * v0.y + v1.x * v1.x + v1.y * v1.y;
            // MODE 1: e = v' - v = (conj(tx) - tx) (.) H_ls  (exactly zero for real tx);  er = e^H rx, ev = e^H v
            T erx = 0, ery = 0, evx = 0, evy = 0;
            if (MODE == 1) {
                const cx<T> e0 = mk<T>((T)2 * x0.y * h0.y, (T)-2 * x0.y * h0.x), e1 = mk<T>((T)2 * x1.y * h1.y, (T)-2 * x1.y * h1.x);
                erx = e0.x * r0.x + e0.y * r0.y + e1.x * r1.x + e1.y * r1.y;
                ery = e0.x * r0.y - e0.y * r0.x + e1.x * r1.y - e1.y * r1.x;
                evx = e0.x * v0.x + e0.y * v0.y + e1.x * v1.x + e1.y * v1.y;
                evy = e0.x * v0.y - e0.y * v0.x + e1.x * v1.y - e1.y * v1.x;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a3x += __shfl_xor_sync(0xffffffffu, a3x, o); a3y += __shfl_xor_sync(0xffffffffu, a3y, o);
                vv += __shfl_xor_sync(0xffffffffu, vv, o);
                if (MODE == 1) {
                    erx += __shfl_xor_sync(0xffffffffu, erx, o); ery += __shfl_xor_sync(0xffffffffu, ery, o);
                    evx += __shfl_xor_sync(0xffffffffu, evx, o); evy += __shfl_xor_sync(0xffffffffu, evy, o);
                }
            }
            const T den = s2 + vv;
            if (MODE == 0) {
                c = mk<T>(a3x / den, a3y / den);
            } else {
                // v'^H w = [a1 ow2 + (a1 vv - a2 a3)] / (ow2 (ow2 + vv)) with a1 = a3 + er, a2 = vv + ev:
                //        = a3 / (ow2 + vv) + [er (ow2 + vv) - ev a3] / (ow2 (ow2 + vv))       (second term: complex tx only)
                const T dx = erx * den - (evx * a3x - evy * a3y), dy = ery * den - (evx * a3y + evy * a3x);
                c.x += a3x / den + dx / (s2 * den);
                c.y += a3y / den + dy / (s2 * den);
            }
        }
        if (MODE == 1) c = mk<T>(c.x * (T)0.25, c.y * (T)0.25);
        st_stream(H + f * NSC + lane, cmul(h0, c));
        if (second) st_stream(H + f * NSC + lane + 32, cmul(h1, c));
    }
}

cudaError_t launch_mmse_rank1(wifi_dtype dt, int matlab, const void *tx, const void *rx, int64_t frame_stride, const void *ow2, const void *Hls,
                              void *H, int64_t n_frames, cudaStream_t s)
{
    g_last_launches = 0;
    if (n_frames == 0) return cudaSuccess;
    g_last_launches = 1;
    const unsigned grid = (unsigned)std::min<int64_t>((n_frames + 7) / 8, (int64_t)148 * 8);
#define R1(T, M) mmse_rank1_kernel<T, M><<<grid, 256, 0, s>>>((const cx<T> *)tx, (const cx<T> *)rx, frame_stride, (const T *)ow2, (const cx<T> *)Hls, (cx<T> *)H, n_frames)
    if (dt == WIFI_F32) { if (matlab) R1(float, 1); else R1(float, 0); }
    else { if (matlab) R1(double, 1); else R1(double, 0); }
#undef R1
    return cudaGetLastError();
}

}  // namespace wifi
