// wifi_common.cuh -- shared device helpers of the sm_100a kernels (complex arithmetic on
// interleaved float2/double2, streaming vector loads/stores, frame constants).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/wifi_b200.h"

#define NSC WIFI_NSC
#define NBLK WIFI_NBLK
#define FRAME WIFI_FRAME
#define DCBIN WIFI_DC

namespace wifi {

template <typename T> struct V2;
template <> struct V2<float> { using type = float2; };
template <> struct V2<double> { using type = double2; };
template <typename T> using cx = typename V2<T>::type;

template <typename T> __host__ __device__ __forceinline__ cx<T> mk(T re, T im) { cx<T> r; r.x = re; r.y = im; return r; }
template <typename C> __device__ __forceinline__ C cadd(C a, C b) { a.x += b.x; a.y += b.y; return a; }
template <typename C> __device__ __forceinline__ C csub(C a, C b) { a.x -= b.x; a.y -= b.y; return a; }
template <typename C> __device__ __forceinline__ C cmul(C a, C b) { C r; r.x = a.x * b.x - a.y * b.y; r.y = a.x * b.y + a.y * b.x; return r; }
template <typename C> __device__ __forceinline__ C cconj(C a) { a.y = -a.y; return a; }
// acc += a * b
template <typename C> __device__ __forceinline__ void cfma(C &acc, C a, C b)
{
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
}
// acc -= a * b
template <typename C> __device__ __forceinline__ void cfms(C &acc, C a, C b)
{
    acc.x = fma(-a.x, b.x, acc.x); acc.x = fma(a.y, b.y, acc.x);
    acc.y = fma(-a.x, b.y, acc.y); acc.y = fma(-a.y, b.x, acc.y);
}
template <typename C> __device__ __forceinline__ auto cabs2(C a) -> decltype(a.x) { return a.x * a.x + a.y * a.y; }
// a / b, textbook formula (what rx/tx of main.c:83 computes, up to rounding); 0/0 -> NaN like the reference
template <typename C> __device__ __forceinline__ C cdiv(C a, C b)
{
    auto den = b.x * b.x + b.y * b.y;
    C r;
    r.x = (a.x * b.x + a.y * b.y) / den;
    r.y = (a.y * b.x - a.x * b.y) / den;
    return r;
}
template <typename C> __device__ __forceinline__ C crecip(C b)
{
    auto den = b.x * b.x + b.y * b.y;
    C r; r.x = b.x / den; r.y = -b.y / den; return r;
}
template <typename C, typename S> __device__ __forceinline__ C cscale(C a, S s) { a.x *= s; a.y *= s; return a; }

// streaming (evict-first) global accesses: every LS/interp byte is touched exactly once
template <typename V> __device__ __forceinline__ V ld_stream(const V *p) { return __ldcs(p); }
template <typename V> __device__ __forceinline__ void st_stream(V *p, V v) { __stcs(p, v); }

// isolated gather load (pilot sub-carriers): streaming, and ask L2 to fetch only the 64-byte half-line that holds the
// value -- by default every touched sector promotes to a full 128-byte line from HBM (ncu: 929 B/frame for 8 pilots)
__device__ __forceinline__ float2 ld_gather(const float2 *p)
{
    float2 v;
    asm volatile("ld.global.cs.L2::64B.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
__device__ __forceinline__ double2 ld_gather(const double2 *p)
{
    double2 v;
    asm volatile("ld.global.cs.L2::64B.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
    return v;
}

// Packed FP32 pairs (sm_100a: FFMA2 / FADD2 -- one issue slot for two operations; the FMA *rate* is not higher than FFMA's, so this
// pays where a kernel is issue- or latency-bound, not FMA-pipe-bound).  A pair lives in a 64-bit register: a float2 read as one word.
using f32x2 = unsigned long long;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ void unpack2(f32x2 v, uint32_t &lo, uint32_t &hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 ffma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 fsub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// main.c:69-72 as written: c = Re(tx) - Im(tx) (a real scalar), H = (c*rx)/(c*tx).  Re(tx) == Im(tx)
// gives 0/0 = NaN exactly like the reference.
template <typename T> __device__ __forceinline__ cx<T> lt_ls_one(cx<T> tx, cx<T> rx)
{
    T c = tx.x - tx.y;
    return cdiv(mk<T>(c * rx.x, c * rx.y), mk<T>(c * tx.x, c * tx.y));
}

// r / h for the FP32 equalizer with one reciprocal (MUFU.RCP, 1 ulp, branch-free); the IEEE '/' of cdiv() is ~12 instructions
// and a slow-path branch per real divide.  0/0 still yields NaN.
__device__ __forceinline__ float2 eq_div(float2 a, float2 b)
{
    const float den = fmaf(b.x, b.x, b.y * b.y);
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(den));
    return make_float2(fmaf(a.x, b.x, a.y * b.y) * r, fmaf(a.y, b.x, -a.x * b.y) * r);
}
__device__ __forceinline__ double2 eq_div(double2 a, double2 b) { return cdiv(a, b); }

}  // namespace wifi
