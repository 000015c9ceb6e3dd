// wifi_peaks.cu -- on-box ceilings the rooflines are quoted against ("the bench must measure them on-box"):
// FP32 FFMA, FP64 DFMA, FP64 DMMA (mma.sync m8n8k4) throughput and a streaming copy.  Each kernel is a register-only
// loop of independent accumulator chains in the 4 x 4 outer-product form the elimination update has (operands reused),
// timed with CUDA events on the context's stream.
#include "wifi_common.cuh"
#include "wifi_internal.h"

namespace wifi {

template <typename T> __global__ void __launch_bounds__(256) peak_fma_kernel(T *out, int iters, T seed)
{
    T acc[16], x[4], y[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = seed * (T)i;
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = seed + (T)(i + threadIdx.x); y[i] = seed - (T)i; }
#pragma unroll 8
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = fma(x[i >> 2], y[i & 3], acc[i]);
    T s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) peak_dmma_kernel(double *out, int iters, double seed)
{
    double c[8][2], a = seed + threadIdx.x, b = seed - threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = seed * i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) peak_copy_kernel(const float4 *__restrict__ src, float4 *__restrict__ dst, int64_t n)
{
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        st_stream(dst + i, ld_stream(src + i));
}

// which: 0 FP32 FMA TFLOP/s, 1 FP64 FMA TFLOP/s, 2 FP64 DMMA TFLOP/s, 3 streaming copy GB/s (read + write)
cudaError_t measure_peak(int which, double *value, cudaStream_t s)
{
    int sms = 148, dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaEvent_t e0, e1;
    cudaError_t err;
    if ((err = cudaEventCreate(&e0)) != cudaSuccess) return err;
    if ((err = cudaEventCreate(&e1)) != cudaSuccess) return err;
    const int grid = sms * 4;
    void *buf = nullptr;
    const size_t copy_bytes = (size_t)1 << 30;
    if ((err = cudaMalloc(&buf, which == 3 ? 2 * copy_bytes : (size_t)grid * 256 * 8)) != cudaSuccess) return err;
    float best = 1e30f;
    double work = 0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0, s);
        if (which == 0) { peak_fma_kernel<float><<<grid, 256, 0, s>>>((float *)buf, 1 << 15, 1e-3f); work = 2.0 * grid * 256 * 16 * (1 << 15); }
        else if (which == 1) { peak_fma_kernel<double><<<grid, 256, 0, s>>>((double *)buf, 1 << 14, 1e-3); work = 2.0 * grid * 256 * 16 * (1 << 14); }
        else if (which == 2) { peak_dmma_kernel<<<grid, 256, 0, s>>>((double *)buf, 1 << 12, 1e-3); work = 512.0 * grid * 8 * 8 * (1 << 12); }
        else { peak_copy_kernel<<<sms * 16, 256, 0, s>>>((const float4 *)buf, (float4 *)((char *)buf + copy_bytes), (int64_t)(copy_bytes / 16)); work = 2.0 * copy_bytes; }
        cudaEventRecord(e1, s);
        if ((err = cudaEventSynchronize(e1)) != cudaSuccess) break;
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaFree(buf);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (err != cudaSuccess) return err;
    *value = which == 3 ? work / (best * 1e-3) / 1e9 : work / (best * 1e-3) / 1e12;
    return cudaGetLastError();
}

}  // namespace wifi
