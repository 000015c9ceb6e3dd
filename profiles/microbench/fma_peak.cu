// fma_peak.cu -- measured FP32 / FP32x2 / FP64 FMA throughput of the GPU this runs on (the FLOP ceilings the per-frame
// MMSE solve is reported against; SURVEY 8(d) asks for an on-box FMA microbenchmark next to the spec values).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_peak fma_peak.cu && ./fma_peak
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE> __global__ void __launch_bounds__(256) fma_kernel(float *out, int iters, float seed)
{
    // 16 independent chains per thread; operands are distinct loop-invariant registers (the 3-register FFMA form)
    if (MODE == 0) {
        float acc[16], x[16], y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc[i] = seed * i; x[i] = seed + i + threadIdx.x; y[i] = seed - i; }
#pragma unroll 8
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(x[i], y[i], acc[i]);
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 1) {
        float2 acc[8], x[8], y[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i] = make_float2(seed * i, seed); x[i] = make_float2(seed + i + threadIdx.x, seed - i); y[i] = make_float2(seed - i, seed + 2 * i); }
#pragma unroll 8
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] = __ffma2_rn(x[i], y[i], acc[i]);
        float s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 3) {
        // 4 x 4 outer-product form (what the elimination update looks like): acc[i][j] += x[i] * y[j]
        float acc[16], x[4], y[4];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = seed * i;
#pragma unroll
        for (int i = 0; i < 4; ++i) { x[i] = seed + i + threadIdx.x; y[i] = seed - i; }
#pragma unroll 8
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(x[i >> 2], y[i & 3], acc[i]);
        float s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (MODE == 4) {
        double acc[16], x[4], y[4];
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = seed * i;
#pragma unroll
        for (int i = 0; i < 4; ++i) { x[i] = seed + i + threadIdx.x; y[i] = seed - i; }
#pragma unroll 8
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(x[i >> 2], y[i & 3], acc[i]);
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
    } else {
        double acc[16], x[16], y[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { acc[i] = seed * i; x[i] = seed + i + threadIdx.x; y[i] = seed - i; }
#pragma unroll 8
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fma(x[i], y[i], acc[i]);
        double s = 0;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += acc[i];
        out[blockIdx.x * blockDim.x + threadIdx.x] = (float)s;
    }
}

template <int MODE> static double run(const char *name, int blocks_per_sm, int iters, float *out, int sms, double clock_ghz)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * blocks_per_sm;
    fma_kernel<MODE><<<grid, 256>>>(out, iters, 1e-3f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fma_kernel<MODE><<<grid, 256>>>(out, iters, 1e-3f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double fmas = (double)grid * 256 * 16 * iters;      // scalar FMAs (an f32x2 instruction counts as two)
    double tflops = 2 * fmas / (best * 1e-3) / 1e12;
    printf("{\"kernel\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"TFLOPs\": %.2f, \"fma_per_clk_per_sm_at_%.3fGHz\": %.1f}\n", name, blocks_per_sm, best, tflops,
           clock_ghz, fmas / (best * 1e-3) / sms / (clock_ghz * 1e9));
    return tflops;
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    double ghz = khz * 1e-6;
    float *out;
    cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_ghz\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
    for (int b : {2, 4, 8}) {
        run<0>("ffma (16 chains, all operands distinct)", b, 1 << 16, out, p.multiProcessorCount, ghz);
        run<3>("ffma (4x4 outer product, operands reused)", b, 1 << 16, out, p.multiProcessorCount, ghz);
        run<1>("ffma2 (fma.rn.f32x2)", b, 1 << 16, out, p.multiProcessorCount, ghz);
        run<2>("dfma (16 chains, all operands distinct)", b, 1 << 14, out, p.multiProcessorCount, ghz);
        run<4>("dfma (4x4 outer product, operands reused)", b, 1 << 14, out, p.multiProcessorCount, ghz);
    }
    return 0;
}
