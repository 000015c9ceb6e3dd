// dmma_peak.cu -- FP64 tensor-core (mma.sync ... f64) throughput on this GPU, per instruction shape.
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE> __global__ void __launch_bounds__(256) dmma_kernel(double *out, int iters, double seed)
{
    // 8 independent accumulator tiles per warp
    if (SHAPE == 0) {           // m8n8k4: A 1 reg, B 1 reg, C 2 regs
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
    } else if (SHAPE == 1) {    // m16n8k4: A 2 regs, B 1 reg, C 4 regs
        double c[8][4], a0 = seed + threadIdx.x, a1 = seed * 3, b = seed - threadIdx.x;
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = seed * i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(b));
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else if (SHAPE == 2) {    // m16n8k8: A 4 regs, B 2 regs, C 4 regs
        double c[8][4], a0 = seed + threadIdx.x, a1 = seed * 3, a2 = seed * 5, a3 = seed * 7, b0 = seed - threadIdx.x, b1 = seed * 2;
#pragma unroll
        for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = seed * i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3]) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {                    // m16n8k16: A 8 regs, B 4 regs, C 4 regs
        double c[8][4], a[8], b[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) { c[i][0] = c[i][1] = c[i][2] = c[i][3] = seed * i; a[i] = seed + i + threadIdx.x; }
#pragma unroll
        for (int i = 0; i < 4; ++i) b[i] = seed - i;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int i = 0; i < 8; ++i)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]), "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        double s = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

template <int SHAPE> static void run(const char *name, double flops_per_mma, int blocks_per_sm, int iters, double *out, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * blocks_per_sm;
    dmma_kernel<SHAPE><<<grid, 256>>>(out, iters, 1e-3);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        dmma_kernel<SHAPE><<<grid, 256>>>(out, iters, 1e-3);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double mmas = (double)grid * 8 /*warps*/ * 8 * iters;
    printf("{\"kernel\": \"%s\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"TFLOPs\": %.2f}\n", name, blocks_per_sm, best, mmas * flops_per_mma / (best * 1e-3) / 1e12);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    double *out;
    cudaMalloc(&out, sizeof(double) * p.multiProcessorCount * 8 * 256);
    for (int b : {1, 2, 4}) {
        run<0>("dmma m8n8k4", 2.0 * 8 * 8 * 4, b, 4096, out, p.multiProcessorCount);
        run<1>("dmma m16n8k4", 2.0 * 16 * 8 * 4, b, 4096, out, p.multiProcessorCount);
        run<2>("dmma m16n8k8", 2.0 * 16 * 8 * 8, b, 2048, out, p.multiProcessorCount);
        run<3>("dmma m16n8k16", 2.0 * 16 * 8 * 16, b, 1024, out, p.multiProcessorCount);
    }
    return 0;
}
