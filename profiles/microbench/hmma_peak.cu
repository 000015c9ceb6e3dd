// hmma_peak.cu -- throughput of the warp-level (legacy) tensor path, mma.sync kind tf32 / bf16, on this GPU: the ceiling of a
// register-resident 3xTF32 elimination (batched inverse / per-frame solve), where tcgen05's 128-row tiles do not fit.
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC> __global__ void __launch_bounds__(256) hmma_kernel(float *out, int iters, float seed)
{
    float c[NACC][4];
    unsigned a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(seed + i + threadIdx.x);
    b[0] = __float_as_uint(seed - threadIdx.x); b[1] = __float_as_uint(seed * 2);
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = seed * i;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (SHAPE == 0)        // tf32 m16n8k8
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else if (SHAPE == 1)   // tf32 m16n8k4
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(b[0]));
            else                   // bf16 m16n8k16
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int SHAPE, int NACC> static void run(const char *name, double flops_per_mma, int blocks_per_sm, int iters, float *out, int sms)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    int grid = sms * blocks_per_sm;
    hmma_kernel<SHAPE, NACC><<<grid, 256>>>(out, iters, 1e-3f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        hmma_kernel<SHAPE, NACC><<<grid, 256>>>(out, iters, 1e-3f);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double mmas = (double)grid * 8 /*warps*/ * NACC * iters;
    printf("{\"bench\": \"hmma_peak\", \"shape\": \"%s\", \"independent_acc\": %d, \"blocks_per_sm\": %d, \"ms\": %.4f, \"tflops\": %.2f, \"cycles_per_mma_per_scheduler\": %.2f}\n",
           name, NACC, blocks_per_sm, best, mmas * flops_per_mma / (best * 1e-3) / 1e12,
           (best * 1e-3 * 1.965e9) / (mmas / (sms * 4.0)));
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    float *out; cudaMalloc(&out, sizeof(float) * p.multiProcessorCount * 8 * 256);
    const int it = 4096;
    run<0, 8>("tf32.m16n8k8", 2.0 * 16 * 8 * 8, 2, it, out, p.multiProcessorCount);
    run<0, 8>("tf32.m16n8k8", 2.0 * 16 * 8 * 8, 1, it, out, p.multiProcessorCount);
    run<0, 2>("tf32.m16n8k8", 2.0 * 16 * 8 * 8, 1, it, out, p.multiProcessorCount);
    run<0, 1>("tf32.m16n8k8", 2.0 * 16 * 8 * 8, 1, it, out, p.multiProcessorCount);
    run<1, 8>("tf32.m16n8k4", 2.0 * 16 * 8 * 4, 2, it, out, p.multiProcessorCount);
    run<1, 1>("tf32.m16n8k4", 2.0 * 16 * 8 * 4, 1, it, out, p.multiProcessorCount);
    run<2, 8>("bf16.m16n8k16", 2.0 * 16 * 8 * 16, 2, it, out, p.multiProcessorCount);
    run<2, 1>("bf16.m16n8k16", 2.0 * 16 * 8 * 16, 1, it, out, p.multiProcessorCount);
    cudaFree(out);
    return 0;
}
