// latency.cu -- dependent-issue latencies (cycles) that set the per-step critical path of the per-frame solve.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void lat_kernel(long long *out, double xd, float xf, int iters)
{
    __shared__ double2 sm[256];
    long long t0, t1;
    sm[threadIdx.x] = make_double2(xd + threadIdx.x, xd);
    __syncthreads();
    // DFMA chain
    double a = xd;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) a = fma(a, xd, xd);
    t1 = clock64();
    if (threadIdx.x == 0) out[0] = t1 - t0;
    // FFMA chain
    float b = xf;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) b = fmaf(b, xf, xf);
    t1 = clock64();
    if (threadIdx.x == 0) out[1] = t1 - t0;
    // LDS.128 pointer chase
    int idx = threadIdx.x;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) { double2 v = sm[idx]; idx = ((int)v.y + idx) & 255; }
    t1 = clock64();
    if (threadIdx.x == 0) out[2] = t1 - t0;
    // bar.sync with 64 threads
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) asm volatile("bar.sync 1, 64;" ::: "memory");
    t1 = clock64();
    if (threadIdx.x == 0) out[3] = t1 - t0;
    // rcp.approx.ftz.f64 + dependent DFMA
    double r = xd;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) { double q; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(q) : "d"(r)); r = q; }
    t1 = clock64();
    if (threadIdx.x == 0) out[4] = t1 - t0;
    // shfl chain
    float s = xf;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < iters; ++i) s = __shfl_sync(0xffffffffu, s, (threadIdx.x + 1) & 31);
    t1 = clock64();
    if (threadIdx.x == 0) out[5] = t1 - t0;
    // STS -> syncwarp -> LDS round trip
    t0 = clock64();
    float w = xf;
    float *fs = (float *)sm;
#pragma unroll 16
    for (int i = 0; i < iters; ++i) { fs[threadIdx.x] = w; __syncwarp(); w = fs[(threadIdx.x + 1) & 31]; __syncwarp(); }
    t1 = clock64();
    if (threadIdx.x == 0) out[6] = t1 - t0;
    if (a + b + idx + r + s + w == 12345.678) out[7] = 1;
}

int main()
{
    long long *d, h[8];
    cudaMalloc(&d, 64);
    const int iters = 4096;
    for (int rep = 0; rep < 2; ++rep) {
        lat_kernel<<<1, 64>>>(d, 1.0000001, 1.0001f, iters);
        cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    }
    const char *names[] = {"dfma dependent", "ffma dependent", "lds.128 dependent", "bar.sync 64 threads (2 warps)", "mufu.rcp64h dependent",
                           "shfl dependent", "sts->syncwarp->lds->syncwarp"};
    for (int i = 0; i < 7; ++i) printf("{\"op\": \"%s\", \"cycles\": %.1f}\n", names[i], (double)h[i] / iters);
    return 0;
}
