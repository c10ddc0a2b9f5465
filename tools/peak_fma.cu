// Micro-benchmark: measured FP32 FFMA, packed FFMA2 (fma.rn.f32x2) and INT32 IMAD issue peaks
// on this B200 -- the compute roofs DESIGN.md quotes next to the HBM roof.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peak_fma tools/peak_fma.cu
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b)
{
    float r[16];
    unsigned u[16];
#pragma unroll
    for (int i = 0; i < 16; i++) { r[i] = threadIdx.x * 0.001f + i; u[i] = threadIdx.x + i; }
    const unsigned ua = __float_as_uint(a), ub = __float_as_uint(b);
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            if (MODE == 0) r[i] = fmaf(r[i], a, b);
            if (MODE == 1) u[i] = u[i] * ua + ub;
        }
        if (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) {
                unsigned long long d, x, y, z;
                asm("mov.b64 %0, {%1, %2};" : "=l"(x) : "f"(r[i]), "f"(r[i + 1]));
                asm("mov.b64 %0, {%1, %2};" : "=l"(y) : "f"(a), "f"(a));
                asm("mov.b64 %0, {%1, %2};" : "=l"(z) : "f"(b), "f"(b));
                asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(x), "l"(y), "l"(z));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(r[i]), "=f"(r[i + 1]) : "l"(d));
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) s += r[i] + __uint_as_float(u[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
static void run(const char *name, int sms)
{
    float *out;
    const int grid = sms * 8, iters = 20000;
    cudaMalloc(&out, grid * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double ops = (double)grid * 256 * iters * 16;
    printf("{\"probe\": \"%s\", \"ms\": %.3f, \"Tmac_per_s\": %.2f, \"Tflop_per_s\": %.2f}\n", name, best, ops / best / 1e9,
           2 * ops / best / 1e9);
    cudaFree(out);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{\"sms\": %d}\n", sms);
    run<0>("ffma", sms);
    run<2>("ffma2_f32x2", sms);
    run<1>("imad", sms);
    return 0;
}
