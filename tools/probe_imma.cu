// Micro-benchmark: legacy warp-level int8 tensor MMA (mma.sync m16n8k32 s8 x s8 -> s32) on sm_100a.
// Informs the next step for the bit-exact int16 FIR (limb-split Toeplitz GEMM): is the legacy
// IMMA path fast enough, or does it have to be tcgen05 kind::i8?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_imma tools/probe_imma.cu
#include <cuda_runtime.h>
#include <cstdio>

__global__ void __launch_bounds__(256) k(int *out, int iters, unsigned a0, unsigned b0)
{
    int c[8][4];
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + i + j;
    unsigned a[4] = {a0 + threadIdx.x, a0 * 3 + threadIdx.x, a0 * 5, a0 * 7}, b[2] = {b0 + threadIdx.x, b0 * 3};
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+r"(c[i][0]), "+r"(c[i][1]), "+r"(c[i][2]), "+r"(c[i][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
    }
    int s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int *out;
    const int grid = sms * 8, iters = 4000;
    cudaMalloc(&out, grid * 256 * sizeof(int));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<grid, 256>>>(out, iters, 0x01020304u, 0x05060708u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<<<grid, 256>>>(out, iters, 0x01020304u, 0x05060708u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // one mma = 16*8*32 MACs per warp
    const double macs = (double)grid * 8 /*warps*/ * iters * 8 * (16.0 * 8 * 32);
    printf("{\"probe\": \"mma.sync m16n8k32 s8\", \"ms\": %.3f, \"Tmac_per_s\": %.1f, \"Tops\": %.1f, \"err\": \"%s\"}\n", best,
           macs / best / 1e9, 2 * macs / best / 1e9, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
