// Micro-benchmark: does packed fp32x2 arithmetic (FFMA2/FADD2/FMUL2) free issue slots on sm_100a?
// The fused overlap-save FIR kernel is issue/L1-bound with the FMA pipe at 40 %: if a packed
// instruction occupies one issue slot for two lanes' worth of FMA-pipe work, the other slot can
// feed the ALU / LSU pipes.  Modes mix FP work with independent integer (ALU-pipe) or LDS work.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe_issue tools/probe_issue.cu
#include <cuda_runtime.h>
#include <cstdio>

__device__ __forceinline__ unsigned long long pk(float a, float b)
{
    unsigned long long r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
    return r;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b)
{
    unsigned long long d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// FP: 0 = 16 scalar FFMA, 1 = 8 FFMA2, 2 = 16 scalar FADD, 3 = 8 FADD2, 4 = 8 FMUL2, 5 = none
// SIDE: 0 = none, 1 = 8 integer ALU ops (LOP3), 2 = 4 LDS.64, 3 = 16 integer ALU ops
template <int FP, int SIDE>
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b, unsigned key)
{
    __shared__ float2 sm[1024];
    float r[16];
    unsigned long long p[8];
    unsigned u[16];
    float2 acc = make_float2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 16; i++) { r[i] = threadIdx.x * 0.001f + i; u[i] = threadIdx.x + i; }
#pragma unroll
    for (int i = 0; i < 8; i++) p[i] = pk(r[2 * i], r[2 * i + 1]);
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = make_float2(i, -i);
    __syncthreads();
    const unsigned long long pa = pk(a, a), pb = pk(b, b);
    int idx = threadIdx.x;
    for (int it = 0; it < iters; it++) {
        if (FP == 0) {
#pragma unroll
            for (int i = 0; i < 16; i++) r[i] = fmaf(r[i], a, b);
        } else if (FP == 1) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = fma2(p[i], pa, pb);
        } else if (FP == 2) {
#pragma unroll
            for (int i = 0; i < 16; i++) r[i] = r[i] + b;
        } else if (FP == 3) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = add2(p[i], pb);
        } else if (FP == 4) {
#pragma unroll
            for (int i = 0; i < 8; i++) p[i] = mul2(p[i], pa);
        }
        if (SIDE == 1 || SIDE == 3) {
#pragma unroll
            for (int i = 0; i < (SIDE == 1 ? 8 : 16); i++) u[i] = (u[i] ^ key) & (u[i] | 0x5a5a5a5au);   // one LOP3
        } else if (SIDE == 2) {
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float2 v = sm[(idx + 256 * i) & 1023];
                acc.x += v.x; acc.y += v.y;      // 2 FADD per LDS keep the loads alive
            }
            idx = (idx + 1) & 255;
        }
    }
    float s = acc.x + acc.y;
#pragma unroll
    for (int i = 0; i < 16; i++) s += r[i] + __uint_as_float(u[i]);
#pragma unroll
    for (int i = 0; i < 8; i++) s += __uint_as_float((unsigned)(p[i] >> 32)) + __uint_as_float((unsigned)p[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int FP, int SIDE>
static void run(const char *name, int sms)
{
    float *out;
    const int grid = sms * 8, iters = 20000;
    cudaMalloc(&out, grid * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<FP, SIDE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f, 0x1234567u);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(e0);
        k<FP, SIDE><<<grid, 256>>>(out, iters, 1.0001f, 0.5f, 0x1234567u);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    // cycles per loop iteration per SM sub-partition at 1.965 GHz: 8 CTAs * 8 warps / 4 SMSPs = 16 warps each
    const double cyc = best * 1e-3 * 1.965e9 / iters / 16.0;
    printf("{\"probe\": \"%s\", \"ms\": %.3f, \"cycles_per_warp_iter_at_1965MHz\": %.2f}\n", name, best, cyc);
    cudaFree(out);
}

int main()
{
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    printf("{\"sms\": %d}\n", sms);
    run<0, 0>("16 FFMA", sms);
    run<1, 0>("8 FFMA2", sms);
    run<2, 0>("16 FADD", sms);
    run<3, 0>("8 FADD2", sms);
    run<4, 0>("8 FMUL2", sms);
    run<5, 1>("8 LOP3", sms);
    run<5, 3>("16 LOP3", sms);
    run<0, 1>("16 FFMA + 8 LOP3", sms);
    run<1, 1>("8 FFMA2 + 8 LOP3", sms);
    run<0, 3>("16 FFMA + 16 LOP3", sms);
    run<1, 3>("8 FFMA2 + 16 LOP3", sms);
    run<3, 3>("8 FADD2 + 16 LOP3", sms);
    run<5, 2>("4 LDS.64 + 8 FADD", sms);
    run<0, 2>("16 FFMA + 4 LDS.64 + 8 FADD", sms);
    run<1, 2>("8 FFMA2 + 4 LDS.64 + 8 FADD", sms);
    return 0;
}
