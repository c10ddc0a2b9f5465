// How long does one tcgen05.mma.kind::i8 (M = 128, K = 32) take, by operand placement, N, and accumulator rotation?
// One thread issues `iters` MMAs back to back, commits, waits; cycles per MMA = (t1 - t0) / iters.
//   A: shared memory (32-byte-swizzled rows, the Hankel data planes of fir_umma32_kernel; or canonical no-swizzle tiles)
//      or tensor memory (fir_umma32t_kernel's tap tiles);
//   B: shared memory, 32-byte-swizzled rows or canonical tiles;  N = 64 .. 256;
//   nacc accumulators used round-robin (1 = every MMA depends on the previous one).
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_umma_rate tools/probe_umma_rate.cu && /tmp/probe_umma_rate
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long desc(unsigned addr, unsigned lbo, unsigned sbo, unsigned layout)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((addr >> 4) & 0x3FFF);
    d |= (unsigned long long)((lbo >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;
    d |= (unsigned long long)(layout & 7) << 61;
    return d;
}

struct Mode { int a_tmem, a_sw32, b_sw32, N, nacc, step_rows; };

template <int NACC, bool ATMEM>
__global__ void probe(Mode m, int iters, long long *out)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    __shared__ unsigned tbase;
    __shared__ __align__(8) unsigned long long bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 72 * 1024 / 4; i += blockDim.x) reinterpret_cast<unsigned *>(sm)[i] = 0x01010101u;
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tbase;
    {   // A tile in tensor memory: columns 504..511
        unsigned w = 0x01010101u;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(tb + ((unsigned)((tid >> 5) * 32) << 16) + 504), "r"(w) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        const unsigned idesc = (2u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(m.N >> 3) << 17) | ((128u >> 4) << 24);
        // a_sw32 == 2: the 16-byte-pitch rows of fir_umma_kernel / fir_ummap_kernel (no swizzle, LBO 16, SBO 128: row m = plane[16 m ..])
        const unsigned long long ad0 = m.a_sw32 == 2 ? desc(smem_u32(sm), 16, 128, 0) : m.a_sw32 ? desc(smem_u32(sm), 16, 256, 6) : desc(smem_u32(sm), 128, 256, 0);
        const unsigned long long bd0 = m.b_sw32 ? desc(smem_u32(sm + 24576), 16, 256, 6) : desc(smem_u32(sm + 24576), 128, 256, 0);
        // everything per MMA is a compile-time choice among registers set up here (the kernels do the same)
        unsigned dcol[NACC];
        unsigned long long ad[5], bd[5];
#pragma unroll
        for (int k = 0; k < NACC; k++) dcol[k] = tb + (unsigned)(k * m.N);
#pragma unroll
        for (int k = 0; k < 5; k++) {
            // operands advance like the kernels' k-blocks: one 32-byte row (swizzled) or one tile (canonical) per step
            ad[k] = ad0 + (m.a_sw32 ? k * 2ull : k * (unsigned long long)((128 * 32) >> 4));
            bd[k] = bd0 + (m.b_sw32 ? k * 2ull : k * ((unsigned long long)(m.N * 32) >> 4));
        }
        const long long t0 = clock64();
        for (int i = 0; i < iters; i += 20) {
#pragma unroll
            for (int k = 0; k < 20; k++) {
                if (ATMEM)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(dcol[k % NACC]), "r"(tb + 504), "l"(bd[k % 5]), "r"(idesc) : "memory");
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.eq.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol[k % NACC]), "l"(ad[k % 5]), "l"(bd[k % 5]), "r"(idesc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
        out[0] = clock64() - t0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tb) : "memory");
}

int main()
{
    long long *d, h;
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(probe<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(probe<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(probe<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(probe<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(probe<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    cudaFuncSetAttribute(probe<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 72 * 1024);
    const int iters = 2000;
    const Mode modes[] = {
        // a_tmem a_sw32 b_sw32 N nacc
        {0, 1, 0, 128, 1, 0}, {0, 1, 0, 128, 2, 0}, {0, 1, 0, 64, 1, 0}, {0, 1, 0, 64, 4, 0}, {0, 1, 0, 256, 1, 0}, {0, 1, 0, 256, 2, 0},
        {0, 0, 0, 128, 1, 0}, {0, 0, 0, 128, 2, 0}, {0, 0, 0, 256, 1, 0}, {0, 0, 0, 256, 2, 0}, {0, 0, 0, 64, 1, 0}, {0, 0, 0, 64, 4, 0},
        {1, 0, 1, 96, 1, 0},  {1, 0, 1, 96, 2, 0},  {1, 0, 1, 96, 4, 0},  {1, 0, 1, 128, 1, 0}, {1, 0, 1, 128, 2, 0}, {1, 0, 1, 192, 1, 0},
        {1, 0, 1, 192, 2, 0}, {1, 0, 1, 240, 2, 0}, {1, 0, 1, 64, 1, 0},  {1, 0, 1, 64, 4, 0},  {1, 0, 1, 32, 1, 0},  {1, 0, 1, 32, 4, 0},
        {1, 0, 0, 96, 2, 0},  {1, 0, 0, 128, 2, 0}, {1, 0, 0, 240, 2, 0}, {1, 0, 0, 64, 4, 0},
        {0, 2, 0, 64, 1, 0},  {0, 2, 0, 96, 2, 0},  {0, 2, 0, 128, 2, 0}, {0, 2, 0, 192, 2, 0},
    };
    std::printf("A operand        B operand        N  accumulators  cycles/MMA  MAC/cycle\n");
    for (const Mode &m : modes) {
        auto run = [&] {
            if (m.a_tmem) {
                if (m.nacc == 1) probe<1, true><<<1, 128, 72 * 1024>>>(m, iters, d);
                else if (m.nacc == 2) probe<2, true><<<1, 128, 72 * 1024>>>(m, iters, d);
                else probe<4, true><<<1, 128, 72 * 1024>>>(m, iters, d);
            } else {
                if (m.nacc == 1) probe<1, false><<<1, 128, 72 * 1024>>>(m, iters, d);
                else if (m.nacc == 2) probe<2, false><<<1, 128, 72 * 1024>>>(m, iters, d);
                else probe<4, false><<<1, 128, 72 * 1024>>>(m, iters, d);
            }
        };
        run();
        if (cudaDeviceSynchronize() != cudaSuccess) { std::printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        run();
        cudaDeviceSynchronize();
        cudaMemcpy(&h, d, sizeof(h), cudaMemcpyDeviceToHost);
        const double c = (double)h / iters;
        std::printf("%-16s %-16s %3d  %5d  %10.1f  %9.0f\n", m.a_tmem ? "tensor memory" : m.a_sw32 == 2 ? "smem 16B-pitch" : m.a_sw32 ? "smem 32B-swizzle" : "smem canonical",
                    m.b_sw32 ? "smem 32B-swizzle" : "smem canonical", m.N, m.nacc, c, 128.0 * m.N * 32 / c);
    }
    return 0;
}
