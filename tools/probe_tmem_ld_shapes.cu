// Which (lane, column) of tensor memory does each register of  tcgen05.ld.16x256b.x2  hold?
// Every TMEM lane L writes value L * 256 + c into column c (32x32b store: thread = lane); warp 1 then reads
// 16 columns with the 16-lane shape at lane offsets 32 and 48 of its quadrant and prints one line per thread.
// Expected (the mma.m16n8 accumulator fragment): reg 4 g + 2 h + e of thread t = lane base + t / 4 + 8 h,
// column 8 g + 2 (t % 4) + e.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/probe_tmem_ld tools/probe_tmem_ld_shapes.cu && /tmp/probe_tmem_ld
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(unsigned *out)
{
    __shared__ unsigned tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tbase;
    for (int c0 = 0; c0 < 16; c0 += 8) {
        unsigned w[8];
        for (int c = 0; c < 8; c++) w[c] = tid * 256 + c0 + c;
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tb + ((unsigned)(warp * 32) << 16) + c0),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 1)
        for (int half = 0; half < 2; half++) {
            unsigned v[8];
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                         : "r"(tb + ((unsigned)(32 + 16 * half) << 16))
                         : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int k = 0; k < 8; k++) out[(half * 32 + (tid & 31)) * 8 + k] = v[k];
        }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb) : "memory");
}

int main()
{
    unsigned *d, h[2 * 32 * 8];
    cudaMalloc(&d, sizeof(h));
    cudaMemset(d, 0xff, sizeof(h));
    probe<<<1, 128>>>(d);
    if (cudaDeviceSynchronize() != cudaSuccess) { std::printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int half = 0; half < 2; half++)
        for (int t = 0; t < 32; t++) {
            std::printf("lane base %d thread %2d:", 32 + 16 * half, t);
            for (int k = 0; k < 8; k++) {
                const unsigned v = h[(half * 32 + t) * 8 + k];
                std::printf(" (%u,%u)", v >> 8, v & 255);
                const int g = k >> 2, hh = (k >> 1) & 1, e = k & 1;
                if ((int)(v >> 8) != 32 + 16 * half + t / 4 + 8 * hh || (int)(v & 255) != 8 * g + 2 * (t % 4) + e) bad++;
            }
            std::printf("\n");
        }
    std::printf("registers that differ from the m16n8 accumulator-fragment layout: %d of 512\n", bad);
    return 0;
}
