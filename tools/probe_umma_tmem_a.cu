// Probe for the operand-swap plan of DESIGN.md section 7: how must an int8 A operand be laid out in
// TENSOR MEMORY for  tcgen05.mma.kind::i8 [d], [a_tmem], b_desc ...  ?
// Each thread (TMEM lane m) writes 32 bytes with tcgen05.st.32x32b.x8 (8 consecutive 32-bit columns),
// byte j of lane m = pattern(m, j); B is a 32 x 32 identity in shared memory, so D[m][n] is the byte the
// tensor core used as A[m][k = n].  The host reports whether A[m][k] = byte k of lane m (natural packing).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_tmem_a probe_umma_tmem_a.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__host__ __device__ inline unsigned pat(int m, int j) { return (unsigned)((m * 5 + j * 3 + 1) & 0x7f); }

__global__ void probe(unsigned *out)
{
    __shared__ __align__(1024) unsigned char bt[32 * 32];
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tbase;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32 * 32; i += blockDim.x) {
        const int n = i / 32, k = i % 32;
        bt[(n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + (k % 16)] = (n == k) ? 1 : 0;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tbase, a_col = 32;      // D: columns 0..31, A: columns 32..39
    {
        unsigned w[8];
        for (int c = 0; c < 8; c++) w[c] = pat(tid, 4 * c) | (pat(tid, 4 * c + 1) << 8) | (pat(tid, 4 * c + 2) << 16) | (pat(tid, 4 * c + 3) << 24);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(tb + ((unsigned)(warp * 32) << 16) + a_col),
                     "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
                     : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        unsigned long long bd = 0;
        const unsigned bs = smem_u32(bt);
        bd |= (unsigned long long)((bs >> 4) & 0x3FFF);
        bd |= (unsigned long long)((128u >> 4) & 0x3FFF) << 16;
        bd |= (unsigned long long)((256u >> 4) & 0x3FFF) << 32;
        bd |= 1ull << 46;
        const unsigned idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 0, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tb), "r"(tb + a_col), "l"(bd),
                     "r"(idesc)
                     : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int c = 0; c < 32; c += 4) {
        unsigned v0, v1, v2, v3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(tb + ((unsigned)(warp * 32) << 16) + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[tid * 32 + c] = v0; out[tid * 32 + c + 1] = v1; out[tid * 32 + c + 2] = v2; out[tid * 32 + c + 3] = v3;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tb) : "memory");
}

int main()
{
    unsigned *d;
    cudaMalloc(&d, 128 * 32 * 4);
    cudaMemset(d, 0xff, 128 * 32 * 4);
    probe<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<unsigned> h(128 * 32);
    cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
    int natural = 0, total = 0;
    for (int m = 0; m < 128; m++)
        for (int k = 0; k < 32; k++) { natural += h[m * 32 + k] == pat(m, k); total++; }
    printf("A operand in tensor memory (kind::i8, M = 128, K = 32): A[m][k] == byte k of lane m's 8 columns for %d/%d entries\n", natural, total);
    if (natural != total) {
        printf("first rows as read back (row: 32 values), expected natural pattern in brackets:\n");
        for (int m = 0; m < 4; m++) { for (int k = 0; k < 32; k++) printf("%u[%u] ", h[m * 32 + k], pat(m, k)); printf("\n"); }
    }
    return 0;
}
