// Probe: which shared-memory bytes does tcgen05.mma read for a K-major SWIZZLE_32B / 128B A operand
// whose start address is advanced by whole rows (the Hankel trick of fir_umma.cu)?
// One MMA  D[128 x 32] = A[128 x 32] . I[32 x 32]  (kind::i8, u8 x s8), so D[m][k] is the byte the
// tensor core fetched for (row m, k).  The plane is filled with byte(o) = (o * 7 + (o >> 8)) & 0x7f
// at PHYSICAL offset o; the host decodes, for every (m, k), which physical offset was read and
// compares it with the address-based model  phys = L ^ swz(L),  L = start + pitch m + k.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o probe_umma_swizzle probe_umma_swizzle.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__global__ void probe(int layout /*6: 32B, 2: 128B, 0: none*/, int start_off, int base_off, unsigned sbo, unsigned lbo, unsigned *out)
{
    extern __shared__ __align__(1024) unsigned char sm[];
    unsigned char *plane = sm;                 // 32 KB, 1024-aligned
    unsigned char *bt = sm + 32768;            // B: 32 x 32 identity, K-major no swizzle: [n/8][chunk][n%8][16]
    __shared__ __align__(8) unsigned long long bar;
    __shared__ unsigned tbase;
    const int tid = threadIdx.x;
    for (int o = tid; o < 32768; o += blockDim.x) plane[o] = (unsigned char)((o * 7 + (o >> 8)) & 0x7f);
    for (int i = tid; i < 32 * 32; i += blockDim.x) {
        const int n = i / 32, k = i % 32;
        bt[(n / 8) * 256 + (k / 16) * 128 + (n % 8) * 16 + (k % 16)] = (n == k) ? 1 : 0;
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tbase)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tb = tbase;
    if (tid == 0) {
        unsigned long long ad = 0, bd = 0;
        const unsigned as = smem_u32(plane) + start_off, bs = smem_u32(bt);
        ad |= (unsigned long long)((as >> 4) & 0x3FFF);
        ad |= (unsigned long long)((lbo >> 4) & 0x3FFF) << 16;
        ad |= (unsigned long long)((sbo >> 4) & 0x3FFF) << 32;
        ad |= 1ull << 46;
        ad |= (unsigned long long)(base_off & 7) << 49;
        ad |= (unsigned long long)(layout & 7) << 61;
        bd |= (unsigned long long)((bs >> 4) & 0x3FFF);
        bd |= (unsigned long long)((128u >> 4) & 0x3FFF) << 16;
        bd |= (unsigned long long)((256u >> 4) & 0x3FFF) << 32;
        bd |= 1ull << 46;
        const unsigned idesc = (2u << 4) | (0u << 7) | (1u << 10) | ((32u >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tb), "l"(ad),
                     "l"(bd), "r"(idesc), "r"(0u)
                     : "memory");
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5;
    for (int c = 0; c < 32; c += 4) {
        unsigned v0, v1, v2, v3;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3) : "r"(tb + ((unsigned)(warp * 32) << 16) + c) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        out[tid * 32 + c] = v0; out[tid * 32 + c + 1] = v1; out[tid * 32 + c + 2] = v2; out[tid * 32 + c + 3] = v3;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tb) : "memory");
}

static unsigned char val(int o) { return (unsigned char)((o * 7 + (o >> 8)) & 0x7f); }

int main()
{
    unsigned *d;
    cudaMalloc(&d, 128 * 32 * 4);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40 * 1024);
    std::vector<unsigned> h(128 * 32);
    struct Case { const char *name; int layout, pitch, sbo, lbo; };
    const Case cases[] = {{"none", 0, 16, 128, 16}, {"sw32", 6, 32, 256, 16}, {"sw128", 2, 128, 1024, 16}};
    for (const Case &cs : cases)
        for (int start : {0, 32, 64, 96, 128, 160, 256, 384, 512, 1024 + 32})
            for (int bo_mode = 0; bo_mode < 2; bo_mode++) {
                const int bo = bo_mode ? ((start >> 7) & 7) : 0;
                if (bo_mode && bo == 0) continue;
                cudaMemset(d, 0xff, 128 * 32 * 4);
                probe<<<1, 128, 36 * 1024>>>(cs.layout, start, bo, cs.sbo, cs.lbo, d);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("%s start %d bo %d: CUDA error %s\n", cs.name, start, bo, cudaGetErrorString(e)); return 1; }
                cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
                // models: absolute  phys = L ^ f(L);  relative  phys = start + (R ^ f(R + bo * 128)), R = pitch m + k
                int ok_abs = 0, ok_rel = 0, ok_plain = 0, total = 0;
                for (int m = 0; m < 128; m++)
                    for (int k = 0; k < 32; k++) {
                        const int R = cs.pitch * m + k, L = start + R;
                        auto f = [&](int a) { return cs.layout == 6 ? ((a >> 7) & 1) << 4 : cs.layout == 2 ? ((a >> 7) & 7) << 4 : 0; };
                        const unsigned got = h[m * 32 + k];
                        ok_abs += got == val(L ^ f(L));
                        ok_rel += got == val(start + (R ^ f(R + bo * 128)));
                        ok_plain += got == val(L);
                        total++;
                    }
                printf("%-6s start %4d base_offset %d : absolute-address model %4d/%d, start-relative model %4d/%d, unswizzled %4d/%d\n", cs.name, start, bo,
                       ok_abs, total, ok_rel, total, ok_plain, total);
            }
    return 0;
}
