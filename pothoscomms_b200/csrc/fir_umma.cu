// Bit-exact int16 / complex-int16 FIR (L = M = 1) on the 5th-generation tensor cores:
// tcgen05.mma kind::i8 with the accumulators in tensor memory (TMEM).
//
// Same byte-limb algebra as fir_imma.cu (fir_imma.hpp): x = xl + 2^8 xh, Q16 tap q = sum_l 2^(8l) q_l
// with balanced signed digits, every limb product an exact int8 GEMM, recombined mod 2^32.
//
// GEMM shape.  One MMA is  D[128 x N] += A[128 x 32] . B[32 x N]  (int32 += u8/s8 . s8):
//  * A is never materialised.  Row m of A is the 32-byte window  plane[16 m + 32 b + (0..31)]  of one
//    byte plane of the input (re/im x lo/hi).  In the K-major no-swizzle canonical layout the
//    operand is a grid of 8-row x 16-byte core matrices addressed by two strides (SBO between
//    8-row groups, LBO between 16-byte k chunks): SBO = 128 B and LBO = 16 B alias that grid onto
//    the plain byte plane, overlapping windows included -- the Hankel structure of the data
//    matrix is expressed by the shared-memory descriptor alone.
//  * B (host-built at setTaps) stacks, for each tap limb plane q, the 16-column Toeplitz slice
//    B[j][NQ n + q] = digit_q(h[n + K-1 - j]); N = 16 x NQ, NQ = (tap components) x (tap limbs).
//  * D[m][NQ n + q] is the partial sum of output y[16 m + n] for (data plane, tap plane q); one
//    accumulator region of N TMEM columns per data plane, NB = ceil((K + 15) / 32) k-blocks each.
// 128 x 16 = 2048 outputs per tile; the epilogue reads the accumulators with tcgen05.ld (thread =
// TMEM lane = row m), recombines the limb sums with shifts and stores 16 consecutive outputs.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bulk.cuh"
#include "fir_imma.hpp"
#include "umma.cuh"

namespace b200c {

struct FirUmmaArgs {
    const void *in;
    void *out;
    const void *btile;   // [NB][N x 32 B] canonical K-major no-swizzle B tiles
    long long n_in, n_out, ntiles;
    int K, NB, PL;       // PL: bytes per plane = 2048 + 32 NB
    int R;               // warp-specialised kernel: depth of the bulk-copy landing ring
    long long *dbg;      // optional [grid][8] cycle counters of the roles' barrier waits (B200C_UMMA_DBG)
};

constexpr int kUmmaTile = 2048;    // outputs per CTA tile: 128 rows x 16
constexpr int kUmmaThreads = 128;

// Epilogue loads for CH consecutive outputs n0.. of one row: columns are output-major
// (col = n NQ + q), so the CH x NQ accumulators of one data plane are adjacent.
template <int NPL, int NQ, int CH>
__device__ __forceinline__ void load_acc(unsigned taddr, int N, int n0, unsigned (&v)[NPL][CH * NQ])
{
    static_assert(CH * NQ == 16 || CH * NQ == 4, "chunk must be 4 or 16 columns per plane");
#pragma unroll
    for (int p = 0; p < NPL; p++) {
        if constexpr (CH * NQ == 16) tmem_ld16(taddr + (unsigned)(p * N + n0 * NQ), v[p]);
        else tmem_ld4(taddr + (unsigned)(p * N + n0 * NQ), v[p]);
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Recombine the limb sums of CH outputs mod 2^32, keep bits [16, 32) of each (fromQ and the
// narrowing store, filter/FIRFilter.cpp:300) and store them at output index o.
template <int DC, int TC, int NLT, int CH>
__device__ __forceinline__ void finish_chunk(const unsigned (&v)[DC * 2][CH * TC * NLT], const long long o, const FirUmmaArgs &a)
{
    constexpr int NQ = TC * NLT;
    unsigned res[CH];                                 // complex: packed (re, im) int16; real: one int16 in the low half
#pragma unroll
    for (int i = 0; i < CH; i++) {
        // term(dc, tc) = sum_{dl, l} A[2 dc + dl][tc NLT + l] << 8 (dl + l), shifts >= 32 vanish
        unsigned term[DC][TC];
#pragma unroll
        for (int dc = 0; dc < DC; dc++)
#pragma unroll
            for (int tc = 0; tc < TC; tc++) {
                unsigned s = 0;
#pragma unroll
                for (int dl = 0; dl < 2; dl++)
#pragma unroll
                    for (int l = 0; l < NLT; l++)
                        if (dl + l < 4) s += v[2 * dc + dl][i * NQ + tc * NLT + l] << (8 * (dl + l));
                term[dc][tc] = s;
            }
        if constexpr (DC == 2) {
            const unsigned yr = TC == 2 ? term[0][0] - term[1][TC - 1] : term[0][0];
            const unsigned yi = TC == 2 ? term[0][TC - 1] + term[1][0] : term[1][0];
            res[i] = prmt_u(yr, yi, 0x7632);
        } else {
            res[i] = term[0][0] >> 16;
        }
    }
    if constexpr (DC == 2) {
        unsigned *out32 = static_cast<unsigned *>(a.out);
        if (o + CH <= a.n_out && (reinterpret_cast<unsigned long long>(out32) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < CH; i += 4) __stcg(reinterpret_cast<uint4 *>(out32 + o + i), make_uint4(res[i], res[i + 1], res[i + 2], res[i + 3]));
        } else {
#pragma unroll
            for (int i = 0; i < CH; i++)
                if (o + i < a.n_out) out32[o + i] = res[i];
        }
    } else {
        static_assert(DC == 2 || CH == 8, "real data: 8 outputs per chunk");
        unsigned short *out16 = static_cast<unsigned short *>(a.out);
        if (o + CH <= a.n_out && (reinterpret_cast<unsigned long long>(out16) & 15) == 0) {
            __stcg(reinterpret_cast<uint4 *>(out16 + o), make_uint4(res[0] | (res[1] << 16), res[2] | (res[3] << 16), res[4] | (res[5] << 16),
                                                                   res[6] | (res[7] << 16)));
        } else {
#pragma unroll
            for (int i = 0; i < CH; i++)
                if (o + i < a.n_out) out16[o + i] = (unsigned short)res[i];
        }
    }
}

template <int DC, int TC, int NLT>
__global__ void __launch_bounds__(kUmmaThreads, 1) fir_umma_kernel(const FirUmmaArgs a)
{
    extern __shared__ __align__(128) unsigned char smem_u[];
    constexpr int NPL = DC * 2;                    // data byte planes [dc][lo, hi]
    constexpr int NQ = TC * NLT;                   // tap limb planes
    constexpr int N = 16 * NQ;                     // MMA N
    constexpr int COLS = NPL * N;                  // accumulator columns in use
    constexpr int ALLOC = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
    constexpr int ESZ = DC * 2;                    // bytes per input sample
    constexpr int CH = 16 / NQ;                    // outputs per epilogue chunk: 16 accumulator columns per data plane
    static_assert(COLS <= 512, "accumulators exceed tensor memory");
    const int NB = a.NB, PL = a.PL;
    unsigned char *btile = smem_u;                                   // NB * N * 32 bytes
    unsigned char *planes = btile + (size_t)NB * N * 32;             // NPL * PL
    unsigned char *raw = planes + (size_t)NPL * PL;                  // landing zone of the bulk prefetch
    __shared__ __align__(8) unsigned long long bar_in, bar_mma;
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < NB * N * 2; i += kUmmaThreads)             // B tiles: 16-byte pieces
        reinterpret_cast<uint4 *>(btile)[i] = __ldg(static_cast<const uint4 *>(a.btile) + i);
    if (tid == 0) { mbar_init(&bar_in, 1); mbar_init(&bar_mma, 1); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ALLOC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;

    const bool al = (reinterpret_cast<unsigned long long>(a.in) & 15) == 0;
    auto bulk_ok = [&](long long tile) { return al && tile < a.ntiles && tile * kUmmaTile + PL <= a.n_in; };
    bool pending = bulk_ok(blockIdx.x);
    if (pending && tid == 0)
        bulk_load(raw, static_cast<const unsigned char *>(a.in) + (size_t)blockIdx.x * kUmmaTile * ESZ, (unsigned)(NPL * PL), &bar_in);
    unsigned par_in = 0, par_mma = 0;

    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const long long o0 = tile * kUmmaTile;
        // ---- stage: de-interleave the tile's input window into byte planes
        if (pending) { mbar_wait(&bar_in, par_in); par_in ^= 1; }
        if constexpr (DC == 2) {
            const unsigned *__restrict__ in32 = static_cast<const unsigned *>(a.in);
            for (int q = tid; q < PL / 4; q += kUmmaThreads) {
                const long long s = o0 + 4LL * q;
                unsigned s0, s1, s2, s3;
                if (pending) {
                    const uint4 v = reinterpret_cast<const uint4 *>(raw)[q];
                    s0 = v.x; s1 = v.y; s2 = v.z; s3 = v.w;
                } else {
                    s0 = s < a.n_in ? __ldg(in32 + s) : 0u;
                    s1 = s + 1 < a.n_in ? __ldg(in32 + s + 1) : 0u;
                    s2 = s + 2 < a.n_in ? __ldg(in32 + s + 2) : 0u;
                    s3 = s + 3 < a.n_in ? __ldg(in32 + s + 3) : 0u;
                }
                const unsigned t01 = prmt_u(s0, s1, 0x5140), t23 = prmt_u(s2, s3, 0x5140);
                const unsigned u01 = prmt_u(s0, s1, 0x7362), u23 = prmt_u(s2, s3, 0x7362);
                unsigned *p = reinterpret_cast<unsigned *>(planes) + q;
                p[0] = prmt_u(t01, t23, 0x5410);
                p[PL / 4] = prmt_u(t01, t23, 0x7632);
                p[2 * (PL / 4)] = prmt_u(u01, u23, 0x5410);
                p[3 * (PL / 4)] = prmt_u(u01, u23, 0x7632);
            }
        } else {
            const unsigned short *__restrict__ in16 = static_cast<const unsigned short *>(a.in);
            for (int q = tid; q < PL / 4; q += kUmmaThreads) {
                const long long s = o0 + 4LL * q;
                unsigned w0, w1;
                if (pending) {
                    const uint2 v = reinterpret_cast<const uint2 *>(raw)[q];
                    w0 = v.x; w1 = v.y;
                } else {
                    const unsigned x0 = s < a.n_in ? __ldg(in16 + s) : 0u, x1 = s + 1 < a.n_in ? __ldg(in16 + s + 1) : 0u;
                    const unsigned x2 = s + 2 < a.n_in ? __ldg(in16 + s + 2) : 0u, x3 = s + 3 < a.n_in ? __ldg(in16 + s + 3) : 0u;
                    w0 = x0 | (x1 << 16); w1 = x2 | (x3 << 16);
                }
                unsigned *p = reinterpret_cast<unsigned *>(planes) + q;
                p[0] = prmt_u(w0, w1, 0x6420);
                p[PL / 4] = prmt_u(w0, w1, 0x7531);
            }
        }
        // generic-proxy writes (planes, B tiles) -> visible to the tensor core's async proxy
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        pending = bulk_ok(tile + gridDim.x);
        if (tid == 0) {
            if (pending)
                bulk_load(raw, static_cast<const unsigned char *>(a.in) + (size_t)(tile + gridDim.x) * kUmmaTile * ESZ,
                          (unsigned)(NPL * PL), &bar_in);
            // ---- one thread issues the tile's MMAs: (data plane) x (k-block), accumulating over k-blocks
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const unsigned planes_s = smem_u32(planes), btile_s = smem_u32(btile);
            for (int b = 0; b < NB; b++) {
                const unsigned long long bdesc = umma_smem_desc(btile_s + (unsigned)b * N * 32, 128, 256);
#pragma unroll
                for (int p = 0; p < NPL; p++) {
                    const unsigned long long adesc = umma_smem_desc(planes_s + (unsigned)p * PL + 32u * b, 16, 128);
                    umma_i8(tmem_base + (unsigned)(p * N), adesc, bdesc, umma_idesc_i8((p & 1) != 0, N), b > 0);
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
        }
        mbar_wait(&bar_mma, par_mma);
        par_mma ^= 1;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

        // ---- epilogue: thread = TMEM lane = row m = outputs o0 + 16 m + (0..15)
        const unsigned lane_addr = tmem_base + ((unsigned)(warp * 32) << 16);
        const long long orow = o0 + 16LL * tid;
#pragma unroll 1
        for (int n0 = 0; n0 < 16; n0 += CH) {
            unsigned v[NPL][CH * NQ];
            load_acc<NPL, NQ, CH>(lane_addr, N, n0, v);
            finish_chunk<DC, TC, NLT, CH>(v, orow + n0, a);
        }
        // accumulators and planes are free once every thread is past its TMEM loads
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ALLOC) : "memory");
}

// (A warp-specialised form of this kernel -- bulk-copy issuer / stagers / MMA issuer / epilogue on mbarrier pipelines --
// was the stepping stone to fir_umma32_kernel in round 1; it never beat the 32-outputs-per-row kernel that took over its
// structure and was removed in round 2.  This single-role kernel stays for filters whose tables exceed what
// fir_umma32_kernel keeps in shared memory: complex int16 from ~300 to ~1000 taps, 2.8x the mma.sync kernel there,
// profiles/r02_sweep_dispatch.jsonl.)

// ------------------------------------------------------------------------------- host ---
static int8_t umma_digit(int32_t q, int l)
{
    long long r = q, d = 0;
    for (int i = 0; i <= l; i++) {
        d = ((r + 128) & 255) - 128;
        r = (r - d) >> 8;
    }
    return (int8_t)d;
}

int fir_umma_configure(FirUmmaPlan &p, const FirImmaPlan &base, const double *taps, bool force)
{
    p.ready = false;
    // B200C_UMMA=0 keeps int16 streams on the mma.sync kernel (B200C_FIR_ALGO=umma still forces this one)
    const bool enabled = [] { const char *e = std::getenv("B200C_UMMA"); return !e || std::atoi(e) != 0; }();
    if (!base.ready || !(enabled || force) || base.nlt != 2) return B200C_OK;
    const int K = base.K, tc = base.tc, nlt = base.nlt, NQ = tc * nlt, N = 16 * NQ;
    const int NB = (K + 15 + 31) / 32;
    const size_t smem = (size_t)NB * N * 32 + 2 * ((size_t)base.dc * 2 * (kUmmaTile + 32 * NB));
    if (smem > 96 * 1024) return B200C_OK;                  // very long filters stay on the mma.sync kernel
    std::vector<uint8_t> bt((size_t)NB * N * 32, 0);
    for (int b = 0; b < NB; b++)
        for (int c = 0; c < tc; c++)
            for (int l = 0; l < nlt; l++)
                for (int n = 0; n < 16; n++)
                    for (int jj = 0; jj < 32; jj++) {
                        const int d = n + K - 1 - (32 * b + jj);
                        if (d < 0 || d >= K) continue;
                        const int32_t q = (int32_t)(long long)std::ldexp(taps[(size_t)d * tc + c], 16);
                        const int col = n * NQ + (c * nlt + l);   // output-major: the accumulators of one output are adjacent
                        // canonical K-major no-swizzle: [col / 8][k chunk][col % 8][16 bytes]
                        const size_t at = (size_t)b * N * 32 + (size_t)(col / 8) * 256 + (size_t)(jj / 16) * 128 + (size_t)(col % 8) * 16 + (jj % 16);
                        bt[at] = (uint8_t)umma_digit(q, l);
                    }
    if (bt.size() > p.capacity) {
        if (p.d_btile) cudaFree(p.d_btile);
        p.d_btile = nullptr; p.capacity = 0;
        B200C_CUDA_TRY(cudaMalloc(&p.d_btile, bt.size()));
        p.capacity = bt.size();
    }
    B200C_CUDA_TRY(cudaMemcpy(p.d_btile, bt.data(), bt.size(), cudaMemcpyHostToDevice));
    p.K = K; p.NB = NB; p.dc = base.dc; p.tc = tc; p.nlt = nlt;
    p.ready = true;
    return B200C_OK;
}

void fir_umma_destroy(FirUmmaPlan &p)
{
    if (p.d_btile) cudaFree(p.d_btile);
    p.d_btile = nullptr; p.capacity = 0; p.ready = false;
}

template <int DC, int TC, int NLT>
static int launch_umma(const FirUmmaArgs &a, size_t smem, int sm_count, cudaStream_t stream)
{
    int dev = 0;
    B200C_CUDA_TRY(cudaGetDevice(&dev));
    // Tensor memory holds 512 columns per SM: the CTAs resident on one SM must not ask for more, or
    // tcgen05.alloc would wait forever.  Shared memory is padded so that at most `per_sm` CTAs fit.
    auto kern = fir_umma_kernel<DC, TC, NLT>;
    static thread_local bool configured[16] = {false};
    if (dev < 16 && !configured[dev]) {
        B200C_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        configured[dev] = true;
    }
    constexpr int COLS = DC * 2 * 16 * TC * NLT;
    constexpr int ALLOC = COLS <= 32 ? 32 : COLS <= 64 ? 64 : COLS <= 128 ? 128 : COLS <= 256 ? 256 : 512;
    const int per_sm = std::min(512 / ALLOC, 4);
    smem = std::max(smem, (size_t)(227 * 1024) / (per_sm + 1) + 1024);   // more than a (per_sm + 1)-th of the SM
    const int grid = (int)std::min<long long>(a.ntiles, (long long)sm_count * per_sm);
    kern<<<grid, kUmmaThreads, smem, stream>>>(a);
    B200C_CUDA_TRY(cudaGetLastError());
    return B200C_OK;
}

int fir_umma_launch(const FirUmmaPlan &p, const void *d_in, size_t in_elems, void *d_out, size_t n_out, int sm_count,
                    cudaStream_t stream)
{
    if (n_out == 0) return B200C_OK;
    FirUmmaArgs a;
    a.in = d_in; a.out = d_out; a.btile = p.d_btile;
    a.n_in = (long long)in_elems; a.n_out = (long long)n_out;
    a.ntiles = ((long long)n_out + kUmmaTile - 1) / kUmmaTile;
    a.K = p.K; a.NB = p.NB; a.PL = kUmmaTile + 32 * p.NB; a.R = 2; a.dbg = nullptr;
    const size_t smem = (size_t)p.NB * 16 * p.tc * p.nlt * 32 + 2 * ((size_t)p.dc * 2 * a.PL) + 128;
    if (p.dc == 1) return launch_umma<1, 1, 2>(a, smem, sm_count, stream);
    if (p.tc == 1) return launch_umma<2, 1, 2>(a, smem, sm_count, stream);
    return launch_umma<2, 2, 2>(a, smem, sm_count, stream);
}

} // namespace b200c
