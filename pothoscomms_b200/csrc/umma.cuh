// tcgen05 (5th-generation tensor core) building blocks shared by the int16 FIR kernels:
// shared-memory matrix descriptors, the kind::i8 instruction descriptor, MMA issue, TMEM loads.
// Field layouts follow CUTLASS cute/arch/mma_sm100_desc.hpp (SmemDescriptor, InstrDescriptor).
// Addressing facts these kernels rely on were measured with tools/probe_umma_swizzle.cu
// (profiles/r01d_probe_umma_swizzle.txt): with base_offset = 0 the tensor core applies the
// swizzle XOR to the ABSOLUTE shared-memory address, for any 16-byte aligned start address.
#pragma once
#include "bulk.cuh"

namespace b200c {

__device__ __forceinline__ unsigned prmt_u(unsigned a, unsigned b, unsigned sel)
{
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// One lane of a converged warp (elect.sync): unlike `lane == 0` it tells ptxas that exactly one thread runs the
// region, so the tensor-core instructions inside take their operands from uniform registers without the
// per-distinct-value issue loop it otherwise wraps around each of them.
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "elect.sync _|p, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(pred));
    return pred != 0;
}

// K-major operand: start address, LBO (no swizzle: between the 16-byte k chunks), SBO (between 8-row
// groups), layout type (0: no swizzle, 6: 32-byte, 4: 64-byte, 2: 128-byte swizzle)
__device__ __forceinline__ unsigned long long umma_smem_desc(unsigned smem_addr, unsigned lbo_bytes, unsigned sbo_bytes, unsigned layout = 0)
{
    unsigned long long d = 0;
    d |= (unsigned long long)((smem_addr >> 4) & 0x3FFF);
    d |= (unsigned long long)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (unsigned long long)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ull << 46;                                         // descriptor version: Blackwell
    d |= (unsigned long long)(layout & 7) << 61;
    return d;                                                // base offset 0
}

// kind::i8 instruction descriptor: D = int32, A = u8 or s8, B = s8, both K-major, M = 128
__host__ __device__ constexpr unsigned umma_idesc_i8(bool a_signed, int N)
{
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | (1u << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
}

// the same with both signednesses free: the swapped formulation has the s8 tap digits as A and the u8 / s8 data limbs as B
__host__ __device__ constexpr unsigned umma_idesc_i8_ab(bool a_signed, bool b_signed, int N)
{
    return (2u << 4) | ((a_signed ? 1u : 0u) << 7) | ((b_signed ? 1u : 0u) << 10) | ((unsigned)(N >> 3) << 17) | ((128u >> 4) << 24);
}

__device__ __forceinline__ void umma_i8(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc, bool accumulate)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((unsigned)accumulate)
        : "memory");
}

// the same with the accumulate flag fixed at compile time: D = A.B (first) / D += A.B (acc)
__device__ __forceinline__ void umma_i8_first(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_i8_acc(unsigned d_tmem, unsigned long long adesc, unsigned long long bdesc, unsigned idesc)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc)
        : "memory");
}

// A operand in tensor memory (lane m = row m, 8 columns = 32 k bytes), B through a shared-memory descriptor
__device__ __forceinline__ void umma_i8_ts_first(unsigned d_tmem, unsigned a_tmem, unsigned long long bdesc, unsigned idesc)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc)
        : "memory");
}
__device__ __forceinline__ void umma_i8_ts_acc(unsigned d_tmem, unsigned a_tmem, unsigned long long bdesc, unsigned idesc)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc)
        : "memory");
}

__device__ __forceinline__ void tmem_ld4(unsigned taddr, unsigned (&v)[4])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}

// 16 consecutive columns of this thread's TMEM lane: one wide load instead of four narrow ones
// (tcgen05.ld cost is per instruction: 4-column loads made the epilogue the longest phase)
__device__ __forceinline__ void tmem_ld16(unsigned taddr, unsigned (&v)[16])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}

// 16 lanes x 16 columns in the accumulator-fragment layout of an m16n8 tile, twice: register 4 g + 2 h + e of
// thread t = lane (t / 4 + 8 h), column 8 g + 2 (t % 4) + e  (tools/probe_tmem_ld_shapes.cu)
__device__ __forceinline__ void tmem_ld16x256b_x2(unsigned taddr, unsigned (&v)[8])
{
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Watchdog of the warp-specialised kernels: a barrier wait longer than ~25 s (a time-sliced GPU can park a kernel for a
// while; no pipeline hand-off takes milliseconds) is a deadlock, and the kernel traps
// (the launch fails loudly) instead of hanging the stream.  With B200C_UMMA_DBG the stuck waits are first
// recorded in host-mapped memory: entry [block * 32 + warp] = shared address of the barrier << 8 | parity << 1 | 1.
static __device__ unsigned long long *g_umma_watch = nullptr;

// suspend-time hint: the hardware parks the thread until the phase completes or the hint (ns) runs out, instead of
// returning after its short default slice -- the spin loops around it were the kernel's largest instruction stream
// (ncu r02ac: branch_resolving the top stall, 2.4 warp instructions per sample against ~0.8 of useful work)
__device__ __forceinline__ bool mbar_try_wait(unsigned long long *bar, unsigned parity, unsigned hint_ns = 1000000u)
{
    unsigned ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(hint_ns) : "memory");
    return ok != 0;
}


__device__ __forceinline__ void timed_wait(unsigned long long *bar, unsigned parity, long long &acc)
{
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {}
    acc += clock64() - t0;
}

__device__ __forceinline__ void watched_wait(bool timed, unsigned long long *bar, unsigned parity, long long &acc)
{
    // the phase is usually complete, or completes within the first suspended try: no clock reads on that path unless
    // the role timers are on
    if (!timed && mbar_try_wait(bar, parity)) return;
    const long long t0 = clock64();
    bool noted = false;
    while (!mbar_try_wait(bar, parity)) {
        const long long waited = clock64() - t0;
        if (waited > (1ll << 35) && !noted) {
            noted = true;
            if (g_umma_watch) {
                g_umma_watch[blockIdx.x * 32 + (threadIdx.x >> 5)] = ((unsigned long long)smem_u32(bar) << 8) | (parity << 1) | 1ull;
                __threadfence_system();
            }
        }
        if (waited > (3ll << 34)) __trap();
    }
    acc += clock64() - t0;
}


} // namespace b200c
