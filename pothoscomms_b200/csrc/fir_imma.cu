// Bit-exact int16 / complex-int16 FIR (L = M = 1) on the int8 tensor cores.
//
// The reference nest (filter/FIRFilter.cpp:286-302) for int16 data multiplies by Q16 taps held in
// int32 (floatToQ, :348), accumulates in wrapping int32 (QType, :381) and keeps bits [16, 32) of
// the sum (fromQ >> 16 and the narrowing store, :300).  That is arithmetic mod 2^32, so it can be
// done exactly in byte limbs (fir_imma.hpp): every limb product is an int8 x int8 -> int32
// Toeplitz GEMM on `mma.sync.m16n8k32` (IMMA), 30x the IMAD rate of the direct kernel
// (tools/probe_imma.cu), and the limb sums are recombined with shifts in the epilogue.
//
// GEMM shape.  Row m of an A tile is the 32-byte window  in[8 m + 32 b + (0..31)]  of one byte
// plane of the input (re/im x lo/hi, de-interleaved into shared memory once per CTA tile), column
// n of the B block b is the Toeplitz slice  B[j][n] = h[n + K-1 - j]  of one tap limb plane, so
//   D[m][n] = sum_j in[8 m + j] h[n + K-1 - j] = y[8 m + n]            (in[] starts K-1 early)
// over NB = ceil((K + 7) / 32) k-blocks.  16 x 8 = 128 consecutive outputs per MMA tile; a lane's
// accumulators are output pairs (8 g + 2 t, +1), so results are stored straight from registers,
// coalesced.  The k index of a block is permuted (slot 4t+i <-> j = 8t+i, slot 16+4t+i <->
// j = 8t+4+i, the same on both operands) so that a lane's two A registers of a row are ONE
// aligned 64-bit shared-memory load of the plain byte plane: the Hankel structure of A costs no
// materialisation at all.  B fragments are built per lane on the host at setTaps().
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "bulk.cuh"
#include "fir_imma.hpp"

namespace b200c {

struct FirImmaArgs {
    const void *in;
    void *out;
    const uint2 *frag;
    long long n_in, n_out, ntiles;
    int K, NB, PL;      // PL: bytes per plane = kImmaTile + 32 NB
};

constexpr int kImmaTile = 4096;    // outputs per CTA tile
constexpr int kImmaWarps = 4;

template <bool A_SIGNED>
__device__ __forceinline__ void imma16832(int (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1)
{
    if constexpr (A_SIGNED)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    else
        asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                     : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ unsigned prmt(unsigned a, unsigned b, unsigned sel)
{
    unsigned r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// accumulator class of the product (data component dc) x (tap component tc):
//   complex x complex: (re,re) -> 0 [+yr], (im,im) -> 1 [-yr], mixed -> 2 [yi]
//   complex x real:    re -> 0 [yr], im -> 1 [yi];   real x real: 0
template <int DC, int TC> __host__ __device__ constexpr int imma_cls(int dc, int tc)
{
    return TC == 2 ? (dc == tc ? dc : 2) : dc;
}

template <int DC, int TC, int NLT, int R, int MINB>
__global__ void __launch_bounds__(32 * kImmaWarps, MINB) fir_imma_kernel(const FirImmaArgs a)
{
    extern __shared__ __align__(16) unsigned char smem_i[];
    constexpr int NPL = DC * 2;                              // byte planes [dc][lo, hi]
    constexpr int NG = NLT + 1 > 4 ? 4 : NLT + 1;            // shift groups 2^(8 s), s < 4
    constexpr int NCLS = DC == 1 ? 1 : (TC == 2 ? 3 : 2);
    constexpr int NTH = 32 * kImmaWarps;
    constexpr int UNITS = kImmaTile / (R * 128);
    const int NB = a.NB, PL = a.PL;
    uint2 *fragS = reinterpret_cast<uint2 *>(smem_i);
    unsigned char *planes = smem_i + (size_t)NB * TC * NLT * 32 * sizeof(uint2);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5, g = lane >> 2, t = lane & 3;

    unsigned char *raw = planes + (size_t)NPL * PL;          // landing zone of the bulk prefetch
    __shared__ __align__(8) unsigned long long bar;
    constexpr int ESZ = DC * 2;                              // bytes per input sample

    for (int i = tid; i < NB * TC * NLT * 32; i += NTH) fragS[i] = a.frag[i];
    // A tile whose whole PL-sample window lies inside the stream (and a 16-byte aligned base) is
    // fetched by ONE bulk copy issued before the previous tile's MMA phase; edge tiles and
    // unaligned streams use guarded loads.
    const bool al = (reinterpret_cast<unsigned long long>(a.in) & 15) == 0;
    auto bulk_ok = [&](long long tile) { return al && tile < a.ntiles && tile * kImmaTile + PL <= a.n_in; };
    if (tid == 0) mbar_init(&bar, 1);
    __syncthreads();
    bool pending = bulk_ok(blockIdx.x);
    if (pending && tid == 0)
        bulk_load(raw, static_cast<const unsigned char *>(a.in) + (size_t)blockIdx.x * kImmaTile * ESZ, (unsigned)(NPL * PL), &bar);
    unsigned parity = 0;

    for (long long tile = blockIdx.x; tile < a.ntiles; tile += gridDim.x) {
        const long long o0 = tile * kImmaTile;
        // ---- stage: de-interleave the tile's input window into byte planes
        if (pending) {
            mbar_wait(&bar, parity);
            parity ^= 1;
        }
        if constexpr (DC == 2) {
            const unsigned *__restrict__ in32 = static_cast<const unsigned *>(a.in);
            for (int q = tid; q < PL / 4; q += NTH) {
                const long long s = o0 + 4LL * q;
                unsigned s0, s1, s2, s3;
                if (pending) {
                    const uint4 v = reinterpret_cast<const uint4 *>(raw)[q];
                    s0 = v.x; s1 = v.y; s2 = v.z; s3 = v.w;
                } else if (al && s + 4 <= a.n_in) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4 *>(in32 + s));
                    s0 = v.x; s1 = v.y; s2 = v.z; s3 = v.w;
                } else {
                    s0 = s < a.n_in ? __ldg(in32 + s) : 0u;
                    s1 = s + 1 < a.n_in ? __ldg(in32 + s + 1) : 0u;
                    s2 = s + 2 < a.n_in ? __ldg(in32 + s + 2) : 0u;
                    s3 = s + 3 < a.n_in ? __ldg(in32 + s + 3) : 0u;
                }
                const unsigned t01 = prmt(s0, s1, 0x5140), t23 = prmt(s2, s3, 0x5140);   // (re_lo, re_lo, re_hi, re_hi)
                const unsigned u01 = prmt(s0, s1, 0x7362), u23 = prmt(s2, s3, 0x7362);   // (im_lo, im_lo, im_hi, im_hi)
                unsigned *p = reinterpret_cast<unsigned *>(planes) + q;
                p[0] = prmt(t01, t23, 0x5410);
                p[PL / 4] = prmt(t01, t23, 0x7632);
                p[2 * (PL / 4)] = prmt(u01, u23, 0x5410);
                p[3 * (PL / 4)] = prmt(u01, u23, 0x7632);
            }
        } else {
            const unsigned short *__restrict__ in16 = static_cast<const unsigned short *>(a.in);
            for (int q = tid; q < PL / 4; q += NTH) {
                const long long s = o0 + 4LL * q;
                unsigned w0, w1;
                if (pending) {
                    const uint2 v = reinterpret_cast<const uint2 *>(raw)[q];
                    w0 = v.x; w1 = v.y;
                } else if (al && s + 4 <= a.n_in) {
                    const uint2 v = __ldg(reinterpret_cast<const uint2 *>(in16 + s));
                    w0 = v.x; w1 = v.y;
                } else {
                    const unsigned x0 = s < a.n_in ? __ldg(in16 + s) : 0u, x1 = s + 1 < a.n_in ? __ldg(in16 + s + 1) : 0u;
                    const unsigned x2 = s + 2 < a.n_in ? __ldg(in16 + s + 2) : 0u, x3 = s + 3 < a.n_in ? __ldg(in16 + s + 3) : 0u;
                    w0 = x0 | (x1 << 16); w1 = x2 | (x3 << 16);
                }
                unsigned *p = reinterpret_cast<unsigned *>(planes) + q;
                p[0] = prmt(w0, w1, 0x6420);
                p[PL / 4] = prmt(w0, w1, 0x7531);
            }
        }
        __syncthreads();                                     // planes complete, landing zone consumed
        pending = bulk_ok(tile + gridDim.x);
        if (pending && tid == 0)
            bulk_load(raw, static_cast<const unsigned char *>(a.in) + (size_t)(tile + gridDim.x) * kImmaTile * ESZ,
                      (unsigned)(NPL * PL), &bar);

        // ---- R x 128 outputs per warp pass
        for (int u = w; u < UNITS; u += kImmaWarps) {
            const long long ob = o0 + (long long)u * (R * 128);
            if (ob >= a.n_out) break;
            int acc[R][NCLS][NG][4];
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int c = 0; c < NCLS; c++)
#pragma unroll
                    for (int s = 0; s < NG; s++)
#pragma unroll
                        for (int i = 0; i < 4; i++) acc[r][c][s][i] = 0;
            const unsigned char *rowp = planes + u * (R * 128) + 8 * g + 8 * t;
            const uint2 *fp = fragS + lane;
#pragma unroll 1
            for (int b = 0; b < NB; b++) {
                uint2 tf[TC][NLT];
#pragma unroll
                for (int c = 0; c < TC; c++)
#pragma unroll
                    for (int l = 0; l < NLT; l++) tf[c][l] = fp[((b * TC + c) * NLT + l) * 32];
#pragma unroll
                for (int r = 0; r < R; r++) {
                    uint2 x0[NPL], x1[NPL];                  // rows g and g + 8 of every plane
#pragma unroll
                    for (int pl = 0; pl < NPL; pl++) {
                        const unsigned char *q = rowp + (size_t)pl * PL + r * 128 + 32 * b;
                        x0[pl] = *reinterpret_cast<const uint2 *>(q);
                        x1[pl] = *reinterpret_cast<const uint2 *>(q + 64);
                    }
#pragma unroll
                    for (int dc = 0; dc < DC; dc++)
#pragma unroll
                        for (int tc = 0; tc < TC; tc++) {
                            const int cls = imma_cls<DC, TC>(dc, tc);
#pragma unroll
                            for (int l = 0; l < NLT; l++) {
                                // x_lo q_l -> 2^(8 l);  x_hi q_l -> 2^(8 (l + 1)), dropped at 2^32
                                imma16832<false>(acc[r][cls][l], x0[2 * dc].x, x1[2 * dc].x, x0[2 * dc].y, x1[2 * dc].y, tf[tc][l].x,
                                                 tf[tc][l].y);
                                if (l + 1 < NG)
                                    imma16832<true>(acc[r][cls][(l + 1) % NG], x0[2 * dc + 1].x, x1[2 * dc + 1].x, x0[2 * dc + 1].y,
                                                    x1[2 * dc + 1].y, tf[tc][l].x, tf[tc][l].y);
                            }
                        }
                }
            }
            // ---- recombine the limb sums mod 2^32, keep bits [16, 32) (fromQ), store
#pragma unroll
            for (int r = 0; r < R; r++)
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    unsigned v[2][NCLS];
#pragma unroll
                    for (int e = 0; e < 2; e++)
#pragma unroll
                        for (int c = 0; c < NCLS; c++) {
                            unsigned sum = (unsigned)acc[r][c][0][2 * h + e];
#pragma unroll
                            for (int s = 1; s < NG; s++) sum += (unsigned)acc[r][c][s][2 * h + e] << (8 * s);
                            v[e][c] = sum;
                        }
                    const long long o = ob + r * 128 + 8 * (g + 8 * h) + 2 * t;
                    if constexpr (DC == 2) {
                        unsigned pk[2];
#pragma unroll
                        for (int e = 0; e < 2; e++) {
                            const unsigned yr = TC == 2 ? v[e][0] - v[e][1] : v[e][0];
                            const unsigned yi = TC == 2 ? v[e][2] : v[e][1];
                            pk[e] = prmt(yr, yi, 0x7632);
                        }
                        unsigned *out32 = static_cast<unsigned *>(a.out);
                        if (o + 1 < a.n_out && (reinterpret_cast<unsigned long long>(out32) & 7) == 0) {
                            __stcg(reinterpret_cast<uint2 *>(out32 + o), make_uint2(pk[0], pk[1]));
                        } else {
                            if (o < a.n_out) out32[o] = pk[0];
                            if (o + 1 < a.n_out) out32[o + 1] = pk[1];
                        }
                    } else {
                        unsigned short *out16 = static_cast<unsigned short *>(a.out);
                        if (o + 1 < a.n_out && (reinterpret_cast<unsigned long long>(out16) & 3) == 0) {
                            __stcg(reinterpret_cast<unsigned *>(out16 + o), prmt(v[0][0], v[1][0], 0x7632));
                        } else {
                            if (o < a.n_out) out16[o] = (unsigned short)(v[0][0] >> 16);
                            if (o + 1 < a.n_out) out16[o + 1] = (unsigned short)(v[1][0] >> 16);
                        }
                    }
                }
        }
        __syncthreads();                                     // planes are free for the next tile
    }
}

// ------------------------------------------------------------------------------- host ---
// Q16 tap exactly as the direct path stores it (fir.cu store_tap: trunc(ldexp(v, 16)) -> int32)
static int32_t q16_tap(double v) { return (int32_t)(long long)std::ldexp(v, 16); }

static int digits_needed(int32_t q)
{
    long long r = q;
    int n = 0;
    while (r != 0 && n < 4) {
        const long long d = ((r + 128) & 255) - 128;
        r = (r - d) >> 8;
        n++;
    }
    return n;   // a non-zero remainder after 4 digits is a multiple of 2^32: irrelevant
}

static int8_t digit(int32_t q, int l)
{
    long long r = q, d = 0;
    for (int i = 0; i <= l; i++) {
        d = ((r + 128) & 255) - 128;
        r = (r - d) >> 8;
    }
    return (int8_t)d;
}

int fir_imma_configure(FirImmaPlan &p, int dtype, const double *taps, size_t ntaps, bool complex_taps, size_t M, size_t L,
                       bool force)
{
    p.ready = false;
    if ((dtype != B200C_I16 && dtype != B200C_CI16) || M != 1 || L != 1) return B200C_OK;
    if (ntaps > kFirImmaMaxTaps || (!force && ntaps < kFirImmaMinTaps)) return B200C_OK;
    const int K = (int)ntaps, tc = complex_taps ? 2 : 1;
    std::vector<int32_t> q((size_t)K * tc);
    int nlt = 2;
    for (size_t i = 0; i < q.size(); i++) {
        q[i] = q16_tap(taps[i]);
        nlt = std::max(nlt, digits_needed(q[i]));
    }
    const int NB = (K + 7 + 31) / 32;
    std::vector<uint32_t> frag((size_t)NB * tc * nlt * 32 * 2, 0u);
    for (int b = 0; b < NB; b++)
        for (int c = 0; c < tc; c++)
            for (int l = 0; l < nlt; l++)
                for (int lane = 0; lane < 32; lane++) {
                    const int g = lane >> 2, t = lane & 3;
                    uint32_t regs[2] = {0u, 0u};
                    for (int half = 0; half < 2; half++)
                        for (int i = 0; i < 4; i++) {
                            const int j = 32 * b + 8 * t + 4 * half + i;     // window position of this k slot
                            const int d = g + K - 1 - j;                      // tap index, column n = g
                            if (d < 0 || d >= K) continue;
                            const uint8_t by = (uint8_t)digit(q[(size_t)d * tc + c], l);
                            regs[half] |= (uint32_t)by << (8 * i);
                        }
                    const size_t at = ((((size_t)b * tc + c) * nlt + l) * 32 + lane) * 2;
                    frag[at] = regs[0];
                    frag[at + 1] = regs[1];
                }
    const size_t bytes = frag.size() * sizeof(uint32_t);
    if (bytes > p.frag_capacity) {
        if (p.d_frag) cudaFree(p.d_frag);
        p.d_frag = nullptr; p.frag_capacity = 0;
        B200C_CUDA_TRY(cudaMalloc(&p.d_frag, bytes));
        p.frag_capacity = bytes;
    }
    B200C_CUDA_TRY(cudaMemcpy(p.d_frag, frag.data(), bytes, cudaMemcpyHostToDevice));
    p.K = K; p.NB = NB; p.nlt = nlt; p.dc = dtype == B200C_CI16 ? 2 : 1; p.tc = tc;
    p.ready = true;
    return B200C_OK;
}

void fir_imma_destroy(FirImmaPlan &p)
{
    if (p.d_frag) cudaFree(p.d_frag);
    p.d_frag = nullptr; p.frag_capacity = 0; p.ready = false;
}

template <int DC, int TC, int NLT>
static int launch_imma(const FirImmaArgs &a, size_t smem, int sm_count, cudaStream_t stream)
{
    constexpr int R = 2, MINB = NLT == 2 ? 4 : 3;
    auto kern = fir_imma_kernel<DC, TC, NLT, R, MINB>;
    static thread_local bool configured[16] = {false};
    int dev = 0;
    B200C_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 16 && !configured[dev]) {
        B200C_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[dev] = true;
    }
    static const int forced = [] { const char *e = std::getenv("B200C_IMMA_PERSM"); return e ? std::atoi(e) : 0; }();
    const long long per_sm = std::max<long long>(1, std::min<long long>(forced > 0 ? forced : MINB, (220 * 1024) / (long long)(smem + 1024)));
    const int grid = (int)std::min<long long>(a.ntiles, (long long)sm_count * per_sm);
    kern<<<grid, 32 * kImmaWarps, smem, stream>>>(a);
    B200C_CUDA_TRY(cudaGetLastError());
    return B200C_OK;
}

int fir_imma_launch(const FirImmaPlan &p, const void *d_in, size_t in_elems, void *d_out, size_t n_out, int sm_count,
                    cudaStream_t stream)
{
    if (n_out == 0) return B200C_OK;
    FirImmaArgs a;
    a.in = d_in; a.out = d_out; a.frag = static_cast<const uint2 *>(p.d_frag);
    a.n_in = (long long)in_elems; a.n_out = (long long)n_out;
    a.ntiles = ((long long)n_out + kImmaTile - 1) / kImmaTile;
    a.K = p.K; a.NB = p.NB; a.PL = kImmaTile + 32 * p.NB;
    const size_t smem = (size_t)p.NB * p.tc * p.nlt * 32 * sizeof(uint2) + 2 * ((size_t)p.dc * 2 * a.PL);   // fragments, planes, landing zone
    if (smem > 200 * 1024) { set_error("fir_imma: tap count too large for the tensor-core path"); return B200C_ERR_UNSUPPORTED; }
#define IMMA_NLT(DC, TC)                                                                 \
    (p.nlt == 2 ? launch_imma<DC, TC, 2>(a, smem, sm_count, stream)                        \
     : p.nlt == 3 ? launch_imma<DC, TC, 3>(a, smem, sm_count, stream)                      \
                  : launch_imma<DC, TC, 4>(a, smem, sm_count, stream))
    if (p.dc == 1) return IMMA_NLT(1, 1);
    if (p.tc == 1) return IMMA_NLT(2, 1);
    return IMMA_NLT(2, 2);
#undef IMMA_NLT
}

} // namespace b200c
