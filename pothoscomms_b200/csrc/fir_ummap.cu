// int16 / complex-int16 polyphase resampler (interpolation L <= 4, decimation M <= 4) on tcgen05.
//
// The reference nest (filter/FIRFilter.cpp:286-302) in polyphase form (fir.hpp): with the buffer
// index t = d_p + K-1 - k of tap k of output slot p (j_p, d_p as in fir.hpp), e = t mod M, a = t div M,
//   y[q L + p] = sum_k taps[j_p + k L] . plane_e[q + a],        plane_e[i] = buf[i M + e]
// i.e. per input residue e a Toeplitz GEMM over the de-interleaved stream plane_e whose columns
// are (u, p): output block q = 16 m + u of GEMM row m, slot p.  Same exact byte-limb algebra and
// the same descriptor-only Hankel operand as fir_umma.cu (K-major, no swizzle, SBO 128 B, LBO 16 B:
// row m = plane[16 m + 32 b ..]); as in fir_umma32.cu every (residue, component) plane of one data
// limb accumulates into ONE tensor-memory region through its own signed tap-digit matrix:
//   N = 16 u x L slots x (components x 2 digits),  regions: lo / hi data limb,
//   MMAs per tile of 2048 blocks q: 2 limbs x components x M residues x NB k-blocks.
// A GEMM row's columns are 16 L CONSECUTIVE outputs, so the epilogue stores them as they come.
// Two accumulator stages when 4 N <= 512 TMEM columns, one otherwise (L >= 3 complex).  A tile can also be
// computed as two column passes of N / 2 (the first / last eight output blocks of every row, the same A
// planes against half of each B tile), each pass its own unit of a two-stage MMA -> epilogue pipeline;
// measured slower (the MMA count doubles and an MMA costs about the same at half the width), so opt-in.
// Warp-specialised like fir_umma32_kernel: bulk-copy issuer -> stagers -> MMA issuer -> epilogue.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "fir_imma.hpp"
#include "umma.cuh"

namespace b200c {

struct FirUmmaPArgs {
    const void *in;
    void *out;
    const void *bmat;    // [M][DC][NB][N x 32 B] canonical K-major no-swizzle B tiles
    long long n_in;      // buffer elements available (history included); beyond: zeros
    long long nq;        // output blocks q to produce (outputs = nq * L)
    long long ntiles;
    int L, M, NB, N;     // N = 16 * L * DC * 2
    int PL, PLa;         // plane bytes in use (2048 + 32 NB) / allocated (multiple of 128)
    int R, nstage;       // landing ring depth; accumulator stages: 1 or 2
    int npass;           // column passes per tile: 1, or 2 halves of N / 2 columns
    int pstage;          // plane stages: 2 (staging overlaps the MMAs) unless shared memory is short
    long long *dbg;      // optional [grid][8] barrier-wait cycle counters of the roles (B200C_UMMA_DBG)
};

constexpr int kUPTile = 2048;      // blocks q per tile: 128 rows x 16
constexpr int kUPEpiWarps = 8, kUPStageWarps = 8, kUPMaxRing = 6;
constexpr int kUPThreads = 32 * (kUPEpiWarps + kUPStageWarps + 2);

template <int DC>
__global__ void __launch_bounds__(kUPThreads, 1) fir_ummap_kernel(const FirUmmaPArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_w[];
    constexpr int NQ = DC * 2, NPL = DC * 2, ESZ = DC * 2;
    constexpr int CH = 16 / NQ;                    // outputs per 16-column epilogue chunk
    const int L = a.L, M = a.M, NB = a.NB, N = a.N, PL = a.PL, PLa = a.PLa, R = a.R, NS = a.nstage;
    const int NPASS = a.npass, Nh = N / NPASS;     // columns of one pass
    const int COLS = 2 * Nh;                       // lo and hi regions of one stage
    const unsigned ALLOC = (unsigned)(NS * COLS) <= 32 ? 32 : (NS * COLS) <= 64 ? 64 : (NS * COLS) <= 128 ? 128 : (NS * COLS) <= 256 ? 256 : 512;
    const size_t bm_bytes = (size_t)M * DC * NB * N * 32, stage_bytes = (size_t)M * NPL * PLa, raw_bytes = (size_t)PL * M * ESZ;
    const int nmma_pass = 2 * M * DC * NB, nmma = NPASS * nmma_pass;  // MMAs of one column pass / of one tile
    unsigned char *bmat = smem_w;
    uint4 *mma_tab = reinterpret_cast<uint4 *>(bmat + bm_bytes);     // [nmma] issue list of one tile (see the MMA issuer)
    unsigned char *planes = reinterpret_cast<unsigned char *>(mma_tab) + (((size_t)nmma * 16 + 127) & ~(size_t)127);   // [NP][M][NPL][PLa]
    const int NP = a.pstage;
    unsigned char *raw = planes + NP * stage_bytes;                  // [R][PL * M * ESZ]
    __shared__ __align__(8) unsigned long long raw_full[kUPMaxRing], raw_empty[kUPMaxRing], planes_full[2], planes_empty[2], acc_full[2], acc_empty[2];
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (size_t i = tid; i < bm_bytes / 16; i += kUPThreads)
        reinterpret_cast<uint4 *>(bmat)[i] = __ldg(static_cast<const uint4 *>(a.bmat) + i);
    // The MMAs of a tile are the same list for every tile up to the plane stage (A start address) and the
    // accumulator stage (D column): one 16-byte entry per MMA, in issue order (pass, limb, residue, component,
    // k-block) -- the issuing thread then spends one shared-memory load and two adds per MMA instead of
    // rebuilding descriptors (which, not the tensor core, set the pace: ~350 cycles per operand plane).
    //   .x/.y = B descriptor, .z = A start-address offset from the plane stage (16-byte units),
    //   .w = D column offset within the stage | hi limb << 16 | first MMA of its accumulator << 17
    {
        const unsigned long long b_base = umma_smem_desc(smem_u32(bmat), 128, 256, 0);
        const unsigned long long kBStep = (unsigned long long)((N * 32) >> 4);
        for (int j = tid; j < nmma; j += kUPThreads) {
            int t = j;
            const int b = t % NB; t /= NB;
            const int dc = t % DC; t /= DC;
            const int e = t % M; t /= M;
            const int dl = t & 1, h = t >> 1;
            const unsigned long long bd = b_base + (unsigned long long)((e * DC + dc) * NB + b) * kBStep + (unsigned long long)(h * ((Nh * 32) >> 4));
            const unsigned a_off = (unsigned)((((e * NPL) + 2 * dc + dl) * PLa + 32 * b) >> 4);
            const unsigned meta = (unsigned)(dl * Nh) | ((unsigned)dl << 16) | ((e == 0 && dc == 0 && b == 0) ? 1u << 17 : 0u);
            mma_tab[j] = make_uint4((unsigned)bd, (unsigned)(bd >> 32), a_off, meta);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        for (int r = 0; r < R; r++) { mbar_init(&raw_full[r], 1); mbar_init(&raw_empty[r], 32 * kUPStageWarps); }
        for (int s = 0; s < 2; s++) {
            mbar_init(&planes_full[s], 32 * kUPStageWarps); mbar_init(&planes_empty[s], 1);
            mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 32 * kUPEpiWarps);
        }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ALLOC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;
    const long long first = blockIdx.x, step = gridDim.x;
    const int ntl = first < a.ntiles ? (int)((a.ntiles - first + step - 1) / step) : 0;
    const bool al = (reinterpret_cast<unsigned long long>(a.in) & 15) == 0;
    // the tile's window: buffer elements [q0 M, (q0 + PL) M)
    auto bulk_ok = [&](long long tile) { return al && (tile * kUPTile + PL) * M <= a.n_in; };
    // accumulator stage / phase of the i-th tile of this CTA for an NS-deep ring; the planes are their
    // own ring (2-deep when shared memory allows) so that staging tile i+1 overlaps the MMAs of tile i
    long long w0 = 0, w1 = 0;
    const long long t_begin = clock64();
    // (the unit of that ring is a column pass: u = tile index * NPASS + pass)
    auto stage_of = [&](int u, int &s, unsigned &ph) { s = NS == 2 ? (u & 1) : 0; ph = (unsigned)(NS == 2 ? (u >> 1) : u) & 1; };

    if (warp == kUPEpiWarps + kUPStageWarps + 1) {
        // ================================================================ bulk-copy issuer
        if (lane == 0)
            for (int i = 0, r = 0, ph = 0; i < ntl; i++) {
                const long long tile = first + (long long)i * step;
                timed_wait(&raw_empty[r], (unsigned)ph ^ 1, w0);
                if (bulk_ok(tile))
                    bulk_load(raw + (size_t)r * raw_bytes, static_cast<const unsigned char *>(a.in) + (size_t)tile * kUPTile * M * ESZ,
                              (unsigned)raw_bytes, &raw_full[r]);
                else
                    mbar_arrive(&raw_full[r]);
                if (++r == R) { r = 0; ph ^= 1; }
            }
    } else if (warp == kUPEpiWarps + kUPStageWarps) {
        // ====================================================================== MMA issuer
        if (elect_one()) {
            const unsigned idesc_lo = umma_idesc_i8(false, Nh), idesc_hi = umma_idesc_i8(true, Nh);
            // A: plane (e, dc, dl): row m = plane[16 m + 32 b ..] (SBO 128, LBO 16), start address = stage + table offset
            const unsigned long long a_stage0 = umma_smem_desc(smem_u32(planes), 16, 128, 0);
            const unsigned long long a_stage1 = umma_smem_desc(smem_u32(planes + stage_bytes), 16, 128, 0);
            for (int i = 0; i < ntl; i++) {
                const int sp = NP == 2 ? (i & 1) : 0;
                const unsigned php = (unsigned)(NP == 2 ? (i >> 1) : i) & 1;
                timed_wait(&planes_full[sp], php, w0);
                const unsigned long long a_stage = sp ? a_stage1 : a_stage0;
                const uint4 *tab = mma_tab;
                for (int h = 0; h < NPASS; h++) {                          // column pass: columns [h Nh, (h + 1) Nh) of every B tile
                    int s; unsigned ph;
                    stage_of(i * NPASS + h, s, ph);
                    timed_wait(&acc_empty[s], ph ^ 1, w1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const unsigned d_stage = tmem_base + (unsigned)(s * COLS);
                    uint4 cur = tab[0];
#pragma unroll 4
                    for (int k = 0; k < nmma_pass; k++) {
                        const uint4 nxt = tab[k + 1 < nmma_pass ? k + 1 : k];   // loaded ahead of the MMA it follows
                        const unsigned long long bd = ((unsigned long long)cur.y << 32) | cur.x;
                        umma_i8(d_stage + (cur.w & 0xffffu), a_stage + cur.z, bd, (cur.w & 0x10000u) ? idesc_hi : idesc_lo, !(cur.w & 0x20000u));
                        cur = nxt;
                    }
                    tab += nmma_pass;
                    if (h == NPASS - 1)
                        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&planes_empty[sp])) : "memory");
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&acc_full[s])) : "memory");
                }
            }
        }
    } else if (warp >= kUPEpiWarps) {
        // ========================================================================= stagers
        // one thread = 4 consecutive positions i of every residue plane = 4 M consecutive buffer elements
        const int st = tid - 32 * kUPEpiWarps;
        constexpr int NST = 32 * kUPStageWarps;
        const int nchunk = PL / 4;
        for (int i = 0, r = 0, rph = 0; i < ntl; i++) {
            const int s = NP == 2 ? (i & 1) : 0;
            const unsigned ph = (unsigned)(NP == 2 ? (i >> 1) : i) & 1;
            const long long tile = first + (long long)i * step;
            const long long e0 = tile * kUPTile * M;                  // first buffer element of the tile's window
            const bool landed = bulk_ok(tile);
            timed_wait(&raw_full[r], (unsigned)rph, w0);
            timed_wait(&planes_empty[s], ph ^ 1, w1);
            unsigned *pl = reinterpret_cast<unsigned *>(planes + (size_t)s * stage_bytes);
            const unsigned char *rw = raw + (size_t)r * raw_bytes;
            for (int c = st; c < nchunk; c += NST) {
                // element j of the chunk (j < 4 M): buffer element e0 + 4 c M + j, residue j % M, position 4 c + j / M
                unsigned x[16];                                       // complex: one word per element; real: two elements per word
                constexpr int EPW = DC == 2 ? 1 : 2;                  // elements per 32-bit word
                const int nwords = 4 * M / EPW;
                if (landed) {
                    const unsigned *src = reinterpret_cast<const unsigned *>(rw) + (size_t)c * nwords;
                    if ((nwords & 3) == 0) {   // 128-bit loads: word loads at this 4 nwords-byte lane stride are 8-way bank conflicts
#pragma unroll
                        for (int j = 0; j < 4; j++) {
                            const uint4 v = 4 * j < nwords ? reinterpret_cast<const uint4 *>(src)[j] : make_uint4(0, 0, 0, 0);
                            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; j++) x[j] = j < nwords ? src[j] : 0u;
                    }
                } else {
                    const long long g0 = e0 + 4LL * c * M;
                    if constexpr (DC == 2) {
                        const unsigned *__restrict__ in32 = static_cast<const unsigned *>(a.in);
#pragma unroll
                        for (int j = 0; j < 16; j++) x[j] = (j < nwords && g0 + j < a.n_in) ? __ldg(in32 + g0 + j) : 0u;
                    } else {
                        const unsigned short *__restrict__ in16 = static_cast<const unsigned short *>(a.in);
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            const unsigned h0 = (2 * j < 4 * M && g0 + 2 * j < a.n_in) ? __ldg(in16 + g0 + 2 * j) : 0u;
                            const unsigned h1 = (2 * j + 1 < 4 * M && g0 + 2 * j + 1 < a.n_in) ? __ldg(in16 + g0 + 2 * j + 1) : 0u;
                            x[j] = h0 | (h1 << 16);
                        }
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    if (e >= M) break;
                    unsigned *p = pl + (size_t)(e * NPL) * (PLa / 4) + c;
                    if constexpr (DC == 2) {
                        // samples e, e + M, e + 2M, e + 3M (dynamic M: select among the unrolled candidates)
                        unsigned s4[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            unsigned v = 0;
#pragma unroll
                            for (int m = 1; m <= 4; m++) if (M == m) v = x[(e + m * k) & 15];
                            s4[k] = v;
                        }
                        const unsigned t01 = prmt_u(s4[0], s4[1], 0x5140), t23 = prmt_u(s4[2], s4[3], 0x5140);
                        const unsigned u01 = prmt_u(s4[0], s4[1], 0x7362), u23 = prmt_u(s4[2], s4[3], 0x7362);
                        p[0] = prmt_u(t01, t23, 0x5410);
                        p[PLa / 4] = prmt_u(t01, t23, 0x7632);
                        p[2 * (PLa / 4)] = prmt_u(u01, u23, 0x5410);
                        p[3 * (PLa / 4)] = prmt_u(u01, u23, 0x7632);
                    } else {
                        unsigned h4[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            unsigned v = 0;
#pragma unroll
                            for (int m = 1; m <= 4; m++)
                                if (M == m) { const int j = e + m * k; v = (x[(j >> 1) & 15] >> (16 * (j & 1))) & 0xffffu; }
                            h4[k] = v;
                        }
                        const unsigned w01 = h4[0] | (h4[1] << 16), w23 = h4[2] | (h4[3] << 16);
                        p[0] = prmt_u(w01, w23, 0x6420);
                        p[PLa / 4] = prmt_u(w01, w23, 0x7531);
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&planes_full[s]);
            mbar_arrive(&raw_empty[r]);
            if (++r == R) { r = 0; rph ^= 1; }
        }
    } else {
        // ======================================================================== epilogue
        // thread = (row m = TMEM lane, half of the row's N / 16 column chunks); a chunk is CH consecutive
        // outputs of the row's 16 L:  y = lo_d0 + ((lo_d1 + hi_d0) << 8) + (hi_d1 << 16) mod 2^32, bits [16, 32)
        const int m = 32 * (warp & 3) + lane, half = warp >> 2;
        const unsigned lane_addr = tmem_base + ((unsigned)(32 * (warp & 3)) << 16);
        const int nch = Nh / 16, c0 = half * (nch / 2), c1 = c0 + nch / 2;
        const long long n_out = a.nq * L;
        for (int u = 0; u < ntl * NPASS; u++) {
            int s; unsigned ph;
            stage_of(u, s, ph);
            const int i = u / NPASS, h = u - i * NPASS;
            // pass h holds output blocks [8 h, 8 h + 8) of the row when NPASS == 2: CH * nch outputs further on
            const long long tile = first + (long long)i * step, orow = (tile * kUPTile + 16LL * m) * L + (long long)h * CH * nch;
            timed_wait(&acc_full[s], ph, w0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // chunk c + 1 is loaded from tensor memory while chunk c is combined and stored (two register sets)
            auto load = [&](int c, unsigned (&lo)[16], unsigned (&hi)[16]) {
                tmem_ld16(lane_addr + (unsigned)(s * COLS + 16 * c), lo);
                tmem_ld16(lane_addr + (unsigned)(s * COLS + Nh + 16 * c), hi);
            };
            auto landed_all = [&] {   // every tensor-memory read of this unit is done: the stage can be overwritten
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                mbar_arrive(&acc_empty[s]);
            };
            auto emit = [&](int c, const unsigned (&lo)[16], const unsigned (&hi)[16]) {
                unsigned res[CH];
#pragma unroll
                for (int k = 0; k < CH; k++) {
                    unsigned y[DC];
#pragma unroll
                    for (int cls = 0; cls < DC; cls++)
                        y[cls] = lo[k * NQ + 2 * cls] + ((lo[k * NQ + 2 * cls + 1] + hi[k * NQ + 2 * cls]) << 8) + (hi[k * NQ + 2 * cls + 1] << 16);
                    if constexpr (DC == 2) res[k] = prmt_u(y[0], y[DC - 1], 0x7632);
                    else res[k] = y[0] >> 16;
                }
                const long long o = orow + (long long)CH * c;
                if constexpr (DC == 2) {
                    unsigned *out32 = static_cast<unsigned *>(a.out);
                    if (o + CH <= n_out && ((reinterpret_cast<unsigned long long>(out32) + 4 * (unsigned long long)o) & 15) == 0) {
                        __stcg(reinterpret_cast<uint4 *>(out32 + o), make_uint4(res[0], res[1], res[2], res[3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < CH; k++)
                            if (o + k < n_out) out32[o + k] = res[k];
                    }
                } else {
                    unsigned short *out16 = static_cast<unsigned short *>(a.out);
                    if (o + CH <= n_out && ((reinterpret_cast<unsigned long long>(out16) + 2 * (unsigned long long)o) & 15) == 0) {
                        __stcg(reinterpret_cast<uint4 *>(out16 + o), make_uint4(res[0] | (res[1] << 16), res[2] | (res[3] << 16),
                                                                               res[4 % CH] | (res[5 % CH] << 16), res[6 % CH] | (res[7 % CH] << 16)));
                    } else {
#pragma unroll
                        for (int k = 0; k < CH; k++)
                            if (o + k < n_out) out16[o + k] = (unsigned short)res[k];
                    }
                }
            };
            unsigned loA[16], hiA[16], loB[16], hiB[16];
            load(c0, loA, hiA);
#pragma unroll 1
            for (int c = c0; c < c1; c += 2) {
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c + 1 < c1) load(c + 1, loB, hiB); else landed_all();
                emit(c, loA, hiA);
                if (c + 1 < c1) {
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    if (c + 2 < c1) load(c + 2, loA, hiA); else landed_all();
                    emit(c + 1, loB, hiB);
                }
            }
        }
    }
    if (a.dbg && lane == 0) {
        long long *d = a.dbg + (size_t)blockIdx.x * 8;
        if (warp == 0) { d[0] = clock64() - t_begin; d[1] = ntl; d[7] = w0; }
        if (warp == kUPEpiWarps) { d[5] = w0; d[6] = w1; }
        if (warp == kUPEpiWarps + kUPStageWarps) { d[3] = w0; d[4] = w1; }
        if (warp == kUPEpiWarps + kUPStageWarps + 1) d[2] = w0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ALLOC) : "memory");
}

// ------------------------------------------------------------------------------- host ---
static bool two_digits_p(long long q, int8_t &d0, int8_t &d1)
{
    const long long lo = ((q + 128) & 255) - 128, hi = (q - lo) >> 8;
    if (hi < -128 || hi > 127) return false;
    d0 = (int8_t)lo; d1 = (int8_t)hi;
    return true;
}

int fir_ummap_configure(FirUmmaPPlan &p, int dtype, const double *taps, size_t ntaps, bool complex_taps, size_t M, size_t L, bool force)
{
    p.ready = false;
    const bool enabled = [] { const char *e = std::getenv("B200C_UMMAP"); return !e || std::atoi(e) != 0; }();
    if ((dtype != B200C_I16 && dtype != B200C_CI16) || !(enabled || force)) return B200C_OK;
    if ((M == 1 && L == 1) || M > 4 || L > 4 || ntaps < 2) return B200C_OK;
    const int dc = dtype == B200C_CI16 ? 2 : 1, tc = complex_taps ? 2 : 1, NQ = dc * 2;
    const int N = 16 * (int)L * NQ;
    const long long K = (long long)((ntaps + L - 1) / L);               // filter/FIRFilter.cpp:335
    // buffer offsets t = d_p + K-1 - k >= 0; window positions a = t / M
    long long t_max = 0;
    for (size_t ps = 0; ps < L; ps++) t_max = std::max(t_max, (long long)(((ps + 1) * M - 1) / L) + K - 1);
    const long long a_max = t_max / (long long)M;
    const int NB = (int)((15 + a_max + 1 + 31) / 32);
    const int PL = kUPTile + 32 * NB, PLa = (PL + 127) / 128 * 128;
    // Two accumulator stages of (lo, hi) regions fit the 512 tensor-memory columns when 4 N <= 512.  Beyond
    // (L >= 3 complex) the tile keeps ONE stage: splitting it into two column passes of N / 2 (opt-in,
    // B200C_UMMAP_NPASS=2) does overlap MMAs and epilogue but doubles the MMA count, and an MMA costs
    // ~105 + 125 per new operand plane cycles almost independently of N: measured 199 against 256 Gsamples/s.
    static const bool two_pass = [] { const char *e = std::getenv("B200C_UMMAP_NPASS"); return e && std::atoi(e) == 2; }();
    const int npass = (4 * N > 512 && two_pass) ? 2 : 1;
    const int nstage = 4 * (N / npass) <= 512 ? 2 : 1;
    const size_t tab_bytes = ((size_t)npass * 2 * M * dc * NB * 16 + 127) & ~(size_t)127;            // the MMA issue list
    const size_t bm_bytes = (size_t)M * dc * NB * N * 32 + tab_bytes, stage = (size_t)M * dc * 2 * PLa, one = (size_t)PL * M * dc * 2;
    // tables + planes (two stages if they fit) + a landing ring of at least two slots
    const int pstage = bm_bytes + 2 * stage + 3 * one + 1024 <= 210 * 1024 ? 2 : 1;
    if (bm_bytes + pstage * stage + 2 * one + 1024 > 210 * 1024) return B200C_OK;
    std::vector<uint8_t> bm(bm_bytes - tab_bytes, 0);
    for (size_t ps = 0; ps < L; ps++) {
        const long long ii = (long long)((ps + 1) * M - 1), jp = ii % (long long)L, dp = ii / (long long)L;
        for (long long k = 0; jp + k * (long long)L < (long long)ntaps; k++) {
            const size_t ti = (size_t)(jp + k * (long long)L);
            const long long t = dp + K - 1 - k, e = t % (long long)M, aa = t / (long long)M;
            for (int c = 0; c < tc; c++) {
                const long long q = (long long)(int32_t)(long long)std::ldexp(taps[ti * tc + c], 16);
                int8_t pos[2], neg[2];
                if (!two_digits_p(q, pos[0], pos[1]) || !two_digits_p(-q, neg[0], neg[1])) return B200C_OK;
                struct Use { int dcx, cls; bool negate; };
                Use uses[2];
                int nuse = 0;
                if (dc == 1) { uses[nuse++] = {0, 0, false}; }
                else if (c == 0) { uses[nuse++] = {0, 0, false}; uses[nuse++] = {1, 1, false}; }
                else { uses[nuse++] = {0, 1, false}; uses[nuse++] = {1, 0, true}; }
                for (int u = 0; u < nuse; u++)
                    for (int uu = 0; uu < 16; uu++) {
                        const long long j = uu + aa;                         // window position of this tap for block q = 16 m + uu
                        const int b = (int)(j / 32), jj = (int)(j % 32);
                        for (int l = 0; l < 2; l++) {
                            const int col = ((uu * (int)L + (int)ps) * dc + uses[u].cls) * 2 + l;
                            const size_t at = ((size_t)(e * dc + uses[u].dcx) * NB + b) * N * 32 + (size_t)(col / 8) * 256 + (size_t)(jj / 16) * 128 +
                                              (size_t)(col % 8) * 16 + (jj % 16);
                            bm[at] = (uint8_t)(uses[u].negate ? neg[l] : pos[l]);
                        }
                    }
            }
        }
    }
    if (bm.size() > p.capacity) {
        if (p.d_bmat) cudaFree(p.d_bmat);
        p.d_bmat = nullptr; p.capacity = 0;
        B200C_CUDA_TRY(cudaMalloc(&p.d_bmat, bm.size()));
        p.capacity = bm.size();
    }
    B200C_CUDA_TRY(cudaMemcpy(p.d_bmat, bm.data(), bm.size(), cudaMemcpyHostToDevice));
    p.L = (int)L; p.M = (int)M; p.NB = NB; p.N = N; p.dc = dc; p.nstage = nstage; p.pstage = pstage; p.npass = npass;
    p.ready = true;
    return B200C_OK;
}

void fir_ummap_destroy(FirUmmaPPlan &p)
{
    if (p.d_bmat) cudaFree(p.d_bmat);
    p.d_bmat = nullptr; p.capacity = 0; p.ready = false;
}

template <int DC>
static int launch_up(FirUmmaPArgs a, int sm_count, cudaStream_t stream)
{
    auto kern = fir_ummap_kernel<DC>;
    static thread_local bool configured[16] = {false};
    int dev = 0;
    B200C_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 16 && !configured[dev]) {
        B200C_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured[dev] = true;
    }
    const size_t tab_bytes = ((size_t)a.npass * 2 * a.M * DC * a.NB * 16 + 127) & ~(size_t)127;
    const size_t fixed = (size_t)a.M * DC * a.NB * a.N * 32 + tab_bytes + (size_t)a.pstage * a.M * DC * 2 * a.PLa + 1024, one = (size_t)a.PL * a.M * DC * 2;
    a.R = (int)std::max<size_t>(2, std::min<size_t>(kUPMaxRing, (216 * 1024 - fixed) / one));
    const size_t smem = std::max<size_t>(fixed + a.R * one, 116 * 1024);   // > half an SM: one CTA per SM (tensor memory)
    const int grid = (int)std::min<long long>(a.ntiles, (long long)sm_count);
    static const bool dbg = std::getenv("B200C_UMMA_DBG") != nullptr;
    if (dbg) {
        B200C_CUDA_TRY(cudaMalloc(&a.dbg, (size_t)grid * 8 * sizeof(long long)));
        B200C_CUDA_TRY(cudaMemset(a.dbg, 0, (size_t)grid * 8 * sizeof(long long)));
    }
    kern<<<grid, kUPThreads, smem, stream>>>(a);
    B200C_CUDA_TRY(cudaGetLastError());
    if (dbg) {
        std::vector<long long> h((size_t)grid * 8);
        B200C_CUDA_TRY(cudaStreamSynchronize(stream));
        B200C_CUDA_TRY(cudaMemcpy(h.data(), a.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(a.dbg);
        double s8[8] = {0};
        for (int g = 0; g < grid; g++) for (int k = 0; k < 8; k++) s8[k] += (double)h[(size_t)g * 8 + k] / grid;
        std::fprintf(stderr, "ummap: cycles/tile %.0f | issuer wait raw_empty %.0f | mma wait planes_full %.0f acc_empty %.0f | stager wait raw_full %.0f planes_empty %.0f | epilogue wait acc_full %.0f (tiles/CTA %.1f, R %d, N %d, NB %d, acc stages %d, plane stages %d, column passes %d)\n",
                     s8[0] / s8[1], s8[2] / s8[1], s8[3] / s8[1], s8[4] / s8[1], s8[5] / s8[1], s8[6] / s8[1], s8[7] / s8[1], s8[1], a.R, a.N, a.NB, a.nstage, a.pstage, a.npass);
    }
    return B200C_OK;
}

int fir_ummap_launch(const FirUmmaPPlan &p, const void *d_in, size_t in_elems, void *d_out, size_t nq, int sm_count, cudaStream_t stream)
{
    if (nq == 0) return B200C_OK;
    FirUmmaPArgs a;
    a.in = d_in; a.out = d_out; a.bmat = p.d_bmat;
    a.n_in = (long long)in_elems; a.nq = (long long)nq;
    a.ntiles = ((long long)nq + kUPTile - 1) / kUPTile;
    a.L = p.L; a.M = p.M; a.NB = p.NB; a.N = p.N;
    a.PL = kUPTile + 32 * p.NB; a.PLa = (a.PL + 127) / 128 * 128; a.R = 2; a.nstage = p.nstage; a.pstage = p.pstage; a.npass = p.npass; a.dbg = nullptr;
    return p.dc == 1 ? launch_up<1>(a, sm_count, stream) : launch_up<2>(a, sm_count, stream);
}

} // namespace b200c
