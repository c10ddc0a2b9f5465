// int16 / complex-int16 FIR on tcgen05, second formulation: 32 outputs per GEMM row.
//
// fir_umma.cu aliases the Hankel data matrix onto the byte planes with 16-byte row pitch (no
// swizzle), which ties N to 16 outputs x 4 tap planes = 64 columns per MMA; measured there
// (DESIGN.md): every M = 128 MMA costs ~67 cycles of A-operand fetch whatever N is.  This kernel
// doubles the work per fetched A tile:
//  * the planes are stored with the 32-byte swizzle (byte offset o lives at o ^ ((o >> 7 & 1) << 4)),
//    so a K-major SWIZZLE_32B descriptor sees rows of 32-byte pitch: row m is the window
//    plane[32 m + 32 b + (0..31)], advanced by ONE ROW per k-block b.  That start address is not
//    aligned to the 256-byte swizzle pattern; tools/probe_umma_swizzle.cu shows the tensor core
//    swizzles on the absolute shared-memory address (base_offset 0), which makes it exact.
//  * 32 outputs per row, and the re and im data planes ACCUMULATE INTO THE SAME tensor-memory
//    region through different B matrices: B_re = [h_re digits -> y_re columns | h_im digits ->
//    y_im columns], B_im = [-h_im digits | h_re digits].  Only the data limb (lo / hi byte, weights
//    2^0 and 2^8) needs separate regions: N = 32 outputs x 2 components x 2 tap digits = 128 columns,
//    2 regions, 8 accumulators per output instead of 16; two stages fill the 512 TMEM columns.
//  * NB = ceil((K + 31) / 32) k-blocks; a tile is 128 rows x 32 = 4096 outputs.
// Warp-specialised: bulk-copy issuer -> stagers -> MMA issuer -> epilogue, connected by mbarrier rings.
//
// Three kernels share the body below:
//   fir_umma32_kernel<DC>      the formulation above (data planes = A operand, tap tiles = B operand in shared memory):
//                              shared-memory-bandwidth bound (8 KB of operands per MMA); what stays on it: filters beyond
//                              the tap counts of the next two
//   fir_umma32t_kernel<NB>     operands swapped, complex data: tap tiles = A operand resident in TENSOR MEMORY, data planes =
//                              B operand, N = 96 windows of 32 outputs; tensor-pipe bound
//   fir_umma32tr_kernel<NB>    the same on real data: windows of 64 outputs over 64-byte-swizzled planes; HBM bound
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

#include "fir_imma.hpp"
#include "umma.cuh"

namespace b200c {

struct FirUmma32Args {
    const void *in;
    void *out;
    const void *bmat;    // [DC][NB][N x 32 B] canonical K-major no-swizzle B tiles; swapped kernels: [DC][NB][128 rows x 32 B] A tiles
    long long n_in, n_out, ntiles;
    int K, NB, PL, PLa;  // PL: bytes per plane in use = 4096 + 32 NB; PLa: allocated (multiple of 256)
    int R;               // depth of the bulk-copy landing ring
    long long *dbg;
};

constexpr int kU32Tile = 4096;
// operand-swapped variant (TS = true), complex int16: the tap tiles are the A operand and live in
// TENSOR MEMORY, the data planes are the B operand with N = NW windows per tile.  Columns: two stages x
// two data limbs x NW accumulators + 8 per (data component, k-block) tap tile <= 512.
constexpr int kU32tNW = 96, kU32tTile = 32 * kU32tNW, kU32tMaxNB = (512 - 4 * kU32tNW) / 16;
// The same for REAL int16 data: one component leaves half of the 128 rows free, so a window is 64 outputs (rows = 64 outputs
// x 2 tap digits) and the plane rows are 64 bytes (SWIZZLE_64B), read 32 bytes -- one k-block -- at a time; a tile is
// 96 x 64 = 6144 outputs, still 3072 32-bit words of output.
constexpr int kU32trTile = 64 * kU32tNW;
constexpr int kU32trMaxNB = 16;    // one data component: 8 columns of tap tiles per k-block, 128 columns beside the accumulators
constexpr int kU32EpiWarps = 8, kU32StageWarps = 8, kU32MaxRing = 8;
constexpr int kU32Batch = 5;      // stager loads in flight per thread: their latency under the MMA's operand traffic is long
constexpr int kU32Threads = 32 * (kU32EpiWarps + kU32StageWarps + 2);
// the swapped kernel is bound by instruction issue (ncu r02ad: 1.3 warp instructions per sample, 47 % of them the
// stagers' -- mostly per-tile overhead of eight warps converting three items each): compile-time sizes, and six
// stager warps.  Measured at C2 (profiles/r02ah): 4 / 6 / 8 stager warps and loading all of a tile's accumulators at once
// instead of in three chunks all land within 0.70-0.71 of the HBM roofline: the tile time (~1270 cycles) is the 960 cycles
// of the 20 MMAs (8192 MAC / clk, tools/probe_umma_rate.cu) plus the barrier hand-offs around them
constexpr int kU32tStageWarps = 6;
constexpr int kU32tPlaneStages = 3;   // the stagers run one tile further ahead of the MMAs than the two accumulator stages allow
constexpr int kU32tThreads = 32 * (kU32EpiWarps + kU32tStageWarps + 2);

template <int DC, bool TS, int NBT, int SWT = kU32StageWarps>   // NBT: the k-block count at compile time (0: a.NB)
__device__ __forceinline__ void fir_umma32_body(const FirUmma32Args &a)
{
    extern __shared__ __align__(1024) unsigned char smem_v[];
    constexpr int NQ = DC * 2;                     // (output component, tap digit) per output
    constexpr int N = 32 * NQ;                     // MMA N = columns of one data-limb region
    constexpr int NW = kU32tNW;                    // TS: windows per tile = MMA N
    constexpr bool TR = TS && DC == 1;             // swapped formulation on real data
    constexpr int TILE = TR ? kU32trTile : TS ? kU32tTile : kU32Tile;
    constexpr int COLS = TS ? 2 * NW : 2 * N;      // lo and hi regions
    constexpr int ALLOC = TS ? 512 : 2 * COLS;     // two stages: 512 (complex) / 256 (real) columns
    constexpr int ACOL = 2 * COLS;                 // TS: first column of the tap tiles
    constexpr int SW = SWT, NTHR = 32 * (kU32EpiWarps + SW + 2);
    constexpr int PS = TS ? kU32tPlaneStages : 2;  // plane stages (the accumulators have two)
    constexpr int NPL = DC * 2, ESZ = DC * 2;
    constexpr int CH = 16 / NQ;                    // outputs per 16-column epilogue chunk
    const int NB = NBT ? NBT : a.NB, PL = a.PL, PLa = a.PLa, R = a.R;
    unsigned char *bmat = smem_v;                                    // DC * NB * N * 32 bytes (multiple of 1024); TS: none
    unsigned char *planes = bmat + (TS ? 0 : (size_t)DC * NB * N * 32);   // [2][NPL][PLa], 256-byte aligned, swizzled
    unsigned char *raw = planes + PS * (size_t)NPL * PLa;            // [R][PL * ESZ (+ 128: swapped kernels, misaligned streams)]
    __shared__ __align__(8) unsigned long long raw_full[kU32MaxRing], raw_empty[kU32MaxRing], planes_full[4], planes_empty[4], acc_full[2], acc_empty[2];
    static_assert(PS <= 4, "planes_full / planes_empty hold four stages");
    __shared__ unsigned tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if constexpr (!TS) {
        for (int i = tid; i < DC * NB * N * 2; i += NTHR)
            reinterpret_cast<uint4 *>(bmat)[i] = __ldg(static_cast<const uint4 *>(a.bmat) + i);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (tid == 0) {
        // every stager / epilogue thread waits and arrives itself: one lane per warp followed by a warp barrier measured
        // 2x slower (profiles/r02ab: the stagers' proxy fence + arrival went from ~80 to ~1800 cycles per tile)
        for (int r = 0; r < R; r++) { mbar_init(&raw_full[r], 1); mbar_init(&raw_empty[r], 32 * SW); }
        for (int s = 0; s < PS; s++) { mbar_init(&planes_full[s], 32 * SW); mbar_init(&planes_empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 32 * kU32EpiWarps); }
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(ALLOC) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem_base = tmem_base_s;
    if constexpr (TS) {
        // tap tiles -> tensor memory: row m' (TMEM lane) of tile t is 32 bytes = 8 columns (tools/probe_umma_tmem_a.cu)
        if (warp < 4) {
            const unsigned at = tmem_base + ((unsigned)(32 * warp) << 16) + ACOL;
            const uint4 *am = static_cast<const uint4 *>(a.bmat) + (32 * warp + lane) * 2;
            for (int t = 0; t < DC * NB; t++) {
                const uint4 v0 = __ldg(am + t * 256), v1 = __ldg(am + t * 256 + 1);
                asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(at + 8u * t), "r"(v0.x),
                             "r"(v0.y), "r"(v0.z), "r"(v0.w), "r"(v1.x), "r"(v1.y), "r"(v1.z), "r"(v1.w)
                             : "memory");
            }
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    const long long first = blockIdx.x, step = gridDim.x;
    const int ntl = first < a.ntiles ? (int)((a.ntiles - first + step - 1) / step) : 0;
    // A stream that does not start on a 16-byte boundary (the block layer's read pointer after any work() call: it has
    // advanced by n - (K - 1) elements) is still fetched by bulk copies in the swapped kernels: from the aligned address below,
    // `mis` bytes early and one 16-byte unit longer, and the stagers shift the bytes back (every tile starts a multiple of
    // 16 bytes after the first, so `mis` is one number per launch).  The original kernel takes guarded element loads.
    const int mis = (int)(reinterpret_cast<unsigned long long>(a.in) & 15);
    const bool al = mis == 0;
    const int slot_bytes = PL * ESZ + (TS ? 128 : 0);      // one more 16-byte unit, kept on the 128-byte grid of the slots
    auto bulk_ok = [&](long long tile) {
        if (TS) return (tile * TILE + PL) * ESZ + (mis ? 16 - mis : 0) <= a.n_in * ESZ;
        return al && tile * TILE + PL <= a.n_in;
    };
    long long w0 = 0, w1 = 0, t_conv = 0, t_fence = 0;
    const long long t_begin = clock64();
    if (g_umma_watch && blockIdx.x == 0 && tid == 0) {     // barrier addresses, to decode the watchdog's records
        unsigned long long *w = g_umma_watch + 32 * gridDim.x;
        w[0] = smem_u32(raw_full); w[1] = smem_u32(raw_empty); w[2] = smem_u32(planes_full); w[3] = smem_u32(planes_empty);
        w[4] = smem_u32(acc_full); w[5] = smem_u32(acc_empty);
    }

    if (warp == kU32EpiWarps + SW + 1) {
        // ================================================================ bulk-copy issuer
        if (lane == 0)
            for (int i = 0, r = 0, ph = 0; i < ntl; i++) {
                const long long tile = first + (long long)i * step;
                watched_wait(a.dbg != nullptr, &raw_empty[r], (unsigned)ph ^ 1, w0);
                if (bulk_ok(tile))
                    bulk_load(raw + (size_t)r * slot_bytes, static_cast<const unsigned char *>(a.in) + (size_t)tile * TILE * ESZ - mis,
                              (unsigned)(PL * ESZ + (mis ? 16 : 0)), &raw_full[r]);
                else
                    mbar_arrive(&raw_full[r]);
                if (++r == R) { r = 0; ph ^= 1; }
            }
    } else if (warp == kU32EpiWarps + SW) {
        // ====================================================================== MMA issuer
        if (elect_one()) {
            // One thread feeds the tensor core: everything per MMA beyond two 64-bit adds is hoisted out
            // of the loop (with descriptors rebuilt per MMA the issue rate, not the MMA, set the pace).
            unsigned long long a_base[2][NPL], b_base[DC];
#pragma unroll
            for (int s = 0; s < 2; s++)
#pragma unroll
                for (int p = 0; p < NPL; p++)      // plane p of stage s, rows of 32-byte pitch (SWIZZLE_32B); TR: 64-byte pitch (SWIZZLE_64B)
                    a_base[s][p] = TR ? umma_smem_desc(smem_u32(planes + ((size_t)s * NPL + p) * PLa), 16, 512, 4)
                                      : umma_smem_desc(smem_u32(planes + ((size_t)s * NPL + p) * PLa), 16, 256, 6);
            if constexpr (!TS) {
#pragma unroll
                for (int dc = 0; dc < DC; dc++) b_base[dc] = umma_smem_desc(smem_u32(bmat + (size_t)dc * NB * N * 32), 128, 256, 0);
            }
            constexpr unsigned long long kAStep = 32 >> 4, kBStep = (N * 32) >> 4;   // start-address field is in 16-byte units
            // the tile loop is unrolled over the two stages so that every index below is a compile-time
            // constant (dynamically indexed descriptor arrays would live in local memory)
            // ps / pph: plane stage and its phase (== S, ph with two plane stages)
            auto issue_tile = [&](auto stage_c, const unsigned ph, const int ps, const unsigned pph) {
                constexpr int S = decltype(stage_c)::value;
                watched_wait(a.dbg != nullptr, &planes_full[ps], pph, w0);
                watched_wait(a.dbg != nullptr, &acc_empty[S], ph ^ 1, w1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if constexpr (TS) {
                    // D'[m' = (output j, component, digit)][n' = window] += taps[m'][k] . plane[32 (n' + b) + k]:
                    // the data planes are the B operand (same Hankel descriptor, NW rows), 3 KB of shared
                    // memory per MMA instead of 8
#pragma unroll
                    for (int dl = 0; dl < 2; dl++) {
                        const unsigned d = tmem_base + (unsigned)(S * COLS + dl * NW);
                        constexpr unsigned idesc_lo = umma_idesc_i8_ab(true, false, NW), idesc_hi = umma_idesc_i8_ab(true, true, NW);
#pragma unroll
                        for (int dc = 0; dc < DC; dc++) {
                            unsigned long long bd = a_base[0][2 * dc + dl] + (unsigned long long)ps * (unsigned long long)((NPL * PLa) >> 4);
                            unsigned at = tmem_base + (unsigned)(ACOL + 8 * dc * NB);
                            if (dc == 0) { umma_i8_ts_first(d, at, bd, dl ? idesc_hi : idesc_lo); bd += kAStep; at += 8; }
                            if constexpr (NBT > 0) {
#pragma unroll
                                for (int b = dc == 0 ? 1 : 0; b < NBT; b++) {
                                    umma_i8_ts_acc(d, at, bd, dl ? idesc_hi : idesc_lo);
                                    bd += kAStep; at += 8;
                                }
                            } else {
#pragma unroll 4
                                for (int b = dc == 0 ? 1 : 0; b < NB; b++) {
                                    umma_i8_ts_acc(d, at, bd, dl ? idesc_hi : idesc_lo);
                                    bd += kAStep; at += 8;
                                }
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int dl = 0; dl < 2; dl++) {                       // data limb: its own accumulator region
                        const unsigned d = tmem_base + (unsigned)(S * COLS + dl * N);
                        constexpr unsigned idesc_lo = umma_idesc_i8(false, N), idesc_hi = umma_idesc_i8(true, N);
#pragma unroll
                        for (int dc = 0; dc < DC; dc++) {
                            // k-block b: the A window starts one 32-byte row later (Hankel), B is the next tile
                            unsigned long long ad = a_base[S][2 * dc + dl], bd = b_base[dc];
                            if (dc == 0) { umma_i8_first(d, ad, bd, dl ? idesc_hi : idesc_lo); ad += kAStep; bd += kBStep; }
#pragma unroll 4
                            for (int b = dc == 0 ? 1 : 0; b < NB; b++) {
                                umma_i8_acc(d, ad, bd, dl ? idesc_hi : idesc_lo);
                                ad += kAStep; bd += kBStep;
                            }
                        }
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&planes_empty[ps])) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&acc_full[S])) : "memory");
            };
            int ps = 0;
            unsigned pph = 0;
            auto next_planes = [&] { if (++ps == PS) { ps = 0; pph ^= 1; } };
            for (int i = 0; i < ntl; i += 2) {
                const unsigned ph = (unsigned)(i >> 1) & 1;
                issue_tile(std::integral_constant<int, 0>{}, ph, ps, pph);
                next_planes();
                if (i + 1 < ntl) { issue_tile(std::integral_constant<int, 1>{}, ph, ps, pph); next_planes(); }
            }
        }
    } else if (warp >= kU32EpiWarps) {
        // ========================================================================= stagers
        // raw (re, im) int16 samples -> byte planes, stored with the 32-byte swizzle: the 16-byte chunk
        // c of a plane lives at chunk c ^ (c >> 3 & 1)
        const int st = tid - 32 * kU32EpiWarps;
        constexpr int NST = 32 * SW;
        const int nq = PL / 4;
        for (int i = 0, r = 0, rph = 0, s = 0, ph = 0; i < ntl; i++) {      // s / ph: plane stage and its phase
            const long long tile = first + (long long)i * step, o0 = tile * TILE;
            const bool landed = bulk_ok(tile);
            watched_wait(a.dbg != nullptr, &raw_full[r], (unsigned)rph, w0);
            watched_wait(a.dbg != nullptr, &planes_empty[s], (unsigned)ph ^ 1, w1);
            unsigned *pl = reinterpret_cast<unsigned *>(planes + (size_t)s * NPL * PLa);
            const unsigned char *rw = raw + (size_t)r * slot_bytes;
            const bool timed = a.dbg != nullptr;
            const long long t_c0 = timed ? clock64() : 0;
            // item q of a landed tile: 16 bytes starting `mis` bytes into unit q of the slot
            auto fetch = [&](const uint4 *src) {
                const uint4 lo = src[0];
                const uint4 hi = src[1];
                const unsigned bs = (unsigned)(mis & 3) * 8;     // complex samples are 4 bytes: bs == 0 there
                auto sh = [&](unsigned x, unsigned y) { return DC == 2 ? x : __funnelshift_r(x, y, bs); };
                switch (mis >> 2) {
                case 0: return make_uint4(sh(lo.x, lo.y), sh(lo.y, lo.z), sh(lo.z, lo.w), sh(lo.w, hi.x));
                case 1: return make_uint4(sh(lo.y, lo.z), sh(lo.z, lo.w), sh(lo.w, hi.x), sh(hi.x, hi.y));
                case 2: return make_uint4(sh(lo.z, lo.w), sh(lo.w, hi.x), sh(hi.x, hi.y), sh(hi.y, hi.z));
                default: return make_uint4(sh(lo.w, hi.x), sh(hi.x, hi.y), sh(hi.y, hi.z), sh(hi.z, hi.w));
                }
            };
            if constexpr (TR) {
                // real int16: item q = eight samples = 16 raw bytes -> eight bytes (half a 16-byte chunk) of the lo and of the hi
                // plane, stored with the 64-byte swizzle: chunk c lives at c ^ (c >> 3 & 3).  q + NST is NST / 2 chunks on, a
                // multiple of 32: the swizzle bits do not change.
                constexpr int PLc = kU32trTile + 32 * (NBT ? NBT : 1), PLac = (PLc + 511) / 512 * 512, NQc = PLc / 8;
                constexpr int FULL = NQc / NST, REM = NQc % NST;
                static_assert(NST % 64 == 0, "q + NST must keep the swizzle bits");
                const int c0 = st >> 1;
                unsigned *dst = pl + (((c0 ^ ((c0 >> 3) & 3)) << 2) | ((st & 1) << 1));
                auto split = [&](const uint4 v, unsigned *p) {
                    *reinterpret_cast<uint2 *>(p) = make_uint2(prmt_u(v.x, v.y, 0x6420), prmt_u(v.z, v.w, 0x6420));
                    *reinterpret_cast<uint2 *>(p + PLac / 4) = make_uint2(prmt_u(v.x, v.y, 0x7531), prmt_u(v.z, v.w, 0x7531));
                };
                if (landed) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(rw) + st;
                    uint4 v[FULL];
                    uint4 vt = make_uint4(0, 0, 0, 0);
                    if (mis == 0) {          // one uniform branch per tile, not per item
#pragma unroll
                        for (int u = 0; u < FULL; u++) v[u] = src[u * NST];
                        if (REM && st < REM) vt = src[FULL * NST];
                    } else {
#pragma unroll
                        for (int u = 0; u < FULL; u++) v[u] = fetch(src + u * NST);
                        if (REM && st < REM) vt = fetch(src + FULL * NST);
                    }
#pragma unroll
                    for (int u = 0; u < FULL; u++) split(v[u], dst + u * (NST * 2));
                    if (REM && st < REM) split(vt, dst + FULL * (NST * 2));
                } else {
                    // edge tiles and unaligned streams: guarded element loads
                    const unsigned short *__restrict__ in16 = static_cast<const unsigned short *>(a.in);
#pragma unroll 1
                    for (int u = 0; u <= FULL; u++) {
                        const int q = st + u * NST;
                        if (q >= NQc) break;
                        const long long sm = o0 + 8LL * q;
                        unsigned w[4];
#pragma unroll
                        for (int k = 0; k < 4; k++) {
                            const unsigned x0 = sm + 2 * k < a.n_in ? __ldg(in16 + sm + 2 * k) : 0u;
                            const unsigned x1 = sm + 2 * k + 1 < a.n_in ? __ldg(in16 + sm + 2 * k + 1) : 0u;
                            w[k] = x0 | (x1 << 16);
                        }
                        split(make_uint4(w[0], w[1], w[2], w[3]), dst + u * (NST * 2));
                    }
                }
            } else if (TS && landed) {
                // every size is a compile-time constant here: item q = st + NST u of this thread (four samples: 16 raw
                // bytes -> one word of each plane) sits 16 NST bytes further in the landing slot and 4 NST bytes further
                // in the planes (q + NST is NST / 4 chunks on, a multiple of 16: the swizzle bit (c >> 3) & 1 does not change)
                constexpr int PLc = kU32tTile + 32 * (NBT ? NBT : 1), PLac = (PLc + 255) / 256 * 256, NQc = PLc / 4;
                constexpr int FULL = NQc / NST, REM = NQc % NST;
                static_assert(NST % 64 == 0, "q + NST must keep the swizzle bit");
                const int c0 = st >> 2;
                const uint4 *src = reinterpret_cast<const uint4 *>(rw) + st;
                unsigned *dst = pl + (((c0 ^ ((c0 >> 3) & 1)) << 2) | (st & 3));
                auto split = [&](const uint4 v, unsigned *p) {
                    const unsigned t01 = prmt_u(v.x, v.y, 0x5140), t23 = prmt_u(v.z, v.w, 0x5140);
                    const unsigned u01 = prmt_u(v.x, v.y, 0x7362), u23 = prmt_u(v.z, v.w, 0x7362);
                    p[0] = prmt_u(t01, t23, 0x5410);
                    p[PLac / 4] = prmt_u(t01, t23, 0x7632);
                    p[2 * (PLac / 4)] = prmt_u(u01, u23, 0x5410);
                    p[3 * (PLac / 4)] = prmt_u(u01, u23, 0x7632);
                };
                uint4 v[FULL];
                uint4 vt = make_uint4(0, 0, 0, 0);
                if (mis == 0) {              // one uniform branch per tile, not per item
#pragma unroll
                    for (int u = 0; u < FULL; u++) v[u] = src[u * NST];
                    if (REM && st < REM) vt = src[FULL * NST];
                } else {
#pragma unroll
                    for (int u = 0; u < FULL; u++) v[u] = fetch(src + u * NST);
                    if (REM && st < REM) vt = fetch(src + FULL * NST);
                }
#pragma unroll
                for (int u = 0; u < FULL; u++) split(v[u], dst + u * NST);
                if (REM && st < REM) split(vt, dst + FULL * NST);
            } else if constexpr (DC == 2) {
                const unsigned *__restrict__ in32 = static_cast<const unsigned *>(a.in);
                for (int q0 = st; q0 < nq; q0 += kU32Batch * NST) {
                    uint4 v[kU32Batch];
#pragma unroll
                    for (int u = 0; u < kU32Batch; u++) {
                        const int q = q0 + u * NST;
                        if (q >= nq) { v[u] = make_uint4(0, 0, 0, 0); continue; }
                        if (landed) {
                            v[u] = reinterpret_cast<const uint4 *>(rw)[q];
                        } else {
                            const long long sm = o0 + 4LL * q;
                            v[u].x = sm < a.n_in ? __ldg(in32 + sm) : 0u;
                            v[u].y = sm + 1 < a.n_in ? __ldg(in32 + sm + 1) : 0u;
                            v[u].z = sm + 2 < a.n_in ? __ldg(in32 + sm + 2) : 0u;
                            v[u].w = sm + 3 < a.n_in ? __ldg(in32 + sm + 3) : 0u;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kU32Batch; u++) {
                        const int q = q0 + u * NST;
                        if (q >= nq) continue;
                        const unsigned t01 = prmt_u(v[u].x, v[u].y, 0x5140), t23 = prmt_u(v[u].z, v[u].w, 0x5140);
                        const unsigned u01 = prmt_u(v[u].x, v[u].y, 0x7362), u23 = prmt_u(v[u].z, v[u].w, 0x7362);
                        const int c = q >> 2;
                        unsigned *p = pl + (((c ^ ((c >> 3) & 1)) << 2) | (q & 3));
                        p[0] = prmt_u(t01, t23, 0x5410);
                        p[PLa / 4] = prmt_u(t01, t23, 0x7632);
                        p[2 * (PLa / 4)] = prmt_u(u01, u23, 0x5410);
                        p[3 * (PLa / 4)] = prmt_u(u01, u23, 0x7632);
                    }
                }
            } else {
                const unsigned short *__restrict__ in16 = static_cast<const unsigned short *>(a.in);
                for (int q0 = st; q0 < nq; q0 += kU32Batch * NST) {
                    uint2 v[kU32Batch];
#pragma unroll
                    for (int u = 0; u < kU32Batch; u++) {
                        const int q = q0 + u * NST;
                        if (q >= nq) { v[u] = make_uint2(0, 0); continue; }
                        if (landed) {
                            v[u] = reinterpret_cast<const uint2 *>(rw)[q];
                        } else {
                            const long long sm = o0 + 4LL * q;
                            const unsigned x0 = sm < a.n_in ? __ldg(in16 + sm) : 0u, x1 = sm + 1 < a.n_in ? __ldg(in16 + sm + 1) : 0u;
                            const unsigned x2 = sm + 2 < a.n_in ? __ldg(in16 + sm + 2) : 0u, x3 = sm + 3 < a.n_in ? __ldg(in16 + sm + 3) : 0u;
                            v[u] = make_uint2(x0 | (x1 << 16), x2 | (x3 << 16));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < kU32Batch; u++) {
                        const int q = q0 + u * NST;
                        if (q >= nq) continue;
                        const int c = q >> 2;
                        unsigned *p = pl + (((c ^ ((c >> 3) & 1)) << 2) | (q & 3));
                        p[0] = prmt_u(v[u].x, v[u].y, 0x6420);
                        p[PLa / 4] = prmt_u(v[u].x, v[u].y, 0x7531);
                    }
                }
            }
            const long long t_c1 = timed ? clock64() : 0;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&planes_full[s]);
            mbar_arrive(&raw_empty[r]);
            if (timed) { const long long t_c2 = clock64(); t_conv += t_c1 - t_c0; t_fence += t_c2 - t_c1; }
            if (++r == R) { r = 0; rph ^= 1; }
            if (++s == PS) { s = 0; ph ^= 1; }
        }
    } else if constexpr (TS) {
        // ============================================================= epilogue (swapped operands)
        // TMEM lane m' = 32 quad + 16 cls + 8 digit + jr  (output j = 8 quad + jr), column = window n'.
        // The 16-lane load shape hands thread t the accumulator fragment of an m16n8 tile: rows t / 4 and
        // t / 4 + 8 -- the two tap digits of one (j, component) -- and columns 2 (t % 4), + 1 of every
        // 8-column group (tools/probe_tmem_ld_shapes.cu); the re and im halves are the two 16-lane loads.
        // Per output and component, as in the other kernels:
        //   y = lo_d0 + ((lo_d1 + hi_d0) << 8) + (hi_d1 << 16)  mod 2^32, bits [16, 32) kept (fromQ)
        // warp = (lane quadrant = 8 values of j, half of the windows); a store instruction writes 8 consecutive
        // outputs (32 bytes) of 4 windows.
        const int quad = warp & 3, half = warp >> 2, q = lane & 3;
        const unsigned lane_addr = tmem_base + ((unsigned)(32 * quad) << 16);
        unsigned *out32 = static_cast<unsigned *>(a.out);
        for (int i = 0; i < ntl; i++) {
            const int s = i & 1;
            const unsigned ph = (unsigned)(i >> 1) & 1;
            const long long tile = first + (long long)i * step;
            // output of (window n0 + 2 q, j); the thread's others are whole windows (32 outputs) further on
            // (32-bit words: a complex sample, or -- TR -- two consecutive real outputs: rows ordered so that the second 16-lane
            // load holds the odd outputs of the same window, see fir_umma32_configure)
            const long long o00 = tile * kU32tTile + 32LL * ((NW / 2) * half + 2 * q) + 8 * quad + (lane >> 2);
            const bool whole = (tile + 1) * TILE <= a.n_out && (!TR || (reinterpret_cast<unsigned long long>(a.out) & 3) == 0);
            const unsigned tcol = lane_addr + (unsigned)(s * COLS + (NW / 2) * half);
            unsigned *const ot = out32 + o00;
            watched_wait(a.dbg != nullptr, &acc_full[s], ph, w0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c = 0; c < NW / 32; c++) {
                unsigned lr[8], li[8], hr[8], hi[8];      // limb (lo / hi) x component (re / im)
                tmem_ld16x256b_x2(tcol + 16 * c, lr);
                tmem_ld16x256b_x2(tcol + (16u << 16) + 16 * c, li);
                tmem_ld16x256b_x2(tcol + NW + 16 * c, hr);
                tmem_ld16x256b_x2(tcol + (16u << 16) + NW + 16 * c, hi);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == NW / 32 - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&acc_empty[s]);
                }
                unsigned *o = ot + 32 * 16 * c;
                unsigned res[4];
#pragma unroll
                for (int g = 0; g < 2; g++)
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int k = 4 * g + e;                              // digit 0 at k, digit 1 at k + 2
                        const unsigned yr = lr[k] + ((lr[k + 2] + hr[k]) << 8) + (hr[k + 2] << 16);
                        const unsigned yi = li[k] + ((li[k + 2] + hi[k]) << 8) + (hi[k + 2] << 16);
                        const int w = 32 * (8 * g + e);                       // window 8 g + 2 q + e of the chunk
                        res[2 * g + e] = prmt_u(yr, yi, 0x7632);
                        if (whole) __stcg(o + w, res[2 * g + e]);
                    }
                if (!whole)       // the stream's last tile (TR: or an output pointer that is not word-aligned)
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const long long wi = o00 + 32 * 16 * c + 32 * (8 * (k >> 1) + (k & 1));
                        if constexpr (TR) {
                            unsigned short *o16 = static_cast<unsigned short *>(a.out);
                            if (2 * wi < a.n_out) o16[2 * wi] = (unsigned short)res[k];
                            if (2 * wi + 1 < a.n_out) o16[2 * wi + 1] = (unsigned short)(res[k] >> 16);
                        } else {
                            if (wi < a.n_out) o[32 * (8 * (k >> 1) + (k & 1))] = res[k];
                        }
                    }
            }
        }
    } else {
        // ======================================================================== epilogue
        // thread = (row m = TMEM lane, half of the row's 32 outputs); per output and component:
        //   y = lo_d0 + ((lo_d1 + hi_d0) << 8) + (hi_d1 << 16)  mod 2^32, bits [16, 32) kept (fromQ)
        const int m = 32 * (warp & 3) + lane, half = warp >> 2;
        const unsigned lane_addr = tmem_base + ((unsigned)(32 * (warp & 3)) << 16);
        for (int i = 0; i < ntl; i++) {
            const int s = i & 1;
            const unsigned ph = (unsigned)(i >> 1) & 1;
            const long long tile = first + (long long)i * step, orow = tile * TILE + 32LL * m + 16 * half;
            watched_wait(a.dbg != nullptr, &acc_full[s], ph, w0);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int c = 0; c < 16 / CH; c++) {
                const int n0 = 16 * half + CH * c;
                unsigned lo[16], hi[16];
                tmem_ld16(lane_addr + (unsigned)(s * COLS + n0 * NQ), lo);
                tmem_ld16(lane_addr + (unsigned)(s * COLS + N + n0 * NQ), hi);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (c == 16 / CH - 1) {
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                    mbar_arrive(&acc_empty[s]);
                }
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
                const long long o = orow + CH * c;
                if constexpr (DC == 2) {
                    unsigned *out32 = static_cast<unsigned *>(a.out);
                    if (o + CH <= a.n_out && (reinterpret_cast<unsigned long long>(out32) & 15) == 0) {
                        __stcg(reinterpret_cast<uint4 *>(out32 + o), make_uint4(res[0], res[1], res[2], res[3]));
                    } else {
#pragma unroll
                        for (int k = 0; k < CH; k++)
                            if (o + k < a.n_out) out32[o + k] = res[k];
                    }
                } else {
                    unsigned short *out16 = static_cast<unsigned short *>(a.out);
                    if (o + CH <= a.n_out && (reinterpret_cast<unsigned long long>(out16) & 15) == 0) {
                        __stcg(reinterpret_cast<uint4 *>(out16 + o), make_uint4(res[0] | (res[1] << 16), res[2] | (res[3] << 16),
                                                                               res[4 % CH] | (res[5 % CH] << 16), res[6 % CH] | (res[7 % CH] << 16)));
                    } else {
#pragma unroll
                        for (int k = 0; k < CH; k++)
                            if (o + k < a.n_out) out16[o + k] = (unsigned short)res[k];
                    }
                }
            }
        }
    }
    if (a.dbg && lane == 0) {
        long long *d = a.dbg + (size_t)blockIdx.x * 8;
        if (warp == 0) { d[0] = clock64() - t_begin; d[1] = ntl; d[7] = w0; }
        if (warp == kU32EpiWarps) { d[5] = w0; d[6] = w1; a.dbg[(size_t)gridDim.x * 8 + 2 * blockIdx.x] = t_conv; a.dbg[(size_t)gridDim.x * 8 + 2 * blockIdx.x + 1] = t_fence; }
        if (warp == kU32EpiWarps + SW) { d[3] = w0; d[4] = w1; }
        if (warp == kU32EpiWarps + SW + 1) d[2] = w0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ALLOC) : "memory");
}

template <int DC>
__global__ void __launch_bounds__(kU32Threads, 1) fir_umma32_kernel(const FirUmma32Args a) { fir_umma32_body<DC, false, 0>(a); }
template <int NBT>
__global__ void __launch_bounds__(kU32tThreads, 1) fir_umma32t_kernel(const FirUmma32Args a) { fir_umma32_body<2, true, NBT, kU32tStageWarps>(a); }
template <int NBT>   // real int16 data
__global__ void __launch_bounds__(kU32tThreads, 1) fir_umma32tr_kernel(const FirUmma32Args a) { fir_umma32_body<1, true, NBT, kU32tStageWarps>(a); }

// ------------------------------------------------------------------------------- host ---
// balanced byte digits (d0, d1) of q = d0 + 256 d1, both in [-128, 127]; false if q does not fit
static bool two_digits(long long q, int8_t &d0, int8_t &d1)
{
    const long long lo = ((q + 128) & 255) - 128, hi = (q - lo) >> 8;
    if (hi < -128 || hi > 127) return false;
    d0 = (int8_t)lo; d1 = (int8_t)hi;
    return true;
}

static size_t u32_fixed_smem(int dc, int NB, int PLa, bool swapped = false)
{
    return (swapped ? 0 : (size_t)dc * NB * (32 * dc * 2) * 32) + (swapped ? kU32tPlaneStages : 2) * ((size_t)dc * 2 * PLa) + 1024;
}

// swap: 1 forces the operand-swapped kernel where it applies, 0 forbids it, -1 follows B200C_UMMA32T
int fir_umma32_configure(FirUmma32Plan &p, const FirImmaPlan &base, const double *taps, bool force, int swap)
{
    p.ready = false; p.swapped = false;
    const bool enabled = [] { const char *e = std::getenv("B200C_UMMA32"); return !e || std::atoi(e) != 0; }();   // B200C_UMMA32=0: stay on fir_umma_kernel
    if (!base.ready || !(enabled || force) || base.nlt != 2) return B200C_OK;
    const int K = base.K, dc = base.dc, tc = base.tc, NQ = dc * 2, N = 32 * NQ;
    static const bool swap_default = [] { const char *e = std::getenv("B200C_UMMA32T"); return !e || std::atoi(e) != 0; }();   // =0: the original formulation
    // k-blocks per window: 32 outputs + K - 1 positions (the swapped kernel on real data: 64 outputs)
    const int NBo = (K + 31 + 31) / 32, NBr = (K + 63 + 31) / 32;
    const bool swapped = (dc == 2 ? NBo <= kU32tMaxNB : NBr <= kU32trMaxNB) && (swap < 0 ? swap_default : swap != 0);
    const bool swapped_real = swapped && dc == 1;
    const int NB = swapped_real ? NBr : NBo;
    const int PL = (swapped_real ? kU32trTile : swapped ? kU32tTile : kU32Tile) + 32 * NB;
    const int PLa = swapped_real ? (PL + 511) / 512 * 512 : (PL + 255) / 256 * 256;
    // B tiles, two plane stages and at least a 3-deep landing ring must fit
    if (u32_fixed_smem(dc, NB, PLa, swapped) + 3 * ((size_t)PL * dc * 2 + 128) > 200 * 1024) return B200C_OK;
    std::vector<uint8_t> bm(swapped_real ? 0 : (size_t)dc * NB * N * 32, 0), am(swapped ? (size_t)dc * NB * 128 * 32 : 0, 0);
    for (int d = 0; d < K; d++)
        for (int c = 0; c < tc; c++) {
            const long long q = (long long)(int32_t)(long long)std::ldexp(taps[(size_t)d * tc + c], 16);
            int8_t pos[2], neg[2];
            if (!two_digits(q, pos[0], pos[1]) || !two_digits(-q, neg[0], neg[1])) return B200C_OK;   // needs 3 digits: other kernels
            // contributions (data component dcx, output component cls, sign) of tap component c:
            //   complex x complex: h_re: (re,yr,+) (im,yi,+);  h_im: (re,yi,+) (im,yr,-)
            //   complex x real:    h_re: (re,yr,+) (im,yi,+);  real x real: (0,0,+)
            struct Use { int dcx, cls; bool negate; };
            Use uses[2];
            int nuse = 0;
            if (dc == 1) { uses[nuse++] = {0, 0, false}; }
            else if (c == 0) { uses[nuse++] = {0, 0, false}; uses[nuse++] = {1, 1, false}; }
            else { uses[nuse++] = {0, 1, false}; uses[nuse++] = {1, 0, true}; }
            if (swapped_real) {
                // 64 outputs per window; TMEM lane of (output n, digit l): 32 quad + 16 (n & 1) + 8 l + (n >> 1 & 7) with
                // quad = n >> 4 -- the epilogue's two 16-lane loads then hold the even and the odd outputs of the same pair
                for (int n = 0; n < 64; n++) {
                    const int j = n + K - 1 - d, b = j / 32, jj = j % 32;
                    for (int l = 0; l < 2; l++) {
                        const int row = 32 * (n >> 4) + 16 * (n & 1) + 8 * l + ((n >> 1) & 7);
                        am[((size_t)b * 128 + row) * 32 + jj] = (uint8_t)pos[l];
                    }
                }
                continue;
            }
            for (int u = 0; u < nuse; u++)
                for (int n = 0; n < 32; n++) {
                    const int j = n + K - 1 - d;                         // window position of tap d for output n
                    const int b = j / 32, jj = j % 32;
                    for (int l = 0; l < 2; l++) {
                        const int col = n * NQ + uses[u].cls * 2 + l;
                        // canonical K-major no-swizzle: [col / 8][k chunk][col % 8][16 bytes]
                        const size_t at = (size_t)(uses[u].dcx * NB + b) * N * 32 + (size_t)(col / 8) * 256 + (size_t)(jj / 16) * 128 +
                                          (size_t)(col % 8) * 16 + (jj % 16);
                        bm[at] = (uint8_t)(uses[u].negate ? neg[l] : pos[l]);
                        // the same element of the swapped kernel's A tile: row-major, 32 bytes per row, rows ordered so
                        // that one thread of the epilogue's 16-lane tensor-memory loads holds both digits
                        if (swapped) {
                            const int row = 32 * (n / 8) + 16 * uses[u].cls + 8 * l + n % 8;   // TMEM lane: see the epilogue
                            am[((size_t)(uses[u].dcx * NB + b) * N + row) * 32 + jj] = bm[at];
                        }
                    }
                }
        }
    if (!bm.empty() && bm.size() > p.capacity) {
        if (p.d_bmat) cudaFree(p.d_bmat);
        p.d_bmat = nullptr; p.capacity = 0;
        B200C_CUDA_TRY(cudaMalloc(&p.d_bmat, bm.size()));
        p.capacity = bm.size();
    }
    if (!bm.empty()) B200C_CUDA_TRY(cudaMemcpy(p.d_bmat, bm.data(), bm.size(), cudaMemcpyHostToDevice));
    if (swapped) {
        if (am.size() > p.a_capacity) {
            if (p.d_amat) cudaFree(p.d_amat);
            p.d_amat = nullptr; p.a_capacity = 0;
            B200C_CUDA_TRY(cudaMalloc(&p.d_amat, am.size()));
            p.a_capacity = am.size();
        }
        B200C_CUDA_TRY(cudaMemcpy(p.d_amat, am.data(), am.size(), cudaMemcpyHostToDevice));
    }
    p.K = K; p.NB = NB; p.dc = dc; p.swapped = swapped;
    p.ready = true;
    return B200C_OK;
}

void fir_umma32_destroy(FirUmma32Plan &p)
{
    if (p.d_bmat) cudaFree(p.d_bmat);
    if (p.d_amat) cudaFree(p.d_amat);
    p.d_bmat = nullptr; p.capacity = 0; p.d_amat = nullptr; p.a_capacity = 0; p.ready = false; p.swapped = false;
}

static_assert(kU32tMaxNB == 8, "launch_u32 instantiates fir_umma32t_kernel for 1..8 k-blocks");
template <int DC, bool TS>
static int launch_u32(FirUmma32Args a, int sm_count, cudaStream_t stream)
{
    void (*kern)(FirUmma32Args) = fir_umma32_kernel<DC>;
    if (TS && DC == 1) {
        switch (a.NB) {      // 64 outputs + K - 1 positions: at least two k-blocks
        case 2: kern = fir_umma32tr_kernel<2>; break;
        case 3: kern = fir_umma32tr_kernel<3>; break;
        case 4: kern = fir_umma32tr_kernel<4>; break;
        case 5: kern = fir_umma32tr_kernel<5>; break;
        case 6: kern = fir_umma32tr_kernel<6>; break;
        case 7: kern = fir_umma32tr_kernel<7>; break;
        case 8: kern = fir_umma32tr_kernel<8>; break;
        case 9: kern = fir_umma32tr_kernel<9>; break;
        case 10: kern = fir_umma32tr_kernel<10>; break;
        case 11: kern = fir_umma32tr_kernel<11>; break;
        case 12: kern = fir_umma32tr_kernel<12>; break;
        case 13: kern = fir_umma32tr_kernel<13>; break;
        case 14: kern = fir_umma32tr_kernel<14>; break;
        case 15: kern = fir_umma32tr_kernel<15>; break;
        default: kern = fir_umma32tr_kernel<16>; break;
        }
    } else if (TS) {
        switch (a.NB) {
        case 1: kern = fir_umma32t_kernel<1>; break;
        case 2: kern = fir_umma32t_kernel<2>; break;
        case 3: kern = fir_umma32t_kernel<3>; break;
        case 4: kern = fir_umma32t_kernel<4>; break;
        case 5: kern = fir_umma32t_kernel<5>; break;
        case 6: kern = fir_umma32t_kernel<6>; break;
        case 7: kern = fir_umma32t_kernel<7>; break;
        default: kern = fir_umma32t_kernel<8>; break;
        }
    }
    const int threads = TS ? kU32tThreads : kU32Threads, slot = TS ? a.NB : 0;
    static_assert(kU32trMaxNB == 16, "launch_u32 instantiates fir_umma32tr_kernel for 2..16 k-blocks");
    static thread_local bool configured[16][17] = {{false}};
    int dev = 0;
    B200C_CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 16 && !configured[dev][slot]) {
        B200C_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        configured[dev][slot] = true;
    }
    static const int ring = [] { const char *e = std::getenv("B200C_UMMA_RING"); return e ? std::atoi(e) : kU32MaxRing; }();
    const size_t fixed = u32_fixed_smem(DC, a.NB, a.PLa, TS), one = (size_t)a.PL * DC * 2 + (TS ? 128 : 0);
    a.R = (int)std::max<size_t>(2, std::min<size_t>((size_t)std::max(2, std::min(ring, kU32MaxRing)), (216 * 1024 - fixed) / one));
    // one CTA per SM: its two accumulator stages take all (complex) or half (real) of tensor memory
    const size_t smem = std::max<size_t>(fixed + a.R * one, 116 * 1024);
    const int grid = (int)std::min<long long>(a.ntiles, (long long)sm_count);
    static const bool dbg = std::getenv("B200C_UMMA_DBG") != nullptr;
    unsigned long long *watch = nullptr;
    if (dbg) {
        B200C_CUDA_TRY(cudaMalloc(&a.dbg, (size_t)grid * 10 * sizeof(long long)));
        B200C_CUDA_TRY(cudaMemset(a.dbg, 0, (size_t)grid * 10 * sizeof(long long)));
        // watchdog records in host-mapped memory: readable after the kernel trapped
        B200C_CUDA_TRY(cudaHostAlloc(&watch, ((size_t)grid * 32 + 8) * sizeof(unsigned long long), cudaHostAllocMapped));
        std::memset(watch, 0, ((size_t)grid * 32 + 8) * sizeof(unsigned long long));
        unsigned long long *dwatch = nullptr;
        B200C_CUDA_TRY(cudaHostGetDevicePointer(&dwatch, watch, 0));
        B200C_CUDA_TRY(cudaMemcpyToSymbol(g_umma_watch, &dwatch, sizeof(dwatch)));
    }
    kern<<<grid, threads, smem, stream>>>(a);
    B200C_CUDA_TRY(cudaGetLastError());
    if (dbg) {
        std::vector<long long> h((size_t)grid * 10);
        if (cudaStreamSynchronize(stream) != cudaSuccess) {
            const unsigned long long *names = watch + (size_t)grid * 32;
            static const char *const what[6] = {"raw_full", "raw_empty", "planes_full", "planes_empty", "acc_full", "acc_empty"};
            int shown = 0;
            for (int g = 0; g < grid && shown < 40; g++)
                for (int w = 0; w < 32 && shown < 40; w++) {
                    const unsigned long long r = watch[(size_t)g * 32 + w];
                    if (!(r & 1)) continue;
                    const unsigned long long addr = r >> 8;
                    int which = -1;
                    for (int k = 0; k < 6; k++) if (addr >= names[k] && (which < 0 || names[k] > names[which])) which = k;
                    std::fprintf(stderr, "umma32 watchdog: block %d warp %d stuck on %s[%llu] parity %llu\n", g, w, which >= 0 ? what[which] : "?",
                                 which >= 0 ? (addr - names[which]) / 8 : 0ull, (r >> 1) & 1);
                    shown++;
                }
            return B200C_ERR_CUDA;
        }
        B200C_CUDA_TRY(cudaMemcpy(h.data(), a.dbg, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        cudaFree(a.dbg);
        {   // the next (untimed) launch must not write through a freed pointer
            unsigned long long *none = nullptr;
            B200C_CUDA_TRY(cudaMemcpyToSymbol(g_umma_watch, &none, sizeof(none)));
        }
        cudaFreeHost(watch);
        double s8[8] = {0};
        for (int g = 0; g < grid; g++) for (int k = 0; k < 8; k++) s8[k] += (double)h[(size_t)g * 8 + k] / grid;
        double tc = 0, tf = 0;
        for (int g = 0; g < grid; g++) { tc += (double)h[(size_t)grid * 8 + 2 * g] / grid; tf += (double)h[(size_t)grid * 8 + 2 * g + 1] / grid; }
        std::fprintf(stderr, "umma32: stager convert %.0f fence+arrive %.0f cycles/tile\n", tc / s8[1], tf / s8[1]);
        std::fprintf(stderr, "umma32: cycles/tile %.0f | issuer wait raw_empty %.0f | mma wait planes_full %.0f acc_empty %.0f | stager wait raw_full %.0f planes_empty %.0f | epilogue wait acc_full %.0f (tiles/CTA %.1f, R %d, smem %zu)\n",
                     s8[0] / s8[1], s8[2] / s8[1], s8[3] / s8[1], s8[4] / s8[1], s8[5] / s8[1], s8[6] / s8[1], s8[7] / s8[1], s8[1], a.R, smem);
    }
    return B200C_OK;
}

int fir_umma32_launch(const FirUmma32Plan &p, const void *d_in, size_t in_elems, void *d_out, size_t n_out, int sm_count,
                      cudaStream_t stream)
{
    if (n_out == 0) return B200C_OK;
    FirUmma32Args a;
    const int tile = p.swapped ? (p.dc == 1 ? kU32trTile : kU32tTile) : kU32Tile;
    a.in = d_in; a.out = d_out; a.bmat = p.swapped ? p.d_amat : p.d_bmat;
    a.n_in = (long long)in_elems; a.n_out = (long long)n_out;
    a.ntiles = ((long long)n_out + tile - 1) / tile;
    a.K = p.K; a.NB = p.NB; a.PL = tile + 32 * p.NB; a.R = 2; a.dbg = nullptr;
    a.PLa = p.swapped && p.dc == 1 ? (a.PL + 511) / 512 * 512 : (a.PL + 255) / 256 * 256;
    if (p.swapped) return p.dc == 1 ? launch_u32<1, true>(a, sm_count, stream) : launch_u32<2, true>(a, sm_count, stream);
    return p.dc == 1 ? launch_u32<1, false>(a, sm_count, stream) : launch_u32<2, false>(a, sm_count, stream);
}

} // namespace b200c
