// sm_100a batched FFT kernels for /comms/fft.
// Replaces FFT<Type>::work() -> FFTAux::transform (fft/FFT.cpp:61-72, fft/FFTAux.h:21-44),
// i.e. kissfft<T> (fft/kissfft.hh) for cf32/cf64 and the Q15 C kiss_fft (fft/kiss_fft.c built
// with FIXED_POINT=16) for complex int16.
//
// The reference recurses (decimation in time); here every transform is done ITERATIVELY by
// one CTA (or a fraction of one): a mixed-radix digit-reversal scatter into shared memory,
// then the radix stages innermost first, each element seeing exactly the reference's
// sequence of arithmetic -- which is what makes the int16 path bit-exact (per-stage
// C_FIXDIV 1/radix scaling, Q15 C_MUL with sround, int16 wrap on every add).
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdlib>
#include <cstring>
#include <type_traits>

#include "fft.hpp"

namespace b200c {

// ---------------------------------------------------------------- element arithmetic ---
template <typename T> struct FloatTraits {
    using E = typename std::conditional<sizeof(T) == 4, float2, double2>::type;
    using Sc = T;
    using S = E;   // storage type (global / shared memory) == register type
    using TwS = S; // storage type of the fast kernel's twiddle tables
    static constexpr bool kFixed = false;
    __device__ static E ld(S s) { return s; }
    __device__ static E ldtw(S s) { return s; }
    __device__ static S st(E e) { return e; }
    __device__ static E mk(T r, T i) { E e; e.x = r; e.y = i; return e; }
    __device__ static E add(E a, E b) { return mk(a.x + b.x, a.y + b.y); }
    __device__ static E sub(E a, E b) { return mk(a.x - b.x, a.y - b.y); }
    __device__ static E mul(E a, E b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
    __device__ static E fixdiv(E a, int) { return a; }
    __device__ static T smul(T a, T b) { return a * b; }
    __device__ static T half(T a) { return a * (T)0.5; }
    __device__ static T wrap(T a) { return a; }
};

struct Q15Traits {   // fft/_kiss_fft_guts.h:44-124 with FIXED_POINT=16
    using E = short2;
    using S = short2;
    using TwS = S;
    using Sc = int;  // int16 values carried in 32-bit registers, wrapped at every assignment
    static constexpr bool kFixed = true;
    __device__ static E ld(S s) { return s; }
    __device__ static E ldtw(S s) { return s; }
    __device__ static S st(E e) { return e; }
    __device__ static int wrap(int a) { return (int)(short)a; }
    __device__ static int sround(int x) { return (int)(short)((x + (1 << 14)) >> 15); }
    __device__ static E mk(int r, int i) { E e; e.x = (short)r; e.y = (short)i; return e; }
    __device__ static E add(E a, E b) { return mk((int)a.x + b.x, (int)a.y + b.y); }
    __device__ static E sub(E a, E b) { return mk((int)a.x - b.x, (int)a.y - b.y); }
    __device__ static E mul(E a, E b)
    {
        return mk(sround((int)((unsigned)((int)a.x * b.x) - (unsigned)((int)a.y * b.y))),
                  sround((int)((unsigned)((int)a.x * b.y) + (unsigned)((int)a.y * b.x))));
    }
    __device__ static E fixdiv(E a, int div)
    {
        const int f = 32767 / div;
        return mk(sround((int)a.x * f), sround((int)a.y * f));
    }
    __device__ static int smul(int a, int b) { return sround(a * b); }
    __device__ static int half(int a) { return a >> 1; }
};

// The same Q15 arithmetic with LAZY wrapping, for the fused 4096-point kernel.  kiss_fft stores every sum in an
// int16 (wrap mod 2^16, fft/_kiss_fft_guts.h:100-124); wrapping is a ring homomorphism, so a chain of adds and
// subtractions may be carried unwrapped in 32-bit registers as long as the value is reduced to its int16
// representative before it is used NON-linearly: as an operand of C_FIXDIV / C_MUL (smul then >> 15,
// guts:64-78), or when it is stored.  Components live in two 32-bit registers (no short2 packing between
// operations: the packed form cost one PRMT per assignment, 636 of the kernel's 2400 instructions); memory keeps
// the reference's packed (re, im) int16 pairs, and packing wraps for free.  Range argument for the products:
// C_FIXDIV by 4 leaves |x| <= 8191, a twiddle is <= 32767 in magnitude, so |re|, |im| of a C_MUL result are
// <= (2 * 8191 * 32767 + 16384) >> 15 = 16382: sround's int16 cast never wraps and is left out; sums of four
// such terms stay below 2^17.
struct Q15Lazy {
    using E = int2;        // (re, im), congruent mod 2^16 to the reference's int16 values
    using S = unsigned;    // packed (re, im) int16 pair as it lies in memory
    using Sc = int;
    static constexpr bool kFixed = true;
    __device__ static int sx(int a) { return (int)(short)a; }
    __device__ static E ld(S u) { return make_int2((int)u, (int)u >> 16); }          // re keeps garbage high bits (lazy)
    using TwS = int2;      // twiddles of the 4096-point kernel are stored unpacked (two ALU-pipe instructions less per load)
    __device__ static E ldtw(TwS u) { return u; }                                    // twiddles multiply: exact values
    __device__ static S st(E e) { return __byte_perm((unsigned)e.x, (unsigned)e.y, 0x5410); }
    __device__ static int wrap(int a) { return a; }
    __device__ static E mk(int r, int i) { return make_int2(r, i); }
    __device__ static E add(E a, E b) { return mk((int)((unsigned)a.x + (unsigned)b.x), (int)((unsigned)a.y + (unsigned)b.y)); }
    __device__ static E sub(E a, E b) { return mk((int)((unsigned)a.x - (unsigned)b.x), (int)((unsigned)a.y - (unsigned)b.y)); }
    // a: exact (a C_FIXDIV result), b: exact twiddle
    __device__ static E mul(E a, E b)
    {
        return mk((int)((unsigned)(a.x * b.x) - (unsigned)(a.y * b.y) + 16384u) >> 15,
                  (int)((unsigned)(a.x * b.y) + (unsigned)(a.y * b.x) + 16384u) >> 15);
    }
    __device__ static E fixdiv(E a, int div)
    {
        const int f = 32767 / div;
        return mk((sx(a.x) * f + 16384) >> 15, (sx(a.y) * f + 16384) >> 15);
    }
};

// radix-2/3/4/5 butterflies on F[0], F[m], ... with twiddles tw[q*tws]
template <typename Tr, typename E>
__device__ __forceinline__ void bfly2(E *F, int m, const E *__restrict__ tw, int tws)
{
    E f0 = Tr::fixdiv(F[0], 2), f1 = Tr::fixdiv(F[m], 2);
    const E t = Tr::mul(f1, tw[tws]);
    F[m] = Tr::sub(f0, t);
    F[0] = Tr::add(f0, t);
}

template <typename Tr, typename E>
__device__ __forceinline__ void bfly4(E *F, int m, const E *__restrict__ tw, int tws, int inverse)
{
    E f0 = Tr::fixdiv(F[0], 4), f1 = Tr::fixdiv(F[m], 4), f2 = Tr::fixdiv(F[2 * m], 4), f3 = Tr::fixdiv(F[3 * m], 4);
    const E s0 = Tr::mul(f1, tw[tws]);
    const E s1 = Tr::mul(f2, tw[2 * tws]);
    const E s2 = Tr::mul(f3, tw[3 * tws]);
    const E s5 = Tr::sub(f0, s1);
    f0 = Tr::add(f0, s1);
    const E s3 = Tr::add(s0, s2);
    const E s4 = Tr::sub(s0, s2);
    F[2 * m] = Tr::sub(f0, s3);
    F[0] = Tr::add(f0, s3);
    // forward: F[m] = s5 + (s4.i, -s4.r), F[3m] = s5 - (s4.i, -s4.r); inverse swaps the two
    const E rot = Tr::mk(s4.y, Tr::wrap(-s4.x));   // wrap(-x) only matters mod 2^16 below
    if (inverse) { F[m] = Tr::sub(s5, rot); F[3 * m] = Tr::add(s5, rot); }
    else { F[m] = Tr::add(s5, rot); F[3 * m] = Tr::sub(s5, rot); }
}

template <typename Tr, typename E>
__device__ __forceinline__ void bfly3(E *F, int m, const E *__restrict__ tw, int tws, E epi3)
{
    E f0 = Tr::fixdiv(F[0], 3), f1 = Tr::fixdiv(F[m], 3), f2 = Tr::fixdiv(F[2 * m], 3);
    const E s1 = Tr::mul(f1, tw[tws]);
    const E s2 = Tr::mul(f2, tw[2 * tws]);
    const E s3 = Tr::add(s1, s2);
    E s0 = Tr::sub(s1, s2);
    f1 = Tr::mk(f0.x - Tr::half(s3.x), f0.y - Tr::half(s3.y));
    s0 = Tr::mk(Tr::smul(s0.x, epi3.y), Tr::smul(s0.y, epi3.y));
    f0 = Tr::add(f0, s3);
    f2 = Tr::mk(f1.x + s0.y, f1.y - s0.x);
    f1 = Tr::mk(f1.x - s0.y, f1.y + s0.x);
    F[0] = f0; F[m] = f1; F[2 * m] = f2;
}

template <typename Tr, typename E>
__device__ __forceinline__ void bfly5(E *F, int m, const E *__restrict__ tw, int tws, E ya, E yb)
{
    using Sc = typename Tr::Sc;
    E f0 = Tr::fixdiv(F[0], 5), f1 = Tr::fixdiv(F[m], 5), f2 = Tr::fixdiv(F[2 * m], 5), f3 = Tr::fixdiv(F[3 * m], 5),
      f4 = Tr::fixdiv(F[4 * m], 5);
    const E c0 = f0;
    const E c1 = Tr::mul(f1, tw[tws]);
    const E c2 = Tr::mul(f2, tw[2 * tws]);
    const E c3 = Tr::mul(f3, tw[3 * tws]);
    const E c4 = Tr::mul(f4, tw[4 * tws]);
    const E c7 = Tr::add(c1, c4), c10 = Tr::sub(c1, c4), c8 = Tr::add(c2, c3), c9 = Tr::sub(c2, c3);
    if constexpr (Tr::kFixed) {
        f0 = Tr::mk((Sc)f0.x + ((Sc)c7.x + c8.x), (Sc)f0.y + ((Sc)c7.y + c8.y));
    } else {
        f0 = Tr::add(f0, c7);
        f0 = Tr::add(f0, c8);
    }
    const E c5 = Tr::mk((Sc)c0.x + (Tr::smul(c7.x, ya.x) + Tr::smul(c8.x, yb.x)),
                        (Sc)c0.y + (Tr::smul(c7.y, ya.x) + Tr::smul(c8.y, yb.x)));
    const E c6 = Tr::mk(Tr::smul(c10.y, ya.y) + Tr::smul(c9.y, yb.y),
                        -Tr::smul(c10.x, ya.y) - Tr::smul(c9.x, yb.y));
    F[m] = Tr::sub(c5, c6);
    F[4 * m] = Tr::add(c5, c6);
    const E c11 = Tr::mk((Sc)c0.x + (Tr::smul(c7.x, yb.x) + Tr::smul(c8.x, ya.x)),
                         (Sc)c0.y + (Tr::smul(c7.y, yb.x) + Tr::smul(c8.y, ya.x)));
    const E c12 = Tr::mk(-Tr::smul(c10.y, yb.y) + Tr::smul(c9.y, ya.y),
                         Tr::smul(c10.x, yb.y) - Tr::smul(c9.x, ya.y));
    F[2 * m] = Tr::add(c11, c12);
    F[3 * m] = Tr::sub(c11, c12);
    F[0] = f0;
}

struct FftArgs {
    const void *in;
    void *out;
    const void *tw;
    const int *scatter;
    void *scratch;
    long long batch;
    int n, inverse, nstages, tpc, has_generic;
    int radix[24], rem[24];
};

template <typename Tr, bool SMEM>
__global__ void __launch_bounds__(256) fft_staged_kernel(const FftArgs a)
{
    using E = typename Tr::E;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int n = a.n, tid = threadIdx.x, nthr = blockDim.x;
    const E *__restrict__ tw = static_cast<const E *>(a.tw);
    const long long ngroups = (a.batch + a.tpc - 1) / a.tpc;

    for (long long grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const long long first = grp * a.tpc;
        const int nt = (int)((a.batch - first) < a.tpc ? (a.batch - first) : a.tpc);
        const int total = nt * n;
        const E *in = static_cast<const E *>(a.in) + first * n;
        E *out = static_cast<E *>(a.out) + first * n;
        E *F, *scr;
        if constexpr (SMEM) {
            F = reinterpret_cast<E *>(smem_raw);
            scr = F + (size_t)a.tpc * n;
        } else {
            F = out;
            scr = static_cast<E *>(a.scratch) + (size_t)blockIdx.x * a.tpc * n;
        }

        // kf_work's leaf copy (fft/kissfft.hh:93-98, fft/kiss_fft.c:277-281) as a scatter
        for (int i = tid; i < total; i += nthr) {
            const int tb = i / n, ii = i - tb * n;
            F[tb * n + a.scatter[ii]] = in[i];
        }
        __syncthreads();

        for (int s = a.nstages - 1; s >= 0; s--) {
            const int p = a.radix[s], m = a.rem[s], span = p * m, fstride = n / span;
            if (p >= 2 && p <= 5) {
                const int per = n / p, nb = nt * per;
                E e1 = Tr::mk(0, 0), e2 = Tr::mk(0, 0);
                if (p == 3) e1 = tw[fstride * m];
                if (p == 5) { e1 = tw[fstride * m]; e2 = tw[2 * fstride * m]; }
                for (int b = tid; b < nb; b += nthr) {
                    const int tb = b / per, bb = b - tb * per;
                    const int g = bb / m, k = bb - g * m;
                    E *Fb = F + tb * n + g * span + k;
                    const int tws = k * fstride;
                    if (p == 4) bfly4<Tr, E>(Fb, m, tw, tws, a.inverse);
                    else if (p == 2) bfly2<Tr, E>(Fb, m, tw, tws);
                    else if (p == 3) bfly3<Tr, E>(Fb, m, tw, tws, e1);
                    else bfly5<Tr, E>(Fb, m, tw, tws, e1, e2);
                }
            } else {
                // kf_bfly_generic (fft/kissfft.hh:263-303, fft/kiss_fft.c:199-235), one output per thread
                for (int i = tid; i < total; i += nthr) scr[i] = Tr::fixdiv(F[i], p);
                __syncthreads();
                for (int i = tid; i < total; i += nthr) {
                    const int tb = i / n, ii = i - tb * n;
                    const int kk = ii % span, g0 = i - kk, u = kk % m;
                    E acc = scr[g0 + u];
                    int twidx = 0;
                    for (int q = 1; q < p; q++) {
                        twidx += fstride * kk;
                        if (twidx >= n) twidx -= n;
                        acc = Tr::add(acc, Tr::mul(scr[g0 + u + q * m], tw[twidx]));
                    }
                    F[i] = acc;
                }
            }
            __syncthreads();
        }

        if constexpr (SMEM) {
            for (int i = tid; i < total; i += nthr) out[i] = F[i];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------- fast path: N = 4096 ---
// kiss_fft plans 4096 as six radix-4 stages (remainders m = 1024,256,64,16,4,1, executed
// innermost first).  Here one 256-thread CTA does one transform in THREE register-resident
// passes of two fused radix-4 stages each (16 points per thread), with two shared-memory
// exchanges instead of six -- every element still sees kiss_fft's exact sequence of
// operations (twiddle products, C_FIXDIV, wraps), so the Q15 variant stays bit-exact.
//
// Slot o = 256*A + 16*B + C (A,B,C in [0,16)), input index i = digit-reversed o:
//   pass 1: thread t = i mod 256 loads in[t + 256*j] (coalesced), does stages m=1 and m=4,
//           i.e. all C for its (A,B), and writes slots 256A+16B+C
//   pass 2: thread (A,C) does stages m=16, m=64 over B
//   pass 3: thread (B,C) does stages m=256, m=1024 over A and stores out[256*A + 16B+C] (coalesced)
// Shared-memory index of slot o is o + (o >> 8): with that one-element pad per 256 every
// one of the six access patterns is bank-conflict free.
// Twiddles are pre-gathered on the host from the reference-order table into per-pass tables
// laid out in consumption order (tw1: uniform; tw2[j][C]; tw3[j][16B+C]) so loads coalesce.
// CONJ = multiply by the conjugate twiddle: lets the float overlap-save FIR run its inverse
// transform from the FORWARD tables (kissfft's inverse twiddles are exactly the conjugates for
// floats; NOT so for the Q15 table, whose floor(.5 + x) rounding is asymmetric -- the Q15 FFT
// therefore always carries its own per-direction tables).
template <typename Tr, bool CONJ, typename E>
__device__ __forceinline__ E twmul(const E f, const E t)
{
    if constexpr (CONJ) return Tr::mk(f.x * t.x + f.y * t.y, f.y * t.x - f.x * t.y);
    else return Tr::mul(f, t);
}

template <typename Tr, bool CONJ, typename E>
__device__ __forceinline__ void bfly4_reg(E &f0, E &f1, E &f2, E &f3, const E t1, const E t2, const E t3, const int inverse)
{
    f0 = Tr::fixdiv(f0, 4); f1 = Tr::fixdiv(f1, 4); f2 = Tr::fixdiv(f2, 4); f3 = Tr::fixdiv(f3, 4);
    const E s0 = twmul<Tr, CONJ, E>(f1, t1);
    const E s1 = twmul<Tr, CONJ, E>(f2, t2);
    const E s2 = twmul<Tr, CONJ, E>(f3, t3);
    const E s5 = Tr::sub(f0, s1);
    f0 = Tr::add(f0, s1);
    const E s3 = Tr::add(s0, s2);
    const E s4 = Tr::sub(s0, s2);
    f2 = Tr::sub(f0, s3);
    f0 = Tr::add(f0, s3);
    const E rot = Tr::mk(s4.y, Tr::wrap(-s4.x));
    if (inverse) { f1 = Tr::sub(s5, rot); f3 = Tr::add(s5, rot); }
    else { f1 = Tr::add(s5, rot); f3 = Tr::sub(s5, rot); }
}

// same butterfly when all three twiddles are tw[0] = 1: floats skip the (exact) products,
// Q15 must still do them (tw[0] = 32767/32768)
template <typename Tr, typename E>
__device__ __forceinline__ void bfly4_unit(E &f0, E &f1, E &f2, E &f3, const E one, const int inverse)
{
    if constexpr (Tr::kFixed) {
        bfly4_reg<Tr, false, E>(f0, f1, f2, f3, one, one, one, inverse);
    } else {
        const E s5 = Tr::sub(f0, f2);
        f0 = Tr::add(f0, f2);
        const E s3 = Tr::add(f1, f3);
        const E s4 = Tr::sub(f1, f3);
        f2 = Tr::sub(f0, s3);
        f0 = Tr::add(f0, s3);
        const E rot = Tr::mk(s4.y, -s4.x);
        if (inverse) { f1 = Tr::sub(s5, rot); f3 = Tr::add(s5, rot); }
        else { f1 = Tr::add(s5, rot); f3 = Tr::sub(s5, rot); }
    }
}

struct Fft4096Args {
    const void *in;
    void *out;
    const void *tw1;   // [16]      tw[256*k*q], entry k*4+q (q=0 unused -> tw[0])
    const void *tw2;   // [15][16]  j<3: tw[64*C*(j+1)]; j=3+3*bm+(q-1): tw[(16*bm+C)*16*q]
    const void *tw3;   // [15][256] j<3: tw[4*kk*(j+1)]; j=3+3*am+(q-1): tw[(256*am+kk)*q]
    long long batch;
    int inverse;
};

// The three passes on one transform.  In: v[4*k4 + k5] = x[t + 256*(k4 + 4*k5)] (thread t of
// 256).  Out: v[A] = X[256*A + t].  F is the CTA's 4096+16 element exchange buffer.
template <typename Tr, bool CONJ, typename E>
__device__ __forceinline__ void fft4096_core(E (&v)[16], typename Tr::S *F, const typename Tr::TwS *__restrict__ tw1_s,
                                             const typename Tr::TwS *__restrict__ tw2_s, const typename Tr::TwS *__restrict__ tw3_s,
                                             const int t, const int inverse)
{
    // twiddle tables and the exchange buffer hold the storage type; registers hold Tr::E
    struct Tw { const typename Tr::TwS *__restrict__ p; __device__ __forceinline__ E operator[](int i) const { return Tr::ldtw(p[i]); } };
    const Tw tw1{tw1_s}, tw2{tw2_s}, tw3{tw3_s};
    // pass-1 slot base: t = k0 + 4k1 + 16k2 + 64k3  ->  A = 4k0 + k1, B = 4k2 + k3
    const int A1 = ((t & 3) << 2) | ((t >> 2) & 3), B1 = (((t >> 4) & 3) << 2) | ((t >> 6) & 3);
    const int base1 = 257 * A1 + 16 * B1;
    // pass 2: thread = 16*A + C ; pass 3: thread = 16*B + C = kk
    const int C2 = t & 15;
    const int base2 = 257 * (t >> 4) + C2;
    // ---- pass 1: stages m=1 (over k5) and m=4 (over k4)
    {
        const E one = tw1[0];
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) bfly4_unit<Tr, E>(v[4 * k4], v[4 * k4 + 1], v[4 * k4 + 2], v[4 * k4 + 3], one, inverse);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            if (k == 0) bfly4_unit<Tr, E>(v[0], v[4], v[8], v[12], one, inverse);
            else bfly4_reg<Tr, CONJ, E>(v[k], v[4 + k], v[8 + k], v[12 + k], tw1[4 * k + 1], tw1[4 * k + 2], tw1[4 * k + 3], inverse);
        }
    }
    __syncthreads();   // earlier readers of F are done
#pragma unroll
    for (int c = 0; c < 16; c++) F[base1 + c] = Tr::st(v[c]);
    __syncthreads();
    // ---- pass 2: stages m=16 (k = C, q = B mod 4) and m=64 (k = 16*(B mod 4) + C, q = B div 4)
#pragma unroll
    for (int b = 0; b < 16; b++) v[b] = Tr::ld(F[base2 + 16 * b]);
    {
        const E t1 = tw2[0 * 16 + C2], t2 = tw2[1 * 16 + C2], t3 = tw2[2 * 16 + C2];
#pragma unroll
        for (int bd = 0; bd < 4; bd++) bfly4_reg<Tr, CONJ, E>(v[4 * bd], v[4 * bd + 1], v[4 * bd + 2], v[4 * bd + 3], t1, t2, t3, inverse);
#pragma unroll
        for (int bm = 0; bm < 4; bm++)
            bfly4_reg<Tr, CONJ, E>(v[bm], v[4 + bm], v[8 + bm], v[12 + bm], tw2[(3 + 3 * bm) * 16 + C2], tw2[(4 + 3 * bm) * 16 + C2],
                                   tw2[(5 + 3 * bm) * 16 + C2], inverse);
    }
#pragma unroll
    for (int b = 0; b < 16; b++) F[base2 + 16 * b] = Tr::st(v[b]);
    __syncthreads();
    // ---- pass 3: stages m=256 (k = kk, q = A mod 4) and m=1024 (k = 256*(A mod 4) + kk, q = A div 4)
#pragma unroll
    for (int A = 0; A < 16; A++) v[A] = Tr::ld(F[257 * A + t]);
    {
        const E t1 = tw3[0 * 256 + t], t2 = tw3[1 * 256 + t], t3 = tw3[2 * 256 + t];
#pragma unroll
        for (int ad = 0; ad < 4; ad++) bfly4_reg<Tr, CONJ, E>(v[4 * ad], v[4 * ad + 1], v[4 * ad + 2], v[4 * ad + 3], t1, t2, t3, inverse);
#pragma unroll
        for (int am = 0; am < 4; am++)
            bfly4_reg<Tr, CONJ, E>(v[am], v[4 + am], v[8 + am], v[12 + am], tw3[(3 + 3 * am) * 256 + t], tw3[(4 + 3 * am) * 256 + t],
                                   tw3[(5 + 3 * am) * 256 + t], inverse);
    }
}

// INVT: -1 = direction read from the arguments; 0 / 1 = compiled in (the Q15 kernel is instruction bound: a run-time
// direction costs ~100 predicated instructions per transform pass set)
template <typename Tr, int INVT = -1>
__global__ void __launch_bounds__(256, 3) fft4096_kernel(const Fft4096Args a)
{
    const int inverse = INVT < 0 ? a.inverse : INVT;
    using E = typename Tr::E;
    using S = typename Tr::S;
    __shared__ S F[4096 + 16];
    const int t = threadIdx.x;
    using TwS = typename Tr::TwS;
    const TwS *__restrict__ tw1 = static_cast<const TwS *>(a.tw1);
    const TwS *__restrict__ tw2 = static_cast<const TwS *>(a.tw2);
    const TwS *__restrict__ tw3 = static_cast<const TwS *>(a.tw3);
    for (long long xf = blockIdx.x; xf < a.batch; xf += gridDim.x) {
        const S *in = static_cast<const S *>(a.in) + xf * 4096;
        S *out = static_cast<S *>(a.out) + xf * 4096;
        E v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[4 * (j & 3) + (j >> 2)] = Tr::ld(in[t + 256 * j]);   // coalesced, digit-reversed by register index
        fft4096_core<Tr, false, E>(v, F, tw1, tw2, tw3, t, inverse);
#pragma unroll
        for (int A = 0; A < 16; A++) out[256 * A + t] = Tr::st(v[A]);
    }
}

// ------------------------------------------------------------------------ host: plan ---
static int make_plan(FftPlan &p)
{
    // fft/kissfft.hh:38-55 (float/double) and fft/kiss_fft.c:308-330 (int16)
    int n = p.n, r = 4, s = 0;
    const double floor_sqrt = std::floor(std::sqrt((double)p.n));
    do {
        while (n % r) {
            switch (r) { case 4: r = 2; break; case 2: r = 3; break; default: r += 2; break; }
            if (p.dtype == B200C_CI16) { if (r > floor_sqrt) r = n; }
            else { if ((long long)r * r > n) r = n; }
        }
        n /= r;
        if (s >= 24) { set_error("FFT: too many radix stages"); return B200C_ERR_UNSUPPORTED; }
        p.radix[s] = r; p.rem[s] = n; s++;
    } while (n > 1);
    p.nstages = s;
    p.has_generic = false;
    for (int i = 0; i < s; i++) if (p.radix[i] > 5 || p.radix[i] < 2) p.has_generic = true;   // n = 1 plans a radix-1 stage
    return B200C_OK;
}

int fft_plan_create(FftPlan &p, int dtype, size_t nbins, int inverse, size_t smem_budget)
{
    if (dtype != B200C_CF32 && dtype != B200C_CF64 && dtype != B200C_CI16) {
        set_error("FFTFactory(dtype=%d): unsupported type", dtype);
        return B200C_ERR_UNSUPPORTED;
    }
    if (nbins < 1 || nbins > (1u << 22)) { set_error("FFT: numBins %zu out of range [1, 2^22]", nbins); return B200C_ERR_INVALID; }
    p.dtype = dtype; p.n = (int)nbins; p.inverse = inverse ? 1 : 0;
    int rc = make_plan(p);
    if (rc) return rc;
    const int n = p.n;
    const size_t esz = dtype_bytes(dtype);

    // scatter[input index] = output slot  (inverse of the leaf-copy permutation)
    std::vector<int> scatter(n);
    for (int o = 0; o < n; o++) {
        int remn = o, idx = 0, fstride = 1;
        for (int s = 0; s < p.nstages; s++) {
            const int k = remn / p.rem[s];
            remn -= k * p.rem[s];
            idx += k * fstride;
            fstride *= p.radix[s];
        }
        scatter[idx] = o;
    }

    // twiddles
    std::vector<uint8_t> tw((size_t)n * esz);
    if (dtype == B200C_CF32) {
        // fft/kissfft.hh:21-26: acos((T)-1) resolves to ::acos(double); phinc rounded once to
        // float, phase i*phinc formed in float, std::exp(std::complex<float>)
        const float phinc = (float)((inverse ? 2 : -2) * std::acos((double)-1) / n);
        auto *t = reinterpret_cast<std::complex<float> *>(tw.data());
        for (int i = 0; i < n; i++) t[i] = std::exp(std::complex<float>(0, i * phinc));
    } else if (dtype == B200C_CF64) {
        const double phinc = (inverse ? 2 : -2) * std::acos((double)-1) / n;
        auto *t = reinterpret_cast<std::complex<double> *>(tw.data());
        for (int i = 0; i < n; i++) t[i] = std::exp(std::complex<double>(0, i * phinc));
    } else {
        // fft/kiss_fft.c:357-363 + fft/_kiss_fft_guts.h:128-129
        auto *t = reinterpret_cast<int16_t *>(tw.data());
        for (int i = 0; i < n; i++) {
            const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
            double phase = -2 * pi * i / n;
            if (inverse) phase *= -1;
            t[2 * i] = (int16_t)std::floor(.5 + 32767 * std::cos(phase));
            t[2 * i + 1] = (int16_t)std::floor(.5 + 32767 * std::sin(phase));
        }
    }

    B200C_CUDA_TRY(cudaMalloc(&p.d_tw, tw.size()));
    B200C_CUDA_TRY(cudaMalloc((void **)&p.d_scatter, sizeof(int) * n));
    B200C_CUDA_TRY(cudaMemcpy(p.d_tw, tw.data(), tw.size(), cudaMemcpyHostToDevice));
    B200C_CUDA_TRY(cudaMemcpy(p.d_scatter, scatter.data(), sizeof(int) * n, cudaMemcpyHostToDevice));

    // fast path tables: gathered from the reference-order twiddle table (exact same values)
    p.fast = 0;
    if (n == 4096 && (dtype == B200C_CF32 || dtype == B200C_CI16)) {
        auto at = [&](int idx) { return tw.data() + (size_t)(idx % n) * esz; };
        std::vector<uint8_t> t1(16 * esz), t2(15 * 16 * esz), t3(15 * 256 * esz);
        for (int k = 0; k < 4; k++)
            for (int q = 0; q < 4; q++) std::memcpy(&t1[(k * 4 + q) * esz], at(q == 0 ? 0 : 256 * k * q), esz);
        for (int C = 0; C < 16; C++) {
            for (int q = 1; q < 4; q++) std::memcpy(&t2[((q - 1) * 16 + C) * esz], at(64 * C * q), esz);
            for (int bm = 0; bm < 4; bm++)
                for (int q = 1; q < 4; q++) std::memcpy(&t2[((3 + 3 * bm + q - 1) * 16 + C) * esz], at((16 * bm + C) * 16 * q), esz);
        }
        for (int kk = 0; kk < 256; kk++) {
            for (int q = 1; q < 4; q++) std::memcpy(&t3[((q - 1) * 256 + kk) * esz], at(4 * kk * q), esz);
            for (int am = 0; am < 4; am++)
                for (int q = 1; q < 4; q++) std::memcpy(&t3[((3 + 3 * am + q - 1) * 256 + kk) * esz], at((256 * am + kk) * q), esz);
        }
        B200C_CUDA_TRY(cudaMalloc(&p.d_fast[0], t1.size()));
        B200C_CUDA_TRY(cudaMalloc(&p.d_fast[1], t2.size()));
        B200C_CUDA_TRY(cudaMalloc(&p.d_fast[2], t3.size()));
        B200C_CUDA_TRY(cudaMemcpy(p.d_fast[0], t1.data(), t1.size(), cudaMemcpyHostToDevice));
        B200C_CUDA_TRY(cudaMemcpy(p.d_fast[1], t2.data(), t2.size(), cudaMemcpyHostToDevice));
        B200C_CUDA_TRY(cudaMemcpy(p.d_fast[2], t3.data(), t3.size(), cudaMemcpyHostToDevice));
        if (dtype == B200C_CI16) {
            // the lazy-wrap kernel reads them unpacked: (re, im) as two int32
            const std::vector<uint8_t> *src[3] = {&t1, &t2, &t3};
            for (int k = 0; k < 3; k++) {
                const size_t cnt = src[k]->size() / 4;
                std::vector<int32_t> w(2 * cnt);
                const int16_t *s16 = reinterpret_cast<const int16_t *>(src[k]->data());
                for (size_t i = 0; i < 2 * cnt; i++) w[i] = s16[i];
                B200C_CUDA_TRY(cudaMalloc(&p.d_fastw[k], w.size() * sizeof(int32_t)));
                B200C_CUDA_TRY(cudaMemcpy(p.d_fastw[k], w.data(), w.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
            }
        }
        p.fast = 4096;
    }

    // execution shape
    const size_t bufs = p.has_generic ? 2 : 1;
    p.tpc = std::max(1, 2048 / n);
    p.smem_bytes = (size_t)p.tpc * n * esz * bufs;
    p.smem = p.smem_bytes <= smem_budget;
    if (!p.smem) { p.tpc = 1; p.smem_bytes = 0; }
    const int work = p.tpc * n / 4;
    p.threads = std::min(256, std::max(32, (work + 31) / 32 * 32));
    return B200C_OK;
}

void fft_plan_destroy(FftPlan &p)
{
    if (p.d_tw) cudaFree(p.d_tw);
    if (p.d_scatter) cudaFree(p.d_scatter);
    if (p.d_scratch) cudaFree(p.d_scratch);
    for (auto &f : p.d_fast) { if (f) cudaFree(f); f = nullptr; }
    for (auto &f : p.d_fastw) { if (f) cudaFree(f); f = nullptr; }
    p.d_tw = nullptr; p.d_scatter = nullptr; p.d_scratch = nullptr;
}

template <typename Tr>
static int launch_staged(FftPlan &p, const FftArgs &a, int grid, cudaStream_t stream)
{
    if (p.smem) {
        auto kern = fft_staged_kernel<Tr, true>;
        if (p.smem_bytes > 48 * 1024)
            B200C_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024)));
        kern<<<grid, p.threads, p.smem_bytes, stream>>>(a);
    } else {
        fft_staged_kernel<Tr, false><<<grid, p.threads, 0, stream>>>(a);
    }
    B200C_CUDA_TRY(cudaGetLastError());
    return B200C_OK;
}

int fft_launch(FftPlan &p, const void *d_in, void *d_out, size_t batch, int sm_count, cudaStream_t stream)
{
    if (batch == 0) return B200C_OK;
    if (d_in == d_out) { set_error("FFT: in-place transforms are not supported (d_in == d_out)"); return B200C_ERR_INVALID; }
    if (p.fast == 4096 && !p.force_staged) {
        Fft4096Args f;
        f.in = d_in; f.out = d_out; f.tw1 = p.d_fast[0]; f.tw2 = p.d_fast[1]; f.tw3 = p.d_fast[2];
        f.batch = (long long)batch; f.inverse = p.inverse;
        const int grid = (int)std::min<long long>((long long)batch, (long long)sm_count * 12);
        if (p.dtype == B200C_CF32) fft4096_kernel<FloatTraits<float>><<<grid, 256, 0, stream>>>(f);
        else {
            // lazy-wrap Q15 (bit-identical; B200C_FFT_Q15=packed keeps the per-assignment wrap of round 1 for A/B runs)
            static const bool packed = [] { const char *e = std::getenv("B200C_FFT_Q15"); return e && std::strcmp(e, "packed") == 0; }();
            if (!packed) { f.tw1 = p.d_fastw[0]; f.tw2 = p.d_fastw[1]; f.tw3 = p.d_fastw[2]; }
            if (packed) fft4096_kernel<Q15Traits><<<grid, 256, 0, stream>>>(f);
            else if (f.inverse) fft4096_kernel<Q15Lazy, 1><<<grid, 256, 0, stream>>>(f);
            else fft4096_kernel<Q15Lazy, 0><<<grid, 256, 0, stream>>>(f);
        }
        B200C_CUDA_TRY(cudaGetLastError());
        return B200C_OK;
    }
    FftArgs a;
    a.in = d_in; a.out = d_out; a.tw = p.d_tw; a.scatter = p.d_scatter; a.scratch = nullptr;
    a.batch = (long long)batch; a.n = p.n; a.inverse = p.inverse; a.nstages = p.nstages; a.tpc = p.tpc;
    a.has_generic = p.has_generic ? 1 : 0;
    for (int i = 0; i < p.nstages; i++) { a.radix[i] = p.radix[i]; a.rem[i] = p.rem[i]; }
    const long long ngroups = ((long long)batch + p.tpc - 1) / p.tpc;
    int grid = (int)std::min<long long>(ngroups, (long long)sm_count * 8);
    if (!p.smem && p.has_generic) {
        const size_t need = (size_t)grid * p.n * dtype_bytes(p.dtype);
        if (need > p.scratch_bytes) {
            B200C_CUDA_TRY(cudaStreamSynchronize(stream));
            if (p.d_scratch) cudaFree(p.d_scratch);
            p.d_scratch = nullptr; p.scratch_bytes = 0;
            B200C_CUDA_TRY(cudaMalloc(&p.d_scratch, need));
            p.scratch_bytes = need;
        }
        a.scratch = p.d_scratch;
    }
    switch (p.dtype) {
    case B200C_CF32: return launch_staged<FloatTraits<float>>(p, a, grid, stream);
    case B200C_CF64: return launch_staged<FloatTraits<double>>(p, a, grid, stream);
    default: return launch_staged<Q15Traits>(p, a, grid, stream);
    }
}

} // namespace b200c
