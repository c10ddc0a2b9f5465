/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the /comms/fft hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may call this.  The product (libb200comms.so) never links or loads it.
 *
 * An independent, ITERATIVE restatement of the two kiss_fft variants the reference ships
 * (the reference recurses; here: mixed-radix digit reversal, then stages innermost first):
 *   float/double : fft/kissfft.hh:21-56 (twiddles in scalar precision, radix plan 4,2,3,5,7..),
 *                  :87-120 (decimation in time), :132-303 (butterflies 2/4/3/5/generic)
 *   complex int16: fft/kiss_fft.c:339-368 (Q15 twiddles from double phase), :308-330 (plan),
 *                  :21-235 (butterflies with C_FIXDIV 1/radix scaling), fft/_kiss_fft_guts.h:44-129
 * PINNED: tests/test_oracle_fft.py checks this restatement bit-for-bit (int16) and
 * bit-for-bit/ulp-level (float, same libm) against oracle/_ref (the reference's own sources
 * compiled here) and against the N=4 goldens of fft/TestFFT.cpp:19-29,95-105.
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define MAXF 64

typedef struct { int n, nstages, p[MAXF], m[MAXF]; } plan_t;

/* kissfft.hh:38-55: rule "if (p*p > n) p = n" on the REMAINING n */
static void plan_hh(plan_t *pl, int nfft)
{
    int n = nfft, p = 4, s = 0;
    do {
        while (n % p) {
            switch (p) { case 4: p = 2; break; case 2: p = 3; break; default: p += 2; break; }
            if (p * p > n) p = n;
        }
        n /= p;
        pl->p[s] = p; pl->m[s] = n; s++;
    } while (n > 1);
    pl->n = nfft; pl->nstages = s;
}

/* kiss_fft.c:308-330: rule "if (p > floor(sqrt(ORIGINAL n))) p = n" */
static void plan_c(plan_t *pl, int nfft)
{
    int n = nfft, p = 4, s = 0;
    const double floor_sqrt = floor(sqrt((double)nfft));
    do {
        while (n % p) {
            switch (p) { case 4: p = 2; break; case 2: p = 3; break; default: p += 2; break; }
            if (p > floor_sqrt) p = n;
        }
        n /= p;
        pl->p[s] = p; pl->m[s] = n; s++;
    } while (n > 1);
    pl->n = nfft; pl->nstages = s;
}

/* output slot o = sum k_s*m_s  <-  input index sum k_s*fstride_s  (kf_work leaf copy) */
static void plan_perm(const plan_t *pl, int *perm)
{
    for (int o = 0; o < pl->n; o++) {
        int rem = o, idx = 0, fstride = 1;
        for (int s = 0; s < pl->nstages; s++) {
            const int k = rem / pl->m[s];
            rem -= k * pl->m[s];
            idx += k * fstride;
            fstride *= pl->p[s];
        }
        perm[o] = idx;
    }
}

int oracle_fft_plan(int nfft, int fixed, int *radix, int *remainder)
{
    plan_t pl;
    if (nfft < 1) return -1;
    if (fixed) plan_c(&pl, nfft); else plan_hh(&pl, nfft);
    for (int s = 0; s < pl.nstages; s++) { radix[s] = pl.p[s]; remainder[s] = pl.m[s]; }
    return pl.nstages;
}

/* ------------------------------------------------------------------ float / double ---- */
#define FFT_FLOAT(NAME, T, CEXP, ACOS)                                                          \
    static void NAME##_twiddles(T _Complex *tw, int nfft, int inverse)                          \
    {                                                                                           \
        /* kissfft.hh:21-26 -- the unqualified acos((T)-1) there resolves to ::acos(double), so   \
         * the increment is formed in double and rounded once to T; the phase i*phinc is in T */  \
        const T phinc = (T)((inverse ? 2 : -2) * ACOS((double)-1) / nfft);                      \
        for (int i = 0; i < nfft; i++) tw[i] = CEXP(__builtin_complex((T)0, (T)(i * phinc)));   \
    }                                                                                           \
    static void NAME##_stage(T _Complex *F, const T _Complex *tw, int N, int p, int m,          \
                             int fstride, int inverse, T _Complex *scratch)                     \
    {                                                                                           \
        switch (p) {                                                                            \
        case 2:                                                                                 \
            for (int k = 0; k < m; k++) {                                                       \
                const T _Complex t = F[m + k] * tw[k * fstride];                                \
                F[m + k] = F[k] - t;                                                            \
                F[k] += t;                                                                      \
            }                                                                                   \
            break;                                                                              \
        case 4: {                                                                               \
            const T neg = inverse ? -1 : 1;                                                     \
            for (int k = 0; k < m; k++) {                                                       \
                const T _Complex s0 = F[k + m] * tw[k * fstride];                               \
                const T _Complex s1 = F[k + 2 * m] * tw[k * fstride * 2];                       \
                const T _Complex s2 = F[k + 3 * m] * tw[k * fstride * 3];                       \
                const T _Complex s5 = F[k] - s1;                                                \
                F[k] += s1;                                                                     \
                const T _Complex s3 = s0 + s2;                                                  \
                T _Complex s4 = s0 - s2;                                                        \
                s4 = __builtin_complex((T)(__imag__ s4 * neg), (T)(-__real__ s4 * neg));        \
                F[k + 2 * m] = F[k] - s3;                                                       \
                F[k] += s3;                                                                     \
                F[k + m] = s5 + s4;                                                             \
                F[k + 3 * m] = s5 - s4;                                                         \
            }                                                                                   \
        } break;                                                                                \
        case 3: {                                                                               \
            const T _Complex epi3 = tw[fstride * m];                                            \
            for (int k = 0; k < m; k++) {                                                       \
                const T _Complex s1 = F[k + m] * tw[k * fstride];                               \
                const T _Complex s2 = F[k + 2 * m] * tw[k * fstride * 2];                       \
                const T _Complex s3 = s1 + s2;                                                  \
                T _Complex s0 = s1 - s2;                                                        \
                F[k + m] = __builtin_complex((T)(__real__ F[k] - (T)(__real__ s3 * .5)),        \
                                             (T)(__imag__ F[k] - (T)(__imag__ s3 * .5)));       \
                s0 = __builtin_complex((T)(__real__ s0 * __imag__ epi3), (T)(__imag__ s0 * __imag__ epi3)); \
                F[k] += s3;                                                                     \
                F[k + 2 * m] = __builtin_complex((T)(__real__ F[k + m] + __imag__ s0),          \
                                                 (T)(__imag__ F[k + m] - __real__ s0));         \
                F[k + m] += __builtin_complex((T)(-__imag__ s0), (T)(__real__ s0));             \
            }                                                                                   \
        } break;                                                                                \
        case 5: {                                                                               \
            const T _Complex ya = tw[fstride * m], yb = tw[fstride * 2 * m];                    \
            const T yar = __real__ ya, yai = __imag__ ya, ybr = __real__ yb, ybi = __imag__ yb; \
            for (int u = 0; u < m; u++) {                                                       \
                T _Complex *F0 = F + u, *F1 = F0 + m, *F2 = F0 + 2 * m, *F3 = F0 + 3 * m, *F4 = F0 + 4 * m; \
                const T _Complex c0 = *F0;                                                      \
                const T _Complex c1 = *F1 * tw[u * fstride];                                    \
                const T _Complex c2 = *F2 * tw[2 * u * fstride];                                \
                const T _Complex c3 = *F3 * tw[3 * u * fstride];                                \
                const T _Complex c4 = *F4 * tw[4 * u * fstride];                                \
                const T _Complex c7 = c1 + c4, c10 = c1 - c4, c8 = c2 + c3, c9 = c2 - c3;       \
                *F0 += c7;                                                                      \
                *F0 += c8;                                                                      \
                const T _Complex c5 = c0 + __builtin_complex(                                   \
                    (T)(__real__ c7 * yar + __real__ c8 * ybr), (T)(__imag__ c7 * yar + __imag__ c8 * ybr)); \
                const T _Complex c6 = __builtin_complex(                                        \
                    (T)(__imag__ c10 * yai + __imag__ c9 * ybi), (T)(-(__real__ c10 * yai) - __real__ c9 * ybi)); \
                *F1 = c5 - c6;                                                                  \
                *F4 = c5 + c6;                                                                  \
                const T _Complex c11 = c0 + __builtin_complex(                                  \
                    (T)(__real__ c7 * ybr + __real__ c8 * yar), (T)(__imag__ c7 * ybr + __imag__ c8 * yar)); \
                const T _Complex c12 = __builtin_complex(                                       \
                    (T)(-(__imag__ c10 * ybi) + __imag__ c9 * yai), (T)(__real__ c10 * ybi - __real__ c9 * yai)); \
                *F2 = c11 + c12;                                                                \
                *F3 = c11 - c12;                                                                \
            }                                                                                   \
        } break;                                                                                \
        default:                                                                                \
            for (int u = 0; u < m; u++) {                                                       \
                int k = u;                                                                      \
                for (int q1 = 0; q1 < p; q1++) { scratch[q1] = F[k]; k += m; }                  \
                k = u;                                                                          \
                for (int q1 = 0; q1 < p; q1++) {                                                \
                    int twidx = 0;                                                              \
                    F[k] = scratch[0];                                                          \
                    for (int q = 1; q < p; q++) {                                               \
                        twidx += fstride * k;                                                   \
                        if (twidx >= N) twidx -= N;                                             \
                        const T _Complex t = scratch[q] * tw[twidx];                            \
                        F[k] += t;                                                              \
                    }                                                                           \
                    k += m;                                                                     \
                }                                                                               \
            }                                                                                   \
        }                                                                                       \
    }                                                                                           \
    int oracle_fft_##NAME(int nfft, int inverse, const void *vin, void *vout, size_t batch)     \
    {                                                                                           \
        if (nfft < 1) return -1;                                                                \
        plan_t pl;                                                                              \
        plan_hh(&pl, nfft);                                                                     \
        int *perm = (int *)malloc(sizeof(int) * nfft);                                          \
        T _Complex *tw = (T _Complex *)malloc(sizeof(T _Complex) * nfft);                       \
        T _Complex *scratch = (T _Complex *)malloc(sizeof(T _Complex) * nfft);                  \
        plan_perm(&pl, perm);                                                                   \
        NAME##_twiddles(tw, nfft, inverse);                                                     \
        for (size_t b = 0; b < batch; b++) {                                                    \
            const T _Complex *in = (const T _Complex *)vin + b * nfft;                          \
            T _Complex *F = (T _Complex *)vout + b * nfft;                                      \
            for (int o = 0; o < nfft; o++) F[o] = in[perm[o]];                                  \
            for (int s = pl.nstages - 1; s >= 0; s--) {                                         \
                const int p = pl.p[s], m = pl.m[s], span = p * m, fstride = nfft / span;        \
                for (int g = 0; g < nfft; g += span) NAME##_stage(F + g, tw, nfft, p, m, fstride, inverse, scratch); \
            }                                                                                   \
        }                                                                                       \
        free(perm); free(tw); free(scratch);                                                    \
        return 0;                                                                               \
    }

FFT_FLOAT(cf32, float, cexpf, acos)
FFT_FLOAT(cf64, double, cexp, acos)

/* ------------------------------------------------------------- Q15 complex int16 ------ */
typedef struct { int16_t r, i; } c16;

/* _kiss_fft_guts.h:64-71 */
static inline int16_t sround(int32_t x) { return (int16_t)((x + (1 << 14)) >> 15); }
static inline int16_t s_mul(int16_t a, int16_t b) { return sround((int32_t)a * b); }
static inline c16 c_mul(c16 a, c16 b)
{
    c16 m;
    m.r = sround((int32_t)((uint32_t)((int32_t)a.r * b.r) - (uint32_t)((int32_t)a.i * b.i)));
    m.i = sround((int32_t)((uint32_t)((int32_t)a.r * b.i) + (uint32_t)((int32_t)a.i * b.r)));
    return m;
}
/* guts:73-78  C_FIXDIV(c, div): c = sround(c * (SAMP_MAX / div)) */
static inline c16 c_fixdiv(c16 c, int div)
{
    const int32_t f = 32767 / div;
    c16 o = { sround((int32_t)c.r * f), sround((int32_t)c.i * f) };
    return o;
}
static inline c16 c_add(c16 a, c16 b) { c16 o = { (int16_t)(a.r + b.r), (int16_t)(a.i + b.i) }; return o; }
static inline c16 c_sub(c16 a, c16 b) { c16 o = { (int16_t)(a.r - b.r), (int16_t)(a.i - b.i) }; return o; }

static void q15_stage(c16 *F, const c16 *tw, int N, int p, int m, int fstride, int inverse, c16 *scratch)
{
    switch (p) {
    case 2: /* kiss_fft.c:21-42 */
        for (int k = 0; k < m; k++) {
            F[k] = c_fixdiv(F[k], 2);
            F[k + m] = c_fixdiv(F[k + m], 2);
            const c16 t = c_mul(F[k + m], tw[k * fstride]);
            F[k + m] = c_sub(F[k], t);
            F[k] = c_add(F[k], t);
        }
        break;
    case 4: /* kiss_fft.c:44-90 */
        for (int k = 0; k < m; k++) {
            c16 f0 = c_fixdiv(F[k], 4), f1 = c_fixdiv(F[k + m], 4), f2 = c_fixdiv(F[k + 2 * m], 4), f3 = c_fixdiv(F[k + 3 * m], 4);
            const c16 s0 = c_mul(f1, tw[k * fstride]);
            const c16 s1 = c_mul(f2, tw[k * fstride * 2]);
            const c16 s2 = c_mul(f3, tw[k * fstride * 3]);
            const c16 s5 = c_sub(f0, s1);
            f0 = c_add(f0, s1);
            const c16 s3 = c_add(s0, s2);
            const c16 s4 = c_sub(s0, s2);
            F[k + 2 * m] = c_sub(f0, s3);
            F[k] = c_add(f0, s3);
            if (inverse) {
                F[k + m].r = (int16_t)(s5.r - s4.i); F[k + m].i = (int16_t)(s5.i + s4.r);
                F[k + 3 * m].r = (int16_t)(s5.r + s4.i); F[k + 3 * m].i = (int16_t)(s5.i - s4.r);
            } else {
                F[k + m].r = (int16_t)(s5.r + s4.i); F[k + m].i = (int16_t)(s5.i - s4.r);
                F[k + 3 * m].r = (int16_t)(s5.r - s4.i); F[k + 3 * m].i = (int16_t)(s5.i + s4.r);
            }
        }
        break;
    case 3: { /* kiss_fft.c:92-134 */
        const c16 epi3 = tw[fstride * m];
        for (int k = 0; k < m; k++) {
            c16 f0 = c_fixdiv(F[k], 3), f1 = c_fixdiv(F[k + m], 3), f2 = c_fixdiv(F[k + 2 * m], 3);
            const c16 s1 = c_mul(f1, tw[k * fstride]);
            const c16 s2 = c_mul(f2, tw[k * fstride * 2]);
            const c16 s3 = c_add(s1, s2);
            c16 s0 = c_sub(s1, s2);
            f1.r = (int16_t)(f0.r - (s3.r >> 1));
            f1.i = (int16_t)(f0.i - (s3.i >> 1));
            s0.r = s_mul(s0.r, epi3.i);
            s0.i = s_mul(s0.i, epi3.i);
            f0 = c_add(f0, s3);
            f2.r = (int16_t)(f1.r + s0.i);
            f2.i = (int16_t)(f1.i - s0.r);
            f1.r = (int16_t)(f1.r - s0.i);
            f1.i = (int16_t)(f1.i + s0.r);
            F[k] = f0; F[k + m] = f1; F[k + 2 * m] = f2;
        }
    } break;
    case 5: { /* kiss_fft.c:136-197 */
        const c16 ya = tw[fstride * m], yb = tw[fstride * 2 * m];
        for (int u = 0; u < m; u++) {
            c16 f0 = c_fixdiv(F[u], 5), f1 = c_fixdiv(F[u + m], 5), f2 = c_fixdiv(F[u + 2 * m], 5),
                f3 = c_fixdiv(F[u + 3 * m], 5), f4 = c_fixdiv(F[u + 4 * m], 5);
            const c16 c0 = f0;
            const c16 c1 = c_mul(f1, tw[u * fstride]);
            const c16 c2 = c_mul(f2, tw[2 * u * fstride]);
            const c16 c3 = c_mul(f3, tw[3 * u * fstride]);
            const c16 c4 = c_mul(f4, tw[4 * u * fstride]);
            const c16 c7 = c_add(c1, c4), c10 = c_sub(c1, c4), c8 = c_add(c2, c3), c9 = c_sub(c2, c3);
            f0.r = (int16_t)(f0.r + (c7.r + c8.r));
            f0.i = (int16_t)(f0.i + (c7.i + c8.i));
            c16 c5, c6, c11, c12;
            c5.r = (int16_t)(c0.r + s_mul(c7.r, ya.r) + s_mul(c8.r, yb.r));
            c5.i = (int16_t)(c0.i + s_mul(c7.i, ya.r) + s_mul(c8.i, yb.r));
            c6.r = (int16_t)(s_mul(c10.i, ya.i) + s_mul(c9.i, yb.i));
            c6.i = (int16_t)(-s_mul(c10.r, ya.i) - s_mul(c9.r, yb.i));
            F[u + m] = c_sub(c5, c6);
            F[u + 4 * m] = c_add(c5, c6);
            c11.r = (int16_t)(c0.r + s_mul(c7.r, yb.r) + s_mul(c8.r, ya.r));
            c11.i = (int16_t)(c0.i + s_mul(c7.i, yb.r) + s_mul(c8.i, ya.r));
            c12.r = (int16_t)(-s_mul(c10.i, yb.i) + s_mul(c9.i, ya.i));
            c12.i = (int16_t)(s_mul(c10.r, yb.i) - s_mul(c9.r, ya.i));
            F[u + 2 * m] = c_add(c11, c12);
            F[u + 3 * m] = c_sub(c11, c12);
            F[u] = f0;
        }
    } break;
    default: /* kiss_fft.c:199-235 */
        for (int u = 0; u < m; u++) {
            int k = u;
            for (int q1 = 0; q1 < p; q1++) { scratch[q1] = c_fixdiv(F[k], p); k += m; }
            k = u;
            for (int q1 = 0; q1 < p; q1++) {
                int twidx = 0;
                F[k] = scratch[0];
                for (int q = 1; q < p; q++) {
                    twidx += fstride * k;
                    if (twidx >= N) twidx -= N;
                    F[k] = c_add(F[k], c_mul(scratch[q], tw[twidx]));
                }
                k += m;
            }
        }
    }
}

/* kiss_fft.c:357-363 + guts:128-129: Q15 twiddles, phase in double, floor(.5 + 32767*cos) */
void oracle_fft_q15_twiddles(int nfft, int inverse, int16_t *tw_ri)
{
    for (int i = 0; i < nfft; i++) {
        const double pi = 3.141592653589793238462643383279502884197169399375105820974944;
        double phase = -2 * pi * i / nfft;
        if (inverse) phase *= -1;
        tw_ri[2 * i] = (int16_t)floor(.5 + 32767 * cos(phase));
        tw_ri[2 * i + 1] = (int16_t)floor(.5 + 32767 * sin(phase));
    }
}

int oracle_fft_ci16(int nfft, int inverse, const void *vin, void *vout, size_t batch)
{
    if (nfft < 1) return -1;
    plan_t pl;
    plan_c(&pl, nfft);
    int *perm = (int *)malloc(sizeof(int) * nfft);
    c16 *tw = (c16 *)malloc(sizeof(c16) * nfft);
    c16 *scratch = (c16 *)malloc(sizeof(c16) * nfft);
    plan_perm(&pl, perm);
    oracle_fft_q15_twiddles(nfft, inverse, (int16_t *)tw);
    for (size_t b = 0; b < batch; b++) {
        const c16 *in = (const c16 *)vin + b * nfft;
        c16 *F = (c16 *)vout + b * nfft;
        for (int o = 0; o < nfft; o++) F[o] = in[perm[o]];
        for (int s = pl.nstages - 1; s >= 0; s--) {
            const int p = pl.p[s], m = pl.m[s], span = p * m, fstride = nfft / span;
            for (int g = 0; g < nfft; g += span) q15_stage(F + g, tw, nfft, p, m, fstride, inverse, scratch);
        }
    }
    free(perm); free(tw); free(scratch);
    return 0;
}

/* dtype codes shared with include/b200comms.h: 1 = cf32, 3 = cf64, 7 = ci16 */
int oracle_fft(int dtype, size_t nbins, int inverse, const void *in, void *out, size_t batch)
{
    switch (dtype) {
    case 1: return oracle_fft_cf32((int)nbins, inverse, in, out, batch);
    case 3: return oracle_fft_cf64((int)nbins, inverse, in, out, batch);
    case 7: return oracle_fft_ci16((int)nbins, inverse, in, out, batch);
    }
    return -1;
}
