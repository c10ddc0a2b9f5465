/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the /comms/fir_filter hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * may call this.  The product (libb200comms.so) never links or loads it.
 *
 * Restates, scalar and in the reference's order of operations:
 *   - filter/FIRFilter.cpp:327-354  updateInternals(): K, per-phase taps, floatToQ
 *   - filter/FIRFilter.cpp:278-302  work(): N, decimation counter, convolution nest, fromQ
 *   - filter/FIRFilter.cpp:263-272  burst flush (K-1 appended zeros)
 *   - filter/FIRFilter.cpp:371-382  the (data, QType, QTapsType) table
 * The reference block itself cannot be compiled here (needs PothosCore, absent), so the FIR
 * oracle is a "port"; its Q-format rounding is PARITY UNPINNED (see qformat.h).
 * Integer arithmetic wraps (two's complement), as the reference's does in practice.
 */
#include <complex.h>
#include <pthread.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "qformat.h"

enum { DT_F32 = 0, DT_CF32, DT_F64, DT_CF64, DT_I8, DT_CI8, DT_I16, DT_CI16, DT_I32, DT_CI32, DT_I64, DT_CI64 };

static int dt_is_complex(int dt) { return dt & 1; }
static size_t dt_scalar_bytes(int dt)
{
    switch (dt >> 1) { case 0: return 4; case 1: return 8; case 2: return 1; case 3: return 2; case 4: return 4; default: return 8; }
}
/* QType scalar bytes, FIRFilter.cpp:377-382: double->double, float->float, int64->int64,
 * int32->int64, int16->int32, int8->int16 */
static size_t dt_q_bytes(int dt)
{
    switch (dt >> 1) { case 0: return 4; case 1: return 8; case 2: return 2; case 3: return 4; case 4: return 8; default: return 8; }
}

/* ---- the convolution nest, FIRFilter.cpp:281-302, one instantiation per type row ------ */

/* floating point rows: QType == data type; fromQ is the identity cast */
#define FIR_FLOAT(NAME, T)                                                                       \
    static size_t NAME##_rr(const T *in, T *y, size_t N, size_t M, size_t L, size_t K,           \
                            T *const *taps, const size_t *nt)                                    \
    {                                                                                            \
        const T *x = in + (K - 1);                                                               \
        size_t decim = M, produced = 0;                                                          \
        for (size_t n = 0; n < N; n++)                                                           \
            for (size_t j = 0; j < L; j++) {                                                     \
                if (--decim != 0) continue;                                                      \
                decim = M;                                                                       \
                T y_n = 0;                                                                       \
                for (size_t k = 0; k < nt[j]; k++) y_n += taps[j][k] * x[(ptrdiff_t)n - (ptrdiff_t)k]; \
                y[produced++] = y_n;                                                             \
            }                                                                                    \
        return produced;                                                                         \
    }                                                                                            \
    static size_t NAME##_cr(const T _Complex *in, T _Complex *y, size_t N, size_t M, size_t L,   \
                            size_t K, T *const *taps, const size_t *nt)                          \
    {                                                                                            \
        const T _Complex *x = in + (K - 1);                                                      \
        size_t decim = M, produced = 0;                                                          \
        for (size_t n = 0; n < N; n++)                                                           \
            for (size_t j = 0; j < L; j++) {                                                     \
                if (--decim != 0) continue;                                                      \
                decim = M;                                                                       \
                T _Complex y_n = 0;                                                              \
                for (size_t k = 0; k < nt[j]; k++) {                                             \
                    /* std::complex<T> * T scales both components (no cross terms) */            \
                    const T _Complex xv = x[(ptrdiff_t)n - (ptrdiff_t)k];                        \
                    const T t = taps[j][k];                                                      \
                    y_n += CMPLX##NAME(t * creal##NAME(xv), t * cimag##NAME(xv));                \
                }                                                                                \
                y[produced++] = y_n;                                                             \
            }                                                                                    \
        return produced;                                                                         \
    }                                                                                            \
    static size_t NAME##_cc(const T _Complex *in, T _Complex *y, size_t N, size_t M, size_t L,   \
                            size_t K, T _Complex *const *taps, const size_t *nt)                 \
    {                                                                                            \
        const T _Complex *x = in + (K - 1);                                                      \
        size_t decim = M, produced = 0;                                                          \
        for (size_t n = 0; n < N; n++)                                                           \
            for (size_t j = 0; j < L; j++) {                                                     \
                if (--decim != 0) continue;                                                      \
                decim = M;                                                                       \
                T _Complex y_n = 0;                                                              \
                /* std::complex<T>::operator* -> the compiler's complex multiply */             \
                for (size_t k = 0; k < nt[j]; k++) y_n += taps[j][k] * x[(ptrdiff_t)n - (ptrdiff_t)k]; \
                y[produced++] = y_n;                                                             \
            }                                                                                    \
        return produced;                                                                         \
    }

#define CMPLXf32(a, b) CMPLXF(a, b)
#define crealf32 crealf
#define cimagf32 cimagf
#define CMPLXf64(a, b) CMPLX(a, b)
#define crealf64 creal
#define cimagf64 cimag
FIR_FLOAT(f32, float)
FIR_FLOAT(f64, double)

/* integer rows: S = sample scalar, Q = accumulator scalar (signed), UQ = same width unsigned
 * (used so that wrap-around is defined behaviour in the oracle), SH = fromQ shift */
/* wrap-around multiply in the accumulator width, free of signed-overflow UB */
#define MULW(UQ, a, b) ((UQ)((uint64_t)(int64_t)(a) * (uint64_t)(int64_t)(b)))
#define FIR_INT(NAME, S, Q, UQ, SH)                                                              \
    static size_t NAME##_rr(const S *in, S *y, size_t N, size_t M, size_t L, size_t K,           \
                            Q *const *taps, const size_t *nt)                                    \
    {                                                                                            \
        const S *x = in + (K - 1);                                                               \
        size_t decim = M, produced = 0;                                                          \
        for (size_t n = 0; n < N; n++)                                                           \
            for (size_t j = 0; j < L; j++) {                                                     \
                if (--decim != 0) continue;                                                      \
                decim = M;                                                                       \
                UQ y_n = 0;                                                                      \
                for (size_t k = 0; k < nt[j]; k++)                                               \
                    y_n += MULW(UQ, taps[j][k], (Q)x[(ptrdiff_t)n - (ptrdiff_t)k]);               \
                y[produced++] = (S)((Q)y_n >> SH);                                               \
            }                                                                                    \
        return produced;                                                                         \
    }                                                                                            \
    static size_t NAME##_cr(const S *in, S *y, size_t N, size_t M, size_t L, size_t K,           \
                            Q *const *taps, const size_t *nt)                                    \
    {                                                                                            \
        const S *x = in + 2 * (K - 1);                                                           \
        size_t decim = M, produced = 0;                                                          \
        for (size_t n = 0; n < N; n++)                                                           \
            for (size_t j = 0; j < L; j++) {                                                     \
                if (--decim != 0) continue;                                                      \
                decim = M;                                                                       \
                UQ yr = 0, yi = 0;                                                               \
                for (size_t k = 0; k < nt[j]; k++) {                                             \
                    const S *xv = x + 2 * ((ptrdiff_t)n - (ptrdiff_t)k);                         \
                    yr += MULW(UQ, taps[j][k], (Q)xv[0]);                                        \
                    yi += MULW(UQ, taps[j][k], (Q)xv[1]);                                        \
                }                                                                                \
                y[2 * produced] = (S)((Q)yr >> SH);                                              \
                y[2 * produced + 1] = (S)((Q)yi >> SH);                                          \
                produced++;                                                                      \
            }                                                                                    \
        return produced;                                                                         \
    }                                                                                            \
    static size_t NAME##_cc(const S *in, S *y, size_t N, size_t M, size_t L, size_t K,           \
                            Q *const *taps, const size_t *nt)                                    \
    {                                                                                            \
        const S *x = in + 2 * (K - 1);                                                           \
        size_t decim = M, produced = 0;                                                          \
        for (size_t n = 0; n < N; n++)                                                           \
            for (size_t j = 0; j < L; j++) {                                                     \
                if (--decim != 0) continue;                                                      \
                decim = M;                                                                       \
                UQ yr = 0, yi = 0;                                                               \
                for (size_t k = 0; k < nt[j]; k++) {                                             \
                    /* libstdc++ generic complex<_Tp>::operator*=: (ac - bd, ad + bc) */         \
                    const S *xv = x + 2 * ((ptrdiff_t)n - (ptrdiff_t)k);                         \
                    const UQ tr = (UQ)taps[j][2 * k], ti = (UQ)taps[j][2 * k + 1];               \
                    const UQ xr = (UQ)(Q)xv[0], xi = (UQ)(Q)xv[1];                               \
                    yr += (UQ)(MULW(UQ, tr, xr) - MULW(UQ, ti, xi));                                                     \
                    yi += (UQ)(MULW(UQ, tr, xi) + MULW(UQ, ti, xr));                                                     \
                }                                                                                \
                y[2 * produced] = (S)((Q)yr >> SH);                                              \
                y[2 * produced + 1] = (S)((Q)yi >> SH);                                          \
                produced++;                                                                      \
            }                                                                                    \
        return produced;                                                                         \
    }

FIR_INT(i8, int8_t, int16_t, uint16_t, 8)
FIR_INT(i16, int16_t, int32_t, uint32_t, 16)
FIR_INT(i32, int32_t, int64_t, uint64_t, 32)
FIR_INT(i64, int64_t, int64_t, uint64_t, 32)

/* ---- updateInternals(), FIRFilter.cpp:327-354 ------------------------------------------ */
typedef struct {
    size_t K, L, M;
    void **taps;   /* L arrays of QTapsType */
    size_t *nt;    /* taps per phase */
} fir_plan;

static void plan_free(fir_plan *p)
{
    if (p->taps) for (size_t j = 0; j < p->L; j++) free(p->taps[j]);
    free(p->taps);
    free(p->nt);
}

static int plan_make(fir_plan *p, int dt, int taps_complex, const double *taps, size_t ntaps, size_t M, size_t L)
{
    if (ntaps == 0 || M == 0 || L == 0) return -1;
    if (taps_complex && !dt_is_complex(dt)) return -1; /* FIRFilter.cpp:373-376: no real data + COMPLEX taps */
    const int is_float = (dt >> 1) < 2;
    const size_t qb = dt_q_bytes(dt), tc = taps_complex ? 2 : 1;
    p->M = M; p->L = L;
    p->K = ntaps / L + ((ntaps % L) == 0 ? 0 : 1);
    p->taps = (void **)calloc(L, sizeof(void *));
    p->nt = (size_t *)calloc(L, sizeof(size_t));
    for (size_t j = 0; j < L; j++) {
        p->taps[j] = calloc(p->K * tc + 1, qb);
        size_t cnt = 0;
        for (size_t k = 0; k < p->K; k++) {
            const size_t i = j + k * L;
            if (i >= ntaps) continue;
            for (size_t c = 0; c < tc; c++) {
                const double v = taps[i * tc + c];
                char *dst = (char *)p->taps[j] + (cnt * tc + c) * qb;
                if (is_float) { if (qb == 4) *(float *)dst = (float)v; else *(double *)dst = v; }
                else {
                    const int64_t q = oracle_float_to_q(v, (int)qb);
                    if (qb == 2) *(int16_t *)dst = (int16_t)q; else if (qb == 4) *(int32_t *)dst = (int32_t)q; else *(int64_t *)dst = q;
                }
            }
            cnt++;
        }
        p->nt[j] = cnt;
    }
    return 0;
}

static size_t run_nest(const fir_plan *p, int dt, int tcx, const void *in, void *out, size_t N)
{
#define ARGS(TI, TO, TT) (const TI *)in, (TO *)out, N, p->M, p->L, p->K, (TT *const *)p->taps, p->nt
    switch (dt) {
    case DT_F32: return f32_rr(ARGS(float, float, float));
    case DT_CF32: return tcx ? f32_cc(ARGS(float _Complex, float _Complex, float _Complex)) : f32_cr(ARGS(float _Complex, float _Complex, float));
    case DT_F64: return f64_rr(ARGS(double, double, double));
    case DT_CF64: return tcx ? f64_cc(ARGS(double _Complex, double _Complex, double _Complex)) : f64_cr(ARGS(double _Complex, double _Complex, double));
    case DT_I8: return i8_rr(ARGS(int8_t, int8_t, int16_t));
    case DT_CI8: return tcx ? i8_cc(ARGS(int8_t, int8_t, int16_t)) : i8_cr(ARGS(int8_t, int8_t, int16_t));
    case DT_I16: return i16_rr(ARGS(int16_t, int16_t, int32_t));
    case DT_CI16: return tcx ? i16_cc(ARGS(int16_t, int16_t, int32_t)) : i16_cr(ARGS(int16_t, int16_t, int32_t));
    case DT_I32: return i32_rr(ARGS(int32_t, int32_t, int64_t));
    case DT_CI32: return tcx ? i32_cc(ARGS(int32_t, int32_t, int64_t)) : i32_cr(ARGS(int32_t, int32_t, int64_t));
    case DT_I64: return i64_rr(ARGS(int64_t, int64_t, int64_t));
    case DT_CI64: return tcx ? i64_cc(ARGS(int64_t, int64_t, int64_t)) : i64_cr(ARGS(int64_t, int64_t, int64_t));
    }
#undef ARGS
    return 0;
}

/* K for (ntaps, L), FIRFilter.cpp:335 */
size_t oracle_fir_K(size_t ntaps, size_t L) { return ntaps / L + ((ntaps % L) == 0 ? 0 : 1); }

/* The per-phase Q taps as updateInternals() would build them, written as int64/double pairs
 * for inspection by tests: out[j*K*tc + k*tc + c], zero padded; nt_out[j] = taps in phase j. */
int oracle_fir_phase_taps(int dt, int taps_complex, const double *taps, size_t ntaps, size_t L,
                          double *out, size_t *nt_out)
{
    fir_plan p;
    if (plan_make(&p, dt, taps_complex, taps, ntaps, 1, L)) return -1;
    const size_t qb = dt_q_bytes(dt), tc = taps_complex ? 2 : 1;
    const int is_float = (dt >> 1) < 2;
    for (size_t j = 0; j < L; j++) {
        nt_out[j] = p.nt[j];
        for (size_t e = 0; e < p.K * tc; e++) {
            const char *src = (const char *)p.taps[j] + e * qb;
            double v;
            if (is_float) v = qb == 4 ? *(const float *)src : *(const double *)src;
            else v = qb == 2 ? (double)*(const int16_t *)src : qb == 4 ? (double)*(const int32_t *)src : (double)*(const int64_t *)src;
            out[j * p.K * tc + e] = e < p.nt[j] * tc ? v : 0.0;
        }
    }
    plan_free(&p);
    return 0;
}

/*
 * One work() call's arithmetic.  `in` holds in_elems elements: K-1 history then new data
 * (FIRFilter.cpp:281).  zero_tail != 0 appends K-1 zero elements first (burst flush,
 * FIRFilter.cpp:265-272).  N = min((elems-(K-1))/M, out_cap/L)*M (:278).
 * Returns 0 and sets *consumed = N, *produced = (N/M)*L (:307-308).
 */
int oracle_fir_run(int dt, int taps_complex, const double *taps, size_t ntaps, size_t M, size_t L,
                   const void *in, size_t in_elems, void *out, size_t out_cap, int zero_tail,
                   size_t *consumed, size_t *produced)
{
    fir_plan p;
    *consumed = *produced = 0;
    if (plan_make(&p, dt, taps_complex, taps, ntaps, M, L)) return -1;
    const size_t esz = dt_scalar_bytes(dt) * (dt_is_complex(dt) ? 2 : 1);
    void *flush = NULL;
    size_t elems = in_elems;
    if (zero_tail) {
        elems = in_elems + p.K - 1;
        flush = calloc(elems ? elems : 1, esz);
        memcpy(flush, in, in_elems * esz);
        in = flush;
    }
    if (elems >= p.K - 1 + M) {
        size_t a = (elems - (p.K - 1)) / M, b = out_cap / L;
        const size_t N = (a < b ? a : b) * M;
        const size_t got = run_nest(&p, dt, taps_complex, in, out, N);
        *consumed = N;
        *produced = got; /* == (N/M)*L */
    }
    free(flush);
    plan_free(&p);
    return 0;
}

/* ---- all-host-cores driver for the CPU baseline (one contiguous segment per thread, each
 * segment start a multiple of M so the decimation counter phase is preserved) -- */
typedef struct {
    const fir_plan *p; int dt, tcx; const char *in; char *out; size_t n0, N, esz;
} seg_arg;

static void *seg_main(void *v)
{
    seg_arg *a = (seg_arg *)v;
    run_nest(a->p, a->dt, a->tcx, a->in + a->n0 * a->esz, a->out + (a->n0 / a->p->M) * a->p->L * a->esz, a->N);
    return NULL;
}

int oracle_fir_run_mt(int nthreads, int dt, int taps_complex, const double *taps, size_t ntaps, size_t M,
                      size_t L, const void *in, size_t in_elems, void *out, size_t out_cap,
                      size_t *consumed, size_t *produced)
{
    fir_plan p;
    *consumed = *produced = 0;
    if (nthreads < 1) nthreads = 1;
    if (plan_make(&p, dt, taps_complex, taps, ntaps, M, L)) return -1;
    if (in_elems >= p.K - 1 + M) {
        size_t a = (in_elems - (p.K - 1)) / M, b = out_cap / L;
        const size_t blocks = a < b ? a : b; /* N/M */
        const size_t esz = dt_scalar_bytes(dt) * (dt_is_complex(dt) ? 2 : 1);
        pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
        seg_arg *args = (seg_arg *)calloc(nthreads, sizeof(seg_arg));
        for (int t = 0; t < nthreads; t++) {
            const size_t b0 = blocks * t / nthreads, b1 = blocks * (t + 1) / nthreads;
            args[t] = (seg_arg){&p, dt, taps_complex, (const char *)in, (char *)out, b0 * M, (b1 - b0) * M, esz};
            pthread_create(&th[t], NULL, seg_main, &args[t]);
        }
        for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
        free(th); free(args);
        *consumed = blocks * M;
        *produced = blocks * L;
    }
    plan_free(&p);
    return 0;
}
