/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the HBM-resident neighbours of the filter path
 * (SURVEY.md section 8f rank 4): /comms/scale, /comms/rotate, /comms/signal_probe.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may use anything under
 * oracle/.  Scalar C restatement of
 *   math/Scale.cpp:15-23 (arrayScale), :41-45 (setFactor), :147-153 (type table)
 *   math/Rotate.cpp:15-23 (arrayRotate), :71-75 (setPhase), :150-155 (type table)
 *   utility/SignalProbe.cpp:128-160 (VALUE / RMS / MEAN)
 * with the Q-format helpers of qformat.h (PothosCore, external; rounding direction parity
 * unpinned, see that header).  The reference tests these blocks hold (math/TestScale.cpp:47-53,
 * math/TestRotate.cpp:48-54: result within 1 of Type(input * factor)) are re-run against this
 * file in tests/test_oracle_math.py.
 *
 * dtype codes: (class << 1) | complex, class 0 f32, 1 f64, 2 i8, 3 i16, 4 i32, 5 i64; the Q type
 * of a class is f32, f64, i16, i32, i64, i64 (Scale.cpp:150-153 / Rotate.cpp:150-155).
 */
#include <complex.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>

#include "qformat.h"

static int q_bytes(int cls) { return cls == 2 ? 2 : cls == 3 ? 4 : 8; }

/* wrap a 64-bit value to the Q type of an integer class (two's complement narrowing) */
static int64_t wrap_q(int64_t v, int cls)
{
    if (cls == 2) return (int16_t)v;
    if (cls == 3) return (int32_t)v;
    return v;
}

/* fromQ<Type>(q): q >> 4*sizeof(Q scalar), narrowed to the data type */
static void store_from_q(void *out, size_t i, int64_t q, int cls)
{
    const int64_t s = q >> (4 * q_bytes(cls));
    if (cls == 2) ((int8_t *)out)[i] = (int8_t)s;
    else if (cls == 3) ((int16_t *)out)[i] = (int16_t)s;
    else if (cls == 4) ((int32_t *)out)[i] = (int32_t)s;
    else ((int64_t *)out)[i] = s;
}

static int64_t load_int(const void *in, size_t i, int cls)
{
    if (cls == 2) return ((const int8_t *)in)[i];
    if (cls == 3) return ((const int16_t *)in)[i];
    if (cls == 4) return ((const int32_t *)in)[i];
    return ((const int64_t *)in)[i];
}

/* out[i] = fromQ(floatToQ(factor) * Q(in[i])) over `n_scalars` scalars (complex data: the real
 * factor multiplies both parts, so re and im are just consecutive scalars). */
int oracle_scale(int dtype, double factor, const void *in, void *out, size_t n_scalars)
{
    const int cls = dtype >> 1;
    if (cls < 0 || cls > 5) return -1;
    if (cls == 0) {
        const float f = (float)factor;
        for (size_t i = 0; i < n_scalars; i++) ((float *)out)[i] = f * ((const float *)in)[i];
    } else if (cls == 1) {
        for (size_t i = 0; i < n_scalars; i++) ((double *)out)[i] = factor * ((const double *)in)[i];
    } else {
        const int64_t fq = wrap_q(oracle_float_to_q(factor, q_bytes(cls)), cls);
        for (size_t i = 0; i < n_scalars; i++) {
            /* product in the Q type: wrapping (unsigned arithmetic avoids signed-overflow UB) */
            const int64_t tmp = wrap_q((int64_t)((uint64_t)fq * (uint64_t)load_int(in, i, cls)), cls);
            store_from_q(out, i, tmp, cls);
        }
    }
    return 0;
}

/* out[i] = fromQ(floatToQ(polar(1, phase)) * Q(in[i])), complex types only */
int oracle_rotate(int dtype, double phase, const void *in, void *out, size_t n_elems)
{
    const int cls = dtype >> 1;
    if (!(dtype & 1) || cls < 0 || cls > 5) return -1;
    const double c = cos(phase), s = sin(phase);            /* std::polar(1.0, phase) */
    if (cls == 0) {
        const float pr = (float)c, pi = (float)s;
        const float *x = (const float *)in;
        float *y = (float *)out;
        for (size_t i = 0; i < n_elems; i++) {
            const float a = x[2 * i], b = x[2 * i + 1];
            y[2 * i] = pr * a - pi * b;
            y[2 * i + 1] = pr * b + pi * a;
        }
    } else if (cls == 1) {
        const double *x = (const double *)in;
        double *y = (double *)out;
        for (size_t i = 0; i < n_elems; i++) {
            const double a = x[2 * i], b = x[2 * i + 1];
            y[2 * i] = c * a - s * b;
            y[2 * i + 1] = c * b + s * a;
        }
    } else {
        const int qb = q_bytes(cls);
        const uint64_t pr = (uint64_t)wrap_q(oracle_float_to_q(c, qb), cls), pi = (uint64_t)wrap_q(oracle_float_to_q(s, qb), cls);
        for (size_t i = 0; i < n_elems; i++) {
            const uint64_t a = (uint64_t)load_int(in, 2 * i, cls), b = (uint64_t)load_int(in, 2 * i + 1, cls);
            /* std::complex<Q> multiply: (pr a - pi b, pr b + pi a), each wrapped to the Q type */
            store_from_q(out, 2 * i, wrap_q((int64_t)(pr * a - pi * b), cls), cls);
            store_from_q(out, 2 * i + 1, wrap_q((int64_t)(pr * b + pi * a), cls), cls);
        }
    }
    return 0;
}

static double load_double(const void *in, size_t i, int cls)
{
    if (cls == 0) return ((const float *)in)[i];
    if (cls == 1) return ((const double *)in)[i];
    return (double)load_int(in, i, cls);
}

/* mode 0 VALUE (last element), 1 RMS, 2 MEAN over `n_elems` elements; value = {re, im} */
int oracle_probe(int dtype, int mode, const void *in, size_t n_elems, double *value)
{
    const int cls = dtype >> 1, cx = dtype & 1, nc = cx ? 2 : 1;
    if (cls < 0 || cls > 5 || n_elems == 0) return -1;
    value[0] = value[1] = 0.0;
    if (mode == 0) {
        value[0] = load_double(in, (n_elems - 1) * nc, cls);
        if (cx) value[1] = load_double(in, (n_elems - 1) * nc + 1, cls);
    } else if (mode == 1) {
        double acc = 0.0;
        for (size_t n = 0; n < n_elems; n++) {
            const double v = cx ? cabs(CMPLX(load_double(in, 2 * n, cls), load_double(in, 2 * n + 1, cls))) : fabs(load_double(in, n, cls));
            acc += v * v;
        }
        value[0] = sqrt(acc / (double)n_elems);
    } else if (mode == 2) {
        double mr = 0.0, mi = 0.0;
        for (size_t n = 0; n < n_elems; n++) {
            mr += load_double(in, n * nc, cls);
            if (cx) mi += load_double(in, n * nc + 1, cls);
        }
        value[0] = mr / (double)n_elems;
        value[1] = mi / (double)n_elems;
    } else {
        return -1;
    }
    return 0;
}
