/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the /comms/fir_filter hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference
 * legs may use anything under oracle/.  The product path never links or calls this.
 *
 * Q-format helpers restating PothosCore include/Pothos/Util/QFormat.hpp (EXTERNAL to
 * /root/reference; PothosCore >= 0.6.0 per reference CMakeLists.txt:8).  Call sites in the
 * reference: filter/FIRFilter.cpp:300 (fromQ) and :348 (floatToQ).
 *
 * PARITY UNPINNED for the rounding direction: the only reference tests touching these
 * helpers (math/TestScale.cpp:47-53, math/TestRotate.cpp:48-54) pin that the number of
 * fractional bits is half the Q scalar word (4*sizeof(Q scalar)) and that floatToQ and
 * fromQ agree on it; every case they use is exactly representable, so truncate-vs-round
 * is not pinned by any in-tree test.  The published header does:
 *     floatToQ<T>(x, n = 4*sizeof(T scalar)) : integer T -> T(std::ldexp(x, n))  (C++ trunc toward 0)
 *                                              float   T -> T(x)
 *     fromQ<T>(q,   n = 4*sizeof(Q scalar))  : integer Q -> T(q >> n)            (arithmetic shift, wraps on narrowing)
 *                                              float   Q -> T(q)
 * applied per component for std::complex.  This one header is the single place the
 * assumption lives.
 */
#ifndef B200C_ORACLE_QFORMAT_H
#define B200C_ORACLE_QFORMAT_H
#include <math.h>
#include <stdint.h>

/* floatToQ for an integer Q scalar of `qbytes` bytes: trunc(ldexp(x, 4*qbytes)). The
 * double->integer conversion of an out-of-range value is UB in C++; the oracle (and the
 * product) saturate-free wrap through int64 for |x*2^n| < 2^63, which covers every sane tap. */
static inline int64_t oracle_float_to_q(double x, int qbytes)
{
    const double s = ldexp(x, 4 * qbytes);
    return (int64_t)s; /* C conversion truncates toward zero, as T(double) does in C++ */
}

#endif
