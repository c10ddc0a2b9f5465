// TEST INFRASTRUCTURE ONLY -- stand-in for PothosCore's include/Pothos/Util/QFormat.hpp, which
// is EXTERNAL to /root/reference (PothosCore >= 0.6.0, reference CMakeLists.txt:8; no commit is
// pinned) and absent from this image.  It exists so that the reference's own
// filter/FIRFilter.cpp compiles UNMODIFIED into oracle/_ref/libfirref.so (oracle/Makefile).
//
// RECALLED, not copied: this is the single header of the reference FIR build that does not come
// from /root/reference, hence the single place where FIR parity stays "unpinned":
//   fromQ<T>(q, n = 4*sizeof(Q scalar))    integral Q -> T(q >> n)          floating Q -> T(q)
//   floatToQ<T>(x, n = 4*sizeof(T scalar)) integral T -> T(std::ldexp(x,n)) floating T -> T(x)
// applied per component for std::complex.  The in-tree tests that touch these helpers
// (math/TestScale.cpp:47-53, math/TestRotate.cpp:48-54) pin the bit counts, not the rounding.
// Call sites: filter/FIRFilter.cpp:300 (fromQ<OutType>(y_n)), :348 (floatToQ<QTapsType>(tap)).
#pragma once
#include <cmath>
#include <complex>
#include <type_traits>

namespace Pothos {
namespace Util {

namespace Detail {
template <typename T, typename U> T fromQImpl(const U &in, const int, std::false_type) { return T(in); }
template <typename T, typename U> T fromQImpl(const U &in, const int n, std::true_type) { return T(in >> n); }
template <typename T, typename U> T floatToQImpl(const U &in, const int, std::false_type) { return T(in); }
template <typename T, typename U> T floatToQImpl(const U &in, const int n, std::true_type) { return T(std::ldexp(in, n)); }
} // namespace Detail

template <typename T, typename U> T fromQ(const U &in, const int n = sizeof(U) * 4)
{
    return Detail::fromQImpl<T>(in, n, std::is_integral<U>());
}

template <typename T, typename U> T fromQ(const std::complex<U> &in, const int n = sizeof(U) * 4)
{
    typedef typename T::value_type S;
    return T(fromQ<S, U>(in.real(), n), fromQ<S, U>(in.imag(), n));
}

template <typename T, typename U> T floatToQ(const U &in, const int n = sizeof(T) * 4)
{
    return Detail::floatToQImpl<T>(in, n, std::is_integral<T>());
}

template <typename T, typename U> T floatToQ(const std::complex<U> &in, const int n = sizeof(typename T::value_type) * 4)
{
    typedef typename T::value_type S;
    return T(floatToQ<S, U>(in.real(), n), floatToQ<S, U>(in.imag(), n));
}

} // namespace Util
} // namespace Pothos
