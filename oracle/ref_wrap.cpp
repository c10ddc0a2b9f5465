/*
 * TEST INFRASTRUCTURE ONLY.  Thin extern "C" driver around the REFERENCE's own FFT sources,
 * compiled where they lie (-I/root/reference/fft; nothing is copied into this repo):
 *   fft/FFTAux.h:16-48 (the exact helper the /comms/fft block holds, fft/FFT.cpp:77),
 *   fft/kissfft.hh (cf32/cf64) and fft/kiss_fft.c built with -DFIXED_POINT=16 (complex int16),
 *   as fft/CMakeLists.txt:19-28 does.
 * Output goes only to oracle/_ref/ (git-ignored).  One transform per call, exactly like
 * FFT<Type>::work() (fft/FFT.cpp:61-72).
 */
#include <complex>
#include <cstddef>
#include <cstdint>
#include <thread>
#include <vector>
#include "FFTAux.h"

template <typename T>
static void run(size_t nbins, bool inverse, const void *in, void *out, size_t batch)
{
    FFTAux<std::complex<T>> aux(nbins, inverse);
    auto src = static_cast<const std::complex<T> *>(in);
    auto dst = static_cast<std::complex<T> *>(out);
    for (size_t b = 0; b < batch; b++) aux.transform(src + b * nbins, dst + b * nbins);
}

static int dispatch(int dtype, size_t nbins, int inverse, const void *in, void *out, size_t batch)
{
    switch (dtype) {
    case 1: run<float>(nbins, inverse != 0, in, out, batch); return 0;           /* complex_float32 */
    case 3: run<double>(nbins, inverse != 0, in, out, batch); return 0;          /* complex_float64 */
    case 7: run<kiss_fft_scalar>(nbins, inverse != 0, in, out, batch); return 0; /* complex_int16 */
    }
    return -1; /* FFTFactory throws for anything else, fft/FFT.cpp:92 */
}

extern "C" int ref_fft(int dtype, size_t nbins, int inverse, const void *in, void *out, size_t batch)
{
    return dispatch(dtype, nbins, inverse, in, out, batch);
}

/* all-host-cores driver for the CPU baseline: independent block instances, one per thread */
extern "C" int ref_fft_mt(int nthreads, int dtype, size_t nbins, int inverse, const void *in, void *out, size_t batch)
{
    if (nthreads < 1) nthreads = 1;
    const size_t esz = dtype == 1 ? 8 : dtype == 3 ? 16 : dtype == 7 ? 4 : 0;
    if (!esz) return -1;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; t++) {
        const size_t b0 = batch * t / nthreads, b1 = batch * (t + 1) / nthreads;
        th.emplace_back([=] {
            dispatch(dtype, nbins, inverse, static_cast<const char *>(in) + b0 * nbins * esz,
                     static_cast<char *>(out) + b0 * nbins * esz, b1 - b0);
        });
    }
    for (auto &t : th) t.join();
    return 0;
}
