/*
 * TEST INFRASTRUCTURE ONLY -- CPU oracle for the stream sources next to the filter path
 * (SURVEY.md section 8f rank 4): /comms/waveform_source and /comms/noise_source.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline legs may use anything under
 * oracle/.  Restatement of
 *   waveform/WaveformSource.cpp:184-260 (updateTable: table size search, step, CONST / SINE / RAMP /
 *                                        SQUARE fill), :262-272 (setElem), :98-108 (work)
 *   waveform/NoiseSource.cpp:188-226 (updateTable), :228-251 (setElem, _laplace), :103-125 (work, fast
 *                                        mode: the only mode a caller can reach, _fast has no setter)
 * C++ because the noise tables are draws of libstdc++'s std::mt19937 through its
 * std::uniform_real / normal / poisson / uniform_int distributions: the reference's stream IS that
 * library's output, so the oracle calls the same library with the draw order written out.
 * The reference seeds from std::random_device (NoiseSource.cpp:84); here the seed is an argument.
 * PARITY NOTE: the reference writes std::complex<double>(dist(gen), dist(gen)); C++ leaves the order
 * of the two draws unspecified and g++ (x86-64) evaluates the SECOND argument first, so the imaginary
 * part is drawn before the real part.  That order is stated explicitly below; no reference test pins it
 * (parity unpinned), the device block layer keeps the reference's expression and is compiled by the
 * same g++.
 *
 * dtype codes: (class << 1) | complex, class 0 f32, 1 f64, 2 i8, 3 i16, 4 i32, 5 i64.
 */
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <random>
#include <string>
#include <vector>

namespace {

const size_t kDefaultWaveTable = 4096, kMaxWaveTable = 1024 * 1024, kMinTableStep = 16;   // WaveformSource.cpp:10-12
const size_t kNoiseTable = 4096;                                                          // NoiseSource.cpp:11

size_t scalar_bytes(int cls) { return cls == 0 ? 4 : cls == 1 ? 8 : cls == 2 ? 1 : cls == 3 ? 2 : cls == 4 ? 4 : 8; }
size_t elem_bytes(int dtype) { return scalar_bytes(dtype >> 1) * ((dtype & 1) ? 2 : 1); }

// Type(double): float types round, integer types truncate toward zero (C++ conversion)
void store_scalar(void *t, size_t k, int cls, double v)
{
    switch (cls) {
    case 0: static_cast<float *>(t)[k] = (float)v; break;
    case 1: static_cast<double *>(t)[k] = v; break;
    case 2: static_cast<int8_t *>(t)[k] = (int8_t)v; break;
    case 3: static_cast<int16_t *>(t)[k] = (int16_t)v; break;
    case 4: static_cast<int32_t *>(t)[k] = (int32_t)v; break;
    default: static_cast<int64_t *>(t)[k] = (int64_t)v; break;
    }
}

// setElem(): out = Type(scalar * val + offset) (complex types) or its real part (real types)
void set_elem(void *table, size_t i, int dtype, std::complex<double> scalar, std::complex<double> offset, std::complex<double> val)
{
    const std::complex<double> v = scalar * val + offset;
    if (dtype & 1) {
        store_scalar(table, 2 * i, dtype >> 1, v.real());
        store_scalar(table, 2 * i + 1, dtype >> 1, v.imag());
    } else {
        store_scalar(table, i, dtype >> 1, v.real());
    }
}

} // namespace

extern "C" {

/* WaveformSource::updateTable().  Returns 0, -1 (InvalidArgumentException: unknown wave, or a
 * non-zero frequency whose step rounds to 0), -2 (cap_entries too small; *entries is still set). */
int oracle_waveform_table(int dtype, const char *wave, double freq, double rate, double res, double ampl_re, double ampl_im,
                          double off_re, double off_im, void *table, size_t cap_entries, size_t *entries, uint64_t *step)
{
    const std::complex<double> scalar(ampl_re, ampl_im), offset(off_re, off_im);
    const double frac = ((res == 0.0) ? freq : res) / rate;                    // :190
    size_t n = kDefaultWaveTable;
    while (true) {                                                             // :193-202
        const long long delta = std::llround(frac * n);
        if (frac == 0.0) break;
        if ((size_t)std::llabs(delta) >= kMinTableStep) break;
        if (n * 2 > kMaxWaveTable) break;
        n *= 2;
    }
    *entries = n;
    *step = (uint64_t)(size_t)std::llround((freq / rate) * n);                 // :208 (negative frequencies wrap)
    if (*step == 0 && freq != 0.0) return -1;                                  // :209-212
    const std::string w(wave);
    if (w != "CONST" && w != "SINE" && w != "RAMP" && w != "SQUARE") return -1;   // :259
    if (n > cap_entries) return -2;
    for (size_t i = 0; i < n; i++) {
        std::complex<double> val;
        const size_t q = (i + (3 * n) / 4) % n;
        if (w == "CONST") val = 1.0;
        else if (w == "SINE") val = std::polar(1.0, 2 * M_PI * i / n);
        else if (w == "RAMP") val = std::complex<double>(2.0 * i / (n - 1) - 1.0, 2.0 * q / (n - 1) - 1.0);
        else val = std::complex<double>((i < n / 2) ? 0.0 : 1.0, (q < n / 2) ? 0.0 : 1.0);
        set_elem(table, i, dtype, scalar, offset, val);
    }
    return 0;
}

/* work(): out[i] = table[index & mask]; index += step  (entries a power of two) */
void oracle_table_walk(int dtype, const void *table, size_t entries, uint64_t index, uint64_t step, void *out, size_t n)
{
    const size_t esz = elem_bytes(dtype);
    const uint64_t mask = entries - 1;
    for (size_t i = 0; i < n; i++, index += step)
        std::memcpy(static_cast<char *>(out) + i * esz, static_cast<const char *>(table) + (size_t)(index & mask) * esz, esz);
}

/* A NoiseSource seeded with `seed`: activate() (table fill), then per work call w: optionally
 * (refill_before[w] != 0) a setter call -- which refills the table from the running generator --
 * then work() producing work_elems[w] elements.  `out` receives the concatenated stream,
 * `table_out` (may be NULL) the table as it stands at the end.  Returns -1 for an unknown wave. */
int oracle_noise_stream(int dtype, const char *wave, double mean, double b, double ampl_re, double ampl_im, double off_re,
                        double off_im, uint32_t seed, const size_t *work_elems, const int *refill_before, size_t nwork, void *out,
                        void *table_out)
{
    const std::complex<double> scalar(ampl_re, ampl_im), offset(off_re, off_im);
    const std::string w(wave);
    std::mt19937 gen(seed);
    std::uniform_int_distribution<size_t> waveIndex(0, kNoiseTable - 1);
    std::uniform_real_distribution<> uniform;
    std::normal_distribution<> normal;
    std::poisson_distribution<> poisson;
    const size_t esz = elem_bytes(dtype);
    std::vector<char> table(kNoiseTable * esz);

    auto laplace = [&]() {                                                     // :238-245
        const double num = uniform(gen);
        if (num < 0) return mean + b * std::log(1 + num);
        return mean - b * std::log(1 - num);
    };
    auto fill = [&]() -> int {
        if (w == "UNIFORM") uniform = std::uniform_real_distribution<>(mean - b, mean + b);
        else if (w == "NORMAL") normal = std::normal_distribution<>(mean, b);
        else if (w == "LAPLACE") uniform = std::uniform_real_distribution<>(mean - b, mean + b);
        else if (w == "POISSON") poisson = std::poisson_distribution<>(mean);
        else return -1;
        for (size_t i = 0; i < kNoiseTable; i++) {
            double re, im;                                                     // imaginary part first: see the header
            if (w == "UNIFORM") { im = uniform(gen); re = uniform(gen); }
            else if (w == "NORMAL") { im = normal(gen); re = normal(gen); }
            else if (w == "LAPLACE") { im = laplace(); re = laplace(); }
            else { im = poisson(gen); re = poisson(gen); }
            set_elem(table.data(), i, dtype, scalar, offset, std::complex<double>(re, im));
        }
        return 0;
    };

    if (fill()) return -1;
    size_t index = 0, at = 0;
    for (size_t k = 0; k < nwork; k++) {
        if (refill_before && refill_before[k] && fill()) return -1;
        index += waveIndex(gen);                                               // :108
        for (size_t i = 0; i < work_elems[k]; i++, index++, at++)
            std::memcpy(static_cast<char *>(out) + at * esz, table.data() + (index % kNoiseTable) * esz, esz);
    }
    if (table_out) std::memcpy(table_out, table.data(), table.size());
    return 0;
}

} // extern "C"
