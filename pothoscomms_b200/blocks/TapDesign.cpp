// Tap and window design (host, double precision).  See TapDesign.hpp for scope and provenance.
#include "TapDesign.hpp"

#include <algorithm>
#include <cmath>
#include <stdexcept>

namespace b200c_design {

static const double kPi = 3.14159265358979323846264338327950288;

static double sinc(double x) { return x == 0.0 ? 1.0 : std::sin(kPi * x) / (kPi * x); }

// modified Bessel function of the first kind, order 0 (power series; converges fast for |x| < ~50)
static double bessel_i0(double x)
{
    double sum = 1.0, term = 1.0;
    const double q = x * x / 4.0;
    for (int k = 1; k < 500; k++) {
        term *= q / ((double)k * (double)k);
        sum += term;
        if (term < 1e-17 * sum) break;
    }
    return sum;
}

// ------------------------------------------------------------------------ windows ---
static std::vector<double> cosine_sum_window(size_t n, const std::vector<double> &a, bool open_ends)
{
    // w[i] = sum_k (-1)^k a_k cos(2 pi k t); open_ends: t = (i+1)/(n+1) so that no tap is zeroed
    std::vector<double> w(n);
    for (size_t i = 0; i < n; i++) {
        const double t = open_ends ? (double)(i + 1) / (double)(n + 1) : (n > 1 ? (double)i / (double)(n - 1) : 0.5);
        double v = 0.0, sgn = 1.0;
        for (size_t k = 0; k < a.size(); k++) { v += sgn * a[k] * std::cos(2.0 * kPi * (double)k * t); sgn = -sgn; }
        w[i] = v;
    }
    return w;
}

static std::vector<double> chebyshev_window(size_t n, double atten_db)
{
    // Dolph-Chebyshev: equal side lobes `atten_db` below the main lobe; frequency sampling of
    // T_{n-1}(x0 cos(pi k / n)) followed by an inverse DFT
    if (n == 1) return {1.0};
    if (atten_db <= 0.0) atten_db = 50.0;
    const double order = (double)(n - 1);
    const double r = std::pow(10.0, atten_db / 20.0);
    const double x0 = std::cosh(std::acosh(r) / order);
    auto cheb = [&](double x) {
        if (x > 1.0) return std::cosh(order * std::acosh(x));
        if (x < -1.0) return ((n - 1) % 2 ? -1.0 : 1.0) * std::cosh(order * std::acosh(-x));
        return std::cos(order * std::acos(x));
    };
    std::vector<double> p(n), w(n);
    for (size_t k = 0; k < n; k++) p[k] = cheb(x0 * std::cos(kPi * (double)k / (double)n));
    // inverse transform of the sampled polynomial; even lengths sit on a half-sample grid
    const bool odd = (n % 2) == 1;
    const size_t half = odd ? (n + 1) / 2 : n / 2 + 1;
    std::vector<double> W(half);
    for (size_t j = 0; j < half; j++) {
        double acc = 0.0;
        for (size_t k = 0; k < n; k++)
            acc += p[k] * std::cos((odd ? 0.0 : kPi * (double)k / (double)n) - 2.0 * kPi * (double)j * (double)k / (double)n);
        W[j] = acc;
    }
    if (odd) {
        for (size_t j = 0; j < half; j++) w[half - 1 - j] = w[half - 1 + j] = W[j];
    } else {
        for (size_t j = 1; j < half; j++) w[half - 1 - j] = w[half - 2 + j] = W[j];
    }
    double peak = 0.0;
    for (double v : w) peak = std::max(peak, std::fabs(v));
    for (auto &v : w) v /= peak;
    return w;
}

std::vector<double> design_window(const std::string &type, size_t n, double arg)
{
    if (n == 0) return {};
    if (type == "rectangular" || type == "rect" || type == "boxcar") return std::vector<double>(n, 1.0);
    if (type == "hann" || type == "hanning") return cosine_sum_window(n, {0.5, 0.5}, true);
    if (type == "hamming") return cosine_sum_window(n, {0.54, 0.46}, false);
    if (type == "blackman") return cosine_sum_window(n, {0.42, 0.5, 0.08}, true);
    if (type == "flattop") {
        auto w = cosine_sum_window(n, {0.21557895, 0.41663158, 0.277263158, 0.083578947, 0.006947368}, false);
        return w;
    }
    if (type == "bartlett") {
        std::vector<double> w(n);
        for (size_t i = 0; i < n; i++) w[i] = 1.0 - std::fabs(2.0 * (double)(i + 1) / (double)(n + 1) - 1.0);
        return w;
    }
    if (type == "kaiser") {
        std::vector<double> w(n);
        const double den = bessel_i0(arg);
        for (size_t i = 0; i < n; i++) {
            const double t = n > 1 ? 2.0 * (double)i / (double)(n - 1) - 1.0 : 0.0;
            w[i] = bessel_i0(arg * std::sqrt(std::max(0.0, 1.0 - t * t))) / den;
        }
        return w;
    }
    if (type == "chebyshev") return chebyshev_window(n, arg);
    throw std::runtime_error("design_window: unknown window type '" + type + "'");
}

// ------------------------------------------------------------- Parks-McClellan ---
namespace {

struct Bary {   // barycentric interpolation through (x_i, y_i)
    std::vector<double> x, y, b;
    double eval(double xv) const
    {
        double num = 0.0, den = 0.0;
        for (size_t i = 0; i < x.size(); i++) {
            const double d = xv - x[i];
            if (std::fabs(d) < 1e-15) return y[i];
            const double c = b[i] / d;
            num += c * y[i];
            den += c;
        }
        return num / den;
    }
};

// b_i = 1 / prod_{j != i} (x_i - x_j), returned with a common scale removed (only ratios are used)
std::vector<double> bary_weights(const std::vector<double> &x)
{
    const size_t m = x.size();
    std::vector<double> lg(m, 0.0), sg(m, 1.0), b(m);
    for (size_t i = 0; i < m; i++)
        for (size_t j = 0; j < m; j++) {
            if (i == j) continue;
            const double d = x[i] - x[j];
            lg[i] -= std::log(std::fabs(d));
            if (d < 0) sg[i] = -sg[i];
        }
    const double mx = *std::max_element(lg.begin(), lg.end());
    for (size_t i = 0; i < m; i++) b[i] = sg[i] * std::exp(lg[i] - mx);
    return b;
}

// cosine-series coefficients a_k (k < r) of the degree r-1 polynomial in cos(w) given by `p`
std::vector<double> cosine_coeffs(const Bary &p, size_t r)
{
    const size_t P = 2 * r - 1;
    std::vector<double> A(P), a(r);
    for (size_t m = 0; m < P; m++) A[m] = p.eval(std::cos(2.0 * kPi * (double)m / (double)P));
    for (size_t k = 0; k < r; k++) {
        double acc = 0.0;
        for (size_t m = 0; m < P; m++) acc += A[m] * std::cos(2.0 * kPi * (double)m * (double)k / (double)P);
        a[k] = acc / (double)P * (k == 0 ? 1.0 : 2.0);
    }
    return a;
}

} // namespace

std::vector<double> remez_lowpass(size_t n, double pass_edge, double stop_edge, double stop_weight)
{
    if (n == 0) return {};
    if (n < 3) return std::vector<double>(n, 1.0 / (double)n);
    if (!(pass_edge > 0.0) || !(stop_edge > pass_edge) || !(stop_edge < 0.5) || !(stop_weight > 0.0))
        throw std::runtime_error("remez: band edges must satisfy 0 < pass < stop < 0.5 and the weight must be positive");
    const bool odd = (n % 2) == 1;
    const size_t r = odd ? (n + 1) / 2 : n / 2;              // number of cosine basis functions
    // dense grid over the two bands; even length: H = cos(w/2) A'(w), so fit D / Q with weight W Q
    const size_t dens = 16;
    std::vector<double> gf, gD, gW;
    const double span = pass_edge + (0.5 - stop_edge);
    const size_t total = std::max<size_t>(dens * r, 4 * (r + 2));
    const size_t np = std::max<size_t>(2, (size_t)std::lround((double)total * pass_edge / span));
    const size_t ns = std::max<size_t>(2, total - np);
    const double top = odd ? 0.5 : 0.5 - 0.25 / (double)total;   // w = pi is a forced zero of even-length filters
    for (size_t i = 0; i < np; i++) { gf.push_back(pass_edge * (double)i / (double)(np - 1)); gD.push_back(1.0); gW.push_back(1.0); }
    for (size_t i = 0; i < ns; i++) { gf.push_back(stop_edge + (top - stop_edge) * (double)i / (double)(ns - 1)); gD.push_back(0.0); gW.push_back(stop_weight); }
    const size_t G = gf.size();
    std::vector<double> gx(G);
    for (size_t g = 0; g < G; g++) {
        gx[g] = std::cos(2.0 * kPi * gf[g]);
        if (!odd) { const double Q = std::cos(kPi * gf[g]); gD[g] /= Q; gW[g] *= Q; }
    }
    std::vector<size_t> ext(r + 1);
    for (size_t i = 0; i <= r; i++) ext[i] = (size_t)((double)i * (double)(G - 1) / (double)r);
    Bary p;
    std::vector<double> E(G);
    for (int iter = 0; iter < 60; iter++) {
        p.x.resize(r + 1); p.y.resize(r + 1);
        for (size_t i = 0; i <= r; i++) p.x[i] = gx[ext[i]];
        p.b = bary_weights(p.x);
        double num = 0.0, den = 0.0, sgn = 1.0;
        for (size_t i = 0; i <= r; i++) { num += p.b[i] * gD[ext[i]]; den += p.b[i] * sgn / gW[ext[i]]; sgn = -sgn; }
        const double delta = num / den;
        sgn = 1.0;
        for (size_t i = 0; i <= r; i++) { p.y[i] = gD[ext[i]] - sgn * delta / gW[ext[i]]; sgn = -sgn; }
        for (size_t g = 0; g < G; g++) E[g] = gW[g] * (gD[g] - p.eval(gx[g]));
        // local extrema of the weighted error, alternating in sign
        std::vector<size_t> cand;
        for (size_t g = 0; g < G; g++) {
            const double e = E[g], l = g > 0 ? E[g - 1] : (e > 0 ? -1e300 : 1e300), h = g + 1 < G ? E[g + 1] : (e > 0 ? -1e300 : 1e300);
            const bool mx = e > 0 && e >= l && e >= h, mn = e < 0 && e <= l && e <= h;
            if (!mx && !mn) continue;
            if (!cand.empty() && (E[cand.back()] > 0) == (e > 0)) {
                if (std::fabs(e) > std::fabs(E[cand.back()])) cand.back() = g;
            } else {
                cand.push_back(g);
            }
        }
        if (cand.size() < r + 1) break;                      // numerically converged / degenerate: keep the current fit
        while (cand.size() > r + 1) {
            if (std::fabs(E[cand.front()]) < std::fabs(E[cand.back()])) cand.erase(cand.begin());
            else cand.pop_back();
        }
        if (cand == ext) break;
        ext = cand;
    }
    const std::vector<double> a = cosine_coeffs(p, r);
    std::vector<double> h(n, 0.0);
    if (odd) {
        const size_t mid = r - 1;
        h[mid] = a[0];
        for (size_t k = 1; k < r; k++) h[mid - k] = h[mid + k] = a[k] / 2.0;
    } else {
        // H(w) = sum_{j=1..r} d_j cos((j - 1/2) w),  d_j = 2 h[n/2 - j]
        std::vector<double> d(r + 1, 0.0);
        for (size_t j = 1; j <= r; j++) {
            const double bjm1 = a[j - 1], bj = j < r ? a[j] : 0.0;
            d[j] = j == 1 ? a[0] + (r > 1 ? a[1] / 2.0 : 0.0) : (bjm1 + bj) / 2.0;
        }
        for (size_t j = 1; j <= r; j++) h[n / 2 - j] = h[n / 2 - 1 + j] = d[j] / 2.0;
    }
    return h;
}

static double ripple_delta_pass(double pass_db) { const double x = std::pow(10.0, pass_db / 20.0); return (x - 1.0) / (x + 1.0); }
static double ripple_delta_stop(double stop_db) { return std::pow(10.0, -stop_db / 20.0); }

size_t remez_estimate_num_taps(double trans_bw, double pass_db, double stop_db)
{
    // Kaiser: N ~ (-20 log10 sqrt(d1 d2) - 13) / (14.6 df) + 1
    const double d1 = ripple_delta_pass(pass_db), d2 = ripple_delta_stop(stop_db);
    const double v = (-20.0 * std::log10(std::sqrt(d1 * d2)) - 13.0) / (14.6 * trans_bw) + 1.0;
    return v < 1.0 ? 1 : (size_t)std::ceil(v);
}
double remez_estimate_weight(double pass_db, double stop_db) { return ripple_delta_pass(pass_db) / ripple_delta_stop(stop_db); }
double remez_estimate_bw(size_t num_taps, double pass_db, double stop_db)
{
    const double d1 = ripple_delta_pass(pass_db), d2 = ripple_delta_stop(stop_db);
    return (-20.0 * std::log10(std::sqrt(d1 * d2)) - 13.0) / (14.6 * std::max<double>(1.0, (double)num_taps - 1.0));
}
double remez_estimate_atten(size_t num_taps, double trans_bw, double pass_db)
{
    // invert Kaiser's formula for d2 given N, df and d1; returned as positive dB
    const double d1 = ripple_delta_pass(pass_db);
    const double a = 14.6 * trans_bw * ((double)num_taps - 1.0) + 13.0;   // = -20 log10 sqrt(d1 d2)
    return 2.0 * a + 20.0 * std::log10(d1);
}

// -------------------------------------------------------------- low-pass prototypes ---
static std::vector<double> sinc_lowpass(size_t n, double fc)
{
    std::vector<double> h(n);
    const double mid = ((double)n - 1.0) / 2.0;
    for (size_t i = 0; i < n; i++) h[i] = 2.0 * fc * sinc(2.0 * fc * ((double)i - mid));
    return h;
}

static void unit_dc_gain(std::vector<double> &h)
{
    double s = 0.0;
    for (double v : h) s += v;
    if (std::fabs(s) > 1e-300) for (double &v : h) v /= s;
}

static std::vector<double> raised_cosine(size_t n, double alpha, double T)
{
    // T = samples per symbol; -6 dB point at 1/(2T)
    std::vector<double> h(n);
    const double mid = ((double)n - 1.0) / 2.0;
    for (size_t i = 0; i < n; i++) {
        const double t = ((double)i - mid) / T, den = 1.0 - 4.0 * alpha * alpha * t * t;
        h[i] = std::fabs(den) < 1e-9 ? (kPi / 4.0) * sinc(1.0 / (2.0 * alpha)) : sinc(t) * std::cos(kPi * alpha * t) / den;
    }
    unit_dc_gain(h);
    return h;
}

static std::vector<double> root_raised_cosine(size_t n, double alpha, double T)
{
    std::vector<double> h(n);
    const double mid = ((double)n - 1.0) / 2.0;
    for (size_t i = 0; i < n; i++) {
        const double t = ((double)i - mid) / T;
        double v;
        if (std::fabs(t) < 1e-12) v = 1.0 - alpha + 4.0 * alpha / kPi;
        else if (alpha > 0.0 && std::fabs(std::fabs(4.0 * alpha * t) - 1.0) < 1e-9)
            v = alpha / std::sqrt(2.0) * ((1.0 + 2.0 / kPi) * std::sin(kPi / (4.0 * alpha)) + (1.0 - 2.0 / kPi) * std::cos(kPi / (4.0 * alpha)));
        else
            v = (std::sin(kPi * t * (1.0 - alpha)) + 4.0 * alpha * t * std::cos(kPi * t * (1.0 + alpha))) /
                (kPi * t * (1.0 - 16.0 * alpha * alpha * t * t));
        h[i] = v;
    }
    unit_dc_gain(h);
    return h;
}

static std::vector<double> gaussian(size_t n, double bt)
{
    // Gaussian pulse with 3 dB bandwidth `bt` (cycles per sample): h(t) ~ exp(-2 pi^2 bt^2 t^2 / ln 2)
    std::vector<double> h(n);
    const double mid = ((double)n - 1.0) / 2.0, c = 2.0 * kPi * kPi * bt * bt / std::log(2.0);
    for (size_t i = 0; i < n; i++) { const double t = (double)i - mid; h[i] = std::exp(-c * t * t); }
    unit_dc_gain(h);
    return h;
}

static std::vector<double> maxflat(size_t n, double fc)
{
    // Herrmann's maximally flat linear-phase low-pass: with x = sin^2(w/2) and half order Nh,
    //   H(x) = (1 - x)^K sum_{d < Nh+1-K} C(K-1+d, d) x^d,
    // K zeros at w = pi and Nh-K vanishing derivatives at w = 0.  Only the integer K is free, so the
    // half-power frequency moves in discrete steps; K is picked so that H(x_c) is closest to 1/2.
    if (n < 3) return std::vector<double>(n, 1.0 / (double)std::max<size_t>(n, 1));
    const size_t nodd = (n % 2) ? n : n - 1, Nh = (nodd - 1) / 2;
    auto response = [&](size_t K, double x) {
        double sum = 0.0, term = 1.0;                        // term_d = C(K-1+d, d) x^d
        for (size_t d = 0; d + K < Nh + 1; d++) {
            sum += term;
            term *= x * (double)(K + d) / (double)(d + 1);
        }
        return std::pow(1.0 - x, (double)K) * sum;
    };
    const double xc = std::pow(std::sin(kPi * fc), 2.0);
    size_t bestK = 1;
    double best = 1e300;
    for (size_t K = 1; K <= Nh; K++) {
        const double e = std::fabs(response(K, xc) - 0.5);
        if (e < best) { best = e; bestK = K; }
    }
    // H is a polynomial of degree Nh in cos(w): nodd uniform samples determine the taps exactly
    std::vector<double> A(nodd), h(n, 0.0);
    for (size_t m = 0; m < nodd; m++) A[m] = response(bestK, std::pow(std::sin(kPi * (double)m / (double)nodd), 2.0));
    for (size_t i = 0; i < nodd; i++) {
        double acc = 0.0;
        for (size_t m = 0; m < nodd; m++) acc += A[m] * std::cos(2.0 * kPi * (double)m * ((double)i - (double)Nh) / (double)nodd);
        h[i] = acc / (double)nodd;
    }
    return h;
}

static std::vector<double> prototype(const std::string &type, size_t n, double bw, double alpha, double weight)
{
    if (!(bw > 0.0) || !(bw < 0.5)) throw std::runtime_error("design_fir: prototype bandwidth must lie in (0, 0.5)");
    if (type == "sinc") return sinc_lowpass(n, bw);
    if (type == "raised_cosine" || type == "raisedcosine") return raised_cosine(n, alpha, 0.5 / bw);
    if (type == "root_raised_cosine" || type == "rootraisedcosine") return root_raised_cosine(n, alpha, 0.5 / bw);
    if (type == "gaussian") return gaussian(n, bw);
    if (type == "maxflat") return maxflat(n, bw);
    if (type == "remez") {
        // the pass band ends half a transition width below the nominal edge, the stop band starts half above
        const double lo = bw - alpha / 2.0, hi = bw + alpha / 2.0;
        if (!(lo > 0.0) || !(hi < 0.5)) throw std::runtime_error("design_fir: remez transition band leaves (0, 0.5)");
        return remez_lowpass(n, lo, hi, weight > 0.0 ? weight : 1.0);
    }
    throw std::runtime_error("design_fir: unknown filter type '" + type + "'");
}

std::vector<double> design_fir(const std::string &type, const std::string &band, size_t n, double fl, double fu, double alpha,
                               double weight)
{
    if (n == 0) throw std::runtime_error("design_fir: no taps");
    const double mid = ((double)n - 1.0) / 2.0;
    if (band == "LOW_PASS") return prototype(type, n, fl, alpha, weight);
    if (band == "HIGH_PASS") {
        // low-pass of width 1/2 - fl moved to the Nyquist frequency
        std::vector<double> h = prototype(type, n, 0.5 - fl, alpha, weight);
        for (size_t i = 0; i < n; i++) h[i] *= std::cos(kPi * ((double)i - mid));
        return h;
    }
    if (band == "BAND_PASS" || band == "BAND_STOP") {
        if (!(fu > fl)) throw std::runtime_error("design_fir: upper frequency must exceed the lower one");
        const double bw = (fu - fl) / 2.0, fc = (fu + fl) / 2.0;
        std::vector<double> h = prototype(type, n, bw, alpha, weight);
        for (size_t i = 0; i < n; i++) h[i] *= 2.0 * std::cos(2.0 * kPi * fc * ((double)i - mid));
        if (band == "BAND_STOP") {
            if (n % 2 == 0) throw std::runtime_error("design_fir: band stop needs an odd number of taps");
            for (double &v : h) v = -v;
            h[(n - 1) / 2] += 1.0;                           // delta - band pass
        }
        return h;
    }
    throw std::runtime_error("design_fir: unknown band type '" + band + "'");
}

std::vector<std::complex<double>> design_complex_fir(const std::string &type, const std::string &band, size_t n, double fl,
                                                     double fu, double alpha, double weight)
{
    if (n == 0) throw std::runtime_error("design_complex_fir: no taps");
    if (band != "COMPLEX_BAND_PASS" && band != "COMPLEX_BAND_STOP") throw std::runtime_error("design_complex_fir: unknown band type '" + band + "'");
    if (!(fu > fl)) throw std::runtime_error("design_complex_fir: upper frequency must exceed the lower one");
    const double bw = (fu - fl) / 2.0, fc = (fu + fl) / 2.0, mid = ((double)n - 1.0) / 2.0;
    const std::vector<double> p = prototype(type, n, bw, alpha, weight);
    std::vector<std::complex<double>> h(n);
    for (size_t i = 0; i < n; i++) h[i] = p[i] * std::polar(1.0, 2.0 * kPi * fc * ((double)i - mid));
    if (band == "COMPLEX_BAND_STOP") {
        if (n % 2 == 0) throw std::runtime_error("design_complex_fir: band stop needs an odd number of taps");
        for (auto &v : h) v = -v;
        h[(n - 1) / 2] += 1.0;
    }
    return h;
}

} // namespace b200c_design
