// CPU check of the index logic of fir_os64_kernel (pothoscomms_b200/csrc/fir_os.cu): replays
// one overlap-save block with the kernel's own register-transform code (os64_core.cuh compiled
// for the host) and compares with a direct double-precision convolution.
//   g++ -std=c++17 -O2 -I pothoscomms_b200/csrc -o /tmp/os64_check tools/os64_host_check.cpp && /tmp/os64_check
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "os64_core.cuh"
using namespace b200c;
typedef std::complex<double> cd;

static c2 from(cd z) { return pk((float)z.real(), (float)z.imag()); }
static cd to(c2 p) { float a, b; upk(p, a, b); return cd(a, b); }

int main()
{
    const int N = 4096, K = 200;
    std::mt19937 rng(1);
    std::normal_distribution<double> g;
    std::vector<cd> x(N), h(K);
    for (auto &v : x) v = cd(g(rng), g(rng));
    for (auto &v : h) v = cd(g(rng), g(rng)) / (double)K;
    const double PI = 3.14159265358979323846;
    std::vector<c2> hf(N), twa(8 * 64), twb(8 * 64);
    for (int f = 0; f < N; f++) {
        cd s = 0;
        for (int k = 0; k < K; k++) s += h[k] * std::polar(1.0, -2 * PI * ((long long)f * k % N) / N);
        hf[f] = from(s / (double)N);
    }
    for (int i = 0; i < 8; i++)
        for (int t = 0; t < 64; t++) {
            twa[i * 64 + t] = from(std::polar(1.0, -2 * PI * ((8 * i * t) % N) / N));
            twb[i * 64 + t] = from(std::polar(1.0, -2 * PI * ((i * t) % N) / N));
        }
    static c2 v[64][64];
    std::vector<c2> F(kOs64SmemElems);
    auto step_tw = [&](c2 (&r)[64], int t, bool conj, bool slot_rev) {
        for (int a = 0; a < 8; a++)
            for (int b = 0; b < 8; b++) {
                const int j = 8 * a + b, s = slot_rev ? rev64(j) : j;
                if (a) r[s] = conj ? cmul_p<true>(r[s], twa[a * 64 + t]) : cmul_p<false>(r[s], twa[a * 64 + t]);
                if (b) r[s] = conj ? cmul_p<true>(r[s], twb[b * 64 + t]) : cmul_p<false>(r[s], twb[b * 64 + t]);
            }
    };
    for (int t = 0; t < 64; t++) {
        for (int n1 = 0; n1 < 64; n1++) v[t][rev64(n1)] = from(x[64 * n1 + t]);
        dft64_dit<false>(v[t]);
        step_tw(v[t], t, false, false);
        for (int k1 = 0; k1 < 64; k1++) F[k1 * kOs64Stride + t] = v[t][k1];
    }
    // check the forward transform on the way
    double ferr = 0, fref = 0;
    for (int t = 0; t < 64; t++) {
        for (int n2 = 0; n2 < 64; n2++) v[t][rev64(n2)] = F[t * kOs64Stride + n2];
        dft64_dit<false>(v[t]);
        for (int k2 = 0; k2 < 64; k2 += 13) {
            const int k = t + 64 * k2;
            cd s = 0;
            for (int n = 0; n < N; n++) s += x[n] * std::polar(1.0, -2 * PI * ((long long)n * k % N) / N);
            ferr += std::norm(to(v[t][k2]) - s); fref += std::norm(s);
        }
        for (int k2 = 0; k2 < 64; k2++) v[t][k2] = cmul_p<false>(v[t][k2], hf[64 * k2 + t]);
        dft64_dif<true>(v[t]);
        step_tw(v[t], t, true, true);
    }
    for (int t = 0; t < 64; t++)
        for (int n2 = 0; n2 < 64; n2++) F[t * kOs64Stride + n2] = v[t][rev64(n2)];
    double err = 0, ref = 0;
    for (int t = 0; t < 64; t++) {
        for (int k1 = 0; k1 < 64; k1++) v[t][k1] = F[k1 * kOs64Stride + t];
        dft64_dif<true>(v[t]);
        for (int n1 = 0; n1 < 64; n1++) {
            const int i = 64 * n1 + t;
            if (i < K - 1) continue;
            cd s = 0;
            for (int k = 0; k < K; k++) s += h[k] * x[i - k];
            err += std::norm(to(v[t][rev64(n1)]) - s); ref += std::norm(s);
        }
    }
    printf("forward rel err %.3g, overlap-save rel err %.3g\n", std::sqrt(ferr / fref), std::sqrt(err / ref));
    return (std::sqrt(err / ref) < 2e-6 && std::sqrt(ferr / fref) < 2e-6) ? 0 : 1;
}
