// CPU check of the 32 x 32 (1024-point) overlap-save index logic of fir_os32_kernel.
//   g++ -std=c++17 -O2 -I pothoscomms_b200/csrc -o /tmp/os32_check tools/os32_host_check.cpp && /tmp/os32_check
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
    const int N = 1024, K = 100;
    std::mt19937 rng(1);
    std::normal_distribution<double> g;
    std::vector<cd> x(N), h(K);
    for (auto &v : x) v = cd(g(rng), g(rng));
    for (auto &v : h) v = cd(g(rng), g(rng)) / (double)K;
    const double PI = 3.14159265358979323846;
    std::vector<c2> hf(N), tw(32 * 32);
    for (int f = 0; f < N; f++) {
        cd s = 0;
        for (int k = 0; k < K; k++) s += h[k] * std::polar(1.0, -2 * PI * ((long long)f * k % N) / N);
        hf[f] = from(s / (double)N);
    }
    for (int j = 0; j < 32; j++)
        for (int t = 0; t < 32; t++) tw[j * 32 + t] = from(std::polar(1.0, -2 * PI * ((j * t) % N) / N));
    static c2 v[32][32];
    std::vector<c2> F(kOs32SmemElems);
    for (int t = 0; t < 32; t++) {
        for (int n1 = 0; n1 < 32; n1++) v[t][rev32(n1)] = from(x[32 * n1 + t]);
        dft32_dit<false>(v[t]);
        for (int k1 = 0; k1 < 32; k1++) v[t][k1] = cmul_p<false>(v[t][k1], tw[k1 * 32 + t]);
        for (int k1 = 0; k1 < 32; k1++) F[k1 * kOs32Stride + t] = v[t][k1];
    }
    double ferr = 0, fref = 0;
    for (int t = 0; t < 32; t++) {
        for (int n2 = 0; n2 < 32; n2++) v[t][rev32(n2)] = F[t * kOs32Stride + n2];
        dft32_dit<false>(v[t]);
        for (int k2 = 0; k2 < 32; k2 += 5) {
            const int k = t + 32 * k2;
            cd s = 0;
            for (int n = 0; n < N; n++) s += x[n] * std::polar(1.0, -2 * PI * ((long long)n * k % N) / N);
            ferr += std::norm(to(v[t][k2]) - s); fref += std::norm(s);
        }
        for (int k2 = 0; k2 < 32; k2++) v[t][k2] = cmul_p<false>(v[t][k2], hf[32 * k2 + t]);
        dft32_dif<true>(v[t]);
        for (int n2 = 0; n2 < 32; n2++) v[t][rev32(n2)] = cmul_p<true>(v[t][rev32(n2)], tw[n2 * 32 + t]);
    }
    for (int t = 0; t < 32; t++)
        for (int n2 = 0; n2 < 32; n2++) F[t * kOs32Stride + n2] = v[t][rev32(n2)];
    double err = 0, ref = 0;
    for (int t = 0; t < 32; t++) {
        for (int k1 = 0; k1 < 32; k1++) v[t][k1] = F[k1 * kOs32Stride + t];
        dft32_dif<true>(v[t]);
        for (int n1 = 0; n1 < 32; n1++) {
            const int i = 32 * n1 + t;
            if (i < K - 1) continue;
            cd s = 0;
            for (int k = 0; k < K; k++) s += h[k] * x[i - k];
            err += std::norm(to(v[t][rev32(n1)]) - s); ref += std::norm(s);
        }
    }
    printf("forward rel err %.3g, overlap-save rel err %.3g\n", std::sqrt(ferr / fref), std::sqrt(err / ref));
    return (std::sqrt(err / ref) < 2e-6 && std::sqrt(ferr / fref) < 2e-6) ? 0 : 1;
}
