// CPU check of the index logic of fir_os128_kernel (pothoscomms_b200/csrc/fir_os.cu): replays one overlap-save
// block with the kernel's own register-transform code (os128_core.cuh compiled for the host), 128 "threads" run one
// after the other, and compares with a direct double-precision convolution.
//   g++ -std=c++17 -O2 -I pothoscomms_b200/csrc -o /tmp/os128_check tools/os128_host_check.cpp && /tmp/os128_check
#include <cmath>
#include <complex>
#include <cstdio>
#include <random>
#include <vector>

#include "os128_core.cuh"
using namespace b200c;
typedef std::complex<double> cd;

static c2 from(cd z) { return pk((float)z.real(), (float)z.imag()); }
static cd to(c2 p) { float a, b; upk(p, a, b); return cd(a, b); }

template <int R> static void fwd_rows(c2 (&v)[32], const c2 *row) { os128_fwd_gather<R>(v, row); dft32_dit<false>(v); }
template <int R> static void inv_rows(c2 (&v)[32], c2 *row) { dft32_dif<true>(v); os128_inv_scatter<R>(v, row); }

int main()
{
    const int N = 4096, K = 1000;
    std::mt19937 rng(1);
    std::normal_distribution<double> g;
    std::vector<cd> x(N), h(K);
    for (auto &v : x) v = cd(g(rng), g(rng));
    for (auto &v : h) v = cd(g(rng), g(rng)) / (double)K;
    const double PI = 3.14159265358979323846;
    std::vector<c2> hf(N), tw1(32 * 128);
    for (int f = 0; f < N; f++) {
        cd s = 0;
        for (int k = 0; k < K; k++) s += h[k] * std::polar(1.0, -2 * PI * ((long long)f * k % N) / N);
        hf[f] = from(s / (double)N);
    }
    for (int k1 = 0; k1 < 32; k1++)
        for (int n2 = 0; n2 < 128; n2++) tw1[k1 * 128 + n2] = from(std::polar(1.0, -2 * PI * ((k1 * n2) % N) / N));
    static c2 v[128][32];
    std::vector<c2> T(kOs128SmemElems);
    // step 1: thread n2
    for (int n2 = 0; n2 < 128; n2++) {
        for (int n1 = 0; n1 < 32; n1++) v[n2][rev32(n1)] = from(x[128 * n1 + n2]);
        dft32_dit<false>(v[n2]);
        for (int k1 = 1; k1 < 32; k1++) v[n2][k1] = cmul_p<false>(v[n2][k1], tw1[k1 * 128 + n2]);
        for (int k1 = 0; k1 < 32; k1++) T[k1 * kOs128Stride + n2] = v[n2][k1];
    }
    // step 2: thread (k1, r) = tid k1 + 32 r
    double ferr = 0, fref = 0;
    for (int tid = 0; tid < 128; tid++) {
        const int k1 = tid & 31, r = tid >> 5;
        const c2 *row = T.data() + k1 * kOs128Stride;
        switch (r) { case 0: fwd_rows<0>(v[tid], row); break; case 1: fwd_rows<1>(v[tid], row); break;
                     case 2: fwd_rows<2>(v[tid], row); break; default: fwd_rows<3>(v[tid], row); break; }
        for (int m = 0; m < 32; m += 7) {
            const int k = k1 + 32 * (4 * m + r);
            cd s = 0;
            for (int n = 0; n < N; n++) s += x[n] * std::polar(1.0, -2 * PI * ((long long)n * k % N) / N);
            ferr += std::norm(to(v[tid][m]) - s); fref += std::norm(s);
        }
        for (int m = 0; m < 32; m++) v[tid][m] = cmul_p<false>(v[tid][m], hf[k1 + 32 * r + 128 * m]);
    }
    // step 2': all threads have read the tile (a barrier in the kernel) before it is overwritten
    for (int tid = 0; tid < 128; tid++) {
        const int k1 = tid & 31, r = tid >> 5;
        c2 *row = T.data() + k1 * kOs128Stride;
        switch (r) { case 0: inv_rows<0>(v[tid], row); break; case 1: inv_rows<1>(v[tid], row); break;
                     case 2: inv_rows<2>(v[tid], row); break; default: inv_rows<3>(v[tid], row); break; }
    }
    // step 1': thread n2 = j + 32 q
    double err = 0, ref = 0;
    for (int n2 = 0; n2 < 128; n2++) {
        const int j = n2 & 31, q = n2 >> 5;
        for (int k1 = 0; k1 < 32; k1++) {
            c2 y;
            switch (q) { case 0: y = os128_inv_col<0>(T.data(), k1, j); break; case 1: y = os128_inv_col<1>(T.data(), k1, j); break;
                         case 2: y = os128_inv_col<2>(T.data(), k1, j); break; default: y = os128_inv_col<3>(T.data(), k1, j); break; }
            v[n2][k1] = k1 ? cmul_p<true>(y, tw1[k1 * 128 + n2]) : y;
        }
        dft32_dif<true>(v[n2]);
        for (int n1 = 0; n1 < 32; n1++) {
            const int i = 128 * n1 + n2;
            if (i < K - 1) continue;
            cd s = 0;
            for (int k = 0; k < K; k++) s += h[k] * x[i - k];
            err += std::norm(to(v[n2][rev32(n1)]) - s); ref += std::norm(s);
        }
    }
    printf("forward rel err %.3g, overlap-save rel err %.3g\n", std::sqrt(ferr / fref), std::sqrt(err / ref));
    return (std::sqrt(err / ref) < 2e-6 && std::sqrt(ferr / fref) < 2e-6) ? 0 : 1;
}
