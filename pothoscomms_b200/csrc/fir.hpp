// Host-side description of one configured FIR / polyphase resampler (libb200comms.so).
// Mirrors the state a reference FIRFilter block holds (filter/FIRFilter.cpp:356-363) in the
// form the sm_100a kernels consume.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "common.hpp"

namespace b200c {

// Polyphase decomposition of the reference loop nest (filter/FIRFilter.cpp:286-302).
//
// Output m = q*L + p (p = "output slot", q = block of M inputs / L outputs).  The reference
// emits an output when the virtual up-sampled index i = n*L + j satisfies (i+1) % M == 0, so
//   i = (m+1)*M - 1 = q*L*M + (p+1)*M - 1,  j_p = ((p+1)*M - 1) % L,  d_p = ((p+1)*M - 1) / L,
//   y[q*L + p] = sum_k h_{j_p}[k] * x[q*M + d_p - k],            h_j[k] = taps[j + k*L].
// Splitting the input by residue e = (d_p - k) mod M turns every (p, e) pair into a plain
// stride-1 correlation over the de-interleaved stream x_e[i] = x[i*M + e]:
//   y[q*L + p] = sum_e sum_t G_{p,e}[t] * x_e[q + off_{p,e} + t],   t = 0..S-1
// with G stored in forward order and zero padded to a common length S (a multiple of the
// kernel's register block R).  One table entry per (p, e): `nsub = L*M` of them.
struct FirTable {
    int dtype = B200C_CF32;
    int taps_kind = B200C_TAPS_REAL;
    size_t M = 1, L = 1, K = 1, ntaps = 1;
    int R = 9;       // outputs per thread (odd => conflict-free shared-memory windows)
    int S = 0;       // padded sub-filter length, multiple of R
    int nsub = 1;    // L*M
    int lo = 0;      // min over (p,e) of off
    int hi = 0;      // max over (p,e) of off
    std::vector<int> off;           // [nsub]
    std::vector<uint8_t> taps;      // [nsub][S] tap elements in the accumulator type
    size_t tap_elem_bytes = 0;      // sizeof(accumulator scalar) * (COMPLEX ? 2 : 1)
    bool smem_path = true;          // false => generic one-output-per-thread kernel
    int nrb = 8;                    // 32*R*nrb blocks (q) per CTA tile
};

// accumulator scalar width per data type, filter/FIRFilter.cpp:377-382 (int8 -> int16 math is
// done mod 2^16 inside 32-bit registers)
inline size_t acc_scalar_bytes(int dt)
{
    switch (dt >> 1) { case 0: return 4; case 1: return 8; case 2: return 4; case 3: return 4; default: return 8; }
}
// QTapsType scalar width (what floatToQ targets), filter/FIRFilter.cpp:377-382
inline size_t qtaps_scalar_bytes(int dt)
{
    switch (dt >> 1) { case 0: return 4; case 1: return 8; case 2: return 2; case 3: return 4; default: return 8; }
}

// updateInternals(): builds the table from double taps.  Returns B200C_* status.
int fir_build_table(FirTable &t, int dtype, int taps_kind, const double *taps, size_t ntaps, size_t M, size_t L,
                    size_t smem_budget);

struct FirDeviceState {
    void *d_taps = nullptr;  // [nsub][S]
    int *d_off = nullptr;    // [nsub]
    size_t taps_capacity = 0, off_capacity = 0;
};

// Launch the convolution nest for `nblocks` = N/M blocks on `stream`.
int fir_launch(const FirTable &t, const FirDeviceState &ds, const void *d_in, size_t in_elems, void *d_out,
               size_t nblocks, int sm_count, cudaStream_t stream);

size_t fir_smem_bytes(const FirTable &t, int nrb);

} // namespace b200c
