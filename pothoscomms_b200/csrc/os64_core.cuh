// Register-resident 64-point transforms and the 64 x 64 decomposition of the 4096-point
// transform used by the fused overlap-save FIR kernel (fir_os.cu).
//
// 4096 = 64 * 64 (four-step):  n = 64*n1 + n2,  k = k1 + 64*k2
//   X[k1 + 64 k2] = sum_n2 W64^(n2 k2) * { W4096^(n2 k1) * sum_n1 x[64 n1 + n2] W64^(n1 k1) }
// One thread owns 64 points (128 registers), 64 threads own one transform:
//   step 1: thread n2 transforms x[64 n1 + n2] over n1           -> Y[n2][k1]   (registers)
//           and multiplies by the step twiddle W4096^(n2 k1)
//   exchange through shared memory (the only one per transform)
//   step 2: thread k1 transforms Y[.][k1] over n2                -> X[k1 + 64 k2]
// The inverse runs the mirror image (decimation in frequency, conjugate twiddles), so the
// forward's step 2 and the inverse's first step meet in the SAME registers: per overlap-save
// block there are two shared-memory exchanges instead of the four (2 x 2) of the 16-point-
// per-thread kernel -- the L1/shared data pipe was that kernel's limiter (ncu: 75 % busy).
//
// All register indices are compile-time constants after unrolling, so the base-4 digit
// reversal the in-place transforms need is free (it only renames registers).
#pragma once
#include "packed.cuh"

namespace b200c {

// cos / sin of 2*pi*e/64, e = 0..63, rounded to float
B200C_HD constexpr float cos64(int e)
{
    constexpr float t[64] = {
        1.0f, 0.9951847195625305f, 0.9807852506637573f, 0.9569403529167175f, 0.9238795042037964f, 0.8819212913513184f,
        0.8314695954322815f, 0.7730104327201843f, 0.7071067690849304f, 0.6343932747840881f, 0.5555702447891235f,
        0.4713967442512512f, 0.3826834261417389f, 0.290284663438797f, 0.19509032368659973f, 0.0980171412229538f, 0.0f,
        -0.0980171412229538f, -0.19509032368659973f, -0.290284663438797f, -0.3826834261417389f, -0.4713967442512512f,
        -0.5555702447891235f, -0.6343932747840881f, -0.7071067690849304f, -0.7730104327201843f, -0.8314695954322815f,
        -0.8819212913513184f, -0.9238795042037964f, -0.9569403529167175f, -0.9807852506637573f, -0.9951847195625305f, -1.0f,
        -0.9951847195625305f, -0.9807852506637573f, -0.9569403529167175f, -0.9238795042037964f, -0.8819212913513184f,
        -0.8314695954322815f, -0.7730104327201843f, -0.7071067690849304f, -0.6343932747840881f, -0.5555702447891235f,
        -0.4713967442512512f, -0.3826834261417389f, -0.290284663438797f, -0.19509032368659973f, -0.0980171412229538f, 0.0f,
        0.0980171412229538f, 0.19509032368659973f, 0.290284663438797f, 0.3826834261417389f, 0.4713967442512512f,
        0.5555702447891235f, 0.6343932747840881f, 0.7071067690849304f, 0.7730104327201843f, 0.8314695954322815f,
        0.8819212913513184f, 0.9238795042037964f, 0.9569403529167175f, 0.9807852506637573f, 0.9951847195625305f};
    return t[e & 63];
}
B200C_HD constexpr float sin64(int e) { return cos64(e + 48); }   // sin(x) = cos(x - pi/2)

// f * W64^e, W64 = exp(-2 pi i / 64); INV: f * conj(W64^e).  `e` is a compile-time constant
// at every call site (fully unrolled loops), so the branches and table reads fold away.
template <bool INV> B200C_HD c2 mul_w64(c2 f, int e)
{
    e &= 63;
    if (e == 0) return f;
    if (e == 16) return rot_p<INV>(f);          // forward: * (-i)
    if (e == 48) return rot_p<!INV>(f);
    const float c = cos64(e), s = INV ? sin64(e) : -sin64(e);
    return cmul_s(f, c, s);
}

// position of element n (base-4 digits n = a + 4b + 16c) in the digit-reversed register order
B200C_HD constexpr int rev64(int n) { return 16 * (n & 3) + 4 * ((n >> 2) & 3) + ((n >> 4) & 3); }

// In-place 64-point DFT, decimation in time: input element n in v[rev64(n)], output k in v[k].
template <bool INV> B200C_HD void dft64_dit(c2 (&v)[64])
{
#pragma unroll
    for (int g = 0; g < 16; g++) dft4_p<INV>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = 16 * g + k;
            v[i + 4] = mul_w64<INV>(v[i + 4], 4 * k);
            v[i + 8] = mul_w64<INV>(v[i + 8], 8 * k);
            v[i + 12] = mul_w64<INV>(v[i + 12], 12 * k);
            dft4_p<INV>(v[i], v[i + 4], v[i + 8], v[i + 12]);
        }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        v[k + 16] = mul_w64<INV>(v[k + 16], k);
        v[k + 32] = mul_w64<INV>(v[k + 32], 2 * k);
        v[k + 48] = mul_w64<INV>(v[k + 48], 3 * k);
        dft4_p<INV>(v[k], v[k + 16], v[k + 32], v[k + 48]);
    }
}

// In-place 64-point DFT, decimation in frequency (the transpose of dft64_dit): input element
// n in v[n], output k in v[rev64(k)].
template <bool INV> B200C_HD void dft64_dif(c2 (&v)[64])
{
#pragma unroll
    for (int k = 0; k < 16; k++) {
        dft4_p<INV>(v[k], v[k + 16], v[k + 32], v[k + 48]);
        v[k + 16] = mul_w64<INV>(v[k + 16], k);
        v[k + 32] = mul_w64<INV>(v[k + 32], 2 * k);
        v[k + 48] = mul_w64<INV>(v[k + 48], 3 * k);
    }
#pragma unroll
    for (int g = 0; g < 4; g++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = 16 * g + k;
            dft4_p<INV>(v[i], v[i + 4], v[i + 8], v[i + 12]);
            v[i + 4] = mul_w64<INV>(v[i + 4], 4 * k);
            v[i + 8] = mul_w64<INV>(v[i + 8], 8 * k);
            v[i + 12] = mul_w64<INV>(v[i + 12], 12 * k);
        }
#pragma unroll
    for (int g = 0; g < 16; g++) dft4_p<INV>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
}

// Shared-memory exchange buffer of one transform: element (row r, column c) at r*65 + c.
// Row stride 65 (odd) makes both the row-contiguous and the column accesses conflict free.
constexpr int kOs64Stride = 65;
constexpr int kOs64SmemElems = 64 * kOs64Stride;

// ---------------------------------------------------------------- 32-point transforms ---
// Radix 4, 4, 2.  Element n = q + 2 q' + 8 n'' (q < 2, q' < 4, n'' < 4) sits in register
// rev32(n) = 16 q + 4 q' + n'' for the decimation-in-time input / decimation-in-frequency output.
B200C_HD constexpr int rev32(int n) { return 16 * (n & 1) + 4 * ((n >> 1) & 3) + ((n >> 3) & 3); }

template <bool INV> B200C_HD void dft32_dit(c2 (&v)[32])
{
#pragma unroll
    for (int g = 0; g < 8; g++) dft4_p<INV>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = 16 * g + k;
            v[i + 4] = mul_w64<INV>(v[i + 4], 4 * k);
            v[i + 8] = mul_w64<INV>(v[i + 8], 8 * k);
            v[i + 12] = mul_w64<INV>(v[i + 12], 12 * k);
            dft4_p<INV>(v[i], v[i + 4], v[i + 8], v[i + 12]);
        }
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const c2 b = mul_w64<INV>(v[k + 16], 2 * k);   // W32^k
        v[k + 16] = sub2(v[k], b);
        v[k] = add2(v[k], b);
    }
}

template <bool INV> B200C_HD void dft32_dif(c2 (&v)[32])
{
#pragma unroll
    for (int k = 0; k < 16; k++) {
        const c2 d = sub2(v[k], v[k + 16]);
        v[k] = add2(v[k], v[k + 16]);
        v[k + 16] = mul_w64<INV>(d, 2 * k);
    }
#pragma unroll
    for (int g = 0; g < 2; g++)
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int i = 16 * g + k;
            dft4_p<INV>(v[i], v[i + 4], v[i + 8], v[i + 12]);
            v[i + 4] = mul_w64<INV>(v[i + 4], 4 * k);
            v[i + 8] = mul_w64<INV>(v[i + 8], 8 * k);
            v[i + 12] = mul_w64<INV>(v[i + 12], 12 * k);
        }
#pragma unroll
    for (int g = 0; g < 8; g++) dft4_p<INV>(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
}

// exchange buffer of one 1024-point transform (one warp): element (r, c) at r*33 + c
constexpr int kOs32Stride = 33;
constexpr int kOs32SmemElems = 32 * kOs32Stride;

} // namespace b200c
