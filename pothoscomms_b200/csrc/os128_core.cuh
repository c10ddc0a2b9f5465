// 4096-point transform on 128 threads x 32 points (fir_os128_kernel, fir_os.cu): the three-step decomposition
//   n = 128 n1 + n2 (n1 < 32, n2 < 128),   n2 = j + 32 q (j < 32, q < 4),   k = k1 + 32 k2,   k2 = 4 m + r (m < 32, r < 4)
//   X[k1 + 32 (4 m + r)] = sum_j W32^(j m) W128^(j r) sum_q W4^(q r) { W4096^(n2 k1) sum_n1 x[128 n1 + n2] W32^(n1 k1) }
// Thread roles: step 1 (and its inverse mirror) thread = n2 (warp q, lane j); step 2 thread = (lane k1, warp r).  One
// shared-memory tile T[k1][c], c < 128, row stride 129, carries both exchanges: step 1 writes column n2 of every row,
// step 2 reads its row k1 -- ALL 128 entries, combining the four quarters (the radix-4 stage over q) on the way, so the
// four threads of a row need no exchange among themselves; the inverse mirrors it (step 2' writes its 32 values into
// quarter r of row k1, step 1' reads the four quarters of column j and combines them over r).  32 points (64 registers)
// per thread: half the register state of the 64 x 64 kernel (os64_core.cuh), which is what lets 12-16 warps share an SM.
// Compiles for the host as well (tools/os128_host_check.cpp replays one block against a direct convolution).
#pragma once
#include "os64_core.cuh"

namespace b200c {

// cos(2 pi e / 128), e = 0..127, rounded to float
B200C_HD constexpr float cos128(int e)
{
    constexpr float t[128] = {
        1.0f, 0.99879545f, 0.99518472f, 0.989176512f, 0.980785251f, 0.970031261f, 0.956940353f, 0.941544056f, 0.923879504f,
        0.903989315f, 0.881921291f, 0.857728601f, 0.831469595f, 0.803207517f, 0.773010433f, 0.740951121f, 0.707106769f,
        0.671558976f, 0.634393275f, 0.59569931f, 0.555570245f, 0.514102757f, 0.471396744f, 0.427555084f, 0.382683426f,
        0.336889863f, 0.290284663f, 0.242980182f, 0.195090324f, 0.146730468f, 0.0980171412f, 0.0490676761f, 0.0f,
        -0.0490676761f, -0.0980171412f, -0.146730468f, -0.195090324f, -0.242980182f, -0.290284663f, -0.336889863f,
        -0.382683426f, -0.427555084f, -0.471396744f, -0.514102757f, -0.555570245f, -0.59569931f, -0.634393275f,
        -0.671558976f, -0.707106769f, -0.740951121f, -0.773010433f, -0.803207517f, -0.831469595f, -0.857728601f,
        -0.881921291f, -0.903989315f, -0.923879504f, -0.941544056f, -0.956940353f, -0.970031261f, -0.980785251f,
        -0.989176512f, -0.99518472f, -0.99879545f, -1.0f, -0.99879545f, -0.99518472f, -0.989176512f, -0.980785251f,
        -0.970031261f, -0.956940353f, -0.941544056f, -0.923879504f, -0.903989315f, -0.881921291f, -0.857728601f,
        -0.831469595f, -0.803207517f, -0.773010433f, -0.740951121f, -0.707106769f, -0.671558976f, -0.634393275f,
        -0.59569931f, -0.555570245f, -0.514102757f, -0.471396744f, -0.427555084f, -0.382683426f, -0.336889863f,
        -0.290284663f, -0.242980182f, -0.195090324f, -0.146730468f, -0.0980171412f, -0.0490676761f, 0.0f, 0.0490676761f,
        0.0980171412f, 0.146730468f, 0.195090324f, 0.242980182f, 0.290284663f, 0.336889863f, 0.382683426f, 0.427555084f,
        0.471396744f, 0.514102757f, 0.555570245f, 0.59569931f, 0.634393275f, 0.671558976f, 0.707106769f, 0.740951121f,
        0.773010433f, 0.803207517f, 0.831469595f, 0.857728601f, 0.881921291f, 0.903989315f, 0.923879504f, 0.941544056f,
        0.956940353f, 0.970031261f, 0.980785251f, 0.989176512f, 0.99518472f, 0.99879545f};
    return t[e & 127];
}
B200C_HD constexpr float sin128(int e) { return cos128(e + 96); }   // sin(x) = cos(x - pi/2)

// f * W128^e (INV: conjugate), e a compile-time constant at every call site
template <bool INV> B200C_HD c2 mul_w128(c2 f, int e)
{
    e &= 127;
    if (e == 0) return f;
    if (e == 32) return rot_p<INV>(f);
    if (e == 96) return rot_p<!INV>(f);
    if (e == 64) return sub2(pk(0.f, 0.f), f);
    const float c = cos128(e), s = INV ? sin128(e) : -sin128(e);
    return cmul_s(f, c, s);
}

// sum_q z_q W4^(q R)  (INV: conjugate powers), R a compile-time constant
template <int R, bool INV> B200C_HD c2 comb4(c2 z0, c2 z1, c2 z2, c2 z3)
{
    const c2 a = (R & 1) ? sub2(z0, z2) : add2(z0, z2);
    const c2 b = (R & 1) ? sub2(z1, z3) : add2(z1, z3);
    if (R == 0) return add2(a, b);
    if (R == 2) return sub2(a, b);
    if (R == 1) return add2(a, rot_p<INV>(b));      // forward: a - i b
    return sub2(a, rot_p<INV>(b));                  // R == 3, forward: a + i b
}

constexpr int kOs128Stride = 129;                            // row stride of the exchange tile (c2 elements)
constexpr int kOs128SmemElems = 32 * kOs128Stride;           // 4128 >= the 4098 elements a bulk copy lands

// step 2 of the forward transform for the thread (lane k1, warp R): gathers row k1 of the tile -- the radix-4 stage over
// the four quarters and the W128^(j R) twiddle -- into v[rev32(j)]; a dft32_dit then leaves X[k1 + 32 (4 m + R)] in v[m]
template <int R> B200C_HD void os128_fwd_gather(c2 (&v)[32], const c2 *row)
{
#pragma unroll
    for (int j = 0; j < 32; j++) {
        const c2 u = comb4<R, false>(row[j], row[j + 32], row[j + 64], row[j + 96]);
        v[rev32(j)] = mul_w128<false>(u, j * R);
    }
}

// the mirror image: after a dft32_dif of the (filtered) spectrum bins k1 + 32 (4 m + R), v[rev32(j)] holds the thread's 32
// values; they go, twiddled, into quarter R of row k1
template <int R> B200C_HD void os128_inv_scatter(const c2 (&v)[32], c2 *row)
{
#pragma unroll
    for (int j = 0; j < 32; j++) row[j + 32 * R] = mul_w128<true>(v[rev32(j)], j * R);
}

// step 1' for the thread n2 = j + 32 Q: element (row k1) of column j, combined over the four quarters
template <int Q> B200C_HD c2 os128_inv_col(const c2 *tile, int k1, int j)
{
    const c2 *p = tile + k1 * kOs128Stride + j;
    return comb4<Q, true>(p[0], p[32], p[64], p[96]);
}

} // namespace b200c
