// HBM-resident neighbours of the filter path: /comms/scale,
// /comms/rotate and /comms/signal_probe as streaming kernels, so that a topology such as
// source -> scale -> fir_filter -> probe (filter/TestFIRFilter.cpp:49-51) never leaves HBM.
//
//   scale : out = fromQ(floatToQ(factor) * Q(in))              math/Scale.cpp:15-23,41-45
//   rotate: out = fromQ(floatToQ(polar(1, phase)) * Q(in))     math/Rotate.cpp:15-23,71-75
//   probe : VALUE / RMS / MEAN of a window                     utility/SignalProbe.cpp:140-160
// Q types per class as in the reference's factories (Scale.cpp:150-153): f32, f64, int8 -> int16,
// int16 -> int32, int32 -> int64, int64 -> int64; integer products wrap in the Q type, fromQ is an
// arithmetic shift by half the Q word and a wrapping narrow.  Pure streams: 128-bit loads and
// stores with streaming cache hints, grid a multiple of the SM count; algorithmic bytes are
// 2 x sizeof(element) per element (scale, rotate) and 1 x (probe).
#include <cmath>
#include <cstring>

#include "common.hpp"

namespace b200c {

template <int CLS> struct QT;
template <> struct QT<0> { typedef float T; typedef float Q; };
template <> struct QT<1> { typedef double T; typedef double Q; };
template <> struct QT<2> { typedef int8_t T; typedef int16_t Q; };
template <> struct QT<3> { typedef int16_t T; typedef int32_t Q; };
template <> struct QT<4> { typedef int32_t T; typedef int64_t Q; };
template <> struct QT<5> { typedef int64_t T; typedef int64_t Q; };

// Q-type multiply with the reference's wrap (int16 x int16 promotes to int and is truncated on
// the store to the Q type; 64-bit products wrap): done on unsigned values, which is defined.
template <typename Q> __device__ __forceinline__ Q qmul(Q a, Q b)
{
    if constexpr (sizeof(Q) == 2) return (Q)(uint16_t)((uint32_t)(int32_t)a * (uint32_t)(int32_t)b);
    else if constexpr (sizeof(Q) == 4 && !std::is_floating_point<Q>::value) return (Q)((uint32_t)a * (uint32_t)b);
    else if constexpr (sizeof(Q) == 8 && !std::is_floating_point<Q>::value) return (Q)((uint64_t)a * (uint64_t)b);
    else return a * b;
}
template <typename Q> __device__ __forceinline__ Q qsub(Q a, Q b)
{
    if constexpr (std::is_floating_point<Q>::value) return a - b;
    else if constexpr (sizeof(Q) == 8) return (Q)((uint64_t)a - (uint64_t)b);
    else return (Q)((uint32_t)(int32_t)a - (uint32_t)(int32_t)b);
}
template <typename Q> __device__ __forceinline__ Q qadd(Q a, Q b)
{
    if constexpr (std::is_floating_point<Q>::value) return a + b;
    else if constexpr (sizeof(Q) == 8) return (Q)((uint64_t)a + (uint64_t)b);
    else return (Q)((uint32_t)(int32_t)a + (uint32_t)(int32_t)b);
}
template <typename T, typename Q> __device__ __forceinline__ T from_q(Q q)
{
    if constexpr (std::is_floating_point<Q>::value) return (T)q;
    else return (T)(q >> (4 * (int)sizeof(Q)));
}

template <int CLS> struct ScaleOp {
    typedef typename QT<CLS>::T T;
    typedef typename QT<CLS>::Q Q;
    static constexpr int UNIT = 1;
    Q f;
    __device__ __forceinline__ void operator()(T *e) const { e[0] = from_q<T, Q>(qmul<Q>(f, (Q)e[0])); }
};
template <int CLS> struct RotateOp {
    typedef typename QT<CLS>::T T;
    typedef typename QT<CLS>::Q Q;
    static constexpr int UNIT = 2;
    Q pr, pi;
    __device__ __forceinline__ void operator()(T *e) const
    {
        const Q a = (Q)e[0], b = (Q)e[1];
        // no FMA contraction across the two products: the oracle (and the reference's -O2 build on
        // x86-64) rounds each product before the add
        if constexpr (std::is_same<Q, float>::value) {
            e[0] = __fsub_rn(__fmul_rn(pr, a), __fmul_rn(pi, b));
            e[1] = __fadd_rn(__fmul_rn(pr, b), __fmul_rn(pi, a));
        } else if constexpr (std::is_same<Q, double>::value) {
            e[0] = __dsub_rn(__dmul_rn(pr, a), __dmul_rn(pi, b));
            e[1] = __dadd_rn(__dmul_rn(pr, b), __dmul_rn(pi, a));
        } else {
            e[0] = from_q<T, Q>(qsub<Q>(qmul<Q>(pr, a), qmul<Q>(pi, b)));
            e[1] = from_q<T, Q>(qadd<Q>(qmul<Q>(pr, b), qmul<Q>(pi, a)));
        }
    }
};

// n = number of scalars (a multiple of Op::UNIT); VEC: 16-byte vectors when both pointers allow it
template <typename Op, bool VEC>
__global__ void __launch_bounds__(256) map_kernel(const typename Op::T *__restrict__ in, typename Op::T *__restrict__ out, size_t n, const Op op)
{
    typedef typename Op::T T;
    constexpr int V = VEC ? 16 / (int)sizeof(T) : Op::UNIT;
    const size_t nv = n / V, stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += stride) {
        if constexpr (VEC) {
            uint4 u = __ldcs(reinterpret_cast<const uint4 *>(in) + i);
            T *e = reinterpret_cast<T *>(&u);
#pragma unroll
            for (int k = 0; k < V; k += Op::UNIT) op(e + k);
            __stcs(reinterpret_cast<uint4 *>(out) + i, u);
        } else {
            T e[Op::UNIT];
#pragma unroll
            for (int k = 0; k < Op::UNIT; k++) e[k] = in[i * Op::UNIT + k];
            op(e);
#pragma unroll
            for (int k = 0; k < Op::UNIT; k++) out[i * Op::UNIT + k] = e[k];
        }
    }
    if constexpr (VEC) {   // tail shorter than one vector
        const size_t done = nv * V, left = (n - done) / Op::UNIT;
        const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
        if (i < left) {
            T e[Op::UNIT];
#pragma unroll
            for (int k = 0; k < Op::UNIT; k++) e[k] = in[done + i * Op::UNIT + k];
            op(e);
#pragma unroll
            for (int k = 0; k < Op::UNIT; k++) out[done + i * Op::UNIT + k] = e[k];
        }
    }
}

static int sm_count_of(int device, int &sms)
{
    B200C_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    return B200C_OK;
}

template <typename Op>
static int launch_map(const Op &op, const void *d_in, void *d_out, size_t n_scalars, int device, cudaStream_t stream)
{
    if (n_scalars == 0) return B200C_OK;
    int sms = 0, rc;
    if ((rc = sm_count_of(device, sms))) return rc;
    typedef typename Op::T T;
    const bool vec = ((reinterpret_cast<uintptr_t>(d_in) | reinterpret_cast<uintptr_t>(d_out)) & 15) == 0;
    const size_t items = vec ? n_scalars / (16 / sizeof(T)) + 1 : n_scalars / Op::UNIT;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((items + 255) / 256, (size_t)sms * 16));
    if (vec) map_kernel<Op, true><<<grid, 256, 0, stream>>>(static_cast<const T *>(d_in), static_cast<T *>(d_out), n_scalars, op);
    else map_kernel<Op, false><<<grid, 256, 0, stream>>>(static_cast<const T *>(d_in), static_cast<T *>(d_out), n_scalars, op);
    B200C_CUDA_TRY(cudaGetLastError());
    return B200C_OK;
}

// floatToQ<Q>(x): integer Q -> trunc(ldexp(x, 4 sizeof Q)) narrowed (wrapping) to Q; float Q -> cast
template <typename Q> static Q float_to_q(double x)
{
    if constexpr (std::is_floating_point<Q>::value) return (Q)x;
    else return (Q)(long long)std::ldexp(x, 4 * (int)sizeof(Q));
}

template <int CLS> static int scale_cls(double factor, const void *d_in, void *d_out, size_t n_scalars, int device, cudaStream_t s)
{
    ScaleOp<CLS> op;
    op.f = float_to_q<typename QT<CLS>::Q>(factor);
    return launch_map(op, d_in, d_out, n_scalars, device, s);
}
template <int CLS> static int rotate_cls(double phase, const void *d_in, void *d_out, size_t n_scalars, int device, cudaStream_t s)
{
    RotateOp<CLS> op;
    op.pr = float_to_q<typename QT<CLS>::Q>(std::cos(phase));   // std::polar(1.0, phase), math/Rotate.cpp:74
    op.pi = float_to_q<typename QT<CLS>::Q>(std::sin(phase));
    return launch_map(op, d_in, d_out, n_scalars, device, s);
}

// ------------------------------------------------------------------------------ probe ---
// acc[0] += sum of |x|^2 (RMS) or sum of re (MEAN), acc[1] += sum of im (MEAN).  8- and 16-bit
// integers are summed exactly in 64-bit integers per thread (a square pair is < 2^31), everything
// else in double; 128-bit loads when the window is 16-byte aligned, four in flight per thread.
template <typename T> struct ProbeAcc { typedef double type; };
template <> struct ProbeAcc<int8_t> { typedef long long type; };
template <> struct ProbeAcc<int16_t> { typedef long long type; };

template <typename T, int NC, int MODE, typename A>
__device__ __forceinline__ void probe_accumulate(const T *e, int count, A &a0, A &a1)
{
#pragma unroll
    for (int k = 0; k < count; k += NC) {
        if constexpr (MODE == 1) {
            if constexpr (std::is_same<A, long long>::value) {
                const int re = e[k], im = NC == 2 ? e[k + NC - 1] : 0;
                a0 += (long long)(re * re + im * im);
            } else {
                const double re = (double)e[k], im = NC == 2 ? (double)e[k + NC - 1] : 0.0;
                a0 += re * re + im * im;
            }
        } else {
            a0 += (A)e[k];
            if (NC == 2) a1 += (A)e[k + NC - 1];
        }
    }
}

template <typename T, int NC, int MODE, bool VEC>
__global__ void __launch_bounds__(256) probe_kernel(const T *__restrict__ in, size_t n_elems, double *acc)
{
    typedef typename ProbeAcc<T>::type A;
    A a0 = 0, a1 = 0;
    const size_t n = n_elems * NC, stride = (size_t)gridDim.x * blockDim.x, tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if constexpr (VEC) {
        constexpr int V = 16 / (int)sizeof(T);               // scalars per vector: a whole number of elements
        const size_t nv = n / V;
        const uint4 *vin = reinterpret_cast<const uint4 *>(in);
        for (size_t i = tid; i < nv; i += 4 * stride) {
            uint4 u[4];
#pragma unroll
            for (int j = 0; j < 4; j++) u[j] = i + j * stride < nv ? __ldcs(vin + i + j * stride) : make_uint4(0, 0, 0, 0);
#pragma unroll
            for (int j = 0; j < 4; j++) probe_accumulate<T, NC, MODE, A>(reinterpret_cast<const T *>(&u[j]), V, a0, a1);
        }
        for (size_t i = nv * V + tid * NC; i < n; i += stride * NC) probe_accumulate<T, NC, MODE, A>(in + i, NC, a0, a1);
    } else {
        for (size_t i = tid * NC; i < n; i += stride * NC) probe_accumulate<T, NC, MODE, A>(in + i, NC, a0, a1);
    }
    double d0 = (double)a0, d1 = (double)a1;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        d0 += __shfl_xor_sync(0xffffffffu, d0, o);
        d1 += __shfl_xor_sync(0xffffffffu, d1, o);
    }
    __shared__ double s0[8], s1[8];
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { s0[w] = d0; s1[w] = d1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t0 = 0.0, t1 = 0.0;
        for (int k = 0; k < 8; k++) { t0 += s0[k]; t1 += s1[k]; }
        atomicAdd(acc, t0);
        if (MODE == 2 && NC == 2) atomicAdd(acc + 1, t1);
    }
}

template <typename T, int NC, int MODE>
static void launch_probe(const T *in, size_t n, double *d_acc, int grid, cudaStream_t s)
{
    if ((reinterpret_cast<uintptr_t>(in) & 15) == 0) probe_kernel<T, NC, MODE, true><<<grid, 256, 0, s>>>(in, n, d_acc);
    else probe_kernel<T, NC, MODE, false><<<grid, 256, 0, s>>>(in, n, d_acc);
}

template <typename T> static int probe_t(bool cx, int mode, const void *d_in, size_t n, double *value, int device, cudaStream_t s)
{
    const int nc = cx ? 2 : 1;
    if (mode == 0) {   // VALUE: the last element of the window
        T last[2] = {0, 0};
        B200C_CUDA_TRY(cudaMemcpyAsync(last, static_cast<const T *>(d_in) + (n - 1) * nc, sizeof(T) * nc, cudaMemcpyDeviceToHost, s));
        B200C_CUDA_TRY(cudaStreamSynchronize(s));
        value[0] = (double)last[0]; value[1] = cx ? (double)last[1] : 0.0;
        return B200C_OK;
    }
    int sms = 0, rc;
    if ((rc = sm_count_of(device, sms))) return rc;
    // per-thread, per-device scratch (two doubles on the device, two pinned on the host), allocated once:
    // the actor calling work() is the only user, and a probe window is small, so call overhead matters
    static thread_local double *scratch_d[16] = {nullptr}, *scratch_h[16] = {nullptr};
    if (device >= 16) { set_error("b200c_probe: device index above 15"); return B200C_ERR_UNSUPPORTED; }
    if (!scratch_d[device]) {
        B200C_CUDA_TRY(cudaMalloc(&scratch_d[device], 2 * sizeof(double)));
        B200C_CUDA_TRY(cudaMallocHost(&scratch_h[device], 2 * sizeof(double)));
    }
    double *d_acc = scratch_d[device], *h = scratch_h[device];
    cudaError_t e = cudaMemsetAsync(d_acc, 0, 2 * sizeof(double), s);
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((n * nc * sizeof(T) / 16 + 1023) / 1024, (size_t)sms * 8));
    const T *in = static_cast<const T *>(d_in);
    if (e == cudaSuccess) {
        if (mode == 1) { if (cx) launch_probe<T, 2, 1>(in, n, d_acc, grid, s); else launch_probe<T, 1, 1>(in, n, d_acc, grid, s); }
        else { if (cx) launch_probe<T, 2, 2>(in, n, d_acc, grid, s); else launch_probe<T, 1, 2>(in, n, d_acc, grid, s); }
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(h, d_acc, 2 * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) { set_error("b200c_probe failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return B200C_ERR_CUDA; }
    if (mode == 1) { value[0] = std::sqrt(h[0] / (double)n); value[1] = 0.0; }
    else { value[0] = h[0] / (double)n; value[1] = h[1] / (double)n; }
    return B200C_OK;
}

} // namespace b200c

using namespace b200c;

static int check_device(int device)
{
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        set_error("no usable CUDA device: the B200 path has no CPU fallback");
        return B200C_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range", device); return B200C_ERR_INVALID; }
    return B200C_OK;
}

extern "C" {

int b200c_scale(int dtype, double factor, const void *d_in, void *d_out, size_t elems, int device, void *stream)
{
    if (!dtype_valid(dtype)) { set_error("scaleFactory(): unsupported type"); return B200C_ERR_UNSUPPORTED; }
    if (elems == 0) return B200C_OK;
    if (!d_in || !d_out) { set_error("b200c_scale: null device buffer"); return B200C_ERR_INVALID; }
    int rc;
    if ((rc = check_device(device))) return rc;
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    const size_t n = elems * (dtype_is_complex(dtype) ? 2 : 1);
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype >> 1) {
    case 0: return scale_cls<0>(factor, d_in, d_out, n, device, s);
    case 1: return scale_cls<1>(factor, d_in, d_out, n, device, s);
    case 2: return scale_cls<2>(factor, d_in, d_out, n, device, s);
    case 3: return scale_cls<3>(factor, d_in, d_out, n, device, s);
    case 4: return scale_cls<4>(factor, d_in, d_out, n, device, s);
    default: return scale_cls<5>(factor, d_in, d_out, n, device, s);
    }
}

int b200c_rotate(int dtype, double phase, const void *d_in, void *d_out, size_t elems, int device, void *stream)
{
    if (!dtype_valid(dtype) || !dtype_is_complex(dtype)) { set_error("rotateFactory(): unsupported type"); return B200C_ERR_UNSUPPORTED; }
    if (elems == 0) return B200C_OK;
    if (!d_in || !d_out) { set_error("b200c_rotate: null device buffer"); return B200C_ERR_INVALID; }
    int rc;
    if ((rc = check_device(device))) return rc;
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    const size_t n = elems * 2;
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype >> 1) {
    case 0: return rotate_cls<0>(phase, d_in, d_out, n, device, s);
    case 1: return rotate_cls<1>(phase, d_in, d_out, n, device, s);
    case 2: return rotate_cls<2>(phase, d_in, d_out, n, device, s);
    case 3: return rotate_cls<3>(phase, d_in, d_out, n, device, s);
    case 4: return rotate_cls<4>(phase, d_in, d_out, n, device, s);
    default: return rotate_cls<5>(phase, d_in, d_out, n, device, s);
    }
}

int b200c_probe(int dtype, int mode, const void *d_in, size_t elems, double *value, int device, void *stream)
{
    if (!dtype_valid(dtype)) { set_error("signalProbeFactory(): unsupported type"); return B200C_ERR_UNSUPPORTED; }
    if (mode < 0 || mode > 2 || !value) { set_error("b200c_probe: bad mode or null result"); return B200C_ERR_INVALID; }
    if (elems == 0 || !d_in) { set_error("b200c_probe: empty window"); return B200C_ERR_INVALID; }
    int rc;
    if ((rc = check_device(device))) return rc;
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    const bool cx = dtype_is_complex(dtype);
    cudaStream_t s = (cudaStream_t)stream;
    switch (dtype >> 1) {
    case 0: return probe_t<float>(cx, mode, d_in, elems, value, device, s);
    case 1: return probe_t<double>(cx, mode, d_in, elems, value, device, s);
    case 2: return probe_t<int8_t>(cx, mode, d_in, elems, value, device, s);
    case 3: return probe_t<int16_t>(cx, mode, d_in, elems, value, device, s);
    case 4: return probe_t<int32_t>(cx, mode, d_in, elems, value, device, s);
    default: return probe_t<int64_t>(cx, mode, d_in, elems, value, device, s);
    }
}

} // extern "C"
