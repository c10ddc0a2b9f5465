// HBM-resident stream sources: the work() loops of
// /comms/waveform_source and /comms/noise_source,
//   out[i] = table[(index + i * step) & mask]        waveform/WaveformSource.cpp:98-108
//   out[i] = table[(index + i) % 4096]               waveform/NoiseSource.cpp:108-117 (fast mode)
// as one table-walk kernel.  The tables themselves (4096 ... 2^20 entries of the stream's element
// type) are filled on the host by the block layer exactly as updateTable() does and live in HBM; a
// table of up to 64 KB is staged in shared memory once per CTA, larger ones are gathered through L1/L2.
// The stream is write-only: algorithmic bytes = sizeof(element) per element, stored as 128-bit
// streaming vectors, grid a multiple of the SM count.
#include <algorithm>

#include "common.hpp"

namespace b200c {

constexpr int kSrcThreads = 256;
constexpr size_t kSrcSmemTable = 64 * 1024;

template <typename E> struct SrcVec { static constexpr int V = 16 / (int)sizeof(E); };

// E = an unsigned carrier of the element's size (1, 2, 4, 8, 16 bytes); the table walk only moves bits.
// W > 1: an element is W carriers (buffers aligned to the scalar but not to the complex element).
template <typename E, bool SMEM, bool VEC, int W = 1>
__global__ void __launch_bounds__(kSrcThreads) table_source_kernel(const E *__restrict__ table, unsigned long long mask,
                                                                   unsigned long long index, unsigned long long step,
                                                                   E *__restrict__ out, size_t n)
{
    extern __shared__ uint4 src_smem[];
    const E *tab = table;
    if constexpr (SMEM) {
        E *st = reinterpret_cast<E *>(src_smem);
        const size_t bytes = (size_t)(mask + 1) * sizeof(E) * W;
        if (bytes % 16 == 0 && (reinterpret_cast<uintptr_t>(table) & 15) == 0) {
            const uint4 *g = reinterpret_cast<const uint4 *>(table);
            for (size_t i = threadIdx.x; i < bytes / 16; i += kSrcThreads) src_smem[i] = __ldg(g + i);
        } else {
            for (size_t i = threadIdx.x; i < (mask + 1) * W; i += kSrcThreads) st[i] = table[i];
        }
        __syncthreads();
        tab = st;
    }
    constexpr int V = SrcVec<E>::V;
    const size_t stride = (size_t)gridDim.x * kSrcThreads, tid = (size_t)blockIdx.x * kSrcThreads + threadIdx.x;
    if constexpr (VEC && V > 1 && W == 1) {
        const size_t nv = n / V;
        for (size_t i = tid; i < nv; i += stride) {
            uint4 u;
            E *e = reinterpret_cast<E *>(&u);
            unsigned long long at = index + (unsigned long long)(i * V) * step;
#pragma unroll
            for (int k = 0; k < V; k++, at += step) e[k] = tab[at & mask];
            __stcs(reinterpret_cast<uint4 *>(out) + i, u);
        }
        for (size_t i = nv * V + tid; i < n; i += stride) out[i] = tab[(index + (unsigned long long)i * step) & mask];
    } else {
        for (size_t i = tid; i < n; i += stride) {
            const size_t at = (size_t)((index + (unsigned long long)i * step) & mask);
#pragma unroll
            for (int w = 0; w < W; w++) out[i * W + w] = tab[at * W + w];
        }
    }
}

template <typename E, int W = 1>
static int launch_source(const void *d_table, size_t table_elems, unsigned long long index, unsigned long long step, void *d_out,
                         size_t n, int sms, cudaStream_t s)
{
    const bool smem = table_elems * sizeof(E) * W <= kSrcSmemTable;
    const bool vec = W == 1 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_table) & 15) == 0;
    const size_t items = vec ? n / SrcVec<E>::V + 1 : n;
    const int grid = (int)std::max<size_t>(1, std::min<size_t>((items + kSrcThreads - 1) / kSrcThreads, (size_t)sms * 8));
    const size_t sh = smem ? std::max<size_t>(16, table_elems * sizeof(E) * W) : 0;
    const E *t = static_cast<const E *>(d_table);
    E *o = static_cast<E *>(d_out);
    const unsigned long long mask = table_elems - 1;
    if (smem && sh > 48 * 1024) {
        if (vec) B200C_CUDA_TRY(cudaFuncSetAttribute(table_source_kernel<E, true, true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSrcSmemTable));
        else B200C_CUDA_TRY(cudaFuncSetAttribute(table_source_kernel<E, true, false, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSrcSmemTable));
    }
    if (smem) {
        if (vec) table_source_kernel<E, true, true, 1><<<grid, kSrcThreads, sh, s>>>(t, mask, index, step, o, n);
        else table_source_kernel<E, true, false, W><<<grid, kSrcThreads, sh, s>>>(t, mask, index, step, o, n);
    } else {
        if (vec) table_source_kernel<E, false, true, 1><<<grid, kSrcThreads, 0, s>>>(t, mask, index, step, o, n);
        else table_source_kernel<E, false, false, W><<<grid, kSrcThreads, 0, s>>>(t, mask, index, step, o, n);
    }
    B200C_CUDA_TRY(cudaGetLastError());
    return B200C_OK;
}

} // namespace b200c

using namespace b200c;

extern "C" int b200c_table_source(int dtype, const void *d_table, size_t table_elems, uint64_t index, uint64_t step, void *d_out,
                                  size_t elems, int device, void *stream)
{
    if (!dtype_valid(dtype)) { set_error("waveformSourceFactory(): unsupported type"); return B200C_ERR_UNSUPPORTED; }
    if (table_elems == 0 || (table_elems & (table_elems - 1)) != 0) {
        set_error("b200c_table_source: table of %zu entries (the index mask assumes a power of two)", table_elems);
        return B200C_ERR_INVALID;
    }
    if (elems == 0) return B200C_OK;
    if (!d_table || !d_out) { set_error("b200c_table_source: null device buffer"); return B200C_ERR_INVALID; }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        (void)cudaGetLastError();
        set_error("no usable CUDA device: the B200 path has no CPU fallback");
        return B200C_ERR_CUDA;
    }
    if (device < 0 || device >= count) { set_error("device %d out of range", device); return B200C_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    int sms = 0;
    B200C_CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    cudaStream_t s = (cudaStream_t)stream;
    // complex elements are only guaranteed the alignment of their scalar (std::complex<T>): move them as
    // two scalar carriers when a buffer sits between element boundaries
    const size_t esz = dtype_bytes(dtype);
    const bool whole = ((reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_table)) & (esz - 1)) == 0;
    switch (esz) {
    case 1: return launch_source<uint8_t>(d_table, table_elems, index, step, d_out, elems, sms, s);
    case 2: return whole ? launch_source<uint16_t>(d_table, table_elems, index, step, d_out, elems, sms, s)
                         : launch_source<uint8_t, 2>(d_table, table_elems, index, step, d_out, elems, sms, s);
    case 4: return whole ? launch_source<uint32_t>(d_table, table_elems, index, step, d_out, elems, sms, s)
                         : launch_source<uint16_t, 2>(d_table, table_elems, index, step, d_out, elems, sms, s);
    case 8: return whole ? launch_source<uint2>(d_table, table_elems, index, step, d_out, elems, sms, s)
                         : launch_source<uint32_t, 2>(d_table, table_elems, index, step, d_out, elems, sms, s);
    default: return whole ? launch_source<uint4>(d_table, table_elems, index, step, d_out, elems, sms, s)
                          : launch_source<uint2, 2>(d_table, table_elems, index, step, d_out, elems, sms, s);
    }
}
