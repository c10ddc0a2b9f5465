// Multi-GPU plumbing of the one exchange the FIR path has: the K-1 history samples at a
// segment boundary (filter/FIRFilter.cpp:281,298).  One process per GPU: the left neighbour's segment is
// opened through CUDA IPC (or used directly when both GPUs belong to this process) and the halo is PULLED
// over NVLink by a copy enqueued on the consumer's compute stream -- no collective, no extra kernel, no
// host round trip; ordering against the producer is an interprocess event.  See include/b200comms.h.
#include <cuda.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>

#include "common.hpp"

using namespace b200c;

static_assert(sizeof(cudaIpcMemHandle_t) == 64 && sizeof(cudaIpcEventHandle_t) == 64, "IPC handles are 64 bytes");

struct PeerMapping {   // what b200c_peer_open hands out, so that close knows what to undo
    void *base;        // cudaIpcOpenMemHandle result (nullptr for a same-process pointer)
    int device;
};

extern "C" {

int b200c_peer_export(const void *d_ptr, size_t bytes, int device, b200c_peer_mem *out)
{
    if (!d_ptr || !out) { set_error("b200c_peer_export: null argument"); return B200C_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    std::memset(out, 0, sizeof(*out));
    CUdeviceptr base = 0;
    size_t span = 0;
    // fetched through the runtime: libb200comms.so has no link-time dependency on libcuda.so.1
    typedef CUresult (*fn_range)(CUdeviceptr *, size_t *, CUdeviceptr);
    static fn_range p_range = nullptr;
    if (!p_range) {
        cudaDriverEntryPointQueryResult qr;
        void *fp = nullptr;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &qr) != cudaSuccess || !fp) {
            (void)cudaGetLastError();
            set_error("CUDA driver entry point cuMemGetAddressRange unavailable");
            return B200C_ERR_CUDA;
        }
        p_range = (fn_range)fp;
    }
    B200C_CUDA_TRY(cudaFree(0));
    if (p_range(&base, &span, (CUdeviceptr)d_ptr) != CUDA_SUCCESS) {
        set_error("b200c_peer_export: %p is not a device allocation", d_ptr);
        return B200C_ERR_INVALID;
    }
    const size_t off = (size_t)((CUdeviceptr)d_ptr - base);
    if (off + bytes > span) { set_error("b200c_peer_export: range exceeds its allocation"); return B200C_ERR_INVALID; }
    cudaIpcMemHandle_t mh;
    B200C_CUDA_TRY(cudaIpcGetMemHandle(&mh, (void *)base));
    std::memcpy(out->ipc, &mh, 64);
    out->offset = off;
    out->bytes = bytes;
    out->device = device;
    out->pid = (int64_t)getpid();
    out->local_ptr = (uint64_t)(uintptr_t)d_ptr;
    return B200C_OK;
}

int b200c_peer_open(const b200c_peer_mem *m, int device, void **d_peer_ptr, void **mapping)
{
    if (!m || !d_peer_ptr || !mapping) { set_error("b200c_peer_open: null argument"); return B200C_ERR_INVALID; }
    *d_peer_ptr = nullptr; *mapping = nullptr;
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    PeerMapping *pm = new PeerMapping{nullptr, device};
    if (m->pid == (int64_t)getpid()) {
        // both GPUs in this process: unified addressing makes the pointer usable as it is; direct access
        // (instead of a staged copy) needs peer access enabled once
        if (m->device != device) {
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, device, m->device) == cudaSuccess && can) {
                const cudaError_t e = cudaDeviceEnablePeerAccess(m->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); }
                (void)cudaGetLastError();
            }
        }
        *d_peer_ptr = (void *)(uintptr_t)m->local_ptr;
    } else {
        cudaIpcMemHandle_t mh;
        std::memcpy(&mh, m->ipc, 64);
        void *base = nullptr;
        const cudaError_t e = cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            set_error("cudaIpcOpenMemHandle failed: %s", cudaGetErrorString(e));
            (void)cudaGetLastError();
            delete pm;
            return B200C_ERR_CUDA;
        }
        pm->base = base;
        *d_peer_ptr = static_cast<char *>(base) + m->offset;
    }
    *mapping = pm;
    return B200C_OK;
}

int b200c_peer_close(void *mapping)
{
    PeerMapping *pm = static_cast<PeerMapping *>(mapping);
    if (!pm) return B200C_OK;
    int rc = B200C_OK;
    if (pm->base) {
        DeviceGuard g(pm->device);
        if (cudaIpcCloseMemHandle(pm->base) != cudaSuccess) { (void)cudaGetLastError(); rc = B200C_ERR_CUDA; set_error("cudaIpcCloseMemHandle failed"); }
    }
    delete pm;
    return rc;
}

// An event handle as this library hands it out.  The interprocess event serves neighbours in OTHER processes; a
// neighbour in the SAME process waits on a plain event recorded next to it (cudaStreamWaitEvent on an interprocess
// event from another device's stream of the same process faults inside the driver -- tools/probe_peer_halo.cpp).
struct PeerEvent {
    cudaEvent_t ipc;      // interprocess event (owner) / opened interprocess event (other process) / nullptr (same-process alias)
    cudaEvent_t plain;    // same-process event (owner and same-process aliases) / nullptr
    bool owner;
};

int b200c_peer_event_create(void **event, int device, b200c_peer_event *out)
{
    if (!event || !out) { set_error("b200c_peer_event_create: null argument"); return B200C_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    cudaEvent_t ev, plain;
    B200C_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming | cudaEventInterprocess));
    cudaIpcEventHandle_t eh;
    cudaError_t e = cudaIpcGetEventHandle(&eh, ev);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&plain, cudaEventDisableTiming);
    if (e != cudaSuccess) { cudaEventDestroy(ev); set_error("peer event creation failed: %s", cudaGetErrorString(e)); (void)cudaGetLastError(); return B200C_ERR_CUDA; }
    PeerEvent *pe = new PeerEvent{ev, plain, true};
    std::memset(out, 0, sizeof(*out));
    std::memcpy(out->ipc, &eh, 64);
    out->pid = (int64_t)getpid();
    out->local_event = (uint64_t)(uintptr_t)pe;
    *event = pe;
    return B200C_OK;
}

int b200c_peer_event_open(const b200c_peer_event *e, int device, void **event)
{
    if (!e || !event) { set_error("b200c_peer_event_open: null argument"); return B200C_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    if (e->pid == (int64_t)getpid()) {
        const PeerEvent *own = reinterpret_cast<const PeerEvent *>((uintptr_t)e->local_event);
        *event = new PeerEvent{nullptr, own->plain, false};
        return B200C_OK;
    }
    cudaIpcEventHandle_t eh;
    std::memcpy(&eh, e->ipc, 64);
    cudaEvent_t ev;
    B200C_CUDA_TRY(cudaIpcOpenEventHandle(&ev, eh));
    *event = new PeerEvent{ev, nullptr, false};
    return B200C_OK;
}

int b200c_peer_event_record(void *event, int device, void *stream)
{
    PeerEvent *pe = static_cast<PeerEvent *>(event);
    if (!pe || !pe->owner) { set_error("b200c_peer_event_record: only the creating side records"); return B200C_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    B200C_CUDA_TRY(cudaEventRecord(pe->ipc, (cudaStream_t)stream));
    B200C_CUDA_TRY(cudaEventRecord(pe->plain, (cudaStream_t)stream));
    return B200C_OK;
}

int b200c_peer_event_wait(void *event, int device, void *stream)
{
    PeerEvent *pe = static_cast<PeerEvent *>(event);
    if (!pe) { set_error("b200c_peer_event_wait: null event"); return B200C_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    B200C_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, pe->plain ? pe->plain : pe->ipc, 0));
    return B200C_OK;
}

int b200c_peer_event_destroy(void *event, int device)
{
    PeerEvent *pe = static_cast<PeerEvent *>(event);
    if (!pe) return B200C_OK;
    DeviceGuard g(device);
    if (pe->owner) { cudaEventDestroy(pe->ipc); cudaEventDestroy(pe->plain); }
    else if (pe->ipc) cudaEventDestroy(pe->ipc);
    (void)cudaGetLastError();
    delete pe;
    return B200C_OK;
}

// A one-CTA copy kernel: the halo is at most a few KB, so what matters is latency -- a kernel that loads the
// neighbour's tail straight over NVLink (peer-mapped memory) sits ~2 us in front of the FIR launch where a
// copy-engine cudaMemcpyAsync between devices costs ~10 us (measured on 2 GPUs: 12 us per step with the memcpy).
__global__ void __launch_bounds__(256) halo_copy_kernel(unsigned char *__restrict__ dst, const unsigned char *__restrict__ src, size_t bytes)
{
    const bool al = ((reinterpret_cast<unsigned long long>(dst) | reinterpret_cast<unsigned long long>(src)) & 15) == 0;
    size_t done = 0;
    if (al) {
        const size_t nv = bytes >> 4;
        for (size_t i = threadIdx.x; i < nv; i += blockDim.x)
            reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
        done = nv << 4;
    }
    for (size_t i = done + threadIdx.x; i < bytes; i += blockDim.x) dst[i] = src[i];
}

// can `device` dereference memory owned by `owner` inside a kernel?  (answer cached per pair)
static bool direct_access(int device, int owner)
{
    if (device == owner) return true;
    static thread_local signed char cache[16][16] = {};   // 0 unknown, 1 yes, -1 no
    if (device < 16 && owner >= 0 && owner < 16 && cache[device][owner]) return cache[device][owner] > 0;
    int can = 0;
    bool ok = cudaDeviceCanAccessPeer(&can, device, owner) == cudaSuccess && can;
    if (ok) {
        const cudaError_t e = cudaDeviceEnablePeerAccess(owner, 0);     // no-op when b200c_peer_open already did it
        ok = e == cudaSuccess || e == cudaErrorPeerAccessAlreadyEnabled;
    }
    (void)cudaGetLastError();
    if (device < 16 && owner >= 0 && owner < 16) cache[device][owner] = ok ? 1 : -1;
    return ok;
}

int b200c_halo_exchange(void *d_halo_dst, const void *d_peer_tail, size_t bytes, int device, void *stream)
{
    if (bytes == 0) return B200C_OK;
    if (!d_halo_dst || !d_peer_tail) { set_error("b200c_halo_exchange: null buffer"); return B200C_ERR_INVALID; }
    DeviceGuard g(device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B200C_ERR_CUDA; }
    // whose memory is the tail?  (one attribute query per distinct pointer: segments are long-lived)
    static thread_local const void *last_ptr = nullptr;
    static thread_local int last_owner = -1;
    if (d_peer_tail != last_ptr) {
        cudaPointerAttributes at;
        last_owner = cudaPointerGetAttributes(&at, d_peer_tail) == cudaSuccess && at.type == cudaMemoryTypeDevice ? at.device : -1;
        (void)cudaGetLastError();
        last_ptr = d_peer_tail;
    }
    static const bool force_memcpy = [] { const char *e = std::getenv("B200C_HALO_MEMCPY"); return e && std::atoi(e) != 0; }();
    if (!force_memcpy && last_owner >= 0 && direct_access(device, last_owner)) {
        halo_copy_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(static_cast<unsigned char *>(d_halo_dst),
                                                               static_cast<const unsigned char *>(d_peer_tail), bytes);
        B200C_CUDA_TRY(cudaGetLastError());
        return B200C_OK;
    }
    // no direct access between the two devices: the runtime stages the copy (unified addressing)
    B200C_CUDA_TRY(cudaMemcpyAsync(d_halo_dst, d_peer_tail, bytes, cudaMemcpyDefault, (cudaStream_t)stream));
    return B200C_OK;
}

} // extern "C"
