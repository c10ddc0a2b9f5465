// Same-process exercise of the b200c_peer_* / b200c_halo_exchange entry points on two GPUs, step by step (prints
// before every call so that a crash names its call).  g++ -O1 -o tools/probe_peer_halo tools/probe_peer_halo.cpp -ldl
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <vector>
#include "../include/b200comms.h"

#define LOAD(name) auto p_##name = (decltype(&name))dlsym(lib, #name); if (!p_##name) { printf("missing %s\n", #name); return 2; }
#define STEP(expr) do { printf("-> %s\n", #expr); fflush(stdout); int rc__ = (expr); printf("   rc=%d %s\n", rc__, rc__ ? p_b200c_last_error() : ""); fflush(stdout); } while (0)

int main(int argc, char **argv)
{
    void *lib = dlopen(argc > 1 ? argv[1] : "pothoscomms_b200/libb200comms.so", RTLD_NOW);
    if (!lib) { printf("dlopen: %s\n", dlerror()); return 2; }
    LOAD(b200c_last_error) LOAD(b200c_device_count) LOAD(b200c_dev_alloc) LOAD(b200c_dev_free) LOAD(b200c_copy_h2d) LOAD(b200c_copy_d2h)
    LOAD(b200c_stream_sync) LOAD(b200c_peer_export) LOAD(b200c_peer_open) LOAD(b200c_peer_close) LOAD(b200c_peer_event_create)
    LOAD(b200c_peer_event_open) LOAD(b200c_peer_event_record) LOAD(b200c_peer_event_wait) LOAD(b200c_peer_event_destroy)
    LOAD(b200c_halo_exchange)
    int n = 0;
    p_b200c_device_count(&n);
    printf("devices: %d\n", n);
    if (n < 2) return 0;
    const size_t bytes = 2040;
    void *a = nullptr, *b = nullptr;
    STEP(p_b200c_dev_alloc(&a, 1 << 20, 0));
    STEP(p_b200c_dev_alloc(&b, 1 << 20, 1));
    std::vector<unsigned char> h(bytes), g(bytes, 0);
    for (size_t i = 0; i < bytes; i++) h[i] = (unsigned char)(i * 7 + 3);
    STEP(p_b200c_copy_h2d((char *)a + 4096, h.data(), bytes, 0, nullptr));
    STEP(p_b200c_stream_sync(0, nullptr));
    b200c_peer_mem pm;
    b200c_peer_event pe;
    void *ev = nullptr, *ev_open = nullptr, *peer = nullptr, *mapping = nullptr;
    STEP(p_b200c_peer_export((char *)a + 4096, bytes, 0, &pm));
    printf("   offset=%llu bytes=%llu pid=%lld dev=%d local=%llx\n", (unsigned long long)pm.offset, (unsigned long long)pm.bytes, (long long)pm.pid, pm.device, (unsigned long long)pm.local_ptr);
    STEP(p_b200c_peer_event_create(&ev, 0, &pe));
    printf("   ev=%p local_event=%llx pid=%lld\n", ev, (unsigned long long)pe.local_event, (long long)pe.pid);
    STEP(p_b200c_peer_open(&pm, 1, &peer, &mapping));
    printf("   peer=%p mapping=%p\n", peer, mapping);
    STEP(p_b200c_peer_event_open(&pe, 1, &ev_open));
    printf("   ev_open=%p\n", ev_open);
    STEP(p_b200c_peer_event_record(ev, 0, nullptr));
    STEP(p_b200c_peer_event_wait(ev_open, 1, nullptr));
    STEP(p_b200c_halo_exchange(b, peer, bytes, 1, nullptr));
    STEP(p_b200c_stream_sync(1, nullptr));
    STEP(p_b200c_copy_d2h(g.data(), b, bytes, 1, nullptr));
    STEP(p_b200c_stream_sync(1, nullptr));
    printf("halo bytes %s\n", std::memcmp(g.data(), h.data(), bytes) == 0 ? "MATCH" : "DIFFER");
    STEP(p_b200c_peer_close(mapping));
    STEP(p_b200c_peer_event_destroy(ev_open, 1));
    STEP(p_b200c_peer_event_destroy(ev, 0));
    p_b200c_dev_free(a, 0); p_b200c_dev_free(b, 1);
    return 0;
}
