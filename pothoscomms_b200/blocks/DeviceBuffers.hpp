// Device-resident Pothos::BufferManager implementations over the C-ABI (include/b200comms.h).
// They replace, for the "b200c_hbm" buffer domain, the host managers the reference asks for:
//   Pothos::BufferManager::make("circular")        filter/FIRFilter.cpp:196-199
//   Pothos::BufferManager::make("generic", args)   fft/FFT.cpp:54-59
// so that samples stay in HBM between work() calls (no per-call host<->device copy).
#pragma once
#include <Pothos/Framework.hpp>

#include "../../include/b200comms.h"

namespace b200c_blocks {

static const char *const kHbmDomain = "b200c_hbm";

inline void throwOnError(int rc, const std::string &where)
{
    if (rc == B200C_OK) return;
    const std::string text = b200c_last_error();
    if (rc == B200C_ERR_INVALID || rc == B200C_ERR_UNSUPPORTED) throw Pothos::InvalidArgumentException(where, text);
    throw Pothos::Exception(where, text);
}

// The device blocks hand out HBM buffers, so they can only share them with a neighbour whose port is in the
// same domain.  A host-memory neighbour (domain "": /blocks/feeder_source, /blocks/collector_sink, any CPU
// block) would dereference device pointers: refused with Pothos::PortDomainError, and the topology connects
// it through /b200c/host_to_hbm or /b200c/hbm_to_host (blocks/Bridge.cpp) instead.
inline void requireHbmPeer(const std::string &who, const std::string &domain)
{
    if (domain == kHbmDomain) return;
    if (domain.empty())
        throw Pothos::PortDomainError(who, "host-memory neighbour: connect it through /b200c/host_to_hbm or /b200c/hbm_to_host");
    throw Pothos::PortDomainError(who, "cannot share buffers with domain " + domain);
}

// "circular": one HBM ring mapped twice back to back (CUDA VMM), so the readable window
// (K-1 history + new samples) and the writable window are each always contiguous.
class DeviceCircularBufferManager : public Pothos::BufferManager {
public:
    explicit DeviceCircularBufferManager(int device) : _device(device) {}
    ~DeviceCircularBufferManager() override { b200c_ring_destroy(_ring); }
    void init(const Pothos::BufferManagerArgs &args) override
    {
        throwOnError(b200c_ring_create(&_ring, args.bufferSize * args.numBuffers, _device), "DeviceCircularBufferManager::init()");
        _base = reinterpret_cast<size_t>(b200c_ring_base(_ring));
        _size = b200c_ring_bytes(_ring);
        _rd = _filled = 0;
        update();
    }
    bool empty() const override { return _filled == _size; }
    const Pothos::BufferChunk &front() const override { return _front; }
    void pop(size_t numBytes) override { _filled += numBytes; update(); }                     // producer wrote numBytes
    void push(size_t numBytes) override { _rd = (_rd + numBytes) % _size; _filled -= numBytes; update(); } // consumer released
    std::string domain() const override { return kHbmDomain; }
    // the consumer's contiguous view of everything not yet released
    Pothos::BufferChunk readable() const { return Pothos::BufferChunk(_base + _rd, _filled); }
    size_t capacity() const { return _size; }
    size_t base() const { return _base; }

private:
    void update() { _front = Pothos::BufferChunk(_base + (_rd + _filled) % _size, _size - _filled); }
    int _device;
    b200c_ring *_ring = nullptr;
    size_t _base = 0, _size = 0, _rd = 0, _filled = 0;
    Pothos::BufferChunk _front;
};

// "generic": a pool of equally sized HBM slabs handed out in order.
class DeviceSlabBufferManager : public Pothos::BufferManager {
public:
    explicit DeviceSlabBufferManager(int device) : _device(device) {}
    ~DeviceSlabBufferManager() override
    {
        for (void *p : _slabs) b200c_dev_free(p, _device);
    }
    void init(const Pothos::BufferManagerArgs &args) override
    {
        _bytes = args.bufferSize;
        for (size_t i = 0; i < args.numBuffers; i++) {
            void *p = nullptr;
            throwOnError(b200c_dev_alloc(&p, _bytes, _device), "DeviceSlabBufferManager::init()");
            _slabs.push_back(p);
        }
        _head = 0; _out = 0;
        update();
    }
    bool empty() const override { return _out == _slabs.size(); }
    const Pothos::BufferChunk &front() const override { return _front; }
    void pop(size_t) override { _out++; _head = (_head + 1) % _slabs.size(); update(); }
    void push(size_t) override { if (_out) _out--; update(); }
    std::string domain() const override { return kHbmDomain; }

private:
    void update() { _front = empty() ? Pothos::BufferChunk() : Pothos::BufferChunk(reinterpret_cast<size_t>(_slabs[_head]), _bytes); }
    int _device;
    size_t _bytes = 0, _head = 0, _out = 0;
    std::vector<void *> _slabs;
    Pothos::BufferChunk _front;
};

inline int dtypeCode(const Pothos::DType &dt)
{
    static const std::pair<const char *, int> table[] = {
        {"float32", B200C_F32}, {"complex_float32", B200C_CF32}, {"float64", B200C_F64}, {"complex_float64", B200C_CF64},
        {"int8", B200C_I8}, {"complex_int8", B200C_CI8}, {"int16", B200C_I16}, {"complex_int16", B200C_CI16},
        {"int32", B200C_I32}, {"complex_int32", B200C_CI32}, {"int64", B200C_I64}, {"complex_int64", B200C_CI64}};
    for (const auto &e : table) if (dt.name() == e.first) return e.second;
    return -1;
}

} // namespace b200c_blocks
