// /b200c/host_to_hbm and /b200c/hbm_to_host -- the two copy blocks that join a host-memory neighbour
// (/blocks/feeder_source, /blocks/collector_sink, any CPU block) to the device blocks of this module.
//
// The reference's blocks share host buffers with whatever is upstream and downstream
// (filter/FIRFilter.cpp:196-199, fft/FFT.cpp:54-59); the B200 blocks keep their streams in HBM (buffer
// domain "b200c_hbm") and refuse a host-domain peer with Pothos::PortDomainError, because handing HBM
// pointers to a CPU block would make it dereference device memory.  A topology that mixes the two,
// like the reference's own test (filter/TestFIRFilter.cpp:49-51), becomes
//     feeder_source -> /b200c/host_to_hbm -> /comms/fir_filter -> /b200c/hbm_to_host -> collector_sink
// One copy per buffer at each edge of the device section, none inside it.
#include <Pothos/Framework.hpp>

#include <algorithm>
#include <string>

#include "DeviceBuffers.hpp"

using b200c_blocks::kHbmDomain;
using b200c_blocks::throwOnError;

template <bool ToDevice>
class HbmBridge : public Pothos::Block
{
public:
    HbmBridge(const Pothos::DType &dtype, const int device) : _device(device), _bytes(dtype.size())
    {
        this->setupInput(0, dtype, ToDevice ? "" : kHbmDomain);
        this->setupOutput(0, dtype, ToDevice ? kHbmDomain : "");
    }

    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &domain)
    {
        if (ToDevice)
        {
            //host side: the framework's default manager (or the upstream block's) is fine
            if (not domain.empty()) throw Pothos::PortDomainError("HostToHbm::getInputBufferManager()", "expects a host-memory upstream, got " + domain);
            return Pothos::BufferManager::Sptr();
        }
        b200c_blocks::requireHbmPeer("HbmToHost::getInputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceCircularBufferManager(_device));
    }

    Pothos::BufferManager::Sptr getOutputBufferManager(const std::string &, const std::string &domain)
    {
        if (ToDevice)
        {
            b200c_blocks::requireHbmPeer("HostToHbm::getOutputBufferManager()", domain);
            return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceSlabBufferManager(_device));
        }
        if (not domain.empty()) throw Pothos::PortDomainError("HbmToHost::getOutputBufferManager()", "expects a host-memory downstream, got " + domain);
        return Pothos::BufferManager::Sptr();
    }

    void work(void)
    {
        auto inPort = this->input(0);
        auto outPort = this->output(0);
        const size_t n = std::min(inPort->elements(), outPort->elements());
        if (n == 0) return;
        if (ToDevice)
            throwOnError(b200c_copy_h2d(outPort->buffer().template as<void *>(), inPort->buffer().template as<const void *>(), n*_bytes, _device, nullptr), "HostToHbm::work()");
        else
            throwOnError(b200c_copy_d2h(outPort->buffer().template as<void *>(), inPort->buffer().template as<const void *>(), n*_bytes, _device, nullptr), "HbmToHost::work()");
        //the host buffer goes back to its owner with consume()/produce(): the copy must have finished
        throwOnError(b200c_stream_sync(_device, nullptr), "HbmBridge::work()");
        inPort->consume(n);
        outPort->produce(n);
    }

    //labels ride along unchanged (the default propagation of Pothos::Block)
    void propagateLabels(const Pothos::InputPort *port)
    {
        auto outPort = this->output(0);
        for (const auto &label : port->labels()) outPort->postLabel(label);
    }

private:
    int _device;
    size_t _bytes;
};

static int bridgeDevice(void)
{
    const char *env = std::getenv("B200C_DEVICE");
    return env ? std::atoi(env) : 0;
}

static Pothos::Block *hostToHbmFactory(const Pothos::DType &dtype) { return new HbmBridge<true>(dtype, bridgeDevice()); }
static Pothos::Block *hbmToHostFactory(const Pothos::DType &dtype) { return new HbmBridge<false>(dtype, bridgeDevice()); }

static Pothos::BlockRegistry registerHostToHbm("/b200c/host_to_hbm", &hostToHbmFactory);
static Pothos::BlockRegistry registerHbmToHost("/b200c/hbm_to_host", &hbmToHostFactory);
