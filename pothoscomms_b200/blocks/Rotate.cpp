// /comms/rotate on B200 -- out[n] = in[n] * exp(j*phase) for the six complex types, with the
// reference's registry path, factory and calls (math/Rotate.cpp:25-160: setPhase, getPhase,
// setLabelId, getLabelId) and its label-driven phase changes (:97-119); arrayRotate (:15-23)
// becomes one b200c_rotate() launch over the port's HBM buffer.
#include <Pothos/Framework.hpp>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "DeviceBuffers.hpp"

using b200c_blocks::throwOnError;

class Rotate : public Pothos::Block
{
public:
    Rotate(const Pothos::DType &dtype, const int code, const int device): _code(code), _device(device)
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(Rotate, setPhase));
        this->registerCall(this, POTHOS_FCN_TUPLE(Rotate, getPhase));
        this->registerCall(this, POTHOS_FCN_TUPLE(Rotate, setLabelId));
        this->registerCall(this, POTHOS_FCN_TUPLE(Rotate, getLabelId));
        this->setupInput(0, dtype, b200c_blocks::kHbmDomain);
        this->setupOutput(0, dtype, b200c_blocks::kHbmDomain);
    }

    void setPhase(const double phase) { _phase = phase; }   //floatToQ(polar(1, phase)) happens inside b200c_rotate (:74)
    double getPhase(void) const { return _phase; }
    void setLabelId(const std::string &id) { _labelId = id; }
    std::string getLabelId(void) const { return _labelId; }

    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("Rotate::getInputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceCircularBufferManager(_device));
    }
    Pothos::BufferManager::Sptr getOutputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("Rotate::getOutputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceSlabBufferManager(_device));
    }

    void work(void)
    {
        auto inPort = this->input(0);
        auto outPort = this->output(0);
        size_t elems = std::min(inPort->elements(), outPort->elements());
        if (elems == 0) return;

        if (not _labelId.empty()) for (const auto &label : inPort->labels())
        {
            if (label.index >= elems) break;
            if (label.id != _labelId) continue;
            if (label.index == 0) this->setPhase(label.data.template convert<double>());
            else { elems = label.index; break; }
        }

        throwOnError(b200c_rotate(_code, _phase, inPort->buffer().as<const void *>(), outPort->buffer().as<void *>(), elems, _device, nullptr), "Rotate::work()");
        inPort->consume(elems);
        outPort->produce(elems);
    }

private:
    const int _code, _device;
    double _phase = 0.0;
    std::string _labelId;
};

static Pothos::Block *rotateFactory(const Pothos::DType &dtype)
{
    const int code = b200c_blocks::dtypeCode(dtype);   //complex rows only, math/Rotate.cpp:145-157
    if (code < 0 or not dtype.isComplex()) throw Pothos::InvalidArgumentException("rotateFactory("+dtype.toString()+")", "unsupported type");
    const char *env = std::getenv("B200C_DEVICE");
    return new Rotate(dtype, code, env ? std::atoi(env) : 0);
}
static Pothos::BlockRegistry registerRotate("/comms/rotate", &rotateFactory);
