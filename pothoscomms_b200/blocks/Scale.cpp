// /comms/scale on B200 -- same registry path, factory (dtype) and calls (setFactor, getFactor,
// setLabelId, getLabelId) as the reference (math/Scale.cpp:25-160); arrayScale (:15-23) becomes
// one b200c_scale() launch over the port's HBM buffer.  A label carrying the id set with
// setLabelId() changes the factor from its position on, exactly as in the reference (:86-108):
// at index 0 it is applied, further in it ends this call's span.
// (The API shim has no vector dimension: elements are scalars or complex scalars.)
#include <Pothos/Framework.hpp>

#include <algorithm>
#include <cstdlib>
#include <string>

#include "DeviceBuffers.hpp"

using b200c_blocks::throwOnError;

class Scale : public Pothos::Block
{
public:
    Scale(const Pothos::DType &dtype, const int code, const int device): _code(code), _device(device)
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(Scale, setFactor));
        this->registerCall(this, POTHOS_FCN_TUPLE(Scale, getFactor));
        this->registerCall(this, POTHOS_FCN_TUPLE(Scale, setLabelId));
        this->registerCall(this, POTHOS_FCN_TUPLE(Scale, getLabelId));
        this->setupInput(0, dtype, b200c_blocks::kHbmDomain);
        this->setupOutput(0, dtype, b200c_blocks::kHbmDomain);
    }

    void setFactor(const double factor) { _factor = factor; }   //floatToQ happens inside b200c_scale (Scale.cpp:44)
    double getFactor(void) const { return _factor; }
    void setLabelId(const std::string &id) { _labelId = id; }
    std::string getLabelId(void) const { return _labelId; }

    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("Scale::getInputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceCircularBufferManager(_device));
    }
    Pothos::BufferManager::Sptr getOutputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("Scale::getOutputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceSlabBufferManager(_device));
    }

    void work(void)
    {
        auto inPort = this->input(0);
        auto outPort = this->output(0);
        size_t elems = std::min(inPort->elements(), outPort->elements());   //workInfo().minElements
        if (elems == 0) return;

        if (not _labelId.empty()) for (const auto &label : inPort->labels())
        {
            if (label.index >= elems) break;
            if (label.id != _labelId) continue;
            if (label.index == 0) this->setFactor(label.data.template convert<double>());
            else { elems = label.index; break; }   //the next call starts at this label
        }

        throwOnError(b200c_scale(_code, _factor, inPort->buffer().as<const void *>(), outPort->buffer().as<void *>(), elems, _device, nullptr), "Scale::work()");
        inPort->consume(elems);
        outPort->produce(elems);
    }

private:
    const int _code, _device;
    double _factor = 0.0;
    std::string _labelId;
};

static Pothos::Block *scaleFactory(const Pothos::DType &dtype)
{
    const int code = b200c_blocks::dtypeCode(dtype);   //all twelve rows of math/Scale.cpp:147-153
    if (code < 0) throw Pothos::InvalidArgumentException("scaleFactory("+dtype.toString()+")", "unsupported type");
    const char *env = std::getenv("B200C_DEVICE");
    return new Scale(dtype, code, env ? std::atoi(env) : 0);
}
static Pothos::BlockRegistry registerScale("/comms/scale", &scaleFactory);
