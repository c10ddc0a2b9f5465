// /comms/fir_filter on B200 -- the Pothos block, same registry paths, factory arguments,
// registered calls and work()/label/burst control flow as the reference
// (filter/FIRFilter.cpp:98-389), with the convolution nest (:286-302) replaced by ONE call
// into the sm_100a kernels through the C ABI (include/b200comms.h).  Written against the
// Pothos API subset of SURVEY.md section 8b; builds against the shim in ./shim when PothosCore is
// absent and against the real <Pothos/Framework.hpp> when present.
//
// The reference instantiates 18 class templates, one per (data, taps, Q) type row
// (:373-382); here the arithmetic types live in the kernels, so the block is templated only
// on what its call signatures need: the TapsType of setTaps/getTaps (:374-376).
#include <Pothos/Framework.hpp>

#include <complex>
#include <string>
#include <vector>

#include "DeviceBuffers.hpp"

using b200c_blocks::throwOnError;

template <typename TapsType>
class FIRFilter : public Pothos::Block
{
public:
    FIRFilter(const Pothos::DType &dtype, const int dtypeCode, const int device):
        M(1), L(1), K(1), _inputRequire(1),
        _waitTapsMode(false), _waitTapsArmed(false), _eobSampsLeft(0),
        _device(device), _fir(nullptr)
    {
        const bool complexTaps = not std::is_same<TapsType, double>::value;
        throwOnError(b200c_fir_create(&_fir, dtypeCode, complexTaps ? B200C_TAPS_COMPLEX : B200C_TAPS_REAL, device),
                     "FIRFilterFactory(" + dtype.toString() + ")");
        this->setupInput(0, dtype, b200c_blocks::kHbmDomain);
        this->setupOutput(0, dtype, b200c_blocks::kHbmDomain);
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, setTaps));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, getTaps));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, setDecimation));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, getDecimation));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, setInterpolation));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, getInterpolation));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, setWaitTaps));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, getWaitTaps));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, setFrameStartId));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, getFrameStartId));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, setFrameEndId));
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRFilter, getFrameEndId));
        this->setTaps(std::vector<TapsType>(1, TapsType(1))); //initial update
    }

    ~FIRFilter(void)
    {
        b200c_fir_destroy(_fir);
    }

    void setWaitTaps(const bool waitTaps) { _waitTapsMode = waitTaps; }
    bool getWaitTaps(void) const { return _waitTapsMode; }

    void setTaps(const std::vector<TapsType> &taps)
    {
        if (taps.empty()) throw Pothos::InvalidArgumentException("FIRFilter::setTaps()", "taps cannot be empty");
        //std::complex<double> is layout-compatible with double[2]: the ABI takes interleaved pairs
        throwOnError(b200c_fir_set_taps(_fir, reinterpret_cast<const double *>(taps.data()), taps.size()), "FIRFilter::setTaps()");
        _taps = taps;
        _waitTapsArmed = false; //got taps
        this->updateInternals();
    }
    std::vector<TapsType> getTaps(void) const { return _taps; }

    void setDecimation(const size_t decim)
    {
        if (decim == 0) throw Pothos::InvalidArgumentException("FIRFilter::setDecimation()", "decimation cannot be 0");
        throwOnError(b200c_fir_set_rates(_fir, decim, L), "FIRFilter::setDecimation()");
        M = decim;
        this->updateInternals();
    }
    size_t getDecimation(void) const { return M; }

    void setInterpolation(const size_t interp)
    {
        if (interp == 0) throw Pothos::InvalidArgumentException("FIRFilter::setInterpolation()", "interpolation cannot be 0");
        throwOnError(b200c_fir_set_rates(_fir, M, interp), "FIRFilter::setInterpolation()");
        L = interp;
        this->updateInternals();
    }
    size_t getInterpolation(void) const { return L; }

    void setFrameStartId(std::string id) { _frameStartId = id; }
    std::string getFrameStartId(void) const { return _frameStartId; }
    void setFrameEndId(std::string id) { _frameEndId = id; }
    std::string getFrameEndId(void) const { return _frameEndId; }

    //! always a circular buffer so the sliding window never sees a discontinuity -- in HBM
    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("FIRFilter::getInputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceCircularBufferManager(_device));
    }

    //! output slabs live in HBM as well, so a downstream device block reads them in place
    Pothos::BufferManager::Sptr getOutputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("FIRFilter::getOutputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceSlabBufferManager(_device));
    }

    void activate(void)
    {
        _waitTapsArmed = _waitTapsMode;
        _eobSampsLeft = 0;
    }

    // Burst bookkeeping, reference semantics of filter/FIRFilter.cpp:218-231: while no burst is
    // open, the first matching start label (payload = burst length in units of its width) or
    // end label fixes how many input elements remain until the end of the burst.
    void latchBurstEnd(const Pothos::InputPort *port)
    {
        if (_eobSampsLeft != 0) return;
        const bool useStart = not _frameStartId.empty(), useEnd = not _frameEndId.empty();
        for (const auto &lbl : port->labels())
        {
            if (useStart and lbl.id == _frameStartId and lbl.data.canConvert(typeid(size_t)))
            {
                _eobSampsLeft = lbl.index + lbl.data.template convert<size_t>()*lbl.width;
                return;
            }
            if (useEnd and lbl.id == _frameEndId)
            {
                _eobSampsLeft = lbl.index + lbl.width;
                return;
            }
        }
    }

    // Admission rule of filter/FIRFilter.cpp:237-258.  Returns how many of the `available`
    // elements this call may look at, or 0 after asking the scheduler for a larger reserve.
    size_t admitInput(Pothos::InputPort *port, const size_t available)
    {
        size_t usable = available, want = 0;
        if (_eobSampsLeft != 0)
        {
            if (_eobSampsLeft > available) want = _eobSampsLeft; //whole burst tail must be present
            else usable = _eobSampsLeft;                          //never read past the burst end
        }
        else if (available < _inputRequire) want = _inputRequire; //streaming: M + K - 1 elements
        port->setReserve(want);
        return want ? 0 : usable;
    }

    void work(void)
    {
        if (_waitTapsArmed) return; //setWaitTaps(true): hold the stream until taps arrive
        auto inPort = this->input(0);
        auto outPort = this->output(0);
        if (inPort->elements() == 0) return;

        this->latchBurstEnd(inPort);
        const size_t usable = this->admitInput(inPort, inPort->elements());
        if (usable == 0) return;

        //Burst flush: the reference copies the tail into a temporary buffer followed by K-1
        //zeros (filter/FIRFilter.cpp:265-272); the kernel synthesises that zero tail instead.
        const int zeroTail = (_eobSampsLeft != 0 and _eobSampsLeft < _inputRequire) ? 1 : 0;

        //The convolution nest (filter/FIRFilter.cpp:278-302): one launch, in place in HBM.
        //N = min((elems-(K-1))/M, outElems/L)*M comes back as `taken`, (N/M)*L as `made`.
        size_t taken = 0, made = 0;
        throwOnError(b200c_fir_run(_fir, inPort->buffer().template as<const void *>(), usable,
            outPort->buffer().template as<void *>(), outPort->elements(), zeroTail, &taken, &made, nullptr), "FIRFilter::work()");

        //K-1 elements stay behind in the circular input buffer as the next call's history
        if (_eobSampsLeft != 0) _eobSampsLeft -= taken;
        inPort->consume(taken);
        outPort->produce(made);
    }

    void propagateLabels(const Pothos::InputPort *port)
    {
        auto outputPort = this->output(0);
        for (const auto &label : port->labels())
        {
            auto newLabel = label.toAdjusted(L, M);
            if (label.id == "rxRate" and label.data.type() == typeid(double))
            {
                newLabel.data = Pothos::Object((label.data.template convert<double>()*L)/M);
            }
            outputPort->postLabel(std::move(newLabel));
        }
    }

private:
    void updateInternals(void)
    {
        //K and the input requirement come back from the ABI (filter/FIRFilter.cpp:335,353)
        throwOnError(b200c_fir_info(_fir, &K, &_inputRequire, nullptr, nullptr), "FIRFilter::updateInternals()");
    }

    std::vector<TapsType> _taps;
    size_t M, L, K, _inputRequire;
    bool _waitTapsMode;
    bool _waitTapsArmed;
    std::string _frameStartId;
    std::string _frameEndId;
    size_t _eobSampsLeft;
    int _device;
    b200c_fir *_fir;
};

/***********************************************************************
 * registration -- same paths and factory signature as filter/FIRFilter.cpp:369-389
 **********************************************************************/
static int currentDevice(void)
{
    const char *env = std::getenv("B200C_DEVICE");
    return env ? std::atoi(env) : 0;
}

static Pothos::Block *FIRFilterFactory(const Pothos::DType &dtype, const std::string &tapsType)
{
    const int code = b200c_blocks::dtypeCode(dtype);
    if (code >= 0 and tapsType == "REAL") return new FIRFilter<double>(dtype, code, currentDevice());
    if (code >= 0 and dtype.isComplex() and tapsType == "COMPLEX") return new FIRFilter<std::complex<double>>(dtype, code, currentDevice());
    throw Pothos::InvalidArgumentException("FIRFilterFactory("+dtype.toString()+")", "unsupported types");
}
static Pothos::BlockRegistry registerFIRFilter(
    "/comms/fir_filter", &FIRFilterFactory);

static Pothos::BlockRegistry registerFIRFilterOldPath(
    "/blocks/fir_filter", &FIRFilterFactory);
