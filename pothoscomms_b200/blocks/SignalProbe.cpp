// /comms/signal_probe (+ legacy /blocks/stream_probe) on B200: VALUE / RMS / MEAN of a window of the
// stream, computed where the samples live (HBM) and emitted on the "valueChanged" signal; call
// surface, defaults and rate limiting as in the reference (utility/SignalProbe.cpp:59-170).  This is
// the RMS checker at the end of the reference's FIR test topology (filter/TestFIRFilter.cpp:51,75-78).
#include <Pothos/Framework.hpp>

#include <algorithm>
#include <chrono>
#include <complex>
#include <cstdlib>
#include <string>

#include "DeviceBuffers.hpp"

using b200c_blocks::throwOnError;

template <typename ProbeType>
class SignalProbe : public Pothos::Block
{
public:
    SignalProbe(const Pothos::DType &dtype, const int code, const int device): _code(code), _device(device)
    {
        this->setupInput(0, dtype, b200c_blocks::kHbmDomain);
        this->registerCall(this, POTHOS_FCN_TUPLE(SignalProbe, value));
        this->registerCall(this, POTHOS_FCN_TUPLE(SignalProbe, setMode));
        this->registerCall(this, POTHOS_FCN_TUPLE(SignalProbe, getMode));
        this->registerCall(this, POTHOS_FCN_TUPLE(SignalProbe, setWindow));
        this->registerCall(this, POTHOS_FCN_TUPLE(SignalProbe, getWindow));
        this->registerCall(this, POTHOS_FCN_TUPLE(SignalProbe, setRate));
        this->registerCall(this, POTHOS_FCN_TUPLE(SignalProbe, getRate));
        this->registerProbe("value");
        this->registerSignal("valueChanged");
        this->input(0)->setReserve(1);
    }

    ProbeType value(void) { return _value; }
    void setMode(const std::string &mode) { _mode = mode; }
    std::string getMode(void) const { return _mode; }
    void setWindow(const size_t window) { _window = window; this->input(0)->setReserve(window); }
    size_t getWindow(void) const { return _window; }
    void setRate(const double rate) { _rate = rate; }
    double getRate(void) const { return _rate; }
    void activate(void) { _nextCalc = std::chrono::high_resolution_clock::now(); }

    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("SignalProbe::getInputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceCircularBufferManager(_device));
    }

    void work(void)
    {
        auto inPort = this->input(0);
        const size_t N = std::min(_window, inPort->elements());
        if (N == 0) return;
        const void *x = inPort->buffer().as<const void *>();
        inPort->consume(N);

        //at most `rate` calculations per second; 0 = every window (:136-145)
        const auto now = std::chrono::high_resolution_clock::now();
        if (_rate != 0.0 and now < _nextCalc) return;
        if (_rate != 0.0) _nextCalc += std::chrono::duration_cast<std::chrono::high_resolution_clock::duration>(std::chrono::nanoseconds((long long)(1e9/_rate)));

        const int mode = _mode == "VALUE" ? B200C_PROBE_VALUE : _mode == "RMS" ? B200C_PROBE_RMS : _mode == "MEAN" ? B200C_PROBE_MEAN : -1;
        if (mode >= 0)
        {
            double v[2] = {0.0, 0.0};
            throwOnError(b200c_probe(_code, mode, x, N, v, _device, nullptr), "SignalProbe::work()");
            _value = makeValue(v, ProbeType());
        }
        this->emitSignal("valueChanged", _value);
    }

private:
    static double makeValue(const double (&v)[2], double) { return v[0]; }
    static std::complex<double> makeValue(const double (&v)[2], std::complex<double>) { return std::complex<double>(v[0], v[1]); }

    const int _code, _device;
    ProbeType _value = ProbeType(0);
    std::string _mode = "VALUE";   //defaults: utility/SignalProbe.cpp:63-66
    size_t _window = 1024;
    double _rate = 0.0;
    std::chrono::high_resolution_clock::time_point _nextCalc;
};

static Pothos::Block *signalProbeFactory(const Pothos::DType &dtype)
{
    const int code = b200c_blocks::dtypeCode(dtype);   //utility/SignalProbe.cpp:175-186
    if (code < 0) throw Pothos::InvalidArgumentException("signalProbeFactory("+dtype.toString()+")", "unsupported type");
    const char *env = std::getenv("B200C_DEVICE");
    const int device = env ? std::atoi(env) : 0;
    if (dtype.isComplex()) return new SignalProbe<std::complex<double>>(dtype, code, device);
    return new SignalProbe<double>(dtype, code, device);
}
static Pothos::BlockRegistry registerSignalProbe("/comms/signal_probe", &signalProbeFactory);
static Pothos::BlockRegistry registerSignalProbeOldPath("/blocks/stream_probe", &signalProbeFactory);
