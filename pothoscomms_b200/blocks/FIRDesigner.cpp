// /comms/fir_designer (+ legacy /blocks/fir_designer): host-side tap generator with the
// reference's call surface, defaults, validation and "tapsChanged" signal
// (filter/FIRDesigner.cpp:143-193 calls, :128-141 defaults, :387-477 recalculate).
// The taps stay on the host (north_star); a connected /comms/fir_filter receives them through
// setTaps() and uploads its own device tables.  The maths the reference takes from Spuce lives
// in TapDesign.cpp.
#include <Pothos/Framework.hpp>

#include <algorithm>
#include <cctype>
#include <complex>
#include <iostream>
#include <string>
#include <vector>

#include "TapDesign.hpp"

namespace {

struct DesignSpec {
    std::string filterType = "GAUSSIAN", bandType = "LOW_PASS", windowType = "hann";   // filter/FIRDesigner.cpp:128-130
    std::vector<double> windowArgs;
    double gain = 1.0, sampRate = 1.0, freqLower = 0.1, freqUpper = 0.2, transBw = 0.1, alpha = 0.5;   // :131-136
    double weight = 1.0, stopDB = 60.0, passDB = 0.1;                                                    // :137-139
    size_t numTaps = 51;                                                                                  // :140
};

bool usesUpperFrequency(const std::string &band)
{
    return band == "BAND_PASS" || band == "BAND_STOP" || band == "COMPLEX_BAND_PASS" || band == "COMPLEX_BAND_STOP";
}

[[noreturn]] void fail(const std::string &why) { throw Pothos::Exception("FIRDesigner()", why); }

} // namespace

class FIRDesigner : public Pothos::Block
{
public:
    static Block *make(void) { return new FIRDesigner(); }

    FIRDesigner(void)
    {
#define B200C_PARAM(setter, getter) \
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRDesigner, setter)); \
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRDesigner, getter));
        B200C_PARAM(setBandType, bandType)
        B200C_PARAM(setFilterType, filterType)
        B200C_PARAM(setWindowType, windowType)
        B200C_PARAM(setWindowArgs, windowArgs)
        B200C_PARAM(setSampleRate, sampleRate)
        this->registerCall(this, POTHOS_FCN_TUPLE(FIRDesigner, setFrequencies));
        B200C_PARAM(setFrequencyLower, frequencyLower)
        B200C_PARAM(setFrequencyUpper, frequencyUpper)
        B200C_PARAM(setBandwidthTrans, bandwidthTrans)
        B200C_PARAM(setNumTaps, numTaps)
        B200C_PARAM(setAlpha, alpha)
        B200C_PARAM(setStopDB, stopDB)
        B200C_PARAM(setPassDB, passDB)
        B200C_PARAM(setGain, gain)
#undef B200C_PARAM
        this->registerSignal("tapsChanged");
        this->recalculate();
    }

    // every setter stores its value and re-emits the taps (a no-op until the block is active)
    void setFilterType(const std::string &type)
    {
        // a band name given as filter type is the pre-0.5 usage: SINC prototype of that band (:195-214)
        static const char *bands[] = {"LOW_PASS", "HIGH_PASS", "BAND_PASS", "BAND_STOP", "COMPLEX_BAND_PASS", "COMPLEX_BAND_STOP"};
        if (std::find(std::begin(bands), std::end(bands), type) != std::end(bands)) {
            std::cerr << "FIRDesigner: filter type '" << type << "' should now be used as a band type, with filter type set to 'SINC'" << std::endl;
            _s.filterType = "SINC";
            _s.bandType = type;
        } else {
            _s.filterType = type;
        }
        this->recalculate();
    }
    std::string filterType(void) const { return _s.filterType; }
    void setBandType(const std::string &type) { _s.bandType = type; this->recalculate(); }
    std::string bandType(void) const { return _s.bandType; }
    void setWindowType(const std::string &type) { _s.windowType = type; this->recalculate(); }
    std::string windowType(void) const { return _s.windowType; }
    void setWindowArgs(const std::vector<double> &args) { _s.windowArgs = args; this->recalculate(); }
    std::vector<double> windowArgs(void) const { return _s.windowArgs; }
    void setSampleRate(const double rate) { _s.sampRate = rate; this->recalculate(); }
    double sampleRate(void) const { return _s.sampRate; }
    void setFrequencies(const std::vector<double> &freqs)
    {
        if (freqs.size() > 0) _s.freqLower = freqs[0];
        if (freqs.size() > 1) _s.freqUpper = freqs[1];
        this->recalculate();
    }
    void setFrequencyLower(const double freq) { _s.freqLower = freq; this->recalculate(); }
    double frequencyLower(void) const { return _s.freqLower; }
    void setFrequencyUpper(const double freq) { _s.freqUpper = freq; this->recalculate(); }
    double frequencyUpper(void) const { return _s.freqUpper; }
    void setBandwidthTrans(const double bw) { _s.transBw = bw; this->recalculate(); }
    double bandwidthTrans(void) const { return _s.transBw; }
    void setNumTaps(const size_t num) { _s.numTaps = num; this->recalculate(); }
    size_t numTaps(void) const { return _s.numTaps; }
    void setAlpha(const double alpha) { _s.alpha = alpha; this->recalculate(); }
    double alpha(void) const { return _s.alpha; }
    void setPassDB(const double db) { _s.passDB = db; this->recalculate(); }
    double passDB(void) const { return _s.passDB; }
    void setStopDB(const double db) { _s.stopDB = db; this->recalculate(); }
    double stopDB(void) const { return _s.stopDB; }
    void setGain(const double gain) { _s.gain = gain; this->recalculate(); }
    double gain(void) const { return _s.gain; }

    void activate(void) { this->recalculate(); }

private:
    void validate(void) const;
    void recalculate(void);
    DesignSpec _s;
};

// the reference's parameter checks, in its order and with its messages (filter/FIRDesigner.cpp:395-416)
void FIRDesigner::validate(void) const
{
    const bool cx = _s.bandType.find("COMPLEX") != std::string::npos, stop = _s.bandType.find("STOP") != std::string::npos;
    const double nyq = _s.sampRate / 2;
    if (_s.numTaps == 0) fail("num taps must be positive");
    if (_s.sampRate <= 0) fail("sample rate must be positive");
    if (cx and _s.freqLower <= -nyq) fail("lower frequency below Nyquist range");
    if (not cx and _s.freqLower <= 0) fail("lower frequency must be positive");
    if (_s.freqLower >= nyq) fail("lower frequency above Nyquist range");
    if (usesUpperFrequency(_s.bandType)) {
        if (_s.numTaps % 2 == 0) fail("Band pass or Band stop FIRs must have an odd number of taps");
        if (cx and _s.freqUpper <= -nyq) fail("upper frequency below Nyquist range");
        if (not cx and _s.freqUpper <= 0) fail("upper frequency must be positive");
        if (_s.freqUpper >= nyq) fail("upper frequency above Nyquist range");
        if (_s.freqUpper <= _s.freqLower) fail("upper frequency <= lower frequency");
    }
    if (_s.filterType == "MAXFLAT" and stop) fail("Can not use MAXFLAT as prototype for stop-band filter, please choose another type");
    if (_s.filterType == "REMEZ") {
        if (_s.transBw <= 0) fail("Transition Bandwidth must be > 0");
        if (_s.passDB <= 0) fail("Passband Attenuation must be > 0");
        if (_s.stopDB <= 0) fail("Stopband Attenuation must be > 0");
    }
}

void FIRDesigner::recalculate(void)
{
    if (not this->isActive()) return;
    this->validate();

    if (_s.filterType == "REMEZ") {
        // the transition width rides in alpha, the pass/stop error weight comes from the two dB figures (:420-436)
        _s.alpha = _s.transBw / _s.sampRate;
        const size_t need = b200c_design::remez_estimate_num_taps(_s.alpha, _s.passDB, _s.stopDB);
        if (need > _s.numTaps) {
            std::cerr << "FIRDesigner.Remez: order not large enough to meet the specification: use " << need << " taps, or "
                      << b200c_design::remez_estimate_atten(_s.numTaps, _s.alpha, _s.passDB) << " dB stop band, or a "
                      << b200c_design::remez_estimate_bw(_s.numTaps, _s.passDB, _s.stopDB) * _s.sampRate / 1e3
                      << " kHz transition, or more pass band ripple" << std::endl;
        }
        _s.weight = b200c_design::remez_estimate_weight(_s.passDB, _s.stopDB);
    }

    std::string kind = _s.filterType;
    std::transform(kind.begin(), kind.end(), kind.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    const double fl = _s.freqLower / _s.sampRate, fu = _s.freqUpper / _s.sampRate;
    const bool cx = _s.bandType == "COMPLEX_BAND_PASS" or _s.bandType == "COMPLEX_BAND_STOP";

    std::vector<double> window, real;
    std::vector<std::complex<double>> cplx;
    try {
        if (cx) cplx = b200c_design::design_complex_fir(kind, _s.bandType, _s.numTaps, fl, fu, _s.alpha, _s.weight);
        else real = b200c_design::design_fir(kind, _s.bandType, _s.numTaps, fl, fu, _s.alpha, _s.weight);
        window = b200c_design::design_window(_s.windowType, _s.numTaps, _s.windowArgs.empty() ? 0.0 : _s.windowArgs.at(0));
    }
    catch (const std::runtime_error &err) {
        throw Pothos::InvalidArgumentException(
            "Problem with creating taps for FIRDesigner(" + _s.filterType + "/" + _s.bandType + "):" + err.what(), "problem with input parameters?");
    }

    // gain and window, then the signal carries real or complex taps according to the band type (:456-476)
    if (cx) {
        for (size_t i = 0; i < cplx.size(); i++) cplx[i] *= _s.gain * window[i];
        this->emitSignal("tapsChanged", cplx);
    } else {
        for (size_t i = 0; i < real.size(); i++) real[i] *= _s.gain * window[i];
        this->emitSignal("tapsChanged", real);
    }
}

static Pothos::BlockRegistry registerFIRDesigner("/comms/fir_designer", &FIRDesigner::make);
static Pothos::BlockRegistry registerFIRDesignerOldPath("/blocks/fir_designer", &FIRDesigner::make);
