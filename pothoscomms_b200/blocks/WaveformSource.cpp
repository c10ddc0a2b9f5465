// /comms/waveform_source (+ legacy /blocks/waveform_source) on B200 -- same registry paths, factory
// (dtype) and calls as the reference (waveform/WaveformSource.cpp:66-91,110-182,265-293).  One period
// of CONST / SINE / RAMP / SQUARE is tabulated on the host on activate() and on every setter, in a
// power-of-two table grown until one sample advances at least 16 entries (:184-260), and uploaded to
// HBM; work() (:98-108) is one b200c_table_source() launch into the output port's HBM buffer, so the
// reference's test topology source -> fir_filter -> probe (filter/TestFIRFilter.cpp:49-51) starts in HBM.
#include <Pothos/Framework.hpp>

#include <cmath>
#include <cstdint>
#include <cstdlib>

#include "TableSource.hpp"

namespace {

const size_t kDefaultEntries = 4096, kMaxEntries = 1024*1024, kMinStep = 16;   //waveform/WaveformSource.cpp:10-12

//smallest power-of-two table in which `frac` periods per sample is at least kMinStep entries (:193-202)
size_t tableEntriesFor(const double frac)
{
    size_t entries = kDefaultEntries;
    while (frac != 0.0 and size_t(std::abs(std::llround(frac*entries))) < kMinStep and entries*2 <= kMaxEntries) entries *= 2;
    return entries;
}

} // namespace

template <typename Type>
class WaveformSource : public b200c_blocks::TableSource<WaveformSource<Type>, Type>
{
    typedef b200c_blocks::TableSource<WaveformSource<Type>, Type> Base;
public:
    WaveformSource(const Pothos::DType &dtype, const int code, const int device): Base(dtype, code, device, "CONST")
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(WaveformSource, setFrequency));
        this->registerCall(this, POTHOS_FCN_TUPLE(WaveformSource, getFrequency));
        this->registerCall(this, POTHOS_FCN_TUPLE(WaveformSource, setSampleRate));
        this->registerCall(this, POTHOS_FCN_TUPLE(WaveformSource, getSampleRate));
        this->registerCall(this, POTHOS_FCN_TUPLE(WaveformSource, setResolution));
        this->registerCall(this, POTHOS_FCN_TUPLE(WaveformSource, getResolution));
    }

    void work(void) override
    {
        //size_t wrap-around of the phase is harmless: only its low bits address the table
        _index += _step*this->walk(_index, _step, "WaveformSource::work()");
    }

    void setFrequency(const double &freq) { _freq = freq; this->refresh(); }
    double getFrequency(void) { return _freq; }
    void setSampleRate(const double &rate) { _rate = rate; this->refresh(); }
    double getSampleRate(void) { return _rate; }
    void setResolution(const double &res) { _res = res; this->refresh(); }
    double getResolution(void) { return _res; }

    //one period of the waveform; the table's size and the phase step follow from freq / rate / resolution
    void fillTable(std::vector<Type> &table)
    {
        const size_t n = tableEntriesFor(((_res == 0.0)? _freq : _res)/_rate);
        _step = size_t(std::llround((_freq/_rate)*n));
        if (_step == 0 and _freq != 0.0)
            throw Pothos::InvalidArgumentException("WaveformSource::updateTable()", "step size not achievable");

        enum { CONST, SINE, RAMP, SQUARE } kind;
        if (this->_wave == "CONST") kind = CONST;
        else if (this->_wave == "SINE") kind = SINE;
        else if (this->_wave == "RAMP") kind = RAMP;
        else if (this->_wave == "SQUARE") kind = SQUARE;
        else throw Pothos::InvalidArgumentException("WaveformSource::setWaveform("+this->_wave+")", "unknown waveform setting");

        table.resize(n);
        for (size_t i = 0; i < n; i++)
        {
            const size_t q = (i + (3*n)/4) % n;   //the imaginary part runs three quarters of a period ahead
            std::complex<double> val(1.0);
            switch (kind)
            {
            case CONST: break;
            case SINE: val = std::polar(1.0, 2*M_PI*i/n); break;
            case RAMP: val = std::complex<double>(2.0*i/(n-1) - 1.0, 2.0*q/(n-1) - 1.0); break;
            case SQUARE: val = std::complex<double>((i < n/2)? 0.0 : 1.0, (q < n/2)? 0.0 : 1.0); break;
            }
            table[i] = this->element(val);
        }
    }

private:
    size_t _index = 0, _step = 0;
    double _rate = 1.0, _freq = 0.0, _res = 0.0;   //defaults: waveform/WaveformSource.cpp:72-76
};

static Pothos::Block *waveformSourceFactory(const Pothos::DType &dtype)
{
    const char *env = std::getenv("B200C_DEVICE");
    return b200c_blocks::makeTableSource<WaveformSource>(dtype, "waveformSourceFactory", env ? std::atoi(env) : 0);
}

static Pothos::BlockRegistry registerWaveformSource("/comms/waveform_source", &waveformSourceFactory);
static Pothos::BlockRegistry registerWaveformSourceOldPath("/blocks/waveform_source", &waveformSourceFactory);
