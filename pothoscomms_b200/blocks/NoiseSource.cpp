// /comms/noise_source (+ legacy /blocks/noise_source) on B200 -- same registry paths, factory (dtype)
// and calls as the reference (waveform/NoiseSource.cpp:72-99,127-186,267-289).  As in the reference
// the stream is a 4096-entry pool of UNIFORM / NORMAL / LAPLACE / POISSON draws (redrawn on
// activate() and on every setter, :188-226) entered at a random position on each work() (:108); the
// pool is drawn on the host with libstdc++'s generators -- the reference's own -- and uploaded to
// HBM, and work()'s copy loop (:109-113) is one b200c_table_source() launch into the output port's
// HBM buffer.  (_fast has no setter in the reference, so its per-sample branch :115-125 is unreachable.)
// The reference seeds std::mt19937 from std::random_device; B200C_NOISE_SEED (environment, read when
// the block is made) fixes the seed instead so that tests can replay the stream.
#include <Pothos/Framework.hpp>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <random>

#include "TableSource.hpp"

static const size_t kPoolEntries = 4096;   //waveform/NoiseSource.cpp:11

template <typename Type>
class NoiseSource : public b200c_blocks::TableSource<NoiseSource<Type>, Type>
{
    typedef b200c_blocks::TableSource<NoiseSource<Type>, Type> Base;
public:
    NoiseSource(const Pothos::DType &dtype, const int code, const int device, const unsigned seed):
        Base(dtype, code, device, "NORMAL"), _gen(seed), _entry(0, kPoolEntries-1)
    {
        this->registerCall(this, POTHOS_FCN_TUPLE(NoiseSource, setMean));
        this->registerCall(this, POTHOS_FCN_TUPLE(NoiseSource, getMean));
        this->registerCall(this, POTHOS_FCN_TUPLE(NoiseSource, setB));
        this->registerCall(this, POTHOS_FCN_TUPLE(NoiseSource, getB));
    }

    void work(void) override
    {
        _index += _entry(_gen);   //every work() enters the pool somewhere else
        _index += this->walk(_index, 1, "NoiseSource::work()");
    }

    void setMean(const double mean) { _mean = mean; this->refresh(); }
    double getMean(void) const { return _mean; }
    void setB(const double b) { _b = b; this->refresh(); }
    double getB(void) const { return _b; }

    void fillTable(std::vector<Type> &pool)
    {
        pool.resize(kPoolEntries);
        //Each (re, im) pair is written as the reference writes it -- two draws inside one constructor
        //call -- because the order of those draws is the compiler's choice there too.
        if (this->_wave == "UNIFORM")
        {
            _uniform = std::uniform_real_distribution<>(_mean-_b, _mean+_b);
            for (auto &e : pool) e = this->element(std::complex<double>(_uniform(_gen), _uniform(_gen)));
        }
        else if (this->_wave == "NORMAL")
        {
            _normal = std::normal_distribution<>(_mean, _b);
            for (auto &e : pool) e = this->element(std::complex<double>(_normal(_gen), _normal(_gen)));
        }
        else if (this->_wave == "LAPLACE")
        {
            //the reference draws its Laplace variates from a uniform on mean +/- b, kept as is
            _uniform = std::uniform_real_distribution<>(_mean-_b, _mean+_b);
            for (auto &e : pool) e = this->element(std::complex<double>(laplaceDraw(), laplaceDraw()));
        }
        else if (this->_wave == "POISSON")
        {
            _poisson = std::poisson_distribution<>(_mean);
            for (auto &e : pool) e = this->element(std::complex<double>(_poisson(_gen), _poisson(_gen)));
        }
        else throw Pothos::InvalidArgumentException("NoiseSource::setWaveform("+this->_wave+")", "unknown waveform setting");
    }

private:
    //inverse-CDF Laplace draw (waveform/NoiseSource.cpp:238-245)
    double laplaceDraw(void)
    {
        const double u = _uniform(_gen);
        return (u < 0)? _mean + _b*std::log(1+u) : _mean - _b*std::log(1-u);
    }

    size_t _index = 0;
    double _mean = 0.0, _b = 1.0;   //defaults: waveform/NoiseSource.cpp:76-83
    std::mt19937 _gen;
    std::uniform_int_distribution<size_t> _entry;
    std::uniform_real_distribution<> _uniform;
    std::normal_distribution<> _normal;
    std::poisson_distribution<> _poisson;
};

static Pothos::Block *noiseSourceFactory(const Pothos::DType &dtype)
{
    const char *env = std::getenv("B200C_DEVICE");
    const char *seedEnv = std::getenv("B200C_NOISE_SEED");
    const unsigned seed = seedEnv ? (unsigned)std::strtoul(seedEnv, nullptr, 0) : std::random_device()();
    return b200c_blocks::makeTableSource<NoiseSource>(dtype, "noiseSourceFactory", env ? std::atoi(env) : 0, seed);
}

static Pothos::BlockRegistry registerNoiseSource("/comms/noise_source", &noiseSourceFactory);
static Pothos::BlockRegistry registerNoiseSourceOldPath("/blocks/noise_source", &noiseSourceFactory);
