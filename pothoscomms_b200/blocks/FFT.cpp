// /comms/fft on B200 -- same registry path and factory (dtype, numBins, inverse) as the
// reference (fft/FFT.cpp:39-95); FFTAux::transform (fft/FFTAux.h:21-44, kiss_fft underneath)
// is replaced by the batched sm_100a kernels behind the C ABI.
//
// The reference performs exactly ONE transform per work() call and sizes its output slabs to
// one transform (fft/FFT.cpp:54-72) -- one kernel launch per 32 KiB would starve a GPU.  This
// block does floor(min(inElems, outElems)/numBins) transforms per call and asks for large HBM
// slabs; the element stream is identical.
#include <Pothos/Framework.hpp>

#include <cstdlib>
#include <string>

#include "DeviceBuffers.hpp"

using b200c_blocks::throwOnError;

class FFT : public Pothos::Block
{
public:
    FFT(const Pothos::DType &dtype, const int dtypeCode, const size_t numBins, const bool inverse, const int device):
        _dtype(dtype), _dtypeCode(dtypeCode), _numBins(numBins), _inverse(inverse), _device(device), _fft(nullptr)
    {
        throwOnError(b200c_fft_create(&_fft, dtypeCode, numBins, inverse ? 1 : 0, device), "FFTFactory(" + dtype.toString() + ")");
        this->setupInput(0, dtype, b200c_blocks::kHbmDomain);
        this->setupOutput(0, dtype, b200c_blocks::kHbmDomain);
        this->input(0)->setReserve(_numBins);
        //Not in the reference (its FFT block registers no calls, fft/FFT.cpp:43-51); the
        //direction becomes switchable at run time, the factory argument stays the default.
        this->registerCall(this, POTHOS_FCN_TUPLE(FFT, setInverse));
        this->registerCall(this, POTHOS_FCN_TUPLE(FFT, getInverse));
        this->registerCall(this, POTHOS_FCN_TUPLE(FFT, getNumBins));
    }

    ~FFT(void)
    {
        b200c_fft_destroy(_fft);
    }

    void setInverse(const bool inverse)
    {
        if (inverse == _inverse) return;
        b200c_fft *fresh = nullptr;
        throwOnError(b200c_fft_create(&fresh, _dtypeCode, _numBins, inverse ? 1 : 0, _device), "FFT::setInverse()");
        b200c_fft_destroy(_fft);
        _fft = fresh;
        _inverse = inverse;
    }
    bool getInverse(void) const { return _inverse; }
    size_t getNumBins(void) const { return _numBins; }

    //! HBM slabs holding many transforms each (the reference: one transform per slab)
    Pothos::BufferManager::Sptr getOutputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("FFT::getOutputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceSlabBufferManager(_device));
    }

    Pothos::BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &domain)
    {
        b200c_blocks::requireHbmPeer("FFT::getInputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new b200c_blocks::DeviceCircularBufferManager(_device));
    }

    //! slab size hint for whoever initialises the output manager: a whole number of transforms
    size_t preferredOutputBytes(void) const
    {
        const size_t one = _numBins*_dtype.size();
        const size_t target = size_t(64) << 20;
        return one >= target ? one : (target/one)*one;
    }

    void work(void)
    {
        auto inPort = this->input(0);
        auto outPort = this->output(0);
        const size_t batch = std::min(inPort->elements(), outPort->elements())/_numBins;
        if (batch == 0) return; //the reserve of numBins (set in the constructor) is not met yet

        throwOnError(b200c_fft_run(_fft, inPort->buffer().as<const void *>(), outPort->buffer().as<void *>(), batch, nullptr), "FFT::work()");

        inPort->consume(batch*_numBins);
        outPort->produce(batch*_numBins);
    }

private:
    const Pothos::DType _dtype;
    const int _dtypeCode;
    const size_t _numBins;
    bool _inverse;
    const int _device;
    b200c_fft *_fft;
};

/***********************************************************************
 * registration -- fft/FFT.cpp:83-95
 **********************************************************************/
static Pothos::Block *FFTFactory(const Pothos::DType &dtype, const size_t numBins, const bool inverse)
{
    const int code = b200c_blocks::dtypeCode(dtype);
    //complex double, complex float and complex kiss_fft_scalar (= int16), fft/FFT.cpp:89-91
    if (code == B200C_CF64 or code == B200C_CF32 or code == B200C_CI16)
    {
        const char *env = std::getenv("B200C_DEVICE");
        return new FFT(dtype, code, numBins, inverse, env ? std::atoi(env) : 0);
    }
    throw Pothos::InvalidArgumentException("FFTFactory("+dtype.toString()+")", "unsupported type");
}
static Pothos::BlockRegistry registerFFT(
    "/comms/fft", &FFTFactory);
