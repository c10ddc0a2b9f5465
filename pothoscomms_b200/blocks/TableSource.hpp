// Shared part of the two table-driven sources (/comms/waveform_source, /comms/noise_source): a lookup
// table that is computed on the host on setter calls, kept in HBM, and walked by the device in work()
// (b200c_table_source).  Holds what both reference blocks repeat: the amplitude / offset / waveform
// settings with their calls, the element conversion Type(scalar * val + offset)
// (waveform/WaveformSource.cpp:262-272 == waveform/NoiseSource.cpp:228-238), and the output manager.
#pragma once
#include <Pothos/Framework.hpp>

#include <complex>
#include <string>
#include <vector>

#include "DeviceBuffers.hpp"

namespace b200c_blocks {

template <typename T> struct IsComplex { static const bool value = false; };
template <typename T> struct IsComplex<std::complex<T>> { static const bool value = true; };

template <typename Derived, typename Type>
class TableSource : public Pothos::Block
{
public:
    TableSource(const Pothos::DType &dtype, const int code, const int device, const std::string &wave):
        _code(code), _device(device), _wave(wave)
    {
        this->setupOutput(0, dtype, b200c_blocks::kHbmDomain);
        this->registerCall(this, "setWaveform", &TableSource::setWaveform);
        this->registerCall(this, "getWaveform", &TableSource::getWaveform);
        this->registerCall(this, "setOffset", &TableSource::setOffset);
        this->registerCall(this, "getOffset", &TableSource::getOffset);
        this->registerCall(this, "setAmplitude", &TableSource::setAmplitude);
        this->registerCall(this, "getAmplitude", &TableSource::getAmplitude);
    }
    ~TableSource() override { if (_devTable) b200c_dev_free(_devTable, _device); }

    Pothos::BufferManager::Sptr getOutputBufferManager(const std::string &, const std::string &domain) override
    {
        requireHbmPeer("TableSource::getOutputBufferManager()", domain);
        return Pothos::BufferManager::Sptr(new DeviceSlabBufferManager(_device));
    }

    void activate(void) override { this->refresh(); }

    void setWaveform(const std::string &wave) { _wave = wave; this->refresh(); }
    std::string getWaveform(void) { return _wave; }
    void setOffset(const std::complex<double> &offset) { _offset = offset; this->refresh(); }
    std::complex<double> getOffset(void) { return _offset; }
    void setAmplitude(const std::complex<double> &scalar) { _scalar = scalar; this->refresh(); }
    std::complex<double> getAmplitude(void) { return _scalar; }

protected:
    //every setter recomputes the table once the block is active (the reference's updateTable() guard)
    void refresh(void)
    {
        if (not this->isActive()) return;
        static_cast<Derived *>(this)->fillTable(_table);
        const size_t bytes = _table.size()*sizeof(Type);
        if (bytes > _devBytes)
        {
            if (_devTable) b200c_dev_free(_devTable, _device);
            _devTable = nullptr; _devBytes = 0;
            throwOnError(b200c_dev_alloc(&_devTable, bytes, _device), "TableSource::refresh()");
            _devBytes = bytes;
        }
        //synchronous: this runs on setter calls, never inside work()
        throwOnError(b200c_copy_h2d(_devTable, _table.data(), bytes, _device, nullptr), "TableSource::refresh()");
        throwOnError(b200c_stream_sync(_device, nullptr), "TableSource::refresh()");
    }

    //out = Type(scalar * val + offset); real streams keep the real part
    Type element(const std::complex<double> &val) const { return convert(_scalar * val + _offset, Tag<IsComplex<Type>::value>()); }

    //out[i] = table[(index + i*step) & (entries-1)] into the output port's HBM buffer; returns the element count
    size_t walk(const size_t index, const size_t step, const std::string &where)
    {
        auto outPort = this->output(0);
        const size_t elems = outPort->elements();
        if (elems == 0) return 0;
        throwOnError(b200c_table_source(_code, _devTable, _table.size(), index, step, outPort->buffer().template as<void *>(), elems, _device, nullptr), where);
        outPort->produce(elems);
        return elems;
    }

    const int _code, _device;
    std::string _wave;
    std::complex<double> _offset = 0.0, _scalar = 1.0;

private:
    template <bool C> struct Tag {};
    static Type convert(const std::complex<double> &v, Tag<false>) { return Type(v.real()); }
    static Type convert(const std::complex<double> &v, Tag<true>) { return Type(v); }

    std::vector<Type> _table;
    void *_devTable = nullptr;
    size_t _devBytes = 0;
};

//the twelve rows both factories share (waveform/WaveformSource.cpp:277-286, waveform/NoiseSource.cpp:273-282)
template <template <typename> class Block, typename... Extra>
Pothos::Block *makeTableSource(const Pothos::DType &dtype, const std::string &factoryName, Extra... extra)
{
    const int code = dtypeCode(dtype);
    switch (code)
    {
    case B200C_F32: return new Block<float>(dtype, code, extra...);
    case B200C_CF32: return new Block<std::complex<float>>(dtype, code, extra...);
    case B200C_F64: return new Block<double>(dtype, code, extra...);
    case B200C_CF64: return new Block<std::complex<double>>(dtype, code, extra...);
    case B200C_I8: return new Block<int8_t>(dtype, code, extra...);
    case B200C_CI8: return new Block<std::complex<int8_t>>(dtype, code, extra...);
    case B200C_I16: return new Block<int16_t>(dtype, code, extra...);
    case B200C_CI16: return new Block<std::complex<int16_t>>(dtype, code, extra...);
    case B200C_I32: return new Block<int32_t>(dtype, code, extra...);
    case B200C_CI32: return new Block<std::complex<int32_t>>(dtype, code, extra...);
    case B200C_I64: return new Block<int64_t>(dtype, code, extra...);
    case B200C_CI64: return new Block<std::complex<int64_t>>(dtype, code, extra...);
    default: throw Pothos::InvalidArgumentException(factoryName+"("+dtype.toString()+")", "unsupported type");
    }
}

} // namespace b200c_blocks
