// Minimal stand-in for the subset of <Pothos/Framework.hpp> (PothosCore >= 0.6) that the
// reference's FIR and FFT blocks touch (SURVEY.md section 8b lists it with file:line).
// PothosCore is not available in this build environment; this header exists so that the block
// sources in this directory -- and, for the oracle, the REFERENCE's own filter/FIRFilter.cpp
// (oracle/Makefile ref_fir) -- compile and can be driven from tests.  The blocks use only calls the
// reference blocks use; the BufferManager interface below (pop / push by byte count, readable window)
// is this shim's own, so blocks/DeviceBuffers.hpp needs its glue rewritten against PothosCore's
// ManagedBuffer / SharedBuffer signatures before it loads into a real Pothos process.
#pragma once

#include <algorithm>
#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeindex>
#include <cctype>
#include <typeinfo>
#include <utility>
#include <vector>

namespace Pothos {

// ------------------------------------------------------------------------ exceptions ---
class Exception : public std::runtime_error {
public:
    explicit Exception(const std::string &what, const std::string &arg = "")
        : std::runtime_error(arg.empty() ? what : what + ": " + arg) {}
    std::string message() const { return what(); }
    std::string displayText() const { return what(); }
};
class InvalidArgumentException : public Exception {
public:
    explicit InvalidArgumentException(const std::string &what, const std::string &arg = "")
        : Exception("Invalid argument: " + what, arg) {}
};

// thrown by get{Input,Output}BufferManager when a block cannot share buffers with the neighbour's domain
class PortDomainError : public Exception {
public:
    explicit PortDomainError(const std::string &what, const std::string &arg = "")
        : Exception("Port domain error: " + what, arg) {}
};

// ----------------------------------------------------------------------------- DType ---
class DType {
public:
    DType() = default;
    DType(const std::string &alias) { init(alias); }
    DType(const char *alias) { init(alias); }
    DType(const std::type_info &t)
    {
#define B200C_T(T, NAME) if (t == typeid(T)) { init(NAME); return; } if (t == typeid(std::complex<T>)) { init("complex_" NAME); return; }
        B200C_T(double, "float64") B200C_T(float, "float32") B200C_T(int64_t, "int64") B200C_T(int32_t, "int32")
        B200C_T(int16_t, "int16") B200C_T(int8_t, "int8")
        B200C_T(long long, "int64")
#undef B200C_T
        throw InvalidArgumentException("DType(typeid)", "unsupported type");
    }
    const std::string &name() const { return _name; }
    std::string toString() const { return _name; }
    size_t size() const { return _size; }
    bool isComplex() const { return _complex; }
    bool operator==(const DType &o) const { return _name == o._name; }
    bool operator!=(const DType &o) const { return !(*this == o); }

private:
    void init(const std::string &alias)
    {
        static const std::map<std::string, std::pair<std::string, size_t>> table = {
            {"float64", {"float64", 8}}, {"double", {"float64", 8}}, {"float32", {"float32", 4}}, {"float", {"float32", 4}},
            {"int64", {"int64", 8}}, {"int32", {"int32", 4}}, {"int", {"int32", 4}}, {"int16", {"int16", 2}},
            {"short", {"int16", 2}}, {"int8", {"int8", 1}}, {"char", {"int8", 1}}};
        std::string a = alias;
        _complex = a.rfind("complex_", 0) == 0;
        if (_complex) a = a.substr(8);
        if (a == "complex64") { _complex = true; a = "float32"; }
        if (a == "complex128") { _complex = true; a = "float64"; }
        auto it = table.find(a);
        if (it == table.end()) throw InvalidArgumentException("DType(" + alias + ")", "unknown type alias");
        _name = (_complex ? "complex_" : "") + it->second.first;
        _size = it->second.second * (_complex ? 2 : 1);
    }
    std::string _name = "unspecified";
    size_t _size = 1;
    bool _complex = false;
};

// ---------------------------------------------------------------------------- Object ---
// Type-erased value with the handful of conversions the FIR/FFT call surface needs.
class Object {
public:
    Object() = default;
    template <typename T> Object(const T &v) : _p(std::make_shared<Holder<T>>(v)), _t(&typeid(T)) {}
    Object(const char *s) : Object(std::string(s)) {}
    explicit operator bool() const { return bool(_p); }
    const std::type_info &type() const { return _t ? *_t : typeid(void); }
    bool canConvert(const std::type_info &to) const
    {
        if (!_p) return false;
        if (type() == to) return true;
        return isNumber(type()) && isNumber(to);
    }
    template <typename T> T convert() const
    {
        if (!_p) throw Exception("Object::convert()", "null object");
        if (type() == typeid(T)) return static_cast<const Holder<T> *>(_p.get())->v;
        return Convert<T>::from(*this);
    }
    template <typename T> const T &extract() const { return static_cast<const Holder<T> *>(_p.get())->v; }
    // Pothos::Object's explicit cast, `double(label.data)` at filter/FIRFilter.cpp:319
    template <typename T> explicit operator T() const { return convert<T>(); }

private:
    struct Base { virtual ~Base() = default; };
    template <typename T> struct Holder : Base { explicit Holder(const T &x) : v(x) {} T v; };
    static bool isNumber(const std::type_info &t)
    {
        return t == typeid(double) || t == typeid(float) || t == typeid(int) || t == typeid(long) || t == typeid(long long) ||
               t == typeid(unsigned) || t == typeid(unsigned long) || t == typeid(unsigned long long) || t == typeid(bool) ||
               t == typeid(short) || t == typeid(unsigned short);
    }
    double asDouble() const
    {
#define B200C_N(T) if (type() == typeid(T)) return (double)extract<T>();
        B200C_N(double) B200C_N(float) B200C_N(int) B200C_N(long) B200C_N(long long) B200C_N(unsigned) B200C_N(unsigned long)
        B200C_N(unsigned long long) B200C_N(bool) B200C_N(short) B200C_N(unsigned short)
#undef B200C_N
        throw Exception("Object::convert()", std::string("cannot convert ") + type().name());
    }
    template <typename T, typename = void> struct Convert {
        static T from(const Object &o) { throw Exception("Object::convert()", std::string("cannot convert ") + o.type().name()); }
    };
    template <typename T> struct Convert<T, typename std::enable_if<std::is_arithmetic<T>::value>::type> {
        static T from(const Object &o) { return (T)o.asDouble(); }
    };
    std::shared_ptr<Base> _p;
    const std::type_info *_t = nullptr;
};
// vector<double> -> vector<complex<double>> (a REAL designer wired to a COMPLEX-taps filter)
template <> struct Object::Convert<std::vector<std::complex<double>>, void> {
    static std::vector<std::complex<double>> from(const Object &o)
    {
        if (o.type() == typeid(std::vector<double>)) {
            const auto &v = o.extract<std::vector<double>>();
            return std::vector<std::complex<double>>(v.begin(), v.end());
        }
        throw Exception("Object::convert()", "cannot convert to complex taps");
    }
};

// ----------------------------------------------------------------------------- Label ---
struct Label {
    Label() = default;
    template <typename T> Label(const std::string &id_, const T &data_, unsigned long long index_, size_t width_ = 1)
        : id(id_), data(data_), index(index_), width(width_) {}
    // index*mult/div, width*mult/div (>= 1), as Pothos::Label::toAdjusted
    Label toAdjusted(size_t mult, size_t div) const
    {
        Label l = *this;
        l.index = index * mult / div;
        l.width = std::max<size_t>(1, width * mult / div);
        return l;
    }
    std::string id;
    Object data;
    unsigned long long index = 0;
    size_t width = 1;
};

// ----------------------------------------------------------------------- BufferChunk ---
// A view of (host OR device) memory.  The allocating constructor makes host memory, as the
// reference uses it for the burst flush buffer (filter/FIRFilter.cpp:268); the device FIR
// block never needs it (the kernel synthesises the zero tail).
class BufferChunk {
public:
    BufferChunk() = default;
    BufferChunk(size_t address_, size_t length_, const DType &dt = DType("int8")) : address(address_), length(length_), dtype(dt) {}
    BufferChunk(const DType &dt, size_t numElems) : length(dt.size() * numElems), dtype(dt), _own(new char[std::max<size_t>(dt.size() * numElems, 1)])
    {
        address = reinterpret_cast<size_t>(_own.get());
    }
    size_t elements() const { return length / dtype.size(); }
    template <typename T> T as() const { return reinterpret_cast<T>(address); }
    template <typename T> operator T *() const { return reinterpret_cast<T *>(address); }
    size_t address = 0;
    size_t length = 0;
    DType dtype = DType("int8");

private:
    std::shared_ptr<char[]> _own;
};

// --------------------------------------------------------------------- BufferManager ---
struct BufferManagerArgs {
    size_t bufferSize = 8 * 1024 * 1024;
    size_t numBuffers = 4;
    long nodeAffinity = -1;
};

// The two roles the reference's blocks ask for: "circular" (contiguous sliding window over the
// stream, filter/FIRFilter.cpp:196-199) and "generic" (slabs of bufferSize, fft/FFT.cpp:54-59).
class BufferManager {
public:
    typedef std::shared_ptr<BufferManager> Sptr;
    virtual ~BufferManager() = default;
    virtual void init(const BufferManagerArgs &args) = 0;
    virtual bool empty() const = 0;                  // no room / no buffer to hand out
    virtual const BufferChunk &front() const = 0;    // the next writable region
    virtual void pop(size_t numBytes) = 0;           // `numBytes` of front() were filled
    virtual void push(size_t numBytes) = 0;          // `numBytes` were released by the reader
    virtual std::string domain() const { return ""; } // "" = host memory
    // Named factories, as PothosCore's plugin tree /framework/buffer_manager/<name> provides them.  The
    // shim itself registers none (device blocks bring their own managers); a host that wants the reference's
    // make("circular") (filter/FIRFilter.cpp:198) registers one -- oracle/ref_fir_wrap.cpp does.
    typedef std::function<Sptr(const BufferManagerArgs &)> Factory;
    static void registerFactory(const std::string &name, Factory f) { factories()[name] = std::move(f); }
    static Sptr make(const std::string &name, const BufferManagerArgs &args = BufferManagerArgs())
    {
        auto it = factories().find(name);
        if (it == factories().end()) throw Exception("BufferManager::make(" + name + ")", "no such buffer manager factory in this host");
        return it->second(args);
    }

private:
    static std::map<std::string, Factory> &factories()
    {
        static std::map<std::string, Factory> f;
        return f;
    }
};

// ------------------------------------------------------------------------------ ports ---
class Block;

class InputPort {
public:
    size_t elements() const { return _bytes / _dtype.size(); }
    const std::vector<Label> &labels() const { return _labels; }
    void setReserve(size_t numElements) { _reserve = numElements; }
    size_t reserve() const { return _reserve; }
    // the readable window: history + new data, contiguous (circular manager)
    BufferChunk buffer() const { return BufferChunk(_addr, _bytes, _dtype); }
    void consume(size_t numElements) { _pendingConsume += numElements; }
    const DType &dtype() const { return _dtype; }
    const std::string &domain() const { return _domain; }   // "" = host memory
    unsigned long long totalElements() const { return _totalConsumed; }

private:
    friend class Block;
    friend class Harness;
    DType _dtype;
    std::string _domain;
    size_t _addr = 0, _bytes = 0, _reserve = 0, _pendingConsume = 0;
    unsigned long long _totalConsumed = 0;
    std::vector<Label> _labels;   // indices relative to the front of buffer()
};

class OutputPort {
public:
    size_t elements() const { return _bytes / _dtype.size(); }
    BufferChunk buffer() const { return BufferChunk(_addr, _bytes, _dtype); }
    void produce(size_t numElements) { _pendingProduce += numElements; }
    void postLabel(Label &&l) { _posted.emplace_back(std::move(l)); }
    void postLabel(const Label &l) { _posted.push_back(l); }
    const DType &dtype() const { return _dtype; }
    const std::string &domain() const { return _domain; }   // "" = host memory

private:
    friend class Block;
    friend class Harness;
    DType _dtype;
    std::string _domain;
    size_t _addr = 0, _bytes = 0, _pendingProduce = 0;
    std::vector<Label> _posted;   // indices relative to the element about to be produced
};

// ------------------------------------------------------------------------------ Block ---
class Block {
public:
    virtual ~Block() = default;
    virtual void work() {}
    virtual void activate() {}
    virtual void deactivate() {}
    virtual void propagateLabels(const InputPort *) {}
    virtual BufferManager::Sptr getInputBufferManager(const std::string &, const std::string &) { return BufferManager::Sptr(); }
    virtual BufferManager::Sptr getOutputBufferManager(const std::string &, const std::string &) { return BufferManager::Sptr(); }

    InputPort *input(size_t i) { return &_inputs.at(i); }
    OutputPort *output(size_t i) { return &_outputs.at(i); }
    size_t numInputs() const { return _inputs.size(); }
    size_t numOutputs() const { return _outputs.size(); }
    bool isActive() const { return _active; }

    // Proxy-style call by name (filter.call("setTaps", taps) in the reference's tests)
    Object call(const std::string &name, const std::vector<Object> &args = {})
    {
        auto it = _calls.find(name);
        if (it == _calls.end()) throw Exception("Block::call(" + name + ")", "no such registered call");
        return it->second(args);
    }
    template <typename... A> Object call(const std::string &name, const A &...a) { return call(name, std::vector<Object>{Object(a)...}); }
    bool hasCall(const std::string &name) const { return _calls.count(name) != 0; }

    // Signals: the shim's stand-in for Topology::connect(src, "signal", dst, "slot") -- an emitted
    // signal calls the connected registered calls synchronously and is remembered for inspection.
    void connectSignal(const std::string &signal, Block *dst, const std::string &slot)
    {
        if (!_signals.count(signal)) throw Exception("Block::connectSignal(" + signal + ")", "no such registered signal");
        _signalConns[signal].push_back(std::make_pair(dst, slot));
    }
    bool hasSignal(const std::string &name) const { return _signals.count(name) != 0; }
    const std::vector<Object> *lastSignal(const std::string &name) const
    {
        auto it = _lastSignal.find(name);
        return it == _lastSignal.end() ? nullptr : &it->second;
    }
    size_t signalCount(const std::string &name) const { auto it = _signalCounts.find(name); return it == _signalCounts.end() ? 0 : it->second; }

protected:
    void registerSignal(const std::string &name) { _signals[name] = true; }
    // registerProbe("value"): slot "probeValue" that emits "valueTriggered" with the call's result
    void registerProbe(const std::string &name)
    {
        std::string cap = name;
        if (!cap.empty()) cap[0] = (char)std::toupper((unsigned char)cap[0]);
        const std::string sig = name + "Triggered";
        _signals[sig] = true;
        _calls["probe" + cap] = [this, name, sig](const std::vector<Object> &) {
            const std::vector<Object> args{this->call(name)};
            _lastSignal[sig] = args;
            _signalCounts[sig]++;
            for (auto &c : _signalConns[sig]) c.first->call(c.second, args);
            return Object();
        };
    }
    template <typename... A> void emitSignal(const std::string &name, const A &...a)
    {
        if (!_signals.count(name)) throw Exception("Block::emitSignal(" + name + ")", "no such registered signal");
        const std::vector<Object> args{Object(a)...};
        _lastSignal[name] = args;
        _signalCounts[name]++;
        for (auto &c : _signalConns[name]) c.first->call(c.second, args);
    }
    // third argument: the port's buffer domain, as Pothos::Block::setupInput(name, dtype, domain)
    void setupInput(size_t index, const DType &dt, const std::string &domain = "")
    {
        if (_inputs.size() <= index) _inputs.resize(index + 1);
        _inputs[index]._dtype = dt;
        _inputs[index]._domain = domain;
    }
    void setupOutput(size_t index, const DType &dt, const std::string &domain = "")
    {
        if (_outputs.size() <= index) _outputs.resize(index + 1);
        _outputs[index]._dtype = dt;
        _outputs[index]._domain = domain;
    }
    // registerCall(this, POTHOS_FCN_TUPLE(Class, method))
    template <typename C, typename R, typename... A>
    void registerCall(C *self, const std::string &name, R (C::*fn)(A...))
    {
        _calls[name] = [self, fn, name](const std::vector<Object> &args) { return invoke<C, R, decltype(fn), A...>(self, fn, name, args, std::index_sequence_for<A...>{}); };
    }
    template <typename C, typename R, typename... A>
    void registerCall(C *self, const std::string &name, R (C::*fn)(A...) const)
    {
        _calls[name] = [self, fn, name](const std::vector<Object> &args) { return invoke<C, R, decltype(fn), A...>(self, fn, name, args, std::index_sequence_for<A...>{}); };
    }

private:
    friend class Harness;
    template <typename C, typename R, typename F, typename... A, size_t... I>
    static Object invoke(C *self, F fn, const std::string &name, const std::vector<Object> &args, std::index_sequence<I...>)
    {
        if (args.size() != sizeof...(A)) throw Exception("Block::call(" + name + ")", "wrong number of arguments");
        return ret<R>([&] { return (self->*fn)(args[I].template convert<typename std::decay<A>::type>()...); });
    }
    template <typename R, typename L> static typename std::enable_if<std::is_void<R>::value, Object>::type ret(L &&l) { l(); return Object(); }
    template <typename R, typename L> static typename std::enable_if<!std::is_void<R>::value, Object>::type ret(L &&l) { return Object(l()); }

    std::vector<InputPort> _inputs;
    std::vector<OutputPort> _outputs;
    std::map<std::string, std::function<Object(const std::vector<Object> &)>> _calls;
    std::map<std::string, bool> _signals;
    std::map<std::string, std::vector<std::pair<Block *, std::string>>> _signalConns;
    std::map<std::string, std::vector<Object>> _lastSignal;
    std::map<std::string, size_t> _signalCounts;
    bool _active = false;
};

#define POTHOS_FCN_TUPLE(Class, method) #method, &Class::method

// --------------------------------------------------------------------- BlockRegistry ---
class BlockRegistry {
public:
    typedef std::function<Block *(const std::vector<Object> &)> Factory;
    template <typename... A> BlockRegistry(const std::string &path, Block *(*factory)(A...))
    {
        table()[path] = [factory, path](const std::vector<Object> &args) {
            if (args.size() != sizeof...(A)) throw InvalidArgumentException("BlockRegistry::make(" + path + ")", "wrong number of factory arguments");
            return callFactory(factory, args, std::index_sequence_for<A...>{});
        };
    }
    template <typename... A> static Block *make(const std::string &path, const A &...a) { return makeFromObjects(path, std::vector<Object>{Object(a)...}); }
    static Block *makeFromObjects(const std::string &path, const std::vector<Object> &args)
    {
        auto it = table().find(path);
        if (it == table().end()) throw Exception("BlockRegistry::make(" + path + ")", "no such registry path");
        return it->second(args);
    }
    static bool doesBlockExist(const std::string &path) { return table().count(path) != 0; }
    static std::vector<std::string> paths()
    {
        std::vector<std::string> out;
        for (const auto &kv : table()) out.push_back(kv.first);
        return out;
    }

private:
    template <typename... A, size_t... I>
    static Block *callFactory(Block *(*factory)(A...), const std::vector<Object> &args, std::index_sequence<I...>)
    {
        return factory(args[I].template convert<typename std::decay<A>::type>()...);
    }
    static std::map<std::string, Factory> &table()
    {
        static std::map<std::string, Factory> t;
        return t;
    }
};

} // namespace Pothos
