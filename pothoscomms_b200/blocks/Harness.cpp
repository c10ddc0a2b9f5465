// Test harness for the block layer: a single-block stand-in for what a Pothos::Topology with
// /blocks/feeder_source -> block -> /blocks/collector_sink does in the reference's tests
// (filter/TestFIRFilter.cpp:19-53, fft/TestFFT.cpp:33-46), plus an extern "C" surface so that
// pytest can drive it through ctypes.  NOT part of the product data path: it only feeds host
// data into the block's device buffer managers, runs work() until it stops making progress
// (honouring reserves, labels and propagateLabels like the framework's post-work step) and
// drains produced elements back to the host.
#include <Pothos/Framework.hpp>

#include <chrono>
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "DeviceBuffers.hpp"

namespace Pothos {

class Harness {
public:
    Harness(Block *blk, int device, size_t inBytes, size_t outBytes) : _blk(blk), _device(device)
    {
        if (blk->numInputs() == 0 && blk->numOutputs() == 0) return;   // host-only block (designers): calls and signals only
        if (blk->numInputs() == 0) {   // a source (/comms/waveform_source, /comms/noise_source): output side only
            _outMgr = blk->getOutputBufferManager("0", b200c_blocks::kHbmDomain);
            if (!_outMgr) throw Exception("Harness()", "block does not provide device buffer managers");
            BufferManagerArgs oa;
            oa.bufferSize = outBytes; oa.numBuffers = 2;
            _outMgr->init(oa);
            return;
        }
        _inMgr = std::dynamic_pointer_cast<b200c_blocks::DeviceCircularBufferManager>(blk->getInputBufferManager("0", b200c_blocks::kHbmDomain));
        _sink = blk->numOutputs() == 0;   // a consumer only (/comms/signal_probe): no output side
        if (!_sink) _outMgr = blk->getOutputBufferManager("0", b200c_blocks::kHbmDomain);
        if (!_inMgr || (!_sink && !_outMgr)) throw Exception("Harness()", "block does not provide device buffer managers");
        BufferManagerArgs ia;
        ia.bufferSize = inBytes; ia.numBuffers = 1;
        _inMgr->init(ia);
        if (_sink) return;
        BufferManagerArgs oa;
        oa.bufferSize = outBytes; oa.numBuffers = 2;
        _outMgr->init(oa);
    }

    void activate()
    {
        _blk->_active = true;
        _blk->activate();
    }

    void postLabel(const Label &l) { _inLabels.push_back(l); }   // absolute input element index

    // host -> HBM ring; returns the number of elements accepted (ring may be full)
    size_t feed(const void *host, size_t elems)
    {
        InputPort *in = _blk->input(0);
        const size_t esz = in->dtype().size();
        const size_t room = _inMgr->front().length / esz;
        const size_t n = std::min(room, elems);
        if (n == 0) return 0;
        b200c_blocks::throwOnError(b200c_copy_h2d(_inMgr->front().as<void *>(), host, n * esz, _device, nullptr), "Harness::feed()");
        _inMgr->pop(n * esz);
        _totalFed += n;
        return n;
    }

    void run()
    {
        InputPort *in = _blk->input(0);
        if (_sink) { runSink(in); return; }
        OutputPort *out = _blk->output(0);
        const size_t isz = in->dtype().size(), osz = out->dtype().size();
        for (int guard = 0; guard < 1000000; guard++) {
            const BufferChunk rd = _inMgr->readable();
            in->_addr = rd.address;
            in->_bytes = rd.length / isz * isz;
            // the readable window runs past the end of the ring's first mapping: only the VMM double mapping keeps it contiguous
            if (rd.length && rd.address + rd.length > _inMgr->base() + _inMgr->capacity()) _seamWindows++;
            in->_labels.clear();
            for (const auto &l : _inLabels) {
                if (l.index >= _totalConsumed && l.index - _totalConsumed < in->elements()) {
                    Label r = l;
                    r.index = l.index - _totalConsumed;
                    in->_labels.push_back(r);
                }
            }
            if (_outMgr->empty()) break;
            out->_addr = _outMgr->front().address;
            out->_bytes = _outMgr->front().length / osz * osz;
            in->_pendingConsume = 0;
            out->_pendingProduce = 0;
            out->_posted.clear();
            // the scheduler does not call work() until the reserve is met
            if (in->elements() < std::max<size_t>(in->_reserve, 1)) break;

            _blk->work();
            _workCalls++;

            const size_t c = in->_pendingConsume, p = out->_pendingProduce;
            if (c > in->elements() || p > out->elements()) throw Exception("Harness::run()", "block over-consumed or over-produced");
            if (c) {
                // post-work: labels inside the consumed region are propagated, then dropped
                std::vector<Label> consumed, keep;
                for (const auto &l : in->_labels) if (l.index < c) consumed.push_back(l);
                in->_labels = consumed;
                _blk->propagateLabels(in);
                for (const auto &l : _inLabels) if (!(l.index >= _totalConsumed && l.index - _totalConsumed < c)) keep.push_back(l);
                _inLabels.swap(keep);
                _inMgr->push(c * isz);
                _totalConsumed += c;
                in->_totalConsumed = _totalConsumed;
            }
            for (auto &l : out->_posted) {
                l.index += _totalProduced;   // posted indices are relative to this call's first output
                _outLabels.push_back(l);
            }
            out->_posted.clear();
            if (p) {
                const size_t at = _collected.size();
                _collected.resize(at + p * osz);
                b200c_blocks::throwOnError(b200c_copy_d2h(_collected.data() + at, out->buffer().as<const void *>(), p * osz, _device, nullptr), "Harness::run()");
                b200c_blocks::throwOnError(b200c_stream_sync(_device, nullptr), "Harness::run()");
                _outMgr->pop(p * osz);
                _outMgr->push(p * osz);   // the "collector" is done with the slab
                _totalProduced += p;
            }
            if (c == 0 && p == 0) break;
        }
    }

    // Streaming throughput of the block the way a Pothos scheduler drives it between two DEVICE neighbours: the upstream
    // block has written `chunkElems` new elements into the HBM ring per round (here the ring is filled once with the host
    // pattern and the producer pointer simply advances over it: no host<->device copy inside the loop), work() runs while it
    // makes progress -- one kernel launch per call, K-1 history elements left in the ring -- and the downstream side
    // releases every output slab at once.  Returns wall seconds for `rounds` rounds, device idle on both sides.
    double streamBench(const void *hostPattern, size_t patternElems, size_t chunkElems, size_t rounds)
    {
        InputPort *in = _blk->input(0);
        OutputPort *out = _blk->output(0);
        const size_t isz = in->dtype().size(), osz = out->dtype().size();
        const size_t cap = _inMgr->capacity();
        if (_inMgr->readable().length != 0 || chunkElems * isz * 2 > cap) throw Exception("Harness::streamBench()", "ring must be empty and hold two chunks");
        // fill the whole ring with the pattern (repeated) once
        size_t filled = 0;
        while (filled < cap) {
            const size_t n = std::min(cap - filled, patternElems * isz);
            b200c_blocks::throwOnError(b200c_copy_h2d(reinterpret_cast<void *>(_inMgr->base() + filled), hostPattern, n, _device, nullptr), "Harness::streamBench()");
            filled += n;
        }
        b200c_blocks::throwOnError(b200c_stream_sync(_device, nullptr), "Harness::streamBench()");
        auto round = [&] {
            if (_inMgr->front().length < chunkElems * isz) throw Exception("Harness::streamBench()", "ring full: block is not consuming");
            _inMgr->pop(chunkElems * isz);
            for (int guard = 0; guard < 64; guard++) {
                const BufferChunk rd = _inMgr->readable();
                in->_addr = rd.address;
                in->_bytes = rd.length / isz * isz;
                if (rd.length && rd.address + rd.length > _inMgr->base() + cap) _seamWindows++;
                in->_labels.clear();
                out->_addr = _outMgr->front().address;
                out->_bytes = _outMgr->front().length / osz * osz;
                in->_pendingConsume = 0;
                out->_pendingProduce = 0;
                if (in->elements() < std::max<size_t>(in->_reserve, 1)) break;
                _blk->work();
                _workCalls++;
                const size_t c = in->_pendingConsume, p = out->_pendingProduce;
                if (c) { _inMgr->push(c * isz); _totalConsumed += c; }
                if (p) { _outMgr->pop(p * osz); _outMgr->push(p * osz); _totalProduced += p; }
                if (c == 0 && p == 0) break;
            }
        };
        for (int w = 0; w < 3; w++) round();
        b200c_blocks::throwOnError(b200c_stream_sync(_device, nullptr), "Harness::streamBench()");
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t r = 0; r < rounds; r++) round();
        b200c_blocks::throwOnError(b200c_stream_sync(_device, nullptr), "Harness::streamBench()");
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }

    // a source block: `nwork` calls of work(), each offered room for `elems` elements (0 = the whole slab)
    void runSource(size_t nwork, size_t elems)
    {
        OutputPort *out = _blk->output(0);
        const size_t osz = out->dtype().size();
        for (size_t k = 0; k < nwork; k++) {
            if (_outMgr->empty()) throw Exception("Harness::runSource()", "no output buffer");
            out->_addr = _outMgr->front().address;
            out->_bytes = _outMgr->front().length / osz * osz;
            if (elems && elems * osz < out->_bytes) out->_bytes = elems * osz;
            out->_pendingProduce = 0;
            _blk->work();
            _workCalls++;
            const size_t p = out->_pendingProduce;
            if (p > out->elements()) throw Exception("Harness::runSource()", "block over-produced");
            if (!p) continue;
            const size_t at = _collected.size();
            _collected.resize(at + p * osz);
            b200c_blocks::throwOnError(b200c_copy_d2h(_collected.data() + at, out->buffer().as<const void *>(), p * osz, _device, nullptr), "Harness::runSource()");
            b200c_blocks::throwOnError(b200c_stream_sync(_device, nullptr), "Harness::runSource()");
            _outMgr->pop(p * osz);
            _outMgr->push(p * osz);
            _totalProduced += p;
        }
    }

    // a sink block: work() while the reserve is met and elements are consumed
    void runSink(InputPort *in)
    {
        const size_t isz = in->dtype().size();
        for (int guard = 0; guard < 1000000; guard++) {
            const BufferChunk rd = _inMgr->readable();
            in->_addr = rd.address;
            in->_bytes = rd.length / isz * isz;
            in->_labels.clear();
            in->_pendingConsume = 0;
            if (in->elements() < std::max<size_t>(in->_reserve, 1)) break;
            _blk->work();
            _workCalls++;
            const size_t c = in->_pendingConsume;
            if (c > in->elements()) throw Exception("Harness::run()", "block over-consumed");
            if (c == 0) break;
            _inMgr->push(c * isz);
            _totalConsumed += c;
            in->_totalConsumed = _totalConsumed;
        }
    }

    size_t collect(void *host, size_t maxElems)
    {
        if (_sink) return 0;
        const size_t osz = _blk->output(0)->dtype().size();
        const size_t n = std::min(maxElems, _collected.size() / osz);
        std::memcpy(host, _collected.data(), n * osz);
        _collected.erase(_collected.begin(), _collected.begin() + n * osz);
        return n;
    }

    // The reference's test topology with host neighbours (filter/TestFIRFilter.cpp:49-51), as it has to be wired around the
    // device blocks:  feeder (host) -> /b200c/host_to_hbm -> `mid` (a device block) -> /b200c/hbm_to_host -> collector (host).
    // Buffers: the upstream block writes into the manager the DOWNSTREAM input asks for (mid's HBM ring, the d2h block's HBM
    // ring); the two host ends are plain host memory.  `chunk` elements are fed per round.  Returns elements collected.
    static size_t runHostChain(Block *mid, const char *in, size_t inElems, size_t chunk, char *out, size_t outCap,
                               unsigned long long *bridgeCalls)
    {
        const DType dt = mid->input(0)->dtype();
        std::unique_ptr<Block> h2d(BlockRegistry::make("/b200c/host_to_hbm", dt)), d2h(BlockRegistry::make("/b200c/hbm_to_host", dt));
        const size_t esz = dt.size();
        // the device blocks refuse to share their buffers with a host-memory neighbour
        bool refused = false;
        try { mid->getInputBufferManager("0", ""); } catch (const PortDomainError &) { refused = true; }
        if (!refused) throw Exception("runHostChain()", "device block accepted a host-domain upstream");
        refused = false;
        try { mid->getOutputBufferManager("0", ""); } catch (const PortDomainError &) { refused = true; }
        if (!refused) throw Exception("runHostChain()", "device block accepted a host-domain downstream");
        if (h2d->getInputBufferManager("0", "")) throw Exception("runHostChain()", "host_to_hbm should take the default host manager");
        auto ringA = std::dynamic_pointer_cast<b200c_blocks::DeviceCircularBufferManager>(mid->getInputBufferManager("0", h2d->output(0)->domain()));
        auto ringB = std::dynamic_pointer_cast<b200c_blocks::DeviceCircularBufferManager>(d2h->getInputBufferManager("0", mid->output(0)->domain()));
        if (!ringA || !ringB) throw Exception("runHostChain()", "no device ring between the device blocks");
        BufferManagerArgs ra;
        ra.bufferSize = std::max<size_t>(4 * chunk * esz, 1 << 21); ra.numBuffers = 1;
        ringA->init(ra);
        BufferManagerArgs rb = ra;
        rb.bufferSize = 4 * ra.bufferSize;     // room for interpolated output
        ringB->init(rb);
        for (Block *b : {h2d.get(), mid, d2h.get()}) { b->_active = true; b->activate(); }
        size_t fed = 0, got = 0;
        unsigned long long calls = 0;
        for (int guard = 0; guard < 1000000; guard++) {
            bool progress = false;
            // feeder -> host_to_hbm -> ring A
            {
                InputPort *ip = h2d->input(0); OutputPort *op = h2d->output(0);
                const size_t n = std::min(chunk, inElems - fed);
                ip->_addr = reinterpret_cast<size_t>(in + fed * esz); ip->_bytes = n * esz; ip->_labels.clear();
                op->_addr = ringA->front().address; op->_bytes = ringA->front().length / esz * esz;
                ip->_pendingConsume = 0; op->_pendingProduce = 0;
                if (n && op->_bytes) { h2d->work(); calls++; }
                if (ip->_pendingConsume != op->_pendingProduce) throw Exception("runHostChain()", "bridge must copy 1:1");
                fed += ip->_pendingConsume;
                ringA->pop(op->_pendingProduce * esz);
                progress |= ip->_pendingConsume != 0;
            }
            // ring A -> mid -> ring B
            for (;;) {
                InputPort *ip = mid->input(0); OutputPort *op = mid->output(0);
                const BufferChunk rd = ringA->readable();
                ip->_addr = rd.address; ip->_bytes = rd.length / esz * esz; ip->_labels.clear();
                op->_addr = ringB->front().address; op->_bytes = ringB->front().length / esz * esz;
                ip->_pendingConsume = 0; op->_pendingProduce = 0;
                if (ip->elements() < std::max<size_t>(ip->_reserve, 1) || op->_bytes == 0) break;
                mid->work();
                const size_t c = ip->_pendingConsume, p = op->_pendingProduce;
                ringA->push(c * esz);
                ringB->pop(p * esz);
                if (c == 0 && p == 0) break;
                progress = true;
            }
            // ring B -> hbm_to_host -> collector
            {
                InputPort *ip = d2h->input(0); OutputPort *op = d2h->output(0);
                const BufferChunk rd = ringB->readable();
                ip->_addr = rd.address; ip->_bytes = rd.length / esz * esz; ip->_labels.clear();
                op->_addr = reinterpret_cast<size_t>(out + got * esz); op->_bytes = (outCap - got) * esz;
                ip->_pendingConsume = 0; op->_pendingProduce = 0;
                if (ip->_bytes && op->_bytes) { d2h->work(); calls++; }
                ringB->push(ip->_pendingConsume * esz);
                got += op->_pendingProduce;
                progress |= op->_pendingProduce != 0;
            }
            if (!progress) break;
        }
        if (bridgeCalls) *bridgeCalls = calls;
        return got;
    }

    Block *block() { return _blk.get(); }
    const std::vector<Label> &outLabels() const { return _outLabels; }
    size_t pendingOutput() const { return _sink ? 0 : _collected.size() / _blk->output(0)->dtype().size(); }
    size_t reserve() const { return _blk->_inputs.at(0)._reserve; }
    unsigned long long totalConsumed() const { return _totalConsumed; }
    unsigned long long workCalls() const { return _workCalls; }
    unsigned long long totalProduced() const { return _totalProduced; }
    unsigned long long seamWindows() const { return _seamWindows; }
    size_t ringBytes() const { return _inMgr ? _inMgr->capacity() : 0; }
    std::string inputDomain() const { return _inMgr->domain(); }

private:
    std::unique_ptr<Block> _blk;
    int _device;
    bool _sink = false;
    std::shared_ptr<b200c_blocks::DeviceCircularBufferManager> _inMgr;
    BufferManager::Sptr _outMgr;
    std::vector<Label> _inLabels, _outLabels;
    std::vector<char> _collected;
    unsigned long long _totalFed = 0, _totalConsumed = 0, _totalProduced = 0, _workCalls = 0, _seamWindows = 0;
};

} // namespace Pothos

// ------------------------------------------------------------------ extern "C" test surface ---
static thread_local std::string g_blkErr;

template <typename F> static int guarded(F &&f)
{
    try { f(); return 0; }
    catch (const Pothos::InvalidArgumentException &e) { g_blkErr = e.what(); return -1; }
    catch (const Pothos::Exception &e) { g_blkErr = e.what(); return -2; }
    catch (const std::exception &e) { g_blkErr = e.what(); return -3; }
}

using Pothos::Harness;
using Pothos::Object;

extern "C" {

const char *b200c_blk_last_error(void) { return g_blkErr.c_str(); }

int b200c_blk_registry_has(const char *path) { return Pothos::BlockRegistry::doesBlockExist(path) ? 1 : 0; }

// factory: /comms/fir_filter(dtype, tapsType) when taps_type != NULL, else /comms/fft(dtype, numBins, inverse)
void *b200c_blk_make(const char *path, const char *dtype, const char *taps_type, size_t num_bins, int inverse, size_t in_bytes,
                     size_t out_bytes, int *status)
{
    Harness *h = nullptr;
    const int rc = guarded([&] {
        Pothos::Block *blk = taps_type ? Pothos::BlockRegistry::make(path, Pothos::DType(dtype), std::string(taps_type))
                                       : Pothos::BlockRegistry::make(path, Pothos::DType(dtype), num_bins, inverse != 0);
        const char *env = std::getenv("B200C_DEVICE");
        h = new Harness(blk, env ? std::atoi(env) : 0, in_bytes, out_bytes);
    });
    if (status) *status = rc;
    return h;
}

// factory (dtype) only: /comms/scale, /comms/rotate, /comms/signal_probe, /blocks/stream_probe
void *b200c_blk_make_dtype(const char *path, const char *dtype, size_t in_bytes, size_t out_bytes, int *status)
{
    Harness *h = nullptr;
    const int rc = guarded([&] {
        Pothos::Block *blk = Pothos::BlockRegistry::make(path, Pothos::DType(dtype));
        const char *env = std::getenv("B200C_DEVICE");
        h = new Harness(blk, env ? std::atoi(env) : 0, in_bytes, out_bytes);
    });
    if (status) *status = rc;
    return h;
}
// a call returning double or std::complex<double> (SignalProbe::value): out[0] = re, out[1] = im
int b200c_blk_get_complex(void *h, const char *name, double *out)
{
    return guarded([&] {
        const Object o = static_cast<Harness *>(h)->block()->call(name);
        if (o.type() == typeid(std::complex<double>)) { const auto v = o.extract<std::complex<double>>(); out[0] = v.real(); out[1] = v.imag(); }
        else { out[0] = o.convert<double>(); out[1] = 0.0; }
    });
}
// last payload of a signal carrying one double or std::complex<double> ("valueChanged")
int b200c_blk_last_signal_value(void *h, const char *signal, double *out, size_t *count)
{
    return guarded([&] {
        Pothos::Block *b = static_cast<Harness *>(h)->block();
        *count = b->signalCount(signal);
        out[0] = out[1] = 0.0;
        const auto *args = b->lastSignal(signal);
        if (!args || args->empty()) return;
        const Object &o = args->at(0);
        if (o.type() == typeid(std::complex<double>)) { const auto v = o.extract<std::complex<double>>(); out[0] = v.real(); out[1] = v.imag(); }
        else out[0] = o.convert<double>();
    });
}

void b200c_blk_destroy(void *h) { delete static_cast<Harness *>(h); }

int b200c_blk_call_taps(void *h, const char *name, const double *taps, size_t n, int is_complex)
{
    return guarded([&] {
        if (is_complex) {
            const auto *c = reinterpret_cast<const std::complex<double> *>(taps);
            static_cast<Harness *>(h)->block()->call(name, std::vector<std::complex<double>>(c, c + n));
        } else {
            static_cast<Harness *>(h)->block()->call(name, std::vector<double>(taps, taps + n));
        }
    });
}
int b200c_blk_call_size(void *h, const char *name, size_t v) { return guarded([&] { static_cast<Harness *>(h)->block()->call(name, v); }); }
int b200c_blk_call_bool(void *h, const char *name, int v) { return guarded([&] { static_cast<Harness *>(h)->block()->call(name, v != 0); }); }
int b200c_blk_call_string(void *h, const char *name, const char *v) { return guarded([&] { static_cast<Harness *>(h)->block()->call(name, std::string(v)); }); }

int b200c_blk_get_size(void *h, const char *name, size_t *out)
{
    return guarded([&] { *out = static_cast<Harness *>(h)->block()->call(name).convert<size_t>(); });
}
int b200c_blk_get_bool(void *h, const char *name, int *out)
{
    return guarded([&] { *out = static_cast<Harness *>(h)->block()->call(name).convert<bool>() ? 1 : 0; });
}
int b200c_blk_get_string(void *h, const char *name, char *buf, size_t cap)
{
    return guarded([&] {
        const std::string s = static_cast<Harness *>(h)->block()->call(name).convert<std::string>();
        std::snprintf(buf, cap, "%s", s.c_str());
    });
}
// getTaps(): writes up to cap doubles (interleaved when complex); *n = tap count
int b200c_blk_get_taps(void *h, double *buf, size_t cap, size_t *n, int *is_complex)
{
    return guarded([&] {
        const Object o = static_cast<Harness *>(h)->block()->call("getTaps");
        if (o.type() == typeid(std::vector<double>)) {
            const auto &v = o.extract<std::vector<double>>();
            *n = v.size(); *is_complex = 0;
            for (size_t i = 0; i < v.size() && i < cap; i++) buf[i] = v[i];
        } else {
            const auto &v = o.extract<std::vector<std::complex<double>>>();
            *n = v.size(); *is_complex = 1;
            for (size_t i = 0; i < v.size() && 2 * i + 1 < cap; i++) { buf[2 * i] = v[i].real(); buf[2 * i + 1] = v[i].imag(); }
        }
    });
}
// port-less host blocks: /comms/fir_designer, /comms/window_designer
void *b200c_blk_make_noargs(const char *path, int *status)
{
    Harness *h = nullptr;
    const int rc = guarded([&] { h = new Harness(Pothos::BlockRegistry::make(path), 0, 0, 0); });
    if (status) *status = rc;
    return h;
}
int b200c_blk_call_complex(void *h, const char *name, double re, double im)
{
    return guarded([&] { static_cast<Harness *>(h)->block()->call(name, std::complex<double>(re, im)); });
}
int b200c_blk_run_source(void *h, size_t nwork, size_t elems) { return guarded([&] { static_cast<Harness *>(h)->runSource(nwork, elems); }); }
int b200c_blk_call_double(void *h, const char *name, double v) { return guarded([&] { static_cast<Harness *>(h)->block()->call(name, v); }); }
int b200c_blk_get_double(void *h, const char *name, double *out)
{
    return guarded([&] { *out = static_cast<Harness *>(h)->block()->call(name).convert<double>(); });
}
int b200c_blk_get_doubles(void *h, const char *name, double *buf, size_t cap, size_t *n)
{
    return guarded([&] {
        const auto v = static_cast<Harness *>(h)->block()->call(name).convert<std::vector<double>>();
        *n = v.size();
        for (size_t i = 0; i < v.size() && i < cap; i++) buf[i] = v[i];
    });
}
// Topology::connect(src, signal, dst, slot)
int b200c_blk_connect_signal(void *src, const char *signal, void *dst, const char *slot)
{
    return guarded([&] { static_cast<Harness *>(src)->block()->connectSignal(signal, static_cast<Harness *>(dst)->block(), slot); });
}
// last emitted payload of a taps-carrying signal; *count = emissions so far (0: never emitted, nothing written)
int b200c_blk_last_signal(void *h, const char *signal, double *buf, size_t cap, size_t *n, int *is_complex, size_t *count)
{
    return guarded([&] {
        Pothos::Block *b = static_cast<Harness *>(h)->block();
        *count = b->signalCount(signal); *n = 0; *is_complex = 0;
        const auto *args = b->lastSignal(signal);
        if (!args || args->empty()) return;
        const Object &o = args->at(0);
        if (o.type() == typeid(std::vector<double>)) {
            const auto &v = o.extract<std::vector<double>>();
            *n = v.size();
            for (size_t i = 0; i < v.size() && i < cap; i++) buf[i] = v[i];
        } else {
            const auto &v = o.extract<std::vector<std::complex<double>>>();
            *n = v.size(); *is_complex = 1;
            for (size_t i = 0; i < v.size() && 2 * i + 1 < cap; i++) { buf[2 * i] = v[i].real(); buf[2 * i + 1] = v[i].imag(); }
        }
    });
}
int b200c_blk_has_signal(void *h, const char *name) { return static_cast<Harness *>(h)->block()->hasSignal(name) ? 1 : 0; }

int b200c_blk_has_call(void *h, const char *name) { return static_cast<Harness *>(h)->block()->hasCall(name) ? 1 : 0; }

int b200c_blk_activate(void *h) { return guarded([&] { static_cast<Harness *>(h)->activate(); }); }

// kind: 0 = no payload, 1 = size_t payload, 2 = double payload
int b200c_blk_post_label(void *h, const char *id, int kind, size_t sval, double dval, unsigned long long index, size_t width)
{
    return guarded([&] {
        Pothos::Label l;
        l.id = id; l.index = index; l.width = width;
        if (kind == 1) l.data = Object(sval);
        if (kind == 2) l.data = Object(dval);
        static_cast<Harness *>(h)->postLabel(l);
    });
}

long long b200c_blk_feed(void *h, const void *host, size_t elems)
{
    long long n = -1;
    const int rc = guarded([&] { n = (long long)static_cast<Harness *>(h)->feed(host, elems); });
    return rc ? rc : n;
}
int b200c_blk_run(void *h) { return guarded([&] { static_cast<Harness *>(h)->run(); }); }
long long b200c_blk_pending(void *h) { return (long long)static_cast<Harness *>(h)->pendingOutput(); }
long long b200c_blk_collect(void *h, void *host, size_t max_elems) { return (long long)static_cast<Harness *>(h)->collect(host, max_elems); }
size_t b200c_blk_reserve(void *h) { return static_cast<Harness *>(h)->reserve(); }
unsigned long long b200c_blk_total_consumed(void *h) { return static_cast<Harness *>(h)->totalConsumed(); }
unsigned long long b200c_blk_work_calls(void *h) { return static_cast<Harness *>(h)->workCalls(); }
unsigned long long b200c_blk_total_produced(void *h) { return static_cast<Harness *>(h)->totalProduced(); }
// number of work() calls whose readable window straddled the end of the ring's first mapping (base + bytes)
unsigned long long b200c_blk_seam_windows(void *h) { return static_cast<Harness *>(h)->seamWindows(); }
size_t b200c_blk_ring_bytes(void *h) { return static_cast<Harness *>(h)->ringBytes(); }
// feeder (host) -> /b200c/host_to_hbm -> the harness's block -> /b200c/hbm_to_host -> collector (host); returns elements collected
long long b200c_blk_run_host_chain(void *h, const void *in, size_t in_elems, size_t chunk, void *out, size_t out_cap, unsigned long long *bridge_calls)
{
    long long got = -1;
    const int rc = guarded([&] {
        got = (long long)Harness::runHostChain(static_cast<Harness *>(h)->block(), static_cast<const char *>(in), in_elems,
                                               chunk, static_cast<char *>(out), out_cap, bridge_calls);
    });
    return rc ? rc : got;
}
double b200c_blk_stream_bench(void *h, const void *pattern, size_t pattern_elems, size_t chunk_elems, size_t rounds)
{
    double secs = -1.0;
    const int rc = guarded([&] { secs = static_cast<Harness *>(h)->streamBench(pattern, pattern_elems, chunk_elems, rounds); });
    return rc ? (double)rc : secs;
}
int b200c_blk_input_domain(void *h, char *buf, size_t cap) { std::snprintf(buf, cap, "%s", static_cast<Harness *>(h)->inputDomain().c_str()); return 0; }

long long b200c_blk_num_out_labels(void *h) { return (long long)static_cast<Harness *>(h)->outLabels().size(); }
int b200c_blk_out_label(void *h, size_t i, char *id, size_t idcap, unsigned long long *index, size_t *width, int *kind, double *value)
{
    return guarded([&] {
        const Pothos::Label &l = static_cast<Harness *>(h)->outLabels().at(i);
        std::snprintf(id, idcap, "%s", l.id.c_str());
        *index = l.index; *width = l.width; *kind = 0; *value = 0;
        if (l.data) {
            if (l.data.type() == typeid(double)) { *kind = 2; *value = l.data.convert<double>(); }
            else if (l.data.canConvert(typeid(size_t))) { *kind = 1; *value = (double)l.data.convert<size_t>(); }
        }
    });
}

} // extern "C"
