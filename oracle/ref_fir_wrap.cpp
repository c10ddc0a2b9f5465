/*
 * TEST INFRASTRUCTURE ONLY.  extern "C" feeder/collector around the REFERENCE's own
 * /comms/fir_filter block: filter/FIRFilter.cpp is compiled UNMODIFIED where it lies
 * (#include below resolves through -I/root/reference/filter; nothing is copied into this repo)
 * against
 *   - oracle/ref_include/Pothos/Framework.hpp  -> the repo's restatement of the Pothos API subset
 *     the block touches (PothosCore is absent from this image), and
 *   - oracle/ref_include/Pothos/Util/QFormat.hpp, RECALLED (PothosCore's header is external to
 *     /root/reference): the single unpinned piece of the FIR reference build.
 * Everything that computes a sample -- work() :207-309 incl. the burst flush :263-272,
 * updateInternals() :327-354, the 18-row factory :369-384 -- is the reference's own object code.
 * The build renames the namespace (-DPothos=PothosRefHost) so that this library and the product's
 * block layer can live in one process without sharing inline statics (the block registry).
 * Output goes only to oracle/_ref/libfirref.so (git-ignored).
 *
 * What this file adds is what a Pothos::Topology with /blocks/feeder_source -> fir_filter ->
 * /blocks/collector_sink does around the block (filter/TestFIRFilter.cpp:19-53): present the
 * unconsumed input window and an output buffer, honour setReserve(), call work(), apply
 * consume()/produce(), drop consumed labels.
 */
#include "FIRFilter.cpp" /* the reference's source, found via -I$(REF)/filter */

#include <cstdio>
#include <memory>
#include <thread>

namespace Pothos {

/* make("circular") of filter/FIRFilter.cpp:198 on the host: a buffer of twice the window so that the
 * readable span (K-1 history + new elements) is always contiguous; compaction by memmove stands in for
 * PothosCore's doubly mapped ring. */
class HostCircularBufferManager : public BufferManager {
public:
    void init(const BufferManagerArgs &args) override
    {
        _size = args.bufferSize * args.numBuffers;
        _mem.assign(2 * _size, 0);
        _rd = _filled = 0;
        update();
    }
    bool empty() const override { return _filled == _size; }
    const BufferChunk &front() const override { return _front; }
    void pop(size_t numBytes) override { _filled += numBytes; update(); }
    void push(size_t numBytes) override
    {
        _rd += numBytes; _filled -= numBytes;
        if (_rd >= _size) { std::memmove(_mem.data(), _mem.data() + _rd, _filled); _rd = 0; }
        update();
    }
    BufferChunk readable() const { return BufferChunk(reinterpret_cast<size_t>(_mem.data()) + _rd, _filled); }

private:
    void update() { _front = BufferChunk(reinterpret_cast<size_t>(_mem.data()) + _rd + _filled, _size - _filled); }
    std::vector<char> _mem;
    size_t _size = 0, _rd = 0, _filled = 0;
    BufferChunk _front;
};

/* (the scheduler's init(args) with the topology's buffer size follows in Harness::stream) */
static const bool registeredCircular = (BufferManager::registerFactory("circular", [](const BufferManagerArgs &) {
    return BufferManager::Sptr(std::make_shared<HostCircularBufferManager>());
}), true);

/* the friend the shim's ports name: the scheduler's side of one block */
class Harness {
public:
    Harness(const char *dtype, const char *tapsType)
        : _blk(BlockRegistry::make("/comms/fir_filter", DType(std::string(dtype)), std::string(tapsType)))
    {
    }

    Block *block() { return _blk.get(); }

    void activate() { _blk->_active = true; _blk->activate(); }

    /* Stream `in` (in_elems elements) through the block.  in_chunk elements become available per
     * round (0: all at once), out_chunk output elements are offered per work() (0: all that is left).
     * frame_end != 0 posts a frame-end label on the last element (burst flush, :226-229,263-272). */
    void stream(const char *in, size_t in_elems, char *out, size_t out_cap, size_t in_chunk, size_t out_chunk, bool frame_end,
                size_t *consumed, size_t *produced, size_t *work_calls)
    {
        InputPort *ip = _blk->input(0);
        OutputPort *op = _blk->output(0);
        const size_t esz = ip->dtype().size();
        if (in_chunk == 0) in_chunk = in_elems;
        /* the window lives in the manager the block itself asks for (:196-199) */
        BufferManagerArgs args;
        args.bufferSize = std::max<size_t>((in_chunk + 8192) * esz, 1 << 16);
        args.numBuffers = 1;
        auto mgr = std::dynamic_pointer_cast<HostCircularBufferManager>(_blk->getInputBufferManager("0", ""));
        if (!mgr) throw Exception("ref Harness", "block did not return the host circular manager");
        mgr->init(args);
        size_t fed = 0, cons = 0, prod = 0, calls = 0;
        for (;;) {
            /* feeder: top the ring up by at most one chunk */
            size_t n = std::min(std::min(in_chunk, in_elems - fed), mgr->front().length / esz);
            if (n) { std::memcpy(mgr->front().as<void *>(), in + fed * esz, n * esz); mgr->pop(n * esz); fed += n; }
            bool progress = false;
            for (;;) {
                const BufferChunk rd = mgr->readable();
                ip->_addr = rd.address;
                ip->_bytes = rd.length / esz * esz;
                ip->_labels.clear();
                if (frame_end && fed == in_elems && in_elems > cons) ip->_labels.push_back(Label("eob", true, in_elems - 1 - cons));
                size_t room = out_cap - prod;
                if (out_chunk) room = std::min(room, out_chunk);
                op->_addr = reinterpret_cast<size_t>(out + prod * esz);
                op->_bytes = room * esz;
                ip->_pendingConsume = 0;
                op->_pendingProduce = 0;
                if (ip->elements() < std::max<size_t>(ip->_reserve, 1)) break; /* the scheduler waits for the reserve */
                if (room == 0) break;
                _blk->work();
                calls++;
                const size_t c = ip->_pendingConsume, p = op->_pendingProduce;
                if (c > ip->elements() || p > room) throw Exception("ref Harness", "block over-consumed or over-produced");
                mgr->push(c * esz);
                cons += c;
                prod += p;
                if (c == 0 && p == 0) break;
                progress = true;
            }
            if (n == 0 && !progress) break; /* nothing new to offer and the block is idle (or stalled, :278) */
        }
        *consumed = cons; *produced = prod; *work_calls = calls;
    }

    /* exactly one work() over a caller-owned window (K-1 history + new), as the product's b200c_fir_run sees it */
    void work_once(const char *in, size_t in_elems, char *out, size_t out_cap, size_t *consumed, size_t *produced)
    {
        InputPort *ip = _blk->input(0);
        OutputPort *op = _blk->output(0);
        const size_t esz = ip->dtype().size();
        ip->_addr = reinterpret_cast<size_t>(in);
        ip->_bytes = in_elems * esz;
        ip->_labels.clear();
        op->_addr = reinterpret_cast<size_t>(out);
        op->_bytes = out_cap * esz;
        ip->_pendingConsume = 0;
        op->_pendingProduce = 0;
        _blk->work();
        *consumed = ip->_pendingConsume;
        *produced = op->_pendingProduce;
    }

    size_t reserve() { return _blk->input(0)->_reserve; }

private:
    std::unique_ptr<Block> _blk;
};

} // namespace Pothos

#define FIRREF_API extern "C" __attribute__((visibility("default")))

static thread_local std::string g_err;

static void configure(Pothos::Harness &h, const char *tapsType, const double *taps, size_t ntaps, size_t decim, size_t interp, bool frame_end)
{
    if (std::string(tapsType) == "COMPLEX") {
        std::vector<std::complex<double>> t(ntaps);
        for (size_t i = 0; i < ntaps; i++) t[i] = std::complex<double>(taps[2 * i], taps[2 * i + 1]);
        h.block()->call("setTaps", t);
    } else {
        h.block()->call("setTaps", std::vector<double>(taps, taps + ntaps));
    }
    h.block()->call("setDecimation", decim);
    h.block()->call("setInterpolation", interp);
    if (frame_end) h.block()->call("setFrameEndId", std::string("eob"));
    h.activate();
}

FIRREF_API const char *firref_last_error(void) { return g_err.c_str(); }

/* Stream a whole input through ONE reference block instance; see Harness::stream().  Returns 0, or -1 with
 * firref_last_error() holding the reference's exception text (unsupported type row, empty taps, zero rate). */
FIRREF_API int firref_stream(const char *dtype, const char *tapsType, const double *taps, size_t ntaps, size_t decim, size_t interp,
                             const void *in, size_t in_elems, void *out, size_t out_cap, size_t in_chunk, size_t out_chunk,
                             int frame_end, size_t *consumed, size_t *produced, size_t *work_calls)
{
    try {
        Pothos::Harness h(dtype, tapsType);
        configure(h, tapsType, taps, ntaps, decim, interp, frame_end != 0);
        h.stream(static_cast<const char *>(in), in_elems, static_cast<char *>(out), out_cap, in_chunk, out_chunk, frame_end != 0,
                 consumed, produced, work_calls);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

/* One work() call on a caller-owned window. */
FIRREF_API int firref_work(const char *dtype, const char *tapsType, const double *taps, size_t ntaps, size_t decim, size_t interp,
                           const void *in, size_t in_elems, void *out, size_t out_cap, size_t *consumed, size_t *produced)
{
    try {
        Pothos::Harness h(dtype, tapsType);
        configure(h, tapsType, taps, ntaps, decim, interp, false);
        h.work_once(static_cast<const char *>(in), in_elems, static_cast<char *>(out), out_cap, consumed, produced);
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

/* All-host-cores CPU baseline: `nthreads` independent block instances (the reference runs one actor per block),
 * thread t filtering its own contiguous segment of `seg_elems` elements (K-1 history first) in one work() call.
 * in/out hold nthreads segments back to back; out segments are `out_seg_cap` elements apart. */
FIRREF_API int firref_work_mt(int nthreads, const char *dtype, const char *tapsType, const double *taps, size_t ntaps, size_t decim,
                              size_t interp, const void *in, size_t seg_elems, void *out, size_t out_seg_cap, size_t *consumed,
                              size_t *produced)
{
    try {
        if (nthreads < 1) nthreads = 1;
        std::vector<std::unique_ptr<Pothos::Harness>> hs;
        for (int t = 0; t < nthreads; t++) {
            hs.emplace_back(new Pothos::Harness(dtype, tapsType));
            configure(*hs.back(), tapsType, taps, ntaps, decim, interp, false);
        }
        const size_t esz = hs[0]->block()->input(0)->dtype().size();
        std::vector<size_t> c(nthreads), p(nthreads);
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; t++)
            th.emplace_back([&, t] {
                hs[t]->work_once(static_cast<const char *>(in) + size_t(t) * seg_elems * esz, seg_elems,
                                 static_cast<char *>(out) + size_t(t) * out_seg_cap * esz, out_seg_cap, &c[t], &p[t]);
            });
        for (auto &x : th) x.join();
        *consumed = *produced = 0;
        for (int t = 0; t < nthreads; t++) { *consumed += c[t]; *produced += p[t]; }
        return 0;
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

/* getters of the reference block after configuration, for the host-logic tests: K is not exposed by the block,
 * but _inputRequire = M + K - 1 shows up as the reserve it sets when starved (:248-252). */
FIRREF_API long firref_input_require(const char *dtype, const char *tapsType, const double *taps, size_t ntaps, size_t decim, size_t interp)
{
    try {
        Pothos::Harness h(dtype, tapsType);
        configure(h, tapsType, taps, ntaps, decim, interp, false);
        /* one element available: below any requirement >= 2 -> the block sets its reserve */
        std::vector<char> one(64, 0), out(64, 0);
        size_t c, p;
        h.work_once(one.data(), 1, out.data(), 0, &c, &p);
        return (long)std::max<size_t>(h.reserve(), 1);
    } catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}
