// ktg_api.cu -- the extern "C" boundary declared in include/katome_gpu.h.
#include <fstream>
#include <memory>
#include <atomic>
#include <system_error>
#include <thread>

#include "builder.cuh"
#include "fastq_device.cuh"
#include "host_plan.h"
#include "host_reader.h"
#include "multi.cuh"

using namespace ktg;

struct ktg_builder {
    std::unique_ptr<BuilderBase> impl;   // one GPU
    std::unique_ptr<MultiBuilder> multi; // or ktg_config.n_devices > 1 (multi.cuh)
    uint32_t k = 0;
    // device staging for host batches: two buffers for the full chunks, four small ones for the
    // short chunks a large batch ends in (all of those are copied while the last flush runs)
    // staging buffers of the host batcher: up to N_FULL rotate under the full chunks, the short
    // chunks at the end of a large batch have one each
    static constexpr int N_FULL = 6, N_TAIL = 4, N_STAGE = N_FULL + N_TAIL;
    DeviceBuf st_bases[N_STAGE], st_offs[N_STAGE];
    cudaEvent_t st_free[N_STAGE] = {}; // staging buffer consumed by the compute stream
    cudaEvent_t st_ready[N_STAGE] = {};
};

extern "C" {

const char *ktg_last_error(void) { return last_error_ref().c_str(); }

int ktg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

int ktg_create(const ktg_config *cfg, ktg_builder **out) {
    if (!cfg || !out) return fail(KTG_ERR_INVALID, "null argument");
    *out = nullptr;
    if (cfg->abi_version != KTG_ABI_VERSION)
        return fail(KTG_ERR_INVALID, "ABI version %u, library is %u", cfg->abi_version, KTG_ABI_VERSION);
    // prelude.rs:35 asserts k > 1, compress.rs:19 needs len > 2; 64 is our key width
    if (cfg->k < 3 || cfg->k > 64) return fail(KTG_ERR_BAD_K, "k_mer_size %u outside 3..=64", cfg->k);
    if (cfg->world_size > 1 && cfg->rank >= cfg->world_size) return fail(KTG_ERR_INVALID, "rank >= world_size");
    if (cfg->sub_table_log2_bytes && (cfg->sub_table_log2_bytes < 16 || cfg->sub_table_log2_bytes > 34))
        return fail(KTG_ERR_INVALID, "sub_table_log2_bytes out of range");
    if (ktg_device_count() < 1) return fail(KTG_ERR_NO_DEVICE, "no CUDA device: this library has no CPU fallback");
    if (cfg->n_devices > 1) {
        std::unique_ptr<MultiBuilder> m(new MultiBuilder());
        KTG_TRY(m->init(*cfg));
        ktg_builder *b = new ktg_builder();
        b->multi = std::move(m);
        b->k = cfg->k;
        *out = b;
        return KTG_OK;
    }
    int dev = cfg->device;
    if (cfg->n_devices == 1 && cfg->device_ids) dev = cfg->device_ids[0];
    if (dev < 0) KTG_CUDA(cudaGetDevice(&dev));
    std::unique_ptr<BuilderBase> impl;
    if (cfg->k <= 32) impl.reset(new Builder<uint64_t>());
    else impl.reset(new Builder<u128>());
    impl->cfg = *cfg;
    impl->k = cfg->k;
    impl->rc = cfg->reverse_complement != 0;
    impl->device = dev;
    impl->stream = (cudaStream_t)cfg->stream;
    impl->prof.enabled = (cfg->flags & KTG_FLAG_PROFILE) != 0;
    KTG_TRY(impl->init());
    ktg_builder *b = new ktg_builder();
    b->impl = std::move(impl);
    b->k = cfg->k;
    *out = b;
    return KTG_OK;
}

void ktg_destroy(ktg_builder *b) {
    if (!b) return;
    if (b->multi) {
        delete b;
        return;
    }
    cudaSetDevice(b->impl->device);
    cudaStreamSynchronize(b->impl->stream);
    cudaStreamSynchronize(b->impl->copy_stream);
    for (int i = 0; i < ktg_builder::N_STAGE; ++i) {
        b->st_bases[i].release();
        b->st_offs[i].release();
        if (b->st_free[i]) cudaEventDestroy(b->st_free[i]);
        if (b->st_ready[i]) cudaEventDestroy(b->st_ready[i]);
    }
    delete b;
}

#define KTG_ENTER(b)                                                                           \
    if (!(b)) return fail(KTG_ERR_INVALID, "null builder");                                    \
    KTG_CUDA(cudaSetDevice((b)->multi ? (b)->multi->dev[0] : (b)->impl->device));
// entry points that address one shard's device memory: single-device handles only
#define KTG_SINGLE(b)                                                                          \
    if ((b)->multi) return fail(KTG_ERR_INVALID, "%s is not available on a multi-device handle", __func__);

int ktg_add_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets,
                         uint64_t n_reads, uint64_t total_bases, uint64_t *accepted_reads,
                         uint64_t *accepted_bytes) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    uint64_t r0 = 0, b0 = 0;
    if (accepted_reads || accepted_bytes) KTG_TRY(b->impl->read_counters(&r0, &b0));
    KTG_TRY(b->impl->ingest_device((const uint8_t *)d_bases, (const uint64_t *)d_offsets, n_reads, total_bases));
    if (accepted_reads || accepted_bytes) {
        uint64_t r1 = 0, b1 = 0;
        KTG_TRY(b->impl->read_counters(&r1, &b1));
        if (accepted_reads) *accepted_reads += r1 - r0;
        if (accepted_bytes) *accepted_bytes += b1 - b0;
    }
    return KTG_OK;
}

// Host batch: chunked H2D on the copy stream, double buffered against the
// kernels on the compute stream.
int ktg_add_reads(ktg_builder *b, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                  uint64_t *accepted_reads, uint64_t *accepted_bytes) {
    KTG_ENTER(b);
    if (n_reads == 0) return KTG_OK;
    if (!bases || !offsets) return fail(KTG_ERR_INVALID, "null argument");
    if (b->multi) return b->multi->add_reads(bases, offsets, n_reads, accepted_reads, accepted_bytes);
    BuilderBase *impl = b->impl.get();
    struct Paced { // input at PCIe pace for the duration of this call (the builder's flush cadence)
        BuilderBase *p;
        explicit Paced(BuilderBase *q) : p(q) { p->host_paced = true; }
        ~Paced() { p->host_paced = false; }
    } paced{impl};
    uint64_t r0c = 0, b0c = 0;
    if (accepted_reads || accepted_bytes) KTG_TRY(impl->read_counters(&r0c, &b0c));
    // bytes of bases per chunk: small enough that the H2D copy of chunk i+1 hides the kernels of
    // chunk i and only one chunk's kernels are exposed at the end
    const uint64_t CHUNK = (uint64_t)std::max(1, impl->tune.chunk_mb) << 20;
    for (int i = 0; i < ktg_builder::N_STAGE; ++i) {
        if (!b->st_free[i]) {
            KTG_CUDA(cudaEventCreateWithFlags(&b->st_free[i], cudaEventDisableTiming));
            KTG_CUDA(cudaEventCreateWithFlags(&b->st_ready[i], cudaEventDisableTiming));
        }
    }
    // where the batch is cut and after which chunks the stage is flushed: host_plan.h
    // (measured on C2, 8.3 ms of copies, before the eager page stage: the builder's own cadence, 4 flushes,
    // 12.55 ms per build, one flush at 73 % 12.8, at 58 % 11.8, two (44 %, 73 %) 11.9)
    std::vector<uint64_t> pcts;
    for (int v : {impl->tune.flush_pct, impl->tune.flush_pct2})
        if (v >= 1 && v <= 100) pcts.push_back((uint64_t)v);
    const ChunkPlan plan = plan_chunks(offsets, n_reads, CHUNK, pcts, impl->tune.taper != 0);
    const std::vector<uint64_t> &cut = plan.cut;
    const std::vector<char> &flush_here = plan.flush_here;
    const size_t n_chunks = plan.n_chunks(), tail_first = plan.tail_first;
    const bool large = plan.large;
    const uint64_t total = offsets[n_reads] - offsets[0];
    struct Hold { // the builder keeps its stage until flush_hint() / the end of this call
        BuilderBase *p;
        ~Hold() { p->hold_flush = false; p->call_keys_hint = 0; }
    } hold{impl};
    if (large) {
        impl->hold_flush = true;
        impl->call_keys_hint = total > n_reads * (uint64_t)(impl->k - 1) ? total - n_reads * (uint64_t)(impl->k - 1) : 0;
    }
    // staging buffer of chunk c: the full chunks rotate over n_full buffers, the short ones at the
    // end have one each.  Two are enough: a buffer is free as soon as its chunk is packed, and a large
    // call ends in short chunks with buffers of their own, so the copy engine does not wait for a
    // flush (measured on C2: 3, 4 or 6 buffers change nothing; KTG_STAGE_BUFS to try).
    const size_t n_full = std::min<size_t>(std::max<size_t>((size_t)impl->tune.stage_bufs, 2), ktg_builder::N_FULL);
    auto slot_of = [&](size_t c) -> int {
        return c >= tail_first ? ktg_builder::N_FULL + (int)std::min<size_t>(c - tail_first, ktg_builder::N_TAIL - 1)
                               : (int)(c % n_full);
    };
    // What the offsets on the host already tell about a chunk: one read length or ragged, and how
    // many windows at most; with that the chunk is queued without a device round trip, and the
    // offsets of a uniform chunk are generated on the device instead of copied (8 bytes per read:
    // 8 % of the PCIe traffic of 100 bp reads).
    // (the pass over the offsets takes ~0.6 ms of host time per 64 MiB of 100 bp reads; on the issuing
    // thread that made the host the pace setter of the whole call, so a helper thread runs ahead with it)
    std::vector<BatchHint> hints(n_chunks);
    std::atomic<size_t> hints_done{0};
    auto compute_hint = [&](size_t c) {
        BatchHint &h = hints[c];
        const uint64_t r = cut[c], r1 = cut[c + 1], kk = impl->k, len0 = offsets[r + 1] - offsets[r];
        uint64_t diff = 0; // one length?  (branch free; the window count needs its own pass only for a ragged chunk)
        for (uint64_t i = r; i < r1; ++i) diff |= (offsets[i + 1] - offsets[i]) ^ len0;
        const bool uniform = diff == 0;
        uint64_t wub = 0;
        if (uniform) wub = len0 >= kk ? (r1 - r) * (len0 - kk + 1) : 0;
        else
            for (uint64_t i = r; i < r1; ++i) {
                const uint64_t len = offsets[i + 1] - offsets[i];
                wub += len >= kk ? len - kk + 1 : 0;
            }
        h.ulen = (uniform && len0 <= 0xFFFFFFFFull) ? (uint32_t)len0 : 0;
        h.windows_ub = wub;
    };
    struct Helper { // joined on every way out of this function
        std::thread t;
        ~Helper() {
            if (t.joinable()) t.join();
        }
    } helper;
    if (n_chunks >= 3) {
        try {
            helper.t = std::thread([&] {
                for (size_t c = 0; c < n_chunks; ++c) {
                    compute_hint(c);
                    hints_done.store(c + 1, std::memory_order_release);
                }
            });
        } catch (const std::system_error &) { // no thread to be had: do it here
        }
    }
    auto hint_of = [&](size_t c) -> const BatchHint & {
        if (!helper.t.joinable()) {
            for (size_t d = hints_done.load(std::memory_order_relaxed); d <= c; ++d) {
                compute_hint(d);
                hints_done.store(d + 1, std::memory_order_relaxed);
            }
        }
        else
            while (hints_done.load(std::memory_order_acquire) <= c) std::this_thread::yield();
        return hints[c];
    };
    auto issue_copy = [&](size_t c) -> int {
        const uint64_t r = cut[c], r1 = cut[c + 1], nb = offsets[r1] - offsets[r], nr = r1 - r;
        const int s = slot_of(c);
        // the staging buffer may still be read by the kernels of two chunks ago
        KTG_CUDA(cudaStreamWaitEvent(impl->copy_stream, b->st_free[s], 0));
        if (b->st_bases[s].cap < nb + 64 || b->st_offs[s].cap < (nr + 1) * 8) {
            KTG_CUDA(cudaStreamSynchronize(impl->stream)); // reallocation frees the old buffer
            KTG_CUDA(cudaStreamSynchronize(impl->copy_stream));
            KTG_TRY(b->st_bases[s].ensure((s < ktg_builder::N_FULL ? std::max<uint64_t>(nb, CHUNK) : nb + nb / 4) + 64));
            KTG_TRY(b->st_offs[s].ensure((nr + 1) * 8));
        }
        KTG_CUDA(cudaMemcpyAsync(b->st_bases[s].p, bases + offsets[r], nb, cudaMemcpyHostToDevice, impl->copy_stream));
        const uint32_t ulen = hint_of(c).ulen;
        if (ulen) {
            const int g = (int)std::min<uint64_t>((nr + 1 + 255) / 256, 4096);
            fill_offsets_kernel<<<g, 256, 0, impl->copy_stream>>>((uint64_t *)b->st_offs[s].p, nr + 1, offsets[r], ulen);
            KTG_CUDA(cudaGetLastError());
        }
        else KTG_CUDA(cudaMemcpyAsync(b->st_offs[s].p, offsets + r, (nr + 1) * 8, cudaMemcpyHostToDevice, impl->copy_stream));
        KTG_CUDA(cudaEventRecord(b->st_ready[s], impl->copy_stream));
        trace("copy queued", c);
        return KTG_OK;
    };
    // The copy of chunk i+1 is queued BEFORE the kernels of chunk i are launched: ingest_device
    // may synchronise the compute stream (table sizing at a flush), and the copy engine must not
    // sit idle meanwhile.
    KTG_TRY(issue_copy(0));
    size_t issued = 1;
    for (size_t c = 0; c < n_chunks; ++c) {
        const size_t ahead = (tail_first != (size_t)-1 && c + 1 >= tail_first) ? n_chunks : std::min(c + n_full, n_chunks);
        for (; issued < ahead; ++issued) KTG_TRY(issue_copy(issued));
        const uint64_t r = cut[c], r1 = cut[c + 1], nb = offsets[r1] - offsets[r], nr = r1 - r;
        const int s = slot_of(c);
        const BatchHint hint = hint_of(c);
        KTG_CUDA(cudaStreamWaitEvent(impl->stream, b->st_ready[s], 0));
        // offsets stay absolute: bias the base pointer instead (pack kernel subtracts offsets[0])
        const uint8_t *d_bases = (const uint8_t *)b->st_bases[s].p - offsets[r];
        impl->hint_shift0 = (uint32_t)((uintptr_t)b->st_bases[s].p & 31);
        impl->input_consumed = b->st_free[s];
        int rc_ = impl->ingest_device(d_bases, (const uint64_t *)b->st_offs[s].p, nr, nb, &hint);
        impl->input_consumed = nullptr;
        KTG_TRY(rc_);
        if (flush_here[c]) KTG_TRY(impl->flush_hint());
    }
    if (accepted_reads || accepted_bytes) {
        uint64_t r1c = 0, b1c = 0;
        KTG_TRY(impl->read_counters(&r1c, &b1c));
        if (accepted_reads) *accepted_reads += r1c - r0c;
        if (accepted_bytes) *accepted_bytes += b1c - b0c;
    }
    return KTG_OK;
}

// ---- FASTQ parsed on the device (fastq_device.cuh): the host only moves raw file bytes ----------
namespace {

struct FastqDeviceParser {
    ktg_builder *b;
    BuilderBase *impl;
    size_t chunk;                     // raw bytes per chunk
    uint8_t *pinned[2] = {nullptr, nullptr};
    DeviceBuf raw[2], dense[2], offs[2], nl, bcount, bstart, sstart, slen, tmp, info, hdr, recidx, baseoff;
    cudaEvent_t copied[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
    bool used[2] = {false, false};

    FastqDeviceParser(ktg_builder *b_, size_t chunk_) : b(b_), impl(b_->impl.get()), chunk(chunk_) {}
    ~FastqDeviceParser() {
        cudaStreamSynchronize(impl->stream);
        cudaStreamSynchronize(impl->copy_stream);
        for (int i = 0; i < 2; ++i) {
            if (pinned[i]) cudaFreeHost(pinned[i]);
            if (copied[i]) cudaEventDestroy(copied[i]);
            if (consumed[i]) cudaEventDestroy(consumed[i]);
            raw[i].release(); dense[i].release(); offs[i].release();
        }
        nl.release(); bcount.release(); bstart.release(); sstart.release(); slen.release(); tmp.release(); info.release();
        hdr.release(); recidx.release(); baseoff.release();
    }
    int init() {
        for (int i = 0; i < 2; ++i) {
            KTG_CUDA(cudaEventCreateWithFlags(&copied[i], cudaEventDisableTiming));
            KTG_CUDA(cudaEventCreateWithFlags(&consumed[i], cudaEventDisableTiming));
            KTG_TRY(raw[i].ensure(chunk + 64));
        }
        KTG_TRY(info.ensure(sizeof(FastqChunkInfo)));
        return KTG_OK;
    }
    // the two page-locked buffers of the serial paths (the block reader brings its own)
    int ensure_pinned() {
        for (int i = 0; i < 2; ++i)
            if (!pinned[i]) KTG_CUDA(cudaHostAlloc((void **)&pinned[i], chunk + 64, cudaHostAllocDefault));
        return KTG_OK;
    }

    // newline index of n raw bytes on the device -> nl (positions), *n_lines
    int index_lines(const uint8_t *d_raw, size_t n, uint32_t *n_lines) {
        cudaStream_t st = impl->stream;
        const uint32_t n_blocks = (uint32_t)((n + FQ_BLOCK_BYTES - 1) / FQ_BLOCK_BYTES);
        KTG_TRY(bcount.ensure(((size_t)n_blocks + 1) * 4));
        KTG_TRY(bstart.ensure(((size_t)n_blocks + 1) * 4));
        KTG_CUDA(cudaMemsetAsync((uint32_t *)bcount.p + n_blocks, 0, 4, st));
        fq_count_kernel<<<n_blocks, 256, 0, st>>>(d_raw, n, (uint32_t *)bcount.p);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint32_t *)bcount.p, (uint32_t *)bstart.p, (int)n_blocks + 1, st);
        KTG_TRY(tmp.ensure(tb));
        KTG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, (uint32_t *)bcount.p, (uint32_t *)bstart.p, (int)n_blocks + 1, st));
        KTG_CUDA(cudaMemcpyAsync(n_lines, (uint32_t *)bstart.p + n_blocks, 4, cudaMemcpyDeviceToHost, st));
        KTG_CUDA(cudaStreamSynchronize(st));
        if (*n_lines) {
            KTG_TRY(nl.ensure((size_t)*n_lines * 4));
            fq_positions_kernel<<<n_blocks, 256, 0, st>>>(d_raw, n, (const uint32_t *)bstart.p, (uint32_t *)nl.p);
        }
        return KTG_OK;
    }

    // a record that does not fit the chunk: double the pinned and the raw buffers, keeping what
    // pinned[cur] holds (a FASTA record may be a whole chromosome)
    int grow(int cur, size_t keep) {
        if (chunk >= ((size_t)2 << 30)) return fail(KTG_ERR_BAD_RECORD, "a record is larger than %zu bytes", chunk);
        KTG_CUDA(cudaStreamSynchronize(impl->stream));
        KTG_CUDA(cudaStreamSynchronize(impl->copy_stream));
        const size_t bigger = chunk * 2;
        for (int i = 0; i < 2; ++i) {
            uint8_t *p = nullptr;
            KTG_CUDA(cudaHostAlloc((void **)&p, bigger + 64, cudaHostAllocDefault));
            if (i == cur && keep) memcpy(p, pinned[i], keep);
            cudaFreeHost(pinned[i]);
            pinned[i] = p;
            KTG_TRY(raw[i].ensure(bigger + 64));
        }
        chunk = bigger;
        return KTG_OK;
    }

    // One FASTA file.  Mirrors next_fasta of host_reader.h record for record.
    int parse_fasta(ReadFile &f) {
        KTG_TRY(ensure_pinned());
        cudaStream_t st = impl->stream;
        size_t carry = 0; // bytes of an unfinished record at the front of pinned[cur]
        int cur = 0;
        bool eof = false, first = true;
        while (!eof) {
            if (used[cur]) KTG_CUDA(cudaEventSynchronize(copied[cur]));
            if (carry >= chunk) KTG_TRY(grow(cur, carry));
            uint8_t *h = pinned[cur];
            size_t got = f.read_raw(h + carry, chunk - carry);
            size_t n = carry + got;
            eof = got < chunk - carry;
            if (eof && n && h[n - 1] != '\n') h[n++] = '\n'; // a last line without newline is a line
            if (n == 0) break;
            if (first && h[0] != '>') return fail(KTG_ERR_BAD_RECORD, "Expected > at record start.");
            first = false;
            if (used[cur]) KTG_CUDA(cudaStreamWaitEvent(impl->copy_stream, consumed[cur], 0));
            KTG_CUDA(cudaMemcpyAsync(raw[cur].p, h, n, cudaMemcpyHostToDevice, impl->copy_stream));
            KTG_CUDA(cudaEventRecord(copied[cur], impl->copy_stream));
            KTG_CUDA(cudaStreamWaitEvent(st, copied[cur], 0));
            used[cur] = true;
            const uint8_t *d_raw = (const uint8_t *)raw[cur].p;
            uint32_t n_lines = 0;
            KTG_TRY(index_lines(d_raw, n, &n_lines));
            // per-line facts and their scans
            KTG_TRY(hdr.ensure(((size_t)n_lines + 1) * 4));
            KTG_TRY(recidx.ensure(((size_t)n_lines + 1) * 4));
            KTG_TRY(slen.ensure(((size_t)n_lines + 1) * 8));
            KTG_TRY(baseoff.ensure(((size_t)n_lines + 1) * 8));
            const int lgrid = (int)std::min<uint64_t>(((uint64_t)n_lines + 256) / 256, 148 * 8);
            fa_lines_kernel<<<lgrid, 256, 0, st>>>(d_raw, (const uint32_t *)nl.p, n_lines, (uint32_t *)hdr.p, (uint64_t *)slen.p);
            size_t tb = 0, tb2 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb, (uint32_t *)hdr.p, (uint32_t *)recidx.p, (int)n_lines + 1, st);
            cub::DeviceScan::ExclusiveSum(nullptr, tb2, (uint64_t *)slen.p, (uint64_t *)baseoff.p, (int)n_lines + 1, st);
            KTG_TRY(tmp.ensure(std::max(tb, tb2)));
            KTG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, (uint32_t *)hdr.p, (uint32_t *)recidx.p, (int)n_lines + 1, st));
            KTG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, (uint64_t *)slen.p, (uint64_t *)baseoff.p, (int)n_lines + 1, st));
            uint32_t n_hdr = 0;
            KTG_CUDA(cudaMemcpyAsync(&n_hdr, (uint32_t *)recidx.p + n_lines, 4, cudaMemcpyDeviceToHost, st));
            KTG_CUDA(cudaStreamSynchronize(st));
            // the last record of a chunk may continue in the next one
            const uint64_t n_rec = eof ? n_hdr : (n_hdr ? n_hdr - 1 : 0);
            FastqChunkInfo ci{};
            ci.consumed = eof ? n : 0;
            if (n_rec) {
                KTG_TRY(offs[cur].ensure((n_rec + 1) * 8));
                FastqChunkInfo init{};
                init.bad_header = ~0ull;
                init.min_len = ~0ull;
                init.consumed = eof ? n : 0;
                KTG_CUDA(cudaMemcpyAsync(info.p, &init, sizeof init, cudaMemcpyHostToDevice, st));
                fa_offsets_kernel<<<lgrid, 256, 0, st>>>((const uint32_t *)nl.p, (const uint32_t *)hdr.p, (const uint32_t *)recidx.p,
                                                        (const uint64_t *)baseoff.p, n_lines, n_rec, (uint64_t *)offs[cur].p,
                                                        (FastqChunkInfo *)info.p);
                const int rgrid = (int)std::min<uint64_t>((n_rec + 255) / 256, 148 * 8);
                fa_hint_kernel<<<rgrid, 256, 0, st>>>((const uint64_t *)offs[cur].p, n_rec, impl->k, (FastqChunkInfo *)info.p);
                uint64_t total_bases = 0;
                KTG_CUDA(cudaMemcpyAsync(&ci, info.p, sizeof ci, cudaMemcpyDeviceToHost, st));
                KTG_CUDA(cudaMemcpyAsync(&total_bases, (uint64_t *)offs[cur].p + n_rec, 8, cudaMemcpyDeviceToHost, st));
                KTG_CUDA(cudaStreamSynchronize(st));
                KTG_TRY(dense[cur].ensure(total_bases + 64));
                fa_gather_kernel<<<148 * 8, 256, 0, st>>>(d_raw, (const uint32_t *)nl.p, (const uint32_t *)hdr.p, (const uint32_t *)recidx.p,
                                                         (const uint64_t *)baseoff.p, n_lines, n_rec, (uint8_t *)dense[cur].p);
                BatchHint hint;
                hint.ulen = (ci.min_len == ci.max_len && ci.max_len <= 0xFFFFFFFFull) ? (uint32_t)ci.max_len : 0;
                hint.windows_ub = ci.windows_ub;
                impl->hint_shift0 = (uint32_t)((uintptr_t)dense[cur].p & 31);
                impl->input_consumed = consumed[cur];
                int rc_ = impl->ingest_device((const uint8_t *)dense[cur].p, (const uint64_t *)offs[cur].p, n_rec, total_bases, &hint);
                impl->input_consumed = nullptr;
                KTG_TRY(rc_);
            }
            else KTG_CUDA(cudaEventRecord(consumed[cur], st));
            if (eof) break;
            // what follows the last complete record goes to the front of the other buffer
            const size_t done = n_rec ? (size_t)ci.consumed : 0, rest = n - done;
            if (rest >= chunk) { // one record fills the whole chunk: read on into a larger one
                KTG_TRY(grow(cur, n));
                carry = n;
                continue;
            }
            const int nxt = cur ^ 1;
            if (used[nxt]) KTG_CUDA(cudaEventSynchronize(copied[nxt]));
            memcpy(pinned[nxt], h + done, rest);
            carry = rest;
            cur = nxt;
        }
        return KTG_OK;
    }

    // One chunk of FASTQ text at h[0, n) (page-locked), starting at a record boundary: to the device,
    // records cut, complete records ingested.  *rest_at = where the unfinished record at its end starts.
    int fastq_chunk(const uint8_t *h, size_t n, int cur, bool eof, size_t *rest_at) {
        cudaStream_t st = impl->stream;
        // raw bytes to the device (the device buffer of two chunks ago has been consumed)
        if (used[cur]) KTG_CUDA(cudaStreamWaitEvent(impl->copy_stream, consumed[cur], 0));
        KTG_TRY(raw[cur].ensure(n + 64));
        KTG_CUDA(cudaMemcpyAsync(raw[cur].p, h, n, cudaMemcpyHostToDevice, impl->copy_stream));
        KTG_CUDA(cudaEventRecord(copied[cur], impl->copy_stream));
        KTG_CUDA(cudaStreamWaitEvent(st, copied[cur], 0));
        used[cur] = true;
        const uint8_t *d_raw = (const uint8_t *)raw[cur].p;
        uint32_t n_lines = 0;
        KTG_TRY(index_lines(d_raw, n, &n_lines));
        const uint64_t n_rec = n_lines / 4;
        FastqChunkInfo ci{};
        ci.bad_header = ~0ull;
        if (n_rec) {
            KTG_TRY(sstart.ensure(n_rec * 4));
            KTG_TRY(slen.ensure((n_rec + 1) * 8));
            KTG_TRY(offs[cur].ensure((n_rec + 1) * 8));
            FastqChunkInfo init{};
            init.bad_header = ~0ull;
            init.min_len = ~0ull;
            KTG_CUDA(cudaMemcpyAsync(info.p, &init, sizeof init, cudaMemcpyHostToDevice, st));
            KTG_CUDA(cudaMemsetAsync((uint64_t *)slen.p + n_rec, 0, 8, st));
            const int grid = (int)std::min<uint64_t>((n_rec + 255) / 256, 148 * 8);
            fq_records_kernel<<<grid, 256, 0, st>>>(d_raw, (const uint32_t *)nl.p, n_rec, impl->k, (uint32_t *)sstart.p,
                                                   (uint64_t *)slen.p, (FastqChunkInfo *)info.p);
            size_t tb2 = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, tb2, (uint64_t *)slen.p, (uint64_t *)offs[cur].p, (int)n_rec + 1, st);
            KTG_TRY(tmp.ensure(tb2));
            KTG_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb2, (uint64_t *)slen.p, (uint64_t *)offs[cur].p, (int)n_rec + 1, st));
            uint64_t total_bases = 0;
            KTG_CUDA(cudaMemcpyAsync(&ci, info.p, sizeof ci, cudaMemcpyDeviceToHost, st));
            KTG_CUDA(cudaMemcpyAsync(&total_bases, (uint64_t *)offs[cur].p + n_rec, 8, cudaMemcpyDeviceToHost, st));
            KTG_CUDA(cudaStreamSynchronize(st));
            if (ci.bad_header != ~0ull) return fail(KTG_ERR_BAD_RECORD, "Expected @ at record start.");
            KTG_TRY(dense[cur].ensure(total_bases + 64));
            fq_gather_kernel<<<148 * 8, 256, 0, st>>>(d_raw, (const uint32_t *)sstart.p, (const uint64_t *)offs[cur].p,
                                                     n_rec, (uint8_t *)dense[cur].p);
            BatchHint hint;
            hint.ulen = (ci.min_len == ci.max_len && ci.max_len <= 0xFFFFFFFFull) ? (uint32_t)ci.max_len : 0;
            hint.windows_ub = ci.windows_ub;
            impl->hint_shift0 = (uint32_t)((uintptr_t)dense[cur].p & 31);
            impl->input_consumed = consumed[cur]; // dense, offsets and raw of this slot are free after the pack
            int rc_ = impl->ingest_device((const uint8_t *)dense[cur].p, (const uint64_t *)offs[cur].p, n_rec, total_bases, &hint);
            impl->input_consumed = nullptr;
            KTG_TRY(rc_);
        }
        else KTG_CUDA(cudaEventRecord(consumed[cur], st));
        // what follows the last complete record
        const size_t done = n_rec ? (size_t)ci.consumed : 0;
        *rest_at = done;
        if (eof && n - done) { // lines of an unfinished record: the reader fails on its header or on its missing lines
            if (h[done] != '@') return fail(KTG_ERR_BAD_RECORD, "Expected @ at record start.");
            return fail(KTG_ERR_BAD_RECORD, "Incomplete record. Each FastQ record has to consist of 4 lines.");
        }
        return KTG_OK;
    }

    // One FASTQ file.  Mirrors next_fastq of host_reader.h record for record.  A regular file is read in
    // blocks by reader threads that run ahead (BlockReader); anything else serially.
    int parse(ReadFile &f) {
        const long long fsize = f.regular_size();
        if (fsize >= 0 && chunk >= (64u << 10)) {
            if (fsize == 0) return KTG_OK;
            BlockReader rd(f.fd(), (size_t)fsize, chunk, 4, 3);
            if (rd.init()) return parse_blocks(rd);
            // (no page-locked memory or no thread to be had: the serial path below)
        }
        KTG_TRY(ensure_pinned());
        size_t carry = 0; // bytes of an unfinished record at the front of pinned[cur]
        int cur = 0;
        bool eof = false;
        while (!eof) {
            uint8_t *h = pinned[cur];
            if (used[cur]) KTG_CUDA(cudaEventSynchronize(copied[cur])); // (the carry was written after this wait)
            size_t got = f.read_raw(h + carry, chunk - carry);
            size_t n = carry + got;
            eof = got < chunk - carry;
            if (eof && n && h[n - 1] != '\n') h[n++] = '\n'; // a last line without newline is a line
            if (n == 0) break;
            size_t done = 0;
            KTG_TRY(fastq_chunk(h, n, cur, eof, &done));
            if (eof) break;
            const size_t rest = n - done;
            if (rest >= chunk) return fail(KTG_ERR_BAD_RECORD, "a FASTQ record is larger than %zu bytes", chunk);
            const int nxt = cur ^ 1;
            if (used[nxt]) KTG_CUDA(cudaEventSynchronize(copied[nxt]));
            memcpy(pinned[nxt], h + done, rest);
            carry = rest;
            cur = nxt;
        }
        return KTG_OK;
    }

    int parse_blocks(BlockReader &rd) {
        std::vector<uint8_t> carry;
        const size_t nb = rd.n_blocks();
        for (size_t j = 0; j < nb; ++j) {
            uint8_t *p = nullptr;
            size_t got = 0;
            trace("block wanted", j);
            if (!rd.acquire(j, &p, &got)) return fail(KTG_ERR_IO, "reading the file failed");
            trace("block there", j);
            if (carry.size() > BlockReader::HEAD) return fail(KTG_ERR_BAD_RECORD, "a FASTQ record is larger than %zu bytes", BlockReader::HEAD);
            uint8_t *h = p - carry.size();
            if (!carry.empty()) memcpy(h, carry.data(), carry.size());
            size_t n = carry.size() + got;
            const bool eof = j + 1 == nb;
            if (eof && n && h[n - 1] != '\n') h[n++] = '\n'; // a last line without newline is a line
            size_t done = 0;
            if (n) KTG_TRY(fastq_chunk(h, n, (int)(j & 1), eof, &done));
            carry.assign(h + done, h + n);
            // the copy of this block to the device must have finished before its buffer is read into again
            KTG_CUDA(cudaEventSynchronize(copied[j & 1]));
            rd.release(j);
            trace("block done", j);
        }
        return KTG_OK;
    }
};

} // namespace

int ktg_create_from_files(ktg_builder *b, const char *const *paths, uint32_t n_paths,
                          int file_type, uint64_t *total_bytes) {
    KTG_ENTER(b);
    if (file_type != KTG_FASTQ && file_type != KTG_FASTA)
        return fail(KTG_ERR_INVALID, "unsupported input_file_type %d", file_type);
    uint64_t reads = 0, bytes = 0;
    // check_files + opening every file up front (builder.rs:57-77, 146-149)
    std::vector<std::unique_ptr<ReadFile>> files;
    for (uint32_t i = 0; i < n_paths; ++i) {
        std::unique_ptr<ReadFile> f(new ReadFile());
        std::string why;
        if (!f->open(paths[i], file_type == KTG_FASTA, &why)) return fail(KTG_ERR_IO, "%s", why.c_str());
        files.push_back(std::move(f));
    }
    // records are cut on the device (option host_parse keeps the host reader, its twin)
    trace("create_from_files", n_paths);
    if (!b->multi && !b->impl->tune.host_parse) {
        // small: the two pinned buffers are allocated per call (0.3 ms / MiB)
        const size_t chunk = (size_t)std::max(1, b->impl->tune.fastq_chunk_kb) << 10;
        FastqDeviceParser parser(b, chunk);
        KTG_TRY(parser.init());
        trace("parser ready");
        for (auto &f : files) KTG_TRY(file_type == KTG_FASTA ? parser.parse_fasta(*f) : parser.parse(*f));
        trace("files parsed");
        KTG_TRY(b->impl->read_counters(&reads, &bytes));
        if (total_bytes) *total_bytes = bytes;
        return ktg_finalize(b);
    }
    // Two pinned batches: the parser fills one while the copy / kernels of the other are in flight.
    ReadBatch batch[2];
    cudaEvent_t copied[2] = {nullptr, nullptr};
    int cur = 0, rc_ = KTG_OK;
    for (auto &f : files) {
        for (;;) {
            if (copied[cur]) { // the previous contents of this batch must have left the host
                cudaEventSynchronize(copied[cur]);
            }
            std::string why;
            int st = f->next_batch(&batch[cur], 64u << 20, &why);
            if (st < 0) {
                rc_ = fail(KTG_ERR_BAD_RECORD, "%s", why.c_str());
                break;
            }
            if (batch[cur].n_reads()) {
                rc_ = ktg_add_reads(b, batch[cur].bases, batch[cur].offsets.data(), batch[cur].n_reads(), nullptr, nullptr);
                if (rc_ != KTG_OK) break;
                if (!b->multi) { // (a multi-device add_reads returns with its copies done)
                    if (!copied[cur]) cudaEventCreateWithFlags(&copied[cur], cudaEventDisableTiming);
                    cudaEventRecord(copied[cur], b->impl->copy_stream);
                }
                cur ^= 1;
            }
            if (st == 0) break;
        }
        if (rc_ != KTG_OK) break;
    }
    if (!b->multi) cudaStreamSynchronize(b->impl->copy_stream); // the batches are freed below
    for (int i = 0; i < 2; ++i)
        if (copied[i]) cudaEventDestroy(copied[i]);
    KTG_TRY(rc_);
    KTG_TRY(b->multi ? b->multi->read_counters(&reads, &bytes) : b->impl->read_counters(&reads, &bytes));
    if (total_bytes) *total_bytes = bytes;
    return ktg_finalize(b);
}

// ---- BFCounter input --------------------------------------------------------
// Host arrays: n k-mers of exactly k ASCII bases each (contiguous) and their counts.
int ktg_add_weighted_kmers(ktg_builder *b, const uint8_t *kmers, const uint32_t *weights, uint64_t n,
                           uint32_t minimal_weight_threshold, uint64_t *accepted_kmers, uint64_t *accepted_bytes) {
    KTG_ENTER(b);
    if (n == 0) return KTG_OK;
    if (!kmers || !weights) return fail(KTG_ERR_INVALID, "null argument");
    KTG_SINGLE(b);
    BuilderBase *impl = b->impl.get();
    const uint64_t k = impl->k, step = std::max<uint64_t>(1, (64ull << 20) / k); // 64 MiB of bases at a time
    DeviceBuf d_k, d_w;
    int rc_ = KTG_OK;
    for (uint64_t i = 0; i < n && rc_ == KTG_OK; i += step) {
        const uint64_t m = std::min(step, n - i);
        uint64_t got = 0;
        if ((rc_ = d_k.ensure(m * k)) != KTG_OK || (rc_ = d_w.ensure(m * 4)) != KTG_OK) break;
        if (cudaMemcpyAsync(d_k.p, kmers + i * k, m * k, cudaMemcpyHostToDevice, impl->stream) != cudaSuccess ||
            cudaMemcpyAsync(d_w.p, weights + i, m * 4, cudaMemcpyHostToDevice, impl->stream) != cudaSuccess) {
            rc_ = fail(KTG_ERR_CUDA, "copying BFCounter k-mers failed: %s", cudaGetErrorString(cudaGetLastError()));
            break;
        }
        rc_ = impl->add_weighted_kmers((const uint8_t *)d_k.p, (const uint32_t *)d_w.p, m, minimal_weight_threshold, &got);
        if (accepted_kmers) *accepted_kmers += got;
        if (accepted_bytes) *accepted_bytes += got * k; // total += edge.len() (builder.rs:109)
    }
    cudaStreamSynchronize(impl->stream);
    d_k.release();
    d_w.release();
    return rc_;
}

// create_bfc (builder.rs:79-115): lines "<k-mer>\t<count>".  Errors mirror the reference's panics:
// unopenable file (:88-91), missing field / unparsable count (:100-105), k-mer shorter than k
// ("Read is too short!", pt_graph.rs:319).  A k-mer longer than k is rejected as a bad record
// (the reference would pack it as a longer edge against k-wide nodes).
int ktg_create_from_bfc_files(ktg_builder *b, const char *const *paths, uint32_t n_paths,
                              uint32_t minimal_weight_threshold, uint64_t *total_bytes) {
    KTG_ENTER(b);
    if (!paths && n_paths) return fail(KTG_ERR_INVALID, "null argument");
    KTG_SINGLE(b);
    const uint64_t k = b->impl->k;
    std::vector<FILE *> fs;
    auto close_all = [&]() { for (FILE *f : fs) if (f) fclose(f); };
    for (uint32_t i = 0; i < n_paths; ++i) {
        FILE *f = fopen(paths[i], "rb");
        if (!f) {
            close_all();
            return fail(KTG_ERR_IO, "Couldn't open all files: %s", paths[i]);
        }
        fs.push_back(f);
    }
    std::vector<uint8_t> kmers;
    std::vector<uint32_t> weights;
    uint64_t total = 0;
    int rc_ = KTG_OK;
    auto flush = [&]() -> int {
        if (weights.empty()) return KTG_OK;
        int r = ktg_add_weighted_kmers(b, kmers.data(), weights.data(), weights.size(), 0, nullptr, nullptr);
        kmers.clear();
        weights.clear();
        return r;
    };
    char line[4096];
    for (FILE *f : fs) {
        while (rc_ == KTG_OK && fgets(line, sizeof line, f)) {
            size_t n = strlen(line);
            while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
            char *tab = (char *)memchr(line, '\t', n);
            if (!tab) { rc_ = fail(KTG_ERR_BAD_RECORD, "BFCounter line without a count"); break; }
            char *endp = nullptr;
            const unsigned long long w = strtoull(tab + 1, &endp, 10);
            if (endp == tab + 1 || (*endp && *endp != '\t') || w > 0xFFFFFFFFull || tab[1] == '-' || tab[1] == '+') {
                rc_ = fail(KTG_ERR_BAD_RECORD, "Parse int error in a BFCounter line");
                break;
            }
            if ((uint32_t)w < minimal_weight_threshold) continue; // builder.rs:106-108, before total += len
            const uint64_t len = (uint64_t)(tab - line);
            if (len < k) { rc_ = fail(KTG_ERR_SHORT_READ, "Read is too short!"); break; }
            if (len != k) { rc_ = fail(KTG_ERR_BAD_RECORD, "BFCounter k-mer of %llu bases, k is %llu", (unsigned long long)len, (unsigned long long)k); break; }
            total += len;
            kmers.insert(kmers.end(), (const uint8_t *)line, (const uint8_t *)line + len);
            weights.push_back((uint32_t)w);
            if (weights.size() >= (1u << 20)) rc_ = flush();
        }
        if (rc_ != KTG_OK) break;
    }
    if (rc_ == KTG_OK) rc_ = flush();
    close_all();
    KTG_TRY(rc_);
    if (total_bytes) *total_bytes = total;
    return ktg_finalize(b);
}

int ktg_wait_input(ktg_builder *b) {
    KTG_ENTER(b);
    if (b->multi) return KTG_OK; // its ktg_add_reads returns with the copies done
    KTG_CUDA(cudaStreamSynchronize(b->impl->copy_stream));
    return KTG_OK;
}

int ktg_reset(ktg_builder *b) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->reset();
    return b->impl->reset();
}

int ktg_finalize(ktg_builder *b) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->finalize();
    return b->impl->finalize();
}

static int edge_stats_of(ktg_builder *b, uint32_t threshold, EdgeStats *es) {
    return b->multi ? b->multi->edge_stats(threshold, es) : b->impl->edge_stats(threshold, es);
}
static int node_stats_of(ktg_builder *b, NodeStats *ns) {
    return b->multi ? b->multi->node_stats(ns) : b->impl->node_stats(ns);
}

int ktg_counts(ktg_builder *b, uint64_t *nodes, uint64_t *edges) {
    KTG_ENTER(b);
    if (edges) {
        EdgeStats es;
        KTG_TRY(edge_stats_of(b, 0, &es));
        *edges = es.edges;
    }
    if (nodes) {
        NodeStats ns;
        KTG_TRY(node_stats_of(b, &ns));
        *nodes = ns.nodes;
    }
    return KTG_OK;
}

int ktg_collection_stats(ktg_builder *b, ktg_stats *out) {
    KTG_ENTER(b);
    if (!out) return fail(KTG_ERR_INVALID, "null argument");
    EdgeStats es;
    NodeStats ns;
    KTG_TRY(edge_stats_of(b, 0, &es));
    KTG_TRY(node_stats_of(b, &ns));
    out->node_count = ns.nodes;
    out->edge_count = es.edges;
    out->max_edge_weight = es.max_w;
    out->sum_edge_weight = es.sum_w;
    out->max_in_degree = ns.max_in;
    out->max_out_degree = ns.max_out;
    out->incoming_vert_count = ns.sources;
    out->outgoing_vert_count = ns.sinks;
    return KTG_OK;
}

int ktg_nodes_export_device(ktg_builder *b, void **d_keys, void **d_degrees, uint64_t *n, uint32_t *key_words) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_keys || !d_degrees || !n || !key_words) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->nodes_export(d_keys, d_degrees, n, key_words);
}

int ktg_nodes_stats_from_device(ktg_builder *b, const void *d_keys, const void *d_degrees, uint64_t n, ktg_stats *out) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!out) return fail(KTG_ERR_INVALID, "null argument");
    NodeStats ns;
    KTG_TRY(b->impl->nodes_stats_from(d_keys, d_degrees, n, &ns));
    memset(out, 0, sizeof *out);
    out->node_count = ns.nodes;
    out->max_in_degree = ns.max_in;
    out->max_out_degree = ns.max_out;
    out->incoming_vert_count = ns.sources;
    out->outgoing_vert_count = ns.sinks;
    return KTG_OK;
}

int ktg_edge_sums(ktg_builder *b, uint32_t threshold, uint64_t *sum_w, uint64_t *sum_w_below) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!sum_w || !sum_w_below) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->edge_sums(threshold, sum_w, sum_w_below);
}

int ktg_scale_weights(ktg_builder *b, double ratio, uint32_t threshold) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    return b->impl->scale_weights(ratio, threshold);
}

int ktg_remove_weak_edges(ktg_builder *b, uint32_t threshold) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->remove_weak_edges(threshold);
    return b->impl->remove_weak_edges(threshold);
}

int ktg_remove_single_vertices(ktg_builder *b) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->finalize();
    return b->impl->finalize(); // nodes are implicit: nothing to remove
}

int ktg_standardize_edges(ktg_builder *b, uint64_t genome_len, uint64_t k, uint32_t threshold) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->standardize(genome_len, k, threshold);
    return b->impl->standardize(genome_len, k, threshold);
}

int ktg_export_edges(ktg_builder *b, uint64_t *key_hi, uint64_t *key_lo, uint32_t *weight,
                     uint64_t cap, int sorted, uint64_t *n) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->export_edges(key_hi, key_lo, weight, cap, sorted, n);
    return b->impl->export_edges(key_hi, key_lo, weight, cap, sorted, n);
}

int ktg_export_graph(ktg_builder *b, uint64_t *node_hi, uint64_t *node_lo, uint64_t n_nodes, uint64_t *src,
                     uint64_t *dst, uint32_t *weight, uint8_t *edge_bytes, uint64_t n_edges) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->export_graph(node_hi, node_lo, n_nodes, src, dst, weight, edge_bytes, n_edges);
    return b->impl->export_graph(node_hi, node_lo, n_nodes, src, dst, weight, edge_bytes, n_edges);
}

int ktg_graph_prepare(ktg_builder *b, uint64_t *n_nodes, uint64_t *n_edges) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->graph_prepare(n_nodes, n_edges);
    KTG_TRY(b->impl->finalize());
    return b->impl->graph_prepare(n_nodes, n_edges);
}

int ktg_export_externals(ktg_builder *b, uint64_t *node_ids, uint8_t *kinds, uint64_t cap, uint64_t *n) {
    KTG_ENTER(b);
    if (b->multi) return b->multi->export_externals(node_ids, kinds, cap, n);
    return b->impl->export_externals(node_ids, kinds, cap, n);
}

uint32_t ktg_edge_record_bytes(const ktg_builder *b) { return b ? (b->k + 3) / 4 + 1 : 0; }

int ktg_digest(ktg_builder *b, uint64_t out[4]) {
    KTG_ENTER(b);
    EdgeStats es;
    KTG_TRY(edge_stats_of(b, 0, &es));
    out[0] = es.digest;
    out[1] = es.edges;
    out[2] = es.sum_w;
    out[3] = es.max_w;
    return KTG_OK;
}

uint32_t ktg_key_words(const ktg_builder *b) { return b && b->k > 32 ? 2u : 1u; }

uint32_t ktg_owner_of(const ktg_builder *b, uint64_t key_hi, uint64_t key_lo) {
    if (!b) return 0;
    if (b->multi) return b->multi->use_skm ? b->multi->sh[0]->skm_owner_of(key_hi, key_lo) : b->multi->sh[0]->owner_of(key_hi, key_lo);
    return b->impl->owner_of(key_hi, key_lo);
}

int ktg_partition_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets,
                               uint64_t n_reads, uint64_t total_bases, void **d_keys,
                               uint64_t *counts, uint64_t *accepted_reads,
                               uint64_t *accepted_bytes) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_keys || !counts) return fail(KTG_ERR_INVALID, "null argument");
    uint64_t r0 = 0, b0 = 0;
    if (accepted_reads || accepted_bytes) KTG_TRY(b->impl->read_counters(&r0, &b0));
    KTG_TRY(b->impl->partition_reads((const uint8_t *)d_bases, (const uint64_t *)d_offsets, n_reads, total_bases, d_keys, counts));
    if (accepted_reads || accepted_bytes) {
        uint64_t r1 = 0, b1 = 0;
        KTG_TRY(b->impl->read_counters(&r1, &b1));
        if (accepted_reads) *accepted_reads += r1 - r0;
        if (accepted_bytes) *accepted_bytes += b1 - b0;
    }
    return KTG_OK;
}

int ktg_insert_keys_device(ktg_builder *b, const void *d_keys, uint64_t n) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    return b->impl->insert_keys(d_keys, n);
}

int ktg_partition_keys_device(ktg_builder *b, const void *d_keys, uint64_t n, void **d_out, uint64_t *counts) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_out || !counts) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->partition_keys(d_keys, n, d_out, counts);
}

int ktg_mg_plan(ktg_builder *b, uint64_t max_windows, int *needs_realloc) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!needs_realloc) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_plan(max_windows, needs_realloc);
}

int ktg_mg_prepare(ktg_builder *b, uint64_t max_windows, void **rx_base, uint64_t *rx_bytes,
                   uint64_t *bucket_cap, uint32_t *n_sub) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!rx_base || !rx_bytes || !bucket_cap || !n_sub) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_prepare(max_windows, rx_base, rx_bytes, bucket_cap, n_sub);
}

int ktg_mg_scatter_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets, uint64_t n_reads,
                                uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                                void *send_stream, void **d_cursors) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!peer_rx || !d_cursors) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_scatter_reads((const uint8_t *)d_bases, (const uint64_t *)d_offsets, n_reads, total_bases,
                                     peer_rx, slot, first_of_batch, (cudaStream_t)send_stream, d_cursors);
}

int ktg_mg_insert_buckets(ktg_builder *b, const void *d_bucket_ends, uint64_t n_keys, uint32_t slot) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    return b->impl->mg_insert_buckets(d_bucket_ends, n_keys, slot);
}

int ktg_mg_direct_plan(ktg_builder *b, uint64_t max_windows, int *needs_realloc, uint32_t *n_sub, uint32_t *sub_log2) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!needs_realloc || !n_sub || !sub_log2) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mgd_plan(max_windows, needs_realloc, n_sub, sub_log2);
}

int ktg_mg_direct_prepare(ktg_builder *b, uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!rx_base || !rx_bytes || !bucket_cap) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mgd_prepare(max_windows, rx_base, rx_bytes, bucket_cap);
}

int ktg_mg_direct_scatter_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets, uint64_t n_reads,
                                       uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                                       void *send_stream, void **d_cursors) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!peer_rx || !d_cursors) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mgd_scatter_reads((const uint8_t *)d_bases, (const uint64_t *)d_offsets, n_reads, total_bases, peer_rx,
                                      slot, first_of_batch, (cudaStream_t)send_stream, d_cursors);
}

int ktg_mg_direct_insert(ktg_builder *b, const void *d_bucket_ends, uint64_t n_keys, uint32_t slot) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_bucket_ends) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mgd_insert(d_bucket_ends, n_keys, slot);
}

int ktg_mg_sketch(ktg_builder *b, void **d_regs, uint32_t *n_regs) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_regs || !n_regs) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_sketch(d_regs, n_regs);
}

int ktg_mg_merge_sketch(ktg_builder *b, const void *d_regs) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_regs) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_merge_sketch(d_regs);
}

int ktg_mg_spill(ktg_builder *b, void **d_keys, uint64_t *n) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_keys || !n) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_spill(d_keys, n);
}

int ktg_mg_insert_spill(ktg_builder *b, const void *d_keys, uint64_t n) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    return b->impl->mg_insert_spill(d_keys, n);
}

// ---- the same exchange in super-k-mer records (superkmer.cuh) ----
int ktg_mg_skm_supported(uint32_t k) { return ktg::skm_supported(k) ? 1 : 0; }

int ktg_mg_skm_plan(ktg_builder *b, uint64_t max_windows, int *needs_realloc) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!needs_realloc) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_skm_plan(max_windows, needs_realloc);
}

int ktg_mg_skm_prepare(ktg_builder *b, uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!rx_base || !rx_bytes || !bucket_cap) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_skm_prepare(max_windows, rx_base, rx_bytes, bucket_cap);
}

int ktg_mg_skm_scatter_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets, uint64_t n_reads,
                                    uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                                    void *send_stream, void **d_cursors, void **d_key_counts) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!peer_rx || !d_cursors || !d_key_counts) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_skm_scatter_reads((const uint8_t *)d_bases, (const uint64_t *)d_offsets, n_reads, total_bases,
                                         peer_rx, slot, first_of_batch, (cudaStream_t)send_stream, d_cursors,
                                         d_key_counts);
}

int ktg_mg_skm_insert_buckets(ktg_builder *b, const void *d_bucket_ends, uint64_t n_keys_ub, uint32_t slot) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_bucket_ends) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_skm_insert_buckets(d_bucket_ends, n_keys_ub, slot);
}

int ktg_mg_skm_spill(ktg_builder *b, void **d_records, uint64_t *n) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_records || !n) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_skm_spill(d_records, n);
}

int ktg_mg_skm_partition_records(ktg_builder *b, const void *d_records, uint64_t n, void **d_out, uint64_t *counts) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    if (!d_out || !counts) return fail(KTG_ERR_INVALID, "null argument");
    return b->impl->mg_skm_partition_records(d_records, n, d_out, counts);
}

int ktg_mg_skm_insert_records(ktg_builder *b, const void *d_records, uint64_t n) {
    KTG_ENTER(b);
    KTG_SINGLE(b);
    return b->impl->mg_skm_insert_records(d_records, n);
}

uint32_t ktg_mg_skm_owner_of(const ktg_builder *b, uint64_t key_hi, uint64_t key_lo) {
    if (!b || !b->impl) return 0;
    return b->impl->skm_owner_of(key_hi, key_lo);
}

// Host-only twin of the sending kernel's per-item logic (no device needed): cuts the windows
// of n_items work items into super-k-mer records.
int ktg_skm_items_host(const uint64_t *packed, const uint64_t *pos, const uint32_t *valid, uint64_t n_items, uint32_t k,
                       uint32_t world, uint64_t *records_lo_hi, uint64_t cap, uint64_t *n_records) {
    if (!packed || !pos || !valid || !records_lo_hi || !n_records) return fail(KTG_ERR_INVALID, "null argument");
    if (!ktg::skm_supported(k) || world == 0 || world > (uint32_t)ktg::MAX_P2P_WORLD)
        return fail(KTG_ERR_INVALID, "unsupported k or world");
    uint64_t n = 0;
    for (uint64_t i = 0; i < n_items; ++i) {
        ktg::u128 out[ktg::SKM_W];
        const uint32_t c = ktg::skm_item_host(packed, pos[i], valid[i] & 0xFFFFu, k, world, out);
        for (uint32_t j = 0; j < c; ++j, ++n) {
            if (n >= cap) return fail(KTG_ERR_INVALID, "record buffer too small");
            records_lo_hi[2 * n] = (uint64_t)out[j];
            records_lo_hi[2 * n + 1] = (uint64_t)(out[j] >> 64);
        }
    }
    *n_records = n;
    return KTG_OK;
}

uint32_t ktg_skm_owner_of_kmer(uint64_t kmer, uint32_t k, uint32_t world) {
    if (!ktg::skm_supported(k) || world == 0) return 0;
    return ktg::skm_owner(ktg::skm_minimizer_of_kmer(kmer, k), world);
}

int ktg_ipc_get_handle(const void *dev_ptr, uint8_t handle[64]) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    if (!dev_ptr || !handle) return fail(KTG_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    KTG_CUDA(cudaIpcGetMemHandle(&h, (void *)dev_ptr));
    memcpy(handle, &h, 64);
    return KTG_OK;
}

int ktg_ipc_open(const uint8_t handle[64], void **dev_ptr) {
    if (!dev_ptr || !handle) return fail(KTG_ERR_INVALID, "null argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    KTG_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return KTG_OK;
}

int ktg_ipc_close(void *dev_ptr) {
    if (!dev_ptr) return KTG_OK;
    KTG_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return KTG_OK;
}

int ktg_host_alloc(void **p, size_t bytes) {
    if (!p) return fail(KTG_ERR_INVALID, "null argument");
    KTG_CUDA(cudaHostAlloc(p, bytes, cudaHostAllocDefault));
    return KTG_OK;
}

void ktg_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int ktg_synth_reads_device(void *d_out, uint64_t seed_g, uint64_t genome_len, uint32_t read_len,
                           uint32_t err_ppm, uint64_t r0, uint64_t r1, void *stream) {
    if (!d_out || r1 < r0 || read_len == 0 || genome_len < read_len)
        return fail(KTG_ERR_INVALID, "bad synthetic read configuration");
    if (r1 == r0) return KTG_OK;
    const uint64_t thr = (uint64_t)err_ppm * 18446744073709ull; // err_ppm * floor(2^64 / 1e6)
    uint64_t work = (r1 - r0) * ((read_len + 3) / 4);
    int grid = (int)std::min<uint64_t>((work + 255) / 256, 148 * 16);
    synth_reads_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((uint8_t *)d_out, seed_g, genome_len, read_len, thr, r0, r1 - r0);
    KTG_CUDA(cudaGetLastError());
    return KTG_OK;
}

int ktg_random_access_probe(uint64_t bytes, uint64_t n_updates, uint32_t slot_bytes, float *ms) {
    if (!ms || (slot_bytes != 16 && slot_bytes != 32) || bytes < slot_bytes)
        return fail(KTG_ERR_INVALID, "bad probe configuration");
    void *p = nullptr;
    KTG_CUDA(cudaMalloc(&p, bytes));
    KTG_CUDA(cudaMemset(p, 0, bytes));
    unsigned long long *sink = nullptr;
    KTG_CUDA(cudaMalloc(&sink, 8));
    cudaEvent_t a, c;
    cudaEventCreate(&a);
    cudaEventCreate(&c);
    uint64_t n_slots = bytes / slot_bytes;
    int grid = 148 * 8;
    for (int it = 0; it < 2; ++it) { // first pass warms the cache / TLB
        cudaEventRecord(a);
        if (slot_bytes == 16) random_access_probe_kernel<16><<<grid, 256>>>((unsigned char *)p, n_slots, n_updates, sink);
        else random_access_probe_kernel<32><<<grid, 256>>>((unsigned char *)p, n_slots, n_updates, sink);
        cudaEventRecord(c);
    }
    cudaError_t e = cudaDeviceSynchronize();
    cudaEventElapsedTime(ms, a, c);
    cudaEventDestroy(a);
    cudaEventDestroy(c);
    cudaFree(p);
    cudaFree(sink);
    if (e != cudaSuccess) return fail(KTG_ERR_CUDA, "probe failed: %s", cudaGetErrorString(e));
    return KTG_OK;
}

// profile of a multi-device handle: the first shard's (every shard runs the same schedule)
static BuilderBase *profiled(ktg_builder *b) { return b->multi ? b->multi->sh[0].get() : b->impl.get(); }

int ktg_get_profile(ktg_builder *b, ktg_kernel_profile *out, uint32_t cap, uint32_t *n) {
    KTG_ENTER(b);
    KTG_CUDA(cudaStreamSynchronize(profiled(b)->stream));
#ifdef KTG_PHASE_TIMERS
    {
        unsigned long long h[8];
        cudaMemcpyFromSymbol(h, g_phase_cycles, sizeof h);
        fprintf(stderr, "[ktg phases] load %llu rank %llu scan %llu place %llu copy %llu\n", h[0], h[1], h[2], h[3], h[4]);
        memset(h, 0, sizeof h);
        cudaMemcpyToSymbol(g_phase_cycles, h, sizeof h);
    }
#endif
    profiled(b)->prof.resolve();
    const auto &es = profiled(b)->prof.entries;
    if (n) *n = (uint32_t)es.size();
    for (uint32_t i = 0; i < es.size() && i < cap; ++i) {
        memset(&out[i], 0, sizeof out[i]);
        strncpy(out[i].name, es[i].name.c_str(), sizeof(out[i].name) - 1);
        out[i].launches = es[i].launches;
        out[i].total_ms = es[i].ms;
        out[i].units = es[i].units;
    }
    return KTG_OK;
}

int ktg_reset_profile(ktg_builder *b) {
    KTG_ENTER(b);
    if (b->multi)
        return b->multi->parallel([&](uint32_t i) -> int {
            KTG_TRY(b->multi->sh[i]->sync_stream());
            b->multi->sh[i]->prof.reset();
            return KTG_OK;
        });
    KTG_CUDA(cudaStreamSynchronize(b->impl->stream));
    b->impl->prof.reset();
    return KTG_OK;
}

int ktg_item_reads(const uint32_t *items, uint64_t n, uint32_t items_per_read, uint32_t *out) {
    if (!items || !out || items_per_read == 0) return fail(KTG_ERR_INVALID, "bad request");
    ReadView v{};
    v.set_ipr(items_per_read);
    for (uint64_t i = 0; i < n; ++i) out[i] = v.ipr_magic ? div_magic(items[i], v.ipr_magic) : items[i];
    return KTG_OK;
}

int ktg_plan_chunks(const uint64_t *offsets, uint64_t n_reads, uint64_t chunk_bytes, const uint32_t *flush_pcts,
                    uint32_t n_pcts, uint64_t *cuts, uint8_t *flush_after, uint32_t cap, uint32_t *n_chunks,
                    int64_t *tail_first) {
    if (!offsets || !n_chunks || chunk_bytes == 0 || n_reads == 0) return fail(KTG_ERR_INVALID, "bad plan request");
    std::vector<uint64_t> pcts(flush_pcts, flush_pcts + (flush_pcts ? n_pcts : 0));
    const ChunkPlan p = plan_chunks(offsets, n_reads, chunk_bytes, pcts, true);
    *n_chunks = (uint32_t)p.n_chunks();
    if (tail_first) *tail_first = p.tail_first == (size_t)-1 ? -1 : (int64_t)p.tail_first;
    if (p.n_chunks() > cap) return fail(KTG_ERR_INVALID, "plan has %zu chunks, room for %u", p.n_chunks(), cap);
    for (size_t c = 0; c <= p.n_chunks(); ++c)
        if (cuts) cuts[c] = p.cut[c];
    for (size_t c = 0; c < p.n_chunks(); ++c)
        if (flush_after) flush_after[c] = (uint8_t)p.flush_here[c];
    return KTG_OK;
}

int ktg_host_parse_file(const char *path, int file_type, uint64_t batch_bytes, uint64_t *n_records,
                        uint64_t *total_bases, uint64_t *checksum) {
    if (!path || (file_type != KTG_FASTQ && file_type != KTG_FASTA)) return fail(KTG_ERR_INVALID, "bad argument");
    ReadFile f;
    std::string why;
    if (!f.open(path, file_type == KTG_FASTA, &why)) return fail(KTG_ERR_IO, "%s", why.c_str());
    ReadBatch batch;
    uint64_t n = 0, bases = 0, h = 0xcbf29ce484222325ull; // FNV-1a over every sequence followed by '\n'
    for (;;) {
        const int st = f.next_batch(&batch, batch_bytes ? batch_bytes : (64u << 20), &why);
        if (st < 0) return fail(KTG_ERR_BAD_RECORD, "%s", why.c_str());
        for (uint64_t r = 0; r < batch.n_reads(); ++r) {
            for (uint64_t i = batch.offsets[r]; i < batch.offsets[r + 1]; ++i) h = (h ^ batch.bases[i]) * 0x100000001b3ull;
            h = (h ^ (uint64_t)'\n') * 0x100000001b3ull;
        }
        n += batch.n_reads();
        bases += batch.size;
        if (st == 0) break;
    }
    if (n_records) *n_records = n;
    if (total_bases) *total_bases = bases;
    if (checksum) *checksum = h;
    return KTG_OK;
}

int ktg_set_profile(ktg_builder *b, int enabled) {
    KTG_ENTER(b);
    if (b->multi)
        return b->multi->parallel([&](uint32_t i) -> int {
            KTG_TRY(b->multi->sh[i]->sync_stream());
            b->multi->sh[i]->prof.enabled = enabled != 0;
            return KTG_OK;
        });
    KTG_CUDA(cudaStreamSynchronize(b->impl->stream));
    b->impl->prof.resolve(); // what was timed so far stays in the table
    b->impl->prof.enabled = enabled != 0;
    return KTG_OK;
}

static int set_option_on(BuilderBase *impl, const char *name, int64_t value) {
    Tuning &t = impl->tune;
    const struct { const char *n; int *p; } ints[] = {
        {"page_threads", &t.page_threads}, {"page_nbuf", &t.page_nbuf}, {"page_log2", &t.page_log2},
        {"l2s_variant", &t.l2s_variant}, {"l1_ctas", &t.l1_ctas}, {"l1_big", &t.l1_big}, {"p2p_ctas", &t.p2p_ctas},
        {"chunk_mb", &t.chunk_mb}, {"stage_bufs", &t.stage_bufs}, {"flush_pct", &t.flush_pct},
        {"flush_pct2", &t.flush_pct2}, {"taper", &t.taper}, {"eager_pages", &t.eager_pages},
        {"stage_factor_milli", &t.stage_factor_milli}, {"host_parse", &t.host_parse},
        {"fastq_chunk_kb", &t.fastq_chunk_kb}, {"mg_pad", &t.mg_pad}, {"mg_direct", &t.mg_direct}, {"trace", &t.trace},
    };
    if (!strcmp(name, "stage_max_keys")) {
        t.stage_max_keys = value;
        return KTG_OK;
    }
    for (const auto &o : ints)
        if (!strcmp(name, o.n)) {
            *o.p = (int)value;
            if (o.p == &t.trace) trace_enabled() = value != 0;
            return KTG_OK;
        }
    return fail(KTG_ERR_INVALID, "unknown option %s", name);
}

int ktg_set_option(ktg_builder *b, const char *name, int64_t value) {
    KTG_ENTER(b);
    if (!name) return fail(KTG_ERR_INVALID, "null argument");
    if (b->multi) return b->multi->set_option(name, value, set_option_on);
    return set_option_on(b->impl.get(), name, value);
}

int ktg_get_info(ktg_builder *b, ktg_info *out) {
    KTG_ENTER(b);
    if (!out) return fail(KTG_ERR_INVALID, "null argument");
    if (b->multi) return b->multi->info(out);
    return b->impl->info(out);
}

} // extern "C"
