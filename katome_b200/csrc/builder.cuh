// builder.cuh -- host-side state of one GIR build on one GPU: table sizing and
// growth, the per-batch kernel schedule, profiling.  Included by ktg_api.cu.
#pragma once
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <vector>

#include "../../include/katome_gpu.h"
#include "kernels.cuh"
#include "superkmer.cuh"

namespace ktg {

// ------------------------------------------------------------------ errors
inline std::string &last_error_ref() {
    static thread_local std::string s;
    return s;
}
inline int fail(int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_ref() = buf;
    return code;
}
#define KTG_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess)                                                                 \
            return ::ktg::fail(KTG_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                               cudaGetErrorString(e_), __FILE__, __LINE__);                    \
    } while (0)
#define KTG_TRY(expr)                                                                          \
    do {                                                                                       \
        int ktg_try_rc__ = (expr);                                                             \
        if (ktg_try_rc__ != KTG_OK) return ktg_try_rc__;                                       \
    } while (0)

} // namespace ktg
#include "export.cuh"
namespace ktg {

// host-side timeline for tuning (KTG_TRACE=1): label + microseconds since the first event
inline bool &trace_enabled() {
    static bool on = false;
    return on;
}
inline void trace(const char *label, uint64_t v = 0) {
    if (!trace_enabled()) return;
    static timespec t0{};
    timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    if (t0.tv_sec == 0 && t0.tv_nsec == 0) t0 = t;
    fprintf(stderr, "[ktg %9.1f us] %s %llu\n", (t.tv_sec - t0.tv_sec) * 1e6 + (t.tv_nsec - t0.tv_nsec) * 1e-3, label,
            (unsigned long long)v);
}

// grow-only device buffer
struct DeviceBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap) return KTG_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess)
            return fail(KTG_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
        cap = want;
        return KTG_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
};

// ---------------------------------------------------------------- profiling
struct Profiler {
    struct Entry {
        std::string name;
        uint64_t launches = 0, units = 0;
        double ms = 0;
    };
    struct Pending {
        cudaEvent_t a, b;
        int entry;
    };
    bool enabled = false;
    std::vector<Entry> entries;
    std::vector<Pending> pending;
    std::vector<cudaEvent_t> pool;
    uint64_t total_launches = 0;

    int entry_of(const char *name) {
        for (size_t i = 0; i < entries.size(); ++i)
            if (entries[i].name == name) return (int)i;
        entries.push_back(Entry());
        entries.back().name = name;
        return (int)entries.size() - 1;
    }
    cudaEvent_t get_event() {
        if (!pool.empty()) {
            cudaEvent_t e = pool.back();
            pool.pop_back();
            return e;
        }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
    void begin(const char *name, uint64_t units, cudaStream_t s) {
        ++total_launches;
        if (!enabled) return;
        Pending p;
        p.entry = entry_of(name);
        entries[p.entry].launches++;
        entries[p.entry].units += units;
        p.a = get_event();
        p.b = get_event();
        cudaEventRecord(p.a, s);
        pending.push_back(p);
    }
    void end(cudaStream_t s) {
        if (!enabled) return;
        cudaEventRecord(pending.back().b, s);
    }
    void resolve() { // caller has synchronised the stream
        for (Pending &p : pending) {
            float ms = 0;
            if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) entries[p.entry].ms += ms;
            pool.push_back(p.a);
            pool.push_back(p.b);
        }
        pending.clear();
    }
    void reset() {
        resolve();
        entries.clear();
        total_launches = 0;
    }
    ~Profiler() {
        for (Pending &p : pending) {
            cudaEventDestroy(p.a);
            cudaEventDestroy(p.b);
        }
        for (cudaEvent_t e : pool) cudaEventDestroy(e);
    }
};

struct DeviceProps {
    int sms = 148;
    size_t l2_bytes = 126u << 20;
};

constexpr uint32_t MAX_BINS = 2048;      // sub-tables (smem bins of the scatter kernel)
constexpr uint32_t MAX_PAGES_PER_SUB = 2048; // smem bins of the level-2 scatter
constexpr double LOAD_MAX = 0.70;        // grow before a batch could exceed this
constexpr double LOAD_TARGET = 0.50;     // load right after sizing / growing
constexpr uint64_t OVF_CAP = 1u << 20;   // replay buffer entries

// What the host batcher already knows about a batch from its host-side offsets.  With it a
// batch is queued without any host<->device round trip (the counters are read at finalize).
struct BatchHint {
    uint32_t ulen = 0;       // every read has this length (0: ragged)
    uint64_t windows_ub = 0; // windows if every read is accepted
};

// Options a caller may change through ktg_set_option (tests and measurements; the defaults are the
// measured winners).  Nothing in the library reads the environment.
struct Tuning {
    int page_threads = 0;    // update_pages: threads per CTA (0: the default of the key width)
    int page_nbuf = 0;       // update_pages: 1 = two CTAs per SM, 2 = one pipelined CTA per SM (0: default)
    int page_log2 = 0;       // slots per page, log2 (0: PageGeom)
    int l2s_variant = -1;    // level-2 scatter geometry (-1: the default of the key width)
    int l1_ctas = 0;         // cap on the level-1 scatter's CTAs per SM (0: occupancy)
    int l1_big = -1;         // level-1 scatter with the big tile (two extraction passes): -1 = when a thread of the one-pass kernel would have more than one bin, 0 / 1 = never / always
    int p2p_ctas = 0;        // same for the fused exchange kernels
    int chunk_mb = 64;       // host batcher: bytes of bases per chunk
    int stage_bufs = 2;      // host batcher: staging buffers under the full chunks
    int flush_pct = 55;      // host batcher: flush on the way after this share of a large call (0: never)
    int flush_pct2 = 0;      //   and a second one
    int taper = 1;           // host batcher: a large call ends in short chunks
    int eager_pages = 1;     // host batcher: level-2 scatter chunk by chunk
    int stage_factor_milli = 0; // keys per flush in thousandths of the capacity (0: 750 host paced, 3000 resident)
    long long stage_max_keys = 0; // most keys one stage may hold (0: what memory allows)
    int host_parse = 0;      // ktg_create_from_files: FASTQ / FASTA records cut by the host reader
    int fastq_chunk_kb = 8 << 10; // device parser: raw bytes per chunk (page-locked memory costs ~1 ms per MiB to allocate:
                                  // four 16 MiB blocks were 74 ms of a 157 ms Build::create from a 950 MB FASTQ file)
    int mg_pad = 1;          // fused exchange: runs padded to 128-byte lines
    int mg_direct = -1;      // multi-device handle: the direct exchange (sender bins by owner and sub-table) where the shards agree on
                             // their geometry: -1 = below 4 devices, 0 = never, 1 = whenever possible
    int trace = 0;           // host timeline on stderr
};

// ------------------------------------------------------------ abstract base
struct BuilderBase {
    Tuning tune;
    ktg_config cfg{};
    uint32_t k = 0;
    bool rc = false;
    int device = 0;
    cudaStream_t stream = nullptr, copy_stream = nullptr;
    bool own_stream = false;
    DeviceProps props;
    Profiler prof;
    int deferred_error = KTG_OK;
    uint32_t hint_shift0 = 0; // flat position of the first base of a hinted batch (set by the host batcher)
    // recorded on the compute stream as soon as the pack kernel has consumed the caller's
    // read buffer (the host batcher reuses its staging buffer then, not a whole flush later)
    cudaEvent_t input_consumed = nullptr;
    uint64_t windows_inserted = 0;
    uint32_t grow_events = 0;
    // set by the host batcher for the duration of one large ktg_add_reads call: the keys the call
    // will offer, and "keep staging until flush_hint() or the stage is full"
    uint64_t call_keys_hint = 0;
    bool hold_flush = false;
    // set by the host batcher for the duration of any ktg_add_reads call: input arrives at PCIe pace,
    // so flushes hide behind the following copies and should be small; device-resident input is
    // better served by as few sweeps of the table as memory allows
    bool host_paced = false;

    virtual ~BuilderBase() {}
    virtual int init() = 0;
    virtual int ingest_device(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                              uint64_t total_bases, const BatchHint *hint = nullptr) = 0;
    virtual int read_counters(uint64_t *reads, uint64_t *bytes) = 0;
    virtual int finalize() = 0;
    // the host batcher knows that little input is left: a good moment to empty the stage
    virtual int flush_hint() = 0;
    virtual int reset() = 0;
    virtual int edge_stats(uint32_t threshold, EdgeStats *out) = 0;
    virtual int node_stats(NodeStats *out) = 0;
    virtual int nodes_export(void **d_keys, void **d_deg, uint64_t *n, uint32_t *key_words) = 0;
    virtual int nodes_stats_from(const void *d_keys, const void *d_deg, uint64_t n, NodeStats *out) = 0;
    virtual int edge_sums(uint32_t threshold, uint64_t *sum_w, uint64_t *sum_below) = 0;
    virtual int scale_weights(double p, uint32_t t) = 0;
    virtual int remove_weak_edges(uint32_t t) = 0;
    virtual int standardize(uint64_t G, uint64_t k_, uint32_t t) = 0;
    virtual int export_edges(uint64_t *hi, uint64_t *lo, uint32_t *w, uint64_t cap, int sorted,
                             uint64_t *n) = 0;
    virtual int export_graph(uint64_t *node_hi, uint64_t *node_lo, uint64_t n_nodes, uint64_t *src, uint64_t *dst,
                             uint32_t *weight, uint8_t *edge_bytes, uint64_t n_edges) = 0;
    // the export in steps, for a sharded handle (multi.cuh)
    virtual int compact_edges_into(uint64_t *d_hi, uint64_t *d_lo, uint32_t *d_w, uint64_t cap) = 0;
    virtual int edges_to_host(uint64_t *d_hi, uint64_t *d_lo, uint32_t *d_w, uint64_t ne, int sorted, uint64_t *hi,
                              uint64_t *lo, uint32_t *w, uint64_t cap) = 0;
    virtual int graph_build(Scratch &from, uint64_t *d_ehi, uint64_t *d_elo, uint32_t *d_w, uint64_t ne) = 0;
    virtual int graph_prepare(uint64_t *n_nodes, uint64_t *n_edges) = 0;
    virtual bool graph_ready() const = 0;
    virtual void touch() = 0;
    virtual int graph_to_host(uint64_t *node_hi, uint64_t *node_lo, uint64_t n_nodes, uint64_t *src, uint64_t *dst,
                              uint32_t *weight, uint8_t *edge_bytes, uint64_t n_edges) = 0;
    virtual int externals_to_host(uint64_t *ids, uint8_t *kinds, uint64_t cap, uint64_t *n_out) = 0;
    virtual int export_externals(uint64_t *ids, uint8_t *kinds, uint64_t cap, uint64_t *n_out) = 0;
    virtual int sync_stream() = 0;
    virtual int partition_reads(const uint8_t *d_bases, const uint64_t *d_offsets,
                                uint64_t n_reads, uint64_t total_bases, void **d_keys,
                                uint64_t *counts) = 0;
    virtual int insert_keys(const void *d_keys, uint64_t n) = 0;
    virtual int add_weighted_kmers(const uint8_t *d_kmers, const uint32_t *d_weights, uint64_t n, uint32_t threshold,
                                   uint64_t *accepted) = 0;
    virtual int partition_keys(const void *d_keys, uint64_t n, void **d_out, uint64_t *counts) = 0;
    virtual int mg_plan(uint64_t max_windows, int *needs_realloc) = 0;
    virtual int mg_prepare(uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap,
                           uint32_t *n_sub) = 0;
    virtual int mg_insert_spill(const void *d_keys, uint64_t n) = 0;
    virtual int mg_scatter_reads(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                                 uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                                 cudaStream_t send_stream, void **d_cursors, const BatchHint *hint = nullptr) = 0;
    virtual int mg_insert_buckets(const void *d_bucket_ends, uint64_t n_keys, uint32_t slot) = 0;
    // the direct exchange: the sender partitions by (owner, sub-table), the owner goes straight to the page level
    virtual int mgd_plan(uint64_t max_windows, int *needs_realloc, uint32_t *n_sub, uint32_t *sub_log2) = 0;
    virtual int mgd_prepare(uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap) = 0;
    virtual int mgd_scatter_reads(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads, uint64_t total_bases,
                                  void *const *peer_rx, uint32_t slot, int first_of_batch, cudaStream_t send_stream,
                                  void **d_cursors, const BatchHint *hint = nullptr) = 0;
    virtual int mgd_insert(const void *d_bucket_ends, uint64_t n_keys, uint32_t slot) = 0;
    virtual int mg_sketch(void **d_regs, uint32_t *n_regs) = 0;
    virtual int mg_merge_sketch(const void *d_regs) = 0;
    virtual int mg_spill(void **d_keys, uint64_t *n) = 0;
    // the same exchange in super-k-mer records (superkmer.cuh); 23 <= k <= 31 only
    virtual int mg_skm_plan(uint64_t max_windows, int *needs_realloc) = 0;
    virtual int mg_skm_prepare(uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap) = 0;
    virtual int mg_skm_scatter_reads(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                                     uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                                     cudaStream_t send_stream, void **d_cursors, void **d_key_counts,
                                     const BatchHint *hint = nullptr) = 0;
    virtual int mg_skm_insert_buckets(const void *d_bucket_ends, uint64_t n_keys_ub, uint32_t slot) = 0;
    virtual int mg_skm_spill(void **d_records, uint64_t *n) = 0;
    virtual int mg_skm_partition_records(const void *d_records, uint64_t n, void **d_out, uint64_t *counts) = 0;
    virtual int mg_skm_insert_records(const void *d_records, uint64_t n) = 0;
    virtual uint32_t skm_owner_of(uint64_t hi, uint64_t lo) = 0;
    virtual uint32_t owner_of(uint64_t hi, uint64_t lo) = 0;
    virtual int info(ktg_info *out) = 0;
};

template <class F> inline int grid_for(F kernel, int block, size_t smem, const DeviceProps &p) {
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, block, smem);
    if (per_sm < 1) per_sm = 1;
    return p.sms * per_sm; // a whole number of CTAs per SM: one full wave, grid-stride inside
}

// ---------------------------------------------------------------- the builder
template <class K> struct Builder : BuilderBase {
    typedef KeyTraits<K> T;

    Table<K> tab{};
    bool fresh = true;        // the table is logically empty and its memory undefined (see ensure_init)
    uint32_t page_updates = 0;
    uint64_t occupied_ub = 0; // upper bound on occupied slots
    uint64_t hll_base = 0;    // exact occupancy when the sketch was (re)started
    bool sketch_complete = true; // the sketch covers every key inserted since hll_base
    DeviceBuf b_hll;
    DeviceBuf b_spill;
    DeviceBuf b_packed, b_keys, b_keys2, b_hist, b_ovf_keys, b_ovf_inc, b_small;
    DeviceBuf b_pkeys, b_pcur; // level-2 (page) buckets and their cursors
    PackCounters *d_ctr = nullptr;        // accumulates over the whole build
    unsigned long long *d_ovf_count = nullptr;
    unsigned long long *d_scratch = nullptr; // 16 u64 of scratch (stats, cursors)
    unsigned long long *d_lost = nullptr;    // keys dropped because a spill list overflowed (voids the build)
    unsigned long long *d_stage_spill = nullptr; // cursor of the stage's spill list (level-1 and page overflow)
    // lazily built node ((k-1)-mer) table
    bool nodes_valid = false;
    NodeStats node_cache{};
    // The graph Convert::create_from loads, kept on the device between ktg_graph_prepare (which reports its
    // sizes) and ktg_export_graph / ktg_export_externals (which copy it out); any change of the table drops it.
    struct GraphCache {
        Scratch sc;
        KeyArr edges{nullptr, nullptr}, nodes{nullptr, nullptr};
        uint32_t *w = nullptr;
        uint64_t *src = nullptr, *dst = nullptr, ne = 0, nn = 0;
        bool valid = false;
        void clear() {
            Scratch empty;
            std::swap(sc.ptrs, empty.ptrs); // frees what the cache held
            valid = false;
        }
    } graph;
    void touch() override { // the table changed
        nodes_valid = false;
        if (graph.valid) graph.clear();
    }

    ~Builder() override {
        cudaSetDevice(device);
        if (stream) cudaStreamSynchronize(stream);
        if (tab.base) cudaFree(tab.base);
        b_empty_page.release();
        b_packed.release(); b_bad.release(); b_valid.release(); b_wstart.release(); b_keys.release(); b_keys2.release();
        b_hist.release(); b_hll.release(); b_spill.release(); b_ovf_keys.release(); b_ovf_inc.release(); b_small.release();
        b_pkeys.release(); b_pcur.release(); b_stage_cur.release();
        b_rx.release(); b_mg_cur.release(); b_mg_spill.release(); b_skm_cnt.release(); b_skm_part.release(); b_bfc_ctr.release(); b_node_keys.release(); b_node_deg.release();
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (own_stream && stream) cudaStreamDestroy(stream);
    }

    // ---- table geometry ------------------------------------------------------
    // balance == true (no explicit sub-table size): sub-tables start at 16 MiB (they stay L2
    // resident for the atomic path) and double until there are no more sub-tables (level-1 bins)
    // than pages per sub-table (level-2 bins).  A table of P pages then costs two scatter passes
    // with about sqrt(P) bins each instead of one pass with P / 128 bins, whose runs would shrink
    // to a few keys (C3: 632 sub-tables made the level-1 scatter 3x slower per key).
    // world > 1: this is one shard of `world`; the sender of the direct exchange bins by (owner, sub-table),
    // so it is world x n_sub that must stay within what one scatter pass handles well
    static void geometry(uint64_t need_slots, uint32_t sub_log2_bytes, bool balance, uint32_t world, uint32_t *n_sub,
                         uint32_t *sub_log2) {
        uint32_t slot_log2 = sizeof(K) == 8 ? 4 : 5; // sub-table sizes are quoted for 16 / 32-byte slots (12 / 20 now)
        uint32_t sl = sub_log2_bytes - slot_log2; // slots per sub-table (log2)
        if (balance) {
            while (sl < 28 && sl > PageGeom<K>::LOG2) {
                const uint64_t ns = (need_slots + (1ull << sl) - 1) >> sl;
                if (ns * world <= 128 || ns * world <= (1ull << (sl - PageGeom<K>::LOG2))) break;
                ++sl;
            }
        }
        if (need_slots < 1024) need_slots = 1024;
        if (need_slots <= (1ull << sl)) {
            uint32_t l = 10;
            while ((1ull << l) < need_slots) ++l;
            *n_sub = 1;
            *sub_log2 = l;
            return;
        }
        uint64_t ns = (need_slots + (1ull << sl) - 1) >> sl;
        while (ns > MAX_BINS) {
            ++sl;
            ns = (need_slots + (1ull << sl) - 1) >> sl;
        }
        *n_sub = (uint32_t)ns;
        *sub_log2 = sl;
    }

    // do_init == false leaves the memory undefined: the caller marks the table `fresh`
    int alloc_table(uint64_t need_slots, Table<K> *out, bool do_init) {
        Table<K> t = tab;
        uint32_t slb = cfg.sub_table_log2_bytes ? cfg.sub_table_log2_bytes : 24;
        geometry(need_slots, slb, cfg.sub_table_log2_bytes == 0, std::max<uint32_t>(1, tab.world), &t.n_sub, &t.sub_log2);
        t.sub_mask = (uint32_t)((1ull << t.sub_log2) - 1);
        uint32_t pl = PageGeom<K>::LOG2;
        if (tune.page_log2) pl = std::min<uint32_t>(pl, (uint32_t)std::max(8, tune.page_log2));
        t.page_log2 = std::min<uint32_t>(pl, t.sub_log2);
        t.page_mask = (1u << t.page_log2) - 1;
        t.max_probe = std::min<uint32_t>(1u << t.page_log2, 2048);
        size_t bytes = t.bytes();
        void *p = nullptr;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(KTG_ERR_TABLE_FULL, "cannot allocate a %zu byte table: %s", bytes,
                        cudaGetErrorString(e));
        }
        t.base = (unsigned char *)p;
        if (do_init) {
            prof.begin("init_table", t.capacity() + 1, stream);
            init_table_kernel<K><<<props.sms * 8, 256, 0, stream>>>(t);
            prof.end(stream);
        }
        else KTG_TRY(init_special_slot(t));
        KTG_TRY(make_empty_page(t));
        *out = t;
        return KTG_OK;
    }

    // one page of empty slots: what the page update loads instead of a page of a fresh table
    DeviceBuf b_empty_page;
    uint32_t empty_page_log2 = 0;
    int make_empty_page(const Table<K> &t) {
        if (b_empty_page.p && empty_page_log2 == t.page_log2) return KTG_OK;
        KTG_TRY(b_empty_page.ensure(t.page_bytes()));
        KTG_CUDA(cudaMemsetAsync(b_empty_page.p, 0xFF, (size_t)sizeof(K) << t.page_log2, stream));
        KTG_CUDA(cudaMemsetAsync((char *)b_empty_page.p + ((size_t)sizeof(K) << t.page_log2), 0, (size_t)4 << t.page_log2, stream));
        empty_page_log2 = t.page_log2;
        return KTG_OK;
    }

    // the one slot past the end (all-ones key at full width) is always kept defined
    int init_special_slot(const Table<K> &t) {
        KTG_CUDA(cudaMemsetAsync(t.special_w(), 0, 16, stream));
        return KTG_OK;
    }

    // A table that was just allocated / reset is only marked `fresh`; the page
    // update initialises pages in shared memory as it sweeps, every other user of
    // the table memory (L2-atomic inserts, scans) materialises the empty table first.
    int ensure_init() {
        if (!fresh) return KTG_OK;
        prof.begin("init_table", tab.capacity() + 1, stream);
        init_table_kernel<K><<<props.sms * 8, 256, 0, stream>>>(tab);
        prof.end(stream);
        fresh = false;
        return KTG_OK;
    }

    int init() override {
        KTG_CUDA(cudaSetDevice(device));
        cudaDeviceProp dp;
        KTG_CUDA(cudaGetDeviceProperties(&dp, device));
        props.sms = dp.multiProcessorCount;
        props.l2_bytes = (size_t)dp.l2CacheSize;
        if (!stream) {
            KTG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
            own_stream = true;
        }
        KTG_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        KTG_TRY(b_small.ensure(4096));
        KTG_CUDA(cudaMemsetAsync(b_small.p, 0, 4096, stream));
        d_ctr = (PackCounters *)b_small.p;
        d_ovf_count = (unsigned long long *)((char *)b_small.p + 256);
        d_scratch = (unsigned long long *)((char *)b_small.p + 512);
        d_lost = (unsigned long long *)((char *)b_small.p + 1024);
        d_stage_spill = d_lost + 1;
        KTG_TRY(b_hll.ensure(HLL_M * 4));
        KTG_CUDA(cudaMemsetAsync(b_hll.p, 0, HLL_M * 4, stream));
        KTG_TRY(b_ovf_keys.ensure(OVF_CAP * sizeof(K)));
        KTG_TRY(b_ovf_inc.ensure(OVF_CAP * sizeof(uint32_t)));
        tab.world = cfg.world_size ? cfg.world_size : 1;
        tab.rank = cfg.rank;
        tab.ovf_keys = (K *)b_ovf_keys.p;
        tab.ovf_inc = (uint32_t *)b_ovf_inc.p;
        tab.ovf_count = d_ovf_count;
        tab.ovf_cap = OVF_CAP;
        // capacity hint counts edges the way the reference does (both strands)
        uint64_t entries = cfg.capacity_hint_edges;
        if (rc) entries = (entries + 1) / 2;
        entries = (entries + tab.world - 1) / tab.world;
        uint64_t need = entries ? (uint64_t)((double)entries / LOAD_TARGET) + 1 : (1u << 20);
        KTG_TRY(alloc_table(need, &tab, false));
        fresh = true;
        // opt in to large dynamic shared memory for the scatter kernels
        set_smem_attrs();
        return KTG_OK;
    }

    template <class F> static void allow_smem(F kernel, size_t bytes) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    }
    static size_t mx_scatter_smem() { return ScatterSmem<K, SCATTER_TILE>::bytes(MAX_BINS, true); }
    void set_smem_attrs() {
        const size_t mx = ScatterSmem<K, SCATTER_TILE>::bytes(MAX_BINS, true);
        allow_smem(scatter_buckets_kernel<K, 2>, ScatterSmem<K, L2S_TILE>::bytes(MAX_PAGES_PER_SUB, false));
        allow_smem(scatter_buckets_kernel<K, 1>, ScatterSmem<K, L2S_TILE>::bytes(MAX_BINS, false));
        allow_smem(scatter_keys_kernel<K, true, false>, mx);
        allow_smem(scatter_keys_kernel<K, false, true>, mx);
        allow_smem(scatter_keys_kernel<K, false, false>, mx);
    }

    bool use_partition() const {
        if (cfg.flags & KTG_FLAG_FORCE_DIRECT) return false;
        if (cfg.flags & (KTG_FLAG_FORCE_PARTITION | KTG_FLAG_FORCE_PAGES)) return true;
        return tab.n_sub > 3; // up to ~48 MB of table is L2 resident as a whole
    }
    // Streaming page update or L2 atomics?  The sweep reads and writes every slot
    // (32 B per slot of traffic) and then absorbs ~120 G keys/s, the atomic path runs at
    // ~34 G keys/s whatever the table size: measured on C2 (109 M slots, sweep 0.55 ms) they
    // break even at 26 M keys, a quarter of the slots.
    bool use_pages(uint64_t n_keys) const {
        if (cfg.flags & KTG_FLAG_NO_PAGES) return false;
        if (tab.pages_per_sub() > MAX_PAGES_PER_SUB) return false;
        if (cfg.flags & KTG_FLAG_FORCE_PAGES) return true;
        return 4 * n_keys >= tab.capacity();
    }

    int sync_stream() override { return sync(); }
    int sync() {
        trace("sync>");
        KTG_CUDA(cudaStreamSynchronize(stream));
        trace("sync<");
        prof.resolve();
        return KTG_OK;
    }

    // ---- growth -----------------------------------------------------------------
    int count_occupied(uint64_t *out) {
        if (fresh) {
            *out = 0;
            return KTG_OK;
        }
        KTG_CUDA(cudaMemsetAsync(d_scratch, 0, 8, stream));
        prof.begin("count_occupied", tab.capacity(), stream);
        count_occupied_kernel<K><<<props.sms * 8, 256, 0, stream>>>(tab, d_scratch);
        prof.end(stream);
        unsigned long long v = 0;
        KTG_CUDA(cudaMemcpyAsync(&v, d_scratch, 8, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        *out = v;
        return KTG_OK;
    }

    int grow_to(uint64_t need_slots) {
        Table<K> nt;
        if (fresh) { // nothing to move
            KTG_TRY(alloc_table(need_slots, &nt, false));
            KTG_TRY(sync());
            cudaFree(tab.base);
            tab = nt;
            ++grow_events;
            return KTG_OK;
        }
        KTG_TRY(alloc_table(need_slots, &nt, true));
        prof.begin("rehash", tab.capacity(), stream);
        rehash_kernel<K><<<props.sms * 8, 256, 0, stream>>>(tab, nt);
        prof.end(stream);
        KTG_TRY(sync());
        cudaFree(tab.base);
        tab = nt;
        ++grow_events;
        return KTG_OK;
    }

    // replay inserts that did not fit (after growing); void the build if any were dropped
    int drain_overflow() {
        unsigned long long n = 0;
        KTG_CUDA(cudaMemcpyAsync(&n, d_ovf_count, 8, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        while (n) {
            if (n > OVF_CAP) {
                deferred_error = KTG_ERR_TABLE_FULL;
                return fail(KTG_ERR_TABLE_FULL, "%llu inserts overflowed the replay buffer", n - OVF_CAP);
            }
            KTG_TRY(b_keys2.ensure(n * sizeof(K)));
            DeviceBuf inc;
            KTG_TRY(inc.ensure(n * 4));
            KTG_CUDA(cudaMemcpyAsync(b_keys2.p, tab.ovf_keys, n * sizeof(K), cudaMemcpyDeviceToDevice, stream));
            KTG_CUDA(cudaMemcpyAsync(inc.p, tab.ovf_inc, n * 4, cudaMemcpyDeviceToDevice, stream));
            KTG_CUDA(cudaMemsetAsync(d_ovf_count, 0, 8, stream));
            int rc_ = grow_to(tab.capacity() * 2);
            if (rc_ != KTG_OK) { inc.release(); return rc_; }
            prof.begin("replay_overflow", n, stream);
            replay_overflow_kernel<K><<<props.sms * 4, 256, 0, stream>>>((const K *)b_keys2.p, (const uint32_t *)inc.p, n, tab);
            prof.end(stream);
            KTG_CUDA(cudaMemcpyAsync(&n, d_ovf_count, 8, cudaMemcpyDeviceToHost, stream));
            KTG_TRY(sync());
            inc.release();
        }
        return KTG_OK;
    }

    // HyperLogLog estimate of the distinct keys ever offered to this table
    // (persistent sketch, see kernels.cuh)
    int hll_estimate(double *est) {
        std::vector<uint32_t> regs(HLL_M);
        KTG_CUDA(cudaMemcpyAsync(regs.data(), b_hll.p, HLL_M * 4, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        double m = (double)HLL_M, sum = 0;
        uint32_t zeros = 0;
        for (uint32_t r : regs) {
            sum += ldexp(1.0, -(int)r);
            zeros += r == 0;
        }
        double e = (0.7213 / (1.0 + 1.079 / m)) * m * m / sum;
        if (e <= 2.5 * m && zeros) e = m * log(m / (double)zeros);
        *est = e * HLL_SAMPLE; // only 1/HLL_SAMPLE of the hash space is sketched
        return KTG_OK;
    }

    // Make sure the coming batch cannot push the load past LOAD_MAX.  `bound`
    // is a trivial upper bound on its new keys; when that is not enough to
    // decide, `sketch()` folds the batch into the HyperLogLog sketch and the
    // table is sized from the estimated number of distinct keys.  Estimate
    // errors are caught by the overflow/replay path, never lost.
    template <class F> int reserve(uint64_t bound, F sketch) {
        if ((double)(occupied_ub + bound) <= LOAD_MAX * (double)tab.capacity()) {
            occupied_ub += bound;
            sketch_complete = false; // these keys go in unsketched
            return KTG_OK;
        }
        if (!sketch_complete) {
            uint64_t exact = 0;
            KTG_TRY(count_occupied(&exact));
            occupied_ub = exact;
            if ((double)(exact + bound) <= LOAD_MAX * (double)tab.capacity()) {
                occupied_ub += bound;
                return KTG_OK;
            }
            // restart the sketch here: estimate = exact + distinct(batches from now on)
            hll_base = exact;
            sketch_complete = true;
            KTG_CUDA(cudaMemsetAsync(b_hll.p, 0, HLL_M * 4, stream));
        }
        KTG_TRY(sketch());
        double est = 0;
        KTG_TRY(hll_estimate(&est));
        uint64_t total = hll_base + (uint64_t)(est * 1.08) + 64;
        if ((double)total > LOAD_MAX * (double)tab.capacity()) {
            uint64_t need = (uint64_t)((double)total / LOAD_TARGET) + 1;
            KTG_TRY(grow_to(need));
        }
        occupied_ub = total;
        return KTG_OK;
    }

    // ---- K1 ------------------------------------------------------------------------
    // Packs the batch and returns what the extraction kernels need to know about it:
    // the exact number of windows and whether all reads share one length.
    struct Batch {
        uint64_t windows = 0;
        ReadView v{};
    };
    uint64_t windows_seen = 0; // cumulative PackCounters::windows already accounted for
    DeviceBuf b_bad, b_valid, b_wstart;

    bool counters_stale = false; // batches were queued without reading the counters back

    int pack(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
             uint64_t total_bases, Batch *bt, const BatchHint *hint = nullptr, uint32_t shift0_hint = 0) {
        if (!hint && counters_stale) { // the per-batch window count below is a difference of totals
            PackCounters c0;
            KTG_CUDA(cudaMemcpyAsync(&c0, d_ctr, sizeof c0, cudaMemcpyDeviceToHost, stream));
            KTG_TRY(sync());
            windows_seen = c0.windows;
            counters_stale = false;
        }
        const uint64_t nw_max = (total_bases + 31 + 31) / 32 + 1; // whatever the alignment of the first base
        KTG_TRY(b_packed.ensure((nw_max + 4) * 8));
        KTG_TRY(b_bad.ensure(nw_max * 4));
        KTG_TRY(b_valid.ensure(n_reads + 1));
        // per-batch min/max read length
        KTG_CUDA(cudaMemsetAsync(&d_ctr->min_len, 0xFF, 8, stream));
        KTG_CUDA(cudaMemsetAsync(&d_ctr->max_len, 0, 8, stream));
        prof.begin("pack_reads", total_bases, stream);
        {
            int grid = (int)std::min<uint64_t>((nw_max + 255) / 256, (uint64_t)props.sms * 16);
            pack_flat_kernel<<<grid, 256, 0, stream>>>(d_bases, d_offsets, total_bases, (uint64_t *)b_packed.p,
                                                       (uint32_t *)b_bad.p, d_ctr);
        }
        prof.end(stream);
        prof.begin("check_reads", n_reads, stream);
        {
            // one CTA of 1024 threads per SM: every CTA ends in six atomics on the same six counters, which
            // L2 serialises (~10 ns each); 592 CTAs of 256 were a fixed 30 us per launch, the whole cost
            // of the kernel for a 64 MiB chunk of the host batcher
            int grid = (int)std::min<uint64_t>((n_reads + 1023) / 1024, (uint64_t)props.sms);
            check_reads_kernel<<<grid, 1024, 0, stream>>>(d_offsets, n_reads, d_bases, k, (const uint32_t *)b_bad.p,
                                                         (uint8_t *)b_valid.p, d_ctr);
        }
        prof.end(stream);
        touch();
        PackCounters c{};
        if (hint) { // no round trip: an upper bound on the windows is all the staging needs
            counters_stale = true;
            bt->windows = hint->windows_ub;
            c.shift0 = shift0_hint;
            c.min_len = c.max_len = hint->ulen;
            if (!hint->ulen) c.max_len = 1; // ragged
        }
        else {
            KTG_CUDA(cudaMemcpyAsync(&c, d_ctr, sizeof c, cudaMemcpyDeviceToHost, stream));
            KTG_TRY(sync());
            bt->windows = c.windows - windows_seen;
            windows_seen = c.windows;
            if (c.short_reads) { // hm_gir.rs:40: the reference panics, the build is void
                deferred_error = KTG_ERR_SHORT_READ;
                return fail(KTG_ERR_SHORT_READ, "Read is too short!");
            }
        }
        ReadView &v = bt->v;
        v.packed = (const uint64_t *)b_packed.p;
        v.valid = (const uint8_t *)b_valid.p;
        v.wstart = nullptr;
        v.shift0 = (uint32_t)c.shift0;
        v.n_words = (total_bases + v.shift0 + 31) / 32;
        if (c.min_len == c.max_len) { // every read has the same length: closed-form item map
            const uint32_t L = (uint32_t)c.max_len;
            v.ulen = L;
            v.set_ipr(L >= k ? (L - k + 1 + GRAN - 1) / GRAN : 0);
            v.n_items = n_reads * v.ipr;
        }
        else if (bt->windows) { // ragged: mark the window starts, items are (word, granule) pairs
            KTG_TRY(b_wstart.ensure(nw_max * 4));
            KTG_CUDA(cudaMemsetAsync(b_wstart.p, 0, v.n_words * 4, stream));
            prof.begin("mark_starts", n_reads, stream);
            int grid = (int)std::min<uint64_t>((n_reads + 255) / 256, (uint64_t)props.sms * 16);
            mark_starts_kernel<<<grid, 256, 0, stream>>>(d_offsets, n_reads, d_bases, k, (const uint8_t *)b_valid.p,
                                                         (uint32_t *)b_wstart.p);
            prof.end(stream);
            v.wstart = (const uint32_t *)b_wstart.p;
            v.ulen = 0;
            v.set_ipr(0);
            v.n_items = v.n_words * ITEMS_PER_WORD;
        }
        // nothing after this point reads the caller's bases or offsets
        if (input_consumed) KTG_CUDA(cudaEventRecord(input_consumed, stream));
        return KTG_OK;
    }

    // ---- partition helpers ----------------------------------------------------------
    // b_hist layout: hist[n_bins] | offsets[n_bins + 1] | cursors[n_bins] | spill cursor
    unsigned long long *hist_ptr() { return (unsigned long long *)b_hist.p; }
    unsigned long long *offs_ptr(uint32_t n_bins) { return hist_ptr() + n_bins; }
    unsigned long long *curs_ptr(uint32_t n_bins) { return hist_ptr() + 2 * n_bins + 1; }
    unsigned long long *spill_ptr(uint32_t n_bins) { return hist_ptr() + 3 * n_bins + 1; }
    int ensure_hist(uint32_t n_bins) { return b_hist.ensure((3 * (size_t)n_bins + 4) * 8); }

    static uint64_t bucket_cap_for(uint64_t n_keys, uint32_t n_bins) {
        uint64_t mean = (n_keys + n_bins - 1) / n_bins;
        uint64_t slack = std::max<uint64_t>(4096, (uint64_t)(10.0 * sqrt((double)mean)) + mean / 64);
        uint64_t cap = mean + slack;
        return (cap + L2S_TILE - 1) / L2S_TILE * L2S_TILE; // whole tiles of either consumer
    }
    // page buckets: a page's key count is a sum over its distinct keys of their
    // multiplicities, so its spread grows with the coverage; 1/8 + 1024 is ~3.5 sigma
    // at 100x.  The tail goes to the spill list (exact, just slower).
    static uint64_t page_bucket_cap_for(uint64_t n_keys, uint64_t n_pages) {
        uint64_t mean = (n_keys + n_pages - 1) / n_pages;
        return (mean + mean / 8 + 1024 + 3) & ~3ull;
    }

    int scan_bins_pass(uint32_t n_bins, uint64_t bucket_cap) {
        prof.begin("scan_bins", n_bins, stream);
        scan_bins_kernel<<<1, 1024, 0, stream>>>(hist_ptr(), n_bins, bucket_cap, offs_ptr(n_bins), curs_ptr(n_bins));
        prof.end(stream);
        KTG_CUDA(cudaMemsetAsync(spill_ptr(n_bins), 0, 8, stream));
        return KTG_OK;
    }

    template <bool BY_OWNER> int hist_reads_pass(const Batch &bt, uint32_t n_bins) {
        KTG_TRY(ensure_hist(n_bins));
        KTG_CUDA(cudaMemsetAsync(b_hist.p, 0, n_bins * 8, stream));
        size_t hs = (size_t)n_bins * 4;
        prof.begin("hist_reads", bt.windows, stream);
        if (rc) {
            int g = grid_for(hist_reads_kernel<K, true, BY_OWNER, false>, 256, hs, props);
            hist_reads_kernel<K, true, BY_OWNER, false><<<g, 256, hs, stream>>>(bt.v, k, tab, n_bins, hist_ptr(), nullptr);
        }
        else {
            int g = grid_for(hist_reads_kernel<K, false, BY_OWNER, false>, 256, hs, props);
            hist_reads_kernel<K, false, BY_OWNER, false><<<g, 256, hs, stream>>>(bt.v, k, tab, n_bins, hist_ptr(), nullptr);
        }
        prof.end(stream);
        return KTG_OK;
    }

    ScatterOut scatter_out(uint32_t n_bins, uint64_t bucket_cap, void *out, void *spill, uint64_t spill_cap) {
        ScatterOut o;
        o.cursors = curs_ptr(n_bins);
        o.bucket_cap = bucket_cap;
        o.out = out;
        o.spill_out = spill;
        o.spill_cursor = spill_ptr(n_bins);
        o.spill_cap = spill_cap;
        return o;
    }

    template <int BINS, bool HLL>
    int scatter_reads_pass(const Batch &bt, uint32_t n_bins, const ScatterOut &o, const PeerOut *po = nullptr) {
        uint64_t n_tiles = std::max<uint64_t>(1, (bt.v.n_items + SCATTER_THREADS - 1) / SCATTER_THREADS);
        PeerOut peers{};
        if (po) peers = *po;
        // tuning knobs: CTAs per SM (the fused exchange kernel may share the SMs with the
        // owner-side kernels of the previous chunk when a batch is sent in chunks)
        const int cap_ctas = po ? tune.p2p_ctas : tune.l1_ctas;
        prof.begin(po ? "scatter_reads_p2p" : "scatter_reads", bt.windows, stream);
        if (cap_ctas > 0) n_tiles = std::min<uint64_t>(n_tiles, (uint64_t)props.sms * cap_ctas);
        auto launch = [&](auto kern, int per) {
            const size_t sb = scatter_smem_bytes((size_t)SCATTER_THREADS * per, sizeof(K), n_bins, HLL);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max(sb, mx_scatter_smem()));
            int g = (int)std::min<uint64_t>(grid_for(kern, SCATTER_THREADS, sb, props), n_tiles);
            kern<<<g, SCATTER_THREADS, sb, stream>>>(bt.v, k, tab, n_bins, o, (uint32_t *)b_hll.p, peers);
        };
        // more bins than the one-pass kernel has threads: the big-tile kernel (two extraction passes, 64 KB tiles)
        // (and the sender of the direct exchange, whose bins are owner x sub-table: NVLink wants the longer runs --
        // 100-byte runs reached ~320 GB/s, C3 at N = 2: 20.4 ms with the one-pass sender, 17.3 with this one)
        const bool big = tune.l1_big < 0 ? (n_bins > (uint32_t)SCATTER_THREADS || (po && BINS == BIN_OWNER_PART)) : tune.l1_big != 0;
        auto launch_big = [&](auto kern) {
            const size_t sb = big_scatter_smem(BigGeom<K>::TILE, sizeof(K), n_bins, HLL);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
            uint64_t nt = std::max<uint64_t>(1, (bt.v.n_items + BIG_THREADS * BigGeom<K>::IPT - 1) / (BIG_THREADS * BigGeom<K>::IPT));
            if (cap_ctas > 0) nt = std::min<uint64_t>(nt, (uint64_t)props.sms * cap_ctas);
            int g = (int)std::min<uint64_t>(grid_for(kern, BIG_THREADS, sb, props), nt);
            kern<<<g, BIG_THREADS, sb, stream>>>(bt.v, k, tab, n_bins, o, (uint32_t *)b_hll.p, peers);
        };
        if (big && n_bins <= 2 * BIG_THREADS && big_scatter_smem(BigGeom<K>::TILE, sizeof(K), n_bins, HLL) <= 113 * 1024) {
            if (rc) launch_big(scatter_reads_big_kernel<K, true, BINS, HLL>);
            else launch_big(scatter_reads_big_kernel<K, false, BINS, HLL>);
        }
        else if (rc) launch(scatter_reads_kernel<K, true, BINS, HLL>, ScatterGeom<K>::PER);
        else launch(scatter_reads_kernel<K, false, BINS, HLL>, ScatterGeom<K>::PER);
        prof.end(stream);
        return KTG_OK;
    }

    template <bool BY_OWNER, bool HLL> int scatter_keys_pass(const K *keys, uint64_t n, uint32_t n_bins, const ScatterOut &o) {
        size_t ss = ScatterSmem<K, SCATTER_TILE>::bytes(n_bins, HLL);
        uint64_t n_tiles = std::max<uint64_t>(1, (n + SCATTER_TILE - 1) / SCATTER_TILE);
        int g = (int)std::min<uint64_t>(grid_for(scatter_keys_kernel<K, BY_OWNER, HLL>, SCATTER_THREADS, ss, props), n_tiles);
        prof.begin("scatter_keys", n, stream);
        scatter_keys_kernel<K, BY_OWNER, HLL><<<g, SCATTER_THREADS, ss, stream>>>(keys, n, tab, n_bins, o, (uint32_t *)b_hll.p);
        prof.end(stream);
        return KTG_OK;
    }

    // flat (dense) or bucketed insert with L2 atomics; n_dev: the count lives on the device
    int launch_insert(const K *keys, uint64_t n, const unsigned long long *bin_end, uint64_t bucket_cap,
                      uint32_t n_bins, const unsigned long long *n_dev = nullptr, bool skip_empty = false) {
        uint64_t tiles_per_bin = bin_end ? (bucket_cap + INSERT_TILE - 1) / INSERT_TILE : 0; // the kernel clips a tile at its bin's end
        uint64_t n_tiles = bin_end ? tiles_per_bin * n_bins : (n + INSERT_TILE - 1) / INSERT_TILE;
        if (n_tiles == 0) return KTG_OK;
        KTG_TRY(ensure_init());
        int g = grid_for(insert_keys_kernel<K>, 256, 0, props);
        g = (int)std::min<uint64_t>(g, n_tiles);
        KTG_CUDA(cudaMemsetAsync(d_scratch + 14, 0, 8, stream)); // the tile counter
        prof.begin("insert_keys", n_dev ? 0 : n, stream);
        insert_keys_kernel<K><<<g, 256, 0, stream>>>(keys, n, n_dev, d_lost, k, rc && (k % 2 == 0), skip_empty, tab, d_scratch + 14, bin_end,
                                                      bucket_cap, tiles_per_bin, n_tiles);
        prof.end(stream);
        touch();
        return KTG_OK;
    }
    int launch_insert_keys(const K *keys, uint64_t n) { return launch_insert(keys, n, nullptr, 0, 0); }

    // Level-2 scatter of the level-1 buckets by page, then the streaming page update.
    // Page-bucket overflow goes to a spill list that is inserted with L2 atomics
    // afterwards (count stays on the device: no host round trip).
    // Three steps, so that the host batcher can run the scatter chunk by chunk while the
    // copies are still coming in and leave only the page sweep for the flush (pages_drain):
    //   pages_open    page buckets + cursors for up to `room` keys
    //   pages_scatter level-1 buckets (n_bins of cap1 keys, ends in fill1) -> page buckets;
    //                 sub_mod != 0: bucket q belongs to sub-table q % sub_mod (receive buckets,
    //                 one set per source rank)
    //   pages_update  the sweep + the spill list
    uint64_t pg_cap2 = 0;
    int pages_open(uint64_t room) {
        const uint64_t n_pages = tab.n_pages();
        pg_cap2 = page_bucket_cap_for(room, n_pages);
        KTG_TRY(b_pkeys.ensure(n_pages * pg_cap2 * sizeof(K) + 64));
        KTG_TRY(b_pcur.ensure(n_pages * 8));
        init_cursors_kernel<<<(int)std::min<uint64_t>((n_pages + 255) / 256, props.sms * 8), 256, 0, stream>>>(
            (unsigned long long *)b_pcur.p, n_pages, pg_cap2);
        return KTG_OK;
    }
    int pages_scatter(uint32_t n_bins, uint64_t cap1, uint64_t n_keys, const unsigned long long *fill1,
                      const K *keys1, uint32_t sub_mod) {
        ScatterOut o;
        o.cursors = (unsigned long long *)b_pcur.p;
        o.bucket_cap = pg_cap2;
        o.out = b_pkeys.p;
        // a page bucket that overflows (a heavy hitter: poly-A reads, satellite repeats) spills into the
        // stage's own list, which is as large as the stage: nothing can be lost whatever the skew
        o.spill_out = b_spill.p;
        o.spill_cursor = stage_spill_cursor();
        o.spill_cap = stage_spill_cap;
        const uint64_t tiles_per_bin = cap1 / L2S_TILE, n_tiles = tiles_per_bin * n_bins;
        const size_t ss = ScatterSmem<K, L2S_TILE>::bytes(tab.pages_per_sub(), false);
        // 256-thread CTAs (four per SM) overlap the phases of a tile better than 512-thread ones
        // (1.51 vs 1.60 ms on C2); u128 keys need the registers of the larger block
        // u128 keys: 256 threads x 16 keys (126 registers, no spill; the 512 x 8 geometry spills 128 bytes at its
        // 64-register cap): 10.7 against 10.9 ms on C3 k = 63
        int variant = sizeof(K) == 16 ? 4 : tab.pages_per_sub() <= 256 ? 1 : 0; // u64: at most one bin per thread
        if (tune.l2s_variant >= 0) variant = tune.l2s_variant;
        prof.begin("scatter_pages", n_keys, stream);
        auto launch_v = [&](auto kern, int threads, int per) {
            const uint64_t tile = (uint64_t)threads * per, tpb = cap1 / tile, nt = tpb * n_bins;
            const size_t sb = scatter_smem_bytes(tile, sizeof(K), tab.pages_per_sub(), false);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sb);
            int gg = (int)std::min<uint64_t>(grid_for(kern, threads, sb, props), nt);
            kern<<<gg, threads, sb, stream>>>(keys1, fill1, cap1, tpb, nt, sub_mod, false, tab, o, nullptr);
        };
        if (variant == 1) launch_v(scatter_buckets_kernel<K, 2, 256, 8, 4>, 256, 8);
        else if (variant == 2) launch_v(scatter_buckets_kernel<K, 2, 512, 4, 3>, 512, 4);
        else if (variant == 3) launch_v(scatter_buckets_kernel<K, 2, 1024, 4, 2>, 1024, 4);
        else if (variant == 4) launch_v(scatter_buckets_kernel<K, 2, 256, 16, 2>, 256, 16);
        // (also measured on C2: five CTAs of 256 at 51 registers 1.75 ms, eight of 128 threads 1.54, six of
        // 256 x 4 keys 2.05, against 1.51 for variant 1)
        else {
            int g = (int)std::min<uint64_t>(grid_for(scatter_buckets_kernel<K, 2>, L2S_THREADS, ss, props), n_tiles);
            scatter_buckets_kernel<K, 2><<<g, L2S_THREADS, ss, stream>>>(keys1, fill1, cap1, tiles_per_bin, n_tiles, sub_mod, false, tab, o);
        }
        prof.end(stream);
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }
    int pages_update(uint64_t n_keys) {
        const uint64_t n_pages = tab.n_pages(), cap2 = pg_cap2;
        unsigned long long *cur2 = (unsigned long long *)b_pcur.p;
        const bool palin = rc && (k % 2 == 0), special = !rc && 2 * k == 8 * sizeof(K);
        // geometry (two CTAs per SM): u64 keys 640 threads at 46 registers (C2: 1.397 ms, 704 threads 1.425,
        // one pipelined CTA of 1024 threads with two page buffers 1.56); u128 keys need the registers of
        // the smaller block (512 threads 14.2 ms on C3 k=63, 640 threads 15.7)
        int pt = tune.page_threads ? tune.page_threads : (sizeof(K) == 8 ? 640 : 512);
        const int nbuf = tune.page_nbuf ? tune.page_nbuf : 1;
        KTG_TRY(make_empty_page(tab));
        const unsigned char *empty = fresh ? (const unsigned char *)b_empty_page.p : nullptr;
        prof.begin("update_pages", n_keys, stream);
        auto launch_p = [&](auto kern, int threads, int nb) {
            const size_t ps = page_kernel_smem<K>(threads, nb, tab.page_log2);
            cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ps);
            int g = (int)std::min<uint64_t>(grid_for(kern, threads, ps, props), n_pages);
            kern<<<g, threads, ps, stream>>>((const K *)b_pkeys.p, cur2, cap2, k, tab, empty);
        };
        auto launch_m = [&](auto mode) { // the kernel is specialised for what a key may need besides the plain update
            constexpr int M = decltype(mode)::value;
            if (nbuf == 2 && pt >= 1024) launch_p(update_pages_kernel<K, 1024, 2, M>, 1024, 2);
            else if (nbuf == 2) launch_p(update_pages_kernel<K, 768, 2, M>, 768, 2);
            else if (pt >= 704) launch_p(update_pages_kernel<K, 704, 1, M>, 704, 1);
            else if (pt >= 640) launch_p(update_pages_kernel<K, 640, 1, M>, 640, 1);
            else launch_p(update_pages_kernel<K, 512, 1, M>, 512, 1);
        };
        if (palin) launch_m(std::integral_constant<int, PAGE_PALIN>{});
        else if (special) launch_m(std::integral_constant<int, PAGE_SPECIAL>{});
        else launch_m(std::integral_constant<int, PAGE_PLAIN>{});
        prof.end(stream);
        fresh = false;
        touch();
        ++page_updates;
        return KTG_OK;
    }
    int paged_update(uint32_t n_bins, uint64_t cap1, uint64_t n_keys, const unsigned long long *fill1,
                     const K *keys1 = nullptr, uint32_t sub_mod = 0) {
        if (!keys1) keys1 = (const K *)b_keys.p;
        KTG_TRY(pages_open(n_keys));
        KTG_TRY(pages_scatter(n_bins, cap1, n_keys, fill1, keys1, sub_mod));
        return pages_update(n_keys);
    }

    // ---- staging -------------------------------------------------------------------
    // Batches are not inserted one by one: their keys are partitioned by sub-table
    // (level 1) into buckets that persist across batches, and the buckets are flushed
    // into the table when about as many keys are staged as the table has slots (one
    // sweep of the table then serves all of them), when the next batch would not fit,
    // or when the build is finalised / queried.  With host input this also lets the
    // H2D copy of the following chunks overlap the flush.
    uint32_t stage_bins = 0;      // n_sub the buckets are laid out for (0: no open stage)
    uint32_t stage_sub_log2 = 0;
    uint64_t stage_cap1 = 0;      // bucket capacity per sub-table
    uint64_t stage_room = 0;      // keys the open stage was sized for
    uint64_t stage_target = 0;    // flush once this many keys are staged
    uint64_t stage_spill_cap = 0;
    uint64_t staged_keys = 0;     // keys in the buckets + spill list (+ page buckets, see below)
    // Eager page stage (host batcher, large batches).  With hold_flush the batcher has announced how
    // many keys its call will offer (call_keys_hint) and flushes on its own schedule.  Then every
    // chunk's level-1 buckets are moved on to page buckets right away (pages_drain): the level-2
    // scatter runs while the copies are still coming in, when the GPU has time to spare, and a flush
    // is only the page sweep.  The level-1 buckets then hold one chunk; the spill list stays as large
    // as everything the page stage may take, so it still cannot overflow.
    bool pstage_open = false;
    uint64_t l1_keys = 0;         // keys in the level-1 buckets + spill list since the last drain
    uint64_t pstage_room = 0;     // keys the page buckets were sized for
    uint64_t pstaged_keys = 0;    // keys moved to page buckets (upper bound: includes spilled ones)
    uint32_t pstage_n_sub = 0, pstage_sub_log2 = 0, pstage_page_log2 = 0;
    uint64_t pstage_pages = 0;
    bool eager_pages() const {
        if (!hold_flush || !call_keys_hint || !tune.eager_pages) return false;
        return use_partition() && use_pages(call_keys_hint);
    }
    DeviceBuf b_stage_cur;        // cursors[n_bins] | spill cursor | backup of both
    unsigned long long *stage_cursors() { return (unsigned long long *)b_stage_cur.p; }
    unsigned long long *stage_spill_cursor() { return d_stage_spill; }

    // Most keys one stage may hold.  Two limits: bucket positions are 32-bit (2^31 keys plus the
    // slack of the buckets stay below 4e9), and memory: level-1 buckets, their spill list (as large as
    // the stage) and the page buckets make ~3.3 keys of memory per staged key; the stage may take 70 %
    // of what is free now plus what its own buffers already hold (180 GB of HBM: C3 stages all of its
    // 1.84 G keys, 46 GB, beside the 10.6 GB table and is swept once).
    uint64_t stage_max_keys() const {
        if (tune.stage_max_keys > 0) return std::max<uint64_t>(1u << 16, (uint64_t)tune.stage_max_keys);
        uint64_t lim = ~0ull;
        // (cudaMemGetInfo takes ~1.5 ms on this driver: asked once per table allocation, not per stage)
        if (stage_mem_for != (const void *)tab.base) {
            size_t free_b = 0, total_b = 0;
            stage_mem_keys = ~0ull;
            if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) {
                const uint64_t held = b_keys.cap + b_spill.cap + b_pkeys.cap;
                const uint64_t budget = ((uint64_t)free_b + held) / 10 * 7;
                stage_mem_keys = budget / (uint64_t)(3.3 * sizeof(K));
            }
            else (void)cudaGetLastError();
            stage_mem_for = (const void *)tab.base;
        }
        lim = std::min<uint64_t>(lim, stage_mem_keys);
        return std::max<uint64_t>(lim, 1u << 20);
    }
    mutable uint64_t stage_mem_keys = ~0ull;
    mutable const void *stage_mem_for = nullptr; // the table allocation the memory budget was taken beside

    int stage_open(uint64_t batch_keys) {
        if (!sketch_complete) { // keys went in unsketched (direct path): restart from the exact count
            uint64_t exact = 0;
            KTG_TRY(count_occupied(&exact));
            hll_base = exact;
            sketch_complete = true;
            KTG_CUDA(cudaMemsetAsync(b_hll.p, 0, HLL_M * 4, stream));
        }
        const uint32_t n_bins = tab.n_sub;
        // host input: flushes small enough to hide under the following copies (a sweep per 0.75 x capacity
        // keys).  Device-resident input: every sweep reads and writes the whole table, so as many keys per
        // sweep as fit (C3 in 5 batches: three sweeps of the 10.6 GB table were 17.9 of 41 ms, one is 9.6)
        double factor = host_paced ? 0.75 : 3.0;
        if (tune.stage_factor_milli > 0) factor = tune.stage_factor_milli / 1000.0;
        stage_target = std::min<uint64_t>((uint64_t)(factor * (double)tab.capacity()), stage_max_keys());
        if (host_paced) stage_room = batch_keys >= stage_target ? batch_keys : stage_target + batch_keys;
        else stage_room = std::max(batch_keys, stage_target);
        const bool eager = eager_pages();
        if (eager) { // level-1 buckets are drained after every chunk; the page stage takes the whole call
            stage_room = batch_keys;
            pstage_room = std::min<uint64_t>(std::max(call_keys_hint, batch_keys), stage_max_keys());
        }
        else if (hold_flush && call_keys_hint) // the batcher flushes once, after ~60 % of its input
            stage_room = std::max(stage_room, std::min<uint64_t>(call_keys_hint / 4 * 3 + batch_keys, stage_max_keys()));
        stage_cap1 = bucket_cap_for(stage_room, n_bins);
        // as large as the stage itself: even a batch made of one key cannot overflow it, so a
        // batch is staged without looking at the spill cursor (it is read when the stage is flushed)
        stage_spill_cap = (eager ? pstage_room + batch_keys : stage_room) + 64;
        KTG_TRY(b_keys.ensure(stage_cap1 * n_bins * sizeof(K) + 64));
        KTG_TRY(b_spill.ensure(stage_spill_cap * sizeof(K) + 64));
        KTG_TRY(b_stage_cur.ensure(((size_t)n_bins + 1) * 8));
        stage_bins = n_bins;
        stage_sub_log2 = tab.sub_log2;
        init_cursors_kernel<<<(n_bins + 255) / 256, 256, 0, stream>>>(stage_cursors(), n_bins, stage_cap1);
        KTG_CUDA(cudaMemsetAsync(stage_spill_cursor(), 0, 8, stream));
        staged_keys = 0;
        l1_keys = pstaged_keys = 0;
        pstage_open = false;
        if (eager) {
            KTG_TRY(pages_open(pstage_room));
            pstage_open = true;
            pstage_n_sub = tab.n_sub;
            pstage_sub_log2 = tab.sub_log2;
            pstage_page_log2 = tab.page_log2;
            pstage_pages = tab.n_pages();
        }
        return KTG_OK;
    }

    // eager page stage: level-1 buckets -> page buckets; the level-1 buckets are empty again
    // afterwards (the spill list is not touched: its keys wait for the flush)
    int pages_drain() {
        if (!pstage_open || l1_keys == 0) return KTG_OK;
        KTG_TRY(pages_scatter(stage_bins, stage_cap1, l1_keys, stage_cursors(), (const K *)b_keys.p, 0));
        init_cursors_kernel<<<(stage_bins + 255) / 256, 256, 0, stream>>>(stage_cursors(), stage_bins, stage_cap1);
        pstaged_keys += l1_keys;
        l1_keys = 0;
        return KTG_OK;
    }

    // Adds one batch to the stage.  `scatter(n_bins, o)` runs the level-1 scatter kernel
    // (which also feeds the cardinality sketch); `n_keys` is the number of keys it emits (or an
    // upper bound).  Nothing here waits for the device unless the stage has to be flushed.
    template <class S> int stage_add(uint64_t n_keys, S scatter) {
        if (stage_bins != tab.n_sub || stage_sub_log2 != tab.sub_log2 || l1_keys + n_keys > stage_room ||
            (pstage_open && pstaged_keys + l1_keys + n_keys > pstage_room) || (!pstage_open && stage_bins && eager_pages())) {
            KTG_TRY(flush_staged());
            KTG_TRY(stage_open(n_keys));
        }
        ScatterOut o;
        o.cursors = stage_cursors();
        o.bucket_cap = stage_cap1;
        o.out = b_keys.p;
        o.spill_out = b_spill.p;
        o.spill_cursor = stage_spill_cursor();
        o.spill_cap = stage_spill_cap;
        KTG_TRY(scatter(stage_bins, o));
        staged_keys += n_keys;
        l1_keys += n_keys;
        touch();
        KTG_TRY(pages_drain());
        if (staged_keys >= stage_target && !hold_flush) KTG_TRY(flush_staged());
        return KTG_OK;
    }

    // Size the table for everything offered so far, then move the staged keys into it.
    int flush_staged() {
        if (staged_keys == 0) {
            stage_bins = 0;
            pstage_open = false;
            return KTG_OK;
        }
        trace("flush", staged_keys);
        double est = 0;
        KTG_TRY(hll_estimate(&est)); // synchronises the stream
        // fused multi-GPU mode: the sketch was all-reduced, it describes the keys of ALL ranks
        // (super-k-mer exchange: the owner sketches what it receives, i.e. its own shard)
        const uint64_t distinct = hll_base + (uint64_t)(est * (mg_mode && !mg_local_sketch ? 1.10 / tab.world : 1.08)) + 64;
        occupied_ub = distinct;
        bool moved = false;
        if ((double)distinct > LOAD_MAX * (double)tab.capacity()) {
            // at least double, so that a stream of small batches grows O(log) times
            const uint64_t need = std::max<uint64_t>((uint64_t)((double)distinct / LOAD_TARGET) + 1, 2 * tab.capacity());
            KTG_TRY(grow_to(need));
            moved = tab.n_sub != stage_bins || tab.sub_log2 != stage_sub_log2;
        }
        const uint64_t n = staged_keys, n_l1 = l1_keys, n_pg = pstaged_keys;
        const uint32_t bins = stage_bins;
        const bool eager = pstage_open;
        staged_keys = 0;
        l1_keys = pstaged_keys = 0;
        pstage_open = false;
        stage_bins = 0; // the next batch opens a new stage (sized for the table as it is then)
        // After a change of geometry the buckets no longer match the sub-tables: the keys are
        // still all there, they just lose their L2 locality for this one flush.
        if (eager) {
            const bool pmoved = moved || tab.n_pages() != pstage_pages || tab.page_log2 != pstage_page_log2 ||
                                tab.n_sub != pstage_n_sub || tab.sub_log2 != pstage_sub_log2;
            if (!pmoved) {
                if (n_l1) KTG_TRY(pages_scatter(bins, stage_cap1, n_l1, stage_cursors(), (const K *)b_keys.p, 0));
                KTG_TRY(pages_update(n));
            }
            else { // page buckets of a geometry that is gone: plain arrays of keys now
                if (n_l1) KTG_TRY(launch_insert((const K *)b_keys.p, n_l1, stage_cursors(), stage_cap1, bins));
                if (n_pg) KTG_TRY(launch_insert((const K *)b_pkeys.p, n_pg, (const unsigned long long *)b_pcur.p, pg_cap2, (uint32_t)pstage_pages));
            }
        }
        else if (!moved && use_pages(n)) KTG_TRY(paged_update(bins, stage_cap1, n, stage_cursors()));
        else KTG_TRY(launch_insert((const K *)b_keys.p, n, stage_cursors(), stage_cap1, bins));
        // the spill list (level-1 buckets and page buckets that overflowed; usually empty): the kernel
        // reads the count from the device
        KTG_TRY(launch_insert((const K *)b_spill.p, stage_spill_cap, nullptr, 0, 0, stage_spill_cursor()));
        return KTG_OK;
    }

    // ---- one batch of reads, all on the device ---------------------------------------
    int ingest_device(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                      uint64_t total_bases, const BatchHint *hint = nullptr) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        if (tab.world > 1)
            return fail(KTG_ERR_INVALID, "world_size > 1: use ktg_partition_reads_device + ktg_insert_keys_device");
        if (n_reads == 0) return KTG_OK;
        Batch bt;
        trace("ingest", n_reads);
        if (!use_partition()) hint = nullptr; // the direct path sizes the table from the exact counts
        KTG_TRY(pack(d_bases, d_offsets, n_reads, total_bases, &bt, hint, hint_shift0));
        if (bt.windows == 0) return KTG_OK;
        if (!use_partition()) {
            KTG_TRY(reserve(bt.windows, [&]() -> int {
                prof.begin("hll_reads", bt.windows, stream);
                if (rc) hll_reads_kernel<K, true><<<props.sms * 4, 256, 0, stream>>>(bt.v, k, (uint32_t *)b_hll.p);
                else hll_reads_kernel<K, false><<<props.sms * 4, 256, 0, stream>>>(bt.v, k, (uint32_t *)b_hll.p);
                prof.end(stream);
                return KTG_OK;
            }));
        }
        if (!use_partition()) { // still small after a possible grow: fused extract + insert
            KTG_TRY(ensure_init());
            prof.begin("extract_insert", bt.windows, stream);
            if (rc) {
                int g = grid_for(extract_insert_kernel<K, true>, 256, 0, props);
                g = (int)std::min<uint64_t>(g, (bt.v.n_items + 255) / 256);
                extract_insert_kernel<K, true><<<g, 256, 0, stream>>>(bt.v, k, tab);
            }
            else {
                int g = grid_for(extract_insert_kernel<K, false>, 256, 0, props);
                g = (int)std::min<uint64_t>(g, (bt.v.n_items + 255) / 256);
                extract_insert_kernel<K, false><<<g, 256, 0, stream>>>(bt.v, k, tab);
            }
            prof.end(stream);
        }
        else {
            KTG_TRY(stage_add(bt.windows, [&](uint32_t n_bins, const ScatterOut &o) -> int {
                return scatter_reads_pass<BIN_PART, true>(bt, n_bins, o);
            }));
        }
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    int read_counters(uint64_t *reads, uint64_t *bytes) override {
        PackCounters c;
        KTG_CUDA(cudaMemcpyAsync(&c, d_ctr, sizeof c, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        if (reads) *reads = c.accepted_reads;
        if (bytes) *bytes = c.accepted_bytes;
        windows_inserted = c.windows;
        windows_seen = c.windows;
        counters_stale = false;
        if (c.short_reads) {
            deferred_error = KTG_ERR_SHORT_READ;
            return fail(KTG_ERR_SHORT_READ, "Read is too short!");
        }
        return KTG_OK;
    }

    int flush_hint() override {
        if (deferred_error != KTG_OK || 4 * staged_keys < tab.capacity()) return KTG_OK;
        // a call that does not fit one stage is flushed whenever the stage is full: one more sweep
        // of the table on the way would only add its cost (C3 at k = 63: 30 GB per sweep)
        if (call_keys_hint > stage_max_keys()) return KTG_OK;
        return flush_staged();
    }

    int finalize() override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        KTG_TRY(flush_staged());
        KTG_TRY(read_counters(nullptr, nullptr));
        unsigned long long lost = 0;
        KTG_CUDA(cudaMemcpyAsync(&lost, d_lost, 8, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        if (lost) {
            deferred_error = KTG_ERR_TABLE_FULL;
            return fail(KTG_ERR_TABLE_FULL, "%llu keys overflowed the page spill list", lost);
        }
        KTG_TRY(drain_overflow());
        return KTG_OK;
    }

    int reset() override {
        trace("reset");
        staged_keys = 0;
        l1_keys = pstaged_keys = 0;
        pstage_open = false;
        stage_bins = 0;
        fresh = true; // the next user of the table memory initialises it (ensure_init / page update)
        KTG_TRY(init_special_slot(tab));
        KTG_CUDA(cudaMemsetAsync(b_small.p, 0, 4096, stream));
        KTG_CUDA(cudaMemsetAsync(b_hll.p, 0, HLL_M * 4, stream));
        occupied_ub = 0;
        hll_base = 0;
        sketch_complete = true;
        deferred_error = KTG_OK;
        windows_inserted = 0;
        windows_seen = 0;
        touch();
        page_updates = 0;
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // ---- stats ------------------------------------------------------------------------
    int edge_stats(uint32_t threshold, EdgeStats *out) override {
        trace("edge_stats");
        KTG_TRY(finalize());
        KTG_TRY(ensure_init());
        KTG_CUDA(cudaMemsetAsync(d_scratch, 0, sizeof(EdgeStats), stream));
        uint64_t n = tab.capacity() + 1;
        prof.begin("edge_stats", n, stream);
        if (rc) edge_stats_kernel<K, true><<<props.sms * 8, 256, 0, stream>>>(tab, k, threshold, (EdgeStats *)d_scratch);
        else edge_stats_kernel<K, false><<<props.sms * 8, 256, 0, stream>>>(tab, k, threshold, (EdgeStats *)d_scratch);
        prof.end(stream);
        KTG_CUDA(cudaMemcpyAsync(out, d_scratch, sizeof(EdgeStats), cudaMemcpyDeviceToHost, stream));
        return sync();
    }

    // an empty node table for up to max_entries canonical (k-1)-mers (the caller frees nt->base)
    template <class KN> int make_node_table(uint64_t max_entries, Table<KN> *out) {
        Table<KN> nt{};
        uint64_t need = (uint64_t)((double)(max_entries + 1) / 0.8) + 1024;
        uint32_t l = 10;
        while ((1ull << l) < need) ++l;
        nt.n_sub = 1;
        nt.sub_log2 = l;
        nt.sub_mask = (uint32_t)((1ull << l) - 1);
        nt.page_log2 = l; // one page: probing wraps around the whole node table
        nt.page_mask = nt.sub_mask;
        nt.max_probe = (uint32_t)std::min<uint64_t>(1ull << l, 1u << 20);
        nt.world = 1;
        nt.rank = 0;
        nt.ovf_keys = nullptr;
        nt.ovf_inc = nullptr;
        nt.ovf_count = d_scratch + 15; // count only; nothing is stored (ovf_cap = 0)
        nt.ovf_cap = 0;
        void *p = nullptr;
        size_t bytes = nt.bytes();
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e != cudaSuccess) {
            (void)cudaGetLastError();
            return fail(KTG_ERR_CUDA, "cannot allocate %zu bytes for the node table: %s", bytes, cudaGetErrorString(e));
        }
        nt.base = (unsigned char *)p;
        KTG_CUDA(cudaMemsetAsync(d_scratch, 0, 128, stream));
        prof.begin("init_table", nt.capacity() + 1, stream);
        init_table_kernel<KN><<<props.sms * 8, 256, 0, stream>>>(nt);
        prof.end(stream);
        *out = nt;
        return KTG_OK;
    }

    // node table of this shard's edges: prefix and suffix of every (both-strand expanded) edge
    template <class KN> int build_node_table(Table<KN> *nt) {
        KTG_TRY(ensure_init());
        uint64_t occ = 0;
        KTG_TRY(count_occupied(&occ));
        KTG_TRY(make_node_table<KN>(2 * occ, nt)); // distinct canonical nodes <= 2 x live canonical edges
        uint64_t n = tab.capacity() + 1;
        prof.begin("build_nodes", n, stream);
        if (rc) build_nodes_kernel<K, KN, true><<<props.sms * 8, 256, 0, stream>>>(tab, k, *nt);
        else build_nodes_kernel<K, KN, false><<<props.sms * 8, 256, 0, stream>>>(tab, k, *nt);
        prof.end(stream);
        return KTG_OK;
    }

    template <class KN> int node_table_stats(const Table<KN> &nt, NodeStats *out) {
        prof.begin("node_stats", nt.capacity() + 1, stream);
        if (rc) node_stats_kernel<KN, true><<<props.sms * 8, 256, 0, stream>>>(nt, (NodeStats *)d_scratch);
        else node_stats_kernel<KN, false><<<props.sms * 8, 256, 0, stream>>>(nt, (NodeStats *)d_scratch);
        prof.end(stream);
        unsigned long long host[16];
        KTG_CUDA(cudaMemcpyAsync(host, d_scratch, sizeof host, cudaMemcpyDeviceToHost, stream));
        int rc_ = sync();
        cudaFree(nt.base);
        KTG_TRY(rc_);
        if (host[15]) return fail(KTG_ERR_TABLE_FULL, "node table overflow (%llu)", host[15]);
        memcpy(out, host, sizeof(NodeStats));
        return KTG_OK;
    }

    template <class KN> int node_stats_t(NodeStats *out) {
        Table<KN> nt{};
        KTG_TRY(build_node_table<KN>(&nt));
        return node_table_stats<KN>(nt, out);
    }

    // ---- nodes of a sharded table: a node's edges may live on several shards -----------------
    DeviceBuf b_node_keys, b_node_deg;
    template <class KN> int nodes_export_t(void **d_keys, void **d_deg, uint64_t *n_out) {
        Table<KN> nt{};
        KTG_TRY(build_node_table<KN>(&nt));
        uint64_t occ = 0; // exact number of entries: one more scan, this is not a hot path
        {
            KTG_CUDA(cudaMemsetAsync(d_scratch, 0, 8, stream));
            count_occupied_kernel<KN><<<props.sms * 8, 256, 0, stream>>>(nt, d_scratch);
            unsigned long long v = 0;
            KTG_CUDA(cudaMemcpyAsync(&v, d_scratch, 8, cudaMemcpyDeviceToHost, stream));
            int rc_ = sync();
            if (rc_ != KTG_OK) { cudaFree(nt.base); return rc_; }
            occ = v;
        }
        int rc_ = b_node_keys.ensure(occ * sizeof(KN) + 64);
        if (rc_ == KTG_OK) rc_ = b_node_deg.ensure(occ * 4 + 64);
        if (rc_ != KTG_OK) { cudaFree(nt.base); return rc_; }
        KTG_CUDA(cudaMemsetAsync(d_scratch, 0, 8, stream));
        prof.begin("compact_nodes", nt.capacity() + 1, stream);
        compact_nodes_kernel<KN><<<props.sms * 8, 256, 0, stream>>>(nt, (KN *)b_node_keys.p, (uint32_t *)b_node_deg.p, d_scratch);
        prof.end(stream);
        unsigned long long host[16];
        KTG_CUDA(cudaMemcpyAsync(host, d_scratch, sizeof host, cudaMemcpyDeviceToHost, stream));
        rc_ = sync();
        cudaFree(nt.base);
        KTG_TRY(rc_);
        if (host[15]) return fail(KTG_ERR_TABLE_FULL, "node table overflow (%llu)", host[15]);
        *d_keys = b_node_keys.p;
        *d_deg = b_node_deg.p;
        *n_out = host[0];
        return KTG_OK;
    }
    int nodes_export(void **d_keys, void **d_deg, uint64_t *n, uint32_t *key_words) override {
        KTG_TRY(finalize());
        *key_words = k - 1 <= 32 ? 1 : 2;
        if (k - 1 <= 32) return nodes_export_t<uint64_t>(d_keys, d_deg, n);
        return nodes_export_t<u128>(d_keys, d_deg, n);
    }
    template <class KN> int nodes_stats_from_t(const void *d_keys, const void *d_deg, uint64_t n, NodeStats *out) {
        Table<KN> nt{};
        KTG_TRY(make_node_table<KN>(n, &nt));
        if (n) {
            prof.begin("merge_nodes", n, stream);
            merge_nodes_kernel<KN><<<props.sms * 8, 256, 0, stream>>>((const KN *)d_keys, (const uint32_t *)d_deg, n, nt);
            prof.end(stream);
        }
        return node_table_stats<KN>(nt, out);
    }
    int nodes_stats_from(const void *d_keys, const void *d_deg, uint64_t n, NodeStats *out) override {
        if (k - 1 <= 32) return nodes_stats_from_t<uint64_t>(d_keys, d_deg, n, out);
        return nodes_stats_from_t<u128>(d_keys, d_deg, n, out);
    }

    // sums behind standardize_edges and the scaling itself, split so that the sums can be
    // all-reduced over the shards in between (standardizer.rs:42-70,123-127)
    int edge_sums(uint32_t threshold, uint64_t *sum_w, uint64_t *sum_below) override {
        EdgeStats es;
        KTG_TRY(edge_stats(threshold, &es));
        *sum_w = es.sum_w;
        *sum_below = es.sum_w_below;
        return KTG_OK;
    }
    int scale_weights(double p, uint32_t t) override {
        KTG_TRY(finalize());
        KTG_TRY(ensure_init());
        uint64_t n = tab.capacity() + 1;
        prof.begin("standardize", n, stream);
        standardize_kernel<K><<<props.sms * 8, 256, 0, stream>>>(tab, p, t);
        prof.end(stream);
        touch();
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    int node_stats(NodeStats *out) override {
        KTG_TRY(finalize());
        if (!nodes_valid) {
            if (k - 1 <= 32) KTG_TRY(node_stats_t<uint64_t>(&node_cache));
            else KTG_TRY(node_stats_t<u128>(&node_cache));
            nodes_valid = true;
        }
        *out = node_cache;
        return KTG_OK;
    }

    // ---- K4 -----------------------------------------------------------------------------
    int remove_weak_edges(uint32_t t) override {
        KTG_TRY(finalize());
        if (t <= 1) return KTG_OK; // every stored edge has w >= 1
        KTG_TRY(ensure_init());
        uint64_t n = tab.capacity() + 1;
        prof.begin("filter", n, stream);
        filter_kernel<K><<<props.sms * 8, 256, 0, stream>>>(tab, t);
        prof.end(stream);
        touch();
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    int standardize(uint64_t G, uint64_t k_, uint32_t t) override {
        EdgeStats es;
        KTG_TRY(edge_stats(t, &es));
        if (G < k_ || es.sum_w == es.sum_w_below)
            return fail(KTG_ERR_DEGENERATE, "degenerate standardization ratio (G=%llu k=%llu s=%llu l=%llu)",
                        (unsigned long long)G, (unsigned long long)k_, es.sum_w, es.sum_w_below);
        double p = (double)(G - k_) / (double)(es.sum_w - es.sum_w_below); // standardizer.rs:123-127
        uint64_t n = tab.capacity() + 1;
        prof.begin("standardize", n, stream);
        standardize_kernel<K><<<props.sms * 8, 256, 0, stream>>>(tab, p, t);
        prof.end(stream);
        touch();
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // ---- export: what Convert::create_from consumes (hm_gir.rs:156-226) ------------------------------
    // Three steps that a sharded handle (multi.cuh) runs on different devices: every shard compacts its
    // both-strand expanded edges into arrays that may live on another GPU (peer memory), the gathering
    // device sorts them and derives the graph.
    int compact_edges_into(uint64_t *d_hi, uint64_t *d_lo, uint32_t *d_w, uint64_t cap) override {
        KTG_TRY(finalize());
        KTG_TRY(ensure_init());
        KTG_CUDA(cudaMemsetAsync(d_scratch, 0, 8, stream));
        prof.begin("compact_edges", tab.capacity() + 1, stream);
        if (rc) compact_edges_kernel<K, true><<<props.sms * 8, 256, 0, stream>>>(tab, k, 1, d_hi, d_lo, d_w, cap, d_scratch);
        else compact_edges_kernel<K, false><<<props.sms * 8, 256, 0, stream>>>(tab, k, 1, d_hi, d_lo, d_w, cap, d_scratch);
        prof.end(stream);
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // ne unsorted edges in device arrays of THIS device -> sorted by k-mer (arrays owned by sc)
    int sort_device_edges(Scratch &sc, KeyArr *keys, uint32_t **weights, uint64_t ne) {
        if (!ne) return KTG_OK;
        uint32_t *perm;
        KeyArr s_;
        prof.begin("sort_edges", ne, stream);
        KTG_TRY(sort_keys(*keys, ne, 2 * k, sc, stream, &perm, &s_));
        uint32_t *w_s;
        KTG_TRY(sc.alloc(&w_s, ne));
        gather_kernel<uint32_t><<<export_grid(ne), 256, 0, stream>>>(*weights, perm, w_s, ne);
        prof.end(stream);
        *keys = s_;
        *weights = w_s;
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // this handle's edges: count, allocate, compact (+ sort)
    int device_edges(bool sorted, Scratch &sc, KeyArr *keys, uint32_t **weights, uint64_t *n_out) {
        EdgeStats es;
        KTG_TRY(edge_stats(0, &es));
        const uint64_t ne = es.edges;
        *n_out = ne;
        uint64_t *d_hi = nullptr, *d_lo;
        uint32_t *d_w;
        KTG_TRY(sc.alloc(&d_lo, ne));
        KTG_TRY(sc.alloc(&d_w, ne));
        if (T::WORDS == 2) KTG_TRY(sc.alloc(&d_hi, ne));
        KTG_TRY(compact_edges_into(d_hi, d_lo, d_w, ne));
        keys->hi = d_hi;
        keys->lo = d_lo;
        *weights = d_w;
        if (sorted) KTG_TRY(sort_device_edges(sc, keys, weights, ne));
        return KTG_OK;
    }

    int edges_to_host(uint64_t *d_hi, uint64_t *d_lo, uint32_t *d_w, uint64_t ne, int sorted, uint64_t *hi, uint64_t *lo,
                      uint32_t *w, uint64_t cap) override {
        Scratch sc;
        KeyArr keys{d_hi, d_lo};
        if (sorted) KTG_TRY(sort_device_edges(sc, &keys, &d_w, ne));
        const uint64_t m = std::min(ne, cap);
        if (m == 0) return sync();
        KTG_CUDA(cudaMemcpyAsync(lo, keys.lo, m * 8, cudaMemcpyDeviceToHost, stream));
        KTG_CUDA(cudaMemcpyAsync(w, d_w, m * 4, cudaMemcpyDeviceToHost, stream));
        if (hi) {
            if (keys.hi) KTG_CUDA(cudaMemcpyAsync(hi, keys.hi, m * 8, cudaMemcpyDeviceToHost, stream));
            else memset(hi, 0, m * 8);
        }
        return sync(); // before the scratch buffers are freed
    }

    int export_edges(uint64_t *hi, uint64_t *lo, uint32_t *w, uint64_t cap, int sorted,
                     uint64_t *n_out) override {
        if (!lo || !w || cap == 0) { // size query
            EdgeStats es;
            KTG_TRY(edge_stats(0, &es));
            if (n_out) *n_out = es.edges;
            return KTG_OK;
        }
        Scratch sc;
        KeyArr keys;
        uint32_t *d_w;
        uint64_t ne = 0;
        KTG_TRY(device_edges(false, sc, &keys, &d_w, &ne));
        if (n_out) *n_out = ne;
        return edges_to_host(keys.hi, keys.lo, d_w, ne, sorted, hi, lo, w, cap);
    }

    // The input of Convert::create_from in a canonical numbering (export.cuh): sorted nodes,
    // sorted edges with the node indices of their prefix and suffix, edges as compress_edge bytes.
    // From ne unsorted edges in device arrays of THIS device.  n_nodes / n_edges: what the caller
    // allocated for (checked); *got_nodes / *got_edges (optional): what the graph has.
    // sorted edges -> sorted distinct nodes and, if wanted, every edge's (prefix, suffix) node indices
    int device_graph(Scratch &sc, const KeyArr &edges, uint64_t ne, KeyArr *nodes, uint64_t *nn, uint64_t **d_src, uint64_t **d_dst) {
        const bool wide_nodes = k - 1 > 32;
        *nodes = KeyArr{nullptr, nullptr};
        *nn = 0;
        if (!ne) return KTG_OK;
        KeyArr cand{nullptr, nullptr}, cand_sorted;
        KTG_TRY(sc.alloc(&cand.lo, 2 * ne));
        if (wide_nodes) KTG_TRY(sc.alloc(&cand.hi, 2 * ne));
        prof.begin("graph_nodes", ne, stream);
        split_nodes_kernel<<<export_grid(ne), 256, 0, stream>>>(edges.hi, edges.lo, ne, k, cand.hi, cand.lo);
        uint32_t *perm;
        KTG_TRY(sort_keys(cand, 2 * ne, 2 * (k - 1), sc, stream, &perm, &cand_sorted));
        KTG_TRY(unique_sorted(cand_sorted, 2 * ne, sc, stream, nodes, nn));
        prof.end(stream);
        if (d_src) {
            KTG_TRY(sc.alloc(d_src, ne));
            KTG_TRY(sc.alloc(d_dst, ne));
            prof.begin("graph_ids", ne, stream);
            node_ids_kernel<<<export_grid(ne), 256, 0, stream>>>(edges.hi, edges.lo, ne, k, nodes->hi, nodes->lo, *nn, *d_src, *d_dst);
            prof.end(stream);
        }
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // ne unsorted edges in device arrays of THIS device (owned by `from`, which the cache takes over) ->
    // the device graph: sorted edges, sorted nodes, every edge's node indices
    int graph_build(Scratch &from, uint64_t *d_ehi, uint64_t *d_elo, uint32_t *d_w, uint64_t ne) override {
        graph.clear();
        std::swap(graph.sc.ptrs, from.ptrs);
        graph.edges = KeyArr{d_ehi, d_elo};
        graph.w = d_w;
        graph.ne = ne;
        int rc_ = sort_device_edges(graph.sc, &graph.edges, &graph.w, ne);
        if (rc_ == KTG_OK) rc_ = device_graph(graph.sc, graph.edges, ne, &graph.nodes, &graph.nn, &graph.src, &graph.dst);
        if (rc_ != KTG_OK) {
            graph.clear();
            return rc_;
        }
        graph.valid = true;
        return KTG_OK;
    }
    int graph_prepare(uint64_t *n_nodes, uint64_t *n_edges) override {
        if (!graph.valid) {
            Scratch sc;
            KeyArr edges;
            uint32_t *d_w;
            uint64_t ne = 0;
            KTG_TRY(device_edges(false, sc, &edges, &d_w, &ne));
            KTG_TRY(graph_build(sc, edges.hi, edges.lo, d_w, ne));
        }
        if (n_nodes) *n_nodes = graph.nn;
        if (n_edges) *n_edges = graph.ne;
        return KTG_OK;
    }
    bool graph_ready() const override { return graph.valid; }

    int graph_to_host(uint64_t *node_hi, uint64_t *node_lo, uint64_t n_nodes, uint64_t *src, uint64_t *dst, uint32_t *weight,
                      uint8_t *edge_bytes, uint64_t n_edges) override {
        const uint64_t ne = graph.ne, nn = graph.nn;
        if (ne != n_edges) return fail(KTG_ERR_INVALID, "n_edges is %llu, the graph has %llu edges (ktg_counts)",
                                       (unsigned long long)n_edges, (unsigned long long)ne);
        if (nn != n_nodes) return fail(KTG_ERR_INVALID, "n_nodes is %llu, the graph has %llu nodes (ktg_counts)",
                                       (unsigned long long)n_nodes, (unsigned long long)nn);
        if (ne == 0) return sync();
        Scratch sc;
        if (src) KTG_CUDA(cudaMemcpyAsync(src, graph.src, ne * 8, cudaMemcpyDeviceToHost, stream));
        if (dst) KTG_CUDA(cudaMemcpyAsync(dst, graph.dst, ne * 8, cudaMemcpyDeviceToHost, stream));
        if (edge_bytes) {
            const size_t rec = (k + 3) / 4 + 1;
            uint8_t *d_b;
            KTG_TRY(sc.alloc(&d_b, ne * rec));
            prof.begin("edge_bytes", ne, stream);
            edge_bytes_kernel<<<export_grid(ne), 256, 0, stream>>>(graph.edges.hi, graph.edges.lo, ne, k, d_b);
            prof.end(stream);
            KTG_CUDA(cudaMemcpyAsync(edge_bytes, d_b, ne * rec, cudaMemcpyDeviceToHost, stream));
        }
        if (weight) KTG_CUDA(cudaMemcpyAsync(weight, graph.w, ne * 4, cudaMemcpyDeviceToHost, stream));
        if (node_lo) KTG_CUDA(cudaMemcpyAsync(node_lo, graph.nodes.lo, nn * 8, cudaMemcpyDeviceToHost, stream));
        if (node_hi) {
            if (graph.nodes.hi) KTG_CUDA(cudaMemcpyAsync(node_hi, graph.nodes.hi, nn * 8, cudaMemcpyDeviceToHost, stream));
            else memset(node_hi, 0, nn * 8);
        }
        KTG_CUDA(cudaGetLastError());
        return sync(); // before the scratch buffer is freed
    }

    // The seeds of remove_dead_paths (pruner.rs:165-195, `Externals`): in ascending node index (the
    // numbering of graph_to_host), every node without an incoming edge as Input (kind 0), else every
    // node without an outgoing edge as Output (kind 1).
    int externals_to_host(uint64_t *ids, uint8_t *kinds, uint64_t cap, uint64_t *n_out) override {
        const uint64_t ne = graph.ne, nn = graph.nn;
        if (n_out) *n_out = 0;
        if (nn == 0) return sync();
        Scratch sc;
        uint32_t *deg, *flag, *pos;
        KTG_TRY(sc.alloc(&deg, nn));
        KTG_TRY(sc.alloc(&flag, nn + 1));
        KTG_TRY(sc.alloc(&pos, nn + 1));
        KTG_CUDA(cudaMemsetAsync(deg, 0, nn * 4, stream));
        prof.begin("externals", nn, stream);
        mark_degrees_kernel<<<export_grid(ne), 256, 0, stream>>>(graph.src, graph.dst, ne, deg);
        external_flags_kernel<<<export_grid(nn + 1), 256, 0, stream>>>(deg, nn, flag);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, flag, pos, (int)(nn + 1), stream);
        void *tmp;
        KTG_TRY(sc.alloc((uint8_t **)&tmp, tb));
        KTG_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tb, flag, pos, (int)(nn + 1), stream));
        uint32_t total = 0;
        KTG_CUDA(cudaMemcpyAsync(&total, pos + nn, 4, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        if (n_out) *n_out = total;
        const uint64_t m = std::min<uint64_t>(total, cap);
        if (m && ids && kinds) {
            uint64_t *d_ids;
            uint8_t *d_kinds;
            KTG_TRY(sc.alloc(&d_ids, total));
            KTG_TRY(sc.alloc(&d_kinds, total));
            scatter_externals_kernel<<<export_grid(nn), 256, 0, stream>>>(deg, flag, pos, nn, d_ids, d_kinds);
            KTG_CUDA(cudaMemcpyAsync(ids, d_ids, m * 8, cudaMemcpyDeviceToHost, stream));
            KTG_CUDA(cudaMemcpyAsync(kinds, d_kinds, m, cudaMemcpyDeviceToHost, stream));
        }
        prof.end(stream);
        KTG_CUDA(cudaGetLastError());
        return sync();
    }

    int export_graph(uint64_t *node_hi, uint64_t *node_lo, uint64_t n_nodes, uint64_t *src, uint64_t *dst,
                     uint32_t *weight, uint8_t *edge_bytes, uint64_t n_edges) override {
        KTG_TRY(finalize());
        KTG_TRY(graph_prepare(nullptr, nullptr));
        return graph_to_host(node_hi, node_lo, n_nodes, src, dst, weight, edge_bytes, n_edges);
    }
    int export_externals(uint64_t *ids, uint8_t *kinds, uint64_t cap, uint64_t *n_out) override {
        KTG_TRY(finalize());
        KTG_TRY(graph_prepare(nullptr, nullptr));
        return externals_to_host(ids, kinds, cap, n_out);
    }

    // ---- multi-GPU phases -------------------------------------------------------------------
    int partition_reads(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                        uint64_t total_bases, void **d_keys, uint64_t *counts) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        uint32_t W = tab.world;
        for (uint32_t i = 0; i < W; ++i) counts[i] = 0;
        *d_keys = nullptr;
        if (n_reads == 0) return KTG_OK;
        Batch bt;
        KTG_TRY(pack(d_bases, d_offsets, n_reads, total_bases, &bt));
        if (bt.windows == 0) return KTG_OK;
        // NCCL wants dense per-destination ranges: exact two-pass partition by owner
        KTG_TRY(hist_reads_pass<true>(bt, W));
        KTG_TRY(scan_bins_pass(W, 0));
        KTG_TRY(b_keys2.ensure(bt.windows * sizeof(K) + 64));
        KTG_TRY((scatter_reads_pass<BIN_OWNER, false>(bt, W, scatter_out(W, 0, b_keys2.p, nullptr, 0))));
        std::vector<unsigned long long> h(W);
        KTG_CUDA(cudaMemcpyAsync(h.data(), hist_ptr(), W * 8, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync()); // the caller hands the buffer to NCCL on its own stream
        for (uint32_t i = 0; i < W; ++i) counts[i] = h[i];
        *d_keys = b_keys2.p;
        return KTG_OK;
    }

    int insert_keys(const void *d_keys, uint64_t n) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        if (n == 0) return KTG_OK;
        const K *keys = (const K *)d_keys;
        if (!use_partition()) {
            KTG_TRY(reserve(n, [&]() -> int {
                prof.begin("hll_keys", n, stream);
                hll_keys_kernel<K><<<props.sms * 4, 256, 0, stream>>>(keys, n, (uint32_t *)b_hll.p);
                prof.end(stream);
                return KTG_OK;
            }));
        }
        if (!use_partition()) {
            KTG_TRY(launch_insert_keys(keys, n));
        }
        else {
            KTG_TRY(stage_add(n, [&](uint32_t n_bins, const ScatterOut &o) -> int {
                return scatter_keys_pass<false, true>(keys, n, n_bins, o);
            }));
            KTG_TRY(sync()); // the caller may reuse its key buffer
        }
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // BFCounter input: pre-counted k-mers (device arrays), weight[kmer] += weight
    DeviceBuf b_bfc_ctr;
    int add_weighted_kmers(const uint8_t *d_kmers, const uint32_t *d_weights, uint64_t n, uint32_t threshold,
                           uint64_t *accepted) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        if (tab.world > 1) return fail(KTG_ERR_INVALID, "BFCounter input is single-GPU");
        if (accepted) *accepted = 0;
        if (n == 0) return KTG_OK;
        KTG_TRY(flush_staged());
        KTG_TRY(ensure_init());
        if ((double)(occupied_ub + n) > LOAD_MAX * (double)tab.capacity()) { // every line may be a new edge
            uint64_t exact = 0;
            KTG_TRY(count_occupied(&exact));
            occupied_ub = exact;
            if ((double)(exact + n) > LOAD_MAX * (double)tab.capacity())
                KTG_TRY(grow_to((uint64_t)((double)(exact + n) / LOAD_TARGET) + 1));
        }
        occupied_ub += n;
        sketch_complete = false; // these keys go in unsketched
        KTG_TRY(b_bfc_ctr.ensure(16));
        KTG_CUDA(cudaMemsetAsync(b_bfc_ctr.p, 0, 16, stream));
        const int g = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)props.sms * 8);
        prof.begin("insert_weighted_kmers", n, stream);
        if (rc) insert_weighted_kmers_kernel<K, true><<<g, 256, 0, stream>>>(d_kmers, d_weights, n, k, threshold, tab, (unsigned long long *)b_bfc_ctr.p);
        else insert_weighted_kmers_kernel<K, false><<<g, 256, 0, stream>>>(d_kmers, d_weights, n, k, threshold, tab, (unsigned long long *)b_bfc_ctr.p);
        prof.end(stream);
        KTG_CUDA(cudaGetLastError());
        touch();
        unsigned long long c[2] = {0, 0};
        KTG_CUDA(cudaMemcpyAsync(c, b_bfc_ctr.p, 16, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        if (c[1]) {
            deferred_error = KTG_ERR_BAD_RECORD;
            return fail(KTG_ERR_BAD_RECORD, "%llu BFCounter k-mers with a symbol outside ACGT", c[1]);
        }
        if (accepted) *accepted = c[0];
        return KTG_OK;
    }

    // exact partition of an array of keys by owner rank (spill lists of the fused path)
    int partition_keys(const void *d_keys, uint64_t n, void **d_out, uint64_t *counts) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        const uint32_t W = tab.world;
        for (uint32_t i = 0; i < W; ++i) counts[i] = 0;
        *d_out = nullptr;
        if (n == 0) return KTG_OK;
        KTG_TRY(ensure_hist(W));
        KTG_CUDA(cudaMemsetAsync(b_hist.p, 0, W * 8, stream));
        prof.begin("hist_keys", n, stream);
        hist_keys_kernel<K, true, false><<<props.sms * 4, 256, W * 4, stream>>>((const K *)d_keys, n, tab, W, hist_ptr(), nullptr);
        prof.end(stream);
        KTG_TRY(scan_bins_pass(W, 0));
        KTG_TRY(b_keys2.ensure(n * sizeof(K) + 64));
        KTG_TRY((scatter_keys_pass<true, false>((const K *)d_keys, n, W, scatter_out(W, 0, b_keys2.p, nullptr, 0))));
        std::vector<unsigned long long> h(W);
        KTG_CUDA(cudaMemcpyAsync(h.data(), hist_ptr(), W * 8, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        for (uint32_t i = 0; i < W; ++i) counts[i] = h[i];
        *d_out = b_keys2.p;
        return KTG_OK;
    }

    // ---- multi-GPU, fused: the extraction kernel writes straight into the owners' HBM ------
    // Rank r reserves one receive bucket of mg_cap keys per source rank.  Only rank s writes
    // bucket s (cursors are local to the sender: no remote atomics), in runs of tile / world
    // keys, so NVLink sees kilobyte-sized writes.  After the senders' kernels have completed the
    // owner partitions what it received by sub-table (level 1) into its staged buckets and
    // carries on exactly as on one GPU.  Nothing here depends on the table geometry, so ranks
    // may grow their shards independently.
    // MG_SLOTS receive buffers (one allocation, one IPC handle): chunk c of a batch goes to slot
    // c % MG_SLOTS, so that the owner can consume one chunk while the senders write the next.
    static constexpr uint32_t MG_SLOTS = 2;
    DeviceBuf b_rx, b_mg_cur, b_mg_spill;
    uint64_t mg_cap = 0, mg_spill_cap = 0;
    bool mg_mode = false;
    size_t mg_slot_keys() const { return (size_t)mg_cap * tab.world; }
    // run padding of the exchange (PeerOut::pad), in keys: 128 bytes, unless all-ones is a real key
    uint32_t mg_pad() const {
        if (!rc && 2 * k == 8 * sizeof(K)) return 1;
        if (!tune.mg_pad) return 1;
        return 128 / sizeof(K);
    }
    unsigned long long *mg_cursors(uint32_t slot) { return (unsigned long long *)b_mg_cur.p + (size_t)slot * tab.world; }
    unsigned long long *mg_spill_cursor() { return (unsigned long long *)b_mg_cur.p + (size_t)MG_SLOTS * tab.world; }

    int mg_geometry(uint64_t max_windows, uint64_t *cap) {
        const uint32_t W = tab.world;
        if (W > (uint32_t)MAX_P2P_WORLD) return fail(KTG_ERR_INVALID, "fused exchange needs world <= %d", MAX_P2P_WORLD);
        // every (tile, owner) run may be rounded up by pad - 1 filler keys
        const uint64_t fillers = (max_windows / SCATTER_TILE + 1) * W * (mg_pad() - 1);
        *cap = bucket_cap_for(std::max<uint64_t>(max_windows + fillers, 1), W);
        return KTG_OK;
    }
    // would mg_prepare(max_windows) replace the receive buffer?  (peers must unmap it first)
    int mg_plan(uint64_t max_windows, int *needs_realloc) override {
        uint64_t cap = 0;
        KTG_TRY(mg_geometry(max_windows, &cap));
        *needs_realloc = cap > mg_cap || !b_rx.p;
        return KTG_OK;
    }
    int mg_prepare(uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap,
                   uint32_t *n_sub) override {
        uint64_t cap = 0;
        KTG_TRY(mg_geometry(max_windows, &cap));
        const uint32_t W = tab.world;
        if (cap > mg_cap || !b_rx.p) {
            KTG_TRY(sync());
            b_rx.release();
            KTG_TRY(b_rx.ensure((size_t)cap * W * MG_SLOTS * sizeof(K) + 64));
            mg_cap = cap;
            skm_cap = mgd_cap = 0; // the other exchanges have to size the buffer again
            mg_spill_cap = std::max<uint64_t>(1u << 20, max_windows / 4); // per batch of several chunks
            KTG_TRY(b_mg_spill.ensure(mg_spill_cap * sizeof(K) + 64));
            KTG_TRY(b_mg_cur.ensure(((size_t)W * MG_SLOTS + 2) * 8));
        }
        *rx_base = b_rx.p;
        *rx_bytes = (size_t)mg_cap * W * sizeof(K); // of ONE slot; slot s starts at rx_base + s * rx_bytes
        *bucket_cap = mg_cap;
        *n_sub = tab.n_sub;
        return KTG_OK;
    }
    // keys of the exchange spill list that this rank owns: they are in the (global) sketch
    // already, so they bypass staging and sizing
    int mg_insert_spill(const void *d_keys, uint64_t n) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        if (n == 0) return KTG_OK;
        return launch_insert_keys((const K *)d_keys, n);
    }

    // Runs on send_stream (nullptr: the handle's stream), so that it can overlap the owner-side
    // work of the previous chunk; peer_rx are the slot-0 bases of all ranks.
    int mg_scatter_reads(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                         uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                         cudaStream_t send_stream, void **d_cursors, const BatchHint *hint = nullptr) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        if (!b_rx.p) return fail(KTG_ERR_INVALID, "ktg_mg_prepare first");
        if (slot >= MG_SLOTS) return fail(KTG_ERR_INVALID, "slot out of range");
        struct StreamSwap { // pack() and the scatter pass launch on `stream`
            cudaStream_t &ref, saved;
            StreamSwap(cudaStream_t &r, cudaStream_t s) : ref(r), saved(r) { if (s) ref = s; }
            ~StreamSwap() { ref = saved; }
        } swap(stream, send_stream);
        const uint32_t W = tab.world;
        mg_mode = true;
        mg_local_sketch = false;
        mgd_active = false;
        unsigned long long *cur = mg_cursors(slot);
        *d_cursors = cur;
        init_cursors_kernel<<<1, 32, 0, stream>>>(cur, W, mg_cap);
        if (first_of_batch) KTG_CUDA(cudaMemsetAsync(mg_spill_cursor(), 0, 8, stream));
        if (n_reads == 0) return KTG_OK;
        Batch bt;
        KTG_TRY(pack(d_bases, d_offsets, n_reads, total_bases, &bt, hint, hint_shift0));
        if (bt.windows == 0) return KTG_OK;
        PeerOut po{};
        po.world = W;
        po.bins_per_owner = 1;
        po.pad = mg_pad();
        // the sender's virtual position is v = owner * cap + fill; bucket `rank` of the owner's
        // slot starts at slot_base + rank * cap, so bias the base by (rank - owner) * cap
        for (uint32_t o = 0; o < W; ++o)
            po.rxb[o] = (K *)peer_rx[o] + (int64_t)slot * (int64_t)mg_slot_keys() +
                        ((int64_t)tab.rank - (int64_t)o) * (int64_t)mg_cap;
        ScatterOut so;
        so.cursors = cur;
        so.bucket_cap = mg_cap;
        so.out = nullptr;
        so.spill_out = b_mg_spill.p;
        so.spill_cursor = mg_spill_cursor();
        so.spill_cap = mg_spill_cap;
        KTG_TRY((scatter_reads_pass<BIN_OWNER, true>(bt, W, so, &po)));
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // d_bucket_ends[s]: absolute end (in keys, inside the slot) of the bucket that rank s
    // filled, i.e. s * mg_cap + fill; n_keys: the exact total.  The caller has all-reduced
    // (max) the sketch before this call, so that a flush can size the shard.
    int mg_insert_buckets(const void *d_bucket_ends, uint64_t n_keys, uint32_t slot) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        if (slot >= MG_SLOTS) return fail(KTG_ERR_INVALID, "slot out of range");
        const unsigned long long *ends = (const unsigned long long *)d_bucket_ends;
        if (n_keys == 0) return KTG_OK;
        const uint32_t W = tab.world;
        const uint64_t cap = mg_cap;
        const K *rx = (const K *)b_rx.p + (size_t)slot * mg_slot_keys();
        KTG_TRY(stage_add(n_keys, [&](uint32_t n_bins, const ScatterOut &o) -> int {
            prof.begin("scatter_received", n_keys, stream);
            // (the 256-thread tile that wins for level 2 is slower here: 1.97 vs 1.62 ms)
            const uint64_t tiles_per_bin = cap / L2S_TILE, n_tiles = tiles_per_bin * W;
            const size_t ss = ScatterSmem<K, L2S_TILE>::bytes(n_bins, false);
            int g = (int)std::min<uint64_t>(grid_for(scatter_buckets_kernel<K, 1>, L2S_THREADS, ss, props), n_tiles);
            scatter_buckets_kernel<K, 1><<<g, L2S_THREADS, ss, stream>>>(rx, ends, cap, tiles_per_bin, n_tiles, 0, mg_pad() > 1, tab, o);
            prof.end(stream);
            return KTG_OK;
        }));
        KTG_TRY(sync()); // the receive slot may be written again
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // ---- multi-GPU, direct: the sender does the owner's level-1 partition too -------------------------
    // The fused exchange above leaves the owner one more pass than a single GPU has (scatter_received:
    // 4.6 of 20.1 ms per step on C3 at N = 2).  Here the extraction kernel bins every key by (owner,
    // sub-table of the owner) -- world x n_sub bins, as many as the one-GPU build of the same table has,
    // because a shard is 1/world of it -- and writes the runs into the owner's receive bucket for
    // (this sender, that sub-table).  The owner's receive buffer then IS its level-1 stage: it goes
    // straight to the level-2 scatter (bucket q holds keys of sub-table q % n_sub) and the page sweep.
    // Needs the same geometry (n_sub, sub_log2) on every shard: they are sized from the same all-reduced
    // sketch, and the caller checks it per batch (mgd_plan) and falls back to the exchange above otherwise.
    uint64_t mgd_cap = 0;
    uint32_t mgd_n_sub = 0, mgd_sub_log2 = 0;
    bool mgd_active = false; // the last exchange was the direct one (which spill cursor mg_spill reads)
    size_t mgd_bins() const { return (size_t)tab.world * mgd_n_sub; }
    size_t mgd_slot_keys() const { return (size_t)mgd_cap * mgd_bins(); }
    int mgd_check() {
        if (tab.world > (uint32_t)MAX_P2P_WORLD) return fail(KTG_ERR_INVALID, "fused exchange needs world <= %d", MAX_P2P_WORLD);
        if ((uint64_t)tab.world * tab.n_sub > MAX_BINS) return fail(KTG_ERR_INVALID, "too many sub-tables for the direct exchange");
        return KTG_OK;
    }
    int mgd_plan(uint64_t max_windows, int *needs_realloc, uint32_t *n_sub, uint32_t *sub_log2) override {
        KTG_TRY(mgd_check());
        const uint64_t cap = bucket_cap_for(std::max<uint64_t>(max_windows, 1), tab.world * tab.n_sub);
        *needs_realloc = cap > mgd_cap || tab.n_sub != mgd_n_sub || !b_rx.p;
        *n_sub = tab.n_sub;
        *sub_log2 = tab.sub_log2;
        return KTG_OK;
    }
    int mgd_prepare(uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap) override {
        KTG_TRY(mgd_check());
        const uint32_t W = tab.world;
        const uint64_t cap = bucket_cap_for(std::max<uint64_t>(max_windows, 1), W * tab.n_sub);
        if (cap > mgd_cap || tab.n_sub != mgd_n_sub || !b_rx.p) {
            KTG_TRY(sync());
            b_rx.release();
            mgd_cap = cap;
            mgd_n_sub = tab.n_sub;
            KTG_TRY(b_rx.ensure(mgd_slot_keys() * MG_SLOTS * sizeof(K) + 64));
            mg_cap = skm_cap = 0; // the other exchanges have to size the buffer again
            mg_spill_cap = std::max<uint64_t>(1u << 20, max_windows / 4);
            KTG_TRY(b_mg_spill.ensure(mg_spill_cap * sizeof(K) + 64));
            KTG_TRY(b_mg_cur.ensure((mgd_bins() * MG_SLOTS + 2) * 8));
        }
        mgd_sub_log2 = tab.sub_log2;
        *rx_base = b_rx.p;
        *rx_bytes = mgd_slot_keys() * sizeof(K); // of ONE slot
        *bucket_cap = mgd_cap;
        return KTG_OK;
    }
    unsigned long long *mgd_cursors(uint32_t slot) { return (unsigned long long *)b_mg_cur.p + (size_t)slot * mgd_bins(); }
    unsigned long long *mgd_spill_cursor() { return (unsigned long long *)b_mg_cur.p + (size_t)MG_SLOTS * mgd_bins(); }

    int mgd_scatter_reads(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads, uint64_t total_bases,
                          void *const *peer_rx, uint32_t slot, int first_of_batch, cudaStream_t send_stream,
                          void **d_cursors, const BatchHint *hint = nullptr) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        if (!b_rx.p || !mgd_cap) return fail(KTG_ERR_INVALID, "ktg_mg_direct_prepare first");
        if (slot >= MG_SLOTS) return fail(KTG_ERR_INVALID, "slot out of range");
        if (tab.n_sub != mgd_n_sub) return fail(KTG_ERR_INVALID, "the table geometry changed since ktg_mg_direct_prepare");
        struct StreamSwap {
            cudaStream_t &ref, saved;
            StreamSwap(cudaStream_t &r, cudaStream_t s_) : ref(r), saved(r) { if (s_) ref = s_; }
            ~StreamSwap() { ref = saved; }
        } swap(stream, send_stream);
        const uint32_t W = tab.world, n_bins = (uint32_t)mgd_bins();
        mg_mode = true;
        mg_local_sketch = false;
        mgd_active = true;
        unsigned long long *cur = mgd_cursors(slot);
        *d_cursors = cur;
        // a batch may arrive in several calls (chunks of a host batch): the later ones append to the same buckets
        if (first_of_batch) {
            init_cursors_kernel<<<(n_bins + 255) / 256, 256, 0, stream>>>(cur, n_bins, mgd_cap);
            KTG_CUDA(cudaMemsetAsync(mgd_spill_cursor(), 0, 8, stream));
        }
        if (n_reads == 0) return KTG_OK;
        Batch bt;
        KTG_TRY(pack(d_bases, d_offsets, n_reads, total_bases, &bt, hint, hint_shift0));
        if (bt.windows == 0) return KTG_OK;
        PeerOut po{};
        po.world = W;
        po.bins_per_owner = mgd_n_sub;
        po.pad = 1; // runs are a dozen keys: padding them to 128 bytes would be a fifth of the traffic
        // the sender's virtual position is v = (owner * n_sub + sub) * cap + fill; in the owner's slot the
        // bucket of (this sender, sub) starts at (rank * n_sub + sub) * cap
        for (uint32_t o = 0; o < W; ++o)
            po.rxb[o] = (K *)peer_rx[o] + (int64_t)slot * (int64_t)mgd_slot_keys() +
                        ((int64_t)tab.rank - (int64_t)o) * (int64_t)mgd_n_sub * (int64_t)mgd_cap;
        ScatterOut so;
        so.cursors = cur;
        so.bucket_cap = mgd_cap;
        so.out = nullptr;
        so.spill_out = b_mg_spill.p;
        so.spill_cursor = mgd_spill_cursor();
        so.spill_cap = mg_spill_cap;
        KTG_TRY((scatter_reads_pass<BIN_OWNER_PART, true>(bt, n_bins, so, &po)));
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    // d_bucket_ends[q], q = source * n_sub + sub: absolute end (in keys, inside the slot) of that bucket,
    // q * cap + fill; n_keys: the exact total.  The caller has all-reduced (max) the sketch.
    int mgd_insert(const void *d_bucket_ends, uint64_t n_keys, uint32_t slot) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        if (slot >= MG_SLOTS) return fail(KTG_ERR_INVALID, "slot out of range");
        if (n_keys == 0) return KTG_OK;
        const unsigned long long *ends = (const unsigned long long *)d_bucket_ends;
        const K *rx = (const K *)b_rx.p + (size_t)slot * mgd_slot_keys();
        const uint32_t n_bins = (uint32_t)mgd_bins();
        KTG_TRY(flush_staged()); // (keys staged by another route go first; usually nothing)
        // size the table for everything offered so far, as flush_staged does
        double est = 0;
        KTG_TRY(hll_estimate(&est));
        const uint64_t distinct = hll_base + (uint64_t)(est * 1.10 / tab.world) + 64;
        occupied_ub = distinct;
        if ((double)distinct > LOAD_MAX * (double)tab.capacity()) {
            const uint64_t need = std::max<uint64_t>((uint64_t)((double)distinct / LOAD_TARGET) + 1, 2 * tab.capacity());
            KTG_TRY(grow_to(need));
        }
        touch();
        const bool same = tab.n_sub == mgd_n_sub && tab.sub_log2 == mgd_sub_log2;
        if (same && use_pages(n_keys)) {
            // page-bucket overflow goes to the stage's spill list
            stage_spill_cap = n_keys + 64;
            KTG_TRY(b_spill.ensure(stage_spill_cap * sizeof(K) + 64));
            KTG_CUDA(cudaMemsetAsync(stage_spill_cursor(), 0, 8, stream));
            KTG_TRY(pages_open(n_keys));
            KTG_TRY(pages_scatter(n_bins, mgd_cap, n_keys, ends, rx, mgd_n_sub));
            KTG_TRY(pages_update(n_keys));
            KTG_TRY(launch_insert((const K *)b_spill.p, stage_spill_cap, nullptr, 0, 0, stage_spill_cursor()));
        }
        else KTG_TRY(launch_insert(rx, n_keys, ends, mgd_cap, n_bins)); // plain arrays of keys, any geometry
        KTG_TRY(sync()); // the receive slot may be written again
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    int mg_sketch(void **d_regs, uint32_t *n_regs) override {
        *d_regs = b_hll.p;
        *n_regs = HLL_M;
        return KTG_OK;
    }
    // fold an (all-reduced) copy of the sketch back in; atomic, because the sender-side
    // kernel of the next chunk may be updating the sketch at the same time
    int mg_merge_sketch(const void *d_regs) override {
        hll_merge_kernel<<<HLL_M / 256, 256, 0, stream>>>((uint32_t *)b_hll.p, (const uint32_t *)d_regs);
        KTG_CUDA(cudaGetLastError());
        return KTG_OK;
    }

    int mg_spill(void **d_keys, uint64_t *n) override {
        unsigned long long v = 0;
        *d_keys = b_mg_spill.p;
        *n = 0;
        if (!b_mg_cur.p) return KTG_OK;
        KTG_CUDA(cudaMemcpyAsync(&v, mgd_active ? mgd_spill_cursor() : mg_spill_cursor(), 8, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        if (v > mg_spill_cap) {
            deferred_error = KTG_ERR_TABLE_FULL;
            return fail(KTG_ERR_TABLE_FULL, "%llu keys overflowed the exchange spill list", v - mg_spill_cap);
        }
        *n = v;
        return KTG_OK;
    }

    // ---- multi-GPU, fused, in super-k-mer records (superkmer.cuh) ---------------------------
    // Same protocol as above with 16-byte records instead of keys: the receive buffer (same
    // allocation, same two slots) holds one bucket of skm_cap RECORDS per source rank, the sender
    // also counts the k-mers it sent to every owner (the owner needs an upper bound to stage
    // them), and the owner sketches the keys itself while it unrolls the records, so no sketch
    // is exchanged.  Owner of a k-mer = skm_owner(minimizer), see skm_owner_of.
    DeviceBuf b_skm_cnt, b_skm_part;
    uint64_t skm_cap = 0, skm_spill_cap = 0;
    bool mg_local_sketch = false;
    static constexpr uint32_t SKM_PAD = 128 / sizeof(u128);
    size_t skm_slot_records() const { return (size_t)skm_cap * tab.world; }
    static constexpr uint32_t SKM_CNT_STRIDE = MAX_P2P_WORLD + 8; // key counts per owner, then the spill count
    unsigned long long *skm_key_counts(uint32_t slot) { return (unsigned long long *)b_skm_cnt.p + (size_t)slot * SKM_CNT_STRIDE; }
    unsigned long long *skm_scalar(uint32_t i) { return (unsigned long long *)b_skm_cnt.p + (size_t)MG_SLOTS * SKM_CNT_STRIDE + i; }

    int skm_check() {
        if (sizeof(K) != 8 || !skm_supported(k))
            return fail(KTG_ERR_INVALID, "the super-k-mer exchange needs %u <= k <= %u", SKM_K_MIN, SKM_K_MAX);
        if (tab.world > (uint32_t)MAX_P2P_WORLD) return fail(KTG_ERR_INVALID, "fused exchange needs world <= %d", MAX_P2P_WORLD);
        return KTG_OK;
    }
    int skm_geometry(uint64_t max_windows, uint64_t *cap) {
        KTG_TRY(skm_check());
        const uint32_t W = tab.world;
        // random sequence cuts 2.8 records per 16 windows (0.17 per window); 0.25 leaves a third
        // of headroom, low-complexity reads need fewer, and whatever does not fit takes the
        // spill route.  Every (tile, owner) run may be rounded up by SKM_PAD - 1 fillers.
        const uint64_t fillers = (max_windows / (SKM_W * SCATTER_THREADS) + 1) * W * (SKM_PAD - 1);
        *cap = bucket_cap_for(std::max<uint64_t>(max_windows / 4 + fillers, 1), W);
        return KTG_OK;
    }
    int mg_skm_plan(uint64_t max_windows, int *needs_realloc) override {
        uint64_t cap = 0;
        KTG_TRY(skm_geometry(max_windows, &cap));
        *needs_realloc = cap > skm_cap || !b_rx.p;
        return KTG_OK;
    }
    int mg_skm_prepare(uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap) override {
        uint64_t cap = 0;
        KTG_TRY(skm_geometry(max_windows, &cap));
        const uint32_t W = tab.world;
        if (cap > skm_cap || !b_rx.p) {
            KTG_TRY(sync());
            b_rx.release();
            KTG_TRY(b_rx.ensure((size_t)cap * W * MG_SLOTS * sizeof(u128) + 64));
            skm_cap = cap;
            mg_cap = mgd_cap = 0; // the key exchanges have to size the buffer again
            skm_spill_cap = std::max<uint64_t>(1u << 20, max_windows / 16);
            KTG_TRY(b_mg_spill.ensure(skm_spill_cap * sizeof(u128) + 64));
            KTG_TRY(b_mg_cur.ensure(((size_t)W * MG_SLOTS + 2) * 8));
            KTG_TRY(b_skm_cnt.ensure(((size_t)MG_SLOTS * SKM_CNT_STRIDE + 4) * 8));
        }
        *rx_base = b_rx.p;
        *rx_bytes = (size_t)skm_cap * W * sizeof(u128); // of ONE slot
        *bucket_cap = skm_cap;
        return KTG_OK;
    }

    int mg_skm_scatter_reads(const uint8_t *d_bases, const uint64_t *d_offsets, uint64_t n_reads,
                             uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                             cudaStream_t send_stream, void **d_cursors, void **d_key_counts,
                             const BatchHint *hint = nullptr) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        KTG_TRY(skm_check());
        if (!b_rx.p || !skm_cap) return fail(KTG_ERR_INVALID, "ktg_mg_skm_prepare first");
        if (slot >= MG_SLOTS) return fail(KTG_ERR_INVALID, "slot out of range");
        struct StreamSwap {
            cudaStream_t &ref, saved;
            StreamSwap(cudaStream_t &r, cudaStream_t s) : ref(r), saved(r) { if (s) ref = s; }
            ~StreamSwap() { ref = saved; }
        } swap(stream, send_stream);
        const uint32_t W = tab.world;
        mg_mode = true;
        mg_local_sketch = true;
        unsigned long long *cur = mg_cursors(slot), *kc = skm_key_counts(slot);
        *d_cursors = cur;
        *d_key_counts = kc;
        init_cursors_kernel<<<1, 32, 0, stream>>>(cur, W, skm_cap);
        KTG_CUDA(cudaMemsetAsync(kc, 0, SKM_CNT_STRIDE * 8, stream));
        if (first_of_batch) KTG_CUDA(cudaMemsetAsync(mg_spill_cursor(), 0, 8, stream));
        if (n_reads == 0) return KTG_OK;
        Batch bt;
        KTG_TRY(pack(d_bases, d_offsets, n_reads, total_bases, &bt, hint, hint_shift0));
        if (bt.windows == 0) return KTG_OK;
        SkmView sv{};
        sv.v = bt.v;
        if (bt.v.ulen) {
            sv.ipr = (bt.v.ulen - k + 1 + SKM_W - 1) / SKM_W;
            sv.n_items = n_reads * sv.ipr;
        }
        else sv.n_items = bt.v.n_words * 2;
        PeerOut po{};
        po.world = W;
        po.bins_per_owner = 1;
        po.pad = SKM_PAD;
        for (uint32_t o = 0; o < W; ++o)
            po.rxb[o] = (u128 *)peer_rx[o] + (int64_t)slot * (int64_t)skm_slot_records() +
                        ((int64_t)tab.rank - (int64_t)o) * (int64_t)skm_cap;
        ScatterOut so;
        so.cursors = cur;
        so.bucket_cap = skm_cap;
        so.out = nullptr;
        so.spill_out = b_mg_spill.p;
        so.spill_cursor = mg_spill_cursor();
        so.spill_cap = skm_spill_cap;
        const size_t ss = ScatterSmem<u128, SCATTER_TILE>::bytes(W, false);
        const uint64_t n_tiles = std::max<uint64_t>(1, (sv.n_items + SCATTER_THREADS * SKM_IPL - 1) / (SCATTER_THREADS * SKM_IPL));
        int g = (int)std::min<uint64_t>(grid_for(scatter_superkmers_kernel, SCATTER_THREADS, ss, props), n_tiles);
        if (tune.p2p_ctas > 0) g = (int)std::min<uint64_t>(g, (uint64_t)props.sms * tune.p2p_ctas);
        prof.begin("scatter_superkmers_p2p", bt.windows, stream);
        scatter_superkmers_kernel<<<g, SCATTER_THREADS, ss, stream>>>(sv, k, W, so, po, kc);
        prof.end(stream);
        KTG_CUDA(cudaGetLastError());
        // key_counts[world]: the records this rank has spilled so far in this batch
        KTG_CUDA(cudaMemcpyAsync(kc + W, mg_spill_cursor(), 8, cudaMemcpyDeviceToDevice, stream));
        return KTG_OK;
    }

    // records [q*cap, ends[q]) for q < n_buckets -> canonical k-mers -> staged by sub-table (+ sketch),
    // one kernel (superkmer.cuh: scatter_records_kernel)
    int skm_unroll(const u128 *rx, const unsigned long long *ends, uint64_t cap, uint32_t n_buckets, uint64_t n_keys_ub) {
        if constexpr (sizeof(K) == 8) {
            const uint64_t n_rtiles = (cap / SREC_THREADS + 1) * n_buckets;
            if ((double)n_rtiles >= 4.0e9) return fail(KTG_ERR_INVALID, "batch too large: split it");
            KTG_TRY(stage_add(n_keys_ub, [&](uint32_t n_bins, const ScatterOut &o) -> int {
                const size_t ss = ((ScatterSmem<K, SREC_TILE>::bytes(n_bins, false) + 15) & ~(size_t)15) + (size_t)SREC_STAGE * 8;
                auto launch = [&](auto kern) {
                    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(((ScatterSmem<K, SREC_TILE>::bytes(MAX_BINS, false) + 15) & ~(size_t)15) + (size_t)SREC_STAGE * 8));
                    int g = (int)std::min<uint64_t>(grid_for(kern, SREC_THREADS, ss, props), std::max<uint64_t>(n_rtiles, 1));
                    kern<<<g, SREC_THREADS, ss, stream>>>(rx, ends, cap, n_buckets, k, tab, o, (uint32_t *)b_hll.p);
                };
                prof.begin("scatter_records", n_keys_ub, stream);
                if (rc) launch(scatter_records_kernel<true>);
                else launch(scatter_records_kernel<false>);
                prof.end(stream);
                return KTG_OK;
            }));
            KTG_CUDA(cudaGetLastError());
            return KTG_OK;
        }
        else return fail(KTG_ERR_INVALID, "super-k-mer records need k <= 31");
    }

    // d_bucket_ends[s]: absolute end (in records, inside the slot) of the bucket that rank s
    // filled; n_keys_ub: the k-mers they hold (the sum of what the senders counted)
    int mg_skm_insert_buckets(const void *d_bucket_ends, uint64_t n_keys_ub, uint32_t slot) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        KTG_TRY(skm_check());
        if (slot >= MG_SLOTS) return fail(KTG_ERR_INVALID, "slot out of range");
        if (n_keys_ub == 0) return KTG_OK;
        mg_mode = mg_local_sketch = true;
        const u128 *rx = (const u128 *)b_rx.p + (size_t)slot * skm_slot_records();
        KTG_TRY(skm_unroll(rx, (const unsigned long long *)d_bucket_ends, skm_cap, tab.world, n_keys_ub));
        KTG_TRY(sync()); // the receive slot may be written again
        return KTG_OK;
    }

    int mg_skm_spill(void **d_records, uint64_t *n) override {
        unsigned long long v = 0;
        *d_records = b_mg_spill.p;
        *n = 0;
        if (!b_mg_cur.p) return KTG_OK;
        KTG_CUDA(cudaMemcpyAsync(&v, mg_spill_cursor(), 8, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        if (v > skm_spill_cap) {
            deferred_error = KTG_ERR_TABLE_FULL;
            return fail(KTG_ERR_TABLE_FULL, "%llu records overflowed the exchange spill list", v - skm_spill_cap);
        }
        *n = v;
        return KTG_OK;
    }

    // group n records by owner (owner-major copy in a buffer of the handle), counts[world]
    int mg_skm_partition_records(const void *d_records, uint64_t n, void **d_out, uint64_t *counts) override {
        KTG_TRY(skm_check());
        const uint32_t W = tab.world;
        for (uint32_t i = 0; i < W; ++i) counts[i] = 0;
        *d_out = nullptr;
        if (n == 0) return KTG_OK;
        KTG_TRY(b_skm_part.ensure(n * sizeof(u128) + (size_t)2 * MAX_P2P_WORLD * 8 + 64));
        u128 *out = (u128 *)b_skm_part.p;
        unsigned long long *cnt = (unsigned long long *)(out + n), *cur = cnt + MAX_P2P_WORLD;
        KTG_CUDA(cudaMemsetAsync(cnt, 0, MAX_P2P_WORLD * 8, stream));
        const int g = (int)std::min<uint64_t>((n + 255) / 256, (uint64_t)props.sms * 8);
        count_record_owners_kernel<<<g, 256, 0, stream>>>((const u128 *)d_records, n, k, W, cnt);
        unsigned long long h[MAX_P2P_WORLD], start[MAX_P2P_WORLD];
        KTG_CUDA(cudaMemcpyAsync(h, cnt, MAX_P2P_WORLD * 8, cudaMemcpyDeviceToHost, stream));
        KTG_TRY(sync());
        unsigned long long run = 0;
        for (uint32_t i = 0; i < (uint32_t)MAX_P2P_WORLD; ++i) {
            start[i] = run;
            if (i < W) {
                counts[i] = h[i];
                run += h[i];
            }
        }
        KTG_CUDA(cudaMemcpyAsync(cur, start, MAX_P2P_WORLD * 8, cudaMemcpyHostToDevice, stream));
        scatter_record_owners_kernel<<<g, 256, 0, stream>>>((const u128 *)d_records, n, k, W, cur, out);
        KTG_CUDA(cudaGetLastError());
        KTG_TRY(sync()); // `start` is on this stack frame
        *d_out = out;
        return KTG_OK;
    }

    // a flat array of records this rank owns (what the spill route delivered)
    int mg_skm_insert_records(const void *d_records, uint64_t n) override {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error");
        KTG_TRY(skm_check());
        if (n == 0) return KTG_OK;
        KTG_TRY(b_skm_cnt.ensure(((size_t)MG_SLOTS * SKM_CNT_STRIDE + 4) * 8));
        mg_mode = mg_local_sketch = true;
        const unsigned long long end = n;
        KTG_CUDA(cudaMemcpyAsync(skm_scalar(1), &end, 8, cudaMemcpyHostToDevice, stream));
        KTG_TRY(sync()); // `end` is on this stack frame
        const uint64_t cap = (n + SREC_THREADS - 1) / SREC_THREADS * SREC_THREADS;
        KTG_TRY(skm_unroll((const u128 *)d_records, skm_scalar(1), cap, 1, n * SKM_W));
        KTG_TRY(sync()); // the caller may free the records
        return KTG_OK;
    }

    // owner of a k-mer in the super-k-mer exchange
    uint32_t skm_owner_of(uint64_t hi, uint64_t lo) override {
        (void)hi;
        if (sizeof(K) != 8 || !skm_supported(k)) return 0;
        return skm_owner(skm_minimizer_of_kmer(lo, k), tab.world);
    }

    uint32_t owner_of(uint64_t hi, uint64_t lo) override {
        K key = T::make(hi, lo);
        if (rc) {
            K r = revcomp(key, k);
            if (r < key) key = r;
        }
        return place_of(T::place_hash(key), tab.world, tab.n_sub).owner;
    }

    int info(ktg_info *out) override {
        memset(out, 0, sizeof *out);
        KTG_TRY(flush_staged());
        uint64_t occ = 0;
        KTG_TRY(count_occupied(&occ));
        out->capacity_slots = tab.capacity();
        out->occupied_slots = occ;
        out->table_bytes = tab.bytes();
        out->n_sub_tables = tab.n_sub;
        out->slot_bytes = Table<K>::SLOT_BYTES;
        out->windows_inserted = windows_inserted;
        out->kernel_launches = prof.total_launches;
        out->grow_events = grow_events;
        out->partitioned = use_partition();
        out->page_updates = page_updates;
        out->n_pages = (uint32_t)std::min<uint64_t>(tab.n_pages(), 0xFFFFFFFFull);
        return KTG_OK;
    }
};

} // namespace ktg
