// katome_gpu.hpp -- header-only C++17 host side above the C ABI (katome_gpu.h), with the names,
// argument meaning and error behaviour of the reference's traits for this path, so that a C++ host
// (or the Rust wrapper of INTEGRATION.md, of which this is the twin) reads like katome's own code:
//
//   Init::init(edges, nodes, ft)                     algorithms/builder.rs:19-25
//   Build::add_read_fastaq / create                  algorithms/builder.rs:28-55, hm_gir.rs:39-87
//   Clean::remove_weak_edges / remove_single_vertices algorithms/pruner.rs:29-34, 95-119
//   Standardizable::standardize_edges                algorithms/standardizer.rs:33-70
//   Stats<CollectionStats>::stats                    stats/collections.rs:137-208
//   Convert::create_from (its input)                 collections/girs/hm_gir.rs:156-226
//
// The reference panics where this throws katome::Panic with the same text ("Read is too short!",
// hm_gir.rs:40; the file errors of builder.rs:57-77,148).  There is no CPU fallback: without a CUDA
// device the constructor throws.  Citations are relative to /root/reference/src/katome/.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "katome_gpu.h"

namespace katome {

struct Panic : std::runtime_error {
    int code;
    Panic(int c, const std::string &what) : std::runtime_error(what), code(c) {}
};

enum class InputFileType { Fastq = KTG_FASTQ, Fasta = KTG_FASTA }; // config.rs:5-14 (BFCounter: create_bfc below)

// stats/collections.rs:38-57; the capacity field has no meaning here and floats are derived
struct CollectionStats {
    uint64_t node_count = 0, edge_count = 0;
    uint64_t max_edge_weight = 0, sum_edge_weight = 0;
    uint64_t max_in_degree = 0, max_out_degree = 0;
    uint64_t incoming_vert_count = 0, outgoing_vert_count = 0;
    double avg_edge_weight() const { return edge_count ? (double)sum_edge_weight / (double)edge_count : 0.0; }
    double avg_out_degree() const { return node_count ? (double)edge_count / (double)node_count : 0.0; }
};

// what Convert::create_from bulk-loads into petgraph (hm_gir.rs:156-226): sorted nodes, edges as
// (source index, target index, weight) and their compress_edge bytes (compress.rs:250-271)
struct Graph {
    std::vector<uint64_t> node_hi, node_lo;
    std::vector<uint64_t> src, dst;
    std::vector<uint32_t> weight;
    std::vector<uint8_t> edge_bytes;
    uint32_t edge_record_bytes = 0;
};

class GpuGIR {
  public:
    // The GPUs every collection created afterwards is built on -- katome's settings are process
    // globals too (config.rs, prelude.rs:21-25).  Empty or one entry: one GPU.  Several: ONE table
    // hash-sharded over them (ktg_config.n_devices); create / add_read_fastaq / stats / to_graph are
    // called exactly as before, which is the point: Build::create (builder.rs:42-54) stays one call.
    static std::vector<int32_t> &devices() {
        static std::vector<int32_t> d;
        return d;
    }
    static void set_devices(const std::vector<int32_t> &ids) { devices() = ids; }

    // set_global_k_sizes (prelude.rs:34-43) + Init::init: k is per collection, not a process global
    static GpuGIR init(uint32_t k, bool reverse_complement, uint64_t edges_count = 0, int device = -1) {
        return GpuGIR(k, reverse_complement, edges_count, device);
    }
    // Build::create (builder.rs:42-54): -> (collection, total accepted bytes)
    static std::pair<GpuGIR, uint64_t> create(const std::vector<std::string> &input_files, InputFileType ft,
                                              bool reverse_complement, uint32_t /*minimal_weight_threshold*/,
                                              uint32_t k, int device = -1) {
        GpuGIR g(k, reverse_complement, 0, device);
        std::vector<const char *> paths;
        for (const std::string &p : input_files) paths.push_back(p.c_str());
        uint64_t total = 0;
        g.check(ktg_create_from_files(g.h_, paths.data(), (uint32_t)paths.size(), (int)ft, &total));
        return {std::move(g), total};
    }
    // create_bfc (builder.rs:79-115): pre-counted k-mers, threshold pre-filter
    static std::pair<GpuGIR, uint64_t> create_bfc(const std::vector<std::string> &input_files, bool reverse_complement,
                                                  uint32_t minimal_weight_threshold, uint32_t k, int device = -1) {
        GpuGIR g(k, reverse_complement, 0, device);
        std::vector<const char *> paths;
        for (const std::string &p : input_files) paths.push_back(p.c_str());
        uint64_t total = 0;
        g.check(ktg_create_from_bfc_files(g.h_, paths.data(), (uint32_t)paths.size(), minimal_weight_threshold, &total));
        return {std::move(g), total};
    }

    GpuGIR(GpuGIR &&o) noexcept : h_(o.h_), rc_(o.rc_), bases_(std::move(o.bases_)), offsets_(std::move(o.offsets_)) { o.h_ = nullptr; }
    GpuGIR &operator=(GpuGIR &&o) noexcept {
        if (this != &o) {
            if (h_) ktg_destroy(h_);
            h_ = o.h_;
            rc_ = o.rc_;
            bases_ = std::move(o.bases_);
            offsets_ = std::move(o.offsets_);
            o.h_ = nullptr;
        }
        return *this;
    }
    GpuGIR(const GpuGIR &) = delete;
    GpuGIR &operator=(const GpuGIR &) = delete;
    ~GpuGIR() {
        if (h_) ktg_destroy(h_);
    }

    // Build::add_read_fastaq (builder.rs:30): the caller has already applied create_fastq's accept
    // rule in the reference; here the library applies it again per batch (a read with a byte outside
    // "ACGT" is dropped whole, builder.rs:155).  `reverse_complement` must be what init() was given.
    // Reads are gathered and handed over in batches of ~64 MiB; flush() (or any query) sends the rest.
    void add_read_fastaq(const uint8_t *read, size_t len, bool reverse_complement) {
        if (reverse_complement != rc_) throw Panic(KTG_ERR_INVALID, "reverse_complement differs from init()");
        bases_.insert(bases_.end(), read, read + len);
        offsets_.push_back((uint64_t)bases_.size());
        if (bases_.size() >= (size_t)64 << 20) flush();
    }
    void add_read_fastaq(const std::string &read, bool reverse_complement) {
        add_read_fastaq((const uint8_t *)read.data(), read.size(), reverse_complement);
    }
    void flush() {
        if (offsets_.size() > 1) {
            const int rc = ktg_add_reads(h_, bases_.data(), offsets_.data(), offsets_.size() - 1, nullptr, nullptr);
            bases_.clear();
            offsets_.assign(1, 0);
            check(rc);
        }
    }

    // Clean (pruner.rs:29-34)
    void remove_weak_edges(uint32_t threshold) {
        flush();
        check(ktg_remove_weak_edges(h_, threshold));
    }
    void remove_single_vertices() {
        flush();
        check(ktg_remove_single_vertices(h_));
    }
    // Standardizable (standardizer.rs:35-36)
    void standardize_edges(uint64_t original_genome_length, uint64_t k, uint32_t threshold) {
        flush();
        check(ktg_standardize_edges(h_, original_genome_length, k, threshold));
    }
    // Stats<CollectionStats> (stats/mod.rs:6)
    CollectionStats stats() {
        flush();
        ktg_stats s{};
        check(ktg_collection_stats(h_, &s));
        CollectionStats c;
        c.node_count = s.node_count;
        c.edge_count = s.edge_count;
        c.max_edge_weight = s.max_edge_weight;
        c.sum_edge_weight = s.sum_edge_weight;
        c.max_in_degree = s.max_in_degree;
        c.max_out_degree = s.max_out_degree;
        c.incoming_vert_count = s.incoming_vert_count;
        c.outgoing_vert_count = s.outgoing_vert_count;
        return c;
    }
    // node_count / edge_count alone (the GIR-level stats, stats/collections.rs:190-208)
    std::pair<uint64_t, uint64_t> counts() {
        flush();
        uint64_t n = 0, e = 0;
        check(ktg_counts(h_, &n, &e));
        return {n, e};
    }
    // the input of Convert::create_from (girs/mod.rs:26-29)
    Graph to_graph() {
        flush();
        uint64_t nn = 0, ne = 0;
        check(ktg_graph_prepare(h_, &nn, &ne)); // built on the device once; the export below only copies
        Graph g;
        g.edge_record_bytes = ktg_edge_record_bytes(h_);
        g.node_hi.resize(nn);
        g.node_lo.resize(nn);
        g.src.resize(ne);
        g.dst.resize(ne);
        g.weight.resize(ne);
        g.edge_bytes.resize((size_t)ne * g.edge_record_bytes);
        check(ktg_export_graph(h_, g.node_hi.data(), g.node_lo.data(), nn, g.src.data(), g.dst.data(), g.weight.data(),
                               g.edge_bytes.data(), ne));
        return g;
    }
    // the edges sorted by k-mer with their weights (the dump that must be byte-identical
    // on 1, 2, 4 and 8 GPUs); key_hi is all zero for k <= 32
    void sorted_edges(std::vector<uint64_t> &key_hi, std::vector<uint64_t> &key_lo, std::vector<uint32_t> &weight) {
        flush();
        uint64_t n = 0;
        check(ktg_export_edges(h_, nullptr, nullptr, nullptr, 0, 1, &n));
        key_hi.assign(n, 0);
        key_lo.assign(n, 0);
        weight.assign(n, 0);
        if (n) check(ktg_export_edges(h_, key_hi.data(), key_lo.data(), weight.data(), n, 1, &n));
    }
    ktg_builder *handle() { return h_; }

  private:
    GpuGIR(uint32_t k, bool reverse_complement, uint64_t edges_count, int device) : rc_(reverse_complement), offsets_(1, 0) {
        ktg_config cfg{};
        cfg.abi_version = KTG_ABI_VERSION;
        cfg.k = k;
        cfg.reverse_complement = reverse_complement ? 1u : 0u;
        cfg.device = device;
        cfg.capacity_hint_edges = edges_count;
        cfg.world_size = 1;
        cfg.rank = 0;
        const std::vector<int32_t> ids = devices(); // a copy: the handle reads it during ktg_create only
        if (!ids.empty()) {
            cfg.n_devices = (uint32_t)ids.size();
            cfg.device_ids = ids.data();
        }
        const int rc = ktg_create(&cfg, &h_);
        if (rc != KTG_OK) {
            h_ = nullptr;
            throw Panic(rc, ktg_last_error());
        }
    }
    void check(int rc) const {
        if (rc != KTG_OK) throw Panic(rc, ktg_last_error());
    }
    ktg_builder *h_ = nullptr;
    bool rc_ = false;
    std::vector<uint8_t> bases_;
    std::vector<uint64_t> offsets_;
};

} // namespace katome
