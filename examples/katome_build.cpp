// katome_build -- the GIR build stage from a C++ host, through include/katome_gpu.hpp: what
// BasicAsm::assemble_with_gir does up to the conversion (asm/basic_assembler.rs:31-43):
//   set k, T::create(files, ft, reverse_complement, threshold), stats, [remove_weak_edges],
//   [standardize_edges], and the graph that Convert::create_from would load.
// usage: katome_build [--devices 0,1,...] [--dump <file>] <k> <reverse_complement 0|1> <weak-edge threshold,
//        0 = keep all> <genome length, 0 = no standardization> <fastq file>...
//   --devices  the GPUs the ONE table is sharded over (a device may repeat); default: the current one
//   --dump     write "sequence <k-mer> weight <w>" lines sorted by k-mer (hs_gir.rs:288-290),
//              preceded by "bytes nodes edges": byte-identical whatever --devices is
// prints one line: bytes nodes edges max_w sum_w max_in max_out sources sinks graph_nodes graph_edges
// exit code: 0 ok, 2 usage, 3 a panic of the reference (message on stderr)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "katome_gpu.hpp"

int main(int argc, char **argv) {
    std::string dump;
    while (argc > 2 && argv[1][0] == '-' && argv[1][1] == '-') {
        if (!std::strcmp(argv[1], "--devices")) {
            std::vector<int32_t> ids;
            for (const char *p = argv[2]; *p;) {
                char *end = nullptr;
                ids.push_back((int32_t)std::strtol(p, &end, 10));
                p = *end == ',' ? end + 1 : end;
            }
            katome::GpuGIR::set_devices(ids);
        }
        else if (!std::strcmp(argv[1], "--dump")) dump = argv[2];
        else break;
        argv += 2;
        argc -= 2;
    }
    if (argc < 6) {
        std::fprintf(stderr, "usage: %s k rc threshold genome_len file...\n", argv[0]);
        return 2;
    }
    const uint32_t k = (uint32_t)std::atoi(argv[1]);
    const bool rc = std::atoi(argv[2]) != 0;
    const uint32_t threshold = (uint32_t)std::atoi(argv[3]);
    const uint64_t genome_len = std::strtoull(argv[4], nullptr, 10);
    std::vector<std::string> files(argv + 5, argv + argc);
    try {
        auto [gir, bytes] = katome::GpuGIR::create(files, katome::InputFileType::Fastq, rc, threshold, k);
        if (threshold) gir.remove_weak_edges(threshold);
        if (genome_len) gir.standardize_edges(genome_len, k, threshold);
        const katome::CollectionStats s = gir.stats();
        const katome::Graph g = gir.to_graph();
        std::printf("%llu %llu %llu %llu %llu %llu %llu %llu %llu %zu %zu\n", (unsigned long long)bytes,
                    (unsigned long long)s.node_count, (unsigned long long)s.edge_count,
                    (unsigned long long)s.max_edge_weight, (unsigned long long)s.sum_edge_weight,
                    (unsigned long long)s.max_in_degree, (unsigned long long)s.max_out_degree,
                    (unsigned long long)s.incoming_vert_count, (unsigned long long)s.outgoing_vert_count,
                    g.node_lo.size(), g.weight.size());
        for (size_t e = 0; e < g.src.size(); ++e)
            if (g.src[e] >= g.node_lo.size() || g.dst[e] >= g.node_lo.size()) return 4;
        if (!dump.empty()) {
            std::vector<uint64_t> hi, lo;
            std::vector<uint32_t> w;
            gir.sorted_edges(hi, lo, w);
            FILE *f = std::fopen(dump.c_str(), "w");
            if (!f) return 5;
            std::fprintf(f, "%llu %llu %llu\n", (unsigned long long)bytes, (unsigned long long)s.node_count,
                         (unsigned long long)s.edge_count);
            std::string kmer(k, 'A');
            for (size_t e = 0; e < w.size(); ++e) {
                for (uint32_t j = 0; j < k; ++j) {
                    const uint32_t sh = 2 * (k - 1 - j);
                    const uint64_t word = sh >= 64 ? hi[e] >> (sh - 64) : lo[e] >> sh;
                    kmer[j] = "ACGT"[word & 3];
                }
                std::fprintf(f, "sequence %s weight %u\n", kmer.c_str(), w[e]);
            }
            std::fclose(f);
        }
    } catch (const katome::Panic &p) {
        std::fprintf(stderr, "panic (%d): %s\n", p.code, p.what());
        return 3;
    }
    return 0;
}
