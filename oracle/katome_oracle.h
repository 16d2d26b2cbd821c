/*
 * katome_oracle.h -- CPU restatement of katome's De Bruijn graph build stage.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (katome_b200/, include/)
 * may include, link or call this.  Allowed users: tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg.
 *
 * Parity status: pinned against every known-answer vector the reference's own
 * tests hold for this path (see tests/test_oracle_golden.py and
 * tests/golden/reference_pinned.json); the reverse_complement=true path is
 * NOT tested upstream, so for rc=true the oracle is pinned by the code it
 * restates only ("parity unpinned by reference tests" for rc=true).
 *
 * The reference (fuine/katome, Rust) cannot be compiled here (no rustc/cargo,
 * un-vendored crates), so this is a "port", never "reference".
 *
 * All file:line citations are relative to /root/reference/.
 */
#ifndef KATOME_ORACLE_H
#define KATOME_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ko_gir ko_gir;

/* error codes */
#define KO_OK 0
#define KO_ERR_SHORT_READ 1 /* hm_gir.rs:40 "Read is too short!" */
#define KO_ERR_BAD_K 2      /* prelude.rs:35, compress.rs:19 */
#define KO_ERR_IO 3         /* builder.rs:57-77,148,153 */
#define KO_ERR_BAD_RECORD 4 /* bio 0.10 reader error -> unwrap() panic builder.rs:153 */
#define KO_ERR_DEGENERATE 5 /* standardizer.rs:123-127 with G<k or s==l */

/* ---- collection life cycle (Init / Build, builder.rs:19-55) ---- */
ko_gir *ko_new(int k);
void ko_free(ko_gir *g);
int ko_k(const ko_gir *g);

/* add_read_fastaq (hm_gir.rs:39-87): read must already have passed the ACGT
 * filter.  Returns KO_ERR_SHORT_READ if len < k. */
int ko_add_read_fastaq(ko_gir *g, const uint8_t *read, size_t len, int reverse_complement);

/* The loop body of create_fastq (builder.rs:152-160) over a batch of reads:
 * reject reads with any byte outside "ACGT", total += len, add_read_fastaq. */
int ko_add_reads(ko_gir *g, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                 int reverse_complement, uint64_t *accepted_reads, uint64_t *accepted_bytes);

/* Build::create (builder.rs:42-54) for Fastq (file_type 0) / Fasta (1). */
int ko_create_from_files(ko_gir *g, const char *const *paths, int n_paths, int file_type,
                         int reverse_complement, uint64_t *accepted_reads,
                         uint64_t *accepted_bytes);

/* BFCounter input (builder.rs:79-115, pt_graph.rs:201-213,318-329): one pre-counted k-mer. */
int ko_add_read_bfc(ko_gir *g, const uint8_t *kmer, size_t len, uint32_t weight, int reverse_complement);
int ko_create_from_bfc_files(ko_gir *g, const char *const *paths, int n_paths, int reverse_complement,
                             uint32_t minimal_weight_threshold, uint64_t *accepted_kmers,
                             uint64_t *accepted_bytes);

/* ---- Stats (stats/collections.rs:190-208) ---- */
void ko_counts(const ko_gir *g, uint64_t *nodes, uint64_t *edges);

/* CollectionStats as computed for the converted PtGraph
 * (stats/collections.rs:137-168): out[0]=node_count out[1]=edge_count
 * out[2]=max_edge_weight out[3]=sum_edge_weight out[4]=max_in_degree
 * out[5]=max_out_degree out[6]=incoming_vert_count (in-degree 0)
 * out[7]=outgoing_vert_count (out-degree 0) */
void ko_collection_stats(const ko_gir *g, uint64_t out[8]);

/* ---- Clean (pruner.rs:95-119, 127-157; edges.rs:51-58) ---- */
void ko_remove_weak_edges(ko_gir *g, uint32_t threshold);
void ko_remove_single_vertices(ko_gir *g);

/* ---- Standardizable::standardize_edges (standardizer.rs:42-70, 123-127) ---- */
int ko_standardize_edges(ko_gir *g, uint64_t genome_len, uint64_t k, uint32_t threshold);
/* the arithmetic alone, on a plain weight array (for the reference's KAT
 * standardizer.rs:254-279): rewrites w[] in place, returns number kept and
 * compacts survivors to the front preserving order. */
uint64_t ko_standardize_weights(uint32_t *w, uint64_t n, uint64_t genome_len, uint64_t k,
                                uint32_t threshold, int *err);

/* ---- export: edges sorted by k-mer; key = sum code(w[j]) * 4^(k-1-j) as a
 * 128-bit integer (hi, lo).  Returns number of edges (writes at most cap). */
uint64_t ko_export_edges(const ko_gir *g, uint64_t *key_hi, uint64_t *key_lo, uint32_t *weight,
                         uint64_t cap);
/* nodes ((k-1)-mers) sorted, same integer convention */
uint64_t ko_export_nodes(const ko_gir *g, uint64_t *key_hi, uint64_t *key_lo, uint64_t cap);

/* order-independent multiset digest: out[0]=D out[1]=|E| out[2]=sum w mod 2^64
 * out[3]=max w, D = sum_e splitmix64(splitmix64(hi)^lo) * (2w+1) mod 2^64 */
void ko_digest(const ko_gir *g, uint64_t out[4]);

/* "sequence <kmer> weight <w>\n" lines (format of hs_gir.rs:288-290), sorted by
 * k-mer.  Returns bytes needed; writes at most cap bytes. */
uint64_t ko_dump(const ko_gir *g, char *buf, uint64_t cap);

/* ---- "optimistic CPU" (katome_oracle_mt.c): NOT the reference's work shape.  The same edge
 * multiset counted with rolling extraction, canonical keys, an edge-keyed table and n_threads host
 * threads (hash-sharded); returns ko_digest()'s four numbers for the batch.  bench.py reports it
 * beside the faithful single-thread port, tests check it against ko_digest(). */
int ko_mt_build_digest(int k, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                       int reverse_complement, int n_threads, uint64_t out[4],
                       uint64_t *accepted_reads, uint64_t *accepted_bytes);
/* The same counter as an object, so that the BASELINE-size GPU builds can be checked in full: reads
 * are taken batch by batch (memory follows the distinct keys), and the filter
 * (edges.rs:51-58, pruner.rs:109-118) and standardize_edges (standardizer.rs:42-70,123-127) are
 * applied to its tables; ko_mt_digest is ko_digest's four numbers at any point. */
typedef struct ko_mt ko_mt;
ko_mt *ko_mt_new(int k, int reverse_complement, int n_threads);
void ko_mt_free(ko_mt *j);
int ko_mt_add_reads(ko_mt *j, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads);
void ko_mt_counters(const ko_mt *j, uint64_t *accepted_reads, uint64_t *accepted_bytes);
int ko_mt_digest(ko_mt *j, uint64_t out[4]);
int ko_mt_remove_weak_edges(ko_mt *j, uint32_t threshold);
int ko_mt_standardize_edges(ko_mt *j, uint64_t genome_len, uint64_t k, uint32_t threshold);
/* ko_synth_reads over n_threads host threads (same bytes) */
void ko_synth_reads_mt(uint64_t seed_g, uint64_t G, uint32_t L, uint32_t err_ppm, uint64_t r0, uint64_t r1,
                       uint8_t *out, int n_threads);

/* ---- codec (compress.rs) ---- */
uint8_t ko_encode_fasta_symbol(uint8_t symbol, uint8_t carrier);                /* :347-378 */
size_t ko_compress_node(const uint8_t *s, size_t len, uint8_t *out);            /* :55-73  */
size_t ko_compress_kmer(const uint8_t *kmer, size_t k, uint8_t *out);           /* :18-28  */
size_t ko_compress_kmer_with_rev_compl(const uint8_t *kmer, size_t k, uint8_t *out,
                                       uint8_t *rev);                           /* :34-48  */
void ko_reverse_compressed_node(const uint8_t *in, size_t n, size_t remainder,
                                uint8_t *out);                                  /* :153-169 */
void ko_shift_right_bit_array(uint8_t *a, size_t n, size_t shift);              /* :426-442 */
size_t ko_compress_edge(const uint8_t *edge, size_t len, uint8_t *out);         /* :250-271 */
size_t ko_decompress_edge(const uint8_t *edge, size_t n, uint8_t *out);         /* :283-293 */

/* ---- synthetic reads (our generator; counter based) ---- */
uint64_t ko_splitmix64(uint64_t x);
/* reads [r0, r1) of the configuration; out must hold (r1-r0)*L bytes */
void ko_synth_reads(uint64_t seed_g, uint64_t G, uint32_t L, uint32_t err_ppm, uint64_t r0,
                    uint64_t r1, uint8_t *out);
void ko_synth_genome(uint64_t seed_g, uint64_t pos0, uint64_t n, uint8_t *out);

#ifdef __cplusplus
}
#endif
#endif
