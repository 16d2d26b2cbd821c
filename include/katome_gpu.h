/*
 * katome_gpu.h -- C ABI of the B200 (sm_100a) De Bruijn graph build stage.
 *
 * This is the drop-in boundary for katome's GIR producer: every entry point
 * names the reference interface it replaces (file:line under
 * /root/reference/src/katome/).  A Rust `GpuGIR` that forwards its
 * Init/Build/Clean/Standardizable/Stats impls to these symbols is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain pointers and sizes only; the handle owns all device memory
 *   - every function returns 0 (KTG_OK) or a KTG_ERR_* code; the message of
 *     the last failure on the calling thread is ktg_last_error()
 *   - one host thread drives one handle (the reference is single-threaded:
 *     prelude.rs:32-34); the library uses CUDA streams internally
 *   - there is no CPU fallback: without a CUDA device ktg_create fails
 *   - k-mers are integers key(w) = sum code(w[j]) * 4^(k-1-j), A0 C1 G2 T3
 *     (compress.rs:347-378); integer order == the reference's packed-byte
 *     order.  k <= 32 uses one u64 (lo); 33 <= k <= 64 uses (hi, lo).
 */
#ifndef KATOME_GPU_H
#define KATOME_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KTG_ABI_VERSION 2

enum {
    KTG_OK = 0,
    KTG_ERR_SHORT_READ = 1, /* "Read is too short!"          hm_gir.rs:40      */
    KTG_ERR_BAD_K = 2,      /* assert!(k_size > 1) / len > 2 prelude.rs:35, compress.rs:19 */
    KTG_ERR_IO = 3,         /* check_files / open panics     builder.rs:57-77,148 */
    KTG_ERR_BAD_RECORD = 4, /* reader.records() unwrap()     builder.rs:153     */
    KTG_ERR_DEGENERATE = 5, /* G < k or s == l               standardizer.rs:123-127 */
    KTG_ERR_TABLE_FULL = 6, /* device table cannot grow further (never silent loss) */
    KTG_ERR_CUDA = 7,
    KTG_ERR_INVALID = 8,    /* bad argument / wrong state */
    KTG_ERR_NO_DEVICE = 9
};

/* InputFileType (config.rs:5-14); BFCounter input is a "next" row */
enum { KTG_FASTQ = 0, KTG_FASTA = 1 };

typedef struct ktg_builder ktg_builder;

/* Replaces set_global_k_sizes (prelude.rs:34-43) + Init::init's optional
 * size hints (builder.rs:19-25).  k is per handle, not a process global. */
typedef struct ktg_config {
    uint32_t abi_version;        /* KTG_ABI_VERSION */
    uint32_t k;                  /* K_SIZE: edge length, 3..=64 */
    uint32_t reverse_complement; /* Config::reverse_complement (config.rs:33-36) */
    int32_t device;              /* CUDA device ordinal, -1 = current */
    uint64_t capacity_hint_edges; /* Init::init(edges_count, ..): expected distinct
                                     edges as the reference counts them; 0 = grow on demand */
    uint32_t world_size;         /* hash-sharding: this handle owns keys with */
    uint32_t rank;               /*   owner(key) == rank; 1/0 = single GPU     */
    void *stream;                /* cudaStream_t to launch on; NULL = own stream (pass
                                    cudaStreamLegacy for the legacy default stream) */
    uint32_t sub_table_log2_bytes; /* 0 = default (L2-resident partition size) */
    uint32_t flags;              /* KTG_FLAG_* */
    /* ONE handle over several GPUs of this process: n_devices > 1 builds a table that is
     * hash-sharded over device_ids[0..n_devices) -- the reads of every ktg_add_reads call are split over
     * the devices, each extracts its share and writes the k-mers (or super-k-mer records) straight into
     * the owning device's HBM over NVLink peer memory, and every query / export below answers for the
     * whole graph (the export gathers the shards on device_ids[0] and numbers nodes globally).  The
     * caller changes nothing else: Build::create (builder.rs:42-54) stays one call.  world_size / rank /
     * device / stream are ignored then.  A device may be listed more than once (shards then share a
     * GPU; the tests run 2-4 shards on the one GPU of a test box this way).  n_devices <= 1: one GPU. */
    uint32_t n_devices;
    const int32_t *device_ids;
} ktg_config;

#define KTG_FLAG_PROFILE 1u      /* time every kernel launch with CUDA events */
#define KTG_FLAG_FORCE_DIRECT 2u /* never use the partitioned insert */
#define KTG_FLAG_FORCE_PARTITION 4u
#define KTG_FLAG_FORCE_PAGES 8u  /* always partition twice + streaming page update */
#define KTG_FLAG_NO_PAGES 16u    /* never use the page update (L2 atomics only) */

int ktg_create(const ktg_config *cfg, ktg_builder **out);
void ktg_destroy(ktg_builder *b);
const char *ktg_last_error(void);
int ktg_device_count(void);

/* Build::add_read_fastaq over a batch + the accept/reject rule and byte total
 * of create_fastq's loop body (builder.rs:152-160, hm_gir.rs:39-87).
 * bases: concatenated ASCII reads on the HOST, read r = [offsets[r], offsets[r+1]).
 * A read with any byte outside "ACGT" is dropped whole; an accepted read
 * shorter than k returns KTG_ERR_SHORT_READ and voids the build.
 * accepted_* are incremented (not overwritten) when non-NULL.
 * BUFFER LIFETIME: with page-locked buffers (ktg_host_alloc) and both accepted_* NULL the call returns while the
 * copies of `bases` / `offsets` to the device may still be in flight, so that the caller can prepare its next
 * batch meanwhile.  The buffers must stay valid and untouched until ktg_wait_input, ktg_finalize or any query
 * (ktg_counts, ktg_digest, ...) has returned; with accepted_* non-NULL, or pageable memory, they are free on
 * return.  (Multi-device handles always return with the copies done.) */
int ktg_add_reads(ktg_builder *b, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                  uint64_t *accepted_reads, uint64_t *accepted_bytes);

/* Same, inputs already resident in device memory (d_offsets: n_reads+1 u64); single-device handles only.
 * total_bases == offsets[n_reads] - offsets[0].  The packer reads whole aligned 32-byte
 * chunks: d_bases must be readable from the 32-byte boundary at or below its first base to
 * the one at or above its end (true for any sub-range of a cudaMalloc allocation). */
int ktg_add_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets,
                         uint64_t n_reads, uint64_t total_bases, uint64_t *accepted_reads,
                         uint64_t *accepted_bytes);

/* Build::create (builder.rs:42-54): check_files + FASTQ/FASTA reader + batcher.
 * total_bytes = second element of the reference's return tuple. */
int ktg_create_from_files(ktg_builder *b, const char *const *paths, uint32_t n_paths,
                          int file_type, uint64_t *total_bytes);

/* Waits until the device has its own copy of everything passed to ktg_add_reads so far (see BUFFER LIFETIME). */
int ktg_wait_input(ktg_builder *b);

/* Default::default() again (builder.rs:145): empties the collection but keeps the
 * device allocations, so that repeated builds do not pay cudaMalloc. */
int ktg_reset(ktg_builder *b);

/* Waits for all queued work and surfaces deferred errors (short read, table). */
int ktg_finalize(ktg_builder *b);

/* Stats<CollectionStats> for HmGIR/HsGIR (stats/collections.rs:170-208):
 * node_count = |{prefix, suffix of edges}|, edge_count = |edges|. */
int ktg_counts(ktg_builder *b, uint64_t *nodes, uint64_t *edges);

/* CollectionStats of the converted PtGraph (stats/collections.rs:137-168). */
typedef struct ktg_stats {
    uint64_t node_count, edge_count;
    uint64_t max_edge_weight, sum_edge_weight;
    uint64_t max_in_degree, max_out_degree;
    uint64_t incoming_vert_count; /* nodes with in-degree 0  */
    uint64_t outgoing_vert_count; /* nodes with out-degree 0 */
} ktg_stats;
int ktg_collection_stats(ktg_builder *b, ktg_stats *out);

/* Clean::remove_weak_edges / remove_single_vertices for HmGIR
 * (pruner.rs:95-119, 127-157; edges.rs:51-58): keep edges with w >= threshold;
 * nodes are implicit (prefix/suffix of surviving edges), so the orphan-node
 * rule holds by construction and remove_single_vertices is a no-op. */
int ktg_remove_weak_edges(ktg_builder *b, uint32_t threshold);
int ktg_remove_single_vertices(ktg_builder *b);

/* Standardizable::standardize_edges (standardizer.rs:42-70, 123-127). */
int ktg_standardize_edges(ktg_builder *b, uint64_t genome_len, uint64_t k, uint32_t threshold);

/* Edge export = what Convert::create_from consumes (hm_gir.rs:156-226): the
 * both-strand-expanded edge set with weights, compacted on the device
 * (stream compaction) and, if sorted != 0, ordered by k-mer.
 * key_hi may be NULL when k <= 32.  Returns the edge count in *n; copies at
 * most cap entries. */
int ktg_export_edges(ktg_builder *b, uint64_t *key_hi, uint64_t *key_lo, uint32_t *weight,
                     uint64_t cap, int sorted, uint64_t *n);

/* The graph Convert::create_from builds (hm_gir.rs:156-226, hs_gir.rs:205-262), in a canonical
 * numbering, assembled on the device so that the host only bulk-loads it:
 *   nodes       the n_nodes distinct (k-1)-mers, sorted ascending (node_hi may be NULL for k <= 33);
 *               the reference numbers nodes in HashMap iteration order (hm_gir.rs:160-171), which
 *               is outside the parity contract
 *   edges       sorted by k-mer; edge e goes from node src[e] (its prefix) to node dst[e] (its
 *               suffix) with weight[e] -- the (usize, usize, (EdgeSlice, u32)) of graph.add_edge
 *   edge_bytes  every edge in compress_edge format (compress.rs:250-271), ktg_edge_record_bytes()
 *               = 1 + ceil(k/4) bytes each: what SEQUENCES holds after kmer_to_edge
 * n_nodes / n_edges must be the values of ktg_counts; any output pointer may be NULL. */
/* Builds that graph on the device and reports its sizes, so that the caller can allocate: the following
 * ktg_export_graph / ktg_export_externals only copy (the device copy is dropped when the table changes).
 * Optional -- both compute it on demand -- but it saves the separate node count of ktg_counts.  Buffers
 * from ktg_host_alloc (pinned) take the copies at PCIe speed; pageable memory is several times slower. */
int ktg_graph_prepare(ktg_builder *b, uint64_t *n_nodes, uint64_t *n_edges);
int ktg_export_graph(ktg_builder *b, uint64_t *node_hi, uint64_t *node_lo, uint64_t n_nodes, uint64_t *src,
                     uint64_t *dst, uint32_t *weight, uint8_t *edge_bytes, uint64_t n_edges);
uint32_t ktg_edge_record_bytes(const ktg_builder *b);

/* The seeds of remove_dead_paths (pruner.rs:165-195, the `Externals` iterator): in ascending
 * node index -- the numbering of ktg_export_graph -- every node without an incoming edge as
 * VertexType::Input (kinds[i] = KTG_EXTERNAL_INPUT), else every node without an outgoing edge as
 * VertexType::Output (KTG_EXTERNAL_OUTPUT).  The walk along the dead paths is pointer chasing and stays on
 * the host; this is its degree-0 seeding.  Returns the count in *n, copies at most cap entries
 * (node_ids / kinds may be NULL for a size query). */
enum { KTG_EXTERNAL_INPUT = 0, KTG_EXTERNAL_OUTPUT = 1 };
int ktg_export_externals(ktg_builder *b, uint64_t *node_ids, uint8_t *kinds, uint64_t cap, uint64_t *n);

/* ---- BFCounter input (builder.rs:79-115, pt_graph.rs:201-213,318-329).  The
 * reference implements it for PtGraph only (add_read_bfc is unreachable!() on its GIR types,
 * builder.rs:32-36); here the same table takes it: for every k-mer (exactly k ASCII bases, n*k
 * contiguous bytes) with count >= minimal_weight_threshold (builder.rs:106-108),
 * weight[kmer] += count, and weight[revcomp(kmer)] += count with reverse_complement.
 * *accepted_bytes += k per accepted k-mer (builder.rs:109).  BFCounter's k-mers are unique
 * (pt_graph.rs:78); a repeated one adds up here where petgraph would hold parallel edges.
 * DEVIATION: weight 0 means "not an edge" in this table.  A line with count 0 (only reachable with
 * minimal_weight_threshold == 0; BFCounter never writes one) is accepted and counted in accepted_* like the
 * reference does, but adds no edge, where the reference would add an edge of weight 0; likewise an edge whose
 * u32 weight wraps to exactly 0 (2^32 occurrences) disappears.
 * Single GPU.  ktg_create_from_bfc_files reads "<k-mer>\t<count>" lines and finalizes. */
int ktg_add_weighted_kmers(ktg_builder *b, const uint8_t *kmers, const uint32_t *weights, uint64_t n,
                           uint32_t minimal_weight_threshold, uint64_t *accepted_kmers, uint64_t *accepted_bytes);
int ktg_create_from_bfc_files(ktg_builder *b, const char *const *paths, uint32_t n_paths,
                              uint32_t minimal_weight_threshold, uint64_t *total_bytes);

/* Order-independent digest over the expanded edge set:
 * out[0] = sum splitmix64(splitmix64(hi)^lo)*(2w+1), out[1] = |E|,
 * out[2] = sum w, out[3] = max w. */
int ktg_digest(ktg_builder *b, uint64_t out[4]);

/* ---- hash-sharding across GPUs (world_size > 1): the data path is
 * extract+partition on the sender, an all-to-all of keys (done by the host
 * with NCCL), and insert on the owner. ---- */
uint32_t ktg_key_words(const ktg_builder *b); /* 1 (k<=32) or 2 */
/* owner rank of a k-mer (after canonicalisation when reverse_complement) */
uint32_t ktg_owner_of(const ktg_builder *b, uint64_t key_hi, uint64_t key_lo);
/* Pack + validate + extract the batch's canonical keys, grouped by owner rank.
 * On return d_keys (device, owned by the handle, valid until the next call)
 * holds the keys owner-major; counts[world_size] are keys per owner. */
int ktg_partition_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets,
                               uint64_t n_reads, uint64_t total_bases, void **d_keys,
                               uint64_t *counts, uint64_t *accepted_reads,
                               uint64_t *accepted_bytes);
/* Insert n canonical keys (device pointer, ktg_key_words() u64 each, lo first). */
int ktg_insert_keys_device(ktg_builder *b, const void *d_keys, uint64_t n);

/* Whole-graph node statistics and standardize_edges of a SHARDED table.  A node's edges may live on
 * several shards: every shard exports the (canonical (k-1)-mer, degree word) pairs of its own edges
 * (device arrays owned by the handle; keys are key_words u64 each, lo first), the host routes
 * them to an owner of its choice (any function of the key) and the owner merges what it received;
 * node_count / degrees / sources / sinks of the shards then add up (max for the maxima).
 * standardize_edges: all-reduce the two sums, compute p = (G-k)/(s-l) once, scale every shard. */
int ktg_nodes_export_device(ktg_builder *b, void **d_keys, void **d_degrees, uint64_t *n, uint32_t *key_words);
int ktg_nodes_stats_from_device(ktg_builder *b, const void *d_keys, const void *d_degrees, uint64_t n, ktg_stats *out);
int ktg_edge_sums(ktg_builder *b, uint32_t threshold, uint64_t *sum_w, uint64_t *sum_w_below);
int ktg_scale_weights(ktg_builder *b, double ratio, uint32_t threshold);

/* Same grouping for an array of keys (the spill list of the fused exchange below). */
int ktg_partition_keys_device(ktg_builder *b, const void *d_keys, uint64_t n, void **d_out, uint64_t *counts);

/* ---- fused exchange (one node, world_size <= 8): the extraction kernel of every rank
 * writes its keys, grouped by owner, straight into the owners' HBM over NVLink (mapped peer
 * memory); there is no separate all-to-all of keys.  Per batch, on every rank and in this order:
 *   1. ktg_mg_plan / ktg_mg_prepare size the receive buffer (one bucket per source rank) for
 *      the largest per-rank batch (all ranks pass the same max_windows); when it is replaced,
 *      peers unmap the old one first and the new ktg_ipc_get_handle is exchanged
 *      (ktg_ipc_open on the peers);
 *   2. a collective orders "everybody finished reading its receive buffer" before
 *   3. ktg_mg_scatter_reads_device(peer_rx[world]) - pack + extract + scatter into the
 *      peers' buckets; *d_cursors (device, world u64) are the bucket ends it reached
 *      (owner * bucket_cap + fill);
 *   4. the sketch (ktg_mg_sketch) is all-reduced with MAX so that every shard can size itself;
 *   5. an all-to-all of the cursors tells every owner how full its buckets are (and that the
 *      writers are done); ktg_mg_insert_buckets partitions them by sub-table into the staged
 *      buckets, from where they are flushed like any other batch;
 *   6. keys that did not fit a bucket (ktg_mg_spill) are grouped with
 *      ktg_partition_keys_device, exchanged with NCCL and added with ktg_mg_insert_spill.
 * A large batch is sent in chunks: the receive buffer has two slots (slot s starts rx_bytes * s
 * after rx_base), chunk c uses slot c % 2, and steps 3 (on send_stream) and 5 (on the handle's
 * stream) of consecutive chunks overlap.  first_of_batch resets the spill list. */
int ktg_mg_plan(ktg_builder *b, uint64_t max_windows, int *needs_realloc);
int ktg_mg_prepare(ktg_builder *b, uint64_t max_windows, void **rx_base, uint64_t *rx_bytes,
                   uint64_t *bucket_cap, uint32_t *n_sub);
int ktg_mg_scatter_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets, uint64_t n_reads,
                                uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                                void *send_stream, void **d_cursors);
int ktg_mg_insert_buckets(ktg_builder *b, const void *d_bucket_ends, uint64_t n_keys, uint32_t slot);
int ktg_mg_sketch(ktg_builder *b, void **d_regs, uint32_t *n_regs);
/* fold an all-reduced COPY of the sketch back in (atomic max; the live sketch may be written
 * by the next chunk's scatter at the same time, so it is never all-reduced in place) */
int ktg_mg_merge_sketch(ktg_builder *b, const void *d_regs);
int ktg_mg_spill(ktg_builder *b, void **d_keys, uint64_t *n);
int ktg_mg_insert_spill(ktg_builder *b, const void *d_keys, uint64_t n);
/* ---- the DIRECT exchange: the sender does the owner's level-1 partition as well.  Its extraction kernel bins
 * every key by (owner, sub-table of the owner) -- world x n_sub bins, as many as the one-GPU build of the same
 * table has -- and writes the runs into the owner's receive bucket for (this sender, that sub-table); the owner
 * goes straight to its page level (ktg_mg_direct_insert: level-2 scatter + page sweep), without the extra pass
 * over the received keys that ktg_mg_insert_buckets needs.  Every shard must have the same geometry:
 *   ktg_mg_direct_plan      n_sub / sub_log2 of this shard (all-reduce them: use this exchange only when all
 *                           ranks agree) and whether ktg_mg_direct_prepare would replace the receive buffer;
 *   ktg_mg_direct_prepare   bucket_cap keys per (source, sub-table) bucket, world * n_sub buckets per slot;
 *   ktg_mg_direct_scatter_reads_device   *d_cursors: world * n_sub u64 bucket ends, owner-major
 *                           ((owner * n_sub + sub) * bucket_cap + fill); send n_sub of them to every owner.
 *                           A batch may be scattered in several calls (chunks of a host batch, copied while
 *                           the previous chunk is scattered): only the first passes first_of_batch, the
 *                           others append to the same buckets, and the owners insert once at the end;
 *   the sketch is all-reduced as above; ktg_mg_direct_insert(ends[source * n_sub + sub] = (source * n_sub +
 *   sub) * bucket_cap + fill, exact key count); spills: ktg_mg_spill and on as above. */
int ktg_mg_direct_plan(ktg_builder *b, uint64_t max_windows, int *needs_realloc, uint32_t *n_sub, uint32_t *sub_log2);
int ktg_mg_direct_prepare(ktg_builder *b, uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap);
int ktg_mg_direct_scatter_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets, uint64_t n_reads,
                                       uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                                       void *send_stream, void **d_cursors);
int ktg_mg_direct_insert(ktg_builder *b, const void *d_bucket_ends, uint64_t n_keys, uint32_t slot);
/* ---- the fused exchange in SUPER-K-MER records (23 <= k <= 31).  The owner of a k-mer is a
 * function of its minimizer (smallest hashed canonical (k-15)-mer inside it); consecutive windows
 * of a read mostly share it, so a run of n <= 16 windows travels over NVLink as ONE 16-byte
 * record holding its n+k-1 bases (about 2.8 bytes per window instead of 8), and the owner unrolls
 * records into canonical k-mers while it partitions them by sub-table.  Same protocol as above:
 *   ktg_mg_skm_plan / ktg_mg_skm_prepare   bucket_cap is in records (16 bytes each);
 *   ktg_mg_skm_scatter_reads_device        *d_cursors: world u64 bucket ends (owner * bucket_cap
 *       + fill, in records); *d_key_counts: world u64, the k-mers sent to every owner;
 *   all-to-all of both; no sketch exchange (the owner sketches the keys it unrolls);
 *   ktg_mg_skm_insert_buckets(ends, sum of the received key counts);
 *   records that did not fit a bucket: ktg_mg_skm_spill -> ktg_mg_skm_partition_records ->
 *       all-to-all (2 u64 per record, lo first) -> ktg_mg_skm_insert_records.
 * The two exchanges assign k-mers to shards differently; use one per build. */
int ktg_mg_skm_supported(uint32_t k);
int ktg_mg_skm_plan(ktg_builder *b, uint64_t max_windows, int *needs_realloc);
int ktg_mg_skm_prepare(ktg_builder *b, uint64_t max_windows, void **rx_base, uint64_t *rx_bytes, uint64_t *bucket_cap);
int ktg_mg_skm_scatter_reads_device(ktg_builder *b, const void *d_bases, const void *d_offsets, uint64_t n_reads,
                                    uint64_t total_bases, void *const *peer_rx, uint32_t slot, int first_of_batch,
                                    void *send_stream, void **d_cursors, void **d_key_counts);
int ktg_mg_skm_insert_buckets(ktg_builder *b, const void *d_bucket_ends, uint64_t n_keys_ub, uint32_t slot);
int ktg_mg_skm_spill(ktg_builder *b, void **d_records, uint64_t *n);
int ktg_mg_skm_partition_records(ktg_builder *b, const void *d_records, uint64_t n, void **d_out, uint64_t *counts);
int ktg_mg_skm_insert_records(ktg_builder *b, const void *d_records, uint64_t n);
/* owner rank of a k-mer in this exchange (after canonicalisation it is the same for both strands) */
uint32_t ktg_mg_skm_owner_of(const ktg_builder *b, uint64_t key_hi, uint64_t key_lo);
/* Host-only twins of the record cutter (no device needed; used by the CPU tests): work item i is
 * the 16 window starts from flat base position pos[i] of the 2-bit stream `packed` (first base of
 * word j in bits 63:62), valid[i] bit j = window j exists.  Records come back as (lo, hi) pairs. */
int ktg_skm_items_host(const uint64_t *packed, const uint64_t *pos, const uint32_t *valid, uint64_t n_items, uint32_t k,
                       uint32_t world, uint64_t *records_lo_hi, uint64_t cap, uint64_t *n_records);
uint32_t ktg_skm_owner_of_kmer(uint64_t kmer, uint32_t k, uint32_t world);
/* CUDA IPC handles of device allocations (cudaIpcGetMemHandle / OpenMemHandle / CloseMemHandle) */
int ktg_ipc_get_handle(const void *dev_ptr, uint8_t handle[64]);
int ktg_ipc_open(const uint8_t handle[64], void **dev_ptr);
int ktg_ipc_close(void *dev_ptr);

/* ---- pinned host memory for the batcher ---- */
int ktg_host_alloc(void **p, size_t bytes);
void ktg_host_free(void *p);

/* ---- synthetic reads on the device (bench workloads; twin of
 * oracle ko_synth_reads).  Writes (r1-r0)*L ASCII bytes to d_out. */
int ktg_synth_reads_device(void *d_out, uint64_t seed_g, uint64_t genome_len, uint32_t read_len,
                           uint32_t err_ppm, uint64_t r0, uint64_t r1, void *stream);

/* ---- random-access roofline probe: n uniformly random 'load key + atomicAdd'
 * updates over a table of `bytes` bytes; returns elapsed ms. ---- */
int ktg_random_access_probe(uint64_t bytes, uint64_t n_updates, uint32_t slot_bytes, float *ms);

/* ---- profiling (KTG_FLAG_PROFILE) ---- */
typedef struct ktg_kernel_profile {
    char name[32];
    uint64_t launches;
    double total_ms;
    uint64_t units; /* units of work (bases, windows, slots) summed over launches */
} ktg_kernel_profile;
int ktg_get_profile(ktg_builder *b, ktg_kernel_profile *out, uint32_t cap, uint32_t *n);
int ktg_reset_profile(ktg_builder *b);
/* Switch the per-launch timing on or off for the launches that follow (the events around every
 * kernel cost ~20 us of GPU time per launch: 0.8 ms of an 11 ms host-fed C2 build). */
int ktg_set_profile(ktg_builder *b, int enabled);

/* ---- test hook (no GPU needed): the plan ktg_add_reads follows for a batch with these offsets and
 * this chunk size -- cuts[0..n_chunks] (chunk c = reads [cuts[c], cuts[c+1])), flush_after[c] != 0 where
 * the stage is flushed on the way (after the given percentages of the bytes), *tail_first = index of the
 * first of the short chunks a large call ends in, or -1.  cuts needs cap + 1 entries, flush_after cap.
 * The per-read loop this batches is create_fastq, algorithms/builder.rs:152-160. */
int ktg_plan_chunks(const uint64_t *offsets, uint64_t n_reads, uint64_t chunk_bytes, const uint32_t *flush_pcts,
                    uint32_t n_pcts, uint64_t *cuts, uint8_t *flush_after, uint32_t cap, uint32_t *n_chunks,
                    int64_t *tail_first);

/* ---- test hook (no GPU needed): item / items_per_read the way the extraction kernels compute it for a batch
 * of equally long reads (a multiplication by floor(2^64 / d) + 1 instead of a division): out[i] for items[i]. */
int ktg_item_reads(const uint32_t *items, uint64_t n, uint32_t items_per_read, uint32_t *out);

/* ---- test hook (no GPU needed): the host FASTQ / FASTA reader behind ktg_create_from_files alone
 * (csrc/host_reader.h: check_files, builder.rs:57-77, and the record semantics of rust-bio 0.10 that
 * create_fastq / create_fasta rely on, builder.rs:118-165): records read, bases in their sequences
 * (before the ACGT filter, which runs on the device) and an FNV-1a checksum over every sequence
 * followed by '\n'.  batch_bytes: bases per internal batch (0 = 64 MiB).  KTG_ERR_IO / KTG_ERR_BAD_RECORD
 * as ktg_create_from_files returns them. */
int ktg_host_parse_file(const char *path, int file_type, uint64_t batch_bytes, uint64_t *n_records,
                        uint64_t *total_bases, uint64_t *checksum);

/* ---- options (tests and measurements; the defaults are the measured winners and nothing in the
 * library reads the environment).  Names: page_threads, page_nbuf, page_log2, l2s_variant, l1_ctas,
 * l1_big (level-1 scatter with the big tile: -1 when there are more than 256 bins, 0 never, 1 always), p2p_ctas, chunk_mb, stage_bufs, flush_pct, flush_pct2, taper, eager_pages, stage_factor_milli,
 * stage_max_keys, host_parse (ktg_create_from_files: records cut by the host reader instead of the
 * device), fastq_chunk_kb, mg_pad, mg_direct (a handle over several devices: the direct exchange; -1 below 4
 * devices, 0 never, 1 whenever the shards' geometries agree), trace (host timeline on stderr).  Unknown names:
 * KTG_ERR_INVALID. */
int ktg_set_option(ktg_builder *b, const char *name, int64_t value);

typedef struct ktg_info {
    uint64_t capacity_slots, occupied_slots;
    uint64_t table_bytes;
    uint32_t n_sub_tables, slot_bytes;
    uint64_t windows_inserted;
    uint64_t kernel_launches;
    uint32_t grow_events, partitioned;
    uint32_t page_updates, n_pages; /* batches inserted by the streaming page update */
} ktg_info;
int ktg_get_info(ktg_builder *b, ktg_info *out);

#ifdef __cplusplus
}
#endif
#endif
