/*
 * katome_oracle.c -- CPU restatement ("port") of katome's GIR build stage.
 *
 * TEST INFRASTRUCTURE ONLY (see katome_oracle.h).  Single-threaded on purpose:
 * the reference is single-threaded by construction (prelude.rs:32-34,
 * asm/mod.rs:21).  It keeps the reference's *work shape* -- a node-keyed hash
 * map with per-node outgoing lists, per-window re-packing of both (k-1)-mers
 * from ASCII, byte-level reverse complement, source->target chaining -- so that
 * it doubles as the CPU baseline in bench.py.
 *
 * Citations are file:line under /root/reference/src/katome/.
 */
#define _GNU_SOURCE
#include "katome_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define KO_MAXC 16 /* ceil(63/4): largest packed (k-1)-mer we support (k <= 64) */
#define KO_NONE 0xFFFFFFFFu

typedef unsigned __int128 u128;

/* Edge = (Idx target, EdgeWeight weight, u8 last_char)  -- girs/edges.rs:10 */
typedef struct {
    uint32_t target;
    uint32_t weight;
    uint8_t last_char;
} ko_edge;

/* One hash-map entry: key = packed (k-1)-mer (what NodeSlice dereferences to,
 * slices.rs:86-131,138-145), value = Outgoing (girs/edges.rs:12). */
typedef struct {
    uint8_t key[KO_MAXC];
    ko_edge out[4];
    uint8_t nout;
    uint8_t dead; /* removed by remove_single_vertices */
} ko_node;

struct ko_gir {
    int k;       /* K_SIZE              prelude.rs:21 */
    int k1;      /* K1_SIZE             prelude.rs:23 */
    int c;       /* COMPRESSED_K1_SIZE  prelude.rs:25 */
    int padbits; /* zero bits at the bottom of the last carrier */
    ko_node *nodes;
    uint64_t n_nodes, cap_nodes, live_nodes;
    uint32_t *slots; /* open addressing over node ids */
    uint64_t n_slots;
    /* scratch for the stashed reverse complements, hm_gir.rs:53 */
    uint8_t *rev;
    size_t rev_cap;
};

/* ------------------------------------------------------------------ codec */

/* compress.rs:347-378.  A=00 C=01 G=10 T=11, appended as the two LSBs. */
uint8_t ko_encode_fasta_symbol(uint8_t symbol, uint8_t carrier) {
    carrier = (uint8_t)(carrier << 2);
    symbol = (uint8_t)(symbol - 'A');
    symbol >>= 1;
    uint8_t c = (symbol & 2) >> 1, a = (symbol & 8) >> 3, d = symbol & 1;
    uint8_t first = (uint8_t)((c ^ 1) & d), second = (uint8_t)(c | a);
    return (uint8_t)(carrier | (uint8_t)((second << 1) | first));
}

/* compress.rs:55-73: 4 symbols per carrier, MSB first, last carrier left-aligned. */
size_t ko_compress_node(const uint8_t *s, size_t len, uint8_t *out) {
    size_t n = 0;
    for (size_t i = 0; i < len; i += 4) {
        size_t m = len - i < 4 ? len - i : 4;
        uint8_t carrier = 0;
        for (size_t j = 0; j < m; ++j) carrier = ko_encode_fasta_symbol(s[i + j], carrier);
        if (m < 4) carrier = (uint8_t)(carrier << (2 * (4 - m)));
        out[n++] = carrier;
    }
    return n;
}

/* compress.rs:18-28: [pack(kmer[..k-1]) | pack(kmer[1..])] */
size_t ko_compress_kmer(const uint8_t *kmer, size_t k, uint8_t *out) {
    size_t c = ko_compress_node(kmer, k - 1, out);
    ko_compress_node(kmer + 1, k - 1, out + c);
    return 2 * c;
}

/* compress.rs:426-442 */
void ko_shift_right_bit_array(uint8_t *a, size_t n, size_t shift) {
    shift %= 8;
    if (shift == 0) return;
    uint8_t carry = 0;
    uint8_t mask = (uint8_t)((1u << shift) - 1u);
    for (size_t i = 0; i < n; ++i) {
        uint8_t low = a[i] & mask;
        a[i] = (uint8_t)((a[i] >> shift) | (uint8_t)(carry << (8 - shift)));
        carry = low;
    }
}

/* compress.rs:121-131: reverse the four 2-bit symbols inside a byte */
static inline uint8_t rev_symbols_u8(uint8_t x) {
    x = (uint8_t)(((x >> 2) & 0x33) | ((x & 0x33) << 2));
    x = (uint8_t)(((x >> 4) & 0x0F) | ((x & 0x0F) << 4));
    return x;
}

/* compress.rs:153-169: right-align, reverse bytes, reverse symbols in each
 * byte, complement (NOT), clear the padding again. */
void ko_reverse_compressed_node(const uint8_t *in, size_t n, size_t remainder, uint8_t *out) {
    size_t padding = ((4 - remainder) % 4) * 2;
    uint8_t tmp[KO_MAXC + 1];
    memcpy(tmp, in, n);
    ko_shift_right_bit_array(tmp, n, padding);
    for (size_t i = 0; i < n; ++i) out[i] = (uint8_t)~rev_symbols_u8(tmp[n - 1 - i]);
    out[n - 1] &= (uint8_t) ~((1u << padding) - 1u);
}

/* compress.rs:34-48: rc k-mer = [rc(end node) | rc(start node)] */
size_t ko_compress_kmer_with_rev_compl(const uint8_t *kmer, size_t k, uint8_t *out,
                                       uint8_t *rev) {
    size_t c = ko_compress_node(kmer, k - 1, out);
    ko_compress_node(kmer + 1, k - 1, out + c);
    size_t remainder = (k - 1) % 4;
    ko_reverse_compressed_node(out + c, c, remainder, rev);
    ko_reverse_compressed_node(out, c, remainder, rev + c);
    return 2 * c;
}

/* compress.rs:250-271: [padding byte | packed symbols] */
size_t ko_compress_edge(const uint8_t *edge, size_t len, uint8_t *out) {
    size_t n = ko_compress_node(edge, len, out + 1);
    out[0] = (uint8_t)((4 - (len % 4)) % 4);
    /* the reference writes CHARS_PER_CARRIER - last_chunk_len, which is 0 for a
     * full last chunk (4-4) -- same value */
    return n + 1;
}

/* compress.rs:283-293 */
size_t ko_decompress_edge(const uint8_t *edge, size_t n, uint8_t *out) {
    static const char sym[4] = {'A', 'C', 'G', 'T'};
    size_t padding = edge[0], len = (n - 1) * 4 - padding, o = 0;
    for (size_t i = 1; i < n; ++i)
        for (int j = 3; j >= 0 && o < len; --j) out[o++] = (uint8_t)sym[(edge[i] >> (2 * j)) & 3];
    return len;
}

/* ------------------------------------------------------------- hash table */

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* Hash of the packed bytes.  The reference uses MetroHash (hm_gir.rs:22); the
 * hash only decides iteration order, never results. */
static inline uint64_t hash_key(const uint8_t *key, int c) {
    uint64_t h = 0xcbf29ce484222325ull;
    for (int i = 0; i < c; ++i) h = (h ^ key[i]) * 0x100000001b3ull;
    return mix64(h);
}

static void rehash(ko_gir *g, uint64_t n_slots) {
    free(g->slots);
    g->n_slots = n_slots;
    g->slots = (uint32_t *)malloc(n_slots * sizeof(uint32_t));
    memset(g->slots, 0xFF, n_slots * sizeof(uint32_t));
    for (uint64_t id = 0; id < g->n_nodes; ++id) {
        if (g->nodes[id].dead) continue;
        uint64_t s = hash_key(g->nodes[id].key, g->c) & (n_slots - 1);
        while (g->slots[s] != KO_NONE) s = (s + 1) & (n_slots - 1);
        g->slots[s] = (uint32_t)id;
    }
}

static uint32_t find_node(const ko_gir *g, const uint8_t *key) {
    uint64_t s = hash_key(key, g->c) & (g->n_slots - 1);
    for (;;) {
        uint32_t id = g->slots[s];
        if (id == KO_NONE) return KO_NONE;
        if (!g->nodes[id].dead && memcmp(g->nodes[id].key, key, (size_t)g->c) == 0) return id;
        s = (s + 1) & (g->n_slots - 1);
    }
}

/* gir.entry(key): Occupied -> existing id, Vacant -> push + insert(Box::new([]))
 * (hm_gir.rs:99-120, 122-149) */
static uint32_t find_or_insert_node(ko_gir *g, const uint8_t *key) {
    uint32_t id = find_node(g, key);
    if (id != KO_NONE) return id;
    if ((g->n_nodes + 1) * 2 > g->n_slots) rehash(g, g->n_slots * 2);
    if (g->n_nodes == g->cap_nodes) {
        g->cap_nodes *= 2;
        g->nodes = (ko_node *)realloc(g->nodes, g->cap_nodes * sizeof(ko_node));
    }
    id = (uint32_t)g->n_nodes++;
    ko_node *n = &g->nodes[id];
    memset(n, 0, sizeof(*n));
    memcpy(n->key, key, (size_t)g->c);
    g->live_nodes++;
    uint64_t s = hash_key(key, g->c) & (g->n_slots - 1);
    while (g->slots[s] != KO_NONE) s = (s + 1) & (g->n_slots - 1);
    g->slots[s] = id;
    return id;
}

/* hs_gir.rs:192-203 */
static void create_or_modify_edge_w(ko_node *src, uint32_t to, uint8_t last_char, uint32_t inc) {
    for (int i = 0; i < src->nout; ++i) {
        if (src->out[i].target == to) {
            src->out[i].weight += inc; /* u32, wrapping in release (Cargo.toml:56) */
            return;
        }
    }
    ko_edge e = {to, inc, last_char};
    src->out[src->nout++] = e; /* at most 4 distinct successors exist */
}
static void create_or_modify_edge(ko_node *src, uint32_t to, uint8_t last_char) {
    create_or_modify_edge_w(src, to, last_char, 1);
}

/* hm_gir.rs:91-153.  `kmer` = [start node | end node] packed. */
static void add_single_edge(ko_gir *g, int first_node, const uint8_t *kmer, uint32_t *source,
                            uint8_t last_char) {
    if (first_node) *source = find_or_insert_node(g, kmer);
    uint32_t target = find_or_insert_node(g, kmer + g->c);
    create_or_modify_edge(&g->nodes[*source], target, last_char);
    *source = target;
}

/* ---------------------------------------------------------------- builder */

ko_gir *ko_new(int k) {
    if (k < 3 || k > 64) return NULL; /* prelude.rs:35, compress.rs:19; <=64 is our key width */
    ko_gir *g = (ko_gir *)calloc(1, sizeof(ko_gir));
    g->k = k;
    g->k1 = k - 1;
    g->c = (g->k1 + 3) / 4;
    g->padbits = ((4 - g->k1 % 4) % 4) * 2;
    g->cap_nodes = 1024;
    g->nodes = (ko_node *)malloc(g->cap_nodes * sizeof(ko_node));
    g->n_slots = 0;
    g->slots = NULL;
    rehash(g, 4096);
    return g;
}

void ko_free(ko_gir *g) {
    if (!g) return;
    free(g->nodes);
    free(g->slots);
    free(g->rev);
    free(g);
}

int ko_k(const ko_gir *g) { return g->k; }

/* hm_gir.rs:39-87 */
int ko_add_read_fastaq(ko_gir *g, const uint8_t *read, size_t len, int reverse_complement) {
    if (len < (size_t)g->k) return KO_ERR_SHORT_READ; /* :40 */
    size_t n_windows = len - (size_t)g->k + 1;
    uint8_t kmer[2 * KO_MAXC];
    uint32_t source = KO_NONE;
    size_t stride = 2 * (size_t)g->c;
    if (reverse_complement) {
        if (g->rev_cap < n_windows * stride) {
            g->rev_cap = n_windows * stride * 2;
            g->rev = (uint8_t *)realloc(g->rev, g->rev_cap);
        }
        for (size_t i = 0; i < n_windows; ++i) { /* :55-67 */
            const uint8_t *w = read + i;
            ko_compress_kmer_with_rev_compl(w, (size_t)g->k, kmer, g->rev + i * stride);
            add_single_edge(g, i == 0, kmer, &source, w[g->k - 1]);
        }
        /* :69-74 -- last window's reverse complement first, then the rest backwards */
        add_single_edge(g, 1, g->rev + (n_windows - 1) * stride, &source,
                        read[n_windows - 1 + (size_t)g->k - 1]);
        for (size_t i = n_windows - 1; i-- > 0;)
            add_single_edge(g, 0, g->rev + i * stride, &source, read[i + (size_t)g->k - 1]);
    }
    else {
        for (size_t i = 0; i < n_windows; ++i) { /* :77-85 */
            ko_compress_kmer(read + i, (size_t)g->k, kmer);
            add_single_edge(g, i == 0, kmer, &source, read[i + (size_t)g->k - 1]);
        }
    }
    return KO_OK;
}

/* builder.rs:155: seq.iter().all(|&x| "ACGT".bytes().any(|i| i == x)) */
static inline int all_acgt(const uint8_t *s, size_t len) {
    for (size_t i = 0; i < len; ++i) {
        uint8_t x = s[i];
        if (!(x == 'A' || x == 'C' || x == 'G' || x == 'T')) return 0;
    }
    return 1;
}

int ko_add_reads(ko_gir *g, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                 int reverse_complement, uint64_t *accepted_reads, uint64_t *accepted_bytes) {
    uint64_t nr = 0, nb = 0;
    int rc = KO_OK;
    for (uint64_t r = 0; r < n_reads; ++r) {
        const uint8_t *seq = bases + offsets[r];
        size_t len = (size_t)(offsets[r + 1] - offsets[r]);
        if (!all_acgt(seq, len)) continue; /* builder.rs:155-157 */
        nb += len;                         /* builder.rs:158 */
        nr += 1;
        rc = ko_add_read_fastaq(g, seq, len, reverse_complement); /* builder.rs:159 */
        if (rc != KO_OK) break; /* the reference panics: the build is void */
    }
    if (accepted_reads) *accepted_reads += nr;
    if (accepted_bytes) *accepted_bytes += nb;
    return rc;
}

/* ---- file readers: restatement of rust-bio 0.10.0 (Cargo.lock:22-25), which
 * is NOT in /root/reference; behaviour per its published source: strict
 * 4-line FASTQ records, '@' header, seq = line 2 with trailing whitespace
 * trimmed, no seq/qual length check, empty qual line => error; FASTA: '>'
 * header, sequence lines concatenated after trimming trailing whitespace. */
static size_t rtrim(const char *s, size_t n) {
    while (n > 0 && (s[n - 1] == '\n' || s[n - 1] == '\r' || s[n - 1] == ' ' ||
                     s[n - 1] == '\t' || s[n - 1] == '\v' || s[n - 1] == '\f'))
        --n;
    return n;
}

static int feed_read(ko_gir *g, const uint8_t *seq, size_t len, int rc_flag, uint64_t *nr,
                     uint64_t *nb) {
    if (!all_acgt(seq, len)) return KO_OK;
    *nb += len;
    *nr += 1;
    return ko_add_read_fastaq(g, seq, len, rc_flag);
}

static int read_fastq(ko_gir *g, FILE *f, int rc_flag, uint64_t *nr, uint64_t *nb) {
    char *hdr = NULL, *seq = NULL, *sep = NULL, *qual = NULL;
    size_t ch = 0, cs = 0, cp = 0, cq = 0;
    int err = KO_OK;
    for (;;) {
        ssize_t lh = getline(&hdr, &ch, f);
        if (lh <= 0) break; /* empty header line buffer => end of records */
        if (hdr[0] != '@') { err = KO_ERR_BAD_RECORD; break; }
        ssize_t ls = getline(&seq, &cs, f);
        if (ls < 0) ls = 0;
        ssize_t lp = getline(&sep, &cp, f);
        (void)lp;
        ssize_t lq = getline(&qual, &cq, f);
        if (lq <= 0) { err = KO_ERR_BAD_RECORD; break; } /* "Incomplete record" */
        size_t len = rtrim(seq, (size_t)ls);
        err = feed_read(g, (const uint8_t *)seq, len, rc_flag, nr, nb);
        if (err != KO_OK) break;
    }
    free(hdr); free(seq); free(sep); free(qual);
    return err;
}

static int read_fasta(ko_gir *g, FILE *f, int rc_flag, uint64_t *nr, uint64_t *nb) {
    char *line = NULL;
    size_t cl = 0;
    uint8_t *seq = NULL;
    size_t seq_len = 0, seq_cap = 0;
    int err = KO_OK, have = 0;
    ssize_t n = getline(&line, &cl, f);
    while (n > 0) {
        if (line[0] != '>') { err = KO_ERR_BAD_RECORD; break; }
        have = 1;
        seq_len = 0;
        for (;;) {
            n = getline(&line, &cl, f);
            if (n <= 0 || line[0] == '>') break;
            size_t m = rtrim(line, (size_t)n);
            if (seq_len + m > seq_cap) {
                seq_cap = (seq_len + m) * 2 + 64;
                seq = (uint8_t *)realloc(seq, seq_cap);
            }
            memcpy(seq + seq_len, line, m);
            seq_len += m;
        }
        if (have) {
            err = feed_read(g, seq, seq_len, rc_flag, nr, nb);
            if (err != KO_OK) break;
        }
    }
    free(line); free(seq);
    return err;
}

/* Build::create + check_files + create_fastq/create_fasta
 * (builder.rs:42-54, 57-77, 118-140, 142-165) */
int ko_create_from_files(ko_gir *g, const char *const *paths, int n_paths, int file_type,
                         int reverse_complement, uint64_t *accepted_reads,
                         uint64_t *accepted_bytes) {
    uint64_t nr = 0, nb = 0;
    int err = KO_OK;
    FILE **fs = (FILE **)calloc((size_t)n_paths, sizeof(FILE *));
    for (int i = 0; i < n_paths; ++i) { /* all files are opened up front: builder.rs:146-149 */
        fs[i] = fopen(paths[i], "rb");
        if (!fs[i]) { err = KO_ERR_IO; break; }
    }
    for (int i = 0; i < n_paths && err == KO_OK; ++i)
        err = file_type == 1 ? read_fasta(g, fs[i], reverse_complement, &nr, &nb)
                             : read_fastq(g, fs[i], reverse_complement, &nr, &nb);
    for (int i = 0; i < n_paths; ++i)
        if (fs[i]) fclose(fs[i]);
    free(fs);
    if (accepted_reads) *accepted_reads = nr;
    if (accepted_bytes) *accepted_bytes = nb;
    return err;
}

/* ---- BFCounter input.  add_read_bfc (pt_graph.rs:318-329) + add_single_edge_bfc
 * (pt_graph.rs:201-213): one edge of the given weight per line, plus its reverse complement with
 * reverse_complement.  The reference implements it for PtGraph only (GIR types: unreachable!(),
 * builder.rs:32-36) and never merges: BFCounter's k-mers are unique (pt_graph.rs:78).  Restated on
 * the GIR's map, a k-mer that does come twice adds its weight to the one edge (the petgraph
 * multigraph would hold parallel edges; that case is outside the parity contract).  The line's
 * k-mer must be exactly k long: compress_kmer packs the whole string (compress.rs:18-28) while
 * the node width is the global k (prelude.rs:21-25). */
int ko_add_read_bfc(ko_gir *g, const uint8_t *kmer, size_t len, uint32_t weight, int reverse_complement) {
    if (len < (size_t)g->k) return KO_ERR_SHORT_READ; /* pt_graph.rs:319 */
    if (len != (size_t)g->k || !all_acgt(kmer, len)) return KO_ERR_BAD_RECORD;
    uint8_t fw[2 * KO_MAXC], rv[2 * KO_MAXC];
    ko_compress_kmer_with_rev_compl(kmer, len, fw, rv);
    const uint8_t *both[2] = {fw, rv};
    for (int i = 0; i < (reverse_complement ? 2 : 1); ++i) {
        uint32_t source = find_or_insert_node(g, both[i]);
        uint32_t target = find_or_insert_node(g, both[i] + g->c);
        create_or_modify_edge_w(&g->nodes[source], target, kmer[len - 1], weight);
    }
    return KO_OK;
}

/* create_bfc (builder.rs:79-115): lines "<k-mer>\t<count>"; count < minimal_weight_threshold is
 * skipped BEFORE total += len (:106-109). */
int ko_create_from_bfc_files(ko_gir *g, const char *const *paths, int n_paths, int reverse_complement,
                             uint32_t minimal_weight_threshold, uint64_t *accepted_kmers,
                             uint64_t *accepted_bytes) {
    uint64_t nk = 0, nb = 0;
    int err = KO_OK;
    FILE **fs = (FILE **)calloc((size_t)n_paths, sizeof(FILE *));
    for (int i = 0; i < n_paths; ++i) { /* all readers are created up front: builder.rs:88-91 */
        fs[i] = fopen(paths[i], "rb");
        if (!fs[i]) { err = KO_ERR_IO; break; }
    }
    char line[4096];
    for (int i = 0; i < n_paths && err == KO_OK; ++i) {
        while (err == KO_OK && fgets(line, sizeof line, fs[i])) {
            size_t n = strlen(line);
            while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
            char *tab = strchr(line, '\t');
            if (!tab) { err = KO_ERR_BAD_RECORD; break; } /* iter.next().unwrap() :100 */
            *tab = 0;
            char *endp = NULL;
            unsigned long long w = strtoull(tab + 1, &endp, 10);
            if (endp == tab + 1 || (*endp && *endp != '\t') || w > 0xFFFFFFFFull || tab[1] == '-' || tab[1] == '+') {
                err = KO_ERR_BAD_RECORD; /* parse::<EdgeWeight>() error :100-105 */
                break;
            }
            if ((uint32_t)w < minimal_weight_threshold) continue; /* :106-108 */
            nb += (uint64_t)(tab - line);                        /* :109 */
            err = ko_add_read_bfc(g, (const uint8_t *)line, (size_t)(tab - line), (uint32_t)w, reverse_complement);
            if (err == KO_OK) ++nk;
        }
    }
    for (int i = 0; i < n_paths; ++i)
        if (fs[i]) fclose(fs[i]);
    free(fs);
    if (accepted_kmers) *accepted_kmers = nk;
    if (accepted_bytes) *accepted_bytes = nb;
    return err;
}

/* ------------------------------------------------------------------ stats */

/* stats/collections.rs:190-208: node_count = map.len(), edge_count = sum len */
void ko_counts(const ko_gir *g, uint64_t *nodes, uint64_t *edges) {
    uint64_t e = 0;
    for (uint64_t i = 0; i < g->n_nodes; ++i)
        if (!g->nodes[i].dead) e += g->nodes[i].nout;
    if (nodes) *nodes = g->live_nodes;
    if (edges) *edges = e;
}

/* what Stats for PtGraph reports after Convert (stats/collections.rs:137-168) */
void ko_collection_stats(const ko_gir *g, uint64_t out[8]) {
    uint32_t *indeg = (uint32_t *)calloc(g->n_nodes ? g->n_nodes : 1, sizeof(uint32_t));
    uint64_t edges = 0, maxw = 0, sumw = 0, max_out = 0, max_in = 0, n_src = 0, n_sink = 0;
    for (uint64_t i = 0; i < g->n_nodes; ++i) {
        const ko_node *n = &g->nodes[i];
        if (n->dead) continue;
        edges += n->nout;
        if (n->nout > max_out) max_out = n->nout;
        if (n->nout == 0) n_sink++;
        for (int j = 0; j < n->nout; ++j) {
            indeg[n->out[j].target]++;
            sumw += n->out[j].weight;
            if (n->out[j].weight > maxw) maxw = n->out[j].weight;
        }
    }
    for (uint64_t i = 0; i < g->n_nodes; ++i) {
        if (g->nodes[i].dead) continue;
        if (indeg[i] > max_in) max_in = indeg[i];
        if (indeg[i] == 0) n_src++;
    }
    free(indeg);
    out[0] = g->live_nodes; out[1] = edges; out[2] = maxw; out[3] = sumw;
    out[4] = max_in; out[5] = max_out; out[6] = n_src; out[7] = n_sink;
}

/* ------------------------------------------------------------------ clean */

static inline uint8_t node_symbol(const ko_gir *g, const uint8_t *key, int pos) {
    (void)g;
    return (uint8_t)((key[pos >> 2] >> (2 * (3 - (pos & 3)))) & 3);
}

/* pruner.rs:127-157: probe the 4 possible predecessors X + node[..k1-1] and
 * look for an outgoing edge that points at `id`. */
static int has_incoming_edges(const ko_gir *g, uint32_t id) {
    static const char order[4] = {'A', 'C', 'T', 'G'}; /* pruner.rs:144 */
    uint8_t ascii[64], packed[KO_MAXC];
    const uint8_t *key = g->nodes[id].key;
    static const char sym[4] = {'A', 'C', 'G', 'T'};
    for (int i = 0; i < g->k1 - 1; ++i) ascii[i + 1] = (uint8_t)sym[node_symbol(g, key, i)];
    for (int x = 0; x < 4; ++x) {
        ascii[0] = (uint8_t)order[x];
        ko_compress_node(ascii, (size_t)g->k1, packed);
        uint32_t p = find_node(g, packed);
        if (p == KO_NONE) continue;
        for (int j = 0; j < g->nodes[p].nout; ++j)
            if (g->nodes[p].out[j].target == id) return 1;
    }
    return 0;
}

/* pruner.rs:96-107: candidates are collected first, removed afterwards */
void ko_remove_single_vertices(ko_gir *g) {
    uint64_t n_rm = 0;
    uint32_t *rm = (uint32_t *)malloc((g->n_nodes ? g->n_nodes : 1) * sizeof(uint32_t));
    for (uint64_t i = 0; i < g->n_nodes; ++i)
        if (!g->nodes[i].dead && g->nodes[i].nout == 0 && !has_incoming_edges(g, (uint32_t)i))
            rm[n_rm++] = (uint32_t)i;
    for (uint64_t i = 0; i < n_rm; ++i) {
        g->nodes[rm[i]].dead = 1;
        g->live_nodes--;
    }
    free(rm);
    /* dead entries keep their slot as a tombstone; find_node skips them */
}

/* pruner.rs:109-118 + edges.rs:51-58: keep x.1 >= threshold */
void ko_remove_weak_edges(ko_gir *g, uint32_t threshold) {
    for (uint64_t i = 0; i < g->n_nodes; ++i) {
        ko_node *n = &g->nodes[i];
        if (n->dead) continue;
        int m = 0;
        for (int j = 0; j < n->nout; ++j)
            if (n->out[j].weight >= threshold) n->out[m++] = n->out[j];
        n->nout = (uint8_t)m;
    }
    ko_remove_single_vertices(g);
}

/* ------------------------------------------------------------ standardize */

/* `(x as f64 * p).round() as u32`: f64::round is half-away-from-zero == C
 * round(); the float->int cast saturates (Rust >= 1.45 semantics). */
static inline uint32_t scale_weight(uint32_t w, double p, uint32_t threshold) {
    double r = round((double)w * p);
    uint32_t nw;
    if (!(r >= 0.0)) nw = 0;
    else if (r >= 4294967295.0) nw = 0xFFFFFFFFu;
    else nw = (uint32_t)r;
    if (nw == 0 && w >= threshold) nw = 1; /* standardizer.rs:61-66 */
    return nw;
}

/* standardizer.rs:123-127 with the degenerate inputs rejected */
static int std_ratio(uint64_t G, uint64_t k, uint64_t s, uint64_t l, double *p) {
    if (G < k || s == l) return KO_ERR_DEGENERATE;
    *p = (double)(G - k) / (double)(s - l);
    return KO_OK;
}

uint64_t ko_standardize_weights(uint32_t *w, uint64_t n, uint64_t genome_len, uint64_t k,
                                uint32_t threshold, int *err) {
    uint64_t s = 0, l = 0, m = 0;
    for (uint64_t i = 0; i < n; ++i) { /* standardizer.rs:45-54 */
        s += w[i];
        if (w[i] < threshold) l += w[i];
    }
    double p;
    int e = std_ratio(genome_len, k, s, l, &p);
    if (err) *err = e;
    if (e != KO_OK) return n;
    for (uint64_t i = 0; i < n; ++i) {
        uint32_t nw = scale_weight(w[i], p, threshold);
        if (nw >= 1) w[m++] = nw; /* remove_weak_edges(1), standardizer.rs:69 */
    }
    return m;
}

/* standardizer.rs:42-70 applied to the GIR's edge set */
int ko_standardize_edges(ko_gir *g, uint64_t genome_len, uint64_t k, uint32_t threshold) {
    uint64_t s = 0, l = 0;
    for (uint64_t i = 0; i < g->n_nodes; ++i) {
        const ko_node *n = &g->nodes[i];
        if (n->dead) continue;
        for (int j = 0; j < n->nout; ++j) {
            s += n->out[j].weight;
            if (n->out[j].weight < threshold) l += n->out[j].weight;
        }
    }
    double p;
    int e = std_ratio(genome_len, k, s, l, &p);
    if (e != KO_OK) return e;
    for (uint64_t i = 0; i < g->n_nodes; ++i) {
        ko_node *n = &g->nodes[i];
        if (n->dead) continue;
        for (int j = 0; j < n->nout; ++j)
            n->out[j].weight = scale_weight(n->out[j].weight, p, threshold);
    }
    ko_remove_weak_edges(g, 1);
    return KO_OK;
}

/* ----------------------------------------------------------------- export */

static inline u128 node_int(const ko_gir *g, const uint8_t *key) {
    u128 v = 0;
    for (int i = 0; i < g->c; ++i) v = (v << 8) | key[i];
    return v >> g->padbits;
}

typedef struct {
    u128 key;
    uint32_t w;
} ko_kw;

static int cmp_kw(const void *a, const void *b) {
    const ko_kw *x = (const ko_kw *)a, *y = (const ko_kw *)b;
    return x->key < y->key ? -1 : (x->key > y->key ? 1 : 0);
}

static int cmp_u128(const void *a, const void *b) {
    u128 x = *(const u128 *)a, y = *(const u128 *)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* edge k-mer = source (k-1)-mer followed by the last symbol of the target
 * (this is what DebugHsGIR prints, hs_gir.rs:283-290) */
static ko_kw *collect_edges(const ko_gir *g, uint64_t *n_out) {
    uint64_t ne = 0;
    ko_counts(g, NULL, &ne);
    ko_kw *v = (ko_kw *)malloc((ne ? ne : 1) * sizeof(ko_kw));
    uint64_t m = 0;
    for (uint64_t i = 0; i < g->n_nodes; ++i) {
        const ko_node *n = &g->nodes[i];
        if (n->dead) continue;
        u128 src = node_int(g, n->key);
        for (int j = 0; j < n->nout; ++j) {
            u128 tgt = node_int(g, g->nodes[n->out[j].target].key);
            v[m].key = (src << 2) | (tgt & 3);
            v[m].w = n->out[j].weight;
            ++m;
        }
    }
    *n_out = m;
    return v;
}

uint64_t ko_export_edges(const ko_gir *g, uint64_t *key_hi, uint64_t *key_lo, uint32_t *weight,
                         uint64_t cap) {
    uint64_t n;
    ko_kw *v = collect_edges(g, &n);
    qsort(v, n, sizeof(ko_kw), cmp_kw);
    for (uint64_t i = 0; i < n && i < cap; ++i) {
        if (key_hi) key_hi[i] = (uint64_t)(v[i].key >> 64);
        if (key_lo) key_lo[i] = (uint64_t)v[i].key;
        if (weight) weight[i] = v[i].w;
    }
    free(v);
    return n;
}

uint64_t ko_export_nodes(const ko_gir *g, uint64_t *key_hi, uint64_t *key_lo, uint64_t cap) {
    u128 *v = (u128 *)malloc((g->live_nodes ? g->live_nodes : 1) * sizeof(u128));
    uint64_t m = 0;
    for (uint64_t i = 0; i < g->n_nodes; ++i)
        if (!g->nodes[i].dead) v[m++] = node_int(g, g->nodes[i].key);
    qsort(v, m, sizeof(u128), cmp_u128);
    for (uint64_t i = 0; i < m && i < cap; ++i) {
        if (key_hi) key_hi[i] = (uint64_t)(v[i] >> 64);
        if (key_lo) key_lo[i] = (uint64_t)v[i];
    }
    free(v);
    return m;
}

uint64_t ko_splitmix64(uint64_t x) {
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    return mix64(z);
}

void ko_digest(const ko_gir *g, uint64_t out[4]) {
    uint64_t n;
    ko_kw *v = collect_edges(g, &n);
    uint64_t d = 0, sw = 0, mw = 0;
    for (uint64_t i = 0; i < n; ++i) {
        uint64_t hi = (uint64_t)(v[i].key >> 64), lo = (uint64_t)v[i].key;
        d += ko_splitmix64(ko_splitmix64(hi) ^ lo) * (2ull * v[i].w + 1ull);
        sw += v[i].w;
        if (v[i].w > mw) mw = v[i].w;
    }
    free(v);
    out[0] = d; out[1] = n; out[2] = sw; out[3] = mw;
}

uint64_t ko_dump(const ko_gir *g, char *buf, uint64_t cap) {
    static const char sym[4] = {'A', 'C', 'G', 'T'};
    uint64_t n, pos = 0;
    ko_kw *v = collect_edges(g, &n);
    qsort(v, n, sizeof(ko_kw), cmp_kw);
    char line[160];
    for (uint64_t i = 0; i < n; ++i) {
        char kmer[65];
        for (int j = 0; j < g->k; ++j) kmer[j] = sym[(int)((v[i].key >> (2 * (g->k - 1 - j))) & 3)];
        kmer[g->k] = 0;
        int m = snprintf(line, sizeof line, "sequence %s weight %u\n", kmer, v[i].w);
        if (buf && pos + (uint64_t)m <= cap) memcpy(buf + pos, line, (size_t)m);
        pos += (uint64_t)m;
    }
    free(v);
    return pos;
}

/* -------------------------------------------------------- synthetic reads */
/* Our generator (not the reference's -- it has none).  Counter based so the
 * CUDA twin (katome_b200/csrc/synth.cuh) emits identical bytes:
 *   genome code i      = splitmix64(seed_g + i) & 3            (0123 -> ACGT)
 *   seed_r = seed_g ^ 0x5245414453, seed_e = seed_g ^ 0x4552524f52
 *   read r: start  = splitmix64(seed_r + 2r)   mod (G - L + 1)
 *           strand = splitmix64(seed_r + 2r+1) & 1   (1 => reverse complement)
 *   base j: h = splitmix64(seed_e + r*L + j); substituted iff
 *           h < err_ppm * floor(2^64 / 10^6); new = (orig + 1 + splitmix64(h) % 3) & 3 */
#define KO_PPM_UNIT 18446744073709ull

static inline uint8_t genome_code(uint64_t seed_g, uint64_t i) {
    return (uint8_t)(ko_splitmix64(seed_g + i) & 3);
}

void ko_synth_genome(uint64_t seed_g, uint64_t pos0, uint64_t n, uint8_t *out) {
    static const char sym[4] = {'A', 'C', 'G', 'T'};
    for (uint64_t i = 0; i < n; ++i) out[i] = (uint8_t)sym[genome_code(seed_g, pos0 + i)];
}

void ko_synth_reads(uint64_t seed_g, uint64_t G, uint32_t L, uint32_t err_ppm, uint64_t r0,
                    uint64_t r1, uint8_t *out) {
    static const char sym[4] = {'A', 'C', 'G', 'T'};
    const uint64_t seed_r = seed_g ^ 0x5245414453ull, seed_e = seed_g ^ 0x4552524f52ull;
    const uint64_t thr = (uint64_t)err_ppm * KO_PPM_UNIT;
    for (uint64_t r = r0; r < r1; ++r) {
        uint64_t start = ko_splitmix64(seed_r + 2 * r) % (G - L + 1);
        int strand = (int)(ko_splitmix64(seed_r + 2 * r + 1) & 1);
        uint8_t *dst = out + (r - r0) * L;
        for (uint32_t j = 0; j < L; ++j) {
            uint8_t code = strand ? (uint8_t)(3 - genome_code(seed_g, start + (L - 1 - j)))
                                  : genome_code(seed_g, start + j);
            uint64_t h = ko_splitmix64(seed_e + r * L + j);
            if (h < thr) code = (uint8_t)((code + 1 + ko_splitmix64(h) % 3) & 3);
            dst[j] = (uint8_t)sym[code];
        }
    }
}
