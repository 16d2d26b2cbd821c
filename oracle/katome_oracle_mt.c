/*
 * katome_oracle_mt.c -- the "optimistic CPU" line of SURVEY 8(d): what the host cores could do
 * with the reference's *result* but not its work shape -- and, because it finishes the BASELINE
 * configurations in seconds, the checker the full-size GPU builds are compared with.
 *
 * TEST INFRASTRUCTURE ONLY (see katome_oracle.h).  This is NOT a restatement of the reference:
 * katome is single-threaded by construction (static mut K_SIZE, prelude.rs:32-34; one global
 * RwLock around SEQUENCES, asm/mod.rs:21) and re-packs both (k-1)-mers of every window from ASCII
 * (compress.rs:18-73).  Here the same edge multiset -- weight[w] += 1 and weight[revcomp(w)] += 1
 * per window of an accepted read (hm_gir.rs:55-74, builder.rs:155-159) -- is counted the way a
 * CPU k-mer counter would: rolling 2-bit extraction, one canonical key per window, edge-keyed
 * open addressing, and every host thread used.  Input is taken in chunks of reads:
 *
 *   phase 1  thread t takes a slice of the chunk's reads: ACGT filter, rolling k-mers, canonical
 *            key, owner = hash(key) -> T; keys appended to the (t, owner) run;
 *   phase 2  thread o inserts the T runs addressed to it into its private table, which doubles
 *            whenever it is half full (so memory follows the DISTINCT keys, not the windows);
 *   queries  remove_weak_edges (edges.rs:51-58, pruner.rs:109-118), standardize_edges
 *            (standardizer.rs:42-70,123-127) and the digest (both strands expanded) run over the
 *            private tables, one thread each.
 *
 * Its digest must equal ko_digest() of the faithful oracle on the same input, before and after the
 * filter and the standardisation (tests/test_oracle_golden.py), which makes it a third independent
 * implementation of the path.  bench.py reports its rate beside the faithful single-thread port,
 * never instead of it.
 */
#ifndef MT_KEY /* ======================================================== common part */
#define _GNU_SOURCE
#include "katome_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

static inline uint64_t mt_mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t mt_splitmix64(uint64_t x) { return mt_mix64(x + 0x9E3779B97F4A7C15ull); }

/* compress.rs:347-378: A C G T -> 0 1 2 3; anything else rejects the whole read (builder.rs:155) */
static const int8_t MT_CODE[256] = {
    ['A'] = 1, ['C'] = 2, ['G'] = 3, ['T'] = 4, /* code + 1, 0 = not a base */
};

typedef struct {
    void *keys;
    uint32_t *w;
    uint64_t cap, used; /* cap is a power of two */
    uint32_t special;   /* weight of the all-ones key (T...T at full key width without rc) */
} mt_table;

typedef struct {
    void *v;
    uint64_t n, cap;
} mt_run;

struct ko_mt {
    int k, rc, T, wide;
    mt_run *runs; /* T x T, runs[t * T + o]: keys thread t extracted for owner o (reused per chunk) */
    mt_table *tabs;
    uint64_t reads, bytes, shorts;
    volatile int failed;
    /* the chunk being processed */
    const uint8_t *bases;
    const uint64_t *offsets;
    uint64_t r_begin, r_end;
    uint64_t *acc; /* 3 per thread: reads, bytes, short reads */
    /* query parameters */
    uint32_t q_threshold;
    double q_ratio;
    uint64_t (*q_out)[4];
};

typedef struct {
    ko_mt *j;
    int t;
} mt_arg;

static inline uint64_t mt_hash(uint64_t lo, uint64_t hi) {
    return mt_mix64(lo ^ (hi * 0xA24BAED4963EE407ull) ^ 0x9E3779B97F4A7C15ull);
}

/* standardizer.rs:56-66: (w as f64 * p).round() as u32 (saturating cast), 1 if that is 0 but w >= t */
static inline uint32_t mt_scale_weight(uint32_t w, double p, uint32_t threshold) {
    double r = round((double)w * p);
    uint32_t nw;
    if (!(r >= 0.0)) nw = 0;
    else if (r >= 4294967295.0) nw = 0xFFFFFFFFu;
    else nw = (uint32_t)r;
    if (nw == 0 && w >= threshold) nw = 1;
    return nw;
}

#define MT_KEY uint64_t
#define MT_BITS 64
#define MT_NAME(x) x##_64
#include "katome_oracle_mt.c"
#undef MT_KEY
#undef MT_BITS
#undef MT_NAME
#define MT_KEY u128
#define MT_BITS 128
#define MT_NAME(x) x##_128
#include "katome_oracle_mt.c"
#undef MT_KEY
#undef MT_BITS
#undef MT_NAME
#define MT_DONE 1

static int run_phase(ko_mt *j, void *(*fn)(void *)) {
    const int T = j->T;
    pthread_t *th = (pthread_t *)calloc(T, sizeof(pthread_t));
    mt_arg *args = (mt_arg *)calloc(T, sizeof(mt_arg));
    if (!th || !args) {
        free(th);
        free(args);
        return -1;
    }
    int started = 0;
    for (int t = 0; t < T; ++t) {
        args[t].j = j;
        args[t].t = t;
        if (pthread_create(&th[t], NULL, fn, &args[t]) != 0) break;
        ++started;
    }
    for (int t = started; t < T; ++t) fn(&args[t]); /* could not start a thread: do its share here */
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
    free(th);
    free(args);
    return 0;
}

ko_mt *ko_mt_new(int k, int reverse_complement, int n_threads) {
    if (k < 2 || k > 64) return NULL;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    ko_mt *j = (ko_mt *)calloc(1, sizeof(ko_mt));
    if (!j) return NULL;
    j->k = k;
    j->rc = reverse_complement != 0;
    j->T = n_threads;
    j->wide = k > 32;
    j->runs = (mt_run *)calloc((size_t)j->T * j->T, sizeof(mt_run));
    j->tabs = (mt_table *)calloc(j->T, sizeof(mt_table));
    j->acc = (uint64_t *)calloc(3 * (size_t)j->T, sizeof(uint64_t));
    j->q_out = (uint64_t(*)[4])calloc(j->T, sizeof(uint64_t[4]));
    if (!j->runs || !j->tabs || !j->acc || !j->q_out) {
        ko_mt_free(j);
        return NULL;
    }
    return j;
}

void ko_mt_free(ko_mt *j) {
    if (!j) return;
    if (j->runs)
        for (size_t i = 0; i < (size_t)j->T * j->T; ++i) free(j->runs[i].v);
    if (j->tabs)
        for (int t = 0; t < j->T; ++t) {
            free(j->tabs[t].keys);
            free(j->tabs[t].w);
        }
    free(j->runs);
    free(j->tabs);
    free(j->acc);
    free(j->q_out);
    free(j);
}

/* Build::add_read_fastaq over a batch + the accept rule and byte total of create_fastq
 * (builder.rs:152-160, hm_gir.rs:39-87).  KO_ERR_SHORT_READ voids the build (hm_gir.rs:40). */
int ko_mt_add_reads(ko_mt *j, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads) {
    if (!j) return KO_ERR_IO;
    if (j->failed) return KO_ERR_IO;
    j->bases = bases;
    j->offsets = offsets;
    const uint64_t CHUNK_BASES = 48ull << 20; /* bounds the runs: <= 48 Mi keys of 8 / 16 bytes in flight */
    uint64_t r = 0;
    while (r < n_reads) {
        uint64_t e = r + 1;
        while (e < n_reads && offsets[e + 1] - offsets[r] <= CHUNK_BASES) ++e;
        j->r_begin = r;
        j->r_end = e;
        for (size_t i = 0; i < (size_t)j->T * j->T; ++i) j->runs[i].n = 0;
        if (run_phase(j, j->wide ? mt_extract_128 : mt_extract_64)) return KO_ERR_IO;
        if (j->failed) return KO_ERR_IO;
        if (run_phase(j, j->wide ? mt_count_128 : mt_count_64)) return KO_ERR_IO;
        if (j->failed) return KO_ERR_IO;
        for (int t = 0; t < j->T; ++t) {
            j->reads += j->acc[3 * t];
            j->bytes += j->acc[3 * t + 1];
            j->shorts += j->acc[3 * t + 2];
        }
        r = e;
    }
    return j->shorts ? KO_ERR_SHORT_READ : KO_OK;
}

void ko_mt_counters(const ko_mt *j, uint64_t *accepted_reads, uint64_t *accepted_bytes) {
    if (accepted_reads) *accepted_reads = j->reads;
    if (accepted_bytes) *accepted_bytes = j->bytes;
}

/* out: D, |E|, sum w, max w over the both-strand expanded edge set (ko_digest's convention) */
int ko_mt_digest(ko_mt *j, uint64_t out[4]) {
    if (!j || j->failed) return KO_ERR_IO;
    if (run_phase(j, j->wide ? mt_digest_128 : mt_digest_64)) return KO_ERR_IO;
    uint64_t d = 0, ne = 0, sw = 0, mw = 0;
    for (int t = 0; t < j->T; ++t) {
        d += j->q_out[t][0];
        ne += j->q_out[t][1];
        sw += j->q_out[t][2];
        if (j->q_out[t][3] > mw) mw = j->q_out[t][3];
    }
    out[0] = d; out[1] = ne; out[2] = sw; out[3] = mw;
    return KO_OK;
}

/* Edges::remove_weak_edges (edges.rs:51-58) under Clean for HmGIR (pruner.rs:109-118): keep w >= t */
int ko_mt_remove_weak_edges(ko_mt *j, uint32_t threshold) {
    if (!j || j->failed) return KO_ERR_IO;
    j->q_threshold = threshold;
    if (run_phase(j, j->wide ? mt_filter_128 : mt_filter_64)) return KO_ERR_IO;
    return KO_OK;
}

/* standardize_edges (standardizer.rs:42-70) with the ratio of :123-127; degenerate inputs rejected
 * like ko_standardize_edges */
int ko_mt_standardize_edges(ko_mt *j, uint64_t genome_len, uint64_t k, uint32_t threshold) {
    if (!j || j->failed) return KO_ERR_IO;
    j->q_threshold = threshold;
    if (run_phase(j, j->wide ? mt_sums_128 : mt_sums_64)) return KO_ERR_IO;
    uint64_t s = 0, l = 0;
    for (int t = 0; t < j->T; ++t) {
        s += j->q_out[t][0];
        l += j->q_out[t][1];
    }
    if (genome_len < k || s == l) return KO_ERR_DEGENERATE;
    j->q_ratio = (double)(genome_len - k) / (double)(s - l);
    if (run_phase(j, j->wide ? mt_scale_128 : mt_scale_64)) return KO_ERR_IO;
    return KO_OK;
}

int ko_mt_build_digest(int k, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                       int reverse_complement, int n_threads, uint64_t out[4],
                       uint64_t *accepted_reads, uint64_t *accepted_bytes) {
    if (k < 2 || k > 64) return KO_ERR_BAD_K;
    ko_mt *j = ko_mt_new(k, reverse_complement, n_threads);
    if (!j) return KO_ERR_IO;
    int rc_ = ko_mt_add_reads(j, bases, offsets, n_reads);
    if (rc_ == KO_OK || rc_ == KO_ERR_SHORT_READ) {
        int d = ko_mt_digest(j, out);
        if (d != KO_OK) rc_ = d;
        ko_mt_counters(j, accepted_reads, accepted_bytes);
    }
    ko_mt_free(j);
    return rc_;
}

/* ko_synth_reads over n_threads host threads (the generator is counter based, so any split of
 * [r0, r1) writes the same bytes) */
typedef struct {
    uint64_t seed_g, G, r0, r1;
    uint32_t L, err_ppm;
    uint8_t *out;
} synth_arg;
static void *synth_worker(void *p) {
    synth_arg *a = (synth_arg *)p;
    ko_synth_reads(a->seed_g, a->G, a->L, a->err_ppm, a->r0, a->r1, a->out);
    return NULL;
}
void ko_synth_reads_mt(uint64_t seed_g, uint64_t G, uint32_t L, uint32_t err_ppm, uint64_t r0, uint64_t r1,
                       uint8_t *out, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 256) n_threads = 256;
    pthread_t th[256];
    synth_arg args[256];
    const uint64_t n = r1 - r0;
    int started = 0;
    for (int t = 0; t < n_threads; ++t) {
        const uint64_t a = r0 + n / n_threads * t + (n % n_threads) * (uint64_t)t / n_threads;
        const uint64_t b = r0 + n / n_threads * (t + 1) + (n % n_threads) * (uint64_t)(t + 1) / n_threads;
        args[t] = (synth_arg){seed_g, G, a, b, L, err_ppm, out + (a - r0) * L};
        if (started == t && pthread_create(&th[t], NULL, synth_worker, &args[t]) == 0) ++started;
        else synth_worker(&args[t]);
    }
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

#elif !defined(MT_DONE) /* ============================== per key width (included twice above) */

#define MT_EMPTY (~(MT_KEY)0) /* a canonical key is never all ones; without rc it is counted apart */

static inline uint64_t MT_NAME(key_hash)(MT_KEY key) {
#if MT_BITS == 64
    return mt_hash((uint64_t)key, 0);
#else
    return mt_hash((uint64_t)key, (uint64_t)(key >> 64));
#endif
}

static inline MT_KEY MT_NAME(revcomp)(MT_KEY x, int k) { /* == packing the reverse complement string (compress.rs:153-169) */
    MT_KEY r = 0;
    for (int i = 0; i < k; ++i) {
        r = (r << 2) | (3 - (x & 3));
        x >>= 2;
    }
    return r;
}

/* phase 1: extraction (builder.rs:152-160 + the window loop of hm_gir.rs:55-85, rolling) */
static void *MT_NAME(mt_extract)(void *p) {
    mt_arg *a = (mt_arg *)p;
    ko_mt *j = a->j;
    const int t = a->t, T = j->T, k = j->k;
    const MT_KEY mask = 2 * k == MT_BITS ? ~(MT_KEY)0 : (((MT_KEY)1 << (2 * k)) - 1);
    const uint64_t n = j->r_end - j->r_begin;
    const uint64_t r0 = j->r_begin + n / T * t + (n % T) * (uint64_t)t / T;
    const uint64_t r1 = j->r_begin + n / T * (t + 1) + (n % T) * (uint64_t)(t + 1) / T;
    uint64_t nr = 0, nb = 0, shorts = 0;
    for (uint64_t r = r0; r < r1 && !j->failed; ++r) {
        const uint8_t *s = j->bases + j->offsets[r];
        const uint64_t len = j->offsets[r + 1] - j->offsets[r];
        int ok = 1;
        for (uint64_t i = 0; i < len; ++i) ok &= MT_CODE[s[i]] != 0; /* builder.rs:155 */
        if (!ok) continue;
        ++nr;
        nb += len; /* builder.rs:158 */
        if (len < (uint64_t)k) {
            ++shorts;
            continue;
        }
        MT_KEY fw = 0, rv = 0;
        for (uint64_t i = 0; i < len; ++i) {
            const MT_KEY c = (MT_KEY)(MT_CODE[s[i]] - 1);
            fw = ((fw << 2) | c) & mask;
            rv = (rv >> 2) | ((3 - c) << (2 * (k - 1)));
            if (i + 1 < (uint64_t)k) continue;
            const MT_KEY key = (j->rc && rv < fw) ? rv : fw;
            const int o = (int)(((MT_NAME(key_hash)(key) >> 44) * (uint64_t)T) >> 20);
            mt_run *rn = &j->runs[(size_t)t * T + o];
            if (rn->n == rn->cap) {
                const uint64_t nc = rn->cap ? rn->cap * 2 : 4096;
                void *nv = realloc(rn->v, nc * sizeof(MT_KEY));
                if (!nv) {
                    j->failed = 1;
                    break;
                }
                rn->v = nv;
                rn->cap = nc;
            }
            ((MT_KEY *)rn->v)[rn->n++] = key;
        }
    }
    j->acc[3 * t] = nr;
    j->acc[3 * t + 1] = nb;
    j->acc[3 * t + 2] = shorts;
    return NULL;
}

static inline uint64_t MT_NAME(find_slot)(const MT_KEY *keys, uint64_t cap, MT_KEY key) {
    uint64_t i = (MT_NAME(key_hash)(key) >> 4) & (cap - 1); /* the top bits chose the owner */
    while (keys[i] != key && keys[i] != MT_EMPTY) i = (i + 1) & (cap - 1);
    return i;
}

static int MT_NAME(grow)(mt_table *tb) {
    const uint64_t nc = tb->cap ? tb->cap * 2 : 1024;
    MT_KEY *nk = (MT_KEY *)malloc(nc * sizeof(MT_KEY));
    uint32_t *nw = (uint32_t *)calloc(nc, sizeof(uint32_t));
    if (!nk || !nw) {
        free(nk);
        free(nw);
        return -1;
    }
    memset(nk, 0xFF, nc * sizeof(MT_KEY));
    const MT_KEY *ok = (const MT_KEY *)tb->keys;
    for (uint64_t i = 0; i < tb->cap; ++i) {
        if (ok[i] == MT_EMPTY) continue;
        const uint64_t s = MT_NAME(find_slot)(nk, nc, ok[i]);
        nk[s] = ok[i];
        nw[s] = tb->w[i];
    }
    free(tb->keys);
    free(tb->w);
    tb->keys = nk;
    tb->w = nw;
    tb->cap = nc;
    return 0;
}

/* phase 2: thread t owns the keys of runs[*][t] */
static void *MT_NAME(mt_count)(void *p) {
    mt_arg *a = (mt_arg *)p;
    ko_mt *j = a->j;
    const int t = a->t, T = j->T, k = j->k;
    if (j->failed) return NULL;
    mt_table *tb = &j->tabs[t];
    if (!tb->cap && MT_NAME(grow)(tb)) {
        j->failed = 1;
        return NULL;
    }
    for (int s = 0; s < T; ++s) {
        const mt_run *rn = &j->runs[(size_t)s * T + t];
        const MT_KEY *v = (const MT_KEY *)rn->v;
        for (uint64_t q = 0; q < rn->n; ++q) {
            const MT_KEY key = v[q];
            if (key == MT_EMPTY) {
                ++tb->special;
                continue;
            }
            /* a palindrome is inserted twice by hm_gir.rs:55-74 */
            const uint32_t inc = (j->rc && !(k & 1) && MT_NAME(revcomp)(key, k) == key) ? 2u : 1u;
            MT_KEY *keys = (MT_KEY *)tb->keys;
            uint64_t i = MT_NAME(find_slot)(keys, tb->cap, key);
            if (keys[i] == MT_EMPTY) {
                if (2 * (tb->used + 1) > tb->cap) { /* memory follows the distinct keys */
                    if (MT_NAME(grow)(tb)) {
                        j->failed = 1;
                        return NULL;
                    }
                    keys = (MT_KEY *)tb->keys;
                    i = MT_NAME(find_slot)(keys, tb->cap, key);
                }
                keys[i] = key;
                ++tb->used;
            }
            tb->w[i] += inc; /* u32, wrapping: EdgeWeight prelude.rs:9, hs_gir.rs:195 */
        }
    }
    return NULL;
}

/* Visit the live entries of thread t's table: f(key, &w, mult) with mult = strands the entry stands for */
#define MT_FOR_EACH(...)                                                            \
    mt_arg *a = (mt_arg *)p;                                                         \
    ko_mt *j = a->j;                                                                 \
    const int t = a->t, k = j->k;                                                    \
    mt_table *tb = &j->tabs[t];                                                      \
    MT_KEY *keys = (MT_KEY *)tb->keys;                                               \
    for (uint64_t i = 0; i <= tb->cap; ++i) {                                        \
        MT_KEY key;                                                                  \
        uint32_t *wp;                                                                \
        if (i < tb->cap) {                                                           \
            key = keys[i];                                                           \
            wp = &tb->w[i];                                                          \
            if (key == MT_EMPTY) continue;                                           \
        }                                                                            \
        else {                                                                       \
            key = MT_EMPTY;                                                          \
            wp = &tb->special;                                                       \
        }                                                                            \
        if (*wp == 0) continue; /* removed by the filter (or never seen) */          \
        const MT_KEY other = j->rc ? MT_NAME(revcomp)(key, k) : key;                 \
        const int mult = (j->rc && other != key) ? 2 : 1;                            \
        (void)mult;                                                                  \
        __VA_ARGS__                                                                  \
    }

static void *MT_NAME(mt_digest)(void *p) {
    uint64_t d = 0, ne = 0, sw = 0, mw = 0;
    MT_FOR_EACH({
        const MT_KEY both[2] = {key, other};
        for (int q = 0; q < mult; ++q) {
#if MT_BITS == 64
            const uint64_t hi = 0, lo = (uint64_t)both[q];
#else
            const uint64_t hi = (uint64_t)(both[q] >> 64), lo = (uint64_t)both[q];
#endif
            d += mt_splitmix64(mt_splitmix64(hi) ^ lo) * (2ull * *wp + 1ull);
            ++ne;
            sw += *wp;
        }
        if (*wp > mw) mw = *wp;
    })
    j->q_out[t][0] = d;
    j->q_out[t][1] = ne;
    j->q_out[t][2] = sw;
    j->q_out[t][3] = mw;
    return NULL;
}

static void *MT_NAME(mt_filter)(void *p) {
    MT_FOR_EACH({
        if (*wp < j->q_threshold) *wp = 0;
    })
    return NULL;
}

static void *MT_NAME(mt_sums)(void *p) { /* standardizer.rs:45-54 over both strands */
    uint64_t s = 0, l = 0;
    MT_FOR_EACH({
        s += (uint64_t)mult * *wp;
        if (*wp < j->q_threshold) l += (uint64_t)mult * *wp;
    })
    j->q_out[t][0] = s;
    j->q_out[t][1] = l;
    return NULL;
}

static void *MT_NAME(mt_scale)(void *p) { /* standardizer.rs:56-69: scale, then remove_weak_edges(1) */
    MT_FOR_EACH({
        *wp = mt_scale_weight(*wp, j->q_ratio, j->q_threshold);
    })
    return NULL;
}

#undef MT_FOR_EACH
#undef MT_EMPTY
#endif
