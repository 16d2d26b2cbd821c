/*
 * katome_oracle_mt.c -- the "optimistic CPU" line of SURVEY 8(d): what the host cores could do
 * with the reference's *result* but not its work shape.
 *
 * TEST INFRASTRUCTURE ONLY (see katome_oracle.h).  This is NOT a restatement of the reference:
 * katome is single-threaded by construction (static mut K_SIZE, prelude.rs:32-34; one global
 * RwLock around SEQUENCES, asm/mod.rs:21) and re-packs both (k-1)-mers of every window from ASCII
 * (compress.rs:18-73).  Here the same edge multiset -- weight[w] += 1 and weight[revcomp(w)] += 1
 * per window of an accepted read (hm_gir.rs:55-74, builder.rs:155-159) -- is counted the way a
 * CPU k-mer counter would: rolling 2-bit extraction, one canonical key per window, edge-keyed
 * open addressing, and every host thread used:
 *
 *   phase 1  thread t takes reads [n t/T, n (t+1)/T): ACGT filter, rolling k-mers, canonical
 *            key, owner = hash(key) -> T; keys appended to the (t, owner) run;
 *   phase 2  thread o inserts the T runs addressed to it into its private table;
 *   phase 3  per-thread digest terms (both strands expanded), summed by the caller thread.
 *
 * Its digest must equal ko_digest() of the faithful oracle on the same input
 * (tests/test_oracle_golden.py), which makes it a third independent implementation of the path.
 * bench.py reports it beside the faithful single-thread port, never instead of it.
 */
#ifndef MT_KEY /* ======================================================== common part */
#define _GNU_SOURCE
#include "katome_oracle.h"

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
} mt_table;

typedef struct {
    void *v;
    uint64_t n, cap;
} mt_run;

typedef struct mt_job {
    int k, rc, T;
    const uint8_t *bases;
    const uint64_t *offsets;
    uint64_t n_reads;
    mt_run *runs; /* T x T, runs[t * T + o]: keys thread t extracted for owner o */
    mt_table *tabs;
    uint64_t *acc_reads, *acc_bytes, *short_reads; /* per thread */
    uint64_t (*dig)[4];
    volatile int failed;
} mt_job;

typedef struct {
    mt_job *j;
    int t;
} mt_arg;

static inline uint64_t mt_hash(uint64_t lo, uint64_t hi) {
    return mt_mix64(lo ^ (hi * 0xA24BAED4963EE407ull) ^ 0x9E3779B97F4A7C15ull);
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

static int run_phase(mt_job *j, void *(*fn)(void *)) {
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

int ko_mt_build_digest(int k, const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads,
                       int reverse_complement, int n_threads, uint64_t out[4],
                       uint64_t *accepted_reads, uint64_t *accepted_bytes) {
    if (k < 2 || k > 64) return KO_ERR_BAD_K;
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    const int T = n_threads;
    mt_job j;
    memset(&j, 0, sizeof j);
    j.k = k;
    j.rc = reverse_complement != 0;
    j.T = T;
    j.bases = bases;
    j.offsets = offsets;
    j.n_reads = n_reads;
    j.runs = (mt_run *)calloc((size_t)T * T, sizeof(mt_run));
    j.tabs = (mt_table *)calloc(T, sizeof(mt_table));
    j.acc_reads = (uint64_t *)calloc(3 * (size_t)T, sizeof(uint64_t));
    j.dig = (uint64_t(*)[4])calloc(T, sizeof(uint64_t[4]));
    int rc_ = KO_OK;
    if (!j.runs || !j.tabs || !j.acc_reads || !j.dig) rc_ = KO_ERR_IO;
    if (rc_ == KO_OK) {
        j.acc_bytes = j.acc_reads + T;
        j.short_reads = j.acc_bytes + T;
        const int wide = k > 32;
        if (run_phase(&j, wide ? mt_extract_128 : mt_extract_64)) rc_ = KO_ERR_IO;
        if (rc_ == KO_OK && run_phase(&j, wide ? mt_count_128 : mt_count_64)) rc_ = KO_ERR_IO;
        if (rc_ == KO_OK && j.failed) rc_ = KO_ERR_IO;
    }
    if (rc_ == KO_OK) {
        uint64_t d = 0, ne = 0, sw = 0, mw = 0, nr = 0, nb = 0, shorts = 0;
        for (int t = 0; t < T; ++t) {
            d += j.dig[t][0];
            ne += j.dig[t][1];
            sw += j.dig[t][2];
            if (j.dig[t][3] > mw) mw = j.dig[t][3];
            nr += j.acc_reads[t];
            nb += j.acc_bytes[t];
            shorts += j.short_reads[t];
        }
        out[0] = d; out[1] = ne; out[2] = sw; out[3] = mw;
        if (accepted_reads) *accepted_reads = nr;
        if (accepted_bytes) *accepted_bytes = nb;
        if (shorts) rc_ = KO_ERR_SHORT_READ; /* hm_gir.rs:40: the reference panics, the build is void */
    }
    if (j.runs)
        for (size_t i = 0; i < (size_t)T * T; ++i) free(j.runs[i].v);
    if (j.tabs)
        for (int t = 0; t < T; ++t) {
            free(j.tabs[t].keys);
            free(j.tabs[t].w);
        }
    free(j.runs);
    free(j.tabs);
    free(j.acc_reads);
    free(j.dig);
    return rc_;
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
    mt_job *j = a->j;
    const int t = a->t, T = j->T, k = j->k;
    const MT_KEY mask = 2 * k == MT_BITS ? ~(MT_KEY)0 : (((MT_KEY)1 << (2 * k)) - 1);
    const uint64_t r0 = j->n_reads / T * t + (j->n_reads % T) * (uint64_t)t / T;
    const uint64_t r1 = j->n_reads / T * (t + 1) + (j->n_reads % T) * (uint64_t)(t + 1) / T;
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
    j->acc_reads[t] = nr;
    j->acc_bytes[t] = nb;
    j->short_reads[t] = shorts;
    return NULL;
}

/* phases 2 and 3: thread t owns the keys of runs[*][t]; count them, then the digest terms of its
 * table with both strands expanded (ko_digest's convention) */
static void *MT_NAME(mt_count)(void *p) {
    mt_arg *a = (mt_arg *)p;
    mt_job *j = a->j;
    const int t = a->t, T = j->T, k = j->k;
    if (j->failed) return NULL;
    uint64_t mine = 0;
    for (int s = 0; s < T; ++s) mine += j->runs[(size_t)s * T + t].n;
    uint64_t cap = 1024;
    while (cap < 2 * mine) cap <<= 1; /* load <= 0.5 even if every key were new */
    mt_table *tb = &j->tabs[t];
    tb->cap = cap;
    tb->keys = malloc(cap * sizeof(MT_KEY));
    tb->w = (uint32_t *)calloc(cap, sizeof(uint32_t));
    if (!tb->keys || !tb->w) {
        j->failed = 1;
        return NULL;
    }
    MT_KEY *keys = (MT_KEY *)tb->keys;
    memset(keys, 0xFF, cap * sizeof(MT_KEY));
    uint32_t special = 0; /* the all-ones key: T...T at full key width without rc */
    for (int s = 0; s < T; ++s) {
        const mt_run *rn = &j->runs[(size_t)s * T + t];
        const MT_KEY *v = (const MT_KEY *)rn->v;
        for (uint64_t q = 0; q < rn->n; ++q) {
            const MT_KEY key = v[q];
            if (key == MT_EMPTY) {
                ++special;
                continue;
            }
            /* a palindrome is inserted twice by hm_gir.rs:55-74 */
            const uint32_t inc = (j->rc && !(k & 1) && MT_NAME(revcomp)(key, k) == key) ? 2u : 1u;
            uint64_t i = (MT_NAME(key_hash)(key) >> 4) & (cap - 1); /* the top bits chose the owner */
            for (;;) {
                if (keys[i] == key) {
                    tb->w[i] += inc; /* u32, wrapping: EdgeWeight prelude.rs:9, hs_gir.rs:195 */
                    break;
                }
                if (keys[i] == MT_EMPTY) {
                    keys[i] = key;
                    tb->w[i] = inc;
                    ++tb->used;
                    break;
                }
                i = (i + 1) & (cap - 1);
            }
        }
    }
    uint64_t d = 0, ne = 0, sw = 0, mw = 0;
    for (uint64_t i = 0; i <= cap; ++i) {
        MT_KEY key;
        uint32_t w;
        if (i < cap) {
            key = keys[i];
            w = tb->w[i];
            if (key == MT_EMPTY || w == 0) continue;
        }
        else {
            if (!special) break;
            key = MT_EMPTY;
            w = special;
        }
        MT_KEY both[2] = {key, j->rc ? MT_NAME(revcomp)(key, k) : key};
        const int n = (j->rc && both[1] != key) ? 2 : 1;
        for (int q = 0; q < n; ++q) {
#if MT_BITS == 64
            const uint64_t hi = 0, lo = (uint64_t)both[q];
#else
            const uint64_t hi = (uint64_t)(both[q] >> 64), lo = (uint64_t)both[q];
#endif
            d += mt_splitmix64(mt_splitmix64(hi) ^ lo) * (2ull * w + 1ull);
            ++ne;
            sw += w;
        }
        if (w > mw) mw = w;
    }
    j->dig[t][0] = d;
    j->dig[t][1] = ne;
    j->dig[t][2] = sw;
    j->dig[t][3] = mw;
    return NULL;
}

#undef MT_EMPTY
#endif
