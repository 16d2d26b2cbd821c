// multi.cuh -- ONE handle over several GPUs of this process (ktg_config.n_devices / device_ids).
//
// What katome calls is one function: Build::create (algorithms/builder.rs:42-54, from
// asm/basic_assembler.rs:31-45).  With n_devices > 1 that call builds a table that is hash-sharded
// over the devices: the reads of every ktg_add_reads call are split into one contiguous
// share per device; every device packs and extracts its share and writes each k-mer (or
// super-k-mer record, 23 <= k <= 31 from 4 devices on) straight into the HBM of the owning
// device over NVLink peer memory (the fused exchange of builder.cuh / superkmer.cuh); the owner
// partitions what it received and inserts it exactly as on one GPU.  The control plane is this
// file: one host thread per device, a barrier between "all writers done" and "owners consume",
// and a few words of shared host memory where the per-rank process version (katome_b200/dist.py)
// needs NCCL collectives.  No torch, no NCCL, no IPC: peers are mapped with
// cudaDeviceEnablePeerAccess and addressed by plain pointers.
//
// Queries answer for the whole graph: counts / digests / sums add up over the disjoint shards;
// node statistics merge every shard's (node, degree word) pairs on the first device; the exports
// (hm_gir.rs:156-226) gather the shards' edges on the first device -- every shard's compaction
// kernel writes its edges directly into that device's arrays -- where they are sorted and the
// nodes numbered globally.  A device may be listed more than once: the shards then share a GPU,
// which is how the tests run 2-4 shards on a one-GPU box.
#pragma once
#include <condition_variable>
#include <mutex>
#include <thread>

#include "builder.cuh"
#include "host_plan.h"

namespace ktg {

// Reusable barrier of n host threads that also carries a failure flag: every participant keeps
// calling wait() in the same sequence; once anybody reports !ok, every wait returns false and all
// of them leave at the same point (nobody is left waiting for a peer that bailed out).
struct Gate {
    std::mutex m;
    std::condition_variable cv;
    uint32_t n = 0, count = 0;
    uint64_t gen = 0;
    bool failed = false;
    void reset(uint32_t n_) {
        n = n_;
        count = 0;
        failed = false;
    }
    bool wait(bool ok = true) {
        std::unique_lock<std::mutex> lk(m);
        if (!ok) failed = true;
        const uint64_t g = gen;
        if (++count == n) {
            count = 0;
            ++gen;
            cv.notify_all();
        }
        else cv.wait(lk, [&] { return gen != g; });
        return !failed;
    }
};

struct MultiBuilder {
    static constexpr int PEER_FAILED = -1; // internal: "another shard failed, I left at the barrier"
    ktg_config cfg{};
    uint32_t n = 0, k = 0;
    bool rc = false;
    std::vector<int> dev;
    std::vector<std::unique_ptr<BuilderBase>> sh;
    Gate gate;
    bool use_skm = false;
    int deferred_error = KTG_OK;
    std::string deferred_msg;
    // per-shard staging for host input (two buffers each) and small device scratch
    struct Stage {
        DeviceBuf bases[2], offs[2], ends, sketch;
        cudaEvent_t ready[2] = {nullptr, nullptr}, consumed[2] = {nullptr, nullptr};
        bool used[2] = {false, false};
    };
    std::vector<Stage> stage;
    uint64_t exchanged_bytes = 0;

    ~MultiBuilder() {
        for (uint32_t i = 0; i < n && i < stage.size(); ++i) {
            cudaSetDevice(dev[i]);
            if (sh[i]) sh[i]->sync_stream();
            for (int b = 0; b < 2; ++b) {
                stage[i].bases[b].release();
                stage[i].offs[b].release();
                if (stage[i].ready[b]) cudaEventDestroy(stage[i].ready[b]);
                if (stage[i].consumed[b]) cudaEventDestroy(stage[i].consumed[b]);
            }
            stage[i].ends.release();
            stage[i].sketch.release();
            sh[i].reset();
        }
    }

    // f(i) on one host thread per shard (the shard's device current); the first failure wins
    template <class F> int parallel(F f) {
        std::vector<int> rcs(n, KTG_OK);
        std::vector<std::string> msgs(n);
        std::vector<std::thread> th;
        auto body = [&](uint32_t i) {
            if (cudaSetDevice(dev[i]) != cudaSuccess) {
                rcs[i] = KTG_ERR_CUDA;
                msgs[i] = "cudaSetDevice failed";
                return;
            }
            rcs[i] = f(i);
            if (rcs[i] != KTG_OK) msgs[i] = last_error_ref(); // thread local: carry it to the caller
        };
        for (uint32_t i = 1; i < n; ++i) {
            try {
                th.emplace_back(body, i);
            } catch (const std::system_error &) {
                body(i); // only safe for bodies without barriers; a box that cannot start 8 threads has other problems
            }
        }
        body(0);
        for (auto &t : th) t.join();
        cudaSetDevice(dev[0]);
        for (uint32_t i = 0; i < n; ++i) // the shard that failed, not the ones that left because it did
            if (rcs[i] != KTG_OK && rcs[i] != PEER_FAILED) return fail(rcs[i], "%s (device %d, shard %u)", msgs[i].c_str(), dev[i], i);
        for (uint32_t i = 0; i < n; ++i)
            if (rcs[i] != KTG_OK) return fail(KTG_ERR_INVALID, "shard %u stopped without an error of its own", i);
        return KTG_OK;
    }

    int init(const ktg_config &c) {
        cfg = c;
        n = c.n_devices;
        k = c.k;
        rc = c.reverse_complement != 0;
        if (n < 2 || n > (uint32_t)MAX_P2P_WORLD) return fail(KTG_ERR_INVALID, "n_devices must be 2..%d", MAX_P2P_WORLD);
        if (!c.device_ids) return fail(KTG_ERR_INVALID, "device_ids is null");
        int have = 0;
        KTG_CUDA(cudaGetDeviceCount(&have));
        dev.assign(c.device_ids, c.device_ids + n);
        for (int d : dev)
            if (d < 0 || d >= have) return fail(KTG_ERR_NO_DEVICE, "device %d of device_ids does not exist (%d visible)", d, have);
        // every device reads and writes every other device's receive buffer and the first device's export arrays
        for (uint32_t i = 0; i < n; ++i)
            for (uint32_t j = 0; j < n; ++j) {
                if (dev[i] == dev[j]) continue;
                int can = 0;
                KTG_CUDA(cudaDeviceCanAccessPeer(&can, dev[i], dev[j]));
                if (!can) return fail(KTG_ERR_INVALID, "device %d cannot map the memory of device %d (no peer access)", dev[i], dev[j]);
                KTG_CUDA(cudaSetDevice(dev[i]));
                cudaError_t e = cudaDeviceEnablePeerAccess(dev[j], 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) (void)cudaGetLastError();
                else if (e != cudaSuccess) return fail(KTG_ERR_CUDA, "cudaDeviceEnablePeerAccess(%d -> %d): %s", dev[i], dev[j], cudaGetErrorString(e));
            }
        // super-k-mer records from 4 shards on, where the key exchange is NVLink bound
        use_skm = skm_supported(k) && n >= 4;
        sh.resize(n);
        stage.resize(n);
        gate.reset(n);
        return parallel([&](uint32_t i) -> int {
            std::unique_ptr<BuilderBase> b;
            if (k <= 32) b.reset(new Builder<uint64_t>());
            else b.reset(new Builder<u128>());
            b->cfg = c;
            b->cfg.world_size = n;
            b->cfg.rank = i;
            b->cfg.device = dev[i];
            b->cfg.stream = nullptr;
            b->cfg.flags |= KTG_FLAG_FORCE_PARTITION; // the fused exchange has no unpartitioned mode
            b->k = k;
            b->rc = rc;
            b->device = dev[i];
            b->stream = nullptr;
            b->prof.enabled = (c.flags & KTG_FLAG_PROFILE) != 0;
            KTG_TRY(b->init());
            for (int s = 0; s < 2; ++s) {
                KTG_CUDA(cudaEventCreateWithFlags(&stage[i].ready[s], cudaEventDisableTiming));
                KTG_CUDA(cudaEventCreateWithFlags(&stage[i].consumed[s], cudaEventDisableTiming));
            }
            KTG_TRY(stage[i].ends.ensure(64 * 8));
            KTG_TRY(stage[i].sketch.ensure(HLL_M * 4));
            sh[i] = std::move(b);
            return KTG_OK;
        });
    }

    int check_state() {
        if (deferred_error != KTG_OK) return fail(deferred_error, "build is void after an earlier error: %s", deferred_msg.c_str());
        return KTG_OK;
    }
    int poison(int code) { // an error inside the exchange leaves the shards out of step: the build is void
        if (code != KTG_OK && deferred_error == KTG_OK) {
            deferred_error = code;
            deferred_msg = last_error_ref();
        }
        return code;
    }

    // ---- Build::add_read_fastaq over a batch (builder.rs:152-160, hm_gir.rs:39-87), host input ----
    int add_reads(const uint8_t *bases, const uint64_t *offsets, uint64_t n_reads, uint64_t *acc_reads, uint64_t *acc_bytes) {
        KTG_TRY(check_state());
        if (n_reads == 0) return KTG_OK;
        touch();
        uint64_t r0c = 0, b0c = 0;
        if (acc_reads || acc_bytes) KTG_TRY(read_counters(&r0c, &b0c));
        // shares: contiguous blocks of reads with about the same number of bases
        std::vector<uint64_t> cut(n + 1, 0);
        const uint64_t total = offsets[n_reads] - offsets[0];
        for (uint32_t i = 1; i < n; ++i) {
            const uint64_t want = offsets[0] + total / n * i;
            cut[i] = (uint64_t)(std::lower_bound(offsets, offsets + n_reads, want) - offsets);
            if (cut[i] < cut[i - 1]) cut[i] = cut[i - 1];
        }
        cut[n] = n_reads;
        const uint64_t CHUNK = (uint64_t)std::max(1, sh[0]->tune.chunk_mb) << 20;
        // chunks of every share and what their offsets say about them (done by the shard's own thread)
        struct Plan {
            std::vector<uint64_t> c;     // chunk c = reads [c[j], c[j+1])
            std::vector<BatchHint> hint; // one read length or ragged, windows if every read is accepted
            std::vector<char> have;      // hint[c] has been worked out
        };
        auto hint_of = [&offsets, kk = k](Plan &p, uint32_t c) -> const BatchHint & {
            if (!p.have[c]) {
                const uint64_t r = p.c[c], e = p.c[c + 1], len0 = offsets[r + 1] - offsets[r];
                uint64_t diff = 0, wub = 0;
                for (uint64_t q = r; q < e; ++q) diff |= (offsets[q + 1] - offsets[q]) ^ len0;
                if (diff == 0) wub = len0 >= kk ? (e - r) * (len0 - kk + 1) : 0;
                else
                    for (uint64_t q = r; q < e; ++q) {
                        const uint64_t len = offsets[q + 1] - offsets[q];
                        wub += len >= kk ? len - kk + 1 : 0;
                    }
                p.hint[c].ulen = (diff == 0 && len0 <= 0xFFFFFFFFull) ? (uint32_t)len0 : 0;
                p.hint[c].windows_ub = wub;
                p.have[c] = 1;
            }
            return p.hint[c];
        };
        std::vector<Plan> plan(n);
        std::vector<uint64_t> max_win(n, 0), sum_win(n, 0);
        std::vector<uint32_t> rounds_of(n, 0), geo_sub(n, 0), geo_log2(n, 0);
        // published between barriers
        std::vector<void *> rx(n, nullptr);
        std::vector<uint64_t> cap(n, 0);
        std::vector<std::vector<unsigned long long>> cursors(n, std::vector<unsigned long long>(n, 0));
        std::vector<std::vector<unsigned long long>> key_counts(n, std::vector<unsigned long long>(n + 1, 0));
        std::vector<std::vector<uint32_t>> sketches(n, std::vector<uint32_t>(use_skm ? 0 : HLL_M, 0));
        std::vector<uint64_t> n_spill(n, 0);
        std::vector<std::vector<std::vector<uint64_t>>> spill_out(n, std::vector<std::vector<uint64_t>>(n)); // [src][dst] words
        const uint32_t kk = k;
        const bool skm = use_skm;
        gate.reset(n);
        int rc_ = parallel([&](uint32_t i) -> int {
            BuilderBase *b = sh[i].get();
            Stage &st = stage[i];
            int err = KTG_OK;
            auto step = [&](int e) { // remember the first error, keep walking so that the barriers line up
                if (err == KTG_OK && e != KTG_OK) err = e;
                return err == KTG_OK;
            };
            // -- plan my share: chunk boundaries by binary search over the offsets; what a chunk's offsets say
            // about it (one read length or ragged, windows if every read is accepted) is worked out when the
            // chunk comes up, under the previous chunk's kernels (up front it was 15 ms for 7.7 M reads, a
            // third of a C3 build on 2 GPUs); the buffers are sized from the bases, an upper bound of the windows
            {
                Plan &p = plan[i];
                uint64_t r = cut[i];
                p.c.push_back(r);
                while (r < cut[i + 1]) {
                    uint64_t e = (uint64_t)(std::upper_bound(offsets + r + 1, offsets + cut[i + 1] + 1, offsets[r] + CHUNK) - offsets) - 1;
                    if (e <= r) e = r + 1; // a read longer than a chunk travels alone
                    const uint64_t nb = offsets[e] - offsets[r];
                    max_win[i] = std::max(max_win[i], nb);
                    sum_win[i] += nb;
                    p.c.push_back(e);
                    r = e;
                }
                p.hint.assign(p.c.size() - 1, BatchHint());
                p.have.assign(p.c.size() - 1, 0);
                rounds_of[i] = (uint32_t)p.hint.size();
                int needs = 0;
                step(b->mgd_plan(1, &needs, &geo_sub[i], &geo_log2[i]));
            }
            if (!gate.wait(err == KTG_OK)) return err != KTG_OK ? err : PEER_FAILED;
            uint64_t gmax = 0, gsum = 0;
            uint32_t rounds = 0;
            bool same_geo = true;
            for (uint32_t j = 0; j < n; ++j) {
                gmax = std::max(gmax, max_win[j]);
                gsum = std::max(gsum, sum_win[j]);
                rounds = std::max(rounds, rounds_of[j]);
                same_geo = same_geo && geo_sub[j] == geo_sub[0] && geo_log2[j] == geo_log2[0];
            }
            if (gmax == 0) return KTG_OK; // nothing to insert anywhere (everybody sees the same gmax)
            // The direct exchange (every shard has the same geometry, no super-k-mers): the senders stream
            // all their chunks into the owners' (sender, sub-table) buckets without any barrier in between,
            // and the owners go straight to their page level, once, at the end.
            // Default below 4 devices (mg_direct = -1; 0 / 1 = never / whenever possible), like the per-rank
            // processes: C3 over 2 B200s end to end 31.7 ms against 34.4 with the key exchange; from 4 devices on
            // the super-k-mer exchange is faster (the direct sender is NVLink bound).
            const bool want_direct = b->tune.mg_direct < 0 ? n < 4 : b->tune.mg_direct != 0;
            if (!skm && same_geo && want_direct && (uint64_t)n * geo_sub[0] <= 1024) {
                const uint32_t n_sub = geo_sub[0];
                {
                    void *base = nullptr;
                    uint64_t bytes = 0, bc = 0;
                    step(b->mgd_prepare(gsum, &base, &bytes, &bc));
                    rx[i] = base;
                    cap[i] = bc;
                }
                if (!gate.wait(err == KTG_OK)) return err != KTG_OK ? err : PEER_FAILED;
                Plan &p = plan[i];
                std::vector<unsigned long long> cur_host((size_t)n * n_sub, 0);
                void *d_cur = nullptr;
                auto copy_chunk = [&](uint32_t c) -> int {
                    if (c >= p.hint.size()) return KTG_OK;
                    const int s = (int)(c & 1);
                    const uint64_t r = p.c[c], r1 = p.c[c + 1], nb = offsets[r1] - offsets[r], nr = r1 - r;
                    if (st.used[s]) KTG_CUDA(cudaStreamWaitEvent(b->copy_stream, st.consumed[s], 0));
                    if (st.bases[s].cap < nb + 64 || st.offs[s].cap < (nr + 1) * 8) {
                        KTG_TRY(b->sync_stream());
                        KTG_CUDA(cudaStreamSynchronize(b->copy_stream));
                        KTG_TRY(st.bases[s].ensure(std::max<uint64_t>(nb, CHUNK) + 64));
                        KTG_TRY(st.offs[s].ensure((nr + 1) * 8));
                    }
                    KTG_CUDA(cudaMemcpyAsync(st.bases[s].p, bases + offsets[r], nb, cudaMemcpyHostToDevice, b->copy_stream));
                    if (hint_of(p, c).ulen) {
                        const int g = (int)std::min<uint64_t>((nr + 1 + 255) / 256, 4096);
                        fill_offsets_kernel<<<g, 256, 0, b->copy_stream>>>((uint64_t *)st.offs[s].p, nr + 1, offsets[r], hint_of(p, c).ulen);
                        KTG_CUDA(cudaGetLastError());
                    }
                    else KTG_CUDA(cudaMemcpyAsync(st.offs[s].p, offsets + r, (nr + 1) * 8, cudaMemcpyHostToDevice, b->copy_stream));
                    KTG_CUDA(cudaEventRecord(st.ready[s], b->copy_stream));
                    st.used[s] = true;
                    return KTG_OK;
                };
                step(copy_chunk(0));
                if (p.hint.empty() && err == KTG_OK) // nothing of mine: my cursors still have to say so
                    step(b->mgd_scatter_reads(nullptr, nullptr, 0, 0, rx.data(), 0, 1, nullptr, &d_cur));
                for (uint32_t c = 0; c < p.hint.size() && err == KTG_OK; ++c) {
                    step(copy_chunk(c + 1)); // the next chunk's copy runs under this chunk's kernels
                    const int s = (int)(c & 1);
                    const uint64_t r = p.c[c], r1 = p.c[c + 1];
                    step(cudaStreamWaitEvent(b->stream, st.ready[s], 0) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "cudaStreamWaitEvent failed"));
                    b->input_consumed = st.consumed[s];
                    b->hint_shift0 = (uint32_t)((uintptr_t)st.bases[s].p & 31);
                    if (err == KTG_OK)
                        step(b->mgd_scatter_reads((const uint8_t *)st.bases[s].p - offsets[r], (const uint64_t *)st.offs[s].p, r1 - r,
                                                  offsets[r1] - offsets[r], rx.data(), 0, c == 0, nullptr, &d_cur, &hint_of(p, c)));
                    b->input_consumed = nullptr;
                }
                if (err == KTG_OK) {
                    step(cudaMemcpyAsync(cur_host.data(), d_cur, cur_host.size() * 8, cudaMemcpyDeviceToHost, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the bucket cursors failed"));
                    void *regs = nullptr;
                    uint32_t nregs = 0;
                    step(b->mg_sketch(&regs, &nregs));
                    if (err == KTG_OK)
                        step(cudaMemcpyAsync(sketches[i].data(), regs, HLL_M * 4, cudaMemcpyDeviceToHost, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the sketch failed"));
                    step(b->sync_stream()); // my writes into the peers' buckets have landed
                }
                cursors[i].assign(cur_host.begin(), cur_host.end());
                // -- every writer is done; cursors and sketches are published
                if (!gate.wait(err == KTG_OK)) return err != KTG_OK ? err : PEER_FAILED;
                std::vector<unsigned long long> ends((size_t)n * n_sub);
                uint64_t n_keys = 0;
                for (uint32_t s = 0; s < n; ++s)
                    for (uint32_t q = 0; q < n_sub; ++q) {
                        const unsigned long long lo = ((unsigned long long)i * n_sub + q) * cap[i]; // the sender's virtual position
                        const unsigned long long v = cursors[s][(size_t)i * n_sub + q];
                        unsigned long long fill = v > lo ? v - lo : 0;
                        if (fill > cap[i]) fill = cap[i];
                        ends[(size_t)s * n_sub + q] = ((unsigned long long)s * n_sub + q) * cap[i] + fill;
                        n_keys += fill;
                    }
                {
                    std::vector<uint32_t> merged(sketches[0]);
                    for (uint32_t s = 1; s < n; ++s)
                        for (uint32_t q = 0; q < HLL_M; ++q) merged[q] = std::max(merged[q], sketches[s][q]);
                    step(cudaMemcpyAsync(st.sketch.p, merged.data(), HLL_M * 4, cudaMemcpyHostToDevice, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the merged sketch failed"));
                    if (err == KTG_OK) step(b->sync_stream()); // `merged` is a local
                    if (err == KTG_OK) step(b->mg_merge_sketch(st.sketch.p));
                }
                if (err == KTG_OK) {
                    step(st.ends.ensure(ends.size() * 8));
                    if (err == KTG_OK)
                        step(cudaMemcpyAsync(st.ends.p, ends.data(), ends.size() * 8, cudaMemcpyHostToDevice, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the bucket ends failed"));
                    if (err == KTG_OK) step(b->sync_stream());
                    if (err == KTG_OK) step(b->mgd_insert(st.ends.p, n_keys, 0));
                }
                rounds = 0; // (the rounds of the other exchanges below are skipped)
            }
            // -- receive buffers (one bucket per source shard, two slots)
            if (rounds) {
                void *base = nullptr;
                uint64_t bytes = 0, bc = 0;
                uint32_t ns = 0;
                step(skm ? b->mg_skm_prepare(gmax, &base, &bytes, &bc) : b->mg_prepare(gmax, &base, &bytes, &bc, &ns));
                rx[i] = base;
                cap[i] = bc;
            }
            if (rounds && !gate.wait(err == KTG_OK)) return err != KTG_OK ? err : PEER_FAILED;
            Plan &p = plan[i];
            auto issue_copy = [&](uint32_t c) -> int {
                if (c >= p.hint.size()) return KTG_OK;
                const int s = (int)(c & 1);
                const uint64_t r = p.c[c], r1 = p.c[c + 1], nb = offsets[r1] - offsets[r], nr = r1 - r;
                if (st.used[s]) KTG_CUDA(cudaStreamWaitEvent(b->copy_stream, st.consumed[s], 0));
                if (st.bases[s].cap < nb + 64 || st.offs[s].cap < (nr + 1) * 8) {
                    KTG_TRY(b->sync_stream());
                    KTG_CUDA(cudaStreamSynchronize(b->copy_stream));
                    KTG_TRY(st.bases[s].ensure(std::max<uint64_t>(nb, CHUNK) + 64));
                    KTG_TRY(st.offs[s].ensure((nr + 1) * 8));
                }
                KTG_CUDA(cudaMemcpyAsync(st.bases[s].p, bases + offsets[r], nb, cudaMemcpyHostToDevice, b->copy_stream));
                if (hint_of(p, c).ulen) {
                    const int g = (int)std::min<uint64_t>((nr + 1 + 255) / 256, 4096);
                    fill_offsets_kernel<<<g, 256, 0, b->copy_stream>>>((uint64_t *)st.offs[s].p, nr + 1, offsets[r], hint_of(p, c).ulen);
                    KTG_CUDA(cudaGetLastError());
                }
                else KTG_CUDA(cudaMemcpyAsync(st.offs[s].p, offsets + r, (nr + 1) * 8, cudaMemcpyHostToDevice, b->copy_stream));
                KTG_CUDA(cudaEventRecord(st.ready[s], b->copy_stream));
                st.used[s] = true;
                return KTG_OK;
            };
            if (rounds) step(issue_copy(0));
            std::vector<unsigned long long> my_ends(n);
            for (uint32_t c = 0; c < rounds; ++c) {
                const uint32_t slot = c & 1;
                trace("mh round>", i * 1000 + c);
                step(issue_copy(c + 1)); // the next chunk's copy runs under this round
                // -- sender: pack + extract + scatter into the owners' buckets
                void *d_cur = nullptr, *d_kc = nullptr;
                if (err == KTG_OK) {
                    const bool mine = c < p.hint.size();
                    const uint64_t r = mine ? p.c[c] : 0, r1 = mine ? p.c[c + 1] : 0;
                    const uint64_t nb = mine ? offsets[r1] - offsets[r] : 0, nr = r1 - r;
                    const int s = (int)(c & 1);
                    const uint8_t *d_bases = nullptr;
                    const uint64_t *d_offs = nullptr;
                    if (mine) {
                        step(cudaStreamWaitEvent(b->stream, st.ready[s], 0) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "cudaStreamWaitEvent failed"));
                        d_bases = (const uint8_t *)st.bases[s].p - offsets[r]; // offsets stay absolute
                        d_offs = (const uint64_t *)st.offs[s].p;
                        b->input_consumed = st.consumed[s];
                    }
                    const BatchHint *hint = mine ? &hint_of(p, c) : nullptr; // no device round trip to learn the window count
                    if (mine) b->hint_shift0 = (uint32_t)((uintptr_t)st.bases[s].p & 31);
                    if (skm) step(b->mg_skm_scatter_reads(d_bases, d_offs, nr, nb, rx.data(), slot, c == 0, nullptr, &d_cur, &d_kc, hint));
                    else step(b->mg_scatter_reads(d_bases, d_offs, nr, nb, rx.data(), slot, c == 0, nullptr, &d_cur, hint));
                    b->input_consumed = nullptr;
                    if (mine && nr == 0) KTG_CUDA(cudaEventRecord(st.consumed[s], b->stream));
                }
                if (err == KTG_OK) {
                    step(cudaMemcpyAsync(cursors[i].data(), d_cur, n * 8, cudaMemcpyDeviceToHost, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the bucket cursors failed"));
                    if (skm) step(cudaMemcpyAsync(key_counts[i].data(), d_kc, (n + 1) * 8, cudaMemcpyDeviceToHost, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the key counts failed"));
                    else {
                        void *regs = nullptr;
                        uint32_t nregs = 0;
                        step(b->mg_sketch(&regs, &nregs));
                        if (err == KTG_OK)
                            step(cudaMemcpyAsync(sketches[i].data(), regs, HLL_M * 4, cudaMemcpyDeviceToHost, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the sketch failed"));
                    }
                    step(b->sync_stream()); // my writes into the peers' buckets have landed
                }
                trace("mh scattered", i * 1000 + c);
                // -- every writer is done; cursors (and sketches) are published
                if (!gate.wait(err == KTG_OK)) return err != KTG_OK ? err : PEER_FAILED;
                trace("mh barrier<", i * 1000 + c);
                // -- owner: what the shards wrote into my buckets
                uint64_t n_keys = 0;
                for (uint32_t s = 0; s < n; ++s) {
                    const unsigned long long lo = (unsigned long long)i * cap[s]; // the sender's virtual position: owner * its cap + fill
                    unsigned long long fill = cursors[s][i] > lo ? cursors[s][i] - lo : 0;
                    // every shard sized its buckets from the same gmax, so cap[s] == cap[i]
                    if (fill > cap[i]) fill = cap[i];
                    my_ends[s] = (unsigned long long)s * cap[i] + fill;
                    n_keys += skm ? key_counts[s][i] : fill;
                }
                if (!skm) { // all-reduce (max) of the sketches, every shard sizes itself from it
                    std::vector<uint32_t> merged(sketches[0]);
                    for (uint32_t s = 1; s < n; ++s)
                        for (uint32_t q = 0; q < HLL_M; ++q) merged[q] = std::max(merged[q], sketches[s][q]);
                    step(cudaMemcpyAsync(st.sketch.p, merged.data(), HLL_M * 4, cudaMemcpyHostToDevice, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the merged sketch failed"));
                    if (err == KTG_OK) step(b->sync_stream()); // `merged` is a local
                    if (err == KTG_OK) step(b->mg_merge_sketch(st.sketch.p));
                }
                if (err == KTG_OK) {
                    step(cudaMemcpyAsync(st.ends.p, my_ends.data(), n * 8, cudaMemcpyHostToDevice, b->stream) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying the bucket ends failed"));
                    if (err == KTG_OK) step(b->sync_stream());
                    if (err == KTG_OK) step(skm ? b->mg_skm_insert_buckets(st.ends.p, n_keys, slot) : b->mg_insert_buckets(st.ends.p, n_keys, slot));
                }
                trace("mh inserted", i * 1000 + c);
                // (the insert ends with a stream synchronisation: when the next barrier is passed, every
                // shard has consumed this slot, which is written again two rounds from now)
            }
            // -- keys / records that did not fit a receive bucket (skew): routed through the host
            {
                void *d_sp = nullptr;
                uint64_t nsp = 0;
                if (err == KTG_OK) step(skm ? b->mg_skm_spill(&d_sp, &nsp) : b->mg_spill(&d_sp, &nsp));
                n_spill[i] = nsp;
                if (!gate.wait(err == KTG_OK)) return err != KTG_OK ? err : PEER_FAILED;
                uint64_t any = 0;
                for (uint32_t s = 0; s < n; ++s) any += n_spill[s];
                if (any) {
                    const uint32_t words = skm ? 2u : (kk <= 32 ? 1u : 2u);
                    if (nsp) {
                        void *d_grouped = nullptr;
                        std::vector<uint64_t> counts(n, 0);
                        step(skm ? b->mg_skm_partition_records(d_sp, nsp, &d_grouped, counts.data())
                                 : b->partition_keys(d_sp, nsp, &d_grouped, counts.data()));
                        uint64_t off = 0;
                        for (uint32_t d = 0; d < n && err == KTG_OK; ++d) {
                            spill_out[i][d].resize(counts[d] * words);
                            if (counts[d])
                                step(cudaMemcpy(spill_out[i][d].data(), (const uint64_t *)d_grouped + off * words, counts[d] * words * 8,
                                                cudaMemcpyDeviceToHost) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying spilled keys failed"));
                            off += counts[d];
                        }
                    }
                    if (!gate.wait(err == KTG_OK)) return err != KTG_OK ? err : PEER_FAILED;
                    std::vector<uint64_t> mine;
                    for (uint32_t s = 0; s < n; ++s) mine.insert(mine.end(), spill_out[s][i].begin(), spill_out[s][i].end());
                    if (!mine.empty()) {
                        DeviceBuf tmp;
                        step(tmp.ensure(mine.size() * 8));
                        if (err == KTG_OK)
                            step(cudaMemcpy(tmp.p, mine.data(), mine.size() * 8, cudaMemcpyHostToDevice) == cudaSuccess ? KTG_OK : fail(KTG_ERR_CUDA, "copying spilled keys failed"));
                        if (err == KTG_OK) step(skm ? b->mg_skm_insert_records(tmp.p, mine.size() / words) : b->mg_insert_spill(tmp.p, mine.size() / words));
                        if (err == KTG_OK) step(b->sync_stream());
                        tmp.release();
                    }
                    if (!gate.wait(err == KTG_OK)) return err != KTG_OK ? err : PEER_FAILED;
                }
            }
            return err;
        });
        if (rc_ != KTG_OK) return poison(rc_);
        // the caller may reuse its buffers when this returns
        KTG_TRY(parallel([&](uint32_t i) -> int {
            KTG_CUDA(cudaStreamSynchronize(sh[i]->copy_stream));
            return KTG_OK;
        }));
        if (acc_reads || acc_bytes) {
            uint64_t r1c = 0, b1c = 0;
            KTG_TRY(read_counters(&r1c, &b1c));
            if (acc_reads) *acc_reads += r1c - r0c;
            if (acc_bytes) *acc_bytes += b1c - b0c;
        }
        return KTG_OK;
    }

    int read_counters(uint64_t *reads, uint64_t *bytes) {
        std::vector<uint64_t> r(n, 0), by(n, 0);
        int rc_ = parallel([&](uint32_t i) -> int { return sh[i]->read_counters(&r[i], &by[i]); });
        if (rc_ != KTG_OK) return poison(rc_); // "Read is too short!" voids the whole build (hm_gir.rs:40)
        uint64_t sr = 0, sb = 0;
        for (uint32_t i = 0; i < n; ++i) {
            sr += r[i];
            sb += by[i];
        }
        if (reads) *reads = sr;
        if (bytes) *bytes = sb;
        return KTG_OK;
    }

    int finalize() {
        KTG_TRY(check_state());
        int rc_ = parallel([&](uint32_t i) -> int { return sh[i]->finalize(); });
        return rc_ == KTG_ERR_SHORT_READ || rc_ == KTG_ERR_TABLE_FULL ? poison(rc_) : rc_;
    }
    int reset() {
        deferred_error = KTG_OK;
        deferred_msg.clear();
        exchanged_bytes = 0;
        touch();
        return parallel([&](uint32_t i) -> int { return sh[i]->reset(); });
    }
    int set_option(const char *name, int64_t value, int (*apply)(BuilderBase *, const char *, int64_t)) {
        for (uint32_t i = 0; i < n; ++i) KTG_TRY(apply(sh[i].get(), name, value));
        return KTG_OK;
    }

    // ---- whole-graph queries ---------------------------------------------------------------------
    int edge_stats(uint32_t threshold, EdgeStats *out) {
        KTG_TRY(check_state());
        std::vector<EdgeStats> es(n);
        KTG_TRY(parallel([&](uint32_t i) -> int { return sh[i]->edge_stats(threshold, &es[i]); }));
        EdgeStats t{};
        for (const EdgeStats &e : es) { // disjoint shards: sums wrap mod 2^64, the maximum is a maximum
            t.edges += e.edges;
            t.sum_w += e.sum_w;
            t.sum_w_below += e.sum_w_below;
            t.digest += e.digest;
            t.max_w = std::max(t.max_w, e.max_w);
        }
        *out = t;
        return KTG_OK;
    }

    // A node's edges may live on several shards: every shard exports the (canonical (k-1)-mer, degree
    // word) pairs of its edges, the first device gathers and merges them (stats/collections.rs:137-208)
    int node_stats(NodeStats *out) {
        KTG_TRY(check_state());
        std::vector<void *> pk(n, nullptr), pd(n, nullptr);
        std::vector<uint64_t> cnt(n, 0);
        std::vector<uint32_t> kw(n, 1);
        KTG_TRY(parallel([&](uint32_t i) -> int { return sh[i]->nodes_export(&pk[i], &pd[i], &cnt[i], &kw[i]); }));
        uint64_t total = 0;
        for (uint64_t c : cnt) total += c;
        KTG_CUDA(cudaSetDevice(dev[0]));
        DeviceBuf keys, deg;
        const size_t kb = (size_t)kw[0] * 8;
        int rc_ = keys.ensure(total * kb + 64);
        if (rc_ == KTG_OK) rc_ = deg.ensure(total * 4 + 64);
        uint64_t off = 0;
        // on the first shard's stream, which the merge kernels follow (a device-to-device cudaMemcpyPeer
        // does not wait on the host side, and that stream does not synchronise with the default stream)
        for (uint32_t i = 0; i < n && rc_ == KTG_OK; ++i) {
            if (cnt[i]) {
                if (cudaMemcpyPeerAsync((char *)keys.p + off * kb, dev[0], pk[i], dev[i], cnt[i] * kb, sh[0]->stream) != cudaSuccess ||
                    cudaMemcpyPeerAsync((char *)deg.p + off * 4, dev[0], pd[i], dev[i], cnt[i] * 4, sh[0]->stream) != cudaSuccess)
                    rc_ = fail(KTG_ERR_CUDA, "gathering the shards' nodes failed: %s", cudaGetErrorString(cudaGetLastError()));
            }
            off += cnt[i];
        }
        if (rc_ == KTG_OK) rc_ = sh[0]->nodes_stats_from(keys.p, deg.p, total, out);
        keys.release();
        deg.release();
        return rc_;
    }

    int remove_weak_edges(uint32_t t) {
        KTG_TRY(check_state());
        touch();
        return parallel([&](uint32_t i) -> int { return sh[i]->remove_weak_edges(t); });
    }

    // standardize_edges (standardizer.rs:42-70,123-127): the two sums over all shards, one ratio, every shard scaled
    int standardize(uint64_t G, uint64_t k_, uint32_t t) {
        EdgeStats es;
        KTG_TRY(edge_stats(t, &es));
        if (G < k_ || es.sum_w == es.sum_w_below)
            return fail(KTG_ERR_DEGENERATE, "degenerate standardization ratio (G=%llu k=%llu s=%llu l=%llu)",
                        (unsigned long long)G, (unsigned long long)k_, es.sum_w, es.sum_w_below);
        const double p = (double)(G - k_) / (double)(es.sum_w - es.sum_w_below);
        touch();
        return parallel([&](uint32_t i) -> int { return sh[i]->scale_weights(p, t); });
    }

    // Every shard's edges, both strands expanded, gathered in arrays on the first device: the shards'
    // compaction kernels write straight into them (peer memory).  The arrays belong to sc.
    int gather_edges(Scratch &sc, uint64_t **d_hi, uint64_t **d_lo, uint32_t **d_w, uint64_t *ne_out) {
        KTG_TRY(check_state());
        std::vector<EdgeStats> es(n);
        KTG_TRY(parallel([&](uint32_t i) -> int { return sh[i]->edge_stats(0, &es[i]); }));
        uint64_t ne = 0;
        std::vector<uint64_t> off(n, 0);
        for (uint32_t i = 0; i < n; ++i) {
            off[i] = ne;
            ne += es[i].edges;
        }
        KTG_CUDA(cudaSetDevice(dev[0]));
        *d_hi = nullptr;
        KTG_TRY(sc.alloc(d_lo, ne));
        KTG_TRY(sc.alloc(d_w, ne));
        if (k > 32) KTG_TRY(sc.alloc(d_hi, ne));
        KTG_TRY(parallel([&](uint32_t i) -> int {
            KTG_TRY(sh[i]->compact_edges_into(*d_hi ? *d_hi + off[i] : nullptr, *d_lo + off[i], *d_w + off[i], es[i].edges));
            return sh[i]->sync_stream();
        }));
        *ne_out = ne;
        return KTG_OK;
    }

    int export_edges(uint64_t *hi, uint64_t *lo, uint32_t *w, uint64_t cap, int sorted, uint64_t *n_out) {
        if (!lo || !w || cap == 0) { // size query
            EdgeStats es;
            KTG_TRY(edge_stats(0, &es));
            if (n_out) *n_out = es.edges;
            return KTG_OK;
        }
        Scratch sc;
        uint64_t *d_hi, *d_lo, ne = 0;
        uint32_t *d_w;
        KTG_TRY(gather_edges(sc, &d_hi, &d_lo, &d_w, &ne));
        if (n_out) *n_out = ne;
        KTG_CUDA(cudaSetDevice(dev[0]));
        return sh[0]->edges_to_host(d_hi, d_lo, d_w, ne, sorted, hi, lo, w, cap);
    }

    // the device graph of the whole table lives on the first device (its builder's cache)
    int graph_prepare(uint64_t *n_nodes, uint64_t *n_edges) {
        KTG_TRY(check_state());
        if (!sh[0]->graph_ready()) {
            Scratch sc;
            uint64_t *d_hi, *d_lo, ne = 0;
            uint32_t *d_w;
            KTG_TRY(gather_edges(sc, &d_hi, &d_lo, &d_w, &ne));
            KTG_CUDA(cudaSetDevice(dev[0]));
            KTG_TRY(sh[0]->graph_build(sc, d_hi, d_lo, d_w, ne));
        }
        KTG_CUDA(cudaSetDevice(dev[0]));
        return sh[0]->graph_prepare(n_nodes, n_edges);
    }
    int export_graph(uint64_t *node_hi, uint64_t *node_lo, uint64_t n_nodes, uint64_t *src, uint64_t *dst, uint32_t *weight,
                     uint8_t *edge_bytes, uint64_t n_edges) {
        KTG_TRY(graph_prepare(nullptr, nullptr));
        return sh[0]->graph_to_host(node_hi, node_lo, n_nodes, src, dst, weight, edge_bytes, n_edges);
    }
    int export_externals(uint64_t *ids, uint8_t *kinds, uint64_t cap, uint64_t *n_out) {
        KTG_TRY(graph_prepare(nullptr, nullptr));
        return sh[0]->externals_to_host(ids, kinds, cap, n_out);
    }
    void touch() { // any shard changed: the gathered graph is stale
        if (!sh.empty() && sh[0]) sh[0]->touch();
    }

    int info(ktg_info *out) {
        memset(out, 0, sizeof *out);
        std::vector<ktg_info> inf(n);
        KTG_TRY(parallel([&](uint32_t i) -> int { return sh[i]->info(&inf[i]); }));
        for (const ktg_info &x : inf) {
            out->capacity_slots += x.capacity_slots;
            out->occupied_slots += x.occupied_slots;
            out->table_bytes += x.table_bytes;
            out->n_sub_tables += x.n_sub_tables;
            out->windows_inserted += x.windows_inserted;
            out->kernel_launches += x.kernel_launches;
            out->grow_events += x.grow_events;
            out->page_updates += x.page_updates;
            out->n_pages += x.n_pages;
        }
        out->slot_bytes = inf[0].slot_bytes;
        out->partitioned = 1;
        return KTG_OK;
    }
};

} // namespace ktg
