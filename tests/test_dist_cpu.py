"""world_size-2 test of the hash-sharding host logic on CPU (gloo): ownership, routing with
uneven splits, disjoint shards, merged digest.  The GPU kernels are not involved: each rank's
keys come from the oracle, so this covers exactly the code that sits between
ktg_partition_reads_device and ktg_insert_keys_device in katome_b200/dist.py."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests import helpers as H


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _splitmix64(x):
    from katome_b200.hashing import fmix64
    with np.errstate(over="ignore"):
        return fmix64(np.asarray(x, np.uint64) + np.uint64(0x9E3779B97F4A7C15))


def _worker(rank, world, port, k, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from katome_b200 import hashing
        from katome_b200.dist import exchange_keys, merge_digests
        from oracle import oracle as O
        n, L, G = 600, 100, 20000
        reads = O.synth_reads(4242, G, L, 5000, 0, n)
        offsets = np.arange(n + 1, dtype=np.uint64) * L
        # this rank's reads -> canonical key occurrences (what the partition kernel emits)
        half = n // world
        lo_r, hi_r = rank * half, (rank + 1) * half
        seqs = [reads[i * L:(i + 1) * L].tobytes().decode() for i in range(lo_r, hi_r)]
        occ = np.array([H.kmer_int(s[i:i + k]) for s in seqs for i in range(L - k + 1)], dtype=object)
        lo = np.array([int(v) & (2**64 - 1) for v in occ], np.uint64)
        hi = np.array([int(v) >> 64 for v in occ], np.uint64)
        chi, clo = hashing.canonical(hi, lo, k)
        own = hashing.owner_of(hi, lo, k, world, True)
        order = np.argsort(own, kind="stable")
        counts = np.bincount(own, minlength=world).tolist()
        words = 1 if k <= 32 else 2
        if words == 1:
            send = clo[order].view(np.int64)
        else:
            send = np.stack([clo[order], chi[order]], axis=1).reshape(-1).view(np.int64)
        recv, rcounts = exchange_keys(torch.from_numpy(send.copy()), counts, words)
        assert recv.numel() == sum(rcounts) * words
        got = recv.numpy().view(np.uint64).reshape(-1, words)
        glo = got[:, 0]
        ghi = got[:, 1] if words == 2 else np.zeros_like(glo)
        # everything received is owned by this rank
        assert np.all(hashing.owner_of(ghi, glo, k, world, False) == rank)
        # shard = weights of the canonical keys, expanded to both strands like the device export
        keys, w = np.unique(np.stack([ghi, glo], axis=1), axis=0, return_counts=True)
        rhi, rlo = hashing.revcomp(keys[:, 0], keys[:, 1], k)
        pal = (rhi == keys[:, 0]) & (rlo == keys[:, 1])
        w = np.where(pal, 2 * w, w).astype(np.uint64)
        with np.errstate(over="ignore"):
            term = lambda a, b: _splitmix64(_splitmix64(a) ^ b) * (np.uint64(2) * w + np.uint64(1))
            d = int(np.sum(term(keys[:, 0], keys[:, 1]), dtype=np.uint64))
            d2 = int(np.sum(term(rhi, rlo)[~pal], dtype=np.uint64))
        local = ((d + d2) & (2**64 - 1), int(len(keys) + np.count_nonzero(~pal)),
                 int(np.sum(w) + np.sum(w[~pal])), int(w.max()))
        merged = merge_digests(local, torch.device("cpu"))
        cpu = O.OracleGIR(k)
        cpu.add_reads(reads, offsets, True)
        assert merged == cpu.digest(), (merged, cpu.digest())
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("k", [31, 40])
def test_two_rank_exchange_gloo(tmp_path, k):
    world, port = 2, _free_port()
    mp.spawn(_worker, args=(world, port, k, str(tmp_path)), nprocs=world, join=True)
    assert all(os.path.exists(tmp_path / f"ok{r}") for r in range(world))


def test_host_mirror_of_device_key_functions():
    from katome_b200 import hashing
    rng = np.random.default_rng(5)
    for k in (5, 31, 32, 33, 40, 63, 64):
        for _ in range(20):
            s = "".join(rng.choice(list("ACGT"), size=k))
            v, r = H.kmer_int(s), H.kmer_int(H.revcomp(s))
            hi, lo = hashing.revcomp(np.array([v >> 64], np.uint64), np.array([v & (2**64 - 1)], np.uint64), k)
            assert (int(hi[0]) << 64) | int(lo[0]) == r
            chi, clo = hashing.canonical(np.array([v >> 64], np.uint64), np.array([v & (2**64 - 1)], np.uint64), k)
            assert (int(chi[0]) << 64) | int(clo[0]) == min(v, r)
    own = hashing.owner_of(np.zeros(20000, np.uint64), rng.integers(0, 2**62, 20000, dtype=np.uint64), 31, 8)
    assert own.min() == 0 and own.max() == 7 and np.all(np.bincount(own) > 2000)
