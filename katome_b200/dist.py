"""Hash-sharded build across GPUs: one process per GPU, `torch.distributed` for the plumbing.

The table is sharded by owner(key) = hash(canonical k-mer) range-reduced to [0, world)
(SURVEY 8e).  Shards are disjoint by construction, so the merged GIR is their
concatenation and all whole-graph statistics are plain reductions.

Two data paths:
  fused (default on one node, world <= 8): the extraction kernel of every rank writes its
      keys, grouped by owner, straight into the owners' HBM over NVLink (CUDA IPC mapped peer
      memory); the owner partitions what it received by sub-table and carries on as on one
      GPU.  NCCL only carries the small control messages (batch size, sketch, bucket fills)
      and doubles as the ordering between ranks.
  nccl: every rank groups its keys by owner (`ktg_partition_reads_device`), one all-to-all
      routes them, every rank inserts what it received (`ktg_insert_keys_device`).

`exchange_keys` is device agnostic on purpose: with the gloo backend and CPU tensors
it runs the same routing logic in the world_size-2 CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

import os

from .gir import DeviceArray, GpuGIR, ipc_close, ipc_get_handle, ipc_open

MASK64 = (1 << 64) - 1


def exchange_counts(send_counts: Sequence[int], device, group=None) -> List[int]:
    """all-to-all of the per-destination key counts -> per-source counts"""
    world = dist.get_world_size(group)
    assert len(send_counts) == world
    s = torch.tensor(list(send_counts), dtype=torch.int64, device=device)
    r = torch.empty(world, dtype=torch.int64, device=device)
    dist.all_to_all_single(r, s, group=group)
    return [int(x) for x in r.tolist()]


def exchange_keys(keys: torch.Tensor, send_counts: Sequence[int], words: int = 1, group=None,
                  recv_counts: Optional[Sequence[int]] = None) -> Tuple[torch.Tensor, List[int]]:
    """Route owner-major keys (int64 words, `words` per key) to their owners.

    keys[: send_counts[0]*words] goes to rank 0, the next send_counts[1]*words to rank 1, ...
    Returns (received keys, counts per source rank)."""
    if recv_counts is None:
        recv_counts = exchange_counts(send_counts, keys.device, group)
    out = torch.empty(sum(recv_counts) * words, dtype=keys.dtype, device=keys.device)
    dist.all_to_all_single(out, keys, output_split_sizes=[c * words for c in recv_counts],
                           input_split_sizes=[c * words for c in send_counts], group=group)
    return out, list(recv_counts)


def merge_digests(local: Tuple[int, int, int, int], device, group=None) -> Tuple[int, int, int, int]:
    """Digest of the union of disjoint shards: sums wrap mod 2^64, max weight is a max."""
    def to_i64(x):
        return x - (1 << 64) if x >= (1 << 63) else x
    sums = torch.tensor([to_i64(local[0]), to_i64(local[1]), to_i64(local[2])], dtype=torch.int64, device=device)
    mx = torch.tensor([local[3]], dtype=torch.int64, device=device)
    dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    d, e, s = (int(v) & MASK64 for v in sums.tolist())
    return d, e, s, int(mx.item())


class ShardedGIR:
    """One shard of a GIR that is hash-partitioned over the ranks of a process group."""

    def __init__(self, k: int = 40, reverse_complement: bool = True, *, group=None, edges_count: Optional[int] = None,
                 fused: Optional[bool] = None, **kw):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        if fused is None:
            fused = self.world <= 8 and os.environ.get("KTG_EXCHANGE", "fused") != "nccl"
        self.fused = bool(fused)
        if self.fused:
            kw.setdefault("force_partition", True)  # the fused path has no unpartitioned mode
        self.gir = GpuGIR(k, reverse_complement, edges_count=edges_count, world_size=self.world, rank=self.rank,
                          stream=torch.cuda.current_stream().cuda_stream, **kw)
        self.k = int(k)
        self.words = self.gir.key_words()
        self.exchanged_bytes = 0
        self._peers: List[int] = []   # receive buffer of every rank, mapped here (own rank: own pointer)
        self._cap = self._n_sub = 0

    # ---- fused path ------------------------------------------------------------------------
    def _unmap_peers(self):
        for r, p in enumerate(self._peers):
            if r != self.rank and p:
                ipc_close(p)
        self._peers = []

    def _map_peers(self, gmax: int):
        """(re)allocate the receive buffer for batches of up to gmax windows per rank and map
        everybody's buffer; collective."""
        self._unmap_peers()
        dist.barrier(self.group)  # nobody still has the old buffer mapped when it is freed
        base, nbytes, self._cap, self._n_sub = self.gir.mg_prepare(gmax)
        mine = torch.frombuffer(bytearray(ipc_get_handle(base)), dtype=torch.uint8).to(self.device)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=self.group)
        self._peers = [base if r == self.rank else ipc_open(bytes(allh[r].cpu().numpy().tobytes()))
                       for r in range(self.world)]

    def _add_reads_fused(self, d_bases, d_offsets, n_reads: int, total_bases: int):
        W, dev = self.world, self.device
        ub = max(int(total_bases) - int(n_reads) * (self.k - 1), 0)  # windows if every read is accepted
        t = torch.tensor([ub], dtype=torch.int64, device=dev)
        # also orders "every rank has finished reading its receive buffer" before anybody writes
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        gmax = int(t.item())
        if gmax == 0:
            return
        if not self._peers or self.gir.mg_plan(gmax):
            self._map_peers(gmax)
        cur_ptr = self.gir.mg_scatter_reads_device(d_bases, d_offsets, n_reads, total_bases, self._peers)
        sk_ptr, sk_n = self.gir.mg_sketch()
        regs = torch.as_tensor(DeviceArray(sk_ptr, sk_n, "<i4"), device=dev)
        dist.all_reduce(regs, op=dist.ReduceOp.MAX, group=self.group)  # every shard sizes itself from it
        cap = self._cap
        cur = torch.as_tensor(DeviceArray(cur_ptr, W), device=dev)
        got = torch.empty_like(cur)
        dist.all_to_all_single(got, cur, group=self.group)  # also: the writers' kernels have completed
        fill = (got - self.rank * cap).clamp_(max=cap)
        ends = torch.arange(W, dtype=torch.int64, device=dev) * cap + fill
        self.gir.mg_insert_buckets(ends, int(fill.sum().item()))
        self._keep = (ends, got)
        self.exchanged_bytes += int(ub * 8 * self.words * (W - 1) / W)
        # keys that did not fit their bucket (skew): routed the slow way
        sp_ptr, n_sp = self.gir.mg_spill()
        tot = torch.tensor([n_sp], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, group=self.group)
        if int(tot.item()):
            ptr, counts = self.gir.partition_keys_device(sp_ptr, n_sp)
            n = sum(counts)
            keys = torch.as_tensor(DeviceArray(ptr, n * self.words), device=dev) if n else \
                torch.empty(0, dtype=torch.int64, device=dev)
            recv, rcounts = exchange_keys(keys, counts, self.words, self.group)
            self.gir.mg_insert_spill(recv, sum(rcounts))
            self._keep = (ends, got, recv)

    def add_reads_device(self, d_bases, d_offsets, n_reads: int, total_bases: int):
        if self.fused:
            return self._add_reads_fused(d_bases, d_offsets, n_reads, total_bases)
        ptr, counts = self.gir.partition_reads_device(d_bases, d_offsets, n_reads, total_bases)
        n = sum(counts)
        if n:
            keys = torch.as_tensor(DeviceArray(ptr, n * self.words), device=self.device)
        else:
            keys = torch.empty(0, dtype=torch.int64, device=self.device)
        recv, rcounts = exchange_keys(keys, counts, self.words, self.group)
        self.exchanged_bytes += (n - counts[self.rank]) * 8 * self.words
        self.gir.insert_keys_device(recv, sum(rcounts))
        self._keep = recv  # keep the receive buffer alive until the insert has run

    def reset(self):
        self.gir.reset()
        self.exchanged_bytes = 0

    def finalize(self):
        self.gir.finalize()

    def remove_weak_edges(self, t: int):
        self.gir.remove_weak_edges(t)

    def digest(self):
        return merge_digests(self.gir.digest(), self.device, self.group)

    def edge_count(self) -> int:
        return self.digest()[1]

    def standardize_edges(self, genome_len: int, k: int, t: int):
        raise NotImplementedError("multi-GPU standardize_edges needs the all-reduced sums (next round)")

    def close(self):
        if self._peers:
            torch.cuda.synchronize()
            self._unmap_peers()
            dist.barrier(self.group)
        self.gir.close()
