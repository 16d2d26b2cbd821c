"""Hash-sharded build across GPUs: one process per GPU, `torch.distributed` for the plumbing.

The table is sharded by owner(key) = hash(canonical k-mer) range-reduced to [0, world)
(SURVEY 8e).  Every rank extracts the keys of its own reads grouped by owner
(`ktg_partition_reads_device`), the groups are routed with one all-to-all over
NVLink, and every rank inserts what it received into its shard
(`ktg_insert_keys_device`).  Shards are disjoint by construction, so the merged GIR
is their concatenation and all whole-graph statistics are plain reductions.

`exchange_keys` is device agnostic on purpose: with the gloo backend and CPU tensors
it runs the same routing logic in the world_size-2 CPU tests.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .gir import DeviceArray, GpuGIR

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
                 **kw):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        self.gir = GpuGIR(k, reverse_complement, edges_count=edges_count, world_size=self.world, rank=self.rank,
                          stream=torch.cuda.current_stream().cuda_stream, **kw)
        self.words = self.gir.key_words()
        self.exchanged_bytes = 0

    def add_reads_device(self, d_bases, d_offsets, n_reads: int, total_bases: int):
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
        self.gir.close()
