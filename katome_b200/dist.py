"""Hash-sharded build across GPUs: one process per GPU, `torch.distributed` for the plumbing.

The table is sharded by owner(key) = hash(canonical k-mer) range-reduced to [0, world)
(SURVEY 8e).  Shards are disjoint by construction, so the merged GIR is their
concatenation and all whole-graph statistics are plain reductions.

Data paths:
  direct (opt-in, exchange="direct"; needs the same table geometry on every shard): the extraction kernel of
      every rank bins its keys by (owner, sub-table of the owner) and writes the runs straight into the owner's
      HBM over NVLink; what arrives is already the owner's level-1 stage, so the owner goes on with its level-2
      scatter and page sweep.  Measured slower than the key exchange on 2 B200s (short runs over NVLink).
  fused, super-k-mers (default on one node, world <= 8, 23 <= k <= 31): owner(k-mer) is a
      function of its minimizer; the sending kernel cuts every read into runs of windows that
      share a minimizer and writes each run as ONE 16-byte record straight into the owner's HBM
      over NVLink (2.9 bytes per window instead of 8); the owner unrolls the records into
      canonical k-mers while it partitions them by sub-table (csrc/superkmer.cuh).
  fused, keys (other k): the extraction kernel of every rank writes its
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

from .gir import DeviceArray, GpuGIR, ipc_close, ipc_get_handle, ipc_open, skm_supported

MASK64 = (1 << 64) - 1


def bind_to_gpu_numa(device_index: int) -> Optional[int]:
    """Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE it allocates pinned staging
    memory (first touch then places the pages there).  Eight feeding processes that all sit on node 0 halve
    each other's H2D rate (SCALE_r01: 55 -> 23 GB/s per GPU at N = 8).  Best effort: returns the node, or None
    when the topology cannot be read (single-node VMs, containers without sysfs)."""
    try:
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except (OSError, ValueError, AttributeError):
        return None


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
                 fused: Optional[bool] = None, exchange: Optional[str] = None, **kw):
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = torch.device("cuda", torch.cuda.current_device())
        mode = exchange or "fused"  # fused | direct | skm | keys | nccl
        if fused is None:
            fused = self.world <= 8 and mode != "nccl"
        self.fused = bool(fused)
        # Which fused exchange.  Super-k-mer records cost the owner one more pass (unrolling them)
        # and save the sender 5 of every 8 NVLink bytes: measured on B200s that loses at 2 GPUs
        # (74.7 vs 82.3 G k-mers/s on C2), ties at 4 (147.5 vs 142.7) and wins at 8 (288 vs 262), where
        # the key exchange is NVLink bound, so it
        # is the default from SKM_MIN_WORLD ranks on ("skm" / "keys" force one).
        want_skm = mode == "skm" or (mode == "fused" and self.world >= self.SKM_MIN_WORLD)
        self.exchange = "nccl" if not self.fused else ("skm" if skm_supported(k) and want_skm else "keys")
        # The direct exchange (the sender bins by owner AND sub-table with the big-tile level-1 kernel, the
        # owner goes straight to its level-2 scatter) replaces the key exchange below SKM_MIN_WORLD ranks
        # whenever every shard has the same geometry (checked per batch; "keys" forces the key exchange).
        # Measured on B200s, C3 as one job (profiles/r02e_*): N = 2 17.2 ms against 18.5 for the key exchange;
        # N = 8 6.5 ms against 5.8 for super-k-mers -- its sender is NVLink bound (8 bytes per window against
        # 2.9), so from 4 ranks on it stays opt-in ("direct").  With the one-pass sender of the first version
        # (2048-key tiles, ~100-byte runs reached ~320 GB/s over NVLink) it lost at N = 2 as well (20.4 ms).
        self.direct = self.fused and (mode == "direct" or (mode == "fused" and self.world < self.SKM_MIN_WORLD))
        if self.direct:
            self.exchange = "keys"
        self._mapped_for = None
        self.last_exchange = self.exchange
        if self.fused:
            kw.setdefault("force_partition", True)  # the fused path has no unpartitioned mode
        self.gir = GpuGIR(k, reverse_complement, edges_count=edges_count, world_size=self.world, rank=self.rank,
                          stream=torch.cuda.current_stream().cuda_stream, **kw)
        self.k = int(k)
        self.words = self.gir.key_words()
        self.exchanged_bytes = 0
        self._peers: List[int] = []   # receive buffer of every rank, mapped here (own rank: own pointer)
        self._cap = self._slot_bytes = 0
        self._send_stream = None
        self._send_group = None
        self._copy_stream = None
        self._stage = None

    # ---- fused path ------------------------------------------------------------------------
    # A batch can be sent in several chunks (two receive slots, sender on its own stream) so that the
    # owner-side work of chunk c overlaps the exchange of chunk c+1.  Measured on 2 B200s it does not
    # pay (the extra table sweeps and collectives cost more than the overlap gains), so the default
    # is one chunk (the class attribute CHUNKS overrides it).
    CHUNKS = 1
    MIN_CHUNK_READS = 1 << 16
    SKM_MIN_WORLD = 4

    def _unmap_peers(self):
        for r, p in enumerate(self._peers):
            if r != self.rank and p:
                ipc_close(p)
        self._peers = []

    def _map_peers(self, gmax: int):
        """(re)allocate the receive buffer (two slots) for chunks of up to gmax windows per rank and
        map everybody's buffer; collective."""
        self._unmap_peers()
        dist.barrier(self.group)  # nobody still has the old buffer mapped when it is freed
        if self.exchange == "skm":
            base, self._slot_bytes, self._cap = self.gir.mg_skm_prepare(gmax)
        else:
            base, self._slot_bytes, self._cap, _ = self.gir.mg_prepare(gmax)
        self._mapped_for = self.exchange
        mine = torch.frombuffer(bytearray(ipc_get_handle(base)), dtype=torch.uint8).to(self.device)
        allh = [torch.empty_like(mine) for _ in range(self.world)]
        dist.all_gather(allh, mine, group=self.group)
        self._peers = [base if r == self.rank else ipc_open(bytes(allh[r].cpu().numpy().tobytes()))
                       for r in range(self.world)]

    def _add_reads_fused(self, d_bases, d_offsets, n_reads: int, total_bases: int):
        W, dev, k = self.world, self.device, self.k
        main = torch.cuda.current_stream()
        if self._send_stream is None:
            self._send_stream = torch.cuda.Stream()
            # a second communicator: collectives of the two streams must not queue behind each other
            ranks = None if self.group is None else dist.get_process_group_ranks(self.group)
            self._send_group = dist.new_group(ranks=ranks, backend="nccl")
        send, sgroup = self._send_stream, self._send_group
        # chunks of whole reads; the offsets stay absolute (the kernels subtract offsets[chunk start])
        want = self.CHUNKS
        n_chunks = want if n_reads >= want * self.MIN_CHUNK_READS else 1
        per = -(-n_reads // n_chunks) if n_reads else 0
        bounds = [min(i * per, n_reads) for i in range(n_chunks + 1)]
        offs_host = d_offsets[bounds].tolist() if n_reads else [0] * (n_chunks + 1)
        ubs = [max((offs_host[i + 1] - offs_host[i]) - (bounds[i + 1] - bounds[i]) * (k - 1), 0)
               for i in range(n_chunks)]  # windows of a chunk if every read is accepted
        t = torch.tensor([max(ubs), n_chunks], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        gmax, n_chunks_all = (int(x) for x in t.tolist())
        if gmax == 0:
            return
        self.last_exchange = "keys"
        if not self._peers or self._mapped_for != "keys" or self.gir.mg_plan(gmax):
            self._map_peers(gmax)
        cap = self._cap
        send.wait_stream(main)  # the inputs are ready; the previous batch has been consumed
        sk_ptr, sk_n = self.gir.mg_sketch()
        regs = torch.as_tensor(DeviceArray(sk_ptr, sk_n, "<i4"), device=dev)
        slot_free = [None, None]   # event: every rank has finished reading that slot
        sent = [None] * n_chunks_all
        keep = []

        def send_chunk(c):
            slot = c % 2
            if c < n_chunks and bounds[c + 1] > bounds[c]:
                lo, hi, nbases = bounds[c], bounds[c + 1], offs_host[c + 1] - offs_host[c]
            else:  # other ranks have more chunks than this one: take part with an empty chunk
                lo, hi, nbases = 0, 0, 0
            with torch.cuda.stream(send):
                if slot_free[slot] is not None:
                    send.wait_event(slot_free[slot])
                cur_ptr = self.gir.mg_scatter_reads_device(d_bases, d_offsets[lo:hi + 1], hi - lo, nbases,
                                                           self._peers, slot, c == 0, send.cuda_stream)
                cur = torch.as_tensor(DeviceArray(cur_ptr, W), device=dev)
                got = torch.empty_like(cur)
                dist.all_to_all_single(got, cur, group=sgroup)  # also: the writers' kernels have completed
                ev = torch.cuda.Event()
                ev.record(send)
            sent[c] = (got, ev, slot)

        def receive_chunk(c):
            got, ev, slot = sent[c]
            main.wait_event(ev)
            tmp = regs.clone()
            dist.all_reduce(tmp, op=dist.ReduceOp.MAX, group=self.group)  # every shard sizes itself from it
            self.gir.mg_merge_sketch(tmp)
            fill = (got - self.rank * cap).clamp_(max=cap)
            ends = torch.arange(W, dtype=torch.int64, device=dev) * cap + fill
            self.gir.mg_insert_buckets(ends, int(fill.sum().item()), slot)
            token = torch.zeros(1, dtype=torch.int32, device=dev)
            dist.all_reduce(token, group=self.group)  # every rank is done with this slot
            done = torch.cuda.Event()
            done.record(main)
            slot_free[slot] = done
            keep.append((tmp, ends, got, token))

        for c in range(n_chunks_all):
            send_chunk(c)
            if c > 0:
                receive_chunk(c - 1)
        receive_chunk(n_chunks_all - 1)
        main.wait_stream(send)
        self._keep = keep
        self.exchanged_bytes += int(sum(ubs) * 8 * self.words * (W - 1) / W)
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
            self._keep = (keep, recv)

    # The direct exchange in three steps, so that a host batch can be scattered chunk by chunk (every chunk's
    # H2D copy under the previous chunk's kernels) and inserted once:
    def _direct_begin(self, windows_ub: int) -> bool:
        """collective; False (nothing done) when the shards' geometries differ: the caller takes the key exchange"""
        W, dev = self.world, self.device
        _, n_sub, sub_log2 = self.gir.mg_direct_plan(max(windows_ub, 1))
        geo = (n_sub << 8) | sub_log2
        t = torch.tensor([windows_ub, geo, -geo], dtype=torch.int64, device=dev)
        # also orders "every rank has finished reading its receive buffer" before anybody writes again
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        gmax, gmx, gmn = (int(x) for x in t.tolist())
        self._d_n_sub, self._d_any = n_sub, gmax > 0
        if gmax == 0:
            return True
        if gmx != -gmn or W * n_sub > 1024:
            return False
        if not self._peers or self._mapped_for != ("direct", n_sub) or self.gir.mg_direct_plan(gmax)[0]:
            self._unmap_peers()
            dist.barrier(self.group)
            base, self._slot_bytes, self._cap = self.gir.mg_direct_prepare(gmax)
            self._mapped_for = ("direct", n_sub)
            mine = torch.frombuffer(bytearray(ipc_get_handle(base)), dtype=torch.uint8).to(dev)
            allh = [torch.empty_like(mine) for _ in range(W)]
            dist.all_gather(allh, mine, group=self.group)
            self._peers = [base if r == self.rank else ipc_open(bytes(allh[r].cpu().numpy().tobytes())) for r in range(W)]
        self._d_cur = None
        self._d_ub = 0
        self.last_exchange = "direct"
        return True

    def _direct_scatter(self, d_bases, d_offsets, n_reads: int, total_bases: int, first: bool):
        if not self._d_any:
            return
        self._d_cur = self.gir.mg_direct_scatter_reads_device(d_bases, d_offsets, n_reads, total_bases, self._peers, 0, first)
        self._d_ub += max(int(total_bases) - n_reads * (self.k - 1), 0)

    def _direct_finish(self):
        if not self._d_any:
            return
        W, dev, n_sub, cap = self.world, self.device, self._d_n_sub, self._cap
        if self._d_cur is None:  # nothing of mine: my cursors still have to say so
            empty = torch.empty(1, dtype=torch.int64, device=dev)
            self._d_cur = self.gir.mg_direct_scatter_reads_device(0, empty, 0, 0, self._peers, 0, True)
        cur = torch.as_tensor(DeviceArray(self._d_cur, W * n_sub), device=dev)
        got = torch.empty_like(cur)
        dist.all_to_all_single(got, cur, group=self.group)  # n_sub cursors per owner; also: the writers' kernels have completed
        sk_ptr, sk_n = self.gir.mg_sketch()
        regs = torch.as_tensor(DeviceArray(sk_ptr, sk_n, "<i4"), device=dev).clone()
        dist.all_reduce(regs, op=dist.ReduceOp.MAX, group=self.group)  # every shard sizes itself from it
        self.gir.mg_merge_sketch(regs)
        # got[s * n_sub + p]: where sender s stopped in its bucket for (owner = me, sub-table p)
        lo = (self.rank * n_sub + torch.arange(n_sub, dtype=torch.int64, device=dev)).repeat(W) * cap
        fill = (got - lo).clamp_(min=0, max=cap)
        ends = torch.arange(W * n_sub, dtype=torch.int64, device=dev) * cap + fill
        self.gir.mg_direct_insert(ends, int(fill.sum().item()), 0)
        self._keep = (cur, got, regs, ends)
        self.exchanged_bytes += int(self._d_ub * 8 * self.words * (W - 1) / W)
        sp_ptr, n_sp = self.gir.mg_spill()  # keys that did not fit their bucket (skew): routed the slow way
        tot = torch.tensor([n_sp], dtype=torch.int64, device=dev)
        dist.all_reduce(tot, group=self.group)
        if int(tot.item()):
            ptr, counts = self.gir.partition_keys_device(sp_ptr, n_sp)
            n = sum(counts)
            keys = torch.as_tensor(DeviceArray(ptr, n * self.words), device=dev) if n else \
                torch.empty(0, dtype=torch.int64, device=dev)
            recv, rcounts = exchange_keys(keys, counts, self.words, self.group)
            self.gir.mg_insert_spill(recv, sum(rcounts))
            self._keep = (self._keep, recv)

    def _add_reads_direct(self, d_bases, d_offsets, n_reads: int, total_bases: int) -> bool:
        """One device-resident batch through the direct exchange; False when the geometries differ."""
        if not self._direct_begin(max(int(total_bases) - n_reads * (self.k - 1), 0)):
            return False
        self._direct_scatter(d_bases, d_offsets, n_reads, total_bases, True)
        self._direct_finish()
        return True

    def _add_reads_skm(self, d_bases, d_offsets, n_reads: int, total_bases: int):
        """One batch through the super-k-mer exchange (one chunk, everything on the current stream)."""
        W, dev, k = self.world, self.device, self.k
        ub = max(int(total_bases) - n_reads * (k - 1), 0)  # windows if every read is accepted
        t = torch.tensor([ub], dtype=torch.int64, device=dev)
        # also orders "every rank has finished reading its receive buffer" before anybody writes again
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        gmax = int(t.item())
        if gmax == 0:
            return
        if not self._peers or self._mapped_for != "skm" or self.gir.mg_skm_plan(gmax):
            self._map_peers(gmax)
        cap = self._cap
        cur_ptr, kc_ptr = self.gir.mg_skm_scatter_reads_device(d_bases, d_offsets, n_reads, total_bases, self._peers)
        cur = torch.as_tensor(DeviceArray(cur_ptr, W), device=dev)
        kc = torch.as_tensor(DeviceArray(kc_ptr, W + 1), device=dev)  # [W]: records this rank spilled
        send = torch.stack([cur, kc[:W], kc[W:].expand(W)], dim=1).contiguous()
        got = torch.empty_like(send)
        dist.all_to_all_single(got, send, group=self.group)  # also: the writers' kernels have completed
        fill = (got[:, 0] - self.rank * cap).clamp_(max=cap)
        ends = torch.arange(W, dtype=torch.int64, device=dev) * cap + fill
        n_rec, n_keys, n_spilled = (int(x) for x in torch.stack([fill.sum(), got[:, 1].sum(), got[:, 2].sum()]).tolist())
        self.gir.mg_skm_insert_buckets(ends, n_keys, 0)
        self._keep = (send, got, ends)
        self.exchanged_bytes += n_rec * 16 * (W - 1) // W
        if n_spilled:  # records that did not fit their bucket (skew), on any rank: routed the slow way
            sp_ptr, n_sp = self.gir.mg_skm_spill()
            ptr, counts = self.gir.mg_skm_partition_records(sp_ptr, n_sp)
            n = sum(counts)
            recs = torch.as_tensor(DeviceArray(ptr, n * 2), device=dev) if n else \
                torch.empty(0, dtype=torch.int64, device=dev)
            recv, rcounts = exchange_keys(recs, counts, 2, self.group)
            self.gir.mg_skm_insert_records(recv, sum(rcounts))
            self._keep = (send, got, ends, recv)

    def add_reads_host(self, h_bases: torch.Tensor, h_offsets: torch.Tensor, n_reads: int, chunks: int = 4):
        """Reads in (pinned) host memory: h_bases uint8, h_offsets int64 [n_reads + 1].  The batch is
        copied in chunks on a copy stream, chunk i+1 while chunk i is exchanged and inserted."""
        dev, main = self.device, torch.cuda.current_stream()
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream()
        copy = self._copy_stream
        offs = h_offsets[: n_reads + 1]
        if self.direct and self.fused and chunks == 4:
            chunks = 16  # chunks cost nothing in the direct exchange: small ones hide the first copy better
        n_chunks = max(1, min(chunks, n_reads // self.MIN_CHUNK_READS)) if n_reads else 1
        per = -(-n_reads // n_chunks) if n_reads else 0
        bounds = [min(i * per, n_reads) for i in range(n_chunks + 1)]
        ob = [int(offs[b]) for b in bounds]
        max_b = max(ob[i + 1] - ob[i] for i in range(n_chunks)) + 64
        max_r = max(bounds[i + 1] - bounds[i] for i in range(n_chunks)) + 1
        if self._stage is None or self._stage[0][0].numel() < max_b or self._stage[0][1].numel() < max_r:
            torch.cuda.synchronize()
            self._stage = [(torch.empty(max_b, dtype=torch.uint8, device=dev),
                            torch.empty(max_r, dtype=torch.int64, device=dev)) for _ in range(2)]
            self._stage_ev = [[torch.cuda.Event(), torch.cuda.Event()] for _ in range(2)]  # (ready, free)
            self._stage_used = [False, False]

        def issue(c):
            s = c & 1
            lo, hi = bounds[c], bounds[c + 1]
            with torch.cuda.stream(copy):
                if self._stage_used[s]:
                    copy.wait_event(self._stage_ev[s][1])
                self._stage[s][0][: ob[c + 1] - ob[c]].copy_(h_bases[ob[c]: ob[c + 1]], non_blocking=True)
                self._stage[s][1][: hi - lo + 1].copy_(offs[lo: hi + 1], non_blocking=True)
                self._stage_ev[s][0].record(copy)

        # direct exchange: the chunks are scattered as they arrive (no collective in between) and inserted once
        direct = self.direct and self.fused and self._direct_begin(max(ob[-1] - ob[0] - n_reads * (self.k - 1), 0))
        issue(0)
        for c in range(n_chunks):
            if c + 1 < n_chunks:
                issue(c + 1)
            s = c & 1
            lo, hi = bounds[c], bounds[c + 1]
            main.wait_event(self._stage_ev[s][0])
            # the offsets stay absolute: bias the base pointer instead
            args = (int(self._stage[s][0].data_ptr()) - ob[c], self._stage[s][1][: hi - lo + 1], hi - lo, ob[c + 1] - ob[c])
            if direct:
                self._direct_scatter(*args, first=c == 0)
            else:
                self.add_reads_device(*args)
            self._stage_ev[s][1].record(main)
            self._stage_used[s] = True
        if direct:
            self._direct_finish()

    def add_reads_device(self, d_bases, d_offsets, n_reads: int, total_bases: int):
        if self.exchange == "skm":
            return self._add_reads_skm(d_bases, d_offsets, n_reads, total_bases)
        if self.fused:
            if self.direct and self._add_reads_direct(d_bases, d_offsets, n_reads, total_bases):
                return None
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
        """standardize_edges (standardizer.rs:42-70,123-127) over all shards: the two sums are
        all-reduced, every rank computes the same f64 ratio and scales its shard."""
        s, l = self.gir.edge_sums(t)
        tot = torch.tensor([s, l], dtype=torch.int64, device=self.device)
        dist.all_reduce(tot, group=self.group)
        s, l = (int(x) for x in tot.tolist())
        if genome_len < k or s == l:
            from .gir import KatomeError
            from . import _lib as L
            raise KatomeError(L.KTG_ERR_DEGENERATE, f"degenerate standardization ratio (G={genome_len} k={k} s={s} l={l})")
        self.gir.scale_weights(float(genome_len - k) / float(s - l), t)

    # ---- whole-graph statistics ---------------------------------------------------------------
    @staticmethod
    def node_owner(keys: torch.Tensor, world: int) -> torch.Tensor:
        """owner rank of canonical (k-1)-mers given as int64 rows [n, words]; any deterministic
        function of the key will do (it only decides where a node's degree words are merged)"""
        def lsr(x, n):  # logical shift right on int64
            return (x >> n) & ((1 << (64 - n)) - 1)
        h = keys[:, 0] * -7046029254386353131          # 0x9E3779B97F4A7C15 as int64; wraps
        if keys.shape[1] == 2:
            h = h ^ (keys[:, 1] * -4417276706812531889)  # 0xC2B2AE3D27D4EB4F
        h = h ^ lsr(h, 29)
        h = h * -4658895280553007687                   # 0xBF58476D1CE4E5B9
        return (lsr(h, 33) * world) >> 31

    def collection_stats(self) -> dict:
        """CollectionStats of the whole (sharded) graph, stats/collections.rs:137-208.  Edges are
        disjoint over the shards; nodes are not, so every shard sends the (node, degree word) pairs
        of its edges to the node's owner, which merges them."""
        W, dev = self.world, self.device
        pk, pd, n, kw = self.gir.nodes_export_device()
        keys = torch.as_tensor(DeviceArray(pk, n * kw), device=dev).view(n, kw) if n else \
            torch.empty((0, kw), dtype=torch.int64, device=dev)
        deg = torch.as_tensor(DeviceArray(pd, n, "<i4"), device=dev) if n else torch.empty(0, dtype=torch.int32, device=dev)
        own = self.node_owner(keys, W)
        order = torch.argsort(own)
        counts = torch.bincount(own, minlength=W).tolist()
        rk, rcounts = exchange_keys(keys[order].reshape(-1).contiguous(), counts, kw, self.group)
        rd, _ = exchange_keys(deg[order].contiguous(), counts, 1, self.group, recv_counts=rcounts)
        st = self.gir.nodes_stats_from_device(rk, rd, sum(rcounts))
        D, E, S, M = self.digest()
        sums = torch.tensor([st["node_count"], st["incoming_vert_count"], st["outgoing_vert_count"]],
                            dtype=torch.int64, device=dev)
        mx = torch.tensor([st["max_in_degree"], st["max_out_degree"]], dtype=torch.int64, device=dev)
        dist.all_reduce(sums, group=self.group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=self.group)
        nodes, sources, sinks = (int(x) for x in sums.tolist())
        return {"node_count": nodes, "edge_count": E, "max_edge_weight": M, "sum_edge_weight": S,
                "max_in_degree": int(mx[0].item()), "max_out_degree": int(mx[1].item()),
                "incoming_vert_count": sources, "outgoing_vert_count": sinks,
                "avg_edge_weight": S / E if E else float("nan"),
                "avg_out_degree": E / nodes if nodes else float("nan")}

    def counts(self) -> Tuple[int, int]:
        st = self.collection_stats()
        return st["node_count"], st["edge_count"]

    def close(self):
        if self._peers:
            torch.cuda.synchronize()
            self._unmap_peers()
            dist.barrier(self.group)
        self.gir.close()
