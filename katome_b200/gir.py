"""Host-side mirror of katome's GIR interface for the B200 builder.

`GpuGIR` exposes the operations a katome GIR type implements -- Init / Build
(algorithms/builder.rs:19-55), Clean (algorithms/pruner.rs:29-34),
Standardizable::standardize_edges (algorithms/standardizer.rs:33-39),
Stats<CollectionStats> (stats/collections.rs:170-208) and the edge export that
Convert::create_from consumes (collections/girs/hm_gir.rs:156-226) -- with the
same names, argument meaning and error behaviour, by forwarding to the C ABI in
include/katome_gpu.h.  All file:line references are under
/root/reference/src/katome/.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterable, Optional, Sequence, Tuple

import numpy as np

from . import _lib as L


class KatomeError(RuntimeError):
    """A panic of the reference, surfaced as an exception with the same message."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


class ReadTooShort(KatomeError):
    """`assert!(read.len() >= K_SIZE, "Read is too short!")` (hm_gir.rs:40)."""


def _raise(code: int):
    msg = (L.lib().ktg_last_error() or b"").decode(errors="replace")
    if code == L.KTG_ERR_SHORT_READ:
        raise ReadTooShort(code, msg or "Read is too short!")
    raise KatomeError(code, msg or f"katome_gpu error {code}")


def _check(code: int):
    if code != L.KTG_OK:
        _raise(code)


def _ptr(x) -> int:
    """device / host pointer of a torch tensor, numpy array or int"""
    if x is None:
        return 0
    if isinstance(x, int):
        return x
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    if isinstance(x, np.ndarray):
        return int(x.ctypes.data)
    raise TypeError(f"cannot take a pointer of {type(x)}")


class DeviceArray:
    """Zero-copy view of `n` words at a device pointer owned by a handle, for
    `torch.as_tensor(view, device="cuda")` (CUDA array interface; int64 unless `typestr`)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<i8"):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


_FILE_TYPES = {"fastq": L.KTG_FASTQ, "fasta": L.KTG_FASTA}


def _pinned_empty(shape, dtype) -> np.ndarray:
    """numpy array over page-locked host memory (ktg_host_alloc), freed with the array"""
    import weakref
    dt = np.dtype(dtype)
    n = int(np.prod(shape)) if not isinstance(shape, int) else int(shape)
    nbytes = max(n * dt.itemsize, 1)
    p = C.c_void_p()
    _check(L.lib().ktg_host_alloc(C.byref(p), nbytes))
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dt, count=n).reshape(shape)
    weakref.finalize(buf, L.lib().ktg_host_free, p.value)
    return arr


class GpuGIR:
    """De Bruijn graph intermediate representation built on one B200.

    One handle == one collection; unlike the reference (`static mut K_SIZE`,
    prelude.rs:21-25) k is per handle, so several can coexist.
    """

    def __init__(self, k: int = 40, reverse_complement: bool = True, *, edges_count: Optional[int] = None,
                 device: int = -1, stream: Optional[int] = None, world_size: int = 1, rank: int = 0,
                 profile: bool = False, force_direct: bool = False, force_partition: bool = False,
                 force_pages: bool = False, no_pages: bool = False, sub_table_log2_bytes: int = 0,
                 options: Optional[dict] = None, device_ids: Optional[Sequence[int]] = None):
        self._L = L.lib()
        flags = (L.KTG_FLAG_PROFILE if profile else 0) | (L.KTG_FLAG_FORCE_DIRECT if force_direct else 0) | \
                (L.KTG_FLAG_FORCE_PARTITION if force_partition else 0) | \
                (L.KTG_FLAG_FORCE_PAGES if force_pages else 0) | (L.KTG_FLAG_NO_PAGES if no_pages else 0)
        # stream=None: the handle creates its own stream.  A torch stream handle of 0 means the
        # legacy default stream, which the C ABI spells cudaStreamLegacy (0x1), since NULL = "own".
        if stream is not None and int(stream) == 0:
            stream = 1
        # device_ids: ONE handle over several GPUs (hash-sharded table, fused NVLink exchange); a device
        # may be listed more than once (shards sharing a GPU)
        ids = (C.c_int32 * len(device_ids))(*[int(d) for d in device_ids]) if device_ids else None
        cfg = L.KtgConfig(L.KTG_ABI_VERSION, int(k), int(bool(reverse_complement)), int(device),
                          int(edges_count or 0), int(world_size), int(rank), C.c_void_p(stream),
                          int(sub_table_log2_bytes), flags, len(device_ids) if device_ids else 0, ids)
        h = C.c_void_p()
        self._h = None
        _check(self._L.ktg_create(C.byref(cfg), C.byref(h)))
        self._h = h
        self.k = int(k)
        self.reverse_complement = bool(reverse_complement)
        self.world_size, self.rank = int(world_size), int(rank)
        self.device_ids = [int(d) for d in device_ids] if device_ids else None
        for name, value in (options or {}).items():
            self.set_option(name, value)

    def set_option(self, name: str, value: int):
        """ktg_set_option: test / measurement options (include/katome_gpu.h lists them)"""
        _check(self._L.ktg_set_option(self._h, name.encode(), int(value)))

    # ---- Init (builder.rs:19-25) -------------------------------------------------
    @classmethod
    def init(cls, edges_count: Optional[int], nodes_count: Optional[int], ft: str = "fastq", *, k: int = 40,
             reverse_complement: bool = True, **kw) -> "GpuGIR":
        del nodes_count, ft  # nodes are implicit; the file type does not change the table
        return cls(k, reverse_complement, edges_count=edges_count, **kw)

    def close(self):
        if getattr(self, "_h", None):
            self._L.ktg_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ---- Build (builder.rs:28-55) ------------------------------------------------
    def add_read_fastaq(self, read: bytes, reverse_complement: Optional[bool] = None):
        """One read that already passed the ACGT filter (hm_gir.rs:39-87)."""
        if reverse_complement is not None and bool(reverse_complement) != self.reverse_complement:
            raise KatomeError(L.KTG_ERR_INVALID, "reverse_complement is fixed per GpuGIR handle")
        bases = np.frombuffer(bytes(read), dtype=np.uint8)
        offsets = np.array([0, len(bases)], dtype=np.uint64)
        self.add_reads(bases, offsets)
        _check(self._L.ktg_finalize(self._h))  # a short read panics immediately in the reference

    def add_read_bfc(self, read: bytes, weight: int, reverse_complement: Optional[bool] = None):
        """One BFCounter line (builder.rs:30-36, pt_graph.rs:318-329): a k-mer with its count."""
        if reverse_complement is not None and bool(reverse_complement) != self.reverse_complement:
            raise KatomeError(L.KTG_ERR_INVALID, "reverse_complement is fixed per GpuGIR handle")
        if len(read) < self.k:
            raise ReadTooShort(L.KTG_ERR_SHORT_READ, "Read is too short!")
        self.add_weighted_kmers(np.frombuffer(bytes(read), dtype=np.uint8), np.array([weight], dtype=np.uint32))

    def add_weighted_kmers(self, kmers: np.ndarray, weights: np.ndarray, minimal_weight_threshold: int = 0):
        """Batch of BFCounter lines: `kmers` is n*k ASCII bytes, `weights` n counts; lines below the
        threshold are skipped (builder.rs:106-108).  -> (accepted k-mers, accepted bytes)"""
        kmers = np.ascontiguousarray(kmers, dtype=np.uint8)
        weights = np.ascontiguousarray(weights, dtype=np.uint32)
        if kmers.size != weights.size * self.k:
            raise KatomeError(L.KTG_ERR_BAD_RECORD, f"{kmers.size} bases for {weights.size} k-mers of {self.k}")
        nk, nb = C.c_uint64(0), C.c_uint64(0)
        _check(self._L.ktg_add_weighted_kmers(self._h, kmers.ctypes.data, weights.ctypes.data, weights.size,
                                              int(minimal_weight_threshold), C.byref(nk), C.byref(nb)))
        return nk.value, nb.value

    def add_reads(self, bases: np.ndarray, offsets: np.ndarray) -> Tuple[int, int]:
        """Batch form of the create_fastq loop body (builder.rs:152-160): drops reads with a
        byte outside "ACGT", returns (accepted_reads, accepted_bytes) of this batch."""
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nr, nb = C.c_uint64(0), C.c_uint64(0)
        _check(self._L.ktg_add_reads(self._h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                     C.byref(nr), C.byref(nb)))
        return nr.value, nb.value

    def add_reads_host_ptr(self, bases_ptr: int, offsets_ptr: int, n_reads: int, want_counts: bool = False):
        """Same from raw (ideally pinned) host pointers; asynchronous unless counts are requested."""
        if want_counts:
            nr, nb = C.c_uint64(0), C.c_uint64(0)
            _check(self._L.ktg_add_reads(self._h, bases_ptr, offsets_ptr, n_reads, C.byref(nr), C.byref(nb)))
            return nr.value, nb.value
        _check(self._L.ktg_add_reads(self._h, bases_ptr, offsets_ptr, n_reads, None, None))
        return None

    def add_reads_device(self, d_bases, d_offsets, n_reads: int, total_bases: int, want_counts: bool = False):
        """Inputs resident in HBM (torch tensors or raw device pointers)."""
        if want_counts:
            nr, nb = C.c_uint64(0), C.c_uint64(0)
            _check(self._L.ktg_add_reads_device(self._h, _ptr(d_bases), _ptr(d_offsets), n_reads, total_bases,
                                                C.byref(nr), C.byref(nb)))
            return nr.value, nb.value
        _check(self._L.ktg_add_reads_device(self._h, _ptr(d_bases), _ptr(d_offsets), n_reads, total_bases, None, None))
        return None

    @classmethod
    def create(cls, input_files: Sequence[os.PathLike], ft: str = "fastq", reverse_complement: bool = True,
               minimal_weight_threshold: int = 0, *, k: int = 40, **kw) -> Tuple["GpuGIR", int]:
        """Build::create (builder.rs:42-54): returns (collection, total accepted bytes).
        `minimal_weight_threshold` only matters for BFCounter input (`ft="bfcounter"`,
        builder.rs:106-108) and is ignored for Fastq/Fasta, as in the reference."""
        bfc = ft.lower() in ("bfcounter", "bfc")
        if ft.lower() not in _FILE_TYPES and not bfc:
            raise KatomeError(L.KTG_ERR_INVALID, f"unsupported input_file_type {ft!r}")
        g = cls(k, reverse_complement, **kw)
        arr = (C.c_char_p * len(input_files))(*[os.fsencode(f) for f in input_files])
        total = C.c_uint64(0)
        try:
            if bfc:  # create_bfc (builder.rs:79-115)
                _check(g._L.ktg_create_from_bfc_files(g._h, arr, len(input_files), int(minimal_weight_threshold),
                                                      C.byref(total)))
            else:
                _check(g._L.ktg_create_from_files(g._h, arr, len(input_files), _FILE_TYPES[ft.lower()], C.byref(total)))
        except Exception:
            g.close()
            raise
        return g, total.value

    def finalize(self):
        _check(self._L.ktg_finalize(self._h))

    def reset(self):
        """Back to `Default::default()` (builder.rs:145), keeping the device allocations."""
        _check(self._L.ktg_reset(self._h))

    # ---- Stats (stats/collections.rs:170-208 and :137-168) -----------------------
    def counts(self) -> Tuple[int, int]:
        n, e = C.c_uint64(0), C.c_uint64(0)
        _check(self._L.ktg_counts(self._h, C.byref(n), C.byref(e)))
        return n.value, e.value

    def edge_count(self) -> int:
        e = C.c_uint64(0)
        _check(self._L.ktg_counts(self._h, None, C.byref(e)))
        return e.value

    def stats(self) -> dict:
        n, e = self.counts()
        return {"node_count": n, "edge_count": e}

    def collection_stats(self) -> dict:
        s = L.KtgStats()
        _check(self._L.ktg_collection_stats(self._h, C.byref(s)))
        d = {name: int(getattr(s, name)) for name, _ in L.KtgStats._fields_}
        d["avg_edge_weight"] = d["sum_edge_weight"] / d["edge_count"] if d["edge_count"] else float("nan")
        d["avg_out_degree"] = d["edge_count"] / d["node_count"] if d["node_count"] else float("nan")
        return d

    # ---- Clean / Standardizable ---------------------------------------------------
    def remove_weak_edges(self, threshold: int):
        _check(self._L.ktg_remove_weak_edges(self._h, int(threshold)))

    def remove_single_vertices(self):
        _check(self._L.ktg_remove_single_vertices(self._h))

    def standardize_edges(self, original_genome_length: int, k_size: int, threshold: int):
        _check(self._L.ktg_standardize_edges(self._h, int(original_genome_length), int(k_size), int(threshold)))

    # ---- export (what Convert::create_from consumes) -------------------------------
    def export_edges(self, sorted: bool = True):
        ne = self.edge_count()
        hi = np.zeros(ne, np.uint64)
        lo = np.zeros(ne, np.uint64)
        w = np.zeros(ne, np.uint32)
        n = C.c_uint64(0)
        _check(self._L.ktg_export_edges(self._h, hi.ctypes.data, lo.ctypes.data, w.ctypes.data, ne, int(sorted),
                                        C.byref(n)))
        assert n.value == ne
        return hi, lo, w

    def export_graph(self, pinned: bool = False, out: Optional[dict] = None) -> dict:
        """What Convert::create_from builds (hm_gir.rs:156-226), canonical numbering: sorted
        nodes, sorted edges as (src index, dst index, weight) and in compress_edge bytes.
        pinned: the arrays live in page-locked memory (ktg_host_alloc), which takes the copies at PCIe speed;
        out: arrays of an earlier call to write into again (reused when large enough: page-locked memory is
        slow to allocate, ~0.3 ms per MiB)."""
        n, e = C.c_uint64(0), C.c_uint64(0)
        _check(self._L.ktg_graph_prepare(self._h, C.byref(n), C.byref(e)))
        nn, ne = n.value, e.value
        rec = int(self._L.ktg_edge_record_bytes(self._h))
        new = _pinned_empty if pinned else (lambda shape, dt: np.zeros(shape, dt))
        want = {"node_hi": (nn, np.uint64), "node_lo": (nn, np.uint64), "src": (ne, np.uint64), "dst": (ne, np.uint64),
                "weight": (ne, np.uint32), "edge_bytes": (ne * rec, np.uint8)}
        store = out if out is not None else {}
        res = {}
        for name, (cnt, dt) in want.items():
            base = store.get("_" + name)
            if base is None or base.size < cnt or base.dtype != dt:
                base = new(max(cnt + cnt // 8, 1), dt)
                store["_" + name] = base
            res[name] = base[:cnt]
        res["edge_bytes"] = res["edge_bytes"].reshape(ne, rec)
        _check(self._L.ktg_export_graph(self._h, res["node_hi"].ctypes.data, res["node_lo"].ctypes.data, nn,
                                        res["src"].ctypes.data, res["dst"].ctypes.data, res["weight"].ctypes.data,
                                        res["edge_bytes"].ctypes.data, ne))
        return res

    def export_externals(self):
        """`Externals` of remove_dead_paths (pruner.rs:165-195): (node indices ascending, kinds) with kind 0 =
        Input (no incoming edge), 1 = Output (no outgoing edge); indices are those of export_graph."""
        n = C.c_uint64(0)
        _check(self._L.ktg_export_externals(self._h, None, None, 0, C.byref(n)))
        ids, kinds = np.zeros(n.value, np.uint64), np.zeros(n.value, np.uint8)
        if n.value:
            _check(self._L.ktg_export_externals(self._h, ids.ctypes.data, kinds.ctypes.data, n.value, C.byref(n)))
        return ids, kinds

    def digest(self) -> Tuple[int, int, int, int]:
        out = (C.c_uint64 * 4)()
        _check(self._L.ktg_digest(self._h, out))
        return tuple(int(x) for x in out)

    def dump(self) -> str:
        """`sequence <kmer> weight <w>` lines (format of hs_gir.rs:288-290), sorted by k-mer."""
        hi, lo, w = self.export_edges(sorted=True)
        sym = "ACGT"
        lines = []
        for h, l, wt in zip(hi.tolist(), lo.tolist(), w.tolist()):
            v = (h << 64) | l
            kmer = "".join(sym[(v >> (2 * (self.k - 1 - j))) & 3] for j in range(self.k))
            lines.append(f"sequence {kmer} weight {wt}\n")
        return "".join(lines)

    # ---- hash sharding (world_size > 1) ----------------------------------------------
    def key_words(self) -> int:
        return int(self._L.ktg_key_words(self._h))

    def owner_of(self, hi: int, lo: int) -> int:
        return int(self._L.ktg_owner_of(self._h, hi, lo))

    def partition_reads_device(self, d_bases, d_offsets, n_reads: int, total_bases: int):
        """-> (device pointer of the owner-major key array, counts per owner)"""
        keys = C.c_void_p()
        counts = (C.c_uint64 * self.world_size)()
        _check(self._L.ktg_partition_reads_device(self._h, _ptr(d_bases), _ptr(d_offsets), n_reads, total_bases,
                                                  C.byref(keys), counts, None, None))
        return int(keys.value or 0), [int(c) for c in counts]

    def insert_keys_device(self, d_keys, n: int):
        _check(self._L.ktg_insert_keys_device(self._h, _ptr(d_keys), int(n)))

    # ---- whole-graph statistics of a sharded table ---------------------------------------------
    def nodes_export_device(self):
        """-> (keys pointer, degree-word pointer, n, u64 words per key) of this shard's nodes"""
        pk, pd, n, kw = C.c_void_p(), C.c_void_p(), C.c_uint64(), C.c_uint32()
        _check(self._L.ktg_nodes_export_device(self._h, C.byref(pk), C.byref(pd), C.byref(n), C.byref(kw)))
        return int(pk.value or 0), int(pd.value or 0), int(n.value), int(kw.value)

    def nodes_stats_from_device(self, d_keys, d_degrees, n: int) -> dict:
        st = L.KtgStats()
        _check(self._L.ktg_nodes_stats_from_device(self._h, _ptr(d_keys), _ptr(d_degrees), int(n), C.byref(st)))
        return {name: int(getattr(st, name)) for name, _ in L.KtgStats._fields_}

    def edge_sums(self, threshold: int):
        s, l = C.c_uint64(), C.c_uint64()
        _check(self._L.ktg_edge_sums(self._h, int(threshold), C.byref(s), C.byref(l)))
        return int(s.value), int(l.value)

    def scale_weights(self, ratio: float, threshold: int):
        _check(self._L.ktg_scale_weights(self._h, float(ratio), int(threshold)))

    def partition_keys_device(self, d_keys, n: int):
        """-> (device pointer of the owner-major copy of the keys, counts per owner)"""
        out = C.c_void_p()
        counts = (C.c_uint64 * self.world_size)()
        _check(self._L.ktg_partition_keys_device(self._h, _ptr(d_keys), int(n), C.byref(out), counts))
        return int(out.value or 0), [int(c) for c in counts]

    # ---- fused multi-GPU exchange (include/katome_gpu.h, "fused exchange") -----------------
    def mg_plan(self, max_windows: int) -> bool:
        need = C.c_int(0)
        _check(self._L.ktg_mg_plan(self._h, int(max_windows), C.byref(need)))
        return bool(need.value)

    def mg_prepare(self, max_windows: int):
        """-> (receive buffer pointer, bytes, bucket capacity in keys, sub-tables)"""
        base, nbytes, cap, n_sub = C.c_void_p(), C.c_uint64(), C.c_uint64(), C.c_uint32()
        _check(self._L.ktg_mg_prepare(self._h, int(max_windows), C.byref(base), C.byref(nbytes), C.byref(cap),
                                      C.byref(n_sub)))
        return int(base.value), int(nbytes.value), int(cap.value), int(n_sub.value)

    def mg_scatter_reads_device(self, d_bases, d_offsets, n_reads: int, total_bases: int, peer_rx, slot: int = 0,
                                first_of_batch: bool = True, send_stream: Optional[int] = None) -> int:
        """peer_rx: receive buffer pointers of all ranks (own rank included); -> cursors pointer"""
        arr = (C.c_void_p * self.world_size)(*[C.c_void_p(int(p)) for p in peer_rx])
        cur = C.c_void_p()
        _check(self._L.ktg_mg_scatter_reads_device(self._h, _ptr(d_bases), _ptr(d_offsets), int(n_reads),
                                                   int(total_bases), arr, int(slot), int(bool(first_of_batch)),
                                                   C.c_void_p(send_stream or None), C.byref(cur)))
        return int(cur.value)

    def mg_insert_buckets(self, d_bucket_ends, n_keys: int, slot: int = 0):
        _check(self._L.ktg_mg_insert_buckets(self._h, _ptr(d_bucket_ends), int(n_keys), int(slot)))

    # ---- the direct exchange (sender partitions by (owner, sub-table)) ----
    def mg_direct_plan(self, max_windows: int):
        """-> (needs_realloc, n_sub, sub_log2) of this shard"""
        r, ns, sl = C.c_int(0), C.c_uint32(0), C.c_uint32(0)
        _check(self._L.ktg_mg_direct_plan(self._h, int(max_windows), C.byref(r), C.byref(ns), C.byref(sl)))
        return bool(r.value), int(ns.value), int(sl.value)

    def mg_direct_prepare(self, max_windows: int):
        base, nbytes, cap = C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(self._L.ktg_mg_direct_prepare(self._h, int(max_windows), C.byref(base), C.byref(nbytes), C.byref(cap)))
        return int(base.value), int(nbytes.value), int(cap.value)

    def mg_direct_scatter_reads_device(self, d_bases, d_offsets, n_reads: int, total_bases: int, peer_rx, slot: int = 0,
                                       first_of_batch: bool = True, send_stream: Optional[int] = None) -> int:
        arr = (C.c_void_p * self.world_size)(*[C.c_void_p(int(p)) for p in peer_rx])
        cur = C.c_void_p()
        _check(self._L.ktg_mg_direct_scatter_reads_device(self._h, _ptr(d_bases), _ptr(d_offsets), int(n_reads),
                                                          int(total_bases), arr, int(slot), int(bool(first_of_batch)),
                                                          C.c_void_p(send_stream or None), C.byref(cur)))
        return int(cur.value)

    def mg_direct_insert(self, d_bucket_ends, n_keys: int, slot: int = 0):
        _check(self._L.ktg_mg_direct_insert(self._h, _ptr(d_bucket_ends), int(n_keys), int(slot)))

    def mg_sketch(self):
        p, n = C.c_void_p(), C.c_uint32()
        _check(self._L.ktg_mg_sketch(self._h, C.byref(p), C.byref(n)))
        return int(p.value), int(n.value)

    def mg_merge_sketch(self, d_regs):
        _check(self._L.ktg_mg_merge_sketch(self._h, _ptr(d_regs)))

    def mg_spill(self):
        p, n = C.c_void_p(), C.c_uint64()
        _check(self._L.ktg_mg_spill(self._h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def mg_insert_spill(self, d_keys, n: int):
        _check(self._L.ktg_mg_insert_spill(self._h, _ptr(d_keys), int(n)))

    # ---- the same exchange in super-k-mer records (include/katome_gpu.h) ---------------------
    def mg_skm_plan(self, max_windows: int) -> bool:
        need = C.c_int(0)
        _check(self._L.ktg_mg_skm_plan(self._h, int(max_windows), C.byref(need)))
        return bool(need.value)

    def mg_skm_prepare(self, max_windows: int):
        """-> (receive buffer pointer, bytes of one slot, bucket capacity in records)"""
        base, nbytes, cap = C.c_void_p(), C.c_uint64(), C.c_uint64()
        _check(self._L.ktg_mg_skm_prepare(self._h, int(max_windows), C.byref(base), C.byref(nbytes), C.byref(cap)))
        return int(base.value), int(nbytes.value), int(cap.value)

    def mg_skm_scatter_reads_device(self, d_bases, d_offsets, n_reads: int, total_bases: int, peer_rx, slot: int = 0,
                                    first_of_batch: bool = True, send_stream: Optional[int] = None):
        """-> (bucket ends pointer, key counts pointer): world u64 each"""
        arr = (C.c_void_p * self.world_size)(*[C.c_void_p(int(p)) for p in peer_rx])
        cur, kc = C.c_void_p(), C.c_void_p()
        _check(self._L.ktg_mg_skm_scatter_reads_device(self._h, _ptr(d_bases), _ptr(d_offsets), int(n_reads),
                                                       int(total_bases), arr, int(slot), int(bool(first_of_batch)),
                                                       C.c_void_p(send_stream or None), C.byref(cur), C.byref(kc)))
        return int(cur.value), int(kc.value)

    def mg_skm_insert_buckets(self, d_bucket_ends, n_keys_ub: int, slot: int = 0):
        _check(self._L.ktg_mg_skm_insert_buckets(self._h, _ptr(d_bucket_ends), int(n_keys_ub), int(slot)))

    def mg_skm_spill(self):
        p, n = C.c_void_p(), C.c_uint64()
        _check(self._L.ktg_mg_skm_spill(self._h, C.byref(p), C.byref(n)))
        return int(p.value or 0), int(n.value)

    def mg_skm_partition_records(self, d_records, n: int):
        out = C.c_void_p()
        counts = (C.c_uint64 * self.world_size)()
        _check(self._L.ktg_mg_skm_partition_records(self._h, _ptr(d_records), int(n), C.byref(out), counts))
        return int(out.value or 0), [int(c) for c in counts]

    def mg_skm_insert_records(self, d_records, n: int):
        _check(self._L.ktg_mg_skm_insert_records(self._h, _ptr(d_records), int(n)))

    def mg_skm_owner_of(self, hi: int, lo: int) -> int:
        return int(self._L.ktg_mg_skm_owner_of(self._h, hi, lo))

    # ---- observability -----------------------------------------------------------------
    def profile(self) -> dict:
        n = C.c_uint32(0)
        arr = (L.KtgKernelProfile * 64)()
        _check(self._L.ktg_get_profile(self._h, arr, 64, C.byref(n)))
        return {arr[i].name.decode(): {"launches": int(arr[i].launches), "ms": float(arr[i].total_ms),
                                       "units": int(arr[i].units)} for i in range(min(n.value, 64))}

    def reset_profile(self):
        _check(self._L.ktg_reset_profile(self._h))

    def set_profile(self, enabled: bool):
        """Per-launch timing on / off for the launches that follow (ktg_set_profile)."""
        _check(self._L.ktg_set_profile(self._h, int(bool(enabled))))

    def info(self) -> dict:
        i = L.KtgInfo()
        _check(self._L.ktg_get_info(self._h, C.byref(i)))
        return {name: int(getattr(i, name)) for name, _ in L.KtgInfo._fields_}


def skm_supported(k: int) -> bool:
    """can the multi-GPU exchange of this k run in super-k-mer records?"""
    return bool(L.lib().ktg_mg_skm_supported(int(k)))


def ipc_get_handle(dev_ptr: int) -> bytes:
    buf = C.create_string_buffer(64)
    _check(L.lib().ktg_ipc_get_handle(C.c_void_p(int(dev_ptr)), buf))
    return buf.raw


def ipc_open(handle: bytes) -> int:
    p = C.c_void_p()
    _check(L.lib().ktg_ipc_open(C.create_string_buffer(bytes(handle), 64), C.byref(p)))
    return int(p.value)


def ipc_close(dev_ptr: int):
    _check(L.lib().ktg_ipc_close(C.c_void_p(int(dev_ptr))))


def synth_reads_device(d_out, seed_g: int, genome_len: int, read_len: int, err_ppm: int, r0: int, r1: int,
                       stream: int = 0):
    """`stream` is a raw cudaStream_t handle (0 = the legacy default stream)."""
    _check(L.lib().ktg_synth_reads_device(_ptr(d_out), seed_g, genome_len, read_len, err_ppm, r0, r1,
                                          C.c_void_p(stream or None)))


def random_access_probe(table_bytes: int, n_updates: int, slot_bytes: int = 16) -> float:
    ms = C.c_float(0)
    _check(L.lib().ktg_random_access_probe(table_bytes, n_updates, slot_bytes, C.byref(ms)))
    return float(ms.value)
