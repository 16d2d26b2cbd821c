"""ctypes binding of the CPU oracle (oracle/katome_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference leg.  The product package
(katome_b200/) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libkatome_oracle.so")

KO_OK, KO_ERR_SHORT_READ, KO_ERR_BAD_K, KO_ERR_IO, KO_ERR_BAD_RECORD, KO_ERR_DEGENERATE = range(6)


class OracleError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what} (oracle code {code})")
        self.code = code


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in ("katome_oracle.c", "katome_oracle_mt.c", "katome_oracle.h")]
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u8p, u32p, u64p = C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.ko_new.restype = C.c_void_p
        L.ko_new.argtypes = [C.c_int]
        L.ko_free.argtypes = [C.c_void_p]
        L.ko_add_read_fastaq.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.ko_add_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, u64p, u64p]
        L.ko_create_from_files.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_int, u64p, u64p]
        L.ko_add_read_bfc.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t, C.c_uint32, C.c_int]
        L.ko_create_from_bfc_files.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.c_int, C.c_int, C.c_uint32, u64p, u64p]
        L.ko_counts.argtypes = [C.c_void_p, u64p, u64p]
        L.ko_collection_stats.argtypes = [C.c_void_p, u64p]
        L.ko_remove_weak_edges.argtypes = [C.c_void_p, C.c_uint32]
        L.ko_remove_single_vertices.argtypes = [C.c_void_p]
        L.ko_standardize_edges.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
        L.ko_standardize_weights.restype = C.c_uint64
        L.ko_standardize_weights.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_int)]
        L.ko_export_edges.restype = C.c_uint64
        L.ko_export_edges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.ko_export_nodes.restype = C.c_uint64
        L.ko_export_nodes.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.ko_digest.argtypes = [C.c_void_p, u64p]
        L.ko_dump.restype = C.c_uint64
        L.ko_dump.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64]
        L.ko_encode_fasta_symbol.restype = C.c_uint8
        L.ko_encode_fasta_symbol.argtypes = [C.c_uint8, C.c_uint8]
        for name in ("ko_compress_node", "ko_compress_kmer", "ko_compress_edge", "ko_decompress_edge"):
            getattr(L, name).restype = C.c_size_t
            getattr(L, name).argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.ko_compress_kmer_with_rev_compl.restype = C.c_size_t
        L.ko_compress_kmer_with_rev_compl.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_char_p]
        L.ko_reverse_compressed_node.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t, C.c_char_p]
        L.ko_shift_right_bit_array.argtypes = [C.c_char_p, C.c_size_t, C.c_size_t]
        L.ko_splitmix64.restype = C.c_uint64
        L.ko_splitmix64.argtypes = [C.c_uint64]
        L.ko_synth_reads.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p]
        L.ko_synth_genome.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.ko_mt_build_digest.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, u64p, u64p, u64p]
        L.ko_mt_new.restype = C.c_void_p
        L.ko_mt_new.argtypes = [C.c_int, C.c_int, C.c_int]
        L.ko_mt_free.argtypes = [C.c_void_p]
        L.ko_mt_add_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        L.ko_mt_counters.argtypes = [C.c_void_p, u64p, u64p]
        L.ko_mt_digest.argtypes = [C.c_void_p, u64p]
        L.ko_mt_remove_weak_edges.argtypes = [C.c_void_p, C.c_uint32]
        L.ko_mt_standardize_edges.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint32]
        L.ko_synth_reads_mt.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, C.c_void_p, C.c_int]
        _lib = L
    return _lib


_ERR_TEXT = {
    KO_ERR_SHORT_READ: "Read is too short!",
    KO_ERR_BAD_K: "bad k",
    KO_ERR_IO: "Couldn't open all files",
    KO_ERR_BAD_RECORD: "malformed record",
    KO_ERR_DEGENERATE: "degenerate standardization ratio",
}


def _check(code: int):
    if code != KO_OK:
        raise OracleError(code, _ERR_TEXT.get(code, "oracle error"))


class OracleGIR:
    """CPU GIR with the reference's Init/Build/Clean/Standardize/Stats surface."""

    def __init__(self, k: int):
        self._L = lib()
        self._h = self._L.ko_new(k)
        if not self._h:
            raise OracleError(KO_ERR_BAD_K, "k must be in 3..=64")
        self.k = k
        self.accepted_reads = 0
        self.accepted_bytes = 0

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.ko_free(self._h)
            self._h = None

    # -- Build ------------------------------------------------------------
    def add_read_fastaq(self, read: bytes, reverse_complement: bool):
        _check(self._L.ko_add_read_fastaq(self._h, read, len(read), int(reverse_complement)))

    def add_reads(self, bases: np.ndarray, offsets: np.ndarray, reverse_complement: bool):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nr, nb = C.c_uint64(0), C.c_uint64(0)
        code = self._L.ko_add_reads(self._h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                    int(reverse_complement), C.byref(nr), C.byref(nb))
        self.accepted_reads += nr.value
        self.accepted_bytes += nb.value
        _check(code)

    @classmethod
    def create(cls, k: int, files, file_type: str = "fastq", reverse_complement: bool = False):
        g = cls(k)
        arr = (C.c_char_p * len(files))(*[os.fsencode(f) for f in files])
        nr, nb = C.c_uint64(0), C.c_uint64(0)
        code = g._L.ko_create_from_files(g._h, arr, len(files), 1 if file_type.lower() == "fasta" else 0,
                                         int(reverse_complement), C.byref(nr), C.byref(nb))
        g.accepted_reads, g.accepted_bytes = nr.value, nb.value
        _check(code)
        return g, nb.value

    def add_read_bfc(self, kmer: bytes, weight: int, reverse_complement: bool):
        """one BFCounter line (pt_graph.rs:318-329)"""
        _check(self._L.ko_add_read_bfc(self._h, kmer, len(kmer), int(weight), int(reverse_complement)))

    @classmethod
    def create_bfc(cls, k: int, files, reverse_complement: bool = False, minimal_weight_threshold: int = 0):
        """create_bfc (builder.rs:79-115) -> (collection, total accepted bytes)"""
        g = cls(k)
        arr = (C.c_char_p * len(files))(*[os.fsencode(f) for f in files])
        nk, nb = C.c_uint64(0), C.c_uint64(0)
        code = g._L.ko_create_from_bfc_files(g._h, arr, len(files), int(reverse_complement),
                                             int(minimal_weight_threshold), C.byref(nk), C.byref(nb))
        g.accepted_reads, g.accepted_bytes = nk.value, nb.value
        _check(code)
        return g, nb.value

    # -- Stats ------------------------------------------------------------
    def counts(self):
        n, e = C.c_uint64(0), C.c_uint64(0)
        self._L.ko_counts(self._h, C.byref(n), C.byref(e))
        return n.value, e.value

    def collection_stats(self) -> dict:
        out = (C.c_uint64 * 8)()
        self._L.ko_collection_stats(self._h, out)
        names = ("node_count", "edge_count", "max_edge_weight", "sum_edge_weight", "max_in_degree",
                 "max_out_degree", "incoming_vert_count", "outgoing_vert_count")
        d = dict(zip(names, [int(x) for x in out]))
        d["avg_edge_weight"] = d["sum_edge_weight"] / d["edge_count"] if d["edge_count"] else float("nan")
        d["avg_out_degree"] = d["edge_count"] / d["node_count"] if d["node_count"] else float("nan")
        return d

    # -- Clean / Standardize ------------------------------------------------
    def remove_weak_edges(self, threshold: int):
        self._L.ko_remove_weak_edges(self._h, threshold)

    def remove_single_vertices(self):
        self._L.ko_remove_single_vertices(self._h)

    def standardize_edges(self, genome_len: int, k: int, threshold: int):
        _check(self._L.ko_standardize_edges(self._h, genome_len, k, threshold))

    # -- export -------------------------------------------------------------
    def export_edges(self):
        _, ne = self.counts()
        hi = np.empty(ne, np.uint64)
        lo = np.empty(ne, np.uint64)
        w = np.empty(ne, np.uint32)
        n = self._L.ko_export_edges(self._h, hi.ctypes.data, lo.ctypes.data, w.ctypes.data, ne)
        assert n == ne
        return hi, lo, w

    def export_nodes(self):
        nn, _ = self.counts()
        hi = np.empty(nn, np.uint64)
        lo = np.empty(nn, np.uint64)
        n = self._L.ko_export_nodes(self._h, hi.ctypes.data, lo.ctypes.data, nn)
        assert n == nn
        return hi, lo

    def digest(self):
        out = (C.c_uint64 * 4)()
        self._L.ko_digest(self._h, out)
        return tuple(int(x) for x in out)

    def dump(self) -> str:
        need = self._L.ko_dump(self._h, None, 0)
        buf = C.create_string_buffer(int(need) + 1)
        self._L.ko_dump(self._h, buf, need)
        return buf.raw[:need].decode()


# ---- codec helpers (byte level, for the reference's known-answer vectors) ----
def encode_fasta_symbol(sym: int, carrier: int = 0) -> int:
    return lib().ko_encode_fasta_symbol(sym, carrier)


def _bytes_call(fn, data: bytes, cap: int) -> bytes:
    out = C.create_string_buffer(cap)
    n = fn(data, len(data), out)
    return out.raw[:n]


def compress_node(s: bytes) -> bytes:
    return _bytes_call(lib().ko_compress_node, s, len(s) // 4 + 2)


def compress_kmer(s: bytes) -> bytes:
    return _bytes_call(lib().ko_compress_kmer, s, len(s) // 2 + 4)


def compress_edge(s: bytes) -> bytes:
    return _bytes_call(lib().ko_compress_edge, s, len(s) // 4 + 3)


def decompress_edge(s: bytes) -> bytes:
    return _bytes_call(lib().ko_decompress_edge, s, len(s) * 4 + 4)


def compress_kmer_with_rev_compl(s: bytes):
    cap = len(s) // 2 + 4
    out, rev = C.create_string_buffer(cap), C.create_string_buffer(cap)
    n = lib().ko_compress_kmer_with_rev_compl(s, len(s), out, rev)
    return out.raw[:n], rev.raw[:n]


def reverse_compressed_node(b: bytes, remainder: int) -> bytes:
    out = C.create_string_buffer(len(b))
    lib().ko_reverse_compressed_node(b, len(b), remainder, out)
    return out.raw[:len(b)]


def shift_right_bit_array(b: bytes, shift: int) -> bytes:
    buf = C.create_string_buffer(b, len(b))
    lib().ko_shift_right_bit_array(buf, len(b), shift)
    return buf.raw[:len(b)]


def splitmix64(x: int) -> int:
    return lib().ko_splitmix64(x & 0xFFFFFFFFFFFFFFFF)


def synth_reads(seed_g: int, G: int, L: int, err_ppm: int, r0: int, r1: int) -> np.ndarray:
    out = np.empty((r1 - r0) * L, np.uint8)
    lib().ko_synth_reads(seed_g, G, L, err_ppm, r0, r1, out.ctypes.data)
    return out


def host_threads() -> int:
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)


def synth_reads_mt(seed_g: int, G: int, L: int, err_ppm: int, r0: int, r1: int, n_threads: int = 0) -> np.ndarray:
    """synth_reads over every host thread (counter based generator: the same bytes)"""
    out = np.empty((r1 - r0) * L, np.uint8)
    lib().ko_synth_reads_mt(seed_g, G, L, err_ppm, r0, r1, out.ctypes.data, n_threads or host_threads())
    return out


def synth_genome(seed_g: int, pos0: int, n: int) -> np.ndarray:
    out = np.empty(n, np.uint8)
    lib().ko_synth_genome(seed_g, pos0, n, out.ctypes.data)
    return out


def mt_build_digest(k: int, bases: np.ndarray, offsets: np.ndarray, reverse_complement: bool, n_threads: int):
    """The "optimistic CPU" counter (katome_oracle_mt.c): hash-sharded over n_threads host threads.
    Returns (digest tuple as OracleGIR.digest(), accepted_reads, accepted_bytes)."""
    bases = np.ascontiguousarray(bases, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    out = (C.c_uint64 * 4)()
    nr, nb = C.c_uint64(0), C.c_uint64(0)
    _check(lib().ko_mt_build_digest(k, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                                    int(reverse_complement), int(n_threads), out, C.byref(nr), C.byref(nb)))
    return tuple(int(x) for x in out), nr.value, nb.value


class MtCounter:
    """katome_oracle_mt.c as an object: the multi-threaded CPU counter that finishes the BASELINE
    configurations in seconds and is digest-equal to the faithful OracleGIR (tests/test_oracle_golden.py).
    The checker of the full-size GPU builds: reads batch by batch, remove_weak_edges, standardize_edges,
    digest = OracleGIR.digest()'s four numbers."""

    def __init__(self, k: int, reverse_complement: bool, n_threads: int = 0):
        self._h = lib().ko_mt_new(int(k), int(bool(reverse_complement)), int(n_threads or host_threads()))
        if not self._h:
            raise OracleError(KO_ERR_BAD_K, "bad k")
        self.k = int(k)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib().ko_mt_free(h)

    def add_reads(self, bases: np.ndarray, offsets: np.ndarray):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(lib().ko_mt_add_reads(self._h, bases.ctypes.data, offsets.ctypes.data, len(offsets) - 1))

    def add_reads_ptr(self, bases_ptr: int, offsets_ptr: int, n_reads: int):
        _check(lib().ko_mt_add_reads(self._h, bases_ptr, offsets_ptr, n_reads))

    def counters(self):
        nr, nb = C.c_uint64(0), C.c_uint64(0)
        lib().ko_mt_counters(self._h, C.byref(nr), C.byref(nb))
        return nr.value, nb.value

    def digest(self):
        out = (C.c_uint64 * 4)()
        _check(lib().ko_mt_digest(self._h, out))
        return tuple(int(x) for x in out)

    def remove_weak_edges(self, threshold: int):
        _check(lib().ko_mt_remove_weak_edges(self._h, int(threshold)))

    def standardize_edges(self, genome_len: int, k: int, threshold: int):
        _check(lib().ko_mt_standardize_edges(self._h, int(genome_len), int(k), int(threshold)))
