"""ctypes binding of libkatome_gpu.so (the C ABI in include/katome_gpu.h).

There is no CPU fallback: if the shared library is missing this raises, and if
no CUDA device is present `ktg_create` fails with KTG_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "lib", "libkatome_gpu.so")
CSRC = os.path.join(_HERE, "csrc")

KTG_ABI_VERSION = 2
(KTG_OK, KTG_ERR_SHORT_READ, KTG_ERR_BAD_K, KTG_ERR_IO, KTG_ERR_BAD_RECORD, KTG_ERR_DEGENERATE,
 KTG_ERR_TABLE_FULL, KTG_ERR_CUDA, KTG_ERR_INVALID, KTG_ERR_NO_DEVICE) = range(10)
KTG_FASTQ, KTG_FASTA = 0, 1
KTG_FLAG_PROFILE, KTG_FLAG_FORCE_DIRECT, KTG_FLAG_FORCE_PARTITION = 1, 2, 4
KTG_FLAG_FORCE_PAGES, KTG_FLAG_NO_PAGES = 8, 16

# every symbol include/katome_gpu.h declares
SYMBOLS = (
    "ktg_create", "ktg_destroy", "ktg_last_error", "ktg_device_count", "ktg_add_reads",
    "ktg_add_reads_device", "ktg_create_from_files", "ktg_reset", "ktg_finalize", "ktg_counts",
    "ktg_collection_stats", "ktg_remove_weak_edges", "ktg_remove_single_vertices",
    "ktg_standardize_edges", "ktg_export_edges", "ktg_digest", "ktg_key_words", "ktg_owner_of",
    "ktg_partition_reads_device", "ktg_insert_keys_device", "ktg_host_alloc", "ktg_host_free",
    "ktg_synth_reads_device", "ktg_random_access_probe", "ktg_get_profile", "ktg_reset_profile", "ktg_set_profile", "ktg_plan_chunks", "ktg_item_reads", "ktg_host_parse_file",
    "ktg_get_info", "ktg_set_option", "ktg_export_externals", "ktg_graph_prepare", "ktg_wait_input", "ktg_partition_keys_device", "ktg_mg_plan", "ktg_mg_prepare",
    "ktg_mg_scatter_reads_device", "ktg_mg_direct_plan", "ktg_mg_direct_prepare", "ktg_mg_direct_scatter_reads_device",
    "ktg_mg_direct_insert", "ktg_mg_insert_buckets", "ktg_mg_sketch", "ktg_mg_merge_sketch", "ktg_mg_spill", "ktg_mg_insert_spill", "ktg_ipc_get_handle", "ktg_ipc_open", "ktg_ipc_close",
    "ktg_mg_skm_supported", "ktg_mg_skm_plan", "ktg_mg_skm_prepare", "ktg_mg_skm_scatter_reads_device",
    "ktg_mg_skm_insert_buckets", "ktg_mg_skm_spill", "ktg_mg_skm_partition_records", "ktg_mg_skm_insert_records",
    "ktg_mg_skm_owner_of", "ktg_skm_items_host", "ktg_skm_owner_of_kmer",
    "ktg_add_weighted_kmers", "ktg_create_from_bfc_files",
    "ktg_export_graph", "ktg_edge_record_bytes", "ktg_nodes_export_device", "ktg_nodes_stats_from_device", "ktg_edge_sums", "ktg_scale_weights",
)


class KtgConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("k", C.c_uint32), ("reverse_complement", C.c_uint32),
        ("device", C.c_int32), ("capacity_hint_edges", C.c_uint64), ("world_size", C.c_uint32),
        ("rank", C.c_uint32), ("stream", C.c_void_p), ("sub_table_log2_bytes", C.c_uint32),
        ("flags", C.c_uint32), ("n_devices", C.c_uint32), ("device_ids", C.POINTER(C.c_int32)),
    ]


class KtgStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "node_count", "edge_count", "max_edge_weight", "sum_edge_weight", "max_in_degree",
        "max_out_degree", "incoming_vert_count", "outgoing_vert_count")]


class KtgKernelProfile(C.Structure):
    _fields_ = [("name", C.c_char * 32), ("launches", C.c_uint64), ("total_ms", C.c_double),
                ("units", C.c_uint64)]


class KtgInfo(C.Structure):
    _fields_ = [("capacity_slots", C.c_uint64), ("occupied_slots", C.c_uint64),
                ("table_bytes", C.c_uint64), ("n_sub_tables", C.c_uint32), ("slot_bytes", C.c_uint32),
                ("windows_inserted", C.c_uint64), ("kernel_launches", C.c_uint64),
                ("grow_events", C.c_uint32), ("partitioned", C.c_uint32),
                ("page_updates", C.c_uint32), ("n_pages", C.c_uint32)]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the CUDA library in-tree (nvcc, sm_100a).  Works without a GPU."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(_HERE, "..", "include", "katome_gpu.h")]
    stale = (not os.path.exists(SO_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(SO_PATH) for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", CSRC] + (["-B"] if force else []), capture_output=True, text=True)
        if verbose:
            print(r.stdout, r.stderr)
        if r.returncode != 0:
            raise RuntimeError("building libkatome_gpu.so failed:\n" + r.stdout + r.stderr)
    return SO_PATH


_lib = None


def lib():
    """Load the library; raise loudly when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(katome_b200 has no CPU fallback)")
    L = C.CDLL(SO_PATH)
    vp, u64p, u32p = C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)
    L.ktg_last_error.restype = C.c_char_p
    L.ktg_device_count.restype = C.c_int
    L.ktg_create.argtypes = [C.POINTER(KtgConfig), C.POINTER(vp)]
    L.ktg_destroy.argtypes = [vp]
    L.ktg_destroy.restype = None
    L.ktg_add_reads.argtypes = [vp, vp, vp, C.c_uint64, u64p, u64p]
    L.ktg_add_reads_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, u64p, u64p]
    L.ktg_create_from_files.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint32, C.c_int, u64p]
    L.ktg_add_weighted_kmers.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32, u64p, u64p]
    L.ktg_create_from_bfc_files.argtypes = [vp, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint32, u64p]
    L.ktg_finalize.argtypes = [vp]
    L.ktg_wait_input.argtypes = [vp]
    L.ktg_reset.argtypes = [vp]
    L.ktg_counts.argtypes = [vp, u64p, u64p]
    L.ktg_collection_stats.argtypes = [vp, C.POINTER(KtgStats)]
    L.ktg_remove_weak_edges.argtypes = [vp, C.c_uint32]
    L.ktg_remove_single_vertices.argtypes = [vp]
    L.ktg_standardize_edges.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32]
    L.ktg_export_edges.argtypes = [vp, vp, vp, vp, C.c_uint64, C.c_int, u64p]
    L.ktg_digest.argtypes = [vp, u64p]
    L.ktg_export_graph.argtypes = [vp, vp, vp, C.c_uint64, vp, vp, vp, vp, C.c_uint64]
    L.ktg_export_externals.argtypes = [vp, vp, vp, C.c_uint64, u64p]
    L.ktg_graph_prepare.argtypes = [vp, u64p, u64p]
    L.ktg_edge_record_bytes.argtypes = [vp]
    L.ktg_edge_record_bytes.restype = C.c_uint32
    L.ktg_key_words.argtypes = [vp]
    L.ktg_key_words.restype = C.c_uint32
    L.ktg_owner_of.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.ktg_owner_of.restype = C.c_uint32
    L.ktg_partition_reads_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.POINTER(vp), u64p, u64p, u64p]
    L.ktg_insert_keys_device.argtypes = [vp, vp, C.c_uint64]
    L.ktg_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.ktg_host_free.argtypes = [vp]
    L.ktg_host_free.restype = None
    L.ktg_synth_reads_device.argtypes = [vp, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64, C.c_uint64, vp]
    L.ktg_random_access_probe.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.POINTER(C.c_float)]
    L.ktg_get_profile.argtypes = [vp, C.POINTER(KtgKernelProfile), C.c_uint32, u32p]
    L.ktg_reset_profile.argtypes = [vp]
    L.ktg_set_profile.argtypes = [vp, C.c_int]
    L.ktg_host_parse_file.argtypes = [C.c_char_p, C.c_int, C.c_uint64, u64p, u64p, u64p]
    L.ktg_plan_chunks.argtypes = [vp, C.c_uint64, C.c_uint64, vp, C.c_uint32, vp, vp, C.c_uint32, vp, vp]
    L.ktg_item_reads.argtypes = [vp, C.c_uint64, C.c_uint32, vp]
    L.ktg_get_info.argtypes = [vp, C.POINTER(KtgInfo)]
    L.ktg_set_option.argtypes = [vp, C.c_char_p, C.c_int64]
    intp = C.POINTER(C.c_int)
    L.ktg_partition_keys_device.argtypes = [vp, vp, C.c_uint64, C.POINTER(vp), u64p]
    L.ktg_mg_plan.argtypes = [vp, C.c_uint64, intp]
    L.ktg_mg_prepare.argtypes = [vp, C.c_uint64, C.POINTER(vp), u64p, u64p, u32p]
    L.ktg_mg_scatter_reads_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.POINTER(vp), C.c_uint32, C.c_int,
                                              vp, C.POINTER(vp)]
    L.ktg_mg_insert_buckets.argtypes = [vp, vp, C.c_uint64, C.c_uint32]
    L.ktg_mg_direct_plan.argtypes = [vp, C.c_uint64, intp, u32p, u32p]
    L.ktg_mg_direct_prepare.argtypes = [vp, C.c_uint64, C.POINTER(vp), u64p, u64p]
    L.ktg_mg_direct_scatter_reads_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.POINTER(vp), C.c_uint32, C.c_int,
                                                     vp, C.POINTER(vp)]
    L.ktg_mg_direct_insert.argtypes = [vp, vp, C.c_uint64, C.c_uint32]
    L.ktg_mg_sketch.argtypes = [vp, C.POINTER(vp), u32p]
    L.ktg_mg_merge_sketch.argtypes = [vp, vp]
    L.ktg_mg_spill.argtypes = [vp, C.POINTER(vp), u64p]
    L.ktg_mg_insert_spill.argtypes = [vp, vp, C.c_uint64]
    L.ktg_mg_skm_supported.argtypes = [C.c_uint32]
    L.ktg_mg_skm_plan.argtypes = [vp, C.c_uint64, intp]
    L.ktg_mg_skm_prepare.argtypes = [vp, C.c_uint64, C.POINTER(vp), u64p, u64p]
    L.ktg_mg_skm_scatter_reads_device.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint64, C.POINTER(vp), C.c_uint32,
                                                  C.c_int, vp, C.POINTER(vp), C.POINTER(vp)]
    L.ktg_mg_skm_insert_buckets.argtypes = [vp, vp, C.c_uint64, C.c_uint32]
    L.ktg_mg_skm_spill.argtypes = [vp, C.POINTER(vp), u64p]
    L.ktg_mg_skm_partition_records.argtypes = [vp, vp, C.c_uint64, C.POINTER(vp), u64p]
    L.ktg_mg_skm_insert_records.argtypes = [vp, vp, C.c_uint64]
    L.ktg_mg_skm_owner_of.argtypes = [vp, C.c_uint64, C.c_uint64]
    L.ktg_mg_skm_owner_of.restype = C.c_uint32
    L.ktg_skm_items_host.argtypes = [vp, vp, vp, C.c_uint64, C.c_uint32, C.c_uint32, vp, C.c_uint64, u64p]
    L.ktg_skm_owner_of_kmer.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
    L.ktg_skm_owner_of_kmer.restype = C.c_uint32
    L.ktg_ipc_get_handle.argtypes = [vp, C.c_char_p]
    L.ktg_ipc_open.argtypes = [C.c_char_p, C.POINTER(vp)]
    L.ktg_ipc_close.argtypes = [vp]
    L.ktg_nodes_export_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), u64p, u32p]
    L.ktg_nodes_stats_from_device.argtypes = [vp, vp, vp, C.c_uint64, C.POINTER(KtgStats)]
    L.ktg_edge_sums.argtypes = [vp, C.c_uint32, u64p, u64p]
    L.ktg_scale_weights.argtypes = [vp, C.c_double, C.c_uint32]
    _lib = L
    return L
