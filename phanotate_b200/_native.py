"""ctypes binding of the C ABI in include/phanotate_b200.h -- the only way Python reaches the kernels.

``load()`` opens ``phanotate_b200/libpb200.so`` (the sm_100a CUDA build).  There is no CPU fallback:
if the library is missing or no CUDA device can be opened, importing callers get an exception.
Tests may pass an explicit path to the host-simulation build of the same sources
(tests/native/pb200_hostsim.so); the package itself never does.
"""
from __future__ import annotations

import ctypes
import os
from decimal import Decimal

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libpb200.so")

DEC = np.dtype([("c", "<u4", (4,)), ("e", "<i4"), ("neg", "<i4")])
CALL = np.dtype([("contig", "<i4"), ("left", "<i4"), ("right", "<i4"), ("strand", "<i4"),
                 ("weight", DEC), ("score", "<f8")])
CALL24 = np.dtype([("contig", "<i4"), ("left", "<i4"), ("right", "<i4"), ("strand", "<i4"), ("score", "<f8")])   # pb200_call24
ORF = np.dtype([("contig", "<i4"), ("start", "<i4"), ("stop", "<i4"), ("frame", "<i4"), ("rbs_score", "<i4"),
                ("trigger", "<i4"), ("start_weight", "<i4"), ("node", "<i4"), ("pstop", DEC), ("weight", DEC)])
NODE = np.dtype([("contig", "<i4"), ("position", "<i4"), ("kind", "<i4"), ("frame", "<i4"), ("mate", "<i4"),
                 ("orf", "<i4"), ("other_end", "<i4"), ("trigger", "<i4")])
EDGE = np.dtype([("contig", "<i4"), ("src", "<i4"), ("dst", "<i4"), ("kind", "<i4"), ("weight", DEC)])
CONTIG = np.dtype([("length", "<i4"), ("err", "<u4"), ("node_off", "<i4"), ("n_nodes", "<i4"), ("orf_off", "<i4"),
                   ("n_orfs", "<i4"), ("call_off", "<i4"), ("n_calls", "<i4"), ("n_ties", "<i4"), ("wide", "<i4"),
                   ("pstop", DEC), ("pos_max", DEC, (4,)), ("pos_min", DEC, (4,)),
                   ("background_rbs", "<f8", (28,)), ("training_rbs", "<f8", (28,))])
PARAMS = np.dtype([("n_start", "<i4"), ("start_codon", "S4", (8,)), ("start_weight", DEC, (8,)),
                   ("n_stop", "<i4"), ("stop_codon", "S4", (8,)), ("min_orf_len", "<i4"), ("reserved", "<i4")])
assert CALL.itemsize == 48 and ORF.itemsize == 80 and NODE.itemsize == 32 and EDGE.itemsize == 40
assert PARAMS.itemsize == 4 + 32 + 192 + 4 + 32 + 8

ERR_CHAR, ERR_RANGE, ERR_PARALLEL, ERR_OVERFLOW, ERR_NOPATH, ERR_INTERNAL, ERR_LOOKUP, ERR_TIES = 1, 2, 4, 8, 16, 32, 64, 128
INPUT_DEVICE = 1
REUSE_INPUT = 2
LITERAL = 4
SCAN_REFERENCE = 8
CALL_WEIGHTS = 16
SOLVE_WIDE = 32
SOLVE_PLAIN = 64
SOLVE_NOCHUNK = 128
INPUT_PACKED4 = 256
NODE_SOURCE, NODE_TARGET = -2, -3


def dec_to_decimal(rec) -> Decimal:
    """pb200_dec record -> decimal.Decimal with identical sign, digits and exponent."""
    c = int(rec["c"][0]) | (int(rec["c"][1]) << 32) | (int(rec["c"][2]) << 64) | (int(rec["c"][3]) << 96)
    return Decimal((int(rec["neg"]), tuple(int(ch) for ch in str(c)), int(rec["e"])))


def decimal_to_dec(d: Decimal):
    s, digits, e = Decimal(d).as_tuple()
    c = int("".join(map(str, digits)))
    if c >= 10 ** 38:
        raise ValueError("coefficient too large")
    return ((c & 0xFFFFFFFF, (c >> 32) & 0xFFFFFFFF, (c >> 64) & 0xFFFFFFFF, (c >> 96) & 0xFFFFFFFF), e, s)


def load(path: str | None = None) -> ctypes.CDLL:
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError("phanotate_b200: CUDA library %s not built (run `python -c 'import __graft_entry__ as g; "
                           "g.build()'`); there is no CPU fallback" % p)
    lib = ctypes.CDLL(p)
    vp, i32, i64p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p
    lib.pb200_create.argtypes = [ctypes.c_int, ctypes.POINTER(vp)]
    lib.pb200_destroy.argtypes = [vp]
    lib.pb200_destroy.restype = None
    lib.pb200_last_error.argtypes = [vp]
    lib.pb200_last_error.restype = ctypes.c_char_p
    lib.pb200_run.argtypes = [vp, vp, i64p, i32, vp, ctypes.c_uint32]
    lib.pb200_sizes.argtypes = [vp, vp]
    lib.pb200_upload.argtypes = [vp, vp, i64p, i32]
    lib.pb200_pack4.argtypes = [vp, ctypes.c_int64, vp]
    lib.pb200_pack4.restype = ctypes.c_int64
    lib.pb200_upload_packed4.argtypes = [vp, vp, i32, i64p, i32]
    lib.pb200_upload_async.argtypes = [vp, vp, vp, i32, i64p, i32]
    lib.pb200_prefetch_async.argtypes = [vp, vp, vp, i32, ctypes.c_int64]
    lib.pb200_set_contig_base.argtypes = [vp, i32]
    lib.pb200_set_chunking.argtypes = [vp, i32, i32, i32, i32]
    lib.pb200_set_trnas.argtypes = [vp, vp, vp, vp, i32]
    lib.pb200_stats.argtypes = [vp, vp]
    lib.pb200_get_orf_int_weights.argtypes = [vp, vp]
    lib.pb200_get_overlap_int_weights.argtypes = [vp, vp]
    lib.pb200_get_gap_int_weights.argtypes = [vp, vp, vp]
    for f in ("pb200_get_calls", "pb200_get_calls24", "pb200_get_contigs", "pb200_get_orfs", "pb200_get_orf_holds", "pb200_get_nodes", "pb200_get_edges", "pb200_get_orf_holds"):
        getattr(lib, f).argtypes = [vp, vp]
    lib.pb200_build_edges.argtypes = [vp]
    lib.pb200_bellman_ford.argtypes = [vp, i32, i32, vp, vp, vp, i32, i32, vp, vp]
    lib.pb200_connect.argtypes = [vp, vp, vp, i32, vp, ctypes.c_int64, vp]
    lib.pb200_stage_times.argtypes = [vp, vp, vp, ctypes.c_int]
    lib.pb200_stage_gaps.argtypes = [vp, vp, vp, ctypes.c_int]
    lib.pb200_launch_count.argtypes = [vp]
    lib.pb200_last_run_ms.argtypes = [vp]
    lib.pb200_last_run_ms.restype = ctypes.c_float
    lib.pb200_device_calls.argtypes = [vp]
    lib.pb200_device_calls.restype = ctypes.c_void_p
    lib.pb200_pin_host.argtypes = [vp, ctypes.c_size_t]
    lib.pb200_unpin_host.argtypes = [vp]
    lib.pb200_struct_sizes.argtypes = [vp]
    lib.pb200_fasta_count.argtypes = [vp, ctypes.c_int64]
    lib.pb200_fasta_count.restype = ctypes.c_int64
    lib.pb200_fasta_parse.argtypes = [vp, ctypes.c_int64, vp, vp, vp, vp, ctypes.c_int64]
    lib.pb200_fasta_parse.restype = ctypes.c_int64
    lib.pb200_format_tabular.argtypes = [vp, vp, i32, vp, vp, vp, ctypes.c_int64]
    lib.pb200_format_tabular.restype = ctypes.c_int64
    lib.pb200_format_score.argtypes = [ctypes.c_double, vp]
    lib.pb200_mark.argtypes = [vp, i32]
    lib.pb200_elapsed_ms.argtypes = [vp, i32, i32]
    lib.pb200_elapsed_ms.restype = ctypes.c_float
    lib.pb200_elapsed_between_ms.argtypes = [vp, i32, vp, i32]
    lib.pb200_elapsed_between_ms.restype = ctypes.c_float
    lib.pb200_comm_unique_id.argtypes = [vp]
    lib.pb200_comm_init.argtypes = [vp, vp, i32, i32]
    lib.pb200_comm_destroy.argtypes = [vp]
    lib.pb200_comm_gather_calls.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.pb200_comm_gather_calls24.argtypes = [vp, vp, vp, i32, vp, vp, vp]
    lib.pb200_comm_fetch_gathered.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, vp]
    lib.pb200_comm_fetch_begin.argtypes = [vp, ctypes.c_int64, ctypes.c_int64, vp]
    lib.pb200_comm_fetch_wait.argtypes = [vp]
    lib.pb200_comm_allreduce.argtypes = [vp, vp, i32, i32]
    lib.pb200_comm_barrier.argtypes = [vp]
    lib.pb200_comm_nccl_version.argtypes = []
    sz = (ctypes.c_int32 * 8)()
    lib.pb200_struct_sizes(sz)
    want = [DEC.itemsize, PARAMS.itemsize, CALL.itemsize, ORF.itemsize, NODE.itemsize, EDGE.itemsize, CONTIG.itemsize]
    if list(sz)[:7] != want:
        raise RuntimeError("phanotate_b200: struct layout mismatch between %s %s and the binding %s" % (p, list(sz)[:7], want))
    return lib


EXPORTS = ["pb200_create", "pb200_destroy", "pb200_last_error", "pb200_run", "pb200_upload", "pb200_upload_packed4", "pb200_upload_async", "pb200_prefetch_async", "pb200_get_calls24", "pb200_comm_gather_calls24", "pb200_pack4", "pb200_set_contig_base", "pb200_set_chunking", "pb200_set_trnas", "pb200_sizes", "pb200_stats",
           "pb200_get_orf_int_weights", "pb200_get_overlap_int_weights", "pb200_get_gap_int_weights", "pb200_get_calls",
           "pb200_get_contigs", "pb200_get_orfs", "pb200_get_orf_holds", "pb200_get_nodes", "pb200_build_edges", "pb200_get_edges",
           "pb200_bellman_ford", "pb200_connect", "pb200_stage_times", "pb200_stage_gaps", "pb200_launch_count", "pb200_last_run_ms",
           "pb200_device_calls", "pb200_pin_host", "pb200_unpin_host", "pb200_struct_sizes", "pb200_fasta_count",
           "pb200_fasta_parse", "pb200_format_tabular", "pb200_comm_unique_id", "pb200_comm_init", "pb200_comm_destroy",
           "pb200_comm_gather_calls", "pb200_comm_fetch_gathered", "pb200_comm_fetch_begin", "pb200_comm_fetch_wait", "pb200_comm_allreduce", "pb200_comm_barrier",
           "pb200_comm_nccl_version", "pb200_mark", "pb200_elapsed_ms", "pb200_elapsed_between_ms", "pb200_format_score"]
