"""ctypes loader for libgeneakit_b200.so (the C ABI declared in include/geneakit_b200.h).

The library is the product: if it is missing this module raises ImportError -- there is no
Python or CPU fallback for the compute path.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libgeneakit_b200.so")

GK_OK, GK_ERR_ARG, GK_ERR_INDEX, GK_ERR_CUDA, GK_ERR_OOM, GK_ERR_STATE = 0, 1, 2, 3, 4, 5

_I = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_D = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_F = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


class gk_opts(C.Structure):
    _fields_ = [("struct_size", C.c_int32), ("device", C.c_int32), ("verbose", C.c_int32),
                ("compress", C.c_int32), ("order", C.c_int32), ("rank", C.c_int32),
                ("world", C.c_int32), ("flags", C.c_int32), ("stream", C.c_void_p)]


class gk_phi_stats(C.Structure):
    _fields_ = [("n_cuts", C.c_int32), ("n_steps", C.c_int32), ("compressed", C.c_int32),
                ("ranked", C.c_int32), ("tile", C.c_int32), ("m", C.c_int32),
                ("max_cut", C.c_int64), ("max_classes", C.c_int64),
                ("cells_reference", C.c_int64), ("cells_device", C.c_int64),
                ("algorithmic_bytes_reference", C.c_int64), ("algorithmic_bytes_device", C.c_int64),
                ("device_bytes", C.c_int64), ("kernel_launches", C.c_int64),
                ("plan_seconds", C.c_double), ("plan_bytes", C.c_int64),
                ("panel_row_bytes", C.c_int64), ("corrections", C.c_int64),
                ("rank_row_bytes", C.c_int64), ("rank_remote_row_bytes", C.c_int64)]


class gk_step_view(C.Structure):
    _P = C.POINTER(C.c_int32)
    _fields_ = [("K", C.c_int64), ("Kpad", C.c_int64), ("K_prev", C.c_int64), ("Kpad_prev", C.c_int64),
                ("panel_rows", C.c_int64),
                ("np", C.c_int32), ("panel", C.c_int32), ("ranked", C.c_int32), ("ngroups", C.c_int32),
                ("panel_begin", C.c_int32), ("panel_end", C.c_int32),
                ("pF", _P), ("pM", _P), ("flags", _P), ("cls_ind", _P), ("cls_size", _P),
                ("pg_i", _P), ("pg_j", _P), ("pg_off", _P), ("pt_cls", _P), ("pt_mult", _P)]


class gk_gc_stats(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("total_seconds", C.c_double), ("algorithmic_bytes", C.c_int64),
                ("member_rows", C.c_int64), ("column_blocks", C.c_int64), ("block_columns", C.c_int64),
                ("kernel_launches", C.c_int64), ("n_cuts", C.c_int32), ("reserved", C.c_int32)]


GK_FLAG_DIAGONAL_ONLY = 1


def make_opts(device=-1, verbose=False, compress=-1, order=-1, rank=0, world=1, stream=None, flags=0, gpus=None):
    """gpus=N (one-shot entry points only): this process drives GPUs 0 .. N-1 (rank = -1, world = N)."""
    o = gk_opts()
    o.struct_size = C.sizeof(gk_opts)
    o.device, o.verbose, o.compress, o.order = int(device), int(bool(verbose)), int(compress), int(order)
    o.rank, o.world, o.flags = int(rank), int(world), int(flags)
    if gpus is not None and int(gpus) > 1:
        o.rank, o.world = -1, int(gpus)
    o.stream = stream
    return o


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "geneakit_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    OP = C.POINTER(gk_opts)
    L.gk_last_error.restype = C.c_char_p
    L.gk_abi_version.restype = i32
    L.gk_opts_init.argtypes = [OP]; L.gk_opts_init.restype = None
    L.gk_phi_last_timing.argtypes = [np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")]; L.gk_phi_last_timing.restype = None
    L.gk_host_alloc.argtypes = [C.c_uint64]; L.gk_host_alloc.restype = vp
    L.gk_host_free.argtypes = [vp]
    L.gk_phi_f64.argtypes = [i32, _I, _I, i32, _I, vp, C.POINTER(C.c_double), OP]
    L.gk_gc_f64.argtypes = [i32, _I, _I, i32, _I, i32, _I, vp, OP]
    L.gk_phi_required_gb.argtypes = [i32, _I, _I, i32, _I]; L.gk_phi_required_gb.restype = C.c_double
    L.gk_childless.argtypes = [i32, _I, _I, vp]; L.gk_childless.restype = i32
    L.gk_parentless.argtypes = [i32, _I, _I, vp]; L.gk_parentless.restype = i32
    L.gk_dfs_order.argtypes = [i32, _I, _I, _I]
    L.gk_cut_sizes.argtypes = [i32, _I, _I, i32, _I, _I, i32]; L.gk_cut_sizes.restype = i32
    L.gk_shard_rows.argtypes = [i32, i32, i32, C.POINTER(i64), C.POINTER(i64)]; L.gk_shard_rows.restype = None
    L.gk_phi_plan.argtypes = [i32, _I, _I, i32, _I, OP, C.POINTER(vp)]
    for name in ("gk_phi_upload", "gk_phi_run", "gk_phi_sync"):
        getattr(L, name).argtypes = [vp]
    L.gk_phi_mean.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.gk_phi_copy_out.argtypes = [vp, vp]
    L.gk_phi_copy_rows.argtypes = [vp, i64, i64, vp]
    L.gk_phi_device_result.argtypes = [vp, C.POINTER(vp), C.POINTER(i64)]
    L.gk_phi_get_stats.argtypes = [vp, C.POINTER(gk_phi_stats)]
    L.gk_phi_cut_sizes.argtypes = [vp, _I, i32]
    L.gk_phi_class_sizes.argtypes = [vp, _I, i32]
    L.gk_phi_step_times.argtypes = [vp, _F, i32]
    L.gk_phi_step_bytes.argtypes = [vp, np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS"), i32]
    L.gk_phi_step_view.argtypes = [vp, i32, C.POINTER(gk_step_view)]
    L.gk_phi_final_view.argtypes = [vp, C.POINTER(C.POINTER(C.c_int32)), C.POINTER(C.POINTER(C.c_int32)),
                                    C.POINTER(i64)]
    L.gk_phi_ipc_export.argtypes = [vp, vp]
    L.gk_phi_ipc_open.argtypes = [vp, i32, vp]
    L.gk_phi_stream.argtypes = [vp, C.POINTER(vp)]
    L.gk_phi_set_stream.argtypes = [vp, vp]
    L.gk_phi_local_buffers.argtypes = [vp, C.POINTER(vp), C.POINTER(vp)]
    L.gk_phi_set_peer.argtypes = [vp, i32, vp, vp]
    L.gk_phi_peers_ready.argtypes = [vp]
    L.gk_phi_run_step.argtypes = [vp, i32]
    L.gk_phi_run_pull.argtypes = [vp, i32]
    L.gk_phi_finish.argtypes = [vp]
    L.gk_phi_dvec.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(i64), C.POINTER(i64)]
    L.gk_phi_dvec_copy_from.argtypes = [vp, vp, i32]
    L.gk_phi_free.argtypes = [vp]; L.gk_phi_free.restype = None
    L.gk_f_f64.argtypes = [i32, _I, _I, i32, _I, vp, OP]
    for name in ("gk_phi_replan", "gk_phi_reupload", "gk_phi_pass", "gk_phi_pass_begin", "gk_phi_pass_end", "gk_phi_pass_begin_all"):
        getattr(L, name).argtypes = [vp]
    L.gk_phi_seg_count.argtypes = [vp, C.POINTER(i32)]
    L.gk_phi_seg_ipc_export.argtypes = [vp, i32, vp, C.POINTER(i64)]
    L.gk_phi_seg_ipc_open.argtypes = [vp, i32, i32, vp, i64]
    L.gk_phi_pass_step.argtypes = [vp, i32]
    L.gk_phi_clone_rank.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.gk_phi_diag.argtypes = [vp, vp]
    L.gk_phi_f.argtypes = [vp, vp]
    L.gk_phi_over.argtypes = [vp, C.c_double, C.POINTER(i64)]
    L.gk_phi_over_fetch.argtypes = [vp, vp, vp, vp]
    L.gk_phi_sparse.argtypes = [vp, C.POINTER(i64)]
    L.gk_phi_sparse_fetch.argtypes = [vp, vp, vp, vp]
    L.gk_phi_ci_stats.argtypes = [vp, _I, i32, i32, vp]
    L.gk_phi_debug_prof.argtypes = [vp, np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")]
    L.gk_cache_release.argtypes = []; L.gk_cache_release.restype = None
    L.gk_gc_last_stats.argtypes = [C.POINTER(gk_gc_stats)]
    return L


lib = _load()

EXPORTS = [
    "gk_phi_f64", "gk_gc_f64", "gk_phi_required_gb", "gk_childless", "gk_parentless", "gk_dfs_order",
    "gk_cut_sizes", "gk_shard_rows", "gk_last_error", "gk_abi_version", "gk_host_alloc", "gk_host_free",
    "gk_phi_plan", "gk_phi_upload", "gk_phi_run", "gk_phi_sync", "gk_phi_mean", "gk_phi_copy_out",
    "gk_phi_device_result", "gk_phi_copy_rows", "gk_phi_get_stats", "gk_phi_cut_sizes", "gk_phi_class_sizes",
    "gk_phi_step_times", "gk_phi_step_bytes", "gk_phi_free", "gk_phi_step_view", "gk_phi_final_view",
    "gk_phi_ipc_export", "gk_phi_ipc_open", "gk_phi_local_buffers", "gk_phi_set_peer", "gk_phi_peers_ready",
    "gk_phi_run_step", "gk_phi_run_pull", "gk_phi_finish", "gk_phi_dvec", "gk_phi_dvec_copy_from",
    "gk_phi_debug_prof", "gk_cache_release", "gk_gc_last_stats",
    "gk_opts_init", "gk_phi_last_timing", "gk_f_f64", "gk_phi_stream", "gk_phi_set_stream", "gk_phi_replan", "gk_phi_reupload", "gk_phi_pass", "gk_phi_pass_begin", "gk_phi_pass_step", "gk_phi_pass_end", "gk_phi_pass_begin_all", "gk_phi_seg_count", "gk_phi_seg_ipc_export", "gk_phi_seg_ipc_open", "gk_phi_clone_rank", "gk_phi_diag", "gk_phi_f",
    "gk_phi_over", "gk_phi_over_fetch", "gk_phi_sparse", "gk_phi_sparse_fetch", "gk_phi_ci_stats",
]


class GkError(RuntimeError):
    pass


def check(rc):
    """0 -> ok; otherwise raise the exception type the reference's binding would surface."""
    if rc == GK_OK:
        return
    msg = (lib.gk_last_error() or b"").decode()
    if rc == GK_ERR_INDEX:
        raise IndexError(msg)            # nanobind maps std::out_of_range -> IndexError
    if rc == GK_ERR_OOM:
        raise MemoryError(msg)
    if rc == GK_ERR_ARG:
        raise ValueError(msg)
    raise GkError("geneakit_b200 error %d: %s" % (rc, msg))
