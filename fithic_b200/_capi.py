"""ctypes binding of libfithic_b200.so (include/fithic_b200.h).

The library is the only compute path: if it is missing or fails to load this module raises -- there is no CPU or
PyTorch fallback.  PyTorch tensors are used purely as device buffers (data_ptr()) and for the CUDA stream handle.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfithic_b200.so")

FHC_OK = 0
FHC_E_INVALID, FHC_E_CUDA, FHC_E_RANGE, FHC_E_WORKSPACE = -1, -2, -3, -4
FHC_ABI_VERSION = 9
(S_INTRA_INRANGE_SUM, S_INTRA_ALL_SUM, S_INTER_ALL_SUM, S_INTER_ALL_COUNT, S_MAX_COUNT, S_OFFGRID,
 S_INTRA_INRANGE_LINES, S_INTRA_ALL_LINES, S_NONPOS_LINES) = range(9)
N_SCALARS = 9
MODE_INTRA_ONLY, MODE_INTER_ONLY, MODE_ALL = 0, 1, 2
BH_CUT_BUCKETS = 32768
MAX_CHR_RUNS = 1024


class StageIO(ctypes.Structure):
    """fhc_stage_io of include/fithic_b200.h (the host stage between K1 and K3)."""
    _fields_ = [
        ("k1buf", c_void_p), ("D", c_int64), ("grid", c_int32), ("noOfBins", c_int32), ("L", c_int64), ("U", c_int64),
        ("chr_n", c_void_p), ("chr_maxmid", c_void_p), ("nchr", c_int32), ("want_spline", c_int32), ("nthreads", c_int32),
        ("n_rank_slots", c_int32), ("dec", c_void_p), ("lbeta_tab", c_void_p * 2), ("lbeta_cap", c_int64 * 2),
        ("dists", c_void_p), ("sums", c_void_p), ("nseen", c_int64),
        ("bin_lb", c_void_p), ("bin_ub", c_void_p), ("bin_sumcc", c_void_p), ("bin_pairs", c_void_p),
        ("bin_sumdist", c_void_p), ("x_bins", c_void_p), ("y_bins", c_void_p), ("xs", c_void_p), ("ys", c_void_p),
        ("t", c_void_p), ("c", c_void_p), ("splineX", c_void_p), ("table", c_void_p), ("lut", c_void_p), ("m", c_int64),
        ("totals", c_int64 * 4), ("lbeta_ntab", c_int64 * 2),
        ("nb", c_int32), ("nt", c_int32), ("ier", c_int32), ("calls", c_int32), ("status", c_int32), ("bad_index", c_int32),
        ("fp", c_double), ("timings", c_double * 8), ("pairs_rank", c_int32), ("pairs_world", c_int32),
        ("shm", c_void_p),
    ]


class FithicB200Error(RuntimeError):
    def __init__(self, code, message):
        super().__init__("libfithic_b200: %s (code %d)" % (message, code))
        self.code = code


_SIGNATURES = {
    # name: (restype, argtypes)
    "fhc_abi_version": (ctypes.c_int, []),
    "fhc_last_error": (c_char_p, []),
    "fhc_launch_count": (c_int64, []),
    "fhc_profile_enable": (ctypes.c_int, [ctypes.c_int]),
    "fhc_profile_collect": (ctypes.c_int, [c_char_p, c_size_t]),
    "fhc_copy_async": (ctypes.c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "fhc_stream_synchronize": (ctypes.c_int, [c_void_p]),
    "fhc_event_create": (ctypes.c_int, [c_void_p]),
    "fhc_event_record": (ctypes.c_int, [c_void_p, c_void_p]),
    "fhc_event_synchronize": (ctypes.c_int, [c_void_p]),
    "fhc_event_destroy": (ctypes.c_int, [c_void_p]),
    "fhc_comm_handle_bytes": (c_int64, []),
    "fhc_comm_create": (ctypes.c_int, [c_int32, c_int32, c_int64, c_void_p, c_void_p]),
    "fhc_comm_connect": (ctypes.c_int, [c_void_p, c_void_p]),
    "fhc_comm_allreduce_u64": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    "fhc_comm_allgather": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "fhc_comm_world": (c_int32, [c_void_p]),
    "fhc_comm_rank": (c_int32, [c_void_p]),
    "fhc_comm_failed": (ctypes.c_int, [c_void_p]),
    "fhc_comm_destroy": (ctypes.c_int, [c_void_p]),
    "fhc_peak_fp64": (ctypes.c_int, [c_double, c_void_p, c_void_p, c_void_p]),
    "fhc_hist_distance": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p,
                                          c_int64, c_int64, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_int64, c_void_p,
                                          c_int32, c_int32, c_void_p]),
    "fhc_mid_range": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "fhc_host_make_bins": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_void_p]),
    "fhc_host_frag_pairs": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int64, c_int64, c_void_p, c_void_p,
                                            c_int32, c_void_p, c_void_p, c_void_p]),
    "fhc_host_frag_pairs_varsize": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p, c_int32,
                                                    c_void_p, c_void_p, c_void_p, c_void_p]),
    "fhc_frag_pairs_varsize": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p, c_int32,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fhc_host_frag_pairs_varsize_prefix": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p,
                                                           c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fhc_host_fill_f64": (ctypes.c_int, [c_void_p, c_int64, c_double, c_int32]),
    "fhc_host_lbeta_table": (ctypes.c_int, [c_int64, c_void_p, c_int64, c_int32]),
    "fhc_host_stage": (ctypes.c_int, [ctypes.POINTER(StageIO), c_int32]),
    "fhc_shm_open": (ctypes.c_int, [c_char_p, c_int32, c_int32, c_int64, c_void_p]),
    "fhc_shm_allreduce_u64": (ctypes.c_int, [c_void_p, c_void_p, c_int32]),
    "fhc_shm_close": (ctypes.c_int, [c_void_p]),
    "fhc_host_pool_prewarm": (ctypes.c_int, [c_int32]),
    "fhc_host_pool_selftest": (c_double, [c_int32, c_int32, c_int32, c_void_p]),
    "fhc_host_curfit": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_double, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_void_p, c_void_p]),
    "fhc_spline_workspace_bytes": (c_size_t, [c_int64]),
    "fhc_spline_table": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_double, c_double, c_int32,
                                         c_void_p, c_void_p, c_int64, c_void_p, c_size_t, c_void_p]),
    "fhc_spline_eval": (ctypes.c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_int64, c_void_p, c_void_p]),
    "fhc_host_antitonic": (ctypes.c_int, [c_void_p, c_int64]),
    "fhc_spline_lut": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_double, c_double, c_int32, c_void_p, c_int64,
                                       c_void_p]),
    "fhc_lbeta_table": (ctypes.c_int, [c_int64, c_void_p, c_int64, c_void_p]),
    "fhc_host_log_cr": (c_double, [c_double]),
    "fhc_host_lbeta": (c_double, [c_double, c_double]),
    "fhc_host_bdtrc_lists": (c_double, [c_int32, c_int64, c_double]),
    "fhc_host_one_minus_exp": (c_double, [c_double]),
    "fhc_host_tail_sum": (None, [c_int32, c_int64, c_double, c_int32, c_void_p, c_void_p]),
    "fhc_pvalues": (ctypes.c_int, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64,
                                    c_void_p, c_void_p,
                                    c_void_p, c_int32, c_int32, c_int32, c_int64, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                    c_double, c_double, c_double, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                    c_int64, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                    c_void_p]),
    "fhc_pvalues_prepass": (ctypes.c_int, [c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64,
                                            c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int64, c_int64,
                                            c_double, c_double, c_int64, c_void_p, c_void_p, c_void_p]),
    "fhc_pvalues_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "fhc_bdtrc": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p]),
    "fhc_bh_workspace_bytes": (c_size_t, [c_int64]),
    "fhc_bh_qvalues": (ctypes.c_int, [c_void_p, c_int64, c_double, c_int64, c_double, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "fhc_bh_qvalues_hostcount": (ctypes.c_int, [c_void_p, c_int64, c_double, c_int64, c_double, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_void_p, c_size_t, c_int32, c_void_p]),
    "fhc_fill_f64": (ctypes.c_int, [c_void_p, c_int64, c_double, c_void_p]),
    "fhc_bh_prepare": (ctypes.c_int, [c_void_p, c_int64, c_double, c_int64, c_double, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "fhc_bh_p_cut": (c_double, [c_double, c_double]),
    "fhc_bh_cut_hist": (ctypes.c_int, [c_void_p, c_int64, c_double, c_void_p, c_void_p]),
    "fhc_bh_cut_from_hists": (ctypes.c_int, [c_void_p, c_int32, c_int32, c_double, c_double, c_void_p, c_void_p]),
    "fhc_bh_dist_cut": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_double, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fhc_host_bh_cut_find": (c_double, [c_void_p, c_double, c_double, c_double]),
    "fhc_host_bh_cut_bucket": (c_int32, [c_double]),
    "fhc_bh_finish": (ctypes.c_int, [c_int64, c_double, c_int64, c_double, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fhc_bh_sample_keys": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_double, c_void_p, c_void_p]),
    "fhc_bh_key_of": (ctypes.c_uint64, [c_double]),
    "fhc_bh_partition_count": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int32, c_double, c_void_p, c_void_p]),
    "fhc_bh_partition_scatter": (ctypes.c_int, [c_void_p, c_int64, c_void_p, c_int32, c_double, c_void_p, c_void_p,
                                                 c_void_p, c_void_p, c_int32, c_void_p]),
    "fhc_scatter_f64": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "fhc_gather_ne_one": (ctypes.c_int, [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fhc_sort_workspace_bytes": (c_size_t, [c_int64]),
    "fhc_sort_pairs_u64": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_size_t,
                                           c_void_p]),
    "fhc_merge_workspace_bytes": (c_size_t, [c_int64]),
    "fhc_merge_components": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                             c_void_p]),
    "fhc_merge_select": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32,
                                         c_int32, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "fhc_host_merge_components": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p,
                                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fhc_host_merge_select": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32,
                                              c_int32, c_int32, c_void_p, c_void_p]),
    "fhc_kr_partials": (c_int32, []),
    "fhc_kr_spmv": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "fhc_kr_mul": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "fhc_kr_residual": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "fhc_kr_first": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "fhc_kr_direction": (ctypes.c_int, [c_void_p, c_double, c_int32, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "fhc_kr_w": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]),
    "fhc_kr_ynew_minmax": (ctypes.c_int, [c_void_p, c_double, c_void_p, c_int64, c_void_p, c_void_p]),
    "fhc_kr_gamma": (ctypes.c_int, [c_void_p, c_double, c_void_p, c_double, c_int32, c_int64, c_void_p, c_void_p]),
    "fhc_kr_axpy": (ctypes.c_int, [c_void_p, c_double, c_double, c_void_p, c_int64, c_void_p]),
    "fhc_kr_update": (ctypes.c_int, [c_void_p, c_double, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p,
                                      c_void_p]),
    "fhc_io_format_double": (ctypes.c_int, [c_double, ctypes.c_int, c_char_p]),
    "fhc_io_read_contacts": (c_void_p, [c_char_p]),
    "fhc_io_contacts_n": (c_int64, [c_void_p]),
    "fhc_io_contacts_nchrom": (c_int32, [c_void_p]),
    "fhc_io_contacts_chrom": (c_char_p, [c_void_p, c_int32]),
    "fhc_io_contacts_copy": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "fhc_io_free": (None, [c_void_p]),
    "fhc_io_write_significances": (c_int64, [c_char_p, ctypes.POINTER(c_char_p), c_int32, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int64, c_int64,
                                             c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32]),
    "fhc_digest_lines": (ctypes.c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "fhc_outlier_bin_decrements": (ctypes.c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_void_p,
                                                   c_void_p]),
}

_lib = None


def load():
    """Load libfithic_b200.so, bind every symbol of include/fithic_b200.h, check the ABI version."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "fithic_b200: %s is missing -- build it with `python -m fithic_b200.build` (needs nvcc, sm_100a). "
            "There is no CPU fallback." % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    if lib.fhc_abi_version() != FHC_ABI_VERSION:
        raise ImportError("fithic_b200: ABI version mismatch (library %d, binding %d)" %
                          (lib.fhc_abi_version(), FHC_ABI_VERSION))
    _lib = lib
    return lib


def check(rc):
    """Turn a negative status code into an exception carrying fhc_last_error()."""
    if rc < 0:
        raise FithicB200Error(rc, load().fhc_last_error().decode("utf-8", "replace"))
    return rc


def launch_count():
    return int(load().fhc_launch_count())


def dptr(t):
    """Device (or host numpy) pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return c_void_p(t.data_ptr())
    return c_void_p(t.ctypes.data)


def profile_enable(on=True):
    check(load().fhc_profile_enable(1 if on else 0))


def profile_collect():
    """{kernel name: {"ms": total device ms, "launches": n}} since the last collect (synchronises the device)."""
    import json
    buf = ctypes.create_string_buffer(1 << 16)
    check(load().fhc_profile_collect(buf, len(buf)))
    return json.loads(buf.value.decode())
