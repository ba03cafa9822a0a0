"""ctypes binding of include/akugpu.h.  Fails loudly when libakugpu.so is missing."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class AkuGpuError(RuntimeError):
    """Raised for any non-zero return code of the C ABI (the reference's SWIG layer maps
    its `throw std::string` to RuntimeError the same way, aku/swig/PPToolbox.i:14-30)."""

    def __init__(self, code, msg):
        super().__init__("akugpu error %d: %s" % (code, msg))
        self.code = code


def library_path():
    return os.path.join(_HERE, "libakugpu.so")


SYMBOLS = {
    # name: (restype, argtypes)
    "akugpu_create": (C.c_void_p, [C.c_int]),
    "akugpu_destroy": (None, [C.c_void_p]),
    "akugpu_last_error": (C.c_char_p, [C.c_void_p]),
    "akugpu_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "akugpu_synchronize": (C.c_int, [C.c_void_p]),
    "akugpu_launch_count": (C.c_int64, [C.c_void_p]),
    "akugpu_stage_times": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "akugpu_stage_times_reset": (C.c_int, [C.c_void_p, C.c_int]),
    "akugpu_frontend_load_config": (C.c_int, [C.c_void_p, C.c_char_p]),
    "akugpu_frontend_load_config_text": (C.c_int, [C.c_void_p, C.c_char_p]),
    "akugpu_frontend_dim": (C.c_int, [C.c_void_p]),
    "akugpu_frontend_sample_rate": (C.c_int, [C.c_void_p]),
    "akugpu_frontend_frame_rate": (C.c_float, [C.c_void_p]),
    "akugpu_frontend_base_is_pre": (C.c_int, [C.c_void_p]),
    "akugpu_frontend_base_dim": (C.c_int, [C.c_void_p]),
    "akugpu_frontend_pre_legacy": (C.c_int, [C.c_void_p]),
    "akugpu_frontend_num_frames": (C.c_int64, [C.c_void_p, C.c_int64]),
    "akugpu_frontend_set_parameters": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p]),
    "akugpu_features": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "akugpu_features_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_char_p,
                                        C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "akugpu_features_pre": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "akugpu_features_pre_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_char_p,
                                            C.c_void_p, C.c_int, C.POINTER(C.c_int)]),
    "akugpu_model_read": (C.c_int, [C.c_void_p, C.c_char_p]),
    "akugpu_model_read_files": (C.c_int, [C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p]),
    "akugpu_model_load_diag": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "akugpu_model_load_full": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p]),
    "akugpu_model_read_clustering": (C.c_int, [C.c_void_p, C.c_char_p]),
    "akugpu_model_set_clustering": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]),
    "akugpu_model_set_clustering_min_evals": (C.c_int, [C.c_void_p, C.c_double, C.c_double]),
    "akugpu_model_use_clustering": (C.c_int, [C.c_void_p, C.c_int]),
    "akugpu_model_set_cmllr": (C.c_int, [C.c_void_p, C.c_void_p]),
    "akugpu_model_set_cmllr_units": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int, C.POINTER(C.c_char_p), C.c_void_p]),
    "akugpu_model_num_states": (C.c_int, [C.c_void_p]),
    "akugpu_model_dim": (C.c_int, [C.c_void_p]),
    "akugpu_model_num_gaussians": (C.c_int, [C.c_void_p]),
    "akugpu_gmm_score": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_void_p]),
    "akugpu_gmm_logprobs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_double, C.c_void_p]),
    "akugpu_gmm_lna": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "akugpu_phone_probs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                     C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64)]),
    "akugpu_phone_probs_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p]),
    "akugpu_checksum_begin": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64]),
    "akugpu_checksum_update": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]),
    "akugpu_checksum_end": (C.c_int, [C.c_void_p, C.c_void_p]),
    "akugpu_shared_alloc": (C.c_int, [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "akugpu_shared_open": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]),
    "akugpu_shared_release": (C.c_int, [C.c_void_p, C.c_void_p]),
    "akugpu_copy_async": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "akugpu_lna_header": (C.c_int, [C.c_int, C.c_int, C.c_void_p]),
    "akugpu_set_chunk_frames": (C.c_int, [C.c_void_p, C.c_int64]),
    "akugpu_set_scorer_variant": (C.c_int, [C.c_void_p, C.c_int]),
    "akugpu_model_expanded_form_q": (C.c_double, [C.c_void_p]),
    "akugpu_scorer_in_use": (C.c_int, [C.c_void_p]),
    "akugpu_set_streaming": (C.c_int, [C.c_void_p, C.c_int]),
    "akugpu_stream_probe": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "akugpu_stream_open": (C.c_int, [C.c_void_p, C.c_double]),
    "akugpu_stream_close": (C.c_int, [C.c_void_p]),
    "akugpu_stream_latency": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_int, C.POINTER(C.c_double)]),
    "akugpu_stream_logprobs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.POINTER(C.POINTER(C.c_float))]),
    "akugpu_stream_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "akugpu_pipe_rates": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
}


def load_library():
    """Loads libakugpu.so (built in-tree by __graft_entry__.build()) and types every symbol
    include/akugpu.h declares.  Loading needs libcudart's dependencies only; no GPU is touched
    until akugpu_create()."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise AkuGpuError(-1, "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C aaltoasr_b200/csrc); there is no CPU fallback" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib
