"""ctypes binding of librumi_orb.so (the C ABI in include/rumi_orb.h).

The library is the product: there is no Python / CPU implementation behind these calls.  If the shared object is
missing it is built in-tree with nvcc (rumi_slam_b200/build.py); if that fails, or no CUDA device is usable, every
entry point raises -- nothing falls back to the oracle or to the host.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librumi_orb.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])

# every symbol include/rumi_orb.h declares (tests/test_abi.py checks the header against this table)
_u8p, _f32p, _i32p, _u16p, _u64p = (C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_int32),
                                    C.POINTER(C.c_uint16), C.POINTER(C.c_uint64))
_vp = C.c_void_p
SIGNATURES = {
    "rumi_last_error": (C.c_char_p, []),
    "rumi_device_count": (C.c_int, []),
    "rumi_orb_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "rumi_orb_destroy": (None, [_vp]),
    "rumi_orb_levels": (C.c_int, [_vp]),
    "rumi_orb_tables": (C.c_int, [_vp, _f32p, _f32p, _f32p, _f32p, _i32p]),
    "rumi_orb_frame_capacity": (C.c_int, [_vp, C.c_int, C.c_int]),
    "rumi_orb_extract": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, _vp, _vp, C.c_int,
                                   _i32p, _i32p]),
    "rumi_orb_set_pyramid_staging": (C.c_int, [_vp, C.c_int]),
    "rumi_orb_extract_begin": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int]),
    "rumi_orb_extract_end": (C.c_int, [_vp, _vp, _vp, C.c_int, _i32p, _i32p]),
    "rumi_orb_extract_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, C.c_int,
                                         C.c_int, _vp, _vp, C.c_int, _vp, _vp]),
    "rumi_orb_extract_batch_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t,
                                                C.c_int, C.c_int, _vp, _vp, C.c_int, _vp, _vp, C.c_int]),
    "rumi_orb_wait_stream": (C.c_int, [_vp, _vp]),
    "rumi_orb_signal_stream": (C.c_int, [_vp, _vp]),
    "rumi_match_wait_stream": (C.c_int, [_vp, _vp]),
    "rumi_match_signal_stream": (C.c_int, [_vp, _vp]),
    "rumi_vocab_wait_stream": (C.c_int, [_vp, _vp]),
    "rumi_vocab_signal_stream": (C.c_int, [_vp, _vp]),
    "rumi_nccl_unique_id": (C.c_int, [_vp]),
    "rumi_match_comm_init": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "rumi_match_comm_adopt": (C.c_int, [_vp, _vp, C.c_int, C.c_int]),
    "rumi_hamming_top2_sharded": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
    "rumi_orb_describe": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_size_t, _vp, C.c_int, _vp]),
    "rumi_orb_describe_batch": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_int, C.c_size_t, C.c_size_t, _vp, _vp, _vp]),
    "rumi_orb_pyramid_level": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, _i32p, _i32p]),
    "rumi_orb_blurred_level": (C.c_int, [_vp, C.c_int, _vp, C.c_size_t, _i32p, _i32p]),
    "rumi_orb_debug_candidates": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "rumi_orb_debug_selected": (C.c_int, [_vp, C.c_int, _vp, C.c_int]),
    "rumi_orb_debug_fast_tile": (C.c_int, [_vp, C.c_int, _vp, C.c_int, _vp]),
    "rumi_orb_debug_octree_clocks": (C.c_int, [_vp, _vp, C.c_int]),
    "rumi_orb_set_streams": (C.c_int, [_vp, C.c_int]),
    "rumi_orb_debug_skip_stages": (C.c_int, [_vp, C.c_int]),
    "rumi_orb_timer_start": (C.c_int, [_vp]),
    "rumi_orb_timer_stop": (C.c_int, [_vp, _f32p]),
    "rumi_orb_profile": (C.c_int, [_vp, C.c_int]),
    "rumi_orb_profile_read": (C.c_int, [_vp, _vp, _vp, C.c_int]),
    "rumi_orb_launch_count": (C.c_longlong, [_vp, C.c_int]),
    "rumi_match_timer_start": (C.c_int, [_vp]),
    "rumi_match_timer_stop": (C.c_int, [_vp, _f32p]),
    "rumi_match_launch_count": (C.c_longlong, [_vp, C.c_int]),
    "rumi_match_last_path": (C.c_int, [_vp]),
    "rumi_match_create": (C.c_int, [C.POINTER(_vp), C.c_int]),
    "rumi_match_destroy": (None, [_vp]),
    "rumi_hamming_top2": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]),
    "rumi_hamming_top2_pairs": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp]),
    "rumi_hamming_candidates": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "rumi_hamming_top2_device": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
    "rumi_top2_pack_device": (C.c_int, [_vp, _vp, _vp, _vp, C.c_int, _vp, C.c_int]),
    "rumi_top2_merge_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
    "rumi_stereo_best1": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, _vp, C.c_int, C.c_int, C.c_float,
                                    C.c_float, _vp, _vp]),
    "rumi_stereo_match": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_int, _vp, _vp, C.c_int, C.c_float, C.c_float, _vp, _vp,
                                    _i32p]),
    "rumi_descriptor_distance": (C.c_int, [_vp, _vp]),
    "rumi_vocab_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, _vp, _vp, _vp, _vp]),
    "rumi_vocab_destroy": (None, [_vp]),
    "rumi_vocab_words": (C.c_int, [_vp]),
    "rumi_vocab_launch_count": (C.c_longlong, [_vp, C.c_int]),
    "rumi_bow_transform": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp]),
    "rumi_bow_transform_device": (C.c_int, [_vp, _vp, C.c_int, C.c_int, _vp, _vp, _vp, C.c_int]),
    "rumi_distinctive_descriptors": (C.c_int, [_vp, _vp, _vp, C.c_int, _vp, _vp]),
    "rumi_bow_node_distances": (C.c_int, [_vp, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int, _vp, C.c_int,
                                          _vp, C.c_longlong]),
    "rumi_flow_create": (C.c_int, [C.POINTER(_vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_double, C.c_float]),
    "rumi_flow_destroy": (None, [_vp]),
    "rumi_flow_set_prev": (C.c_int, [_vp, _vp, C.c_int, C.c_int, C.c_size_t]),
    "rumi_flow_track_next": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_int, _vp, _vp, _vp, C.c_int]),
    "rumi_flow_track": (C.c_int, [_vp, _vp, _vp, C.c_int, C.c_int, C.c_size_t, _vp, C.c_int, _vp, _vp, _vp]),
    "rumi_flow_levels": (C.c_int, [_vp]),
    "rumi_flow_launches": (C.c_longlong, [_vp]),
    "rumi_flow_debug_level": (C.c_int, [_vp, C.c_int, C.c_int, _vp, _i32p, _i32p]),
    "rumi_flow_debug_deriv": (C.c_int, [_vp, C.c_int, _vp]),
    "rumi_flow_timer_start": (C.c_int, [_vp]),
    "rumi_flow_timer_stop": (C.c_int, [_vp, _f32p]),
}


class RumiError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("librumi_orb error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    """Loads (building first if needed) librumi_orb.so.  Raises if it cannot be built or loaded."""
    global _lib
    if _lib is None:
        path = os.environ.get("RUMI_ORB_LIB")      # A/B runs against another build of the same ABI
        if not path:
            from . import build as _build
            _build.build()                 # no-op when the .so matches the sources (content hash)
            path = LIB_PATH
        L = C.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)          # AttributeError = the library does not export what the header declares
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        raise RumiError(rc, lib().rumi_last_error().decode("utf-8", "replace"))
    return rc


def torch_stream():
    """cudaStream_t of torch's current stream (the stream the caller's tensors are produced / consumed on)."""
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(a):
    """Raw address of a numpy array / torch tensor / int."""
    if a is None:
        return None
    if isinstance(a, int):
        return a
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()       # torch tensor
