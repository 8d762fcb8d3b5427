"""ctypes binding of the C ABI in include/oflib_b200.h (csrc/liboflib_b200.so).

There is no fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('OFB_LIB_OVERRIDE') or os.path.join(_HERE, 'csrc', 'liboflib_b200.so')   # override: tuning sweeps only

_lib = None

_P = ctypes.c_void_p
_I = ctypes.c_int
_L = ctypes.c_int64
_F = ctypes.c_float
_D = ctypes.c_double

# name -> argtypes, exactly the prototypes of include/oflib_b200.h
SIGNATURES = {
    'ofb_abi_version': ([], _I),
    'ofb_error_string': ([_I], ctypes.c_char_p),
    'ofb_launch_count': ([], _L),
    'ofb_launch_count_reset': ([], None),
    'ofb_set_option': ([_I, _I], _I),
    'ofb_selftest_div': ([_F, ctypes.c_uint32, ctypes.c_uint32, _P, _P], _I),
    'ofb_flow_stats': ([_P, _P, _P, _I, _I, _I, _P], _I),
    'ofb_warp_t_fwd': ([_P, _L, _F, _P, _L, _P, _L, _P, _L, _F, _F, _P, _P, _I, _I, _I, _I, _P], _I),
    'ofb_warp_t_bwd': ([_P, _L, _F, _P, _L, _P, _F, _P, _P, _L, _I, _I, _I, _I, _P], _I),
    'ofb_flow_stats_batch': ([_P, _P, _P, _I, _I, _I, _P], _I),
    'ofb_early_decide': ([_P, _P, _I, _I, _P, _P], _I),
    'ofb_warp_t_flags_ints': ([_I, _I, _I], _L),
    'ofb_warp_t_fwd_ex': ([_P, _L, _F, _P, _L, _P, _L, _P, _L, _F, _F, _P, _P, _I, _I, _I, _I, _P, _P, _P], _I),
    'ofb_warp_t_bwd_ex': ([_P, _L, _F, _P, _L, _P, _F, _P, _P, _L, _I, _I, _I, _I, _P, _P], _I),
    'ofb_warp_t_fwd_d': ([_P, _L, _F, _P, _L, _P, _L, _P, _L, _F, _F, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P], _I),
    'ofb_warp_t_bwd_d': ([_P, _L, _F, _P, _L, _P, _F, _P, _P, _L, _I, _I, _I, _I, _P, _P, _I, _P], _I),
    'ofb_warp_t_inset_fwd': ([_P, _L, _P, _L, _P, _L, _P, _L, _F, _P, _P, _I, _I, _I, _I, _P, _P], _I),
    'ofb_warp_t_inset_bwd': ([_P, _L, _P, _L, _P, _P, _P, _L, _I, _I, _I, _I, _P, _P], _I),
    'ofb_warp_t_win_fwd': ([_P, _L, _F, _P, _L, _P, _L, _P, _L, _F, _F, _P, _P, _I, _I, _I, _I, _P, _P], _I),
    'ofb_warp_t_win_bwd': ([_P, _L, _F, _P, _L, _P, _F, _P, _P, _L, _I, _I, _I, _I, _P, _P], _I),
    'ofb_splat_s_workspace': ([_I, _I, _I, _I, _I], _L),
    'ofb_splat_s_fwd': ([_P, _P, _L, _I, _F, _P, _L, _P, _L, _P, _L, _I, _P, _L, _P, _P, _P, _P, _P, _I,
                         _I, _I, _I, _I, _P], _I),
    'ofb_splat_s_bwd': ([_P, _P, _L, _I, _F, _P, _L, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _L,
                         _I, _I, _I, _I, _P], _I),
    'ofb_splat_s_bwd_workspace': ([_I, _I, _I, _I], _L),
    'ofb_splat_s_win_fwd': ([_P, _P, _L, _F, _P, _L, _P, _L, _P, _L, _I, _P, _L, _P, _P, _P, _P, _P, _I,
                             _I, _I, _I, _I, _P, _P], _I),
    'ofb_splat_s_win_bwd': ([_P, _P, _L, _F, _P, _L, _P, _L, _I, _P, _P, _P, _P, _P, _P, _P, _L, _P, _P, _L,
                             _I, _I, _I, _I, _P, _P], _I),
    'ofb_splat_s_kp': ([_I, _I], _I),
    'ofb_splat_s_band_accumulate': ([_P, _P, _L, _I, _F, _P, _L, _P, _L, _P, _L, _I, _P, _I, _I, _I, _I, _P, _P], _I),
    'ofb_splat_s_band_normalise': ([_P, _P, _L, _I, _P, _L, _P, _L, _P, _L, _I, _P, _P, _P, _P, _P, _P, _I,
                                   _I, _I, _I, _I, _P, _P], _I),
    'ofb_splat_s_kg': ([_I], _I),
    'ofb_splat_s_band_bwd_prepass': ([_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    'ofb_splat_s_band_bwd_gather': ([_P, _P, _L, _F, _P, _L, _P, _L, _I, _P, _P, _P, _P, _P, _L, _P,
                                     _I, _I, _I, _I, _P, _P], _I),
    'ofb_add_rows': ([_P, _L, _P, _L, _I, _L, _P], _I),
    'ofb_sample_pts_fwd': ([_P, _L, _P, _P, _I, _I, _I, _I, _P], _I),
    'ofb_sample_pts_bwd': ([_P, _L, _P, _P, _P, _L, _P, _I, _I, _I, _I, _P], _I),
    'ofb_flow_extent': ([_P, _P, _I, _P, _I, _I, _I, _P], _I),
    'ofb_flo_to_planar': ([_P, _P, _I, _I, _I, _P], _I),
    'ofb_kitti_decode': ([_P, _P, _P, _I, _I, _I, _P], _I),
    'ofb_resize_output_size': ([_I, _D], _I),
    'ofb_resize_fwd': ([_P, _P, _D, _D, _P, _P, _I, _I, _I, _P], _I),
    'ofb_resize_bwd': ([_P, _D, _D, _P, _I, _I, _I, _P], _I),
}


def lib():
    """The loaded shared library.  Raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                "oflibpytorch_b200: the CUDA extension %s is missing. Build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (or oflibpytorch_b200/csrc/build.sh). "
                "There is no CPU or PyTorch fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (argtypes, restype) in SIGNATURES.items():
            if not hasattr(handle, name) and os.environ.get('OFB_SKIP_ABI_TABLE'):
                continue                                  # tuning aid only: A/B against an older build
            fn = getattr(handle, name)
            fn.argtypes = argtypes
            fn.restype = restype
        if handle.ofb_abi_version() != 1:
            raise RuntimeError("oflibpytorch_b200: ABI version mismatch, rebuild the extension")
        _lib = handle
    return _lib


OFB_EINVAL, OFB_ESHAPE = -1, -2     # include/oflib_b200.h


def check(status: int, what: str):
    if status != 0:
        msg = lib().ofb_error_string(status)
        raise RuntimeError("oflibpytorch_b200: %s failed: %s (status %d)" %
                           (what, msg.decode() if msg else '?', status))


def launch_count() -> int:
    return int(lib().ofb_launch_count())


def launch_count_reset():
    lib().ofb_launch_count_reset()
