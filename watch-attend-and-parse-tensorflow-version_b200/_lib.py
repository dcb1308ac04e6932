"""ctypes binding of libwap_b200.so (C ABI declared in include/wap_abi.h).

There is NO fallback: if the shared library is missing or does not export a declared symbol,
importing the engine fails loudly.  Build it with `python __graft_entry__.py build` (or
`watch-attend-and-parse-tensorflow-version_b200/csrc/build.sh`).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WAP_B200_LIB") or os.path.join(_HERE, "libwap_b200.so")     # env override: debugging aid (A/B builds)
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "wap_abi.h")

_vp, _i, _i64, _u32, _f = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_float
_ip = C.POINTER(C.c_int)

# name -> (restype, argtypes); must list every function of include/wap_abi.h (tests check this)
SIGNATURES = {
    "wap_abi_version": (_i, []),
    "wap_last_error": (C.c_char_p, []),
    "wap_create": (_i, [C.POINTER(_vp)]),
    "wap_destroy": (_i, [_vp]),
    "wap_make_current": (_i, [_vp]),
    "wap_current": (_vp, []),
    "wap_prepare_batch": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "wap_param_count": (_i, []),
    "wap_param_info": (_i, [_i, C.c_char_p, _i, C.POINTER(_i64), _ip, C.POINTER(_i64), _ip]),
    "wap_param_floats": (_i64, []),
    "wap_decay_floats": (_i64, []),
    "wap_moving_floats": (_i64, []),
    "wap_annotation_hw": (_i, [_i, _i, _ip, _ip]),
    "wap_mask_subsample": (_i, [_vp, _vp, _i, _i, _i, _vp]),
    "wap_encoder_workspace_floats": (_i64, [_i, _i, _i, _i]),
    "wap_encoder_forward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u32, _vp, _vp]),
    "wap_encoder_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _u32, _vp, _vp, _vp]),
    "wap_attn_precompute": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "wap_decoder_step_workspace_floats": (_i64, [_i, _i, _i]),
    "wap_decoder_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wap_one_time_step": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wap_get_context": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "wap_greedy_workspace_floats": (_i64, [_i, _i, _i]),
    "wap_greedy_decode": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _ip]),
    "wap_greedy_persistent_workspace_floats": (_i64, [_i, _i, _i]),
    "wap_greedy_decode_persistent": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "wap_beam_decode_persistent": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "wap_decoder_train_workspace_floats": (_i64, [_i, _i, _i, _i]),
    "wap_decoder_train_forward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u32, _vp, _vp]),
    "wap_decoder_train_backward": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _u32, _vp, _vp]),
    "wap_train_backward": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _i, C.c_uint32, _vp, _vp, _vp]),
    "wap_adadelta_l2_step": (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i64, _f, _f, _f, _f, _f, _vp]),
    "wap_launch_count": (C.c_uint64, []),
    "wap_profile_enable": (_i, [_i]),
    "wap_profile_read": (_i, [_i, C.POINTER(C.c_double), C.POINTER(_i64), C.POINTER(C.c_double)]),
    "wap_profile_read_bytes": (_i, [_i, _vp]),
    "wap_profile_reset": (_i, []),
    "wap_gemm_nn": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i]),
    "wap_gemm_nt": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i]),
    "wap_gemm_tn_atomic": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i]),
    "wap_tc_available": (_i, []),
    "wap_set_conv_precision": (_i, [_i]),
    "wap_get_conv_precision": (_i, []),
    "wap_debug_bn_gates": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "wap_debug_set_tc_mask": (_i, [_i]),
    "wap_debug_capture": (_i, [_i, _vp]),
    "wap_debug_wgrad": (_i, [_vp, _i]),
    "wap_tc_gemm": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp]),
    "wap_tc_gemm_ex": (_i, [_vp, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "wap_tc_conv3x3": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _vp]),
    "wap_tc_conv3x3_rowtile": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp]),
    "wap_tc_wgrad_1x1": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i, _i]),
    "wap_tc_wgrad_3x3": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _i]),
    "wap_tc_gemm_bf16": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "wap_tc_conv3x3_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp]),
    "wap_tc_conv3x3_rowtile_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "wap_tc_wgrad_1x1_bf16": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _vp, _i]),
    "wap_tc_wgrad_3x3_bf16": (_i, [_vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _i]),
    "wap_set_dropout_seed_source": (_i, [_vp]),
    "wap_set_fused_coverage_filter": (_i, [_vp]),
    "wap_set_pdl": (_i, [_i]),
    "wap_get_pdl": (_i, []),
    "wap_debug_tc_trace": (_i, [_vp]),
    "wap_debug_trap_info": (_i, [_vp]),
    "wap_debug_set_fuse_gates": (_i, [_i]),
    "wap_debug_set_greedy_impl": (_i, [_i]),
    "wap_set_overlap": (_i, [_i]),
    "wap_get_overlap": (_i, []),
    "wap_dropout_mask": (_i, [_vp, _u32, _u32, _i64, _vp]),
    "wap_encoder_workspace_offset": (_i64, [_i, _i, _i, _i, _i, _i]),
}


class WapError(RuntimeError):
    pass


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into the in-tree shared library."""
    script = os.path.join(_HERE, "csrc", "build.sh")
    out = subprocess.run(["bash", script], capture_output=True, text=True)
    if verbose or out.returncode != 0:
        print(out.stdout + out.stderr)
    if out.returncode != 0:
        raise WapError("building libwap_b200.so failed")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise WapError(f"{LIB_PATH} is missing: the CUDA extension has not been built "
                           f"(run `python __graft_entry__.py build`). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            try:
                fn = getattr(L, name)
            except AttributeError as e:
                raise WapError(f"libwap_b200.so does not export {name}") from e
            fn.restype, fn.argtypes = res, args
        if L.wap_abi_version() != 1:
            raise WapError("libwap_b200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise WapError(f"{what} failed ({rc}): {lib().wap_last_error().decode(errors='replace')}")


def param_table():
    """[(name, offset, shape, flags)] from the library's own layout."""
    L = lib()
    out = []
    name = C.create_string_buffer(128)
    off, nd, fl = _i64(), _i(), _i()
    dims = (_i64 * 4)()
    for i in range(L.wap_param_count()):
        check(L.wap_param_info(i, name, 128, C.byref(off), C.byref(nd), dims, C.byref(fl)), "wap_param_info")
        out.append((name.value.decode(), int(off.value), tuple(int(dims[j]) for j in range(nd.value)), int(fl.value)))
    return out
