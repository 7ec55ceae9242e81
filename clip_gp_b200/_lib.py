"""ctypes binding of libclipgp.so (the C ABI declared in include/clipgp.h).

There is no CPU fallback: if the shared library is missing, or a compute entry point is called
without a CUDA device, a RuntimeError is raised (the reference's ``try/except Exception`` blocks
around GP pre-training then behave as they do for any other failure).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libclipgp.so")

KERNEL_IDS = {"rbf": 0, "matern": 1, "linear": 2}

c_f32p = C.c_void_p
c_i64 = C.c_int64


class GpArgs(C.Structure):
    """Mirror of ``clipgp_gp_args`` (include/clipgp.h)."""
    _fields_ = [
        ("kernel_type", C.c_int32), ("x_is_z_prefix", C.c_int32),
        ("C", c_i64), ("T", c_i64), ("n", c_i64), ("d", c_i64), ("S", c_i64),
        ("Z", C.c_void_p), ("X", C.c_void_p),
        ("raw_lengthscale", C.c_void_p), ("raw_outputscale", C.c_void_p), ("raw_variance", C.c_void_p),
        ("var_mean", C.c_void_p), ("chol_var", C.c_void_p), ("mean_x", C.c_void_p),
        ("eps", C.c_void_p), ("eps_sc", c_i64), ("eps_st", c_i64), ("eps_ss", c_i64),
        ("rng_state", C.c_void_p), ("s_offset", c_i64), ("S_total", c_i64),
        ("w", C.c_void_p), ("kl", C.c_void_p), ("L", C.c_void_p), ("A", C.c_void_p), ("R", C.c_void_p),
        ("status", C.c_void_p), ("Ksave", C.c_void_p),
        ("c_begin", c_i64), ("c_count", c_i64),
        ("eps_save", C.c_void_p),
        ("proto_E", C.c_void_p), ("proto_D", c_i64), ("proto_P_hat", C.c_void_p), ("proto_norm", C.c_void_p),
        ("proto_bf16", C.c_void_p), ("proto_bf16_ld", c_i64), ("proto_bf16_seg", c_i64), ("proto_bf16_mode", C.c_int32),
        ("proto_mean_hat", C.c_void_p),
    ]


class GpBwdArgs(C.Structure):
    """Mirror of ``clipgp_gp_bwd_args``."""
    _fields_ = [
        ("dw", C.c_void_p), ("dkl", C.c_void_p), ("dkl_scalar", C.c_float),
        ("dZ_last", C.c_void_p), ("draw_lengthscale", C.c_void_p), ("draw_outputscale", C.c_void_p),
        ("draw_variance", C.c_void_p), ("dvar_mean", C.c_void_p), ("dchol_var", C.c_void_p), ("dmean_x", C.c_void_p),
        ("proto_dP", C.c_void_p), ("proto_dP_stride_s", c_i64), ("proto_dP_scale", C.c_float), ("proto_norm", C.c_void_p),
        ("proto_E", C.c_void_p), ("proto_EEt", C.c_void_p), ("proto_D", c_i64), ("dw_out", C.c_void_p),
        ("tl_Z", C.c_void_p), ("tl_Z_ld", c_i64), ("tl_dlT", C.c_void_p), ("tl_dlT_ld", c_i64), ("tl_seg", c_i64), ("tl_B", c_i64),
        ("tl_mode", C.c_int), ("tl_scale", C.c_float),
    ]


PEER_MAX = 8


class PeerArgs(C.Structure):
    """Mirror of ``clipgp_peer_args`` (include/clipgp.h)."""
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32),
        ("g", C.c_void_p * PEER_MAX), ("p", C.c_void_p * PEER_MAX), ("flags", C.c_void_p * PEER_MAX),
        ("m", C.c_void_p), ("v", C.c_void_p),
        ("n", c_i64), ("n_group0", c_i64),
        ("lr_dev", C.c_void_p),
        ("beta1", C.c_float), ("beta2", C.c_float), ("eps", C.c_float), ("weight_decay", C.c_float),
        ("step", C.c_void_p), ("local", C.c_void_p), ("loss_out", C.c_void_p), ("status", C.c_void_p),
        ("timeout_ns", C.c_uint64),
        ("kl", C.c_void_p), ("kl_n", c_i64), ("kl_scale", C.c_float),
    ]


_SIGNATURES = {
    # name: (restype, argtypes)
    "clipgp_adamw_tail": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_void_p, c_i64, c_i64, c_i64, c_i64, C.c_void_p, C.c_void_p, c_i64, C.c_float, C.c_void_p, C.c_void_p, c_i64,
                                    C.c_void_p, C.c_void_p]),
    "clipgp_transpose_f32": (C.c_int, [C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, c_i64, C.c_void_p]),
    "clipgp_peer_alloc": (C.c_int, [c_i64, C.POINTER(C.c_void_p)]),
    "clipgp_peer_free": (C.c_int, [C.c_void_p]),
    "clipgp_ipc_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "clipgp_ipc_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "clipgp_ipc_close": (C.c_int, [C.c_void_p]),
    "clipgp_peer_adamw": (C.c_int, [C.POINTER(PeerArgs), C.c_void_p]),
    "clipgp_last_error": (C.c_char_p, []),
    "clipgp_version": (C.c_int, []),
    "clipgp_launch_count": (c_i64, []),
    "clipgp_calibration_from_logits": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                 C.c_void_p, C.c_void_p]),
    "clipgp_ece_hist": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p]),
    "clipgp_aece_workspace_bytes": (c_i64, [C.c_int]),
    "clipgp_aece_bins": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                   C.c_void_p, C.c_void_p, c_i64, C.c_void_p]),
    "clipgp_gp_smem_bytes": (c_i64, [c_i64, c_i64, c_i64, C.c_int]),
    "clipgp_gp_forward": (C.c_int, [C.POINTER(GpArgs), C.c_void_p]),
    "clipgp_gp_warp_path_ok": (C.c_int, [c_i64, c_i64, c_i64]),
    "clipgp_gp_fused_proto_ok": (C.c_int, [c_i64, c_i64, c_i64, c_i64, c_i64]),
    "clipgp_gp_fused_proto_bwd_ok": (C.c_int, [c_i64, c_i64, c_i64, c_i64, c_i64]),
    "clipgp_gp_backward": (C.c_int, [C.POINTER(GpArgs), C.POINTER(GpBwdArgs), C.c_void_p]),
    "clipgp_proto_forward": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, c_i64, c_i64, c_i64, C.c_void_p, C.c_float,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                       C.c_void_p]),
    "clipgp_proto_backward": (C.c_int, [C.c_void_p, c_i64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p, c_i64, c_i64, c_i64,
                                        c_i64, C.c_void_p, C.c_void_p]),
    "clipgp_gemm_f32": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, c_i64, c_i64, c_i64,
                                  C.c_float, C.c_int, C.c_void_p]),
    "clipgp_rownorm_forward": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "clipgp_rownorm_cast": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, C.c_void_p, C.c_void_p, c_i64, c_i64, C.c_int, C.c_void_p]),
    "clipgp_rownorm_backward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_i64, c_i64, C.c_void_p, C.c_void_p]),
    "clipgp_softmax_ce": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, C.c_void_p, C.c_float,
                                    C.c_void_p, c_i64, C.c_float, C.c_void_p]),
    "clipgp_l2_identity": (C.c_int, [C.c_void_p, c_i64, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "clipgp_adamw_step": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "clipgp_adamw_step_lrptr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64, C.c_void_p, C.c_float, C.c_float,
                                          C.c_float, C.c_float, C.c_void_p, C.c_void_p]),
    "clipgp_adamw_step_cast": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, c_i64, c_i64, C.c_void_p, C.c_float, C.c_float,
                                         C.c_float, C.c_float, C.c_void_p, C.c_void_p, c_i64, c_i64, C.c_int, C.c_void_p]),
    "clipgp_increment": (C.c_int, [C.c_void_p, c_i64, C.c_void_p]),
    "clipgp_sum_accumulate": (C.c_int, [C.c_void_p, c_i64, C.c_float, C.c_void_p, C.c_void_p]),
    "clipgp_tip_forward": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, c_i64, C.c_float, C.c_float, C.c_void_p, c_i64,
                                     C.c_void_p, c_i64, C.c_int, C.c_void_p]),
    "clipgp_tip_backward": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, C.c_float, C.c_float,
                                      C.c_void_p]),
    "clipgp_tc_tip_logits": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                       c_i64, C.c_void_p]),
    "clipgp_cast_bf16": (C.c_int, [C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, c_i64, c_i64, C.c_int, C.c_void_p]),
    "clipgp_cast_bf16_transpose": (C.c_int, [C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, c_i64, c_i64, C.c_int, C.c_void_p]),
    "clipgp_cast_bf16_dual": (C.c_int, [C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, c_i64, c_i64, C.c_int, C.c_void_p, c_i64, c_i64,
                                        C.c_int, C.c_void_p]),
    "clipgp_softmax_ce_stats": (C.c_int, [C.c_void_p, c_i64, C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, C.c_void_p, C.c_float,
                                          C.c_void_p]),
    "clipgp_softmax_grad_bf16_dual": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, c_i64, c_i64, c_i64, C.c_float, C.c_void_p, c_i64,
                                                c_i64, C.c_int, C.c_void_p, c_i64, c_i64, C.c_int, C.c_void_p]),
    "clipgp_softmax_ce_bf16_dual": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, C.c_float, C.c_float, C.c_void_p,
                                              c_i64, c_i64, C.c_int, C.c_void_p, c_i64, c_i64, C.c_int, C.c_void_p]),
    "clipgp_increment2": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, C.c_void_p]),
    "clipgp_step_epilogue": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, c_i64, c_i64, C.c_void_p, C.c_void_p, c_i64, C.c_void_p]),
    "clipgp_row_sqnorm": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, C.c_void_p]),
    "clipgp_pairdist_radix_hist": (C.c_int, [C.c_void_p, C.c_void_p, c_i64, c_i64, C.c_int, C.c_uint32, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_void_p, C.c_void_p]),
    "clipgp_tc_gemm_store": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, c_i64, C.c_float, C.c_void_p, c_i64,
                                       C.c_void_p]),
    "clipgp_tc_gemm_store_splitk": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, c_i64, C.c_float, C.c_void_p, c_i64,
                                              C.c_void_p]),
    "clipgp_tc_gemm_tf32": (C.c_int, [C.c_void_p, C.c_int, c_i64, C.c_void_p, C.c_int, c_i64, c_i64, C.c_float, C.c_void_p, c_i64, C.c_int,
                                      C.c_void_p]),
    "clipgp_tc_logits_calibration_tf32": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, c_i64, c_i64, C.c_float, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, c_i64, C.c_void_p]),
    "clipgp_tc_logits_calibration": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, c_i64, C.c_float, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                               C.c_void_p, C.c_void_p, C.c_void_p, c_i64, C.c_void_p]),
    "clipgp_tc_proj_logits_calibration": (C.c_int, [C.c_void_p, c_i64, c_i64, C.c_void_p, c_i64, c_i64, c_i64, C.c_float, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                                    C.c_void_p, C.c_void_p, C.c_void_p]),
}

_lib: Optional[C.CDLL] = None


def exported_symbols():
    return sorted(_SIGNATURES)


def load() -> C.CDLL:
    """Load libclipgp.so (built in-tree by ``__graft_entry__.build()``)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback for the clipgp kernels)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in _SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "clipgp") -> None:
    if rc != 0:
        msg = load().clipgp_last_error()
        raise RuntimeError(f"{what} failed (status {rc}): {msg.decode() if msg else 'unknown error'}")


def require_cuda(*tensors: torch.Tensor) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("clipgp kernels need CUDA tensors (sm_100a); there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"clipgp: tensors on different devices ({dev} vs {t.device})")
    if dev is None:
        raise RuntimeError("clipgp: no tensor given")
    return dev


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_ptr(dev: torch.device) -> int:
    return torch.cuda.current_stream(dev).cuda_stream


def launch_count() -> int:
    return int(load().clipgp_launch_count())
