"""ctypes binding of libmdtb200.so (include/mdtb200.h).  There is NO fallback: if the library cannot be
loaded (or built with nvcc) every entry point of the package raises."""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

ABI_VERSION = 1

# enums of include/mdtb200.h
VARIANT = {"mdtv": 0, "mdt": 1}
PRECISION = {"fp32": 0, "bf16x3": 1, "bf16": 2}
SAMPLER = {"ddim": 0, "euler": 1, "heun": 2, "dpmpp_2m": 3, "euler_ancestral": 4}
MODALITY_VIS, MODALITY_LANG = 0, 1

EXPORTS = [
    "mdtb200_abi_version", "mdtb200_create", "mdtb200_destroy", "mdtb200_last_error",
    "mdtb200_bind_weight", "mdtb200_commit_weights", "mdtb200_encode", "mdtb200_set_context",
    "mdtb200_denoise", "mdtb200_sample", "mdtb200_sample_host", "mdtb200_sample_ancestral", "mdtb200_launch_count",
    "mdtb200_debug_copy", "mdtb200_debug_gemm", "mdtb200_debug_gemm_time", "mdtb200_debug_ktrace",
    "mdtb200_op_gemm", "mdtb200_op_gemm_tc", "mdtb200_op_gemm_tc_scratch", "mdtb200_op_group_sum", "mdtb200_op_colsum", "mdtb200_op_act", "mdtb200_op_ln_fwd", "mdtb200_op_ln_bwd",
    "mdtb200_op_attn_fwd", "mdtb200_op_attn_bwd", "mdtb200_op_gate_res", "mdtb200_op_gate_res_bwd", "mdtb200_op_dropout", "mdtb200_op_adamw_ema", "mdtb200_op_set_seed_epoch",
    "mdtb200_op_split", "mdtb200_op_split_rows_per_slab", "mdtb200_op_split_multi", "mdtb200_op_gemm16_ws", "mdtb200_op_gemm16", "mdtb200_op_ln_fwd16",
    "mdtb200_op_ln_bwd2", "mdtb200_op_ln_bwd2_partials", "mdtb200_op_attn_fwd16", "mdtb200_op_attn_bwd16", "mdtb200_op_res_drop_fwd", "mdtb200_op_res_drop_bwd", "mdtb200_op_narrow_fwd", "mdtb200_op_narrow_wgrad",
    "mdtb200_perceiver_create", "mdtb200_perceiver_destroy", "mdtb200_perceiver_last_error", "mdtb200_perceiver_bind_weight",
    "mdtb200_perceiver_commit_weights", "mdtb200_perceiver_forward", "mdtb200_perceiver_launch_count",
]


class MdtConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("variant", C.c_int32), ("embed_dim", C.c_int32), ("n_heads", C.c_int32),
        ("n_enc_layers", C.c_int32), ("n_dec_layers", C.c_int32), ("action_dim", C.c_int32),
        ("action_seq_len", C.c_int32), ("goal_dim", C.c_int32), ("obs_dim", C.c_int32),
        ("n_state_tokens", C.c_int32), ("precision", C.c_int32), ("max_batch", C.c_int32),
        ("sigma_data", C.c_float),
    ]


class MdtPerceiverConfig(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("abi_version", "dim", "depth", "heads", "dim_head", "num_latents", "num_time_embeds", "ff_mult",
                                         "max_batch", "max_features")]


_lock = threading.Lock()
_lib = None


def _declare(lib):
    vp, i32, i64, fp = C.c_void_p, C.c_int, C.c_int64, C.c_void_p
    lib.mdtb200_abi_version.restype = i32
    lib.mdtb200_create.argtypes = [C.POINTER(MdtConfig), C.POINTER(vp)]
    lib.mdtb200_create.restype = i32
    lib.mdtb200_destroy.argtypes = [vp]
    lib.mdtb200_destroy.restype = None
    lib.mdtb200_last_error.argtypes = [vp]
    lib.mdtb200_last_error.restype = C.c_char_p
    lib.mdtb200_bind_weight.argtypes = [vp, C.c_char_p, fp, i64]
    lib.mdtb200_bind_weight.restype = i32
    lib.mdtb200_commit_weights.argtypes = [vp, vp]
    lib.mdtb200_commit_weights.restype = i32
    lib.mdtb200_encode.argtypes = [vp, fp, fp, i32, i32, fp, vp]
    lib.mdtb200_encode.restype = i32
    lib.mdtb200_set_context.argtypes = [vp, fp, i32, vp]
    lib.mdtb200_set_context.restype = i32
    lib.mdtb200_denoise.argtypes = [vp, fp, fp, i32, i32, fp, vp]
    lib.mdtb200_denoise.restype = i32
    lib.mdtb200_sample.argtypes = [vp, i32, fp, i32, fp, fp, i32, i32, fp, vp]
    lib.mdtb200_sample.restype = i32
    lib.mdtb200_sample_host.argtypes = [vp, i32, fp, i32, fp, fp, i32, i32, fp, vp]
    lib.mdtb200_sample_host.restype = i32
    lib.mdtb200_sample_ancestral.argtypes = [vp, fp, i32, fp, fp, i32, i32, fp, fp, C.c_float, vp]
    lib.mdtb200_sample_ancestral.restype = i32
    lib.mdtb200_launch_count.argtypes = [vp]
    lib.mdtb200_launch_count.restype = i64
    lib.mdtb200_debug_copy.argtypes = [vp, C.c_char_p, fp, i64, vp]
    lib.mdtb200_debug_copy.restype = i64
    lib.mdtb200_debug_gemm.argtypes = [vp, fp, fp, fp, fp, fp, i32, i32, i32, i32, i32, fp, vp]
    lib.mdtb200_debug_gemm.restype = i32
    lib.mdtb200_debug_gemm_time.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(C.c_float), vp]
    lib.mdtb200_debug_gemm_time.restype = i32
    lib.mdtb200_debug_ktrace.argtypes = [vp, i64, vp, i64]
    lib.mdtb200_debug_ktrace.restype = i64
    lib.mdtb200_op_gemm.argtypes = [i32, fp, fp, fp, fp, i32, i32, i32, i32, vp]
    lib.mdtb200_op_gemm_tc.argtypes = [i32, fp, fp, fp, fp, i32, i32, i32, vp, vp]
    lib.mdtb200_op_gemm_tc.restype = i32
    lib.mdtb200_op_gemm_tc_scratch.argtypes = [i32, i32, i32, i32]
    lib.mdtb200_op_gemm_tc_scratch.restype = i64
    lib.mdtb200_op_group_sum.argtypes = [fp, fp, i32, i32, i32, i32, vp]
    lib.mdtb200_op_colsum.argtypes = [fp, fp, fp, i32, i32, i32, vp]
    lib.mdtb200_op_act.argtypes = [fp, fp, fp, i64, i32, vp]
    lib.mdtb200_op_ln_fwd.argtypes = [fp, fp, fp, fp, fp, i32, i32, i32, i32, fp, vp]
    lib.mdtb200_op_ln_bwd.argtypes = [fp, fp, fp, fp, fp, i32, i32, i32, i32, fp, fp, fp, fp, vp]
    lib.mdtb200_op_attn_fwd.argtypes = [fp, i32, fp, fp, i32, fp, i32, i32, i32, i32, i32, i32, i32, C.c_float, C.c_uint64, vp]
    lib.mdtb200_op_attn_bwd.argtypes = [fp, i32, fp, fp, i32, fp, i32, fp, i32, fp, fp, i32, i32, i32, i32, i32, i32, i32, C.c_float, C.c_uint64, vp]
    lib.mdtb200_op_dropout.argtypes = [fp, fp, i64, C.c_float, C.c_uint64, vp]
    lib.mdtb200_op_gate_res.argtypes = [fp, fp, fp, fp, i32, i32, i32, vp]
    lib.mdtb200_op_gate_res_bwd.argtypes = [fp, fp, fp, fp, fp, i32, i32, i32, vp]
    lib.mdtb200_op_adamw_ema.argtypes = [vp, vp, i32] + [C.c_float] * 6 + [i32, vp, vp]
    lib.mdtb200_op_set_seed_epoch.argtypes = [vp]
    f32, u64 = C.c_float, C.c_uint64
    lib.mdtb200_op_split.argtypes = [vp, vp, i32, vp, vp, i32, i32, vp]
    lib.mdtb200_op_split_rows_per_slab.argtypes = []
    lib.mdtb200_op_split_multi.argtypes = [vp, vp, i32, vp]
    lib.mdtb200_op_gemm16_ws.argtypes = [i32, i32, i32]
    lib.mdtb200_op_gemm16_ws.restype = C.c_int64
    lib.mdtb200_op_gemm16.argtypes = [i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp]
    lib.mdtb200_op_ln_fwd16.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, vp]
    lib.mdtb200_op_ln_bwd2.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp, vp, i32, vp, i32, i32, i32, vp]
    lib.mdtb200_op_ln_bwd2_partials.argtypes = [i32, i32]
    lib.mdtb200_op_ln_bwd2_partials.restype = i32
    lib.mdtb200_op_attn_fwd16.argtypes = [vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, i32, i32, f32, u64, vp]
    lib.mdtb200_op_attn_bwd16.argtypes = [vp, i32, vp, vp, i32, vp, i32, vp, vp, i32, vp, i32, i32, vp, i32, i32, i32, vp, i32,
                                          i32, i32, i32, i32, i32, i32, f32, u64, vp]
    lib.mdtb200_op_attn_bwd16.restype = i32
    lib.mdtb200_op_res_drop_fwd.argtypes = [vp, vp, vp, i32, vp, i32, i32, i32, f32, u64, vp]
    lib.mdtb200_op_res_drop_bwd.argtypes = [vp, vp, vp, i32, vp, vp, i32, vp, i32, i32, i32, f32, u64, vp]
    lib.mdtb200_op_narrow_fwd.argtypes = [vp, vp, vp, vp, i32, i32, i32, vp]
    lib.mdtb200_op_narrow_wgrad.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp]
    for fn in ("split", "split_rows_per_slab", "split_multi", "gemm16", "ln_fwd16", "ln_bwd2", "attn_fwd16", "res_drop_fwd", "res_drop_bwd",
               "narrow_fwd", "narrow_wgrad"):
        getattr(lib, "mdtb200_op_" + fn).restype = i32
    lib.mdtb200_op_set_seed_epoch.restype = i32
    lib.mdtb200_op_adamw_ema.restype = i32
    lib.mdtb200_perceiver_create.argtypes = [C.POINTER(MdtPerceiverConfig), C.POINTER(vp)]
    lib.mdtb200_perceiver_create.restype = i32
    lib.mdtb200_perceiver_destroy.argtypes = [vp]
    lib.mdtb200_perceiver_destroy.restype = None
    lib.mdtb200_perceiver_last_error.argtypes = [vp]
    lib.mdtb200_perceiver_last_error.restype = C.c_char_p
    lib.mdtb200_perceiver_bind_weight.argtypes = [vp, C.c_char_p, fp, i64]
    lib.mdtb200_perceiver_bind_weight.restype = i32
    lib.mdtb200_perceiver_commit_weights.argtypes = [vp, vp]
    lib.mdtb200_perceiver_commit_weights.restype = i32
    lib.mdtb200_perceiver_forward.argtypes = [vp, fp, fp, i32, i32, i32, fp, vp]
    lib.mdtb200_perceiver_forward.restype = i32
    lib.mdtb200_perceiver_launch_count.argtypes = [vp]
    lib.mdtb200_perceiver_launch_count.restype = i64
    for _n in ("gemm", "group_sum", "colsum", "act", "ln_fwd", "ln_bwd", "attn_fwd", "attn_bwd", "gate_res", "gate_res_bwd", "dropout"):
        getattr(lib, "mdtb200_op_" + _n).restype = i32
    return lib


def lib_path() -> str:
    return _build.LIB_PATH


def load():
    """Returns the loaded library; builds it with nvcc if the .so is absent.  Raises on failure."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = _build.LIB_PATH
        if _build.is_stale():             # missing, or built from other sources (content hash); build() is multi-process safe
            try:
                _build.build()
            except Exception as e:  # noqa: BLE001
                raise RuntimeError(
                    f"mdt_policy_b200: CUDA library {path} is missing and could not be built ({e}); "
                    "there is no CPU fallback") from e
        try:
            lib = C.CDLL(path)
        except OSError as e:
            raise RuntimeError(f"mdt_policy_b200: cannot load {path}: {e}; there is no CPU fallback") from e
        _declare(lib)
        if lib.mdtb200_abi_version() != ABI_VERSION:
            raise RuntimeError(f"mdt_policy_b200: {path} has ABI {lib.mdtb200_abi_version()}, expected {ABI_VERSION}")
        _lib = lib
        return lib


def current_stream_ptr(device_index) -> C.c_void_p:
    """raw cudaStream_t of torch's current stream on `device_index` (None: the current device).  torch.cuda.current_stream() builds a
    Stream object per call (~7 us; a training step makes 300+ such calls), torch._C._cuda_getCurrentRawStream returns the handle directly."""
    import torch
    idx = torch.cuda.current_device() if device_index is None else device_index
    raw = getattr(torch._C, "_cuda_getCurrentRawStream", None)
    if raw is not None:
        return C.c_void_p(raw(idx))
    return C.c_void_p(torch.cuda.current_stream(idx).cuda_stream)
