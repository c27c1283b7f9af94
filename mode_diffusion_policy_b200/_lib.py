"""ctypes binding of libmode_engine.so (the C ABI in include/mode_engine.h).

There is deliberately no fallback: if the shared library is missing or fails to load, importing the engine raises.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_DIR = PKG_DIR.parent
LIB_PATH = PKG_DIR / "lib" / "libmode_engine.so"
CSRC = PKG_DIR / "csrc"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # no implicit multiply-add contraction: the fp32 arithmetic of the row kernels is exactly what the source says (explicit
    # fmaf() calls stay fused), so it does not change with inlining context — the persistent small-batch kernel and the
    # per-phase kernels round identically — and it is the arithmetic of the numpy oracle, which never fuses
    "-fmad=false",
    "-shared", "-Xcompiler", "-fPIC",
]


class mode_config_t(C.Structure):
    """Mirror of `mode_config_t` (include/mode_engine.h)."""

    _fields_ = [
        ("obs_dim", C.c_int32),
        ("goal_dim", C.c_int32),
        ("action_dim", C.c_int32),
        ("embed_dim", C.c_int32),
        ("n_layers", C.c_int32),
        ("n_heads", C.c_int32),
        ("n_state_tokens", C.c_int32),
        ("action_seq_len", C.c_int32),
        ("num_experts", C.c_int32),
        ("top_k", C.c_int32),
        ("router_normalize", C.c_int32),
        ("max_batch", C.c_int32),
        ("sigma_data", C.c_float),
        ("rms_eps", C.c_float),
    ]


# every symbol include/mode_engine.h declares: name -> (restype, argtypes)
_P = C.c_void_p
_F = C.c_void_p  # device/host float pointers are passed as raw addresses
SYMBOLS = {
    "mode_create": (C.c_int, [C.POINTER(mode_config_t), C.POINTER(_P)]),
    "mode_destroy": (None, [_P]),
    "mode_last_error": (C.c_char_p, []),
    "mode_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "mode_finalize_weights": (C.c_int, [_P]),
    "mode_set_weight_on_stream": (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.POINTER(C.c_int64), C.c_int, _P]),
    "mode_finalize_weights_on_stream": (C.c_int, [_P, _P]),
    "mode_forward": (C.c_int, [_P, _F, _F, _F, _F, C.c_int, _F, C.c_int, _P]),
    "mode_denoise": (C.c_int, [_P, _F, _F, _F, _F, C.c_int, _F, C.c_int, _P]),
    "mode_profile_eval": (C.c_int, [_P, _F, _F, _F, _F, C.c_int, _F, C.c_int, C.c_int, _P, _P, _P]),
    "mode_loss": (C.c_int, [_P, _F, _F, _F, _F, _F, _F, _F, C.c_int, _P]),
    "mode_train_step": (C.c_int, [_P, _F, _F, _F, _F, _F, _F, _F, C.c_int, _P]),
    "mode_train_set_stochastic": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_ulonglong, C.c_uint]),
    "mode_train_get_token_routing": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "mode_grad_buffer": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "mode_grad_offset": (C.c_int, [_P, C.c_char_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "mode_train_wait_grads": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "mode_optimizer_bind": (C.c_int, [_P, C.c_char_p, C.c_void_p, C.c_int]),
    "mode_optimizer_unbind_all": (C.c_int, [_P]),
    "mode_adamw_step": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_void_p]),
    "mode_optimizer_set_sharding": (C.c_int, [_P, C.c_int, C.c_int]),
    "mode_optimizer_shard_tensors": (C.c_int, [_P, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_int, C.POINTER(C.c_int)]),
    "mode_optimizer_staging": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "mode_optimizer_pack_group": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "mode_weights_record_ready": (C.c_int, [_P, C.c_int, C.c_void_p]),
    "mode_adamw_step_group": (C.c_int, [_P, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p]),
    "mode_optimizer_set_ema": (C.c_int, [_P, C.c_double]),
    "mode_optimizer_ema_state": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "mode_optimizer_ema_mark_seeded": (C.c_int, [_P]),
    "mode_grad_segment_sumsq": (C.c_int, [_P, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]),
    "mode_optimizer_state": (C.c_int, [_P, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]),
    "mode_train_input_grads": (C.c_int, [_P, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "mode_sample": (C.c_int, [_P, C.c_int, _F, _F, _F, C.POINTER(C.c_float), C.c_int, C.c_int, _P]),
    "mode_sample_program": (C.c_int, [_P, _F, _F, _F, C.POINTER(C.c_float), C.POINTER(C.c_int32), C.POINTER(C.c_float), _F,
                                      C.c_int, C.c_int, _P]),
    "mode_sample_ddim": (C.c_int, [_P, _F, _F, _F, C.POINTER(C.c_float), C.c_int, C.c_int, _P]),
    "mode_sample_ddim_host": (C.c_int, [_P, _F, _F, _F, C.POINTER(C.c_float), C.c_int, C.c_int, _P]),
    "mode_block_forward": (C.c_int, [_P, C.c_int, _F, _F, _F, C.c_int, _P]),
    "mode_get_routing": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P]),
    "mode_get_routing_at": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "mode_get_expert_usage": (C.c_int, [_P, C.c_int, _P, _P]),
    "mode_reset_expert_usage": (C.c_int, [_P]),
    "mode_last_launch_count": (C.c_int64, [_P]),
    "mode_resnet_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "mode_resnet_destroy": (None, [_P]),
    "mode_resnet_set_weight": (C.c_int, [_P, C.c_char_p, _P, C.c_int, C.POINTER(C.c_int64), C.c_int]),
    "mode_resnet_finalize": (C.c_int, [_P, _P]),
    "mode_resnet_forward": (C.c_int, [_P, _F, _F, _F, C.c_int, _P]),
    "mode_debug_gemm": (C.c_int, [_F, _F, _F, _F, _F, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mode_debug_wgrad": (C.c_int, [_F, _F, _F, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "mode_debug_attention": (C.c_int, [_F, _F, _F, _F, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, _P]),
}


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile csrc/engine.cu for sm_100a into lib/libmode_engine.so (nvcc cross-compiles without a GPU)."""
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cuh")) + [REPO_DIR / "include" / "mode_engine.h"]
    if not force and LIB_PATH.exists():
        newest = max(p.stat().st_mtime for p in srcs)
        if LIB_PATH.stat().st_mtime >= newest:
            return LIB_PATH
    LIB_PATH.parent.mkdir(parents=True, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), str(CSRC / "engine.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    subprocess.run(cmd, check=True)
    return LIB_PATH


_lib = None


def load() -> C.CDLL:
    """Load the shared library and attach signatures. Raises if it is missing: there is no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The MoDE engine has no CPU or PyTorch fallback."
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class ModeError(RuntimeError):
    pass


def check(rc: int) -> None:
    if rc != 0:
        msg = load().mode_last_error()
        raise ModeError(f"libmode_engine error {rc}: {msg.decode() if msg else '?'}")
