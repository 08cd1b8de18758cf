"""ctypes binding of libpht_b200.so (include/pht_b200.h).

The library is loaded from the package directory (built in-tree by
``pixel_heal_thyself_b200.build``).  There is NO fallback: if the shared object
is missing, importing this module raises, and every op raises ``RuntimeError``
on a non-zero status with the library's error message.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PHT_LIB_PATH") or os.path.join(_PKG, "libpht_b200.so")   # (override: A/B builds)

PHT_F32, PHT_BF16 = 0, 1
PAD_REPLICATE, PAD_REFLECT = 0, 1
EPI_RESID_PRE, EPI_RESID_POST, EPI_MASK, EPI_PADFOLD = 1, 2, 4, 8
EPI_RING1, EPI_RING2, EPI_RING_REFLECT = 16, 32, 64
ABI_VERSION = 3

PAD_MODES = {"replicate": PAD_REPLICATE, "reflect": PAD_REFLECT}
DTYPES = {torch.float32: PHT_F32, torch.bfloat16: PHT_BF16}


class PhtView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("H", C.c_int32), ("W", C.c_int32), ("C", C.c_int32), ("oy", C.c_int32),
                ("ox", C.c_int32), ("dtype", C.c_int32), ("sb", C.c_int64), ("sy", C.c_int64), ("sx", C.c_int64)]


class ConvGemmArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("B", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32), ("N", C.c_int32),
                ("ksize", C.c_int32), ("n_src", C.c_int32), ("flags", C.c_uint32), ("src", PhtView * 3),
                ("w", C.c_void_p), ("bias", C.c_void_p), ("slope", C.c_void_p), ("mslope", C.c_void_p),
                ("resid", PhtView), ("mask", PhtView), ("out1", PhtView), ("out2", PhtView)]


class WgradArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("B", C.c_int32), ("Ho", C.c_int32), ("Wo", C.c_int32), ("N", C.c_int32),
                ("ksize", C.c_int32), ("n_src", C.c_int32), ("dy", PhtView), ("src", PhtView * 3),
                ("dw", C.c_void_p), ("dbias", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t)]


class WgradReduceJob(C.Structure):
    _fields_ = [("partials", C.c_void_p), ("dw", C.c_void_p), ("bias_partials", C.c_void_p), ("dbias", C.c_void_p),
                ("elems", C.c_int64), ("splits", C.c_int32), ("bias_rows", C.c_int32), ("N", C.c_int32), ("pad_", C.c_int32)]


class AttnArgs(C.Structure):
    _fields_ = [("dtype", C.c_int32), ("B", C.c_int32), ("H", C.c_int32), ("W", C.c_int32), ("heads", C.c_int32),
                ("head_dim", C.c_int32), ("block", C.c_int32), ("halo", C.c_int32), ("q", PhtView), ("k", PhtView),
                ("v", PhtView), ("rel_h", C.c_void_p), ("rel_w", C.c_void_p), ("resid", PhtView), ("out", PhtView),
                ("lse", C.c_void_p), ("ring", C.c_int32), ("pad_", C.c_int32)]


class AttnBwdArgs(C.Structure):
    _fields_ = [("fwd", AttnArgs), ("d_out", PhtView), ("dq", PhtView), ("dk", PhtView),
                ("dv", PhtView), ("d_rel_h", C.c_void_p), ("d_rel_w", C.c_void_p), ("workspace", C.c_void_p),
                ("workspace_bytes", C.c_size_t), ("prezeroed", C.c_int32), ("pad_", C.c_int32)]


class PackArgs(C.Structure):
    _fields_ = [("w", C.c_void_p), ("packed", C.c_void_p), ("dtype", C.c_int32), ("O", C.c_int32), ("I", C.c_int32),
                ("ksize", C.c_int32), ("Ntot", C.c_int32), ("Ktot", C.c_int32), ("n_off", C.c_int32),
                ("k_off", C.c_int32), ("transpose", C.c_int32), ("grid", C.c_int32), ("i_begin", C.c_int32),
                ("i_count", C.c_int32), ("scale", C.c_float)]


# every symbol include/pht_b200.h declares: name -> (restype, argtypes)
_vp, _i32, _i64, _f32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_float, C.c_size_t
_PV = C.POINTER(PhtView)
SYMBOLS = {
    "pht_conv_gemm": (C.c_int, [C.POINTER(ConvGemmArgs), _vp]),
    "pht_wgrad_workspace_bytes": (_sz, [C.POINTER(WgradArgs)]),
    "pht_wgrad": (C.c_int, [C.POINTER(WgradArgs), _vp]),
    "pht_border_fill": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "pht_tonemap_u8": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "pht_image_metrics_ws_bytes": (C.c_size_t, [_i32]),
    "pht_image_metrics_u8": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp, C.c_size_t, _vp]),
    "pht_mrse": (C.c_int, [_vp, _vp, _i32, C.c_int64, _i32, _vp, _vp, C.c_size_t, _vp]),
    "pht_film_fwd": (C.c_int, [C.POINTER(PhtView), C.POINTER(PhtView), C.POINTER(PhtView), _i32, _i32, _vp]),
    "pht_film_bwd": (C.c_int, [C.POINTER(PhtView), C.POINTER(PhtView), C.POINTER(PhtView), C.POINTER(PhtView), C.POINTER(PhtView),
                               C.POINTER(PhtView), _i32, _i32, _vp]),
    "pht_pad_fold": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _PV, _PV, _vp, _PV, _PV, _vp]),
    "pht_im2col5": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "pht_attn_fwd": (C.c_int, [C.POINTER(AttnArgs), _vp]),
    "pht_attn_bwd_workspace_bytes": (_sz, [C.POINTER(AttnBwdArgs)]),
    "pht_attn_bwd": (C.c_int, [C.POINTER(AttnBwdArgs), _vp]),
    "pht_attn_bwd_zero": (C.c_int, [C.POINTER(AttnBwdArgs), _vp]),
    "pht_dec_tail_fwd": (C.c_int, [_PV, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pht_dec_tail_bwd_data": (C.c_int, [_vp, _vp, _PV, _PV, _i32, _i32, _i32, _vp]),
    "pht_dec_tail_ws_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "pht_dec_tail_bwd_weight": (C.c_int, [_vp, _PV, _vp, _vp, _vp, _sz, _i32, _i32, _i32, _vp]),
    "pht_l1_loss": (C.c_int, [_vp, _vp, _i64, _f32, _vp, _vp, _vp]),
    "pht_bn_act_ws_bytes": (_sz, [_i32]),
    "pht_colsum_f32": (C.c_int, [_vp, _vp, _i64, _i32, _vp, _sz, _vp]),
    "pht_bn_act_fwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _f32, _vp, _sz, _vp]),
    "pht_bn_act_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _sz, _vp]),
    "pht_bn_act_bwd_bwd": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _f32, _vp, _sz, _vp]),
    "pht_msssim_ws_bytes": (_sz, [_i32, _i32, _i32]),
    "pht_msssim_loss": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _f32, _vp, _vp, _vp, _sz, _vp]),
    "pht_preprocess": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pht_crop_preprocess": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _vp, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "pht_adam": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _i32, _f32, _vp]),
    "pht_adam_dev": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _vp, _vp]),
    "pht_pack_weight": (C.c_int, [C.POINTER(PackArgs), _vp]),
    "pht_unpack_wgrad": (C.c_int, [C.POINTER(PackArgs), _vp]),
    "pht_wgrad_partial": (C.c_int, [C.POINTER(WgradArgs), C.POINTER(WgradReduceJob), _vp]),
    "pht_wgrad_reduce_batched": (C.c_int, [C.POINTER(WgradReduceJob), _i32, _vp, _sz, _i32, _vp]),
    "pht_unpack_wgrads_batched": (C.c_int, [C.POINTER(PackArgs), _i32, _vp, _sz, _i32, _vp]),
    "pht_pack_table_bytes": (_sz, [_i32]),
    "pht_pack_weights_batched": (C.c_int, [C.POINTER(PackArgs), _i32, _vp, _sz, _i32, _vp]),
    "pht_tail_finish": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pht_tail_gather": (C.c_int, [_vp, _i32, _vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pht_tail_im2col_bwd": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _vp]),
    "pht_cast": (C.c_int, [_vp, _i32, _vp, _i32, _i64, _vp]),
    "pht_cast2d": (C.c_int, [_vp, _i32, _i64, _vp, _i32, _i64, _i64, _i64, _vp]),
    "pht_sample_patches": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp]),
    "pht_importance_map_ws_bytes": (_sz, [_i32, _i32, _i32]),
    "pht_importance_map": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _sz, _vp]),
    "pht_importance_sample": (C.c_int, [_vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "pht_abi_version": (C.c_int, []),
    "pht_last_error": (C.c_char_p, []),
    "pht_get_counters": (None, [C.POINTER(C.c_uint64)]),
    "pht_reset_counters": (None, []),
    "pht_add_counters": (None, [C.POINTER(C.c_uint64)]),
    "pht_set_force_simple": (None, [C.c_int]),
    "pht_attn_bwd_trace": (C.c_int, [C.POINTER(C.c_int64), C.c_int32]),
    "pht_conv_gemm_trace": (C.c_int, [C.POINTER(C.c_int64), C.c_int32]),
    "pht_set_option": (C.c_int, [C.c_char_p, C.c_int]),
}


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m pixel_heal_thyself_b200.build` "
            "(the CUDA extension is mandatory, there is no CPU or PyTorch fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.pht_abi_version() != ABI_VERSION:
        raise ImportError(f"libpht_b200.so ABI {lib.pht_abi_version()} != binding ABI {ABI_VERSION}")
    return lib


lib = _load()

# A/B tuning knobs from the environment, e.g. PHT_OPTIONS="tc_cfg=1" (see pht_set_option in include/pht_b200.h)
for _kv in filter(None, os.environ.get("PHT_OPTIONS", "").split(",")):
    _k, _v = _kv.split("=")
    if lib.pht_set_option(_k.strip().encode(), int(_v)) != 0:
        raise ValueError(f"PHT_OPTIONS: unknown option {_k!r}")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.pht_last_error()
        raise RuntimeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def stream_ptr() -> int:
    """cudaStream_t of torch's current stream on the current device (every launch asks: the raw query costs < 1 us, the
    Stream-object route ~16 us, i.e. 2.5 ms of host time per prod step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def require_cuda(*tensors: torch.Tensor) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("pixel_heal_thyself_b200 ops run on CUDA tensors only (no CPU fallback)")


_view_memo: dict[int, tuple] = {}


def view(t: torch.Tensor | None, oy: int = 0, ox: int = 0) -> PhtView:
    """pht_view of a [B, H, W, C] tensor (any pixel strides, channel stride 1).

    The activation arena and its cached slices are the same Python objects every step, so the struct is memoised per
    tensor object (guarded by a weak reference -- ids are recycled -- and by the data pointer): a prod step builds ~400
    views at ~3 us each otherwise."""
    if t is None:
        return PhtView()
    key = id(t)
    ent = _view_memo.get(key)
    ptr = t.data_ptr()
    if ent is not None and ent[0]() is t and ent[1] == ptr and ent[2] == oy and ent[3] == ox:
        return ent[4]
    sh, st = t.shape, t.stride()
    assert len(sh) == 4 and (st[3] == 1 or sh[3] == 1), "view: need channels-last [B,H,W,C]"
    v = PhtView(ptr, sh[1], sh[2], sh[3], oy, ox, DTYPES[t.dtype], st[0], st[1], st[2])
    if len(_view_memo) > 4096:
        _view_memo.clear()
    _view_memo[key] = (weakref.ref(t), ptr, oy, ox, v)
    return v


def ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def raw_counters():
    arr = (C.c_uint64 * 8)()
    lib.pht_get_counters(arr)
    return arr


def counters() -> dict[str, int]:
    arr = (C.c_uint64 * 8)()
    lib.pht_get_counters(arr)
    names = ["gemm_tc", "gemm_simple", "wgrad_tc", "wgrad_simple", "attn_tc", "attn_simple", "other"]
    return {n: int(arr[i]) for i, n in enumerate(names)}
