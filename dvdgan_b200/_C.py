"""ctypes binding of include/dvdgan_b200.h (the C ABI of libdvdgan_b200.so).

There is no fallback: if the CUDA library has not been built (``python -m dvdgan_b200.build`` or
``__graft_entry__.build()``) every op raises.  PyTorch is used only for device memory and streams.
"""
import ctypes
import os
from ctypes import c_float, c_int, c_int64, c_size_t, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libdvdgan_b200.so")
_lib = None
LAUNCHES = 0          # C-ABI calls issued (each enqueues >= 1 kernel); bench.py reports it


class ConvDesc(ctypes.Structure):
    _fields_ = [("N1", c_int), ("N2", c_int), ("Cin", c_int), ("Cout", c_int),
                ("D", c_int), ("H", c_int), ("W", c_int), ("kD", c_int), ("kH", c_int), ("kW", c_int),
                ("x_s1", c_int64), ("x_s2", c_int64), ("x_cs", c_int64),
                ("y_s1", c_int64), ("y_s2", c_int64), ("y_cs", c_int64),
                ("in_relu", c_int), ("in_up", c_int), ("accumulate", c_int), ("out_act", c_int), ("res_up", c_int),
                ("r_s1", c_int64), ("r_s2", c_int64), ("r_cs", c_int64), ("x_kind", c_int)]


P, I, L, F, Z = c_void_p, c_int, c_int64, c_float, c_size_t
_SIGS = {
    "dvd_abi_version": (c_int, []),
    "dvd_launch_count": (ctypes.c_longlong, []),
    "dvd_prof_enable": (I, [I]),
    "dvd_prof_dump": (I, [ctypes.c_char_p]),
    "dvd_prof_read": (I, [I, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                          ctypes.POINTER(ctypes.c_longlong)]),
    "dvd_set_option": (I, [ctypes.c_char_p, I]),
    "dvd_get_option": (I, [ctypes.c_char_p, ctypes.POINTER(c_int)]),
    "dvd_saturation_count": (I, [ctypes.POINTER(ctypes.c_uint), I, P]),
    "dvd_scratch_bytes": (I, [ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong)]),
    "dvd_conv_fwd": (I, [ctypes.POINTER(ConvDesc), P, P, P, P, P, P]),
    "dvd_conv_wgrad": (I, [ctypes.POINTER(ConvDesc), P, P, P, P]),
    "dvd_weight_pack": (I, [P, I, I, I, I, I, I, P, I, P, I, I, I, I, P]),
    "dvd_weight_unpack": (I, [P, I, I, I, I, I, I, I, I, I, P, P]),
    "dvd_bgemm": (I, [I, I, I, I, I, F, P, I, L, P, I, L, F, P, I, L, I, P, P]),
    "dvd_convgru_layer_workspace_bytes": (Z, [I, I, I, I, I, I, I]),
    "dvd_convgru_layer_fwd": (I, [P, L, L, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, P, Z, P]),
    "dvd_convgru_layer_bwd": (I, [P, L, L, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, P, Z, P]),
    "dvd_convgru_layer_range_workspace_bytes": (Z, [I, I, I, I, I, I, I]),
    "dvd_convgru_layer_fwd_range": (I, [P, L, L, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, I, P, Z, P]),
    "dvd_convgru_layer_bwd_range": (I, [P, L, L, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I, I, I, I, I, I, I, I, I,
                                        P, Z, P]),
    "dvd_specnorm_fwd": (I, [P, I, I, P, P, P, P, P]),
    "dvd_specnorm_bwd": (I, [P, P, P, P, P, I, I, P, I, P, P]),
    "dvd_bn_stats": (I, [P, I, I, I, I, F, F, P, P, P, P, P, P, P]),
    "dvd_bn_stats_ex": (I, [P, I, I, I, I, F, F, P, P, P, P, P, P, I, I, P]),
    "dvd_cbn_bwd_ex": (I, [P, P, I, P, P, P, I, I, I, I, I, I, I, P, P, P, I, P]),
    "dvd_cbn_apply": (I, [P, P, I, P, P, I, I, I, I, I, I, P, P]),
    "dvd_cbn_bwd": (I, [P, P, I, P, P, P, I, I, I, I, I, I, I, P, P, P, P]),
    "dvd_attn_fwd": (I, [P, L, P, L, P, L, P, P, L, I, I, I, I, I, I, P]),
    "dvd_attn_bwd": (I, [P, L, P, L, P, L, P, P, P, L, P, L, P, L, P, L, I, I, I, I, I, I, P]),
    "dvd_attn_flash_supported": (I, [I, I, I, I, I, I]),
    "dvd_attn_flash_workspace_bytes": (Z, [I, I, I, I, I, I]),
    "dvd_attn_flash_fwd": (I, [P, L, P, L, P, L, P, L, P, I, I, I, I, I, P, Z, P]),
    "dvd_attn_flash_bwd": (I, [P, L, P, L, P, L, P, L, P, L, P, P, L, P, L, P, L, I, I, I, I, I, P, Z, P]),
    "dvd_avgpool_fwd": (I, [P, L, I, I, I, I, I, I, F, I, P, P]),
    "dvd_avgpool_bwd": (I, [P, L, I, I, I, I, I, I, I, P, P]),
    "dvd_maxpool_fwd": (I, [P, L, I, I, I, I, I, I, P, P]),
    "dvd_maxpool_bwd": (I, [P, P, L, I, I, I, I, I, I, P, P]),
    "dvd_phi_fwd": (I, [P, I, I, I, I, I, P, P]),
    "dvd_phi_bwd": (I, [P, I, I, I, I, I, I, P, P]),
    "dvd_gather_frames_fwd": (I, [P, P, I, I, I, L, P, P]),
    "dvd_gather_frames_bwd": (I, [P, P, I, I, I, L, I, P, P]),
    "dvd_permute5": (I, [P, ctypes.POINTER(c_int), ctypes.POINTER(c_int), P, P]),
    "dvd_permute_bctp": (I, [P, I, I, I, L, P, P]),
    "dvd_act_fwd": (I, [P, L, I, P, P]),
    "dvd_act_bwd": (I, [P, P, L, I, P, P]),
    "dvd_scale_residual_fwd": (I, [P, P, P, L, P, P]),
    "dvd_scale_residual_bwd": (I, [P, P, P, L, P, P, P, P]),
    "dvd_channel_sum": (I, [P, I, I, L, L, I, P, P, P]),
    "dvd_axpby": (I, [P, F, F, L, P, P]),
    "dvd_gather_flat": (I, [ctypes.POINTER(c_void_p), ctypes.POINTER(c_int64), ctypes.POINTER(c_int64), I, P, P]),
    "dvd_embedding_fwd": (I, [P, P, I, I, I, P, P]),
    "dvd_embedding_bwd": (I, [P, P, I, I, I, P, P]),
    "dvd_index_errors": (I, [ctypes.POINTER(ctypes.c_uint), I, P]),
    "dvd_dhead_fwd": (I, [P, I, I, I, I, I, P, P, P, P, P, P, P, P, P]),
    "dvd_dhead_bwd": (I, [P, P, P, I, I, I, I, I, P, P, P, P, P, P, P, P, P, P]),
    "dvd_clip_transform": (I, [P, I, I, I, I, P, P, P, P, I, P, P, I, I, I, F, ctypes.POINTER(c_float),
                               ctypes.POINTER(c_float), P, P]),
    "dvd_gan_loss_fwd": (I, [P, I, F, I, I, P, P]),
    "dvd_gan_loss_bwd": (I, [P, P, I, F, I, P, P]),
    "dvd_adam_step": (I, [P, P, P, P, L, F, F, F, F, I, F, P]),
}
EXPORTS = sorted(list(_SIGS) + ["dvd_last_error"])


def lib():
    """Load libdvdgan_b200.so (built in-tree).  Raises loudly when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"dvdgan_b200: CUDA library not built ({LIB_PATH} missing). Run `python -m dvdgan_b200.build` "
                "or `__graft_entry__.build()`; there is no CPU / PyTorch fallback.")
        l = ctypes.CDLL(LIB_PATH)
        l.dvd_last_error.restype = ctypes.c_char_p
        l.dvd_last_error.argtypes = []
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype = res
            fn.argtypes = args
        _lib = l
        # A/B switches for experiments: DVD_OPTIONS="oneacc=1,pair=0" (the library itself never reads the environment)
        for kv in filter(None, os.environ.get("DVD_OPTIONS", "").split(",")):
            k, _, v = kv.partition("=")
            set_option(k.strip(), int(v))
    return _lib


def set_option(name, value):
    l = _lib if _lib is not None else lib()
    if l.dvd_set_option(name.encode(), int(value)) != 0:
        raise ValueError(l.dvd_last_error().decode())


def get_option(name):
    v = c_int()
    if lib().dvd_get_option(name.encode(), ctypes.byref(v)) != 0:
        raise ValueError(lib().dvd_last_error().decode())
    return v.value


def index_errors(reset=True):
    """Out-of-range class ids seen by the embedding / head kernels on the current device (synchronises)."""
    n = ctypes.c_uint()
    rc = lib().dvd_index_errors(ctypes.byref(n), int(reset), stream())
    if rc != 0:
        raise RuntimeError(lib().dvd_last_error().decode())
    return n.value


def saturation_count(reset=True):
    """8-element operand groups clamped to fp16's range by the forward planes on the current device (synchronises)."""
    n = ctypes.c_uint()
    if lib().dvd_saturation_count(ctypes.byref(n), int(reset), stream()) != 0:
        raise RuntimeError(lib().dvd_last_error().decode())
    return n.value


def scratch_bytes():
    hw, res = ctypes.c_longlong(), ctypes.c_longlong()
    if lib().dvd_scratch_bytes(ctypes.byref(hw), ctypes.byref(res)) != 0:
        raise RuntimeError(lib().dvd_last_error().decode())
    return hw.value, res.value


def ptr(t):
    if t is None:
        return None
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream


def call(name, *args):
    """Invoke a status-returning C-ABI function on the current stream; raise on error."""
    global LAUNCHES
    l = lib()
    rc = getattr(l, name)(*args, stream())
    LAUNCHES += 1
    if rc != 0:
        raise RuntimeError(f"{name}: {l.dvd_last_error().decode()}")


def _require(tensors, dtype, what):
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("dvdgan_b200 ops run on a CUDA device only (there is no CPU fallback); "
                               "got a tensor on " + str(t.device))
        if t.dtype is not dtype:
            raise TypeError(f"dvdgan_b200 kernels read {what}; got {t.dtype} (cast it first)")


def require_cuda(*tensors):
    """Data tensors handed to the C ABI are raw ``const float*`` there: refuse host tensors (no CPU fallback) and any
    dtype other than float32 instead of reinterpreting the bytes."""
    _require(tensors, torch.float32, "float32 data")


def require_index(*tensors):
    """Index tensors (class ids, frame indices) are ``const int64_t*`` in the C ABI."""
    _require(tensors, torch.int64, "int64 indices")
