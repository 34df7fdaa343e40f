"""ctypes binding of libsimulst_b200.so (the C ABI declared in include/simulst_b200.h).

There is exactly one backend: the sm_100a CUDA library built in-tree by
``make -C simulst_b200/csrc`` (or ``__graft_entry__.build()``).  If it is missing or
cannot be loaded, every operator raises -- there is NO CPU or PyTorch fallback.
"""
import ctypes
import os
from ctypes import c_float, c_int, c_longlong, c_uint, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsimulst_b200.so")

F32, BF16, F16 = 0, 1, 2
_DTYPE_ENUM = {torch.float32: F32, torch.bfloat16: BF16, torch.float16: F16}

ST_NAN, ST_RANGE, ST_NEGPROD, ST_NOT_RIGHT_PADDED = 1, 2, 4, 8

MMA_MASS_PRESERVATION = 1
MMA_SOFT = 2
MMA_ENERGY_F16_FILL = 4
MMA_LEFT_PADDING = 8
MMA_RIGHT_PADDING = 16
MMA_MAX_SRC = 16384
SSNT_MAX_SRC = 4096
SOFT_ATTENTION_MAX_SRC = 9600      # stand-alone soft attention: 6 fp32 rows + scratch in 227 KB

_lib = None
_load_error = None

# name -> (restype, argtypes); must list every symbol declared in include/simulst_b200.h
SIGNATURES = {
    "simulst_version": (c_int, []),
    "simulst_error_string": (ctypes.c_char_p, [c_int]),
    "simulst_launch_count": (c_longlong, []),
    "simulst_reset_launch_count": (None, []),
    "simulst_mma_set_config": (c_int, [c_int, c_int]),
    "simulst_mma_set_tma": (c_int, [c_int]),
    "simulst_mma_set_pipeline": (c_int, [c_int]),
    "simulst_mma_set_mask_split": (c_int, [c_int]),
    "simulst_mma_set_cluster": (c_int, [c_int]),
    "simulst_mma_set_cluster_shape": (c_int, [c_int, c_int]),
    "simulst_cif_set_tile_rows": (c_int, [c_int, c_int]),
    "simulst_cif_set_tile": (c_int, [c_int]),
    "simulst_mma_train_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p,
                                      c_int, c_int, c_int, c_float, c_int, c_uint,
                                      c_void_p, c_void_p]),
    "simulst_mma_train_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_void_p, c_int, c_void_p, c_int,
                                      c_int, c_int, c_int, c_float, c_int, c_uint, c_void_p]),
    "simulst_mma_train_fwd_delays": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_int, c_int, c_int, c_float, c_int, c_uint,
                                             c_void_p, c_void_p]),
    "simulst_mma_train_bwd_delays": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_int, c_void_p, c_int,
                                             c_int, c_int, c_int, c_float, c_int, c_uint, c_void_p]),
    "simulst_mma_out_pitch": (c_int, [c_int]),
    "simulst_mma_train_fwd_pitched": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_int, c_longlong, c_void_p,
                                              c_void_p, c_longlong, c_void_p, c_longlong, c_void_p, c_void_p,
                                              c_int, c_int, c_int, c_float, c_int, c_uint,
                                              c_void_p, c_void_p]),
    "simulst_mma_train_bwd_pitched": (c_int, [c_void_p, c_int, c_longlong, c_void_p, c_int, c_longlong, c_void_p,
                                              c_void_p, c_longlong, c_void_p, c_void_p, c_longlong,
                                              c_void_p, c_longlong, c_void_p,
                                              c_void_p, c_int, c_longlong, c_void_p, c_int, c_longlong,
                                              c_int, c_int, c_int, c_float, c_int, c_uint, c_void_p]),
    "simulst_mma_pooled_is_fused": (c_int, [c_int, c_int, c_int, c_int, c_uint, c_int]),
    "simulst_mma_pooled_workspace_bytes": (c_longlong, [c_int, c_int, c_int, c_int]),
    "simulst_mma_set_pooled_grid": (c_int, [c_int]),
    "simulst_mma_train_fwd_pooled": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_int, c_int, c_int, c_float, c_int, c_uint,
                                             c_void_p, c_void_p]),
    "simulst_mma_train_bwd_pooled": (c_int, [c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p,
                                             c_int, c_int, c_int, c_float, c_int, c_uint, c_void_p]),
    "simulst_multimem_allreduce_f32": (c_int, [c_void_p, c_longlong, c_int, c_int, c_int, c_void_p]),
    "simulst_dal_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "simulst_dal_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                c_void_p]),
    "simulst_ssnt_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "simulst_ssnt_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_float, c_float, c_void_p]),
    "simulst_logprob_check": (c_int, [c_void_p, c_int, c_longlong, c_float, c_void_p, c_void_p]),
    "simulst_ctc_workspace_bytes": (c_longlong, [c_int, c_int, c_int]),
    "simulst_ctc_best_alignment": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_int,
                                           c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int, c_int, c_int, c_int, c_void_p]),
    "simulst_soft_attention_fwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                           c_int, c_int, c_int, c_float, c_int, c_uint,
                                           c_void_p, c_void_p]),
    "simulst_soft_attention_bwd": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                           c_void_p, c_void_p,
                                           c_int, c_int, c_int, c_float, c_int, c_uint, c_void_p]),
    "simulst_mass_preservation_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                              c_uint, c_void_p, c_void_p]),
    "simulst_mass_preservation_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_int, c_int, c_int, c_uint, c_void_p]),
    "simulst_moving_sum": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_int, c_int,
                                   c_void_p]),
    "simulst_exclusive_cumprod": (c_int, [c_void_p, c_void_p, c_int, c_longlong, c_int, c_float,
                                          c_int, c_void_p, c_void_p]),
    "simulst_cumprod_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_int, c_float,
                                    c_int, c_void_p]),
    "simulst_p_choose": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_longlong, c_void_p]),
    "simulst_mma_step": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_int, c_int, c_uint, c_void_p]),
    "simulst_cif_workspace_bytes": (c_longlong, [c_int, c_int]),
    "simulst_cif_seg_stride": (c_int, [c_int, c_int, c_int, c_float]),
    "simulst_cif_plan": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_void_p, c_int,
                                 c_int, c_int, c_float, c_void_p, c_void_p]),
    "simulst_cif_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                c_void_p, c_int,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                c_void_p]),
    "simulst_cif_bwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int,
                                c_void_p]),
}


class BackendUnavailable(RuntimeError):
    """The sm_100a extension is missing / not loadable / there is no CUDA device."""


def load():
    """Load the shared library once; raises BackendUnavailable loudly on failure."""
    global _lib, _load_error
    if _lib is not None:
        return _lib
    if _load_error is not None:
        raise BackendUnavailable(_load_error)
    if not os.path.isfile(LIB_PATH):
        _load_error = (f"{LIB_PATH} not found: build it with `make -C simulst_b200/csrc -j8` "
                       "(or __graft_entry__.build()). simulst_b200 has no CPU fallback.")
        raise BackendUnavailable(_load_error)
    try:
        lib = ctypes.CDLL(LIB_PATH)
    except OSError as exc:  # pragma: no cover - depends on the machine
        _load_error = f"cannot load {LIB_PATH}: {exc}"
        raise BackendUnavailable(_load_error) from exc
    missing = []
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing and not os.environ.get("SIMULST_B200_DEV"):
        _load_error = f"{LIB_PATH} lacks symbols declared in include/simulst_b200.h: {missing}"
        raise BackendUnavailable(_load_error)
    _lib = lib
    return lib


def dtype_enum(t: torch.dtype) -> int:
    try:
        return _DTYPE_ENUM[t]
    except KeyError:
        raise TypeError(f"simulst_b200 supports float32/bfloat16/float16 tensors, got {t}")


def ptr(t):
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(*tensors):
    """Every operator needs CUDA tensors on one device; anything else is an error (no fallback)."""
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise BackendUnavailable(
                "simulst_b200 operators run only on CUDA (sm_100a) tensors; got a "
                f"{t.device} tensor. There is no CPU fallback.")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise ValueError("all tensors must be on the same CUDA device")
    return dev


def check(rc: int, what: str):
    if rc != 0:
        msg = load().simulst_error_string(rc).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {rc})")


# ----------------------------------------------------------------------------- status word
_status_words = {}
_strict = os.environ.get("SIMULST_B200_STRICT", "0") not in ("0", "", "false", "False")


def set_strict(flag: bool):
    """strict=True: inspect the device status word after every checked call (one host sync
    per call, like the reference's prob_check).  Default: lazy, see check_status()."""
    global _strict
    _strict = bool(flag)


def is_strict() -> bool:
    return _strict


_right_padding = os.environ.get("SIMULST_B200_RIGHT_PADDING", "0") not in ("0", "", "false", "False")


def assume_right_padding(flag: bool):
    """Promise that every padding mask handed to the MMA training op is a RIGHT-padding mask
    (mask[n, j] == (j >= len_n)) -- what the reference's mass_preservation(left_padding=False) and
    MMACriterion ("Only right padding is supported", mma_criterion.py:166) assume anyway.  Masked
    rows then run the dense backward kernel (about 2x faster than the arbitrary-mask kernel at the
    training shape).  Default off: without the promise arbitrary masks are honoured element by
    element.  Env: SIMULST_B200_RIGHT_PADDING=1."""
    global _right_padding
    _right_padding = bool(flag)


def right_padding_assumed() -> bool:
    return _right_padding


_pitched_outputs = os.environ.get("SIMULST_B200_PITCHED_OUTPUTS", "1") not in ("0", "", "false", "False")


def set_pitched_outputs(flag: bool):
    """MMA training op, source lengths whose rows are not 16-byte multiples (src_len % 8 != 0): allocate
    alpha / beta / the gradients with a 16-byte row pitch and return them as [..., :src_len] views of the
    pitched buffers (default on).  The dense kernels then take such rows with shifted staging; with
    contiguous outputs they fall back to the generic kernels (about 2.5x slower).  A consumer that needs
    contiguous memory calls .contiguous() (one extra pass) or switches this off.
    Env: SIMULST_B200_PITCHED_OUTPUTS=0."""
    global _pitched_outputs
    _pitched_outputs = bool(flag)


def pitched_outputs() -> bool:
    return _pitched_outputs


def status_word(device):
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    w = _status_words.get(key)
    if w is None:
        w = torch.zeros(1, dtype=torch.int32, device=device)
        _status_words[key] = w
    return w


def raise_for_status(bits: int):
    if bits & ST_NAN:
        raise AssertionError("Nan in a probability tensor.")
    if bits & ST_RANGE:
        raise AssertionError("Incorrect values in a probability tensor, 0.0 <= tensor <= 1.0")
    if bits & ST_NEGPROD:
        raise RuntimeError("Safe cumprod can only take non-negative tensors as input."
                           "Consider use torch.cumprod if you want to calculate negative values.")
    if bits & ST_NOT_RIGHT_PADDED:
        raise RuntimeError("simulst_b200.assume_right_padding(True) is set but a padding mask was not a "
                           "right-padding mask; the results of that call are invalid")


def check_status(device=None):
    """Read (one sync) and clear the data-error word of `device`; raises like the reference's
    prob_check / safe_cumprod if any kernel since the last check saw NaN / out-of-range input."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    w = status_word(device)
    bits = int(w.item())
    if bits:
        w.zero_()
        raise_for_status(bits)


def maybe_check(device):
    if _strict:
        check_status(device)
