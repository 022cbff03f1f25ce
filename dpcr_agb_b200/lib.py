"""ctypes binding of ``libb200sparse.so`` (the C ABI declared in ``include/b200sparse.h``).

There is NO fallback: if the library is missing or a call fails, this raises.  PyTorch is used only
for device memory and streams -- every tensor handed to :func:`call` is passed as a raw device
pointer together with ``torch.cuda.current_stream().cuda_stream``.
"""
from __future__ import annotations

import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libb200sparse.so")

_i32, _i64, _f32, _vp = ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p

# name -> (restype, argtypes); mirrors include/b200sparse.h one to one
SIGNATURES = {
    "b2s_last_error": (ctypes.c_char_p, []),
    "b2s_version": (_i32, []),
    "b2s_device_check": (_i32, []),
    "b2s_set_tuning": (_i32, [ctypes.c_char_p, _i32]),
    "b2s_quantize_points": (_i32, [_vp, _i64, _vp, _f32, _vp, _vp, _vp]),
    "b2s_quantize_workspace_bytes": (_i64, [_i64, _vp]),
    "b2s_quantize_count": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp]),
    "b2s_quantize_fill": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
    "b2s_gather_rows": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "b2s_plot_transform": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp]),
    "b2s_compact_points": (_i32, [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2s_select_by_rank": (_i32, [_vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp]),
    "b2s_point_features": (_i32, [_vp, _i64, _vp, _f32, _f32, _vp, _vp]),
    "b2s_coords_augment": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp, _vp]),
    "b2s_hash_capacity": (_i64, [_i64]),
    "b2s_scan_workspace_bytes": (_i64, [_i64]),
    "b2s_coordmap_insert": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp]),
    "b2s_coordmap_fill": (_i32, [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _i64, _vp, _vp]),
    "b2s_kernel_map": (_i32, [_vp, _i64, _vp, _vp, _i64, _vp, _vp, _i32, _vp, _vp]),
    "b2s_kernel_map_dense": (_i32, [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp]),
    "b2s_kernel_map_lines": (_i32, [_vp, _i64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2s_kernel_map_pair_counts": (_i32, [_vp, _i32, _i64, _vp, _vp]),
    "b2s_kernel_map_pairs_fill": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp, _vp]),
    "b2s_conv_workspace_bytes": (_i64, [_i64, _i64, _i32, _i32, _i32, _i32]),
    "b2s_conv_col_stats_elems": (_i64, [_i64, _i32]),
    "b2s_conv_weight_image_bytes": (_i64, [_i32, _i32, _i32]),
    "b2s_conv_weight_image": (_i32, [_vp, _i32, _i32, _i32, _i32, _vp, _i64, _vp]),
    "b2s_bn_bwd_colsum_rows": (_i64, [_i64, _i32]),
    "b2s_sum_rows": (_i32, [_vp, _i64, _i32, _vp, _vp]),
    "b2s_round_tf32": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "b2s_conv_gather_gemm": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i32, _i32, _i32, _i32, _vp, _vp, _i64, _i32,
                                    _vp, _vp]),
    "b2s_conv_wgrad": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _i64, _i32, _i32, _vp]),
    "b2s_conv_lines_supported": (_i32, [_i32, _i32, _vp]),
    "b2s_conv_lines_workspace_bytes": (_i64, [_i64, _i32, _i32, _vp]),
    "b2s_conv_lines_fwd": (_i32, [_vp, _vp, _vp, _vp, _i64, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp, _vp]),
    "b2s_conv_lines_wgrad": (_i32, [_vp, _vp, _vp, _i64, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _vp]),
    "b2s_parity_plan_rows": (_i64, [_i64]),
    "b2s_parity_plan": (_i32, [_vp, _i64, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2s_conv_dgrad_strided_workspace_bytes": (_i64, [_i32, _i32, _i32]),
    "b2s_conv_dgrad_strided": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _i64, _i32, _vp]),
    "b2s_colsum": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "b2s_maxpool_fwd": (_i32, [_vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b2s_maxpool_bwd": (_i32, [_vp, _vp, _i64, _i64, _vp, _i32, _vp, _vp]),
    "b2s_sumpool": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp]),
    "b2s_nbr_inv_counts": (_i32, [_vp, _i32, _i64, _vp, _vp, _vp]),
    "b2s_batch_counts": (_i32, [_vp, _i32, _i64, _vp, _i32, _vp, _vp]),
    "b2s_segment_sum": (_i32, [_vp, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "b2s_segment_max": (_i32, [_vp, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "b2s_segment_max_bwd": (_i32, [_vp, _vp, _i64, _i32, _i32, _vp, _vp]),
    "b2s_segment_bcast": (_i32, [_vp, _vp, _i32, _i64, _vp, _i32, _vp, _vp, _vp]),
    "b2s_bcast_mul_fwd": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _vp]),
    "b2s_bcast_mul_bwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "b2s_bn_stats": (_i32, [_vp, _i64, _vp, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _vp, _vp]),
    "b2s_bn_finalize": (_i32, [_vp, _i64, _vp, _i32, _f32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "b2s_bn_apply": (_i32, [_vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "b2s_bn_bwd_reduce": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _vp, _vp, _vp]),
    "b2s_bn_bwd_apply": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b2s_gelu_fwd": (_i32, [_vp, _i64, _vp, _i32, _vp, _vp]),
    "b2s_gelu_bwd": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _vp]),
    "b2s_se_gate_fwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b2s_se_gate_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                               _vp, _vp]),
    "b2s_gated_add_gelu_fwd": (_i32, [_vp, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "b2s_gated_add_gelu_bwd": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "b2s_bcast_add_": (_i32, [_vp, _vp, _vp, _i32, _i64, _vp, _i32, _vp]),
    "b2s_add_gelu_fwd": (_i32, [_vp, _vp, _i64, _vp, _i32, _vp, _vp, _vp, _vp]),
    "b2s_grad_check": (_i32, [_vp, _i64, _f32, _vp, _vp]),
    "b2s_adabelief_step": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp]),
}

_lib = None
_checked_devices = set()
launch_count = 0  # number of C-ABI compute calls issued (each launches >= 1 kernel of libb200sparse)

# Optional per-entry-point CUDA-event timing (bench.py): name -> list of (start, end) events recorded on the
# stream the kernels are launched on.  ``_profile_names`` None = every entry point.
_profile = None
_profile_names = None


def profile_start(names=None):
    global _profile, _profile_names
    _profile, _profile_names = {}, (set(names) if names is not None else None)


def profile_stop():
    """Returns {key: (num_calls, total_ms)}; synchronises the device."""
    global _profile
    prof, _profile = _profile, None
    if not prof:
        return {}
    torch.cuda.synchronize()
    return {k: (len(v), sum(s.elapsed_time(e) for s, e in v)) for k, v in prof.items()}


class B2SError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (once) and bind every symbol of the header."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise B2SError(
                f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import "
                f"__graft_entry__ as g; g.build()'` (or `make -C dpcr_agb_b200/csrc`). There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here == header and library out of sync
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def set_tuning(key: str, value: int) -> None:
    """Diagnostic knob override (``b2s_set_tuning``); kernel results do not depend on it."""
    lib = load()
    if lib.b2s_set_tuning(key.encode(), int(value)) != 0:
        raise B2SError(lib.b2s_last_error().decode())


def ptr(t):
    """Device (or host) pointer of a tensor / ctypes array / None."""
    if t is None:
        return None
    if isinstance(t, torch.Tensor):
        return t.data_ptr()
    return ctypes.cast(t, ctypes.c_void_p).value


def host_i32(*vals):
    return (ctypes.c_int32 * len(vals))(*[int(v) for v in vals])


def host_f32(*vals):
    return (ctypes.c_float * len(vals))(*[float(v) for v in vals])


def host_f64(*vals):
    return (ctypes.c_double * len(vals))(*[float(v) for v in vals])


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)
_cuda_ok = None
_fn_cache = {}


def stream() -> int:
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ensure_device():
    global _cuda_ok
    if _cuda_ok is None:
        _cuda_ok = torch.cuda.is_available()
    if not _cuda_ok:
        raise B2SError("libb200sparse needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    dev = torch.cuda.current_device()
    if dev not in _checked_devices:
        lib = load()
        if lib.b2s_device_check() != 0:
            raise B2SError(lib.b2s_last_error().decode())
        _checked_devices.add(dev)
    return dev


def call(name: str, *args):
    """Invoke a status-returning entry point on the current stream; raises B2SError on failure.

    The trailing ``stream`` argument of the C function is appended automatically."""
    global launch_count
    dev = _ensure_device()
    fn = _fn_cache.get(name)
    if fn is None:
        fn = _fn_cache[name] = getattr(load(), name)
    lib = _lib
    conv = [a.data_ptr() if isinstance(a, torch.Tensor) else (a if (a is None or not isinstance(a, ctypes.Array))
                                                              else ctypes.cast(a, ctypes.c_void_p).value)
            for a in args]
    stream = (lambda: _raw_stream(dev)) if _raw_stream is not None else globals()["stream"]
    if _profile is not None and (_profile_names is None or name in _profile_names):
        key = name
        if name == "b2s_conv_gather_gemm":            # w_layout bit 0 set == dgrad
            key = name + (":dgrad" if (args[10] & 1) else ":fwd")
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        rc = fn(*conv, stream())
        ev1.record()
        _profile.setdefault(key, []).append((ev0, ev1))
    else:
        rc = fn(*conv, stream())
    launch_count += 1
    if rc != 0:
        raise B2SError(f"{name} failed ({rc}): {lib.b2s_last_error().decode()}")


def query(name: str, *args) -> int:
    """Invoke a size-query entry point (no stream, returns int64)."""
    lib = load()
    conv = [ptr(a) if isinstance(a, (torch.Tensor, ctypes.Array)) else a for a in args]
    return int(getattr(lib, name)(*conv))
