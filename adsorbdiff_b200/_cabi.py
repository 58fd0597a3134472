"""ctypes binding of the C ABI in include/adsorbdiff_b200.h.

PyTorch supplies device memory and streams; every kernel is reached through the plain-C
entry points of `lib/libadsorbdiff_b200.so` (raw device pointers + a cudaStream_t).  There
is deliberately no fallback: if the library is missing or an entry point fails, this raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int32, c_int64, c_void_p

import torch

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libadsorbdiff_b200.so")

ABI_VERSION = 3
SCHED_COLS = 6
STATUS_EMPTY_SYSTEM = 1
STATUS_ROW_OVERFLOW = 2
STATUS_F16_OVERFLOW = 4
STATUS_BAD_ELEMENT = 8
MAX_IMAGES = 2048
MAX_ATOMS_PER_SYSTEM = 1024
ACT_NONE, ACT_SSILU = 0, 1

# name -> (restype, argtypes); mirrors include/adsorbdiff_b200.h one to one
_P = c_void_p
SIGNATURES = {
    "adk_abi_version": (c_int, []),
    "adk_init": (c_int, []),
    "adk_neighbors_smem_bytes": (c_int64, [c_int, c_int, c_int]),
    "adk_message_mma_smem_bytes": (c_int64, [c_int, c_int]),
    "adk_set_tc_pair": (c_int, [c_int]),
    "adk_gather_rows": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "adk_scatter_rows": (c_int, [_P, _P, c_int, c_int, _P, _P]),
    "adk_mark_sources": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P]),
    "adk_neighbors": (c_int, [_P, _P, _P, c_int, c_int, ctypes.POINTER(c_int32), c_float, c_int,
                              _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "adk_export_edges": (c_int, [_P, _P, _P, c_int, ctypes.POINTER(c_int32), c_int, _P, _P, _P, _P,
                                 _P, c_int64, _P, _P, _P, _P, _P]),
    "adk_embed": (c_int, [_P, _P, c_int, c_int, c_int, _P, _P, _P, _P]),
    "adk_layernorm": (c_int, [_P, _P, _P, c_int, c_int, c_float, _P, _P, c_int64, c_float, _P, _P]),
    "adk_linear": (c_int, [_P, c_int64, _P, _P, c_int, c_int, c_int, c_int, _P, c_int64, _P]),
    "adk_split_f16": (c_int, [_P, c_int64, c_int, c_int, c_float, _P, c_int64, _P, _P]),
    "adk_split_f16_multi": (c_int, [_P, c_int, c_float, _P, _P]),
    "adk_linear_tc": (c_int, [_P, c_int64, c_int, _P, c_int, c_int, _P, c_float, c_int, _P, c_int64,
                              _P, c_int64, c_float, _P, _P]),
    "adk_amax_scale": (c_int, [_P, c_int64, c_float, _P, _P, _P]),
    "adk_split_f16_dev": (c_int, [_P, c_int64, c_int, c_int, _P, _P, c_int64, _P, _P]),
    "adk_split_f16_t_dev": (c_int, [_P, c_int64, c_int, c_int, _P, _P, c_int64, c_int64, _P, _P]),
    "adk_linear_tc_dev": (c_int, [_P, c_int64, c_int, _P, c_int, c_int, _P, _P, _P, _P, c_int64, _P, _P]),
    "adk_linear_train_ws_bytes": (c_int64, [c_int, c_int, c_int]),
    "adk_linear_train_fwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P, _P]),
    "adk_linear_train_bwd": (c_int, [_P, _P, _P, c_int, c_int, c_int, c_float, _P, _P, _P, _P, _P, _P, _P, _P]),
    "adk_update_prep_bwd": (c_int, [_P, _P, _P, c_int, c_int, _P, _P, _P]),
    "adk_update_gate_bwd": (c_int, [_P, _P, _P, _P, _P, _P, c_int, c_int, _P, _P, _P, _P, _P]),
    "adk_message": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_int,
                            _P, _P, _P]),
    "adk_split_f16_transpose": (c_int, [_P, c_int, c_int, c_float, _P, _P, _P]),
    "adk_message_mma": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_float, _P, _P, c_int, c_int,
                                c_float, c_int, c_float, _P, _P, _P, c_int64, c_float, _P, _P]),
    "adk_message_t5_smem_bytes": (c_int64, [c_int, c_int]),
    "adk_message_t5": (c_int, [_P, c_int, c_int, _P, _P, _P, _P, _P, _P, _P, _P, c_float, _P, _P, c_int, c_int,
                               c_float, c_int, c_float, _P, _P, _P, c_int64, c_float, _P, _P]),
    "adk_message_bwd_scratch_floats": (c_int64, [c_int, c_int, c_int, _P]),
    "adk_message_bwd_plan_ints": (c_int64, [c_int, c_int64]),
    "adk_message_bwd_plan": (c_int, [_P, _P, _P, _P, c_int, c_int, c_float, c_int, _P, _P]),
    "adk_message_bwd": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, c_int, c_int, c_float, c_int,
                                _P, _P, _P, _P, _P, _P, _P, _P]),
    "adk_update_prep": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, c_int64, c_float, _P, _P]),
    "adk_update_gate": (c_int, [_P, _P, _P, _P, c_int, c_int, _P, _P, _P, c_int64, c_float, _P, _P]),
    "adk_head_prep": (c_int, [_P, _P, c_int, c_int, _P, _P, c_int64, c_float, _P, _P]),
    "adk_head_gate": (c_int, [_P, _P, c_int, c_int, _P, _P, _P, c_int64, c_float, _P, _P]),
    "adk_init_placement": (c_int, [_P, _P, _P, _P, _P, c_int, _P]),
    "adk_se3_step": (c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, c_int, _P, _P, _P, _P]),
    "adk_early_stop": (c_int, [_P, c_int, c_float, _P, _P, _P, _P, c_int64, _P]),
}


class AdkError(RuntimeError):
    pass


class AdkOverflow(AdkError):
    """STATUS_F16_OVERFLOW: an operand exceeded the fp16 range under the current prescales."""


_lib = None
_inited_devices: set[int] = set()
launch_count = 0  # kernels launched through this binding (bench.py reports it)

_LAUNCHES = {  # kernels behind one entry-point call
    "adk_neighbors": 1, "adk_export_edges": 2, "adk_embed": 1, "adk_layernorm": 1, "adk_linear": 1, "adk_split_f16": 1, "adk_split_f16_multi": 1, "adk_linear_tc": 1,
    "adk_message": 1, "adk_message_mma": 1, "adk_message_t5": 1, "adk_message_bwd": 4, "adk_amax_scale": 1, "adk_split_f16_dev": 1, "adk_split_f16_t_dev": 1, "adk_linear_tc_dev": 1, "adk_linear_train_fwd": 5, "adk_update_prep_bwd": 1, "adk_update_gate_bwd": 1, "adk_linear_train_bwd": 8, "adk_message_bwd_plan": 1, "adk_split_f16_transpose": 1, "adk_update_prep": 1, "adk_update_gate": 1, "adk_head_prep": 1, "adk_head_gate": 1, "adk_gather_rows": 1, "adk_scatter_rows": 1, "adk_mark_sources": 2,
    "adk_init_placement": 1, "adk_se3_step": 2, "adk_early_stop": 2,
}


def lib_path() -> str:
    return _LIB_PATH


def load() -> ctypes.CDLL:
    """Load the shared library and bind every declared symbol (no GPU needed)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise AdkError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(adsorbdiff_b200 has no CPU or PyTorch fallback)")
    lib = ctypes.CDLL(_LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype, fn.argtypes = res, args
    if lib.adk_abi_version() != ABI_VERSION:
        raise AdkError(f"ABI mismatch: library {lib.adk_abi_version()} != binding {ABI_VERSION}")
    _lib = lib
    return lib


def _ensure_init(device: torch.device) -> None:
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx in _inited_devices:
        return
    with torch.cuda.device(idx):
        rc = load().adk_init()
    if rc != 0:
        raise AdkError(f"adk_init failed with code {rc}")
    _inited_devices.add(idx)


def ptr(t: torch.Tensor | None):
    if t is None:
        return None
    if not t.is_cuda:
        raise AdkError("adsorbdiff_b200 kernels take CUDA tensors only (no CPU fallback)")
    if not t.is_contiguous():
        raise AdkError("non-contiguous tensor passed to a kernel")
    return t.data_ptr()


def call(name: str, device: torch.device, *args) -> None:
    """Invoke an entry point on torch's current stream of `device`; raise on a non-zero code.
    The launch happens with `device` current (kernels and streams belong to a device), whatever the caller's
    current device is."""
    global launch_count
    _ensure_init(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        with torch.cuda.device(idx):
            return call(name, torch.device("cuda", idx), *args)
    stream = torch.cuda.current_stream(device).cuda_stream
    rc = getattr(load(), name)(*args, stream)
    if rc != 0:
        raise AdkError(f"{name} failed with code {rc}")
    launch_count += _LAUNCHES.get(name, 0)


def rep_array(rep):
    return (c_int32 * 3)(*[int(r) for r in rep])


def capture_graph(fn, device: torch.device) -> "torch.cuda.CUDAGraph":
    """Capture the launches `fn()` enqueues (through `call`) into a CUDA graph.  Unlike the `torch.cuda.graph`
    context manager this does not synchronise the device, collect garbage or empty the caching allocators first --
    that costs ~0.3 s at a 1024-system plan (all cached blocks go back to the driver and must be re-allocated), and
    nothing here needs it: the step allocates nothing, every buffer it touches is owned by the launch plan."""
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream(device)
    cur = torch.cuda.current_stream(device)
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        g.capture_begin()
        try:
            fn()
        finally:
            g.capture_end()
    cur.wait_stream(side)
    return g
