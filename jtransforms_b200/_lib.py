"""ctypes binding of libjtb200.so (the C ABI declared in include/jtb200.h).

There is no CPU fallback: if the CUDA library is missing or no device is present, every transform
raises.  ``use(path)`` exists so the CPU-only test-suite can point the host layer at the
g++-emulated build of the very same sources (tests/emu); the product never does that.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(HERE, "libjtb200.so")

OK, ERR_ARG, ERR_UNSUPPORTED, ERR_CUDA, ERR_OOM, ERR_NCCL = range(6)
FFT, DCT, DST, DHT = range(4)
F64, F32 = 0, 1
(C2C_FORWARD, C2C_INVERSE, R2C_PACKED, R2C_FULL, C2R_PACKED, C2R_FULL, R2R_FORWARD, R2R_INVERSE) = range(8)

SYMBOLS = [
    "jtb_plan_create", "jtb_plan_destroy", "jtb_plan_elements", "jtb_exec", "jtb_exec_batch", "jtb_exec_device",
    "jtb_lines_c2c_device", "jtb_host_alloc", "jtb_host_free", "jtb_host_register", "jtb_host_unregister", "jtb_fill_uniform_device", "jtb_device_count",
    "jtb_launch_count", "jtb_last_error", "jtb_version", "jtb_debug_set_limits",
    "jtb_debug_table_bytes", "jtb_plan_set_devices", "jtb_plan_device_count", "jtb_exec_n",
    "jtb_slab_create", "jtb_slab_destroy", "jtb_slab_export", "jtb_slab_connect_ipc", "jtb_slab_connect_local",
    "jtb_nccl_unique_id", "jtb_slab_nccl_init", "jtb_slab_nccl_init_local", "jtb_slab_set_exchange",
    "jtb_slab_block_elements", "jtb_slab_forward", "jtb_slab_back", "jtb_slab_group_forward", "jtb_slab_group_back",
    "jtb_slab_status", "jtb_slab_profile", "jtb_slab_last_times", "jtb_slab_chunk_times", "jtb_lines_c2c_out_device",
    "jtb_fft3d_k2_scatter", "jtb_fft3d_k2_scatter_chunk", "jtb_fft3d_k1_scatter", "jtb_fft2d_slices_device", "jtb_peer_barrier", "jtb_peer_alloc", "jtb_peer_open", "jtb_peer_close", "jtb_peer_free",
]

_lib = None


class JtbError(RuntimeError):
    """CUDA / resource failure inside libjtb200 (Java shim: IllegalStateException)."""


def _bind(lib):
    i64, vp, ci = C.c_int64, C.c_void_p, C.c_int
    lib.jtb_plan_create.argtypes = [C.POINTER(vp), ci, ci, ci, C.POINTER(i64), ci]
    lib.jtb_plan_destroy.argtypes = [vp]
    lib.jtb_plan_elements.argtypes = [vp, ci]
    lib.jtb_plan_elements.restype = i64
    lib.jtb_exec.argtypes = [vp, ci, vp, i64, ci]
    lib.jtb_exec_batch.argtypes = [vp, ci, vp, i64, i64, i64, ci]
    lib.jtb_exec_device.argtypes = [vp, ci, vp, i64, i64, ci, vp]
    lib.jtb_lines_c2c_device.argtypes = [ci, ci, vp, i64, i64, i64, i64, i64, i64, ci, C.c_double, vp]
    lib.jtb_host_alloc.argtypes = [C.POINTER(vp), i64]
    lib.jtb_host_free.argtypes = [vp]
    lib.jtb_host_register.argtypes = [vp, i64]
    lib.jtb_host_unregister.argtypes = [vp]
    lib.jtb_fill_uniform_device.argtypes = [ci, ci, vp, i64, C.c_uint64, C.c_double, C.c_double, vp]
    lib.jtb_device_count.restype = ci
    lib.jtb_fft3d_k2_scatter.argtypes = [ci, ci, vp, i64, i64, i64, ci, ci, C.POINTER(vp), ci, vp]
    lib.jtb_fft2d_slices_device.argtypes = [ci, ci, vp, i64, i64, i64, ci, ci, C.POINTER(vp), ci, vp]
    lib.jtb_fft3d_k1_scatter.argtypes = [ci, ci, vp, i64, i64, i64, ci, ci, C.POINTER(vp), ci, vp]
    lib.jtb_fft3d_k2_scatter_chunk.argtypes = [ci, ci, vp, i64, i64, i64, i64, ci, C.POINTER(vp), ci, vp]
    lib.jtb_peer_barrier.argtypes = [ci, C.POINTER(vp), ci, ci, i64, vp]
    lib.jtb_peer_alloc.argtypes = [ci, i64, C.POINTER(vp), C.c_char_p]
    lib.jtb_peer_open.argtypes = [ci, C.c_char_p, C.POINTER(vp)]
    lib.jtb_peer_close.argtypes = [ci, vp]
    lib.jtb_peer_free.argtypes = [ci, vp]
    lib.jtb_plan_set_devices.argtypes = [vp, ci, C.POINTER(ci)]
    lib.jtb_plan_device_count.argtypes = [vp]
    lib.jtb_exec_n.argtypes = [vp, ci, vp, i64, i64, ci]
    lib.jtb_slab_create.argtypes = [C.POINTER(vp), ci, i64, i64, i64, ci, ci, ci]
    lib.jtb_slab_destroy.argtypes = [vp]
    lib.jtb_slab_export.argtypes = [vp, C.c_char_p]
    lib.jtb_slab_connect_ipc.argtypes = [vp, C.c_char_p]
    lib.jtb_slab_connect_local.argtypes = [C.POINTER(vp), ci]
    lib.jtb_nccl_unique_id.argtypes = [C.c_char_p]
    lib.jtb_slab_nccl_init.argtypes = [vp, C.c_char_p]
    lib.jtb_slab_nccl_init_local.argtypes = [C.POINTER(vp), ci]
    lib.jtb_slab_set_exchange.argtypes = [vp, ci]
    lib.jtb_slab_block_elements.argtypes = [vp]
    lib.jtb_slab_block_elements.restype = i64
    lib.jtb_slab_forward.argtypes = [vp, vp, ci, ci, C.POINTER(vp), vp]
    lib.jtb_slab_back.argtypes = [vp, vp, ci, C.POINTER(vp), vp]
    lib.jtb_slab_group_forward.argtypes = [C.POINTER(vp), ci, C.POINTER(vp), ci, ci, C.POINTER(vp), C.POINTER(vp)]
    lib.jtb_slab_group_back.argtypes = [C.POINTER(vp), ci, C.POINTER(vp), ci, C.POINTER(vp), C.POINTER(vp)]
    lib.jtb_slab_status.argtypes = [vp]
    lib.jtb_slab_profile.argtypes = [vp, ci]
    lib.jtb_slab_last_times.argtypes = [vp, C.POINTER(C.c_float)]
    lib.jtb_slab_chunk_times.argtypes = [vp, C.POINTER(ci), C.POINTER(C.c_float), C.POINTER(C.c_float)]
    lib.jtb_lines_c2c_out_device.argtypes = [ci, ci, vp, vp, i64, i64, i64, i64, i64, i64, i64, ci, C.c_double, vp]
    lib.jtb_debug_table_bytes.argtypes = [ci]
    lib.jtb_debug_table_bytes.restype = i64
    lib.jtb_debug_set_limits.argtypes = [ci, ci]
    lib.jtb_launch_count.argtypes = [ci]
    lib.jtb_launch_count.restype = i64
    lib.jtb_last_error.restype = C.c_char_p
    lib.jtb_version.restype = C.c_char_p
    return lib


def use(path: str):
    """Load the C-ABI library from an explicit path (test hook)."""
    global _lib
    _lib = _bind(C.CDLL(path))
    return _lib


def get():
    global _lib
    if _lib is None:
        if not os.path.exists(DEFAULT_LIB):
            raise JtbError(
                "libjtb200.so is not built (%s). Run `python -m jtransforms_b200.build`; "
                "jtransforms_b200 has no CPU fallback." % DEFAULT_LIB)
        use(DEFAULT_LIB)
    return _lib


def check(status: int):
    if status == OK:
        return
    msg = get().jtb_last_error().decode("utf-8", "replace")
    if status == ERR_ARG:
        raise ValueError(msg)            # Java shim: IllegalArgumentException(msg)
    if status == ERR_OOM:
        raise MemoryError(msg)
    if status == ERR_NCCL:
        raise JtbError("libjtb200 NCCL error: %s" % msg)
    raise JtbError("libjtb200 error %d: %s" % (status, msg))
