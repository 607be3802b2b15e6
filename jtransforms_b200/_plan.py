"""Plan handle + array marshalling shared by the org.jtransforms mirror classes."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib


def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


class Plan:
    """Owns one jtb_plan*.  Immutable after construction and safe to share between threads (unlike the
    reference's 2-D/3-D objects, fft/DoubleFFT_2D.java:119, fft/DoubleFFT_3D.java:149-162)."""

    def __init__(self, kind: int, prec: int, dims, device: int = 0, devices=None):
        """``devices``: list of CUDA ordinals for a multi-GPU plan (jtb_plan_set_devices): host-array calls of a
        3-D FFT plan are slab-decomposed over them, batches are split between them."""
        if devices is not None and len(devices) > 0:
            device = int(devices[0])
        self.kind, self.prec, self.dims, self.device = kind, prec, tuple(int(d) for d in dims), device
        self.np_dtype = np.float64 if prec == _lib.F64 else np.float32
        self.total = int(np.prod(self.dims))
        self._h = C.c_void_p()
        arr = (C.c_int64 * len(self.dims))(*self.dims)
        _lib.check(_lib.get().jtb_plan_create(C.byref(self._h), kind, prec, len(self.dims), arr, device))
        self.devices = [device]
        if devices is not None and len(devices) > 1:
            self.set_devices(devices)

    def set_devices(self, devices):
        devs = [int(d) for d in devices]
        arr = (C.c_int * len(devs))(*devs)
        _lib.check(_lib.get().jtb_plan_set_devices(self._h, len(devs), arr))
        self.devices, self.device = devs, devs[0]

    def __del__(self):
        try:
            if self._h:
                _lib.get().jtb_plan_destroy(self._h)
                self._h = C.c_void_p()
        except Exception:
            pass

    def elements(self, op: int) -> int:
        return int(_lib.get().jtb_plan_elements(self._h, op))

    def run(self, op: int, a, offa: int = 0, scale: bool = False, howmany: int = 1, dist: int = 0):
        """In-place transform of `a` (numpy array on the host or torch tensor on the plan's GPU)."""
        need = offa + (howmany - 1) * dist + self.elements(op)
        if _is_torch(a):
            import torch
            want = torch.float64 if self.prec == _lib.F64 else torch.float32
            if a.dtype != want or not a.is_contiguous() or not a.is_cuda:
                raise ValueError("expected a contiguous CUDA tensor of dtype %s" % want)
            if a.numel() < need:
                raise IndexError("array of %d elements is too short (need %d)" % (a.numel(), need))
            if a.device.index != self.device:
                raise ValueError("tensor lives on cuda:%s but the plan was made for cuda:%d" % (a.device.index, self.device))
            ptr = a.data_ptr() + offa * a.element_size()
            stream = torch.cuda.current_stream(a.device).cuda_stream
            _lib.check(_lib.get().jtb_exec_device(self._h, op, C.c_void_p(ptr), howmany, dist, int(bool(scale)),
                                                  C.c_void_p(stream)))
            return a
        if not isinstance(a, np.ndarray) or a.dtype != self.np_dtype or not a.flags["C_CONTIGUOUS"] or not a.flags["WRITEABLE"]:
            raise ValueError("expected a writable C-contiguous numpy array of dtype %s" % np.dtype(self.np_dtype).name)
        if a.size < need:
            # the reference surfaces this as ArrayIndexOutOfBoundsException
            raise IndexError("array of %d elements is too short (need %d)" % (a.size, need))
        if howmany == 1:
            # the length travels with the call: the library refuses to touch elements beyond it (jtb_exec_n)
            _lib.check(_lib.get().jtb_exec_n(self._h, op, C.c_void_p(a.ctypes.data), a.size, offa, int(bool(scale))))
            return a
        _lib.check(_lib.get().jtb_exec_batch(self._h, op, C.c_void_p(a.ctypes.data), offa, howmany, dist,
                                             int(bool(scale))))
        return a
