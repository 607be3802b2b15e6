"""Slab-decomposed DoubleFFT_3D / FloatFFT_3D over one process per GPU.

Replaces the shared-memory slice-axis gather of the reference (cdft3db_subth,
fft/DoubleFFT_3D.java:6318-6520) when one transform is spread over P GPUs:

  rank g owns slices [g*S/P, (g+1)*S/P)  ->  local k3 and k2 passes (fft/DoubleFFT_3D.java:5505-5713)
  -> all-to-all re-slabbing over k2  ->  k1 pass on [S][R/P][C]

On GPUs the whole step -- passes, exchange, synchronisation -- runs inside libjtb200 (jtb_slab_*, csrc/jtb_slab.cu):
this module only creates the member, trades the 192 bytes of CUDA IPC handles (and, for ``exchange="nccl"``, the NCCL
unique id) through torch.distributed, and wraps the result pointer as a tensor.  Exchange variants of the library:

  * "p2p"  (default): the k2 kernel itself stores every output row into the receive buffer of the GPU that owns it
    (NVLink peer stores; fused kernels for the power-of-two shapes, a peer-store row copy for every other shape),
    followed by a device-side flag barrier -- no pack/unpack pass.
  * "nccl": k2 in place, then ncclSend/ncclRecv of the strided sub-blocks issued by the library (no pack pass).

torch.distributed's own all_to_all is only used on CPU tensors (the gloo test path of the host logic).

The result is left k2-slabbed: rank h holds out[k1][h*R/P:(h+1)*R/P][k3] as a contiguous [S][R/P][C] block,
which ``scatter_to_host`` places into the caller's natural-order host array with strided copies.  (A single process
that owns all GPUs does not need this class: ``DoubleFFT_3D(..., devices=[...])`` takes the host array directly.)
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.distributed as dist

from . import _lib


class SlabFFT3D:
    def __init__(self, slices: int, rows: int, columns: int, prec: int = _lib.F64, group=None, device_index=None,
                 exchange: str = "auto"):
        self.S, self.R, self.Cn, self.prec, self.group = int(slices), int(rows), int(columns), prec, group
        self.P = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        if self.S % self.P or self.R % self.P:
            raise ValueError("slices and rows must be divisible by the number of ranks")
        self.Ls, self.Rh = self.S // self.P, self.R // self.P
        self.dtype = torch.float64 if prec == _lib.F64 else torch.float32
        self.dev = 0 if device_index is None else int(device_index)
        self.lib = _lib.get()
        self.esize = 8 if prec == _lib.F64 else 4
        on_gpu = self.P > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl"
        if exchange == "auto":
            exchange = "p2p" if on_gpu else "host"
        self.exchange = exchange
        self._m = None                 # jtb_slab* (GPU paths and P == 1)
        if self.P == 1 or on_gpu:
            self._create_member(on_gpu)

    # ---- library member + peer connection
    def _create_member(self, connect: bool):
        lib, P = self.lib, self.P
        m = C.c_void_p()
        _lib.check(lib.jtb_slab_create(C.byref(m), self.prec, self.S, self.R, self.Cn, P, self.rank, self.dev))
        self._m = m
        if not connect:
            return
        # every rank must take the same path: agree on success before committing to peer stores
        ok, err = 1, ""
        try:
            h = C.create_string_buffer(192)
            _lib.check(lib.jtb_slab_export(m, h))
            allh = [None] * P
            dist.all_gather_object(allh, bytes(h.raw), group=self.group)
            _lib.check(lib.jtb_slab_connect_ipc(m, b"".join(allh)))
        except Exception as e:       # no peer access / IPC on this box
            ok, err = 0, repr(e)
        dev = torch.device("cuda", self.dev)
        flag = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        ok = int(flag.item())
        if self.exchange == "p2p" and not ok:
            import warnings
            warnings.warn("peer-mapped exchange unavailable (%s); using the library's NCCL exchange" % (err or "a peer failed"))
            self.exchange = "nccl"
        if self.exchange == "nccl":
            # the library's own communicator (ncclCommInitRank inside libjtb200); the id travels through torch.distributed
            ids = [None]
            if self.rank == 0:
                idb = C.create_string_buffer(128)
                _lib.check(lib.jtb_nccl_unique_id(idb))
                ids[0] = bytes(idb.raw)
            dist.broadcast_object_list(ids, src=0, group=self.group)
            _lib.check(lib.jtb_slab_nccl_init(m, ids[0]))
        _lib.check(lib.jtb_slab_set_exchange(m, 1 if self.exchange == "nccl" else 0))

    def _wrap(self, ptr: int) -> torch.Tensor:
        class _Wrap:
            pass
        wobj = _Wrap()
        wobj.__cuda_array_interface__ = {
            "shape": (2 * self.S * self.Rh * self.Cn,), "typestr": "<f8" if self.esize == 8 else "<f4",
            "data": (ptr, False), "version": 3}
        return torch.as_tensor(wobj, device=torch.device("cuda", self.dev))

    def close(self):
        if self._m:
            if torch.cuda.is_available():
                torch.cuda.synchronize()
            if dist.is_initialized() and self.P > 1:
                dist.barrier(group=self.group)      # nobody unmaps a buffer a peer may still be writing
            self.lib.jtb_slab_destroy(self._m)
            self._m = None
            if dist.is_initialized() and self.P > 1:
                dist.barrier(group=self.group)

    def __del__(self):
        try:
            if self._m and self.P == 1:
                self.lib.jtb_slab_destroy(self._m)
                self._m = None
        except Exception:
            pass

    # number of real elements (doubles/floats) of the local slab, before and after
    def local_elements(self) -> int:
        return 2 * self.Ls * self.R * self.Cn

    def status(self):
        """Raises if a device-side wait of the exchange timed out (a peer never arrived): call after a synchronize."""
        if self._m:
            _lib.check(self.lib.jtb_slab_status(self._m))

    def _lines(self, t, n, nlines, c0, d0, d3, stride, inverse=False, scale=1.0):
        stream = torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else 0
        _lib.check(self.lib.jtb_lines_c2c_device(self.prec, self.dev, C.c_void_p(t.data_ptr()), n, nlines, c0, d0, d3,
                                                 stride, int(inverse), float(scale), C.c_void_p(stream)))

    def _member_call(self, fn, t, *args):
        stream = C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else 0)
        res = C.c_void_p()
        _lib.check(fn(self._m, C.c_void_p(t.data_ptr()), *args, C.byref(res), stream))
        if res.value == t.data_ptr():
            return t
        return self._wrap(res.value)

    def forward(self, a: torch.Tensor, work: torch.Tensor | None = None) -> torch.Tensor:
        """a: local slab [Ls][R][C] interleaved complex (2*Ls*R*C reals), transformed in place for P == 1.
        Returns the tensor holding the k2-slabbed result [S][Rh][C] (``a`` itself when P == 1; otherwise a
        library-owned receive buffer that stays valid until the step after next)."""
        S, R, Cn, P, Ls, Rh = self.S, self.R, self.Cn, self.P, self.Ls, self.Rh
        if self._m is not None and (P == 1 or a.is_cuda):
            return self._member_call(self.lib.jtb_slab_forward, a, 0, 0)
        # host tensors (gloo): passes through the library, exchange through torch.distributed
        self._lines(a, Cn, Ls * R, 1, 0, Cn, 1)                       # k3: contiguous rows
        self._lines(a, R, Cn * Ls, Cn, 1, R * Cn, Cn)                 # k2: columns inside each slice
        send = a.view(Ls, P, Rh, 2 * Cn).permute(1, 0, 2, 3).contiguous()
        recv = work if work is not None else torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        # recv is [g][ls][r][c] = [S][Rh][C]
        self._lines(recv, S, Rh * Cn, Rh * Cn, 1, S * Rh * Cn, Rh * Cn)
        return recv.view(-1)

    def inverse(self, b: torch.Tensor, scale: bool = True, work: torch.Tensor | None = None) -> torch.Tensor:
        """complexInverse of a k2-slabbed spectrum: ``b`` is this rank's [S][Rh][C] block as ``forward`` returns it
        (transformed in place for the k1 pass); returns the k1-slabbed local slab [Ls][R][C] of the inverse
        transform (``b`` itself when P == 1).  ``scale`` divides by S*R*C like the reference
        (fft/DoubleFFT_3D.java:744-760)."""
        S, R, Cn, P, Ls, Rh = self.S, self.R, self.Cn, self.P, self.Ls, self.Rh
        if self._m is not None and (P == 1 or b.is_cuda):
            return self._member_call(self.lib.jtb_slab_back, b, int(bool(scale)))
        sc = 1.0 / (float(S) * float(R) * float(Cn)) if scale else 1.0
        self._lines(b, S, Rh * Cn, Rh * Cn, 1, S * Rh * Cn, Rh * Cn, inverse=True)      # k1: across slices
        recv = work if work is not None else torch.empty_like(b)
        dist.all_to_all_single(recv.view(-1), b.view(-1), group=self.group)              # chunk g = slices of rank g
        # recv is [h][ls][r][c] (source rank h holds rows h*Rh..): reorder to the local slab [ls][h*Rh + r][c]
        local = recv.view(P, Ls, Rh, 2 * Cn).permute(1, 0, 2, 3).contiguous().view(-1)
        self._lines(local, R, Cn * Ls, Cn, 1, R * Cn, Cn, inverse=True)                  # k2: columns inside each slice
        self._lines(local, Cn, Ls * R, 1, 0, Cn, 1, inverse=True, scale=sc)              # k3: contiguous rows
        return local

    def scatter_to_host(self, result: torch.Tensor, host: torch.Tensor):
        """Place this rank's [S][Rh][C] block into the natural-order host array [S][R][C] (strided D2H)."""
        if self.P == 1:
            host.view(-1).copy_(result.view(-1), non_blocking=True)
            return
        hv = host.view(self.S, self.R, 2 * self.Cn)[:, self.rank * self.Rh:(self.rank + 1) * self.Rh, :]
        hv.copy_(result.view(self.S, self.Rh, 2 * self.Cn), non_blocking=True)
