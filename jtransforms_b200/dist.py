"""Slab-decomposed DoubleFFT_3D / FloatFFT_3D over one process per GPU (torch.distributed for the plumbing).

Replaces the shared-memory slice-axis gather of the reference (cdft3db_subth,
fft/DoubleFFT_3D.java:6318-6520) when one transform is spread over P GPUs:

  rank g owns slices [g*S/P, (g+1)*S/P)  ->  local k3 and k2 passes (fft/DoubleFFT_3D.java:5505-5713)
  -> all-to-all re-slabbing over k2  ->  k1 pass on [S][R/P][C]

Two exchange implementations:
  * "p2p"  (default on GPUs): the k2 kernel itself stores every output row into the receive buffer of the GPU
    that owns it (peer-mapped memory over NVLink/NVSwitch, jtb_fft3d_k2_scatter), followed by a device-side
    flag barrier -- the transpose costs no extra pass over HBM and no pack/unpack.
  * "nccl": k2 in place, pack, torch.distributed.all_to_all_single, then k1 (baseline; also the gloo CPU path).

The result is left k2-slabbed: rank h holds out[k1][h*R/P:(h+1)*R/P][k3] as a contiguous [S][R/P][C] block,
which ``scatter_to_host`` places into the caller's natural-order host array with strided copies.
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
        if exchange == "auto":
            exchange = "p2p" if (self.P > 1 and dist.is_initialized() and dist.get_backend(group) == "nccl") else "nccl"
        self.exchange = exchange
        self.step = 0
        self._peer = None
        self._side = None
        import os as _os
        self.chunks = int(_os.environ.get("JTB_SLAB_CHUNKS", "1"))   # >1: pipeline k3 under the exchange (measured: no gain)
        if self.P > 1 and exchange == "p2p":
            # every rank must take the same path: agree on success before committing to peer stores
            ok = 1
            try:
                self._setup_p2p()
            except Exception as e:       # no peer access / IPC on this box
                ok = 0
                self._p2p_error = repr(e)
            flag = torch.tensor([ok], dtype=torch.int32, device=torch.device("cuda", self.dev))
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
            if int(flag.item()) == 0:
                import warnings
                warnings.warn("peer-mapped exchange unavailable (%s); using the NCCL all-to-all"
                              % getattr(self, "_p2p_error", "a peer failed"))
                self._peer = None
                self.exchange = "nccl"

    # ---- peer-mapped receive buffers (double buffered) + barrier flags
    def _setup_p2p(self):
        lib, P = self.lib, self.P
        nbytes = 2 * self.S * self.Rh * self.Cn * self.esize
        mine, handles = [], []
        for size in (nbytes, nbytes, 4096):
            ptr, h = C.c_void_p(), C.create_string_buffer(64)
            _lib.check(lib.jtb_peer_alloc(self.dev, size, C.byref(ptr), h))
            mine.append(ptr.value)
            handles.append(bytes(h.raw))
        allh = [None] * P
        dist.all_gather_object(allh, handles, group=self.group)
        ptrs = [[0] * P for _ in range(3)]
        for r in range(P):
            for b in range(3):
                if r == self.rank:
                    ptrs[b][r] = mine[b]
                else:
                    q = C.c_void_p()
                    _lib.check(lib.jtb_peer_open(self.dev, allh[r][b], C.byref(q)))
                    ptrs[b][r] = q.value
        self._peer = {"mine": mine, "ptrs": ptrs, "nbytes": nbytes,
                      "arr": [(C.c_void_p * P)(*ptrs[b]) for b in range(3)]}
        # (the caller's all_reduce is the barrier that makes every mapping visible before first use)

    def _recv_tensor(self, b: int) -> torch.Tensor:
        class _Wrap:
            pass
        wobj = _Wrap()
        wobj.__cuda_array_interface__ = {
            "shape": (2 * self.S * self.Rh * self.Cn,), "typestr": "<f8" if self.esize == 8 else "<f4",
            "data": (self._peer["mine"][b], False), "version": 3}
        return torch.as_tensor(wobj, device=torch.device("cuda", self.dev))

    def close(self):
        if self._peer:
            torch.cuda.synchronize()
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for b in range(3):
                for r in range(self.P):
                    if r != self.rank:
                        self.lib.jtb_peer_close(self.dev, C.c_void_p(self._peer["ptrs"][b][r]))
            if dist.is_initialized():
                dist.barrier(group=self.group)
            for m in self._peer["mine"]:
                self.lib.jtb_peer_free(self.dev, C.c_void_p(m))
            self._peer = None

    # number of real elements (doubles/floats) of the local slab, before and after
    def local_elements(self) -> int:
        return 2 * self.Ls * self.R * self.Cn

    def _lines(self, t, n, nlines, c0, d0, d3, stride, inverse=False, scale=1.0):
        stream = torch.cuda.current_stream(t.device).cuda_stream if t.is_cuda else 0
        _lib.check(self.lib.jtb_lines_c2c_device(self.prec, self.dev, C.c_void_p(t.data_ptr()), n, nlines, c0, d0, d3,
                                                 stride, int(inverse), float(scale), C.c_void_p(stream)))

    def forward(self, a: torch.Tensor, work: torch.Tensor | None = None) -> torch.Tensor:
        """a: local slab [Ls][R][C] interleaved complex (2*Ls*R*C reals), transformed in place for P == 1.
        Returns the tensor holding the k2-slabbed result [S][Rh][C] (``a`` itself when P == 1)."""
        S, R, Cn, P, Ls, Rh = self.S, self.R, self.Cn, self.P, self.Ls, self.Rh
        stream = C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream if a.is_cuda else 0)
        if P == 1 or self.exchange == "p2p":
            # both in-slice passes in one call (one persistent kernel for 512^2 double slices); with P > 1 the
            # column pass stores ARE the all-to-all (NVLink peer stores)
            if P > 1:
                b = self.step & 1
                self.step += 1
                peers = self._peer["arr"][b]
            else:
                peers = None
            if P > 1 and self.chunks > 1 and Ls % self.chunks == 0 and a.is_cuda:
                # pipeline: k3 of chunk i+1 (HBM-bound, main stream) runs under the fused k2+exchange of chunk i
                # (NVLink-bound, side stream)
                main = torch.cuda.current_stream(a.device)
                if self._side is None:
                    # k3 runs on a HIGH-priority stream so its CTAs are scheduled ahead of the pending CTAs of the
                    # long NVLink-bound exchange kernel of the previous chunk
                    self._side = torch.cuda.Stream(device=a.device, priority=-1)
                    self._evs = [torch.cuda.Event() for _ in range(self.chunks)]
                hp, per = self._side, Ls // self.chunks
                hp.wait_stream(main)
                esz = a.element_size()
                for i in range(self.chunks):
                    ptr = a.data_ptr() + i * per * R * Cn * 2 * esz
                    _lib.check(self.lib.jtb_lines_c2c_device(self.prec, self.dev, C.c_void_p(ptr), Cn, per * R, 1, 0, Cn,
                                                             1, 0, 1.0, C.c_void_p(hp.cuda_stream)))
                    self._evs[i].record(hp)
                    main.wait_event(self._evs[i])
                    _lib.check(self.lib.jtb_fft3d_k2_scatter_chunk(self.prec, self.dev, C.c_void_p(ptr), per,
                                                                   self.rank * Ls + i * per, R, Cn, P, peers, 0,
                                                                   C.c_void_p(main.cuda_stream)))
            else:
                _lib.check(self.lib.jtb_fft2d_slices_device(self.prec, self.dev, C.c_void_p(a.data_ptr()), Ls, R, Cn, P,
                                                            self.rank, peers, 0, stream))
            if P == 1:
                self._lines(a, S, R * Cn, R * Cn, 1, S * R * Cn, R * Cn)  # k1: across slices
                return a
            _lib.check(self.lib.jtb_peer_barrier(self.dev, self._peer["arr"][2], P, self.rank, self.step, stream))
            recv = self._recv_tensor(b)
            self._lines(recv, S, Rh * Cn, Rh * Cn, 1, S * Rh * Cn, Rh * Cn)
            return recv
        self._lines(a, Cn, Ls * R, 1, 0, Cn, 1)                       # k3: contiguous rows
        self._lines(a, R, Cn * Ls, Cn, 1, R * Cn, Cn)                 # k2: columns inside each slice
        if P == 1:
            self._lines(a, S, R * Cn, R * Cn, 1, S * R * Cn, R * Cn)  # k1: across slices
            return a
        # re-slab over k2: block (ls, h, r, c) -> peer h
        send = a.view(Ls, P, Rh, 2 * Cn).permute(1, 0, 2, 3).contiguous()
        recv = work if work is not None else torch.empty_like(send)
        dist.all_to_all_single(recv.view(-1), send.view(-1), group=self.group)
        # recv is [g][ls][r][c] = [S][Rh][C]
        self._lines(recv, S, Rh * Cn, Rh * Cn, 1, S * Rh * Cn, Rh * Cn)
        return recv

    def inverse(self, b: torch.Tensor, scale: bool = True, work: torch.Tensor | None = None) -> torch.Tensor:
        """complexInverse of a k2-slabbed spectrum: ``b`` is this rank's [S][Rh][C] block as ``forward`` returns it
        (transformed in place for the k1 pass); returns the k1-slabbed local slab [Ls][R][C] of the inverse
        transform (``b`` itself when P == 1).  ``scale`` divides by S*R*C like the reference
        (fft/DoubleFFT_3D.java:744-760).  The re-slabbing runs as one all-to-all (NCCL on GPUs, gloo on CPU): the
        send side needs no packing (the block is already ordered by destination rank)."""
        S, R, Cn, P, Ls, Rh = self.S, self.R, self.Cn, self.P, self.Ls, self.Rh
        sc = 1.0 / (float(S) * float(R) * float(Cn)) if scale else 1.0
        if P == 1:
            self._lines(b, S, R * Cn, R * Cn, 1, S * R * Cn, R * Cn, inverse=True)
            self._lines(b, R, Cn * S, Cn, 1, R * Cn, Cn, inverse=True)
            self._lines(b, Cn, S * R, 1, 0, Cn, 1, inverse=True, scale=sc)
            return b
        if self.exchange == "p2p" and self._peer is not None and b.is_cuda:
            # k1 pass whose stores ARE the second all-to-all (peer stores into the owners' [Ls][R][C] slabs)
            stream = C.c_void_p(torch.cuda.current_stream(b.device).cuda_stream)
            buf = self.step & 1
            rc = self.lib.jtb_fft3d_k1_scatter(self.prec, self.dev, C.c_void_p(b.data_ptr()), S, Rh, Cn, P, self.rank,
                                               self._peer["arr"][buf], 1, stream)
            if rc == _lib.OK:
                self.step += 1
                _lib.check(self.lib.jtb_peer_barrier(self.dev, self._peer["arr"][2], P, self.rank, self.step, stream))
                local = self._recv_tensor(buf)
                self._lines(local, R, Cn * Ls, Cn, 1, R * Cn, Cn, inverse=True)
                self._lines(local, Cn, Ls * R, 1, 0, Cn, 1, inverse=True, scale=sc)
                return local
            if rc != _lib.ERR_UNSUPPORTED:      # no fused kernel for this shape is the same answer on every rank
                _lib.check(rc)
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
