"""Multi-GPU plumbing for the outer layers: one process per GPU.

theta-particles (SMC^2, MBP-IBIS) and chains (pMCMC) are partitioned contiguously over ranks.  The small per-particle
scalars (gx, aw) are exchanged with an all-gather; resampled filters move between ranks with one all-to-all of packed
int32 population blocks.  On GPUs every exchange goes through the C ABI (dpomp_comm_*, dpomp_pf_partial_allgather,
dpomp_pf_resample_migrate, dpomp_mbp_resample_migrate: NCCL inside libdpomp on the handles' own streams) -- the same entry
points a Julia host binds; torch.distributed only bootstraps the NCCL unique id.  Under the gloo backend (the CPU tests of
the host logic, world size 2) the exchanges run over torch.distributed instead.  Every rank draws the same host random
numbers (same seed), so all ranks take identical outer-layer decisions without further communication, and particle-filter
streams are keyed by the GLOBAL filter index, which makes results independent of the number of ranks.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np


class Comm:
    """Thin wrapper over torch.distributed; `Comm(None)` is the single-process communicator (no torch needed)."""

    def __init__(self, group="world"):
        self.dist = None
        self.rank, self.world = 0, 1
        if group is not None:
            import torch.distributed as dist

            if dist.is_available() and dist.is_initialized():
                self.dist = dist
                self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._ag_bufs = {}  # staging buffers of allgather_f64, by block size
        self.device = None  # torch device of the exchange buffers (cuda:k under NCCL, cpu under gloo)
        self._h = None      # dpomp_comm handle: NCCL inside libdpomp (set under the nccl backend)
        if self.dist is not None:
            import torch

            self.device = (torch.device("cuda", torch.cuda.current_device())
                           if self.dist.get_backend() == "nccl" else torch.device("cpu"))
            if self.dist.get_backend() == "nccl":
                self._create_library_comm(torch.cuda.current_device())

    def _create_library_comm(self, device: int) -> None:
        """rank 0 draws the NCCL unique id through the C ABI, torch.distributed broadcasts the bytes, every rank joins."""
        import ctypes as C

        from . import _capi

        nbytes = 128  # DPOMP_UNIQUE_ID_BYTES
        box = [None]
        if self.rank == 0:
            buf = (C.c_ubyte * nbytes)()
            _capi.check(_capi.lib().dpomp_comm_unique_id(buf, nbytes))
            box[0] = bytes(buf)
        self.dist.broadcast_object_list(box, src=0)
        uid = (C.c_ubyte * nbytes).from_buffer_copy(box[0])
        h = C.c_void_p()
        _capi.check(_capi.lib().dpomp_comm_create(uid, nbytes, self.rank, self.world, device, C.byref(h)))
        self._h = h

    def __del__(self):
        try:
            from . import _capi

            for name in ("_h", "_h_single"):
                h = getattr(self, name, None)
                if h:
                    _capi.lib().dpomp_comm_destroy(h)
                    setattr(self, name, None)
        except Exception:
            pass

    @property
    def handle(self):
        """dpomp_comm handle (None when the exchanges run over torch.distributed / gloo or in a single process)."""
        return self._h

    def library_handle(self):
        """dpomp_comm handle for entry points that always take one: the NCCL communicator under the nccl backend, a
        world-size-1 communicator (no NCCL) in a single process, None under gloo."""
        if self._h is None and self.dist is None:
            import ctypes as C

            from . import _capi

            h = C.c_void_p()
            _capi.check(_capi.lib().dpomp_comm_create(None, 0, 0, 1, -1, C.byref(h)))
            self._h_single = h
            return h
        return self._h

    # -- partition -------------------------------------------------------------------------------------------
    def bounds(self, n: int) -> Tuple[int, int]:
        """Contiguous block [lo, hi) of `n` items owned by this rank (sizes differ by at most one)."""
        return partition_bounds(n, self.world, self.rank)

    def owner(self, n: int, idx: np.ndarray) -> np.ndarray:
        return partition_owner(n, self.world, idx)

    # -- collectives -----------------------------------------------------------------------------------------
    def allgather_f64(self, local: np.ndarray, n_total: int) -> np.ndarray:
        """Concatenate each rank's contiguous block of doubles (block sizes per `bounds`)."""
        local = np.ascontiguousarray(local, dtype=np.float64)
        if self.dist is None:
            return local
        if self._h is not None:  # C ABI: pinned staging + ncclAllGather inside the library
            from . import _capi

            width = local.shape[1] if local.ndim == 2 else 1
            out = np.empty((n_total, width) if local.ndim == 2 else n_total, dtype=np.float64)
            _capi.check(_capi.lib().dpomp_comm_allgather_f64(self._h, _capi.ptr(local) if local.size else None, int(n_total),
                                                             int(width), _capi.ptr(out)))
            return out
        import torch

        sizes = [partition_bounds(n_total, self.world, r) for r in range(self.world)]
        width = local.shape[1] if local.ndim == 2 else 1
        maxn = max(hi - lo for lo, hi in sizes)
        # one staged copy each way: pinned host block -> device, all_gather_into_tensor, device -> pinned host
        key = (maxn * width, self.world)
        bufs = self._ag_bufs.get(key)
        if bufs is None:
            pin = self.device.type == "cuda"
            bufs = (torch.zeros(maxn * width, dtype=torch.float64, pin_memory=pin),
                    torch.zeros(maxn * width, dtype=torch.float64, device=self.device),
                    torch.zeros(self.world * maxn * width, dtype=torch.float64, device=self.device),
                    torch.zeros(self.world * maxn * width, dtype=torch.float64, pin_memory=pin))
            self._ag_bufs[key] = bufs
        h_in, d_in, d_out, h_out = bufs
        h_in[: local.size] = torch.from_numpy(local.reshape(-1))
        d_in.copy_(h_in, non_blocking=True)
        if self.dist.get_backend() == "nccl":
            self.dist.all_gather_into_tensor(d_out, d_in)
        else:  # gloo builds without all_gather_into_tensor
            parts_t = list(d_out.view(self.world, -1).unbind(0))
            self.dist.all_gather(parts_t, d_in)
        h_out.copy_(d_out, non_blocking=True)
        if self.device.type == "cuda":
            torch.cuda.current_stream().synchronize()
        allv = h_out.numpy().reshape(self.world, maxn * width)
        res = np.concatenate([allv[r, : (hi - lo) * width] for r, (lo, hi) in enumerate(sizes)])
        return res.reshape(-1, width) if local.ndim == 2 else res

    def all_to_all_blocks(self, send, send_counts: Sequence[int], recv_counts: Sequence[int], block_words: int):
        """Exchange packed int32 filter blocks.  `send` is a torch int32 tensor of sum(send_counts) * block_words words,
        ordered by destination rank; returns the received tensor ordered by source rank."""
        import torch

        recv = torch.empty(int(sum(recv_counts)) * block_words, dtype=torch.int32, device=send.device)
        if self.dist is None:
            recv.copy_(send)
            return recv
        rs = [int(c) * block_words for c in recv_counts]
        ss = [int(c) * block_words for c in send_counts]
        if send.device != self.device:  # gloo with CUDA-resident filters (tests): stage through the host
            tmp = torch.empty(recv.numel(), dtype=torch.int32, device=self.device)
            self._all_to_all(tmp, send.to(self.device), rs, ss)
            recv.copy_(tmp)
        else:
            self._all_to_all(recv, send, rs, ss)
        return recv

    def _all_to_all(self, recv, send, rs, ss):
        import torch

        if self.dist.get_backend() == "gloo":  # gloo has no all_to_all_single on every build: pairwise send/recv
            ro = np.concatenate(([0], np.cumsum(rs))); so = np.concatenate(([0], np.cumsum(ss)))
            reqs = []
            for r in range(self.world):
                if r == self.rank:
                    recv[ro[r]:ro[r + 1]] = send[so[r]:so[r + 1]]
                    continue
                if ss[r]:
                    reqs.append(self.dist.isend(send[so[r]:so[r + 1]].contiguous(), r))
            for r in range(self.world):
                if r != self.rank and rs[r]:
                    buf = torch.empty(rs[r], dtype=torch.int32)
                    self.dist.recv(buf, r)
                    recv[ro[r]:ro[r + 1]] = buf
            for q in reqs:
                q.wait()
        else:
            self.dist.all_to_all_single(recv, send, rs, ss)
            # NCCL only enqueues the exchange on torch's stream; the library reads `recv` on its own stream next
            torch.cuda.current_stream().synchronize()

    def all_to_all_v(self, send, send_counts: Sequence[int], recv_counts: Sequence[int]):
        """Variable-size all-to-all of a 1-d torch tensor of any dtype; counts are element counts per rank."""
        import torch

        recv = torch.empty(int(sum(recv_counts)), dtype=send.dtype, device=send.device)
        if self.dist is None:
            recv.copy_(send)
            return recv
        rs, ss = [int(c) for c in recv_counts], [int(c) for c in send_counts]
        if self.dist.get_backend() == "gloo":
            host_send = send.cpu()
            host_recv = torch.empty(recv.numel(), dtype=send.dtype)
            ro = np.concatenate(([0], np.cumsum(rs))); so = np.concatenate(([0], np.cumsum(ss)))
            reqs = []
            for r in range(self.world):
                if r == self.rank:
                    host_recv[ro[r]:ro[r + 1]] = host_send[so[r]:so[r + 1]]
                elif ss[r]:
                    reqs.append(self.dist.isend(host_send[so[r]:so[r + 1]].contiguous(), r))
            for r in range(self.world):
                if r != self.rank and rs[r]:
                    buf = torch.empty(rs[r], dtype=send.dtype)
                    self.dist.recv(buf, r)
                    host_recv[ro[r]:ro[r + 1]] = buf
            for q in reqs:
                q.wait()
            recv.copy_(host_recv)
        else:
            self.dist.all_to_all_single(recv, send, rs, ss)
            torch.cuda.current_stream().synchronize()  # consumers run on the library's own stream
        return recv

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


def partition_bounds(n: int, world: int, rank: int) -> Tuple[int, int]:
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def partition_owner(n: int, world: int, idx: np.ndarray) -> np.ndarray:
    idx = np.asarray(idx, dtype=np.int64)
    base, extra = divmod(n, world)
    cut = extra * (base + 1)
    return np.where(idx < cut, idx // (base + 1), extra + (idx - cut) // max(base, 1)).astype(np.int64)


def migration_plan(nidx0: np.ndarray, n: int, world: int, rank: int):
    """Who sends what after the outer resample `new[p] <- old[nidx0[p]]` (0-based global indices).

    Returns (local_src, send_slots, send_counts, recv_slots, recv_counts):
      local_src[k]   local source slot for local destination k, or k itself when the ancestor is remote (placeholder)
      send_slots     local slots to pack, ordered by destination rank then by destination index
      recv_slots     local destination slots of the received blocks, ordered by source rank then by destination index
    """
    nidx0 = np.asarray(nidx0, dtype=np.int64)
    dst_owner = partition_owner(n, world, np.arange(n))
    src_owner = partition_owner(n, world, nidx0)
    lo, hi = partition_bounds(n, world, rank)
    local_src = np.arange(hi - lo, dtype=np.int64)
    mine = np.arange(lo, hi)
    same = src_owner[mine] == rank
    local_src[same] = nidx0[mine][same] - lo
    send_slots: List[int] = []
    send_counts = [0] * world
    recv_slots: List[int] = []
    recv_counts = [0] * world
    for r in range(world):
        if r == rank:
            continue
        out = np.nonzero((dst_owner == r) & (src_owner == rank))[0]  # destinations on r fed by my filters
        send_slots.extend((nidx0[out] - lo).tolist())
        send_counts[r] = len(out)
        inc = np.nonzero((dst_owner == rank) & (src_owner == r))[0]  # my destinations fed by r
        recv_slots.extend((inc - lo).tolist())
        recv_counts[r] = len(inc)
    return local_src, np.asarray(send_slots, dtype=np.int64), send_counts, np.asarray(recv_slots, dtype=np.int64), recv_counts
