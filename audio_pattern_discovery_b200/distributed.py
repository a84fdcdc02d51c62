"""One process per GPU: shard the work units of the all-pairs matrix over the ranks of a
torch.distributed process group and assemble the full matrix on every rank with ONE
all-gather of the packed per-shard results (NCCL over NVLink/NVSwitch), followed by the
library's scatter kernel.  Pairs are independent, so the all-gather is the only
collective on the path.

torch is used for device buffers, streams and the process group only.
"""
import numpy as np
import torch
import torch.distributed as dist

from .alignments import APD_MODE_STRICT, Context


class _DeviceFloats:
    """Zero-copy torch view of a float32 device buffer owned by the library (__cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f4", "data": (int(ptr), False), "version": 3}


class ShardedAligner:
    def __init__(self, seqs, device=None, group=None, mode=APD_MODE_STRICT):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device)
        self.mode = mode
        # A real (non-default) stream: the library treats stream 0 as "the context's own
        # stream", and torch events only see the stream they are recorded on.
        import time
        t0 = time.perf_counter()
        self.stream = torch.cuda.Stream(self.device)
        self.ctx = Context(device)
        t1 = time.perf_counter()
        self.set_sequences(seqs)           # every rank holds the whole arena (<= ~0.4 GB)
        self.t_create, self.t_set_sequences = t1 - t0, time.perf_counter() - t1
        self.ctx.set_shard(self.rank, self.world)
        self.n = self.ctx.n
        self._packed = None
        self._gathered = None
        self._matrix = None

    def set_sequences(self, seqs):
        """Every rank ends up with the same packed arena.  One rank packs and uploads it; the others only derive the
        (identical) layout from the lengths and receive the bytes with one NCCL broadcast over NVLink -- instead of
        `world` processes packing and uploading the same 0.4 GB side by side on one host."""
        if self.world == 1:
            self.ctx.set_sequences(seqs)
        else:
            uploader = self.rank == 0
            lens = [len(s) for s in seqs]
            dim = (np.asarray(seqs[0]).shape[1] if np.asarray(seqs[0]).ndim == 2 else 1) if len(seqs) else 1
            self.ctx.set_sequences(seqs) if uploader else self.ctx.set_sequences_layout(lens, dim)
            self.ctx.synchronize()                      # the upload (tables on the other ranks) is complete
            ptr, n = self.ctx.arena_device()
            arena = torch.as_tensor(_DeviceFloats(ptr, n), device=self.device)
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            with torch.cuda.stream(self.stream):
                dist.broadcast(arena, src=src, group=self.group)
            if not uploader:
                self.ctx.arena_commit()
        self.n = self.ctx.n

    def align_all_device(self, pct, ins=1.0, dele=1.0, mat=1.0):
        """-> (n, n) float32 CUDA tensor holding the full matrix on this rank."""
        k = self.ctx.packed_len(pct, self.mode)
        if self._packed is None or self._packed.numel() != k:
            self.stream.synchronize()      # an earlier call may still be using the buffers that are about to be replaced
            self._packed = torch.empty(k, dtype=torch.float32, device=self.device)
            self._gathered = (torch.empty(k * self.world, dtype=torch.float32, device=self.device)
                              if self.world > 1 else self._packed)
        if self._matrix is None or self._matrix.shape[0] != self.n:
            self.stream.synchronize()
            self._matrix = torch.empty((self.n, self.n), dtype=torch.float32, device=self.device)
        stream = self.stream
        with torch.cuda.stream(stream):
            self.ctx.align_packed(pct, ins, dele, mat, self.mode, self._packed.data_ptr(), stream.cuda_stream)
            if self.world > 1:
                dist.all_gather_into_tensor(self._gathered, self._packed, group=self.group)
            self.ctx.scatter_packed(self._gathered.data_ptr(), self.world, self._matrix.data_ptr(),
                                    stream.cuda_stream)
        return self._matrix

    def align_all(self, pct, ins=1.0, dele=1.0, mat=1.0, out=None, to_host=True):
        """Full matrix as a host array on the ranks that ask for it (to_host)."""
        m = self.align_all_device(pct, ins, dele, mat)
        stream = self.stream
        res = None
        if to_host:
            if out is None:
                out = torch.empty((self.n, self.n), dtype=torch.float32, pin_memory=True)
            with torch.cuda.stream(stream):
                out.copy_(m, non_blocking=True)
            res = out
        self.ctx.synchronize(stream.cuda_stream)
        return res

    def synchronize(self):
        self.ctx.synchronize(self.stream.cuda_stream)

    def percentile_of_matrix(self, perc):
        """numerics::percentile of the assembled device matrix (local work, no collective)."""
        m = self._matrix
        return self.ctx.percentile_device(m.data_ptr(), m.numel(), perc, self.stream.cuda_stream)

    def matrix_device(self):
        """The matrix the last align_all_device / align_all assembled on this rank (no new work)."""
        return self._matrix

    def stats(self):
        return self.ctx.stats()

    def close(self):
        self.ctx.close()


class GroupAligner:
    """ONE process, several GPUs: the drop-in form of the reference's single blocking
    `workers.align_all(&discover)` (src/main.rs:189-195).  The library owns the group
    (apd_create_multi): one upload of the arena + NVLink copies, the pair space dealt over
    the devices, packed results stored peer-to-peer from inside the DTW kernels, the matrix
    copied back by every device in parallel.  No torch, no process group."""

    def __init__(self, seqs, devices="all", mode=APD_MODE_STRICT):
        self.mode = mode
        self.ctx = Context(devices=devices)
        self.world = self.ctx.group_size
        self.ctx.set_sequences(seqs)
        self.n = self.ctx.n

    def set_sequences(self, seqs):
        """Every rank ends up with the same packed arena.  One rank packs and uploads it; the others only derive the
        (identical) layout from the lengths and receive the bytes with one NCCL broadcast over NVLink -- instead of
        `world` processes packing and uploading the same 0.4 GB side by side on one host."""
        if self.world == 1:
            self.ctx.set_sequences(seqs)
        else:
            uploader = self.rank == 0
            lens = [len(s) for s in seqs]
            dim = (np.asarray(seqs[0]).shape[1] if np.asarray(seqs[0]).ndim == 2 else 1) if len(seqs) else 1
            self.ctx.set_sequences(seqs) if uploader else self.ctx.set_sequences_layout(lens, dim)
            self.ctx.synchronize()                      # the upload (tables on the other ranks) is complete
            ptr, n = self.ctx.arena_device()
            arena = torch.as_tensor(_DeviceFloats(ptr, n), device=self.device)
            src = dist.get_global_rank(self.group, 0) if self.group is not None else 0
            with torch.cuda.stream(self.stream):
                dist.broadcast(arena, src=src, group=self.group)
            if not uploader:
                self.ctx.arena_commit()
        self.n = self.ctx.n

    def align_all(self, pct, ins=1.0, dele=1.0, mat=1.0, out=None):
        """-> (n, n) float32 host array (any host memory; pageable is fine)."""
        return self.ctx.align_all(pct, ins, dele, mat, self.mode, out=out)

    def percentile_of_matrix(self, perc):
        return self.ctx.percentile(perc)

    def stats(self):
        return self.ctx.stats()

    def close(self):
        self.ctx.close()


def partition_check(n_units, world):
    """Units u with u % world == rank: sizes per rank (host-side helper for tests)."""
    return [len(range(r, n_units, world)) for r in range(world)]
