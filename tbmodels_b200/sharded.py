"""Multi-GPU evaluation: one process per GPU, contiguous k-point shards.

Every k-point is independent (reference src/tbmodels/_tb_model.py:1109-1132, :1149), so the batch is cut
into ``world_size`` contiguous ranges -- output order stays input order, which the API requires -- and each
rank evaluates its range on its own GPU.  The data path needs no collective; two optional exchange steps are
provided because callers want them:

* ``broadcast_model``  -- ship the packed model from rank 0 to all ranks (one broadcast per array);
* ``eigenval_allgather`` -- NCCL all-gather of the eigenvalue shards so every rank holds ``[n_k, N]``.

Plumbing is ``torch.distributed`` (NCCL on GPUs, gloo in the CPU unit tests where a stand-in evaluator is
injected); the kernels are reached through :class:`tbmodels_b200.Evaluator` as in the single-GPU case.
"""
from __future__ import annotations

import numpy as np

from ._pack import PackedModel, pack_arrays


def shard_bounds(n_k: int, world_size: int, rank: int):
    """Contiguous range ``[lo, hi)`` of rank ``rank``; the first ``n_k % world_size`` ranks get one extra point."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    base, rem = divmod(int(n_k), world_size)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def max_shard(n_k: int, world_size: int) -> int:
    return -(-int(n_k) // world_size)


def broadcast_model(packed, src: int = 0, group=None, device=None) -> PackedModel:
    """Replicate a packed model from rank ``src`` (other ranks may pass ``None``)."""
    import torch
    import torch.distributed as dist

    rank = dist.get_rank(group)
    dev = torch.device("cpu") if device is None else torch.device(device)
    header = torch.zeros(3, dtype=torch.int64, device=dev)
    if rank == src:
        header[:] = torch.tensor([packed.n_R, packed.size, packed.dim], dtype=torch.int64)
    dist.broadcast(header, src, group=group)
    n_R, size, dim = (int(x) for x in header.tolist())
    if rank == src:
        R = torch.from_numpy(packed.R.copy()).to(dev)
        hop = torch.from_numpy(packed.hop.view(np.float64).copy()).to(dev)
        pos = torch.from_numpy(packed.pos.copy()).to(dev)
    else:
        R = torch.empty((n_R, dim), dtype=torch.int32, device=dev)
        hop = torch.empty((n_R, size, size * 2), dtype=torch.float64, device=dev)
        pos = torch.empty((size, dim), dtype=torch.float64, device=dev)
    for t in (R, hop, pos):
        if t.numel():
            dist.broadcast(t, src, group=group)
    if rank == src:
        return packed
    hop_c = hop.cpu().numpy().view(np.complex128).reshape(n_R, size, size)
    return pack_arrays(R.cpu().numpy(), hop_c, pos.cpu().numpy())


class PeerGather:
    """Result buffer ``[n_rows, n_cols]`` float64 that exists on every GPU of the group and is peer-mapped into every
    process (``torch.distributed._symmetric_memory``: CUDA VMM handles exchanged once at rendezvous).  It is the
    destination of the FUSED gather: ``Evaluator.eigenval_push_device`` writes a rank's rows into
    its own copy and, chunk by chunk, into all peers' copies with plain stores over NVLink / NVSwitch while the next chunk
    computes; :meth:`barrier` (a device-side barrier through the handle's signal pads, enqueued on the current stream)
    then makes every rank's rows visible everywhere.  No NCCL call on the data path."""

    def __init__(self, n_rows: int, n_cols: int, group=None, device=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group if group is not None else dist.group.WORLD
        self.rank = dist.get_rank(self.group)
        self.world_size = dist.get_world_size(self.group)
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.buffer = symm_mem.empty((int(n_rows), int(n_cols)), dtype=torch.float64, device=dev)
        self.handle = symm_mem.rendezvous(self.buffer, self.group)
        ptrs = list(self.handle.buffer_ptrs)
        if len(ptrs) != self.world_size or int(ptrs[self.rank]) != self.buffer.data_ptr():
            raise RuntimeError("symmetric-memory rendezvous returned an unexpected pointer table")
        self.peer_ptrs = [int(p) for r, p in enumerate(ptrs) if r != self.rank]

    def barrier(self) -> None:
        self.handle.barrier()


class ShardedEvaluator:
    """Shards a k-batch over the ranks of a process group.

    ``local`` is the per-rank evaluator: an object with ``eigenval_device(k) -> [n, N]`` and
    ``hamilton_device(k, convention) -> [n, N, N]`` working on torch tensors of this rank's device.  By default
    it is a CUDA :class:`~tbmodels_b200.Evaluator` on ``LOCAL_RANK``'s GPU (no fallback).
    """

    def __init__(self, packed: PackedModel, group=None, local=None, device=None):
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("torch.distributed is not initialised (launch with torchrun, one process per GPU)")
        self.group = group
        self.rank = dist.get_rank(group)
        self.world_size = dist.get_world_size(group)
        self.size = packed.size
        self.dim = packed.dim
        if local is None:
            from ._evaluator import Evaluator

            local = Evaluator(packed, device=device)
        self.local = local

    def bounds(self, n_k: int):
        return shard_bounds(n_k, self.world_size, self.rank)

    def _check(self):
        """Surface deferred device errors (QL non-convergence) of the asynchronous device-pointer calls."""
        check = getattr(self.local, "check", None)
        if check is not None:
            check()

    def eigenval_local(self, k_all, check=True):
        """Evaluate this rank's contiguous shard of the replicated ``k_all`` -> ``(lo, hi, eig[lo:hi])``.

        ``check=True`` synchronises and raises on a deferred device error; pass ``False`` to stay asynchronous (and call
        ``self.local.check()`` yourself)."""
        lo, hi = self.bounds(k_all.shape[0])
        eig = self.local.eigenval_device(k_all[lo:hi].contiguous())
        if check:
            self._check()
        return lo, hi, eig

    def eigenval_mesh_local(self, dims, shift=None):
        """This rank's contiguous range of LINES (runs along the last dimension) of the regular mesh ``dims``
        -> ``(first_point, last_point, eig)``; see :meth:`Evaluator.eigenval_mesh_device`.  No collective."""
        dims = [int(x) for x in dims]
        n_lines = int(np.prod(dims[:-1])) if len(dims) > 1 else 1
        lo, hi = self.bounds(n_lines)
        eig = self.local.eigenval_mesh_device(dims, shift, first_line=lo, n_lines=hi - lo)
        self._check()
        return lo * dims[-1], hi * dims[-1], eig

    def hamilton_local(self, k_all, convention=2):
        lo, hi = self.bounds(k_all.shape[0])
        return lo, hi, self.local.hamilton_device(k_all[lo:hi].contiguous(), convention=convention)

    def eigenval_allgather_fused(self, k_all):
        """Eigenvalues of the whole batch on every rank with the exchange FUSED behind the eigensolver: every workspace
        chunk a rank finishes is stored straight into all peers' result buffers over NVLink while the next chunk
        computes (``tbk_eigenval_push``), then one device-side barrier.  Shards may be uneven (no padding).  Returns the
        symmetric ``[n_k, N]`` tensor, valid until the next call with the same batch size.  Same bits as
        :meth:`eigenval_allgather` (a k-point's result does not depend on the rank that computed it)."""
        n_k = int(k_all.shape[0])
        key = (n_k, self.size)
        cache = getattr(self, "_peer_gather", None)
        if cache is None or cache[0] != key:
            cache = (key, PeerGather(n_k, self.size, group=self.group, device=k_all.device))
            self._peer_gather = cache
        pg = cache[1]
        lo, hi = self.bounds(n_k)
        pg.barrier()  # nobody is still reading the buffer of the previous call
        self.local.eigenval_push_device(k_all[lo:hi].contiguous(), pg.buffer[lo:hi], pg.peer_ptrs, lo)
        pg.barrier()
        self._check()
        return pg.buffer

    def eigenval_allgather(self, k_all):
        """Eigenvalues of the whole batch on every rank: local shards + one NCCL all-gather (padded, then trimmed)."""
        import torch
        import torch.distributed as dist

        n_k = k_all.shape[0]
        lo, hi, eig = self.eigenval_local(k_all)
        nvtx = eig.is_cuda
        if nvtx:
            torch.cuda.nvtx.range_push("tbk:allgather (NCCL)")
        width = max_shard(n_k, self.world_size)
        send = eig
        if hi - lo < width:
            send = torch.zeros((width, self.size), dtype=eig.dtype, device=eig.device)
            send[: hi - lo] = eig
        gathered = torch.empty((self.world_size * width, self.size), dtype=eig.dtype, device=eig.device)
        if gathered.is_cuda:
            dist.all_gather_into_tensor(gathered, send.contiguous(), group=self.group)
        else:  # gloo (CPU unit tests)
            parts = list(gathered.view(self.world_size, width, self.size).unbind(0))
            dist.all_gather(parts, send.contiguous(), group=self.group)
        if nvtx:
            torch.cuda.nvtx.range_pop()
        if width * self.world_size == n_k:
            return gathered
        out = torch.empty((n_k, self.size), dtype=eig.dtype, device=eig.device)
        for r in range(self.world_size):
            rlo, rhi = shard_bounds(n_k, self.world_size, r)
            out[rlo:rhi] = gathered[r * width : r * width + (rhi - rlo)]
        return out
