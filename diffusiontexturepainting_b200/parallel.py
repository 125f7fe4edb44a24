"""Multi-GPU plumbing for independent stamps: one process per GPU, weights broadcast once over NCCL at
load, per-batch scatter of canvases / gather of results. No data-path collective inside a stamp: stamps are independent
(each is a function of its own canvas and the shared brush conditioning), so the path shards embarrassingly."""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None):
    """Initialise torch.distributed from the torchrun environment; returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shard_range(n_items: int, rank: int, world: int):
    """Contiguous balanced partition of n_items stamps: the first n_items % world ranks take one extra."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def broadcast_packed(packed: Optional[Dict[str, torch.Tensor]], device, src: int = 0) -> Dict[str, torch.Tensor]:
    """Rank `src` holds the packed weight dict; every rank returns an identical dict (tensors on `device`). One flat
    buffer per dtype is broadcast (fp16 matrices ~1.9 GB, fp32 vectors a few MB)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return {k: v.to(device) for k, v in packed.items()}
    meta = [[(k, tuple(v.shape), str(v.dtype)) for k, v in packed.items()]] if dist.get_rank() == src else [None]
    dist.broadcast_object_list(meta, src=src)
    meta = meta[0]
    out: Dict[str, torch.Tensor] = {}
    for dt in (torch.float16, torch.float32):
        keys = [(k, shp) for k, shp, d in meta if d == str(dt)]
        total = sum(int(torch.Size(shp).numel()) for _, shp in keys)
        flat = torch.empty(total, dtype=dt, device=device)
        if dist.get_rank() == src:
            off = 0
            for k, shp in keys:
                n = int(torch.Size(shp).numel())
                flat[off:off + n].copy_(packed[k].reshape(-1))
                off += n
        dist.broadcast(flat, src=src)
        off = 0
        for k, shp in keys:
            n = int(torch.Size(shp).numel())
            out[k] = flat[off:off + n].view(shp)
            off += n
    return out


def _shard_sizes(total: Optional[int], world: int, local_n: int):
    if total is None:
        return [local_n] * world
    return [b - a for a, b in (shard_range(total, r, world) for r in range(world))]


def scatter_stamps(full: Optional[torch.Tensor], shape_per_rank, dtype, device, src: int = 0,
                   total: Optional[int] = None) -> torch.Tensor:
    """Rank `src` holds (n, ...) stamps; rank r receives rows shard_range(n, r, world) into a buffer of `shape_per_rank`.
    `total` = n must be passed (on every rank) when n is not a multiple of the world size: the shards are then uneven (the
    first n % world ranks hold one more) and travel point to point; an even split is one scatter."""
    out = torch.empty(shape_per_rank, dtype=dtype, device=device)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out.copy_(full)
        return out
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = _shard_sizes(total, world, int(out.shape[0]))
    if sizes[rank] != out.shape[0]:
        raise ValueError(f"rank {rank}: shard of {sizes[rank]} stamps does not fit a buffer of {out.shape[0]}")
    if len(set(sizes)) == 1:
        if rank == src and full.shape[0] != world * sizes[0]:
            raise ValueError(f"{full.shape[0]} stamps do not split evenly over {world} ranks: pass total=")
        chunks = list(full.chunk(world)) if rank == src else None
        dist.scatter(out, chunks, src=src)
        return out
    if rank == src:
        off, reqs = 0, []
        for r, n in enumerate(sizes):
            part = full[off:off + n]
            off += n
            if r == src:
                out.copy_(part)
            elif n > 0:
                reqs.append(dist.isend(part.contiguous(), dst=r))
        for q in reqs:
            q.wait()
    elif out.shape[0] > 0:
        dist.recv(out, src=src)
    return out


def gather_stamps(local: torch.Tensor, dst: int = 0, total: Optional[int] = None) -> Optional[torch.Tensor]:
    """Inverse of scatter_stamps: rank `dst` returns the stamps of all ranks in rank order (None elsewhere)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = _shard_sizes(total, world, int(local.shape[0]))
    if len(set(sizes)) == 1:
        bufs = [torch.empty_like(local) for _ in range(world)] if rank == dst else None
        dist.gather(local, bufs, dst=dst)
        return torch.cat(bufs) if bufs is not None else None
    if rank == dst:
        bufs = [torch.empty((n,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for n in sizes]
        bufs[dst].copy_(local)
        for r, n in enumerate(sizes):
            if r != dst and n > 0:
                dist.recv(bufs[r], src=r)
        return torch.cat(bufs)
    if local.shape[0] > 0:
        dist.send(local.contiguous(), dst=dst)
    return None
