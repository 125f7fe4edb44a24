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


def scatter_stamps(full: Optional[torch.Tensor], shape_per_rank, dtype, device, src: int = 0) -> torch.Tensor:
    """Rank `src` holds (world * b, ...) ; every rank receives its (b, ...) slice."""
    out = torch.empty(shape_per_rank, dtype=dtype, device=device)
    if not dist.is_initialized() or dist.get_world_size() == 1:
        out.copy_(full)
        return out
    chunks = list(full.chunk(dist.get_world_size())) if dist.get_rank() == src else None
    dist.scatter(out, chunks, src=src)
    return out


def gather_stamps(local: torch.Tensor, dst: int = 0) -> Optional[torch.Tensor]:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    bufs = [torch.empty_like(local) for _ in range(dist.get_world_size())] if dist.get_rank() == dst else None
    dist.gather(local, bufs, dst=dst)
    return torch.cat(bufs) if bufs is not None else None
