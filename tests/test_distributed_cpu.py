"""CPU, world_size 2 over gloo: the multi-GPU plumbing of the stamp path (weight broadcast at load, per-batch scatter of
canvases and gather of results, balanced stamp sharding). No data-path collective exists inside a stamp."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from diffusiontexturepainting_b200 import parallel as par
    from diffusiontexturepainting_b200 import weights as W
    r, w, _ = par.init_from_env("gloo")
    assert (r, w) == (rank, world)
    cfg = W.tiny_config()
    packed = None
    if rank == 0:
        u, v, e = W.synth_model(cfg, 11)
        packed = {"unet." + k: t for k, t in W.pack_unet(W.merge_lora(u)).items()}
        packed.update({"vae." + k: t for k, t in W.pack_vae(v).items()})
    got = par.broadcast_packed(packed, torch.device("cpu"), src=0)
    # every rank must hold bit-identical tensors: compare a checksum of checksums
    digest = torch.tensor([sum(float(t.double().sum()) for t in got.values()), float(len(got))], dtype=torch.float64)
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    assert all(torch.equal(all_d[0], d) for d in all_d)
    if rank == 0:
        assert all(torch.equal(got[k], packed[k]) for k in packed)
    # scatter canvases, "process" them locally, gather results in rank order
    B = 3
    full = torch.arange(world * B * 4 * 4 * 4, dtype=torch.uint8).view(world * B, 4, 4, 4) if rank == 0 else None
    mine = par.scatter_stamps(full, (B, 4, 4, 4), torch.uint8, torch.device("cpu"), src=0)
    lo, hi = par.shard_range(world * B, rank, world)
    expect = torch.arange(world * B * 64, dtype=torch.uint8).view(world * B, 4, 4, 4)[lo:hi]
    assert torch.equal(mine, expect)
    res = par.gather_stamps(mine[..., :3].contiguous() + 1, dst=0)
    if rank == 0:
        assert torch.equal(res, torch.arange(world * B * 64, dtype=torch.uint8).view(world * B, 4, 4, 4)[..., :3] + 1)
    # uneven split: 5 stamps over 2 ranks -> 3 + 2 (shard_range), point to point, gathered back in order
    n = 5
    full = torch.arange(n * 8, dtype=torch.float32).view(n, 2, 4) if rank == 0 else None
    lo, hi = par.shard_range(n, rank, world)
    mine = par.scatter_stamps(full, (hi - lo, 2, 4), torch.float32, torch.device("cpu"), src=0, total=n)
    assert torch.equal(mine, torch.arange(n * 8, dtype=torch.float32).view(n, 2, 4)[lo:hi])
    back = par.gather_stamps(mine * 2, dst=0, total=n)
    if rank == 0:
        assert torch.equal(back, torch.arange(n * 8, dtype=torch.float32).view(n, 2, 4) * 2)
        try:
            par.scatter_stamps(torch.zeros(5, 1), (2, 1), torch.float32, torch.device("cpu"), src=0)
            raise AssertionError("uneven split without total= must be refused")
        except ValueError:
            pass
        ret.put("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_broadcast_scatter_gather_world2():
    world = 2
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert ret.get(timeout=5) == "ok"
