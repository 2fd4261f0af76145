"""world_size-2 gloo tests of the multi-GPU host logic: frame sharding, the gradient
all-reduce (mean) and the per-rank schedules; the same functions run over NCCL on the GPUs."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from thermo_nerf_b200 import dist as tdist
from thermo_nerf_b200.engine import exponential_decay_lr


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        res = {}
        # gradient arena: mean over ranks of rank-dependent values
        arena = torch.arange(1000, dtype=torch.float32) * (rank + 1)
        tdist.allreduce_mean_(arena)
        res["arena_ok"] = bool(torch.allclose(arena, torch.arange(1000, dtype=torch.float32) * (1 + world) / 2))
        # plugin route: parameter gradients that are views of one arena -> one collective
        base = torch.zeros(16 + 8)
        ps = [torch.nn.Parameter(torch.zeros(4, 4)), torch.nn.Parameter(torch.zeros(6))]
        ps[0].grad = base[:16].view(4, 4)
        ps[1].grad = base[16:22]
        base.fill_(float(rank))
        tdist.allreduce_mean_grads_(ps)
        res["grads_ok"] = bool(torch.allclose(ps[0].grad, torch.full((4, 4), (world - 1) / 2)) and
                               torch.allclose(ps[1].grad, torch.full((6,), (world - 1) / 2)))
        # independent gradients (no shared arena)
        ps2 = [torch.nn.Parameter(torch.zeros(3)), torch.nn.Parameter(torch.zeros(5)), torch.nn.Parameter(torch.zeros(2))]
        ps2[0].grad = torch.full((3,), float(rank + 1))
        ps2[1].grad = torch.full((5,), float(10 * rank))
        tdist.allreduce_mean_grads_(ps2)  # ps2[2] has no gradient: skipped
        res["sep_ok"] = bool(torch.allclose(ps2[0].grad, torch.full((3,), (1 + world) / 2)) and
                             torch.allclose(ps2[1].grad, torch.full((5,), 10 * (world - 1) / 2)) and
                             ps2[2].grad is None)
        # frames: rank r renders r, r+world, ...; rank 0 reassembles them in order
        n = 7
        mine = tdist.shard_frames(n, rank, world)
        frames = torch.tensor(mine, dtype=torch.float32).view(-1, 1, 1).expand(-1, 2, 3).contiguous()
        got = tdist.gather_frames(frames, n)
        if rank == 0:
            res["frames_ok"] = bool(torch.equal(got[:, 0, 0], torch.arange(n, dtype=torch.float32)))
        res["seed"] = tdist.rank_seed(77, rank, 3)
        # the step-dependent schedules are pure functions of the step: identical on every rank
        t = torch.tensor([exponential_decay_lr(123)], dtype=torch.float64)
        lo, hi = t.clone(), t.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        res["sched_ok"] = bool(lo.item() == hi.item())
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_shard_frames_partition():
    for world in (1, 2, 3, 8):
        for n in (0, 1, 7, 300):
            parts = [tdist.shard_frames(n, r, world) for r in range(world)]
            flat = sorted(i for p in parts for i in p)
            assert flat == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    with pytest.raises(ValueError):
        tdist.shard_frames(4, 2, 2)


def test_single_process_is_identity():
    t = torch.ones(5)
    assert tdist.allreduce_mean_(t) is t and torch.equal(t, torch.ones(5))
    assert tdist.world_info() == (0, 1)
    f = torch.zeros(3, 2)
    assert tdist.gather_frames(f, 3) is f


@pytest.mark.timeout(120)
def test_gloo_world_size_2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict(q.get(timeout=100) for _ in range(world))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for r in range(world):
        assert out[r]["arena_ok"] and out[r]["grads_ok"] and out[r]["sep_ok"] and out[r]["sched_ok"]
    assert out[0]["frames_ok"]
    assert out[0]["seed"] != out[1]["seed"]
