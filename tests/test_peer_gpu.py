"""The multi-GPU exchange step fused with Adam over NVLink peer memory (tnf_peer_adam_step).

world_size 1 (runs on the single-GPU box): the fused kernel must equal tnf_adam_step bit for bit.
world_size 2 (skipped without two GPUs): two processes map each other's arenas over CUDA IPC; the result
must equal "mean of the two gradient arenas, then Adam" computed independently, be identical on both ranks,
and the flag barriers must never time out."""

import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref_adam(p, g, m, v, lr, step):
    from thermo_nerf_b200 import functional as F

    F.adam_step([p], [g], [m], [v], [lr], step=step, eps=1e-15)


def test_world1_fused_adam_equals_adam_kernel_bitwise():
    from thermo_nerf_b200.dist import PeerArena

    dev = torch.device("cuda:0")
    n = 4 * 12345
    arena = PeerArena(n, dev)
    assert arena.world == 1 and arena.numel == n and arena.shard == n
    gen = torch.Generator(device="cpu").manual_seed(0)
    P = torch.randn(n, generator=gen).to(dev)
    p_ref, m_ref, v_ref = P.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
    arena.params.copy_(P)
    for step in (1, 2, 3):
        G = (torch.randn(n, generator=gen) * 10 ** torch.randint(-8, 1, (n,), generator=gen).float()).to(dev)
        arena.grads.copy_(G)
        arena.adam_step([(0, n, 1e-2, step, True)])
        _ref_adam(p_ref, G.clone(), m_ref, v_ref, 1e-2, step)
        torch.cuda.synchronize()
        assert torch.equal(arena.params, p_ref)
        assert torch.equal(arena.exp_avg, m_ref) and torch.equal(arena.exp_avg_sq, v_ref)
        assert arena.grads.abs().max().item() == 0.0  # consumed
    # two optimizer groups, the first one without gradients this step: it must not move
    cut = 4 * 1000
    before = arena.params.clone()
    arena.grads.copy_(torch.randn(n, generator=gen).to(dev))
    g2 = arena.grads.clone()
    arena.adam_step([(0, cut, 1e-2, 1, False), (cut, n, 5e-3, 4, True)])
    _ref_adam(p_ref[cut:], g2[cut:].clone(), m_ref[cut:], v_ref[cut:], 5e-3, 4)
    torch.cuda.synchronize()
    assert torch.equal(arena.params[:cut], before[:cut])
    assert torch.equal(arena.params[cut:], p_ref[cut:])
    assert arena.timeouts() == 0


def test_engine_with_fused_exchange_trains():
    """TrainEngine(peer_fused=True) on one GPU: parameters live in the peer arena, the fused kernel replaces the
    two Adam launches, and a fixed batch is fitted as with the plain engine."""
    from oracle import make_synthetic_rays
    from tests.helpers import make_pair
    from thermo_nerf_b200.engine import TrainEngine

    losses = {}
    for fused in (False, True):
        _, model = make_pair(trained_like=False, precision="tc_fp16", log2_field=14, log2_prop=12,
                             camera_optimizer_mode="off")
        model.train()
        eng = TrainEngine(model, peer_fused=fused)
        rays = make_synthetic_rays(1024, num_images=8, seed=5)
        gen = torch.Generator().manual_seed(1)
        gt_rgb = torch.rand((1024, 3), generator=gen).mul(0.2).add(0.4).cuda()
        gt_th = torch.rand((1024,), generator=gen).mul(0.2).add(0.6).cuda()
        jit = torch.rand((3, 1024), generator=gen).cuda()
        o, d, c = rays.origins.cuda(), rays.directions.cuda(), rays.camera_indices.cuda().reshape(-1)
        hist = []
        for _ in range(30):
            ls = eng.step(o, d, c, gt_rgb, gt_th, jitter=jit)
            hist.append(float(ls[0] + ls[3]))
        losses[fused] = hist
        if fused:
            assert eng.arena is not None and eng.arena.timeouts() == 0
            # the module's parameters are views of the arena the kernel writes
            p0 = model.field.mlp_base.encoder.hash_table
            assert p0.data_ptr() == eng.arena.params[eng.offsets[10]:].data_ptr()
    for fused in (False, True):
        assert losses[fused][-1] < 0.5 * losses[fused][0], losses[fused]
    assert abs(losses[True][0] - losses[False][0]) < 1e-6  # identical first forward
    assert abs(losses[True][-1] - losses[False][-1]) < 0.2 * losses[False][-1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q, backend: str = "auto") -> None:
    import torch.distributed as dist

    from thermo_nerf_b200.dist import PeerArena

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["TNF_PEER_BACKEND"] = backend
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n = 64 * 6789  # a multiple of 4 * world for every world size up to 16
        arena = PeerArena(n, dev)
        gen = torch.Generator().manual_seed(123)
        P = torch.randn(n, generator=gen).to(dev)  # identical on every rank
        arena.params.copy_(P)
        p_ref, m_ref, v_ref = P.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        ok = True
        for step in (1, 2, 3, 4):
            g_all = [torch.randn(n, generator=torch.Generator().manual_seed(1000 * step + r)) for r in range(world)]
            arena.grads.copy_(g_all[rank].to(dev))
            arena.adam_step([(0, n, 1e-2, step, True)])
            mean = torch.zeros(n, device=dev)
            for g in g_all:  # rank order, as the kernel sums
                mean += g.to(dev)
            mean *= 1.0 / world
            from thermo_nerf_b200 import functional as F

            F.adam_step([p_ref], [mean], [m_ref], [v_ref], [1e-2], step=step, eps=1e-15)
            torch.cuda.synchronize()
            # peer loads sum in rank order like the reference below: bitwise.  The in-switch reduction (multimem) may
            # sum in another order: exact for two ranks (one addition), last-bit differences beyond
            exact = arena.gather != "multimem" or world == 2
            same_p = torch.equal(arena.params, p_ref) if exact else torch.allclose(arena.params, p_ref, rtol=2e-6, atol=1e-7)
            ok = ok and bool(same_p) and arena.grads.abs().max().item() == 0.0
            lo, hi = arena.shard * rank, arena.shard * (rank + 1)
            same_m = (torch.equal(arena.exp_avg, m_ref[lo:hi]) if exact
                      else torch.allclose(arena.exp_avg, m_ref[lo:hi], rtol=2e-6, atol=1e-7))
            ok = ok and bool(same_m)
            if not exact:
                p_ref.copy_(arena.params)  # keep following the kernel's own trajectory
        # every rank holds the same parameters
        chk = arena.params.double().sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        same = all(float(b) == float(both[0]) for b in both)
        q.put((rank, ok, same, arena.timeouts(), arena.gather, bool(arena.multicast)))
    except Exception:  # noqa: BLE001 - reported to the parent instead of a silent time-out
        import traceback

        q.put((rank, False, False, -1, "error: " + traceback.format_exc()[-1500:], False))
        raise
    finally:
        dist.destroy_process_group()


def _worker_pipelined(rank: int, world: int, port: int, q, backend: str = "auto") -> None:
    """adam_step_pipelined (two slices, the second on a side stream) against mean-then-Adam; then two TrainEngines on
    the same batches, pipelined and not: same loss trajectory."""
    import torch.distributed as dist

    from thermo_nerf_b200 import functional as F
    from thermo_nerf_b200.dist import PeerArena

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    os.environ["TNF_PEER_BACKEND"] = backend
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        n, split = 64 * 6789, 64 * 1000
        arena = PeerArena(n, dev, split=split)
        P = torch.randn(n, generator=torch.Generator().manual_seed(123)).to(dev)
        arena.params.copy_(P)
        p_ref, m_ref, v_ref = P.clone(), torch.zeros(n, device=dev), torch.zeros(n, device=dev)
        ok, head_steps = True, 0
        exact = arena.gather != "multimem" or world == 2
        for step in (1, 2, 3, 4, 5):
            head = step % 2 == 1  # the proposal slice is exchanged every other step
            head_steps += int(head)
            g_all = [torch.randn(n, generator=torch.Generator().manual_seed(1000 * step + r)) for r in range(world)]
            arena.wait_zeroed()
            arena.grads.copy_(g_all[rank].to(dev))
            arena.adam_step_pipelined([(0, split, 1e-2, max(head_steps, 1), head), (split, n, 5e-3, step, True)],
                                      side_ctas=64)
            ev = arena.field_ready_event()
            assert ev is not None
            torch.cuda.current_stream().wait_event(ev)
            mean = torch.zeros(n, device=dev)
            for g in g_all:
                mean += g.to(dev)
            mean *= 1.0 / world
            if head:
                F.adam_step([p_ref[:split]], [mean[:split]], [m_ref[:split]], [v_ref[:split]], [1e-2], step=head_steps,
                            eps=1e-15)
            F.adam_step([p_ref[split:]], [mean[split:]], [m_ref[split:]], [v_ref[split:]], [5e-3], step=step, eps=1e-15)
            torch.cuda.synchronize()
            cmp = torch.equal if exact else (lambda a, b: torch.allclose(a, b, rtol=2e-6, atol=1e-7))
            ok = ok and bool(cmp(arena.params, p_ref)) and arena.grads.abs().max().item() == 0.0
            off = 0
            for lo, hi in arena.owned_ranges():
                ok = ok and bool(cmp(arena.exp_avg[off:off + hi - lo], m_ref[lo:hi]))
                off += hi - lo
            if not exact:
                p_ref.copy_(arena.params)
        chk = arena.params.double().sum().reshape(1)
        both = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        same = all(float(b) == float(both[0]) for b in both)

        # engines: pipelined against the single-block exchange, same batches
        from oracle import make_synthetic_rays
        from tests.helpers import make_pair
        from thermo_nerf_b200.engine import TrainEngine

        hist = {}
        for pipe in ("1", "0"):
            os.environ["TNF_PEER_PIPELINE"] = pipe
            _, model = make_pair(trained_like=False, precision="tc_fp16", log2_field=14, log2_prop=12,
                                 camera_optimizer_mode="off", device=f"cuda:{rank}")
            model.train()
            eng = TrainEngine(model, peer_fused=True, world_size=world)
            assert eng.pipelined == (pipe == "1")
            rays = make_synthetic_rays(1024, num_images=8, seed=5 + rank)
            gen = torch.Generator().manual_seed(1 + rank)
            gt_rgb = torch.rand((1024, 3), generator=gen).mul(0.2).add(0.4).to(dev)
            gt_th = torch.rand((1024,), generator=gen).mul(0.2).add(0.6).to(dev)
            jit = torch.rand((3, 1024), generator=gen).to(dev)
            o, d, c = rays.origins.to(dev), rays.directions.to(dev), rays.camera_indices.to(dev).reshape(-1)
            h = []
            for _ in range(24):
                ls = eng.step(o, d, c, gt_rgb, gt_th, jitter=jit)
                h.append(float(ls[0] + ls[3]))
            torch.cuda.synchronize()
            ok = ok and eng.arena.timeouts() == 0
            hist[pipe] = h
            # an eval render right after a step goes through model.tensors(), which waits for the pending slice
            model.eval()
            with torch.no_grad():
                from thermo_nerf_b200 import RayBundle

                out = model.get_outputs(RayBundle(origins=o[:64], directions=d[:64], camera_indices=c[:64, None]))
            ok = ok and bool(torch.isfinite(out["rgb"]).all())
        a, b = hist["1"], hist["0"]
        ok = ok and abs(a[0] - b[0]) < 1e-6 and a[-1] < 0.6 * a[0] and abs(a[-1] - b[-1]) < 0.05 * b[-1]
        q.put((rank, ok, same, arena.timeouts(), arena.gather, (a[0], a[-1], b[-1])))
    except Exception as e:  # noqa: BLE001 - reported to the parent instead of a silent time-out
        import traceback

        q.put((rank, False, False, -1, "error", traceback.format_exc()[-1500:]))
        raise
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.timeout(400)
@pytest.mark.parametrize("backend", ["ipc", "auto"])
def test_world2_pipelined_exchange(backend):
    """The exchange as two slices (proposal networks on the critical path, field on a side stream under the next
    proposal pass): same parameters as mean-then-Adam, and a TrainEngine that trains like the unpipelined one."""
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8) if os.environ.get("TNF_TEST_ALL_GPUS") else 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_pipelined, args=(r, world, port, q, backend)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
    for rank, ok, same, timeouts, gather, info in out:
        assert ok, f"rank {rank} ({gather}): {info}"
        assert same and timeouts == 0
    print(f"backend={backend}: pipelined exchange, flavour {out[0][4]}, losses (first, last, last unpipelined) {out[0][5]}")


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.timeout(300)
@pytest.mark.parametrize("backend", ["ipc", "auto"])
def test_world2_fused_exchange_matches_mean_then_adam(backend):
    """backend "ipc": our own CUDA-IPC mapping, peer loads / stores.  "auto": torch symmetric memory and, where the
    NVSwitch multicast object exists, the in-switch reduction + broadcast (multimem.ld_reduce / multimem.st)."""
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 8) if os.environ.get("TNF_TEST_ALL_GPUS") else 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, backend)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    for rank, ok, same, timeouts, gather, multicast in out:
        assert ok, f"rank {rank}: fused result differs from mean-then-Adam ({gather})"
        assert same and timeouts == 0
    print(f"backend={backend}: exchange flavour {out[0][4]}, multicast {out[0][5]}")
