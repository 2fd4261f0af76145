"""Multi-GPU host logic of the path: one process per GPU, rays shard naturally.

* training: every rank draws its own rays (nerfstudio DDP semantics, seed + rank) and the flat
  gradient arena is all-reduced (mean) once per step - the only exchange step of the path;
* render / eval: frames are dealt round-robin over ranks, no collective in the compute.

The functions take CPU or CUDA tensors and any backend, so the ``gloo`` world_size-2 tests on CPU
exercise exactly the code the NCCL runs use.
"""

from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.distributed as dist
from torch import Tensor


def peer_padding(world_size: int) -> int:
    """Slices of a PeerArena are sharded in runs of 4 floats per rank: their lengths are multiples of this."""
    return 4 * int(world_size)


def world_info(group=None) -> tuple:
    """(rank, world_size), (0, 1) when torch.distributed is not initialised."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def shard_frames(num_frames: int, rank: int, world_size: int) -> List[int]:
    """Frame indices rendered by ``rank``: r, r + world, r + 2 world, ... (every frame exactly once)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return list(range(rank, num_frames, world_size))


def rank_seed(base_seed: int, rank: int, step: int = 0) -> int:
    """Seed of the pixel batch ``step`` on ``rank`` (independent draws per rank, as nerfstudio's per-rank
    datamanagers)."""
    return int(base_seed) + 1000 * int(rank) + int(step)


def allreduce_mean_(flat: Tensor, group=None, world_size: Optional[int] = None) -> Tensor:
    """In-place mean over ranks of one flat buffer (DDP gradient semantics).  NCCL averages inside the
    collective; backends without ReduceOp.AVG (gloo) sum and scale."""
    if world_size is None:
        world_size = world_info(group)[1]
    if world_size <= 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world_size)
    return flat


def allreduce_mean_grads_(params: Sequence[Tensor], group=None, world_size: Optional[int] = None) -> None:
    """Mean over ranks of ``p.grad`` for the plugin (autograd) route.  Gradients that are views of one
    contiguous arena (what ``functional.render``'s backward produces) go out as a single collective."""
    if world_size is None:
        world_size = world_info(group)[1]
    if world_size <= 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    base = grads[0]._base if grads[0]._base is not None else None
    if base is not None and all(g._base is base for g in grads) and base.is_contiguous():
        span = sum((g.numel() + 3) // 4 * 4 for g in grads)
        if span == base.numel():
            allreduce_mean_(base, group, world_size)
            return
    for g in grads:
        allreduce_mean_(g, group, world_size)


def gather_frames(frames: Tensor, num_frames: int, group=None) -> Optional[Tensor]:
    """Rank 0 receives the frames rendered by every rank in frame order ([num_frames, ...]); other ranks get
    None.  ``frames`` holds this rank's share in the order of :func:`shard_frames`."""
    rank, world = world_info(group)
    if world == 1:
        return frames
    per = (num_frames + world - 1) // world
    pad = torch.zeros((per, *frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    pad[: frames.shape[0]] = frames
    bufs = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
    dist.gather(pad, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = torch.empty((num_frames, *frames.shape[1:]), dtype=frames.dtype, device=frames.device)
    for r in range(world):
        idx = shard_frames(num_frames, r, world)
        out[idx] = bufs[r][: len(idx)]
    return out


class _DeviceMemory:
    """``__cuda_array_interface__`` view of memory owned by a PeerArena (keeps the arena alive)."""

    def __init__(self, ptr: int, shape: tuple, typestr: str, owner) -> None:
        self.owner = owner
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2,
                                         "strides": None}


def _wrap_device_memory(ptr: int, shape: tuple, typestr: str, device, owner) -> Tensor:
    return torch.as_tensor(_DeviceMemory(ptr, shape, typestr, owner), device=device)


class PeerArena:
    """Flat gradient + parameter arenas of every rank mapped into this process - torch symmetric memory (with the
    NVSwitch multicast object when there is one) or our own CUDA-IPC mapping - for the gradient mean over ranks fused
    with Adam in one kernel: reduce-scatter -> Adam on the owned shard -> all-gather, by peer loads / stores
    (``tnf_peer_adam_step``) or inside the switch (``tnf_peer_adam_multimem``), bracketed by flag barriers.

    :meth:`adam_step` runs it as one block on the current stream; :meth:`adam_step_pipelined` exchanges two slices of
    the arena separately, the large one on a side stream (``tnf_peer_adam_range``).

    ``torch.distributed`` is only the plumbing here (it maps the buffers once, at construction); the per-step exchange
    is our kernel reading and writing peer memory.  world_size 1 degenerates to a local fused Adam, which is what the
    single-GPU tests exercise."""

    def __init__(self, numel: int, device, group=None, split: Optional[int] = None) -> None:
        """``split``: boundary between two slices of the arena that are exchanged separately
        (:meth:`adam_step_pipelined`): [0, split) and [split, numel), each sharded over the ranks on its own; it must be
        a multiple of 4 * world_size (:func:`peer_padding`)."""
        import ctypes as C

        from . import _lib as L

        self._L, self._C = L, C
        self.lib = L.load()
        self.rank, self.world = world_info(group)
        if self.world > L.TNF_MAX_PEERS:
            raise ValueError(f"at most {L.TNF_MAX_PEERS} ranks")
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("PeerArena needs a CUDA device; there is no CPU path")
        pad = 4 * self.world
        self.numel = (int(numel) + pad - 1) // pad * pad
        self.shard = self.numel // self.world
        self.device = dev
        arena_bytes = self.numel * 4
        self._mapped: dict = {}  # rank -> base pointer of that rank's allocation mapped for this device (IPC path)
        self._symm = None         # (tensor, handle) of the symmetric-memory path
        self.multicast = 0        # multicast base address of the allocation (NVSwitch / NVLS), 0 = none
        import os

        backend = os.environ.get("TNF_PEER_BACKEND", "auto")
        if backend not in ("auto", "symm", "ipc"):
            raise ValueError("TNF_PEER_BACKEND must be 'auto', 'symm' or 'ipc'")
        bases = None
        if self.world > 1 and backend in ("auto", "symm"):
            bases = self._setup_symmetric(dev, group, arena_bytes, strict=(backend == "symm"))
        if bases is None:
            bases = self._setup_ipc(dev, group, arena_bytes)
        ptrs = [(b, b + arena_bytes, b + 2 * arena_bytes) for b in bases]
        a = L.TnfPeerArena()
        for r, (g, p, f) in enumerate(ptrs):
            a.grads[r], a.params[r], a.flags[r] = g, p, f
        a.world_size, a.rank, a.numel = self.world, self.rank, self.numel
        self.struct = a
        # Adam state of the part(s) of the arena this rank owns: one shard of the whole arena, or - with `split` -
        # this rank's part of [0, split) followed by its part of [split, numel)
        self.exp_avg = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(self.shard, dtype=torch.float32, device=dev)
        self.split = None if split is None else int(split)
        if self.split is not None and (self.split % pad or not 0 <= self.split <= self.numel):
            raise ValueError(f"split={split} must be a multiple of 4*world_size={pad} inside the arena")
        self._side: Optional[torch.cuda.Stream] = None
        self._field_ready: Optional[torch.cuda.Event] = None
        self._epoch = [0] * L.TNF_PEER_FLAG_SLOTS
        self.timing: Optional[list] = None
        # the gradient arena is cleared on a side stream so that the memset overlaps the next forward (which never
        # touches it); whoever writes gradients next calls wait_zeroed() first
        self._zero_stream: Optional[torch.cuda.Stream] = None
        self._zero_done: Optional[torch.cuda.Event] = None
        import os

        # exchange flavour: "push" (peer loads + peer stores in one kernel), "pull" (owners keep their shard, peers copy
        # it out in a second kernel; round 1's choice for 8 GPUs, profiles/r1_peer_exchange_phases.json) or "multimem"
        # (reduction and broadcast inside the NVSwitch; needs the multicast object of the symmetric path).
        self.gather = os.environ.get("TNF_PEER_GATHER", "auto")
        if self.gather == "auto":
            # measured on B200 (profiles/): two GPUs - the push kernel (half the arena per direction either way, and
            # plain peer stores run faster than switch reductions); from four GPUs up the in-switch reduction moves
            # 1/N instead of (N-1)/N of the arena per rank and direction
            # without a multicast object (no NVSwitch, or the CUDA-IPC mapping) the push kernel: unpipelined it is on a
            # par with the pull flavour at 8 GPUs (0.946 / 0.964 ms per step), and unlike pull it can be pipelined
            # (4 GPUs, pipelined push: 0.82 ms against 1.00 ms for the unpipelined pull of round 1)
            self.gather = "multimem" if (self.world >= 4 and self.multicast) else "push"
        if self.gather not in ("push", "pull", "multimem"):
            raise ValueError("TNF_PEER_GATHER must be 'auto', 'push', 'pull' or 'multimem'")
        if self.gather == "multimem" and not self.multicast:
            raise RuntimeError("TNF_PEER_GATHER=multimem needs NVSwitch multicast (symmetric-memory backend)")
        if self.world > 1:
            # self-test of the mappings in both directions: one barrier round must complete without a time-out
            self.barrier(0)
            torch.cuda.synchronize(dev)
            if self.timeouts() != 0:
                raise RuntimeError("peer flag barrier timed out: the IPC mappings are not reachable")

    def _setup_symmetric(self, dev, group, arena_bytes: int, strict: bool):
        """[gradients | parameters | flag block] as ONE torch symmetric-memory allocation: torch.distributed maps every
        rank's buffer into this process and, on NVSwitch systems, binds them to a multicast object - plumbing only;
        the exchange itself is our kernel (peer loads / stores, or multimem.ld_reduce / multimem.st on the multicast
        address).  Returns the per-rank base addresses, or None when symmetric memory is unavailable on any rank."""
        L = self._L
        err, buf, hdl = "", None, None
        try:
            import torch.distributed._symmetric_memory as symm_mem

            words = 2 * self.numel + L.TNF_PEER_FLAG_WORDS
            with torch.cuda.device(dev):
                buf = symm_mem.empty(words, dtype=torch.float32, device=dev)
                buf.zero_()
                torch.cuda.synchronize(dev)
                hdl = symm_mem.rendezvous(buf, group if group is not None else dist.group.WORLD)
            bases = [int(p) for p in hdl.buffer_ptrs]
            if len(bases) != self.world or bases[self.rank] != buf.data_ptr():
                raise RuntimeError("symmetric memory returned unexpected buffer pointers")
        except Exception as e:  # noqa: BLE001 - reported collectively below
            err = repr(e)
        ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            if strict:
                raise RuntimeError("PeerArena: TNF_PEER_BACKEND=symm but symmetric memory is unavailable"
                                   + (f" (this rank: {err})" if err else ""))
            return None
        self._symm = (buf, hdl)
        self._base = bases[self.rank]
        self.grads = buf[:self.numel]
        self.params = buf[self.numel:2 * self.numel]
        self.flags = buf[2 * self.numel:].view(torch.int32)
        mc = 0
        try:
            mc = int(hdl.multicast_ptr or 0)
        except Exception:  # noqa: BLE001
            mc = 0
        # every rank must agree (the multicast kernel is collective)
        flag = torch.tensor([1 if mc else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.multicast = mc if int(flag.item()) == 1 else 0
        return bases

    def _setup_ipc(self, dev, group, arena_bytes: int):
        """One cudaMalloc of our own per rank, shared through CUDA IPC handles (tensors of the caching allocator sit
        inside larger segments, and an IPC handle has to name an allocation's base)."""
        L, C = self._L, self._C
        total_bytes = 2 * arena_bytes + L.TNF_PEER_FLAG_WORDS * 4
        handle = C.create_string_buffer(L.TNF_IPC_HANDLE_BYTES)
        base = C.c_void_p()
        with torch.cuda.device(dev):
            L.check(self.lib.tnf_peer_alloc(total_bytes, C.byref(base), handle))
        self._base = int(base.value)
        self.grads = _wrap_device_memory(self._base, (self.numel,), "<f4", dev, self)
        self.params = _wrap_device_memory(self._base + arena_bytes, (self.numel,), "<f4", dev, self)
        self.flags = _wrap_device_memory(self._base + 2 * arena_bytes, (L.TNF_PEER_FLAG_WORDS,), "<i4", dev, self)
        bases = [self._base] * self.world
        if self.world > 1:
            gathered: List[object] = [None] * self.world
            dist.all_gather_object(gathered, (bytes(handle.raw), self.numel), group=group)
            err = ""
            try:
                with torch.cuda.device(dev):  # handles are opened with the consumer device current
                    for r, (h, n) in enumerate(gathered):
                        if n != self.numel:
                            raise RuntimeError("ranks disagree on the arena size")
                        if r == self.rank:
                            continue
                        out = C.c_void_p()
                        L.check(self.lib.tnf_peer_open_handle(h, C.byref(out)))
                        self._mapped[r] = int(out.value)
                        bases[r] = int(out.value)
            except Exception as e:  # noqa: BLE001 - reported collectively below
                err = repr(e)
            # every rank learns whether every mapping succeeded (a rank that raised alone would leave the others
            # waiting in the next collective)
            ok = torch.tensor([0 if err else 1], dtype=torch.int32, device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()) == 0:
                self.close()
                raise RuntimeError("PeerArena: mapping the peers' memory failed on at least one rank"
                                   + (f" (this rank: {err})" if err else ""))
        return bases

    def close(self) -> None:
        """Unmap the peers' allocations (the own allocation lives as long as tensors view it)."""
        for base in self._mapped.values():
            self.lib.tnf_peer_close_handle(self._C.c_void_p(base))
        self._mapped = {}

    def barrier(self, slot: int, stream: Optional["torch.cuda.Stream"] = None) -> None:
        """Stream-ordered barrier over all ranks (flag words in peer memory; no NCCL call)."""
        self._epoch[slot] += 1
        stream = (stream if stream is not None else torch.cuda.current_stream(self.device)).cuda_stream
        with torch.cuda.device(self.device):
            self._L.check(self.lib.tnf_peer_barrier(self._C.byref(self.struct), slot, self._epoch[slot],
                                                    self._C.c_void_p(stream)))

    def adam_step(self, segments: Sequence[tuple], beta1: float = 0.9, beta2: float = 0.999, eps: float = 1e-15,
                  zero_grads: bool = True, async_zero: bool = False) -> None:
        """``segments`` = [(begin, end, lr, step, active), ...] covering [0, numel).  Runs
        barrier -> fused reduce-scatter/Adam/all-gather -> barrier (-> zero this rank's gradient arena).
        ``async_zero``: clear the arena on a side stream instead (the caller must :meth:`wait_zeroed` before the
        next kernel that accumulates gradients) - TrainEngine does, so that the memset runs beside the forward."""
        L, C = self._L, self._C
        segs = (L.TnfAdamSegment * len(segments))()
        for i, (b, e, lr, step, active) in enumerate(segments):
            segs[i].begin, segs[i].end, segs[i].lr = int(b), int(e), float(lr)
            segs[i].step, segs[i].active = int(step), int(bool(active))
        ev = None
        if self.timing is not None:
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
            ev[0].record()
        self.barrier(0)  # every rank's backward has written its gradient arena
        if ev:
            ev[1].record()
        stream = torch.cuda.current_stream(self.device).cuda_stream
        with torch.cuda.device(self.device):
            if self.gather == "multimem":
                L.check(self.lib.tnf_peer_adam_multimem(
                    C.byref(self.struct), C.c_void_p(self.multicast), C.c_void_p(self.multicast + 4 * self.numel),
                    self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), segs, len(segments), float(beta1),
                    float(beta2), float(eps), C.c_void_p(stream)))
            else:
                fn = self.lib.tnf_peer_adam_step if self.gather == "push" else self.lib.tnf_peer_adam_reduce
                L.check(fn(C.byref(self.struct), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(), segs,
                           len(segments), float(beta1), float(beta2), float(eps), C.c_void_p(stream)))
        if ev:
            ev[2].record()
        self.barrier(1)  # push: every shard has landed everywhere; pull: every owner holds its updated shard
        if self.gather == "pull" and self.world > 1:
            with torch.cuda.device(self.device):
                L.check(self.lib.tnf_peer_gather_params(C.byref(self.struct), segs, len(segments), C.c_void_p(stream)))
            # no third barrier: a peer can only overwrite its shard (next step's Adam) after the next barrier(0),
            # which this rank reaches after its pull has completed (stream order)
        if ev:
            ev[3].record()
        if zero_grads and not async_zero:
            self.grads.zero_()
        elif zero_grads:
            if self._zero_stream is None:
                self._zero_stream = torch.cuda.Stream(self.device)
            main = torch.cuda.current_stream(self.device)
            landed = torch.cuda.Event()
            landed.record(main)  # after barrier(1): every peer has finished reading this rank's gradients
            self._zero_stream.wait_event(landed)
            with torch.cuda.stream(self._zero_stream):
                self.grads.zero_()
                self._zero_done = torch.cuda.Event()
                self._zero_done.record(self._zero_stream)
        if ev:
            ev[4].record()
            self.timing.append(ev)

    def owned_ranges(self) -> list:
        """[(begin, end), ...] of the arena whose Adam state this rank holds, in the order of ``exp_avg``."""
        if self.split is None:
            return [(self.shard * self.rank, self.shard * (self.rank + 1))]
        a, b = self.split // self.world, (self.numel - self.split) // self.world
        return [(a * self.rank, a * (self.rank + 1)), (self.split + b * self.rank, self.split + b * (self.rank + 1))]

    def _exchange_range(self, lo: int, hi: int, state_off: int, segs, nseg: int, beta1, beta2, eps, stream,
                        max_ctas: int) -> None:
        L, C = self._L, self._C
        flavour = L.TNF_PEER_MULTIMEM if self.gather == "multimem" else L.TNF_PEER_PUSH
        with torch.cuda.device(self.device):
            L.check(self.lib.tnf_peer_adam_range(
                C.byref(self.struct), flavour, lo, hi, max_ctas,
                C.c_void_p(self.multicast or None), C.c_void_p((self.multicast + 4 * self.numel) if self.multicast else None),
                self.exp_avg.data_ptr() + 4 * state_off, self.exp_avg_sq.data_ptr() + 4 * state_off, segs, nseg,
                float(beta1), float(beta2), float(eps), C.c_void_p(stream.cuda_stream)))

    def adam_step_pipelined(self, segments: Sequence[tuple], beta1: float = 0.9, beta2: float = 0.999,
                            eps: float = 1e-15, side_ctas: int = 0,
                            tail_grads_event: Optional["torch.cuda.Event"] = None) -> None:
        """The exchange as two slices with the second one off the critical path (``split`` must be set; push and
        multimem flavours).

        side stream   [wait for the gradients of [split, numel)] -> barrier -> [split, numel) exchanged -> barrier
                      -> event (:meth:`field_ready_event`) -> gradient arena cleared (:meth:`wait_zeroed`)
        this stream   barrier -> [0, split) exchanged -> barrier          (only when one of its segments is active)

        The side stream starts behind ``tail_grads_event`` when given (recorded by tnf_render_backward_staged between
        the field level and the proposal levels), else behind this stream's first barrier.

        TrainEngine's use: [0, split) are the proposal networks, [split, numel) the field.  The field slice - 7/8 of
        the bytes - crosses the NVSwitch while the proposal backward still runs (update steps) and while the next
        iteration's proposal pass, which reads the proposal networks only, already runs; the next field level waits
        for the event (tnf_render_forward_staged)."""
        if self.split is None or self.gather not in ("push", "multimem"):
            raise RuntimeError("adam_step_pipelined needs PeerArena(split=...) and the push or multimem flavour")
        L = self._L
        segs = (L.TnfAdamSegment * len(segments))()
        head_active = False
        for i, (b, e, lr, step, active) in enumerate(segments):
            segs[i].begin, segs[i].end, segs[i].lr = int(b), int(e), float(lr)
            segs[i].step, segs[i].active = int(step), int(bool(active))
            head_active = head_active or (bool(active) and int(b) < self.split)
        head_active = head_active and self.split > 0
        main = torch.cuda.current_stream(self.device)
        if self._side is None:
            self._side = torch.cuda.Stream(self.device)
        side = self._side
        if tail_grads_event is not None:
            side.wait_event(tail_grads_event)
            self.barrier(3, stream=side)  # every rank's gradients of the big slice are complete
            if head_active:
                self.barrier(0)           # ... and, on this stream, those of the small slice
        else:
            self.barrier(0)  # every rank's backward has written its gradient arena
            start = torch.cuda.Event()
            start.record(main)
            side.wait_event(start)
        a = self.split // self.world
        # side stream: the big slice
        self._exchange_range(self.split, self.numel, a, segs, len(segments), beta1, beta2, eps, side, side_ctas)
        self.barrier(2, stream=side)  # every rank's part of the slice has landed everywhere
        self._field_ready = torch.cuda.Event()
        self._field_ready.record(side)
        # current stream: the small slice, needed first
        if head_active:
            self._exchange_range(0, self.split, 0, segs, len(segments), beta1, beta2, eps, main, 0)
            self.barrier(1)
        # the gradients may be cleared once every peer has read them: the side stream's barrier covers the big
        # slice, the barrier above the small one (nobody reads it on a step without active head segments).  The
        # clearing also has to come after this rank's own backward (the proposal kernel may still be running)
        head_done = torch.cuda.Event()
        head_done.record(main)
        side.wait_event(head_done)
        with torch.cuda.stream(side):
            self.grads.zero_()
            self._zero_done = torch.cuda.Event()
            self._zero_done.record(side)

    def wait_params(self) -> None:
        """Orders the current stream after a still pending exchange of [split, numel) (for readers of the parameters
        other than the staged forward, e.g. an evaluation render between two training iterations)."""
        if self._field_ready is not None:
            torch.cuda.current_stream(self.device).wait_event(self._field_ready)

    def field_ready_event(self) -> Optional["torch.cuda.Event"]:
        """Event of the last :meth:`adam_step_pipelined` after which [split, numel) of the parameters is up to date on
        every rank (None when nothing is pending); handing it out clears it."""
        ev, self._field_ready = self._field_ready, None
        return ev

    def wait_zeroed(self) -> None:
        """Orders the current stream after the asynchronous clearing of the gradient arena started by the last
        :meth:`adam_step`.  Call before launching anything that accumulates into ``grads``."""
        if self._zero_done is not None:
            torch.cuda.current_stream(self.device).wait_event(self._zero_done)
            self._zero_done = None

    def timing_summary(self) -> Optional[dict]:
        """Median milliseconds of the four phases of adam_step (measurement aid: PeerArena.timing = [] enables it)."""
        if not self.timing:
            return None
        torch.cuda.synchronize(self.device)
        names = ("barrier_wait_for_backward", "fused_reduce_adam_gather", "barrier_params_landed", "zero_grads")
        cols = list(zip(*[[e[i].elapsed_time(e[i + 1]) for i in range(4)] for e in self.timing]))
        return {n: sorted(c)[len(c) // 2] for n, c in zip(names, cols)}

    def timeouts(self) -> int:
        """Number of barrier waits that gave up (a peer never arrived); 0 in a healthy run."""
        return int(self.flags[self._L.TNF_PEER_FLAG_TIMEOUT].item())
