"""Batch sharding of the sampling path across the GPUs of one node.

Backbones are independent — there is no cross-sample operation anywhere on the path — so rank r simply owns the global
samples [first, first + count) and runs the whole 501-forward loop without communication; the only collective is the final
gather of the coordinates.  Noise is counter-based (Philox keyed by (seed, global sample index, step, residue)), so a
sample's trajectory does not depend on the world size.  One process per GPU, torch.distributed (NCCL on GPUs; gloo in the
CPU tests of this host logic).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(global_batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced partition of [0, global_batch): returns (first, count) of this rank."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(global_batch, world_size)
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


def gather_samples(local: torch.Tensor, global_batch: int) -> torch.Tensor:
    """All-gather per-rank sample tensors [count_r, ...] into [global_batch, ...] in global-sample order (every rank gets it).
    Ranks may own different counts (ragged), so shards are padded to the largest count for the collective."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    counts = [shard_range(global_batch, world, r)[1] for r in range(world)]
    mx = max(counts)
    pad = local.new_zeros((mx,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:c] for o, c in zip(out, counts)], dim=0)


def sample_sharded(engine, global_batch: int, nres: int, **kw):
    """Run engine.sample_device on this rank's shard and gather the final atom37 / rigids of the whole batch."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    first, count = shard_range(global_batch, world, rank)
    if count == 0:      # more ranks than samples: nothing to run here, but this rank still takes part in the gather
        dev = engine.device
        a37, rig, ms, nl = torch.empty(0, nres, 37, 3, device=dev), torch.empty(0, nres, 7, device=dev), 0.0, 0
    else:
        a37, rig, ms, nl = engine.sample_device(count, nres, first_sample=first, **kw)
    return gather_samples(a37, global_batch), gather_samples(rig, global_batch), ms, nl


# ---- training: data-parallel step with the gradient all-reduce overlapped with the backward ------------------
def grad_buckets():
    """Contiguous [lo, hi) float ranges of the flat gradient arena that each backward stage completes:
    stage 0 = trunk block 3 + torsion head, 1 = block 2, 2 = block 1, 3 = embedders + block 0 (engine.arena_layout order)."""
    from .engine import arena_layout
    lay, total = arena_layout()

    def first(prefix):
        return min(off for n, _, off in lay if n.startswith(prefix))

    b = [first(f"score_model.trunk.ipa_{k}.") for k in range(4)]
    return [(b[3], total), (b[2], b[3]), (b[1], b[2]), (0, b[1])]


class TrainStep:
    """One optimiser step on this rank's micro-batch, DDP semantics (train_se3_diffusion.py:268-286,320-326): forward -> loss ->
    backward in four stages, each finished gradient bucket all-reduced (SUM) asynchronously on NCCL's stream while the next stage
    computes -> gradients scaled by 1/world -> Adam.  Everything on the device runs in libframediff_b200.so; torch.distributed is the
    NCCL plumbing.  With world size 1 (or no process group) it is the plain single-GPU training step."""

    def __init__(self, engine, state, lr=1e-4, betas=(0.9, 0.999), eps=1e-8, exp_conf=None):
        from .engine import flat_from_state
        self.eng = engine
        self.params = flat_from_state(state, engine.device)
        self.grads = torch.zeros_like(self.params)
        self.m = torch.zeros_like(self.params)
        self.v = torch.zeros_like(self.params)
        self.step_no = 0
        self.lr, self.betas, self.eps, self.exp_conf = lr, betas, eps, exp_conf
        self.buckets = grad_buckets()
        engine.train_bind(self.params, self.grads)
        self.world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        self.comm_exposed_ms = 0.0

    def broadcast_parameters(self, src=0):
        if self.world > 1:
            dist.broadcast(self.params, src)

    def __call__(self, feats: dict, batch: dict, want_loss=True):
        eng = self.eng
        out = eng.train_forward(feats)
        loss = eng.loss_forward(out, batch, self.exp_conf)["total_loss"] if want_loss else None
        dout = eng.loss_backward(out, batch, self.exp_conf)
        self.grads.zero_()
        pending = []
        for s, (lo, hi) in enumerate(self.buckets):
            eng.train_backward(dout, s, s)
            if self.world > 1:
                pending.append(dist.all_reduce(self.grads[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
        if pending:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            for w in pending:
                w.wait()          # the compute stream waits for the collectives (no host block)
            ev1.record()
            self._last_wait = (ev0, ev1)
        self.step_no += 1
        eng.adam_step(self.params, self.grads, self.m, self.v, self.step_no, lr=self.lr, betas=self.betas, eps=self.eps, grad_scale=1.0 / self.world)
        return loss

    def exposed_comm_ms(self):
        """Device time the compute stream spent waiting for the last step's all-reduces after the backward had finished."""
        if self.world == 1 or not hasattr(self, "_last_wait"):
            return 0.0
        self._last_wait[1].synchronize()
        return self._last_wait[0].elapsed_time(self._last_wait[1])
