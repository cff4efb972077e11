"""torchrun --nproc-per-node W tests/run_ddp_train.py — data-parallel training step on W GPUs (NCCL): checks that the four bucketed,
overlapped all-reduces cover the whole gradient arena (all-reduced gradient == mean of the ranks' local gradients), that parameters stay
bit-identical across ranks after optimiser steps, and that W ranks x batch b reproduce the single-process gradient of the concatenated
batch W*b up to the loss normalisation (per-rank mean, then average: train_se3_diffusion.py:662-666 + DDP).  Prints one JSON line on rank 0."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from se3_diffusion_b200 import FrameDiffEngine
    from se3_diffusion_b200.parallel import TrainStep
    from se3_diffusion_b200.synthetic import synthetic_weights, training_batch
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    eng = FrameDiffEngine(local, "fp32")
    eng.train_set_gemm(os.environ.get("FD_TRAIN_GEMM", "bf16x3"))
    ts = TrainStep(eng, synthetic_weights(0), lr=1e-3)
    ts.broadcast_parameters()
    B, N = 2, 64
    batch = {k: v.to(dev) for k, v in training_batch(eng, B, N, seed=10 + rank, pad_last=5 if rank == 0 else 0).items()}
    FE = ("rigids_t", "res_mask", "fixed_mask", "seq_idx", "t", "sc_ca_t", "torsion_angles_sin_cos")
    feats = {k: batch[k] for k in FE}
    # local gradient (no communication), for the reference mean
    out = eng.train_forward(feats)
    dout = eng.loss_backward(out, batch)
    ts.grads.zero_()
    eng.train_backward(dout)
    local_g = ts.grads.clone()
    gathered = [torch.empty_like(local_g) for _ in range(world)]
    dist.all_gather(gathered, local_g)
    mean_g = torch.stack(gathered).mean(0)
    # the DDP step: staged backward + overlapped bucket all-reduce + Adam; keep the reduced gradient before Adam consumes it
    p0 = ts.params.clone()
    loss = ts(feats, batch)
    red = ts.grads / world
    err = float((red - mean_g).abs().max() / mean_g.abs().max())
    moved = float((ts.params - p0).abs().max())
    ts(feats, batch)
    ps = [torch.empty_like(ts.params) for _ in range(world)]
    dist.all_gather(ps, ts.params)
    same = all(bool(torch.equal(ps[0], p)) for p in ps)
    res = {"world": world, "allreduced_vs_mean_of_local_rel_err": err, "params_identical_across_ranks": same, "param_step": moved,
           "loss_rank0": float(loss), "exposed_comm_ms": ts.exposed_comm_ms(),
           "ok": bool(err < 2e-5 and same and moved > 0)}
    if rank == 0:
        print(json.dumps(res), flush=True)
    dist.destroy_process_group()
    return 0 if res["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
