#!/usr/bin/env python
"""BASELINE configs[2]: the full blurred WGAN-GP train step at CelebA-128 with the batch sharded over N GPUs.

    python tools/train_step.py [--global-batch 256] [--steps 20] [--warmup 5] [--std 5.0]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/train_step.py ...

One process per GPU; every rank runs `BlurredWGANGP.train_on_batch` (blurred_gan_b200.wgan / .blurred_gan: the torch
restatement of reference wgan.py + blurred_gan.py with the reference's DCGAN nets) on its shard of the global batch and
all-reduces the gradients (NCCL) before each optimiser step.  The networks run on cuDNN (library code); the seven blur
applications of the step run on the B200 kernels.  Prints one JSON line: images/s of the whole job, ms per step, and
the blur's share of the step (device time of the same step with the blur layer replaced by the identity).
Synthetic reals U[-1, 1); random-init weights; sigma follows BlurDecayController from --std.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from blurred_gan_b200 import callbacks, nets  # noqa: E402
from blurred_gan_b200.blurred_gan import BlurredWGANGP  # noqa: E402
from blurred_gan_b200.dist import allreduce_gradients  # noqa: E402
from blurred_gan_b200.gaussian_blur import GaussianBlur2D  # noqa: E402


class _Identity(nn.Module):
    def __init__(self):
        super().__init__()
        self.std = nn.Parameter(torch.zeros(()), requires_grad=False)

    def assign(self, v):
        pass

    def forward(self, x):
        return x


def timed_steps(gan, reals, ctl, steps, warmup, world):
    for _ in range(warmup):
        ctl.on_batch_begin()
        gan.train_on_batch(reals)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        ctl.on_batch_begin()
        gan.train_on_batch(reals)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def run(global_batch=256, steps=20, warmup=5, std=5.0, fused_critic=False):
    """The measurement as a function (bench.py calls it for its `train_step_cfg3` context row).  Uses the current CUDA device and
    the default process group when one is initialised; returns the result dict on every rank."""
    world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    per_rank = global_batch // world
    out = {}
    metrics = None
    for kind in ("blur", "identity"):
        torch.manual_seed(0)  # identical initial weights on every rank
        hp = BlurredWGANGP.HyperParameters(batch_size=per_rank, global_batch_size=global_batch, initial_blur_std=std)
        blur = GaussianBlur2D(std) if kind == "blur" else _Identity()
        gan = BlurredWGANGP(nets.CelebaGenerator(), nets.CelebaDiscriminator(), hp, blur=blur, fused_critic=fused_critic and kind == "blur",
                            grad_sync=allreduce_gradients if world > 1 else None).cuda()
        ctl = callbacks.BlurDecayController(total_n_training_examples=202599 * 10, max_value=std)
        ctl.set_model(gan)
        torch.manual_seed(1000 + rank)
        reals = torch.rand(per_rank, 3, 128, 128, device="cuda") * 2 - 1
        out[kind] = timed_steps(gan, reals, ctl, steps, warmup, world)
        if kind == "blur":
            metrics = gan.metric_results()
        del gan, reals
    ms, ms_id = out["blur"], out["identity"]
    return {
        "metric": "blurred WGAN-GP train step images/s at 128x128x3", "value": global_batch / (ms * 1e-3), "unit": "images/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": ms, "scaling": "strong", "dtype": "f32", "data": "synthetic",
        "config": {"workload": "cfg3: BlurredWGANGP.train_on_batch, CelebA-128 DCGAN nets, global batch %d sharded over %d GPU(s), std %.2f" % (
            global_batch, world, std), "blur_applications_per_step": 6 if fused_critic else 7, "fused_critic": bool(fused_critic)},
        "ms_per_step_without_blur": ms_id, "blur_share_of_step": max(ms - ms_id, 0.0) / ms,
        "last_metrics": metrics}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--global-batch", type=int, default=256)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=5)
    p.add_argument("--std", type=float, default=5.0)
    p.add_argument("--fused-critic", action="store_true",
                   help="SURVEY 8f row N3: one blur launch for the generated and the real batch, interpolates between the blurred batches")
    a = p.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    res = run(a.global_batch, a.steps, a.warmup, a.std, a.fused_critic)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
