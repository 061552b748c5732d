"""Data-parallel gradient exchange: one process per GPU, NCCL all-reduce over NVLink 5 / NVSwitch.

Replaces the reference's single-process nn.DataParallel (codes/models/networks.py:252-254,
365-367: per-forward parameter broadcast + output gather + gradient reduce to GPU 0).  Here the
weights are replicated once (broadcast at construction), every rank runs the full G/D step on its
own batch shard, and the ONLY data-path exchange is an averaging all-reduce of the flat fp32
gradient buffer of G right before optimizer_G.step() and of D right before optimizer_D.step()
(SURVEY.md 8e).  BatchNorm statistics and the relativistic means stay per-rank, which is what
nn.DataParallel does for BN and what `gpu_ids: [0]` does per shard.
"""
import os

import torch
import torch.distributed as dist


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1 or dist.is_initialized():
        return world
    backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
    if backend == "nccl":
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    dist.init_process_group(backend=backend)
    return world


def flat_buffers_of(net):
    """The gradient storage to exchange for `net`: the engines' flat buffers when present (one
    tensor per network) plus any parameter gradient that lives outside them."""
    bufs, covered = [], set()
    for m in net.modules():
        eng = getattr(m, "_engine", None)
        if eng and getattr(eng[0], "grads", None) is not None:
            fg = eng[0].grads
            bufs.append(fg.flat)
            covered.update(id(p) for p in fg.params)
    for p in net.parameters():
        if id(p) not in covered and p.grad is not None:
            bufs.append(p.grad)
    return bufs


class GradExchange:
    def __init__(self, group=None):
        self.group = group

    @property
    def world(self):
        return dist.get_world_size(self.group) if dist.is_available() and dist.is_initialized() else 1

    def broadcast_params(self, nets, src=0):
        if self.world <= 1:
            return
        for net in nets:
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t.data, src=src, group=self.group)

    def all_reduce_grads(self, net):
        """Average gradients over ranks (sum / world, matching a mean loss over the global batch)."""
        w = self.world
        if w <= 1:
            return
        for buf in flat_buffers_of(net):
            dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
            buf.div_(w)
