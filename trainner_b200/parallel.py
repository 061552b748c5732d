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
    """Averaging all-reduce of a network's gradients (+ its optimizer step) OFF the critical path.

    The exchange of G needs nothing from the D step and vice versa: after backward_G the D step only
    needs fake.detach() and D's weights, and after backward_D the next iteration's G forward only needs G's
    weights.  So for world > 1 the all-reduce (ReduceOp.AVG -- no separate division pass) and the optimizer step run on
    a side stream, ordered after the backward that produced the gradients by an event, and the main stream
    waits for them only where the updated weights are first needed (SRModel.optimize_parameters); the optimizer steps move to the side stream on a single GPU too:
      G: exchange + Adam overlap the whole D step;   D: exchange + Adam overlap the next G forward.
    B200_OVERLAP=0 keeps everything on the main stream."""

    def __init__(self, group=None):
        self.group = group
        self._side = None
        self._done = {}
        self.overlap = os.environ.get("B200_OVERLAP", "1") != "0"

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
        """Average gradients over ranks (matching a mean loss over the global batch), on the current stream."""
        if self.world <= 1:
            return
        backend = dist.get_backend(self.group)
        for buf in flat_buffers_of(net):
            if backend == "nccl":
                dist.all_reduce(buf, op=dist.ReduceOp.AVG, group=self.group)
            else:   # gloo (CPU tests) has no AVG
                dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=self.group)
                buf.div_(self.world)

    def step_async(self, key, net, step_fn):
        """all-reduce `net`'s gradients and run step_fn() (clip + optimizer step + zero_grad).  world == 1, CPU or
        B200_OVERLAP=0: inline on the current stream.  Otherwise on the side stream; wait(key) joins it."""
        # (the side stream also pays off on ONE GPU: the fused Adam passes are HBM-bound and overlap the compute-bound
        #  D step / next G forward; measured 29.4 -> 28.7 ms/step)
        use_side = self.overlap and torch.cuda.is_available() and next(net.parameters()).is_cuda
        if not use_side:
            self.all_reduce_grads(net)
            step_fn()
            return
        if self._side is None:
            self._side = torch.cuda.Stream()
        main = torch.cuda.current_stream()
        ready = torch.cuda.Event()
        ready.record(main)
        # Gradients that are ordinary autograd tensors (the discriminator's classifier: everything outside the flat
        # buffers) were allocated on the main stream and are dropped by zero_grad(set_to_none=True) inside step_fn --
        # on the host, long before the side stream has run the optimizer kernels that read them.  Without
        # record_stream the caching allocator hands their memory to the main stream's next allocations and Adam
        # reads whatever landed there (found by tools/stress_repro.py: ~1 in 8 runs differed in D's parameters).
        for p in net.parameters():
            if p.grad is not None:
                p.grad.record_stream(self._side)
        with torch.cuda.stream(self._side):
            self._side.wait_event(ready)
            self.all_reduce_grads(net)
            step_fn()
            done = torch.cuda.Event()
            done.record(self._side)
        self._done[key] = done

    def wait(self, key):
        """make the current stream wait for the pending exchange + optimizer step of `key` (no-op if none)"""
        done = self._done.pop(key, None)
        if done is not None:
            torch.cuda.current_stream().wait_event(done)
