"""View-batch data parallelism for the rasterizer (additive to the reference API; the reference has no
multi-GPU path at all -- SURVEY.md section 2 "Parallelism strategies: none").

The path shards by *views*: every rank holds a full replica of the Gaussian parameters, renders its slice
of the step's camera batch (forward + backward), and the per-rank gradients -- accumulated over the local
views in ONE flat fp32 buffer -- are summed with a single all-reduce per step
(NCCL over NVLink 5 / NVSwitch on the GPU box; gloo in the CPU tests).  No collective sits on the
per-view data path.

    flat = FlatGradBuffer([means3D, scales, rotations, opacities, shs, features])   # .grad are views
    for v in shard_views(n_views, rank, world):
        render(v) -> loss.backward()          # autograd accumulates in place into the flat buffer
    flat.all_reduce()                         # exactly one collective per step
"""
from typing import Iterable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_views(n_views: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment of view indices to ranks (rank r renders r, r+G, r+2G, ...)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    return list(range(rank, n_views, world_size))


class FlatGradBuffer:
    """One contiguous fp32 gradient buffer whose slices are the `.grad` of every parameter."""

    def __init__(self, params: Sequence[torch.Tensor], extra: int = 0):
        params = list(params)
        if not params:
            raise ValueError("no parameters")
        dev, dt = params[0].device, params[0].dtype
        for p in params:
            if p.device != dev or p.dtype != dt or not p.is_leaf or not p.requires_grad:
                raise ValueError("parameters must be leaf tensors requiring grad on one device with one dtype")
        self.params = params
        self.sizes = [p.numel() for p in params]
        self.offsets = [0]
        for n in self.sizes:
            self.offsets.append(self.offsets[-1] + n)
        # `extra` trailing floats for step statistics that ride in the same collective
        self.flat = torch.zeros(self.offsets[-1] + extra, device=dev, dtype=dt)
        for p, o, n in zip(params, self.offsets, self.sizes):
            p.grad = self.flat[o:o + n].view_as(p)

    @property
    def extra(self) -> torch.Tensor:
        return self.flat[self.offsets[-1]:]

    def zero_(self):
        self.flat.zero_()

    def check_views(self):
        """The .grad tensors must still alias the flat buffer (autograd accumulates in place)."""
        base = self.flat.data_ptr()
        for p, o in zip(self.params, self.offsets):
            if p.grad is None or p.grad.data_ptr() != base + o * self.flat.element_size():
                raise RuntimeError("a parameter's .grad no longer aliases the flat gradient buffer")

    def all_reduce(self, group: Optional[dist.ProcessGroup] = None, average: bool = False):
        """The single collective of a step.  No-op without an initialised process group."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group)
            if average:
                self.flat.div_(dist.get_world_size(group))
        return self.flat


def render_views(render_fn, view_ids: Iterable[int], flat: Optional[FlatGradBuffer] = None):
    """Run `render_fn(view_id)` (which must call backward itself and return a scalar tensor) over the local
    views; returns the list of per-view losses.  Convenience wrapper used by bench.py and the tests."""
    losses = []
    for v in view_ids:
        losses.append(render_fn(v))
    if flat is not None:
        flat.check_views()
    return losses
