"""Data parallelism for the propagation path: graphs are independent (a batch is a disjoint union of graphs,
chem_tensorflow_sparse.py:279-280), so ranks own contiguous ranges of GRAPHS and forward propagation needs no
communication.  Training adds exactly ONE all-reduce per step over a single flat buffer holding every trainable
gradient (plus the normalisation weight), NCCL over NVLink on GPUs, gloo in the CPU tests.  The reference has no
distributed code at all; per-variable clip_by_norm (chem_tensorflow.py:186-190) is applied AFTER the
all-reduce so multi-GPU training matches single-GPU semantics on the union batch."""
from __future__ import annotations

from typing import List, Sequence


def world():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(node_counts: Sequence[int], world_size: int) -> List[int]:
    """Contiguous graph ranges [b[r], b[r+1]) per rank, balanced by NODE count (greedy on the cumulative sum)."""
    n = len(node_counts)
    total = float(sum(node_counts))
    bounds, acc, g = [0], 0.0, 0
    for r in range(1, world_size):
        target = total * r / world_size
        while g < n and acc + node_counts[g] / 2.0 <= target:
            acc += node_counts[g]
            g += 1
        g = max(g, bounds[-1])
        bounds.append(g)
    bounds.append(n)
    return bounds


def shard_graphs(graphs: Sequence, rank: int, world_size: int, num_nodes=lambda g: len(g["node_features"])):
    b = shard_bounds([num_nodes(g) for g in graphs], world_size)
    return list(graphs[b[rank]:b[rank + 1]])


def allreduce_gradients(params, weight: float = 1.0):
    """sum_r weight_r * grad_r / sum_r weight_r  for every parameter, with ONE all-reduce.
    ``weight`` is the rank's normalisation mass.  A single-task loss is a masked mean (chem_tensorflow.py:163-166), so the mass
    that reproduces the union batch is the rank's MASK SUM (labelled graphs), not its graph count; with several tasks use
    ``FlatGradients``, which keeps one mass per task.  The flat layout is identical on every rank: a parameter without a gradient
    contributes zeros plus a 0 in its has-gradient slot, and keeps ``grad = None`` only if no rank had a gradient for it."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return
    params = list(params)
    if not params:
        return
    ref = next((p.grad for p in params if p.grad is not None), params[0])
    pieces = [(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1).to(ref.dtype) for p in params]
    flags = torch.tensor([0.0 if p.grad is None else 1.0 for p in params] + [1.0], device=ref.device, dtype=ref.dtype)
    flat = torch.cat(pieces + [flags])
    n_tail = len(params) + 1
    flat[:-n_tail].mul_(float(weight))
    flat[-1] = float(weight)
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)          # the single collective of the step
    total = float(flat[-1])
    has = flat[-n_tail:-1].tolist()
    off = 0
    for p, h in zip(params, has):
        n = p.numel()
        if h > 0:
            g = (flat[off:off + n] / total).view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
        off += n


class FlatGradients:
    """The gradient exchange of one data-parallel training step as ONE persistent flat fp32 buffer and ONE all-reduce, exact for
    the reference's loss (chem_tensorflow.py:161-170): per task  loss_t = ratio_t * sum_g(0.5 * diff^2 * mask) / (sum_g mask + 1e-7),
    summed over tasks.  On the union of all ranks' batches that is  sum_r num_{r,t} / (sum_r den_{r,t} + 1e-7), which is NOT a
    weighted mean of the ranks' own (locally normalised) gradients as soon as tasks have different masks.  So every rank
    back-propagates the UN-normalised numerators, one task at a time, straight into its segment of the buffer

        [ task 0: d num_0 / d theta (all trainables) | task 1: ... | den_0 .. den_{T-1} | active ]

    (``bind(t)`` points every ``p.grad`` at task t's views, so autograd accumulates in place and nothing is concatenated or copied
    back), all-reduces the buffer once, and ``finish`` forms  sum_t (sum_r d num_{r,t}) / (sum_r den_{r,t} + eps)  in task 0's
    segment, where the ``p.grad`` stay bound.  ``active`` counts the ranks that still had a batch: a rank that ran out of data keeps
    calling with zeros until every rank is done, so ranks with different numbers of batches never dead-lock."""

    def __init__(self, params, num_tasks: int):
        import torch
        self.params = list(params)
        self.num_tasks = int(num_tasks)
        self.sizes = [p.numel() for p in self.params]
        self.P = sum(self.sizes)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(self.num_tasks * self.P + self.num_tasks + 1, dtype=torch.float32, device=dev)
        self.views = []
        for t in range(self.num_tasks):
            off, vs = t * self.P, []
            for p, n in zip(self.params, self.sizes):
                vs.append(self.flat[off:off + n].view_as(p))
                off += n
            self.views.append(vs)

    @property
    def payload_bytes(self) -> int:
        return self.flat.numel() * 4

    def zero(self):
        self.flat.zero_()

    def bind(self, task: int):
        for p, v in zip(self.params, self.views[task]):
            p.grad = v

    def set_masses(self, dens, active: bool):
        import torch
        tail = torch.tensor([float(d) for d in dens] + [1.0 if active else 0.0], dtype=torch.float32)
        self.flat[self.num_tasks * self.P:].copy_(tail)

    def allreduce(self):
        import torch.distributed as dist
        rank, ws = world()
        if ws > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)      # the single collective of the step

    def finish(self, eps: float, sync: bool = True):
        """Combine the reduced numerators into the union-batch gradient (in task 0's segment); returns the number of active ranks.
        ``sync=False`` does the division with device-side scalars and returns None: no host round trip in the step."""
        seg0 = self.flat[:self.P]
        if not sync:
            base = self.num_tasks * self.P
            seg0.div_(self.flat[base] + eps)
            for t in range(1, self.num_tasks):
                seg0.addcdiv_(self.flat[t * self.P:(t + 1) * self.P], (self.flat[base + t] + eps).expand(self.P))
            self.bind(0)
            return None
        tail = self.flat[self.num_tasks * self.P:].tolist()
        seg0.div_(tail[0] + eps)
        for t in range(1, self.num_tasks):
            seg0.add_(self.flat[t * self.P:(t + 1) * self.P], alpha=1.0 / (tail[t] + eps))
        self.bind(0)
        return int(round(tail[-1]))
