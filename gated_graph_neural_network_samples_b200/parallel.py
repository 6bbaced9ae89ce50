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
    ``weight`` is the rank's normalisation mass (graphs in its shard): each rank's loss is a mean over its own
    graphs, so the union-batch mean is the weight-averaged combination."""
    import torch
    import torch.distributed as dist
    rank, ws = world()
    if ws == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads] + [torch.ones(1, device=grads[0].device, dtype=grads[0].dtype)])
    flat.mul_(float(weight))
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)          # the single collective of the step
    flat.div_(flat[-1].clone())
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
