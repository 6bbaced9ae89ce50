"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU): graph sharding + the single gradient all-reduce."""
import os
import socket

import numpy as np
import pytest

from gated_graph_neural_network_samples_b200 import parallel, synthetic


def test_shard_bounds_are_contiguous_and_node_balanced():
    counts = [len(m["node_features"]) for m in synthetic.make_molecules(101, seed=4)]
    for ws in (1, 2, 4, 8):
        b = parallel.shard_bounds(counts, ws)
        assert b[0] == 0 and b[-1] == len(counts) and len(b) == ws + 1 and all(x <= y for x, y in zip(b, b[1:]))
        loads = [sum(counts[b[r]:b[r + 1]]) for r in range(ws)]
        assert max(loads) - min(loads) <= max(counts) + 1
    assert parallel.shard_bounds([5], 4) == [0, 0, 0, 1, 1] or parallel.shard_bounds([5], 4)[-1] == 1


def _worker(rank, world, port, out_q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    params = [torch.zeros(3, 4, requires_grad=True), torch.zeros(7, requires_grad=True), torch.zeros(2, 2, requires_grad=True)]
    g = torch.Generator().manual_seed(100 + rank)
    for p in params[:2]:
        p.grad = torch.randn(p.shape, generator=g)
    weight = 3.0 if rank == 0 else 5.0                      # graphs in the rank's shard
    parallel.allreduce_gradients(params, weight)
    out_q.put((rank, [None if p.grad is None else p.grad.numpy().copy() for p in params]))
    dist.barrier()
    dist.destroy_process_group()


def test_single_allreduce_is_weighted_mean_of_rank_gradients():
    import torch
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = []
    for shape in [(3, 4), (7,)]:
        gs = []
        for r in range(2):
            g = torch.Generator().manual_seed(100 + r)
            gs.append(None)
        expect.append(shape)
    g0, g1 = torch.Generator().manual_seed(100), torch.Generator().manual_seed(101)
    for i, shape in enumerate([(3, 4), (7,)]):
        a, b = torch.randn(shape, generator=g0), torch.randn(shape, generator=g1)
        want = (3.0 * a + 5.0 * b) / 8.0
        for r in range(2):
            np.testing.assert_allclose(res[r][i], want.numpy(), rtol=1e-6, atol=1e-7)
    assert res[0][2] is None and res[1][2] is None          # parameters without a gradient are left alone
