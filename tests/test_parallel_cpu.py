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
    assert res[0][2] is None and res[1][2] is None          # a parameter NO rank has a gradient for is left alone


def _dp_worker(rank, world, port, tmp, out_q):
    """tools/dp_check.py on the CPU: the ChemModel mirror (stand-in engine answering from the oracle, see test_chem_model_cpu.py)
    on this rank's shard of graphs, ONE all-reduce, compared with the union batch computed locally."""
    import torch
    import torch.distributed as dist
    from gated_graph_neural_network_samples_b200 import chem_sparse, synthetic
    from tests.test_chem_model_cpu import StandInEngine, StandInPropagation
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chem_sparse.PropagationEngine = StandInEngine
    chem_sparse._propagation_function = lambda: StandInPropagation
    mols = synthetic.make_molecules(40, seed=11)
    args = {"--log_dir": os.path.join(tmp, "r%d" % rank), "--device": "cpu", "--train_data": mols, "--valid_data": mols[:4],
            "--config": {"hidden_size": 12, "batch_size": 100000, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                         "edge_weight_dropout_keep_prob": 1.0, "random_seed": 3}}
    model = chem_sparse.SparseGGNNChemModel(args)            # same seed on every rank -> identical replicas

    def grads_of(graphs):
        proc = model.process_raw_graphs(graphs, is_training_data=False)
        batch = next(iter(model.make_minibatch_iterator(proc, is_training=False)))
        batch["out_layer_dropout_keep_prob"] = 1.0
        for _, v in model._train_vars:
            v.grad = None
        loss, _ = model.forward_batch(batch)
        loss.backward()
        return batch["num_graphs"]

    n_r = grads_of(parallel.shard_graphs(mols, rank, world))
    parallel.allreduce_gradients([v for _, v in model._train_vars], weight=float(n_r))
    got = [None if v.grad is None else v.grad.clone() for _, v in model._train_vars]
    grads_of(mols)                                            # the union batch on one rank
    worst = 0.0
    for (_, v), g in zip(model._train_vars, got):
        if v.grad is not None:
            worst = max(worst, float((g - v.grad).abs().max()) / (float(v.grad.abs().max()) + 1e-12))
    out_q.put((rank, n_r, worst))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_training_step_equals_the_union_batch(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_dp_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert sum(n for _, n, _ in res) == 40 and all(n > 0 for _, n, _ in res)      # every graph on exactly one rank
    for rank, _, worst in res:
        assert worst < 1e-4, (rank, worst)


def _two_task_mols(n, seed):
    """Molecules with two regression targets (the reference's QM9 files carry 13, chem_tensorflow.py:215-217)."""
    mols = synthetic.make_molecules(n, seed=seed)
    rng = np.random.default_rng(seed + 1)
    for m in mols:
        m["targets"] = [m["targets"][0], [float(rng.normal())]]
    return mols


def _flat_worker(rank, world, port, tmp, out_q):
    """Unequal shards (rank 0: 9 graphs, rank 1: 31), two tasks, graphs WITHOUT a label for task 1 spread unevenly over the ranks
    (task_sample_ratios leaves labels None, chem_tensorflow_sparse.py:246-249): the one-all-reduce step must give the gradient of
    the union batch.  The previous exchange (rank gradient weighted by its graph count) must NOT -- the test pins why it was replaced.
    Then a ragged epoch: rank 0 runs out of batches one step before rank 1 and keeps taking part with an empty contribution."""
    import torch
    import torch.distributed as dist
    from gated_graph_neural_network_samples_b200 import chem_sparse
    from tests.test_chem_model_cpu import StandInEngine, StandInPropagation
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    chem_sparse.PropagationEngine = StandInEngine
    chem_sparse._propagation_function = lambda: StandInPropagation
    mols = _two_task_mols(40, seed=11)
    args = {"--log_dir": os.path.join(tmp, "r%d" % rank), "--device": "cpu", "--train_data": mols[:8], "--valid_data": mols[:4],
            "--config": {"hidden_size": 12, "batch_size": 100000, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                         "edge_weight_dropout_keep_prob": 1.0, "random_seed": 3, "task_ids": [0, 1], "learning_rate": 0.01}}
    model = chem_sparse.SparseGGNNChemModel(args)            # same seed on every rank -> identical replicas
    unlabeled = {1, 2, 3, 5, 8, 20}                          # graphs without a task-1 label: five of them on rank 0's shard of nine

    def feed_of(idx):
        proc = model.process_raw_graphs([mols[i] for i in idx], is_training_data=False)
        batch = next(iter(model.make_minibatch_iterator(proc, is_training=False)))
        batch["out_layer_dropout_keep_prob"] = 1.0
        for k, i in enumerate(idx):
            if i in unlabeled:
                batch["target_mask"][1, k] = 0.0
                batch["target_values"][1, k] = 0.0
        return batch

    shard = list(range(0, 9)) if rank == 0 else list(range(9, 40))
    variables = [v for _, v in model._train_vars]
    # (a) the new exchange
    model.forward_batch(feed_of(shard))
    active = model.reduce_gradients(True)
    got = [v.grad.clone() for v in variables]
    payload = model._flat_grads.payload_bytes
    # (b) the old exchange: local mean-normalised gradient, weighted by the shard's graph count
    for v in variables:
        v.grad = None
    loss, _ = model.forward_batch(feed_of(shard))
    loss.backward()
    parallel.allreduce_gradients(variables, weight=float(len(shard)))
    old = [v.grad.clone() for v in variables]
    # (c) the union batch on one rank
    for v in variables:
        v.grad = None
    loss, _ = model.forward_batch(feed_of(list(range(40))))
    loss.backward()
    rel = lambda a, b: float((a - b).abs().max()) / (float(b.abs().max()) + 1e-12)
    worst_new = max(rel(g, v.grad) for g, v in zip(got, variables))
    worst_old = max(rel(g, v.grad) for g, v in zip(old, variables))
    # (d) ragged epoch: rank 1 has one more batch than rank 0
    steps_taken = []
    for step in range(2):
        mine = shard[:4] if step == 0 else (None if rank == 0 else shard[4:12])
        if mine is None:
            steps_taken.append(model.train_step(None))
        else:
            loss, _ = model.forward_batch(feed_of(mine))
            steps_taken.append(model.train_step(loss))
    steps_taken.append(model.train_step(None))               # everybody is out of data: no update, returns 0
    checksum = float(sum(float(v.detach().double().sum()) for v in variables))
    out_q.put((rank, active, worst_new, worst_old, payload, steps_taken, checksum))
    dist.barrier()
    dist.destroy_process_group()


def test_one_allreduce_step_equals_the_union_batch_with_unequal_shards_and_unlabeled_graphs(tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_flat_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, active, worst_new, worst_old, payload, steps_taken, checksum in res:
        assert active == 2
        assert worst_new < 1e-4, (rank, worst_new)           # union-batch gradient, per-task mask sums
        assert worst_old > 1e-2, (rank, worst_old)           # graph-count weighting is a different gradient: the old code fails here
        assert steps_taken == [2, 1, 0], steps_taken         # ranks with different batch counts stay in lock step, then stop together
        assert payload > 0
    assert abs(res[0][6] - res[1][6]) < 1e-9 * max(1.0, abs(res[0][6]))   # replicas are still identical after the ragged epoch
