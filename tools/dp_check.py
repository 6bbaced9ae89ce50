"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL): graphs sharded across ranks, ONE gradient all-reduce,
result equals the single-GPU gradient of the union batch (the loss is a mean over graphs, so ranks are weighted by
their graph counts).  Prints PASS/FAIL on rank 0."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from gated_graph_neural_network_samples_b200 import parallel, synthetic
from gated_graph_neural_network_samples_b200.chem_sparse import SparseGGNNChemModel


def build(mols, tmp):
    args = {"--log_dir": tmp, "--train_data": mols, "--valid_data": mols[:4], "--device": int(os.environ.get("LOCAL_RANK", 0)),
            "--precision": "bf16x3",
            "--config": {"hidden_size": 64, "batch_size": 100000, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                         "edge_weight_dropout_keep_prob": 1.0, "random_seed": 3}}
    return SparseGGNNChemModel(args)


def grads_of(model, data):
    batch = next(iter(model.make_minibatch_iterator(list(data), is_training=False)))
    batch["out_layer_dropout_keep_prob"] = 1.0
    for _, v in model._train_vars:
        v.grad = None
    loss, _ = model.forward_batch(batch)
    loss.backward()
    return float(loss), batch["num_graphs"]


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    world = int(os.environ["WORLD_SIZE"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    mols = synthetic.make_molecules(96, seed=11)
    tmp = "/tmp/dp_check_%d" % rank
    model = build(mols, tmp)                       # same seed on every rank -> identical replicas
    shard = parallel.shard_graphs(mols, rank, world)
    # every rank: gradient of its own shard, then the single all-reduce
    proc = model.process_raw_graphs(shard, is_training_data=False)
    loss_r, n_r = grads_of(model, proc)
    parallel.allreduce_gradients([v for _, v in model._train_vars], weight=float(n_r))
    got = [None if v.grad is None else v.grad.clone() for _, v in model._train_vars]
    # reference: the union batch on one GPU
    proc_all = model.process_raw_graphs(mols, is_training_data=False)
    grads_of(model, proc_all)
    worst = 0.0
    for (name, v), g in zip(model._train_vars, got):
        if v.grad is None:
            continue
        scale = float(v.grad.abs().max()) + 1e-12
        worst = max(worst, float((g - v.grad).abs().max()) / scale)
    t = torch.tensor([worst], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print("dp_check world=%d shard graphs=%d worst relative gradient mismatch %.2e -> %s" % (world, n_r, float(t), "PASS" if float(t) < 2e-4 else "FAIL"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
