"""The BASELINE.json configurations as concrete, seeded synthetic workloads + their algorithmic cost.

Each workload is the input of ONE propagation step batch (``compute_final_node_representations``):
reference-format arrays (``packing.py``), the reference's parameter dict, and seeded weights.
``algorithmic_bytes`` / ``algorithmic_flops`` are SURVEY.md section 8(d)'s per-step figures summed over the
forward -- the numerator of ``roofline.achieved`` in bench.py.
"""
from __future__ import annotations

import numpy as np

from . import packing, synthetic
from .engine import residual_inputs_of_layer

SPARSE_BASE = {"use_edge_bias": False, "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU",
               "graph_rnn_activation": "tanh", "residual_connections": {}}

CONFIGS = {
    # BASELINE.json configs[1] (and configs[0], the CPU-runnable statement of the same thing)
    "cfg2": dict(kind="sparse", molecules=256, edge_types=4,
                 params=dict(SPARSE_BASE, hidden_size=100, layer_timesteps=[4])),
    # the sparse file's TRUE defaults (sparse:44-60): [2,2,1,2,1] + residuals
    "cfg1_true_default": dict(kind="sparse", molecules=256, edge_types=4,
                              params=dict(SPARSE_BASE, hidden_size=100, layer_timesteps=[2, 2, 1, 2, 1],
                                          residual_connections={"2": [0], "4": [0, 2]})),
    "cfg3_dense": dict(kind="dense", molecules=256, edge_types=4, max_nodes=32,
                       params=dict(hidden_size=100, num_timesteps=4, use_edge_bias=True)),
    "cfg4": dict(kind="sparse", molecules=1024, edge_types=8,
                 params=dict(SPARSE_BASE, hidden_size=256, layer_timesteps=[2, 2, 2, 2],
                             residual_connections={"2": [0]})),
    "cfg5_rgcn": dict(kind="single_graph", nodes=10000, undirected_edges=40000, edge_types=4,
                      params=dict(SPARSE_BASE, hidden_size=100, layer_timesteps=[1] * 8, graph_rnn_cell="RNN",
                                  graph_rnn_activation="ReLU")),
    # supplementary HBM-resident point: the reference's real default batch_size = 100 000 nodes (sparse:44)
    "default_batch_100k_nodes": dict(kind="sparse", molecules=5500, edge_types=4,
                                     params=dict(SPARSE_BASE, hidden_size=100, layer_timesteps=[4])),
}


def glorot(shape, rng):
    r = np.sqrt(6.0 / (shape[-2] + shape[-1]))  # utils.py:11-13
    return rng.uniform(-r, r, size=shape).astype(np.float32)


def init_weights(params: dict, num_edge_types: int, seed: int = 1, edge_bias_scale: float = 0.1):
    """Seeded weights with the reference's shapes/initialisers (sparse:86-115; TF-1.3 cell defaults: glorot
    kernels, gate bias 1.0, candidate bias 0).  Keys follow ``ggnn_layer_weights``."""
    rng = np.random.default_rng(seed)
    D, T = int(params["hidden_size"]), int(num_edge_types)
    layers = []
    for l in range(len(params["layer_timesteps"])):
        din = D * (1 + len(residual_inputs_of_layer(params, l)))
        w = {"edge_weights": glorot([T * D, D], rng).reshape(T, D, D)}
        if params.get("use_edge_bias", False):
            w["edge_biases"] = rng.uniform(-edge_bias_scale, edge_bias_scale, size=(T, D)).astype(np.float32)
        if params.get("graph_rnn_cell", "GRU").lower() == "gru":
            w["gate_kernel"] = glorot([din + D, 2 * D], rng)
            w["gate_bias"] = np.ones(2 * D, np.float32)
        w["cand_kernel"] = glorot([din + D, D], rng)
        w["cand_bias"] = np.zeros(D, np.float32)
        layers.append(w)
    return layers


def dense_engine_params(params: dict) -> dict:
    """The dense model (dense:93-117) in the engine's vocabulary: one layer of num_timesteps steps, edge
    bias (times row-sum of A), no averaging, GRU/tanh."""
    return {"hidden_size": int(params["hidden_size"]), "layer_timesteps": [int(params["num_timesteps"])],
            "residual_connections": {}, "use_edge_bias": bool(params.get("use_edge_bias", True)),
            "use_edge_msg_avg_aggregation": False, "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}


def build(name: str, seed: int = 0, scale: float = 1.0, shard=None) -> dict:
    """Materialise a workload.  ``scale`` multiplies the number of molecules (sampling); ``shard=(rank, world)`` keeps this rank's
    contiguous, node-balanced range of the molecule list (parallel.shard_bounds: strong scaling of one fixed batch)."""
    cfg = CONFIGS[name]
    T = cfg["edge_types"]
    params = dict(cfg["params"])
    D = int(params["hidden_size"])
    rng = np.random.default_rng(seed + 77)
    out = {"name": name, "kind": cfg["kind"], "num_edge_types": T, "params": params}
    if cfg["kind"] == "sparse":
        n = max(1, int(round(cfg["molecules"] * scale)))
        mols = synthetic.make_molecules(n, seed=seed, num_bond_types=T)
        if shard is not None:
            from . import parallel
            mols = parallel.shard_graphs(mols, int(shard[0]), int(shard[1]))
            n = len(mols)
        out["molecules"] = mols
        b = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), D, T)
        out["target_values"] = np.asarray([m["targets"][0][0] for m in mols], np.float32)
        out.update(engine_params=params, num_graphs=n, adjacency_lists=b["adjacency_lists"],
                   num_incoming_edges_per_type=b["num_incoming_edges_per_type"],
                   h0=b["initial_node_representation"], graph_nodes_list=b["graph_nodes_list"])
    elif cfg["kind"] == "single_graph":
        adj, indeg = synthetic.random_sparse_graph(cfg["nodes"], cfg["undirected_edges"], T, seed=seed)
        out.update(engine_params=params, num_graphs=1, adjacency_lists=adj, num_incoming_edges_per_type=indeg,
                   h0=rng.normal(0, 0.1, size=(cfg["nodes"], D)).astype(np.float32))
    else:
        n = max(1, int(round(cfg["molecules"] * scale)))
        mols = synthetic.make_molecules(n, seed=seed, num_bond_types=T)
        b = packing.pack_dense_batch(mols, cfg["max_nodes"], D, T)
        out.update(engine_params=dense_engine_params(params), num_graphs=n, adjacency_matrix=b["adjacency_matrix"],
                   h0=b["initial_node_representation"].reshape(-1, D), node_mask=b["node_mask"],
                   dense_shape=(n, cfg["max_nodes"]))
    out["weights"] = init_weights(out["engine_params"], T, seed=1)
    out["V"] = int(out["h0"].shape[0])
    out["M"] = int(sum(a.shape[0] for a in out.get("adjacency_lists", [])))
    out["timesteps"] = int(sum(out["engine_params"]["layer_timesteps"]))
    out["node_updates"] = out["V"] * out["timesteps"]
    return out


def union_of(name: str, seeds) -> dict:
    """ONE batch holding the molecules of ``build(name, seed=s)`` for every s in ``seeds``, in that order: the union batch a
    data-parallel step over those shards must reproduce."""
    cfg = CONFIGS[name]
    assert cfg["kind"] == "sparse"
    T, params = cfg["edge_types"], dict(cfg["params"])
    D = int(params["hidden_size"])
    mols = []
    for s in seeds:
        mols += synthetic.make_molecules(cfg["molecules"], seed=s, num_bond_types=T)
    b = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), D, T)
    out = {"name": name, "kind": "sparse", "num_edge_types": T, "params": params, "engine_params": params, "num_graphs": len(mols),
           "adjacency_lists": b["adjacency_lists"], "num_incoming_edges_per_type": b["num_incoming_edges_per_type"],
           "h0": b["initial_node_representation"], "graph_nodes_list": b["graph_nodes_list"],
           "target_values": np.asarray([m["targets"][0][0] for m in mols], np.float32)}
    out["weights"] = init_weights(params, T, seed=1)
    out["V"] = int(out["h0"].shape[0])
    out["M"] = int(sum(a.shape[0] for a in out["adjacency_lists"]))
    return out


def algorithmic_bytes(w: dict) -> int:
    """per step 4*D*(2V + M + R*V) + 8*M + 4*V*T + 4*(T*D^2 [+T*D] + (Din+D)*G*D + G*D);
    dense: 4*D*2*b*v + 4*b*T*v^2 + weights.  Summed over all layers/steps of one forward."""
    p = w["engine_params"]
    D, T, V, M = int(p["hidden_size"]), w["num_edge_types"], w["V"], w["M"]
    G = 3 if p.get("graph_rnn_cell", "GRU").lower() == "gru" else 1
    total = 0
    for l, steps in enumerate(p["layer_timesteps"]):
        R = len(residual_inputs_of_layer(p, l))
        din = D * (1 + R)
        weights = 4 * (T * D * D + (T * D if p.get("use_edge_bias") else 0) + (din + D) * G * D + G * D)
        if w["kind"] == "dense":
            b, v = w["dense_shape"]
            step = 4 * D * 2 * b * v + 4 * b * T * v * v + weights
        else:
            step = 4 * D * (2 * V + M + R * V) + 8 * M + 4 * V * T + weights
        total += steps * step
    return int(total)


def algorithmic_flops(w: dict) -> int:
    """Reference formulation (gather-then-matmul): 2*M*D^2 + 2*V*(Din+D)*G*D per step (dense: 2*V*D*D*T +
    2*b*T*v*v*D + GRU)."""
    p = w["engine_params"]
    D, T, V, M = int(p["hidden_size"]), w["num_edge_types"], w["V"], w["M"]
    G = 3 if p.get("graph_rnn_cell", "GRU").lower() == "gru" else 1
    total = 0
    for l, steps in enumerate(p["layer_timesteps"]):
        din = D * (1 + len(residual_inputs_of_layer(p, l)))
        cell = 2 * V * (din + D) * G * D
        if w["kind"] == "dense":
            b, v = w["dense_shape"]
            msg = 2 * V * D * D * T + 2 * b * T * v * v * D
        else:
            msg = 2 * M * D * D
        total += steps * (msg + cell)
    return int(total)
