#!/usr/bin/env python
"""Regenerates the fixtures in tests/golden/ (run HERE, in the build container; /root/reference does not
exist on the GPU box, so only the committed .npz/.json outputs travel).

1. ``packing_sparse.npz`` / ``packing_dense.npz`` -- batches produced by the REFERENCE'S OWN NumPy packing
   code (``chem_tensorflow_sparse.py:234-350``, ``chem_tensorflow_dense.py:30-36,132-228``) imported from
   /root/reference with ``tensorflow`` and ``docopt`` stubbed (neither is installable here; the packers
   are pure NumPy/Python and never touch them).  These pin ``packing.py`` and the wire format.
2. ``prop_*.npz`` -- inputs, seeded weights and float64 outputs of ``oracle.ggnn_oracle`` for small
   propagation configs (the reference holds no golden vectors for the hot path -- "parity unpinned",
   SURVEY 8c -- so these freeze the oracle against regressions and give the CUDA tests a fixed target).

Usage:  python tests/golden/make_golden.py
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from gated_graph_neural_network_samples_b200 import synthetic  # noqa: E402
from oracle import ggnn_oracle as O  # noqa: E402


def import_reference():
    for name in ("tensorflow", "docopt"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.docopt = lambda *a, **k: {}
            m.Tensor = object
            sys.modules[name] = m
    sys.path.insert(0, "/root/reference")
    import chem_tensorflow_sparse as ref_sparse  # noqa
    import chem_tensorflow_dense as ref_dense  # noqa
    return ref_sparse, ref_dense


def reference_sparse_batches(ref_sparse, mols, hidden, batch_size_nodes, num_edge_types):
    m = object.__new__(ref_sparse.SparseGGNNChemModel)
    m.params = {"task_ids": [0], "tie_fwd_bkwd": True, "task_sample_ratios": {}, "batch_size": batch_size_nodes,
                "hidden_size": hidden, "graph_state_dropout_keep_prob": 1.0, "edge_weight_dropout_keep_prob": 1.0}
    m.num_edge_types = num_edge_types
    m.annotation_size = len(mols[0]["node_features"][0])
    keys = ["initial_node_representation", "num_incoming_edges_per_type", "graph_nodes_list", "target_values",
            "target_mask", "num_graphs", "graph_state_keep_prob", "edge_weight_dropout_keep_prob"]
    m.placeholders = {k: k for k in keys}
    m.placeholders["adjacency_lists"] = ["adjacency_lists_%d" % e for e in range(num_edge_types)]
    data = m.process_raw_graphs(mols, is_training_data=False)
    return list(m.make_minibatch_iterator(data, is_training=False))


def reference_dense_batches(ref_dense, mols, hidden, batch_size, num_edge_types):
    m = object.__new__(ref_dense.DenseGGNNChemModel)
    m.params = {"task_ids": [0], "tie_fwd_bkwd": True, "task_sample_ratios": {}, "batch_size": batch_size,
                "hidden_size": hidden, "graph_state_dropout_keep_prob": 1.0}
    m.num_edge_types = num_edge_types
    m.annotation_size = len(mols[0]["node_features"][0])
    keys = ["initial_node_representation", "target_values", "target_mask", "num_graphs", "num_vertices",
            "adjacency_matrix", "node_mask", "graph_state_keep_prob", "edge_weight_dropout_keep_prob"]
    m.placeholders = {k: k for k in keys}
    data = m.process_raw_graphs(mols, is_training_data=False)
    return data, list(m.make_minibatch_iterator(data, is_training=False))


def main():
    ref_sparse, ref_dense = import_reference()
    mols = synthetic.make_molecules(40, seed=123)
    with open(os.path.join(HERE, "molecules_40.json"), "w") as f:
        json.dump(mols, f)

    # -------- 1a. sparse packing from the reference's own code
    batches = reference_sparse_batches(ref_sparse, mols, hidden=8, batch_size_nodes=200, num_edge_types=4)
    out = {"num_batches": np.int64(len(batches))}
    for bi, b in enumerate(batches):
        out["b%d_init" % bi] = np.asarray(b["initial_node_representation"], dtype=np.float32)
        out["b%d_indeg" % bi] = np.asarray(b["num_incoming_edges_per_type"], dtype=np.float32)
        out["b%d_gnl" % bi] = np.asarray(b["graph_nodes_list"], dtype=np.int32)
        out["b%d_num_graphs" % bi] = np.int64(b["num_graphs"])
        out["b%d_targets" % bi] = np.asarray(b["target_values"], dtype=np.float32)
        out["b%d_mask" % bi] = np.asarray(b["target_mask"], dtype=np.float32)
        for e in range(4):
            out["b%d_adj%d" % (bi, e)] = np.asarray(b["adjacency_lists_%d" % e], dtype=np.int32).reshape(-1, 2)
    np.savez_compressed(os.path.join(HERE, "packing_sparse.npz"), **out)

    # -------- 1b. dense packing from the reference's own code (batch_size 4 so several buckets fill)
    (bucketed, bucket_sizes, bucket_at_step), dbatches = reference_dense_batches(ref_dense, mols, 8, 4, 4)
    out = {"num_batches": np.int64(len(dbatches)), "bucket_sizes": np.asarray(bucket_sizes)}
    # which molecules went into which batch: recover through the (unique) targets
    tgt_to_idx = {m["targets"][0][0]: i for i, m in enumerate(mols)}
    for bi, b in enumerate(dbatches):
        out["b%d_init" % bi] = np.asarray(b["initial_node_representation"], dtype=np.float32)
        out["b%d_adj" % bi] = np.asarray(b["adjacency_matrix"], dtype=np.float32)
        out["b%d_mask" % bi] = np.asarray(b["node_mask"], dtype=np.float32)
        out["b%d_num_vertices" % bi] = np.int64(b["num_vertices"])
        out["b%d_mol_idx" % bi] = np.asarray([tgt_to_idx[float(t)] for t in np.asarray(b["target_values"])[0]])
    np.savez_compressed(os.path.join(HERE, "packing_dense.npz"), **out)

    # -------- 2. oracle propagation vectors
    from gated_graph_neural_network_samples_b200 import packing
    proc = packing.process_raw_graphs_sparse(mols[:10])
    cases = {
        "gru_bias_avg_res": {"hidden_size": 12, "layer_timesteps": [2, 1, 2], "residual_connections": {"1": [0], "2": [0, 1]},
                             "use_edge_bias": True, "use_edge_msg_avg_aggregation": True,
                             "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"},
        "gru_plain": {"hidden_size": 20, "layer_timesteps": [4], "residual_connections": {},
                      "use_edge_bias": False, "use_edge_msg_avg_aggregation": False,
                      "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"},
        "rgcn_relu": {"hidden_size": 16, "layer_timesteps": [1, 1, 1], "residual_connections": {},
                      "use_edge_bias": False, "use_edge_msg_avg_aggregation": True,
                      "graph_rnn_cell": "RNN", "graph_rnn_activation": "ReLU"},
    }
    for name, params in cases.items():
        batch = packing.pack_sparse_batch(proc, params["hidden_size"], 4)
        rng = np.random.default_rng(7)
        h0 = batch["initial_node_representation"].copy()
        h0 += rng.normal(0, 0.1, size=h0.shape).astype(np.float32)  # exercise every column
        weights = O.init_sparse_weights(params, 4, np.random.default_rng(1))
        states = O.sparse_propagation_loops(h0, batch["adjacency_lists"], batch["num_incoming_edges_per_type"],
                                            weights, params, return_all_layers=True)
        out = {"params_json": np.asarray(json.dumps(params)), "h0": h0,
               "indeg": batch["num_incoming_edges_per_type"], "final": states[-1]}
        for e in range(4):
            out["adj%d" % e] = batch["adjacency_lists"][e]
        for li, w in enumerate(weights):
            for k, v in w.items():
                out["w%d_%s" % (li, k)] = v
        for li, s in enumerate(states):
            out["state%d" % li] = s
        np.savez_compressed(os.path.join(HERE, "prop_sparse_%s.npz" % name), **out)

    # dense vector
    dparams = {"hidden_size": 12, "num_timesteps": 3, "use_edge_bias": True}
    db = packing.pack_dense_batch(mols[:6], 29, 12, 4)
    rng = np.random.default_rng(9)
    h0 = db["initial_node_representation"] + rng.normal(0, 0.1, size=db["initial_node_representation"].shape).astype(np.float32)
    dw = O.init_dense_weights(dparams, 4, np.random.default_rng(2))
    final = O.dense_propagation_loops(h0, db["adjacency_matrix"], dw, dparams)
    out = {"params_json": np.asarray(json.dumps(dparams)), "h0": h0, "adj": db["adjacency_matrix"], "final": final}
    for k, v in dw.items():
        out["w_%s" % k] = v
    np.savez_compressed(os.path.join(HERE, "prop_dense.npz"), **out)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
