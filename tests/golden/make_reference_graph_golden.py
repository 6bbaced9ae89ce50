#!/usr/bin/env python
"""Fixtures computed by the REFERENCE'S OWN graph-building code (run HERE, in the build container; only the committed .npz travel).

``chem_tensorflow_sparse.py`` / ``chem_tensorflow_dense.py`` are imported unmodified from /root/reference with ``tests/golden/tf_shim.py``
registered as ``tensorflow``: ``prepare_specific_graph_model``, ``compute_final_node_representations`` and ``gated_regression`` run as
written, every ``tf.*`` call building a lazy node that is then evaluated with NumPy in float64 on a batch produced by the reference's own
``make_minibatch_iterator``.  The weights the reference code created (``tf.Variable`` values, the cells' kernels) are exported in the
engine's naming together with the outputs:

    refgraph_sparse_<case>.npz   inputs, per-layer weights, final node states [V, D], readout [G]
    refgraph_dense.npz           inputs, weights, final node states [b, v, D], readout [b]

tests/test_oracle.py requires the oracle (NumPy and C) to reproduce these to 1e-12: that pins the oracle's dataflow to the reference's
code.  The arithmetic INSIDE TensorFlow's ops (GRUCell, BasicRNNCell, segment sums) is restated in tf_shim.py from the 1.3 release, as in
the oracle itself -- TensorFlow cannot run here.

Usage:  python tests/golden/make_reference_graph_golden.py
"""
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import tf_shim  # noqa: E402
from gated_graph_neural_network_samples_b200 import synthetic  # noqa: E402


def import_reference():
    sys.modules["tensorflow"] = tf_shim
    tf_shim.register_submodules(sys.modules)
    d = types.ModuleType("docopt")
    d.docopt = lambda *a, **k: {}
    sys.modules["docopt"] = d
    sys.path.insert(0, "/root/reference")
    import chem_tensorflow_dense as ref_dense  # noqa
    import chem_tensorflow_sparse as ref_sparse  # noqa
    import utils as ref_utils  # noqa
    return ref_sparse, ref_dense, ref_utils


def build_model(m):
    """The reference's make_model (chem_tensorflow.py:133-170), unmodified: placeholders, the two hooks inside variable_scope
    "graph_model", the per-task readout MLPs, gated_regression, masked loss and accuracy."""
    m.placeholders, m.weights, m.ops = {}, {}, {}
    m.make_model()
    gate, trans = m.weights["regression_gate_task0"], m.weights["regression_transform_task0"]
    return {"w_gate": gate.params["weights"][0], "b_gate": gate.params["biases"][0],
            "w_trans": trans.params["weights"][0], "b_trans": trans.params["biases"][0]}


def evaluate_model(m, feed):
    feed[m.placeholders["out_layer_dropout_keep_prob"]] = 1.0
    final, loss, acc = tf_shim.evaluate([m.ops["final_node_representations"], m.ops["loss"], m.ops["accuracy_task0"]], feed)
    ro = tf_shim.evaluate(m.output, feed)                         # gated_regression stores its result in self.output
    return final, ro, float(loss), float(acc)


def seeded_weight(index, shape, scale=None):
    """Weights of the BASELINE-width fixtures that are too large to commit (D = 256: ~15 MB): value = float32(uniform(-r, r)) from
    RandomState(9000 + index), r = glorot range of the shape.  tests/_util.py regenerates them from (index, shape) alone."""
    r = np.sqrt(6.0 / (shape[-2] + shape[-1])) if scale is None else scale
    return np.random.RandomState(9000 + index).uniform(-r, r, size=shape).astype(np.float32).astype(np.float64)


def sparse_case_wide(ref_sparse, name, cfg, mols, T, store_weights):
    """A BASELINE-width case (cfg2 / cfg4 shape).  ``store_weights``: commit the float32-rounded weights the reference code created
    (D = 100), or replace them by ``seeded_weight`` values and commit nothing but the recipe (D = 256)."""
    m = object.__new__(ref_sparse.SparseGGNNChemModel)
    m.params = {"task_ids": [0], "tie_fwd_bkwd": True, "task_sample_ratios": {}, "batch_size": 100000, "out_layer_dropout_keep_prob": 1.0,
                "graph_state_dropout_keep_prob": 1.0, "edge_weight_dropout_keep_prob": 1.0, "use_propagation_attention": False, "use_graph": True}
    m.params.update(cfg)
    m.num_edge_types, m.annotation_size = T, len(mols[0]["node_features"][0])
    np.random.seed(21)
    tf_shim.CELL_RNG.seed(5151)
    n_before = len(tf_shim.VARIABLES)
    ro_w = build_model(m)                                         # make_model -> sparse:63-115, 117-218, 220-231, unmodified
    L, D = len(cfg["layer_timesteps"]), cfg["hidden_size"]
    data = m.process_raw_graphs(mols, is_training_data=False)
    feed = next(iter(m.make_minibatch_iterator(data, is_training=False)))
    h0 = np.asarray(feed[m.placeholders["initial_node_representation"]], dtype=np.float64)
    h0 = (h0 + np.random.RandomState(3).normal(0, 0.1, h0.shape)).astype(np.float32).astype(np.float64)   # exactly representable in fp32
    feed[m.placeholders["initial_node_representation"]] = h0
    evaluate_model(m, feed)                                       # first evaluation creates the cells' kernels (their width is only known then)
    edge_vars = [v for v in tf_shim.VARIABLES[n_before:] if "gnn_edge_weights" in str(getattr(v, "name", ""))]
    assert len(edge_vars) == L, [getattr(v, "name", None) for v in tf_shim.VARIABLES[n_before:]]
    idx = 0
    for l in range(L):
        cell = m.gnn_weights.rnn_cells[l]
        cell = getattr(cell, "cell", cell)
        arrays = [("edge", edge_vars[l], None)] + [(k, cell.vars, k) for k in sorted(cell.vars)]
        for tag, holder, key in arrays:
            cur = holder.value if key is None else holder[key]
            if store_weights:
                new = np.asarray(cur, np.float64).astype(np.float32).astype(np.float64)
            elif cur.ndim == 1:
                new = np.asarray(cur, np.float64)                 # biases keep the reference's initial values (gate 1.0, candidate 0)
            else:
                new = seeded_weight(idx, cur.shape)
            idx += 1
            if key is None:
                holder.value = new
            else:
                holder[key] = new
    out_final, out_ro, loss, acc = evaluate_model(m, feed)
    out = {"params_json": np.asarray(json.dumps(cfg)), "num_edge_types": np.int64(T), "h0": h0.astype(np.float32), "final": out_final,
           "readout": out_ro, "loss": np.float64(loss), "accuracy": np.float64(acc),
           "indeg": np.asarray(feed[m.placeholders["num_incoming_edges_per_type"]], np.float32),
           "graph_nodes_list": np.asarray(feed[m.placeholders["graph_nodes_list"]], np.int32),
           "num_graphs": np.int64(feed[m.placeholders["num_graphs"]]), "weights_stored": np.int64(1 if store_weights else 0)}
    for e, ph in enumerate(m.placeholders["adjacency_lists"]):
        out["adj%d" % e] = np.asarray(feed[ph], np.int32).reshape(-1, 2)
    for l in range(L):
        cell = m.gnn_weights.rnn_cells[l]
        cell = getattr(cell, "cell", cell)
        if store_weights:
            out["w%d_edge_weights" % l] = np.asarray(tf_shim.evaluate(m.gnn_weights.edge_weights[l], feed), np.float32)
            for k, v in cell.vars.items():
                out["w%d_%s" % (l, k)] = np.asarray(v, np.float32)
        else:
            for k, v in cell.vars.items():
                if v.ndim == 1:
                    out["w%d_%s" % (l, k)] = np.asarray(v, np.float32)
    for k, v in ro_w.items():
        out["ro_" + k] = v.value
    np.savez_compressed(os.path.join(HERE, "refgraph_sparse_%s.npz" % name), **out)
    print(name, "V=%d final max %.3f readout[:3] %s" % (h0.shape[0], np.abs(out_final).max(), np.round(out_ro[:3], 4)))


def sparse_case(ref_sparse, ref_utils, name, cfg, mols):
    m = object.__new__(ref_sparse.SparseGGNNChemModel)
    m.params = {"task_ids": [0], "tie_fwd_bkwd": True, "task_sample_ratios": {}, "batch_size": 100000, "out_layer_dropout_keep_prob": 1.0,
                "graph_state_dropout_keep_prob": 1.0, "edge_weight_dropout_keep_prob": 1.0, "use_propagation_attention": False, "use_graph": True}
    m.params.update(cfg)
    m.num_edge_types, m.annotation_size = 4, len(mols[0]["node_features"][0])
    np.random.seed(11)
    tf_shim.CELL_RNG.seed(4242)
    ro_w = build_model(m)                                         # make_model -> sparse:63-115, 117-218, 220-231, unmodified
    L, T, D = len(cfg["layer_timesteps"]), 4, cfg["hidden_size"]
    rng = np.random.RandomState(5)
    for b in m.gnn_weights.edge_biases:                           # the reference initialises zeros / ones: perturb so the paths are exercised
        b.value = rng.uniform(-0.1, 0.1, b.value.shape)
    for a in m.gnn_weights.edge_type_attention_weights:
        a.value = 1.0 + 0.5 * rng.uniform(-1, 1, a.value.shape)
    data = m.process_raw_graphs(mols, is_training_data=False)     # sparse:234-252
    feed = next(iter(m.make_minibatch_iterator(data, is_training=False)))   # sparse:278-350: keyed by the placeholder objects
    h0 = np.asarray(feed[m.placeholders["initial_node_representation"]], dtype=np.float64)
    h0 = h0 * 1.5 + np.random.RandomState(3).normal(0, 0.3, h0.shape)      # every column live, attention scores of order 1
    feed[m.placeholders["initial_node_representation"]] = h0
    out_final, out_ro, loss, acc = evaluate_model(m, feed)
    if cfg["graph_rnn_cell"].lower() == "cudnncompatiblegrucell":
        # the cell's variables exist after the first evaluation; its two candidate biases are initialised to zero: perturb, evaluate again
        for cell in m.gnn_weights.rnn_cells:
            cell.vars["cand_bias"] = rng.uniform(-0.1, 0.1, cell.vars["cand_bias"].shape)
            cell.vars["cand_hidden_bias"] = rng.uniform(-0.1, 0.1, cell.vars["cand_hidden_bias"].shape)
        out_final, out_ro, loss, acc = evaluate_model(m, feed)
    out = {"params_json": np.asarray(json.dumps(cfg)), "h0": h0, "final": out_final, "readout": out_ro, "loss": np.float64(loss),
           "accuracy": np.float64(acc), "target_values": np.asarray(feed[m.placeholders["target_values"]], np.float64),
           "target_mask": np.asarray(feed[m.placeholders["target_mask"]], np.float64),
           "indeg": np.asarray(feed[m.placeholders["num_incoming_edges_per_type"]], np.float64),
           "graph_nodes_list": np.asarray(feed[m.placeholders["graph_nodes_list"]], np.int32),
           "num_graphs": np.int64(feed[m.placeholders["num_graphs"]])}
    for e, ph in enumerate(m.placeholders["adjacency_lists"]):
        out["adj%d" % e] = np.asarray(feed[ph], np.int32).reshape(-1, 2)
    for l in range(L):
        out["w%d_edge_weights" % l] = tf_shim.evaluate(m.gnn_weights.edge_weights[l], feed)          # [T, D, D] after sparse:90-91
        if cfg["use_edge_bias"]:
            out["w%d_edge_biases" % l] = m.gnn_weights.edge_biases[l].value
        if cfg.get("use_propagation_attention"):
            out["w%d_edge_type_attention_weights" % l] = m.gnn_weights.edge_type_attention_weights[l].value
        cv = dict(m.gnn_weights.rnn_cells[l].vars or {})
        if "cand_input_kernel" in cv:   # CudnnCompatibleGRUCell: the engine's candidate kernel stacks [input_projection ; hidden_projection]
            cv["cand_kernel"] = np.concatenate([cv.pop("cand_input_kernel"), cv.pop("cand_hidden_kernel")], axis=0)
        for k, v in cv.items():
            out["w%d_%s" % (l, k)] = v
    for k, v in ro_w.items():
        out["ro_" + k] = v.value
    np.savez_compressed(os.path.join(HERE, "refgraph_sparse_%s.npz" % name), **out)
    print(name, "V=%d final max %.3f readout[:3] %s loss %.5f mae %.5f" % (h0.shape[0], np.abs(out_final).max(), np.round(out_ro[:3], 4), loss, acc))


def dense_case(ref_dense, ref_utils, mols):
    cfg = {"hidden_size": 12, "num_timesteps": 3, "use_edge_bias": True}
    m = object.__new__(ref_dense.DenseGGNNChemModel)
    m.params = {"task_ids": [0], "tie_fwd_bkwd": True, "task_sample_ratios": {}, "batch_size": 4, "out_layer_dropout_keep_prob": 1.0,
                "graph_state_dropout_keep_prob": 1.0, "edge_weight_dropout_keep_prob": 1.0, "use_graph": True}
    m.params.update(cfg)
    m.num_edge_types, m.annotation_size = 4, len(mols[0]["node_features"][0])
    np.random.seed(12)
    tf_shim.CELL_RNG.seed(777)
    ro_w = build_model(m)                                         # make_model -> dense:68-91, 93-117, 119-129, unmodified
    m.weights["edge_biases"].value = np.random.RandomState(6).uniform(-0.1, 0.1, m.weights["edge_biases"].value.shape)
    data = m.process_raw_graphs(mols, is_training_data=False)     # dense:132-164 (buckets)
    feed = next(iter(m.make_minibatch_iterator(data, is_training=False)))
    h0 = np.asarray(feed[m.placeholders["initial_node_representation"]], dtype=np.float64)
    h0 = h0 + np.random.RandomState(4).normal(0, 0.1, h0.shape)
    feed[m.placeholders["initial_node_representation"]] = h0
    out_final, out_ro, loss, acc = evaluate_model(m, feed)
    out = {"params_json": np.asarray(json.dumps(cfg)), "h0": h0, "final": out_final, "readout": out_ro, "loss": np.float64(loss),
           "accuracy": np.float64(acc), "target_values": np.asarray(feed[m.placeholders["target_values"]], np.float64),
           "target_mask": np.asarray(feed[m.placeholders["target_mask"]], np.float64),
           "adj": np.asarray(feed[m.placeholders["adjacency_matrix"]], np.float64),
           "node_mask": np.asarray(feed[m.placeholders["node_mask"]], np.float64),
           "w_edge_weights": m.weights["edge_weights"].value, "w_edge_biases": m.weights["edge_biases"].value}
    for k, v in m.weights["node_gru"].vars.items():
        out["w_" + k] = v
    for k, v in ro_w.items():
        out["ro_" + k] = v.value
    np.savez_compressed(os.path.join(HERE, "refgraph_dense.npz"), **out)
    print("dense b,v =", h0.shape[:2], "final max %.3f readout[:3] %s" % (np.abs(out_final).max(), np.round(out_ro[:3], 4)))


def dense_case_wide(ref_dense, mols):
    """BASELINE configs[2] width (hidden 100, 4 timesteps, edge bias) through the reference's dense graph code; the float32-rounded
    weights the reference code created are committed, h0 is exactly representable in fp32."""
    cfg = {"hidden_size": 100, "num_timesteps": 4, "use_edge_bias": True}
    m = object.__new__(ref_dense.DenseGGNNChemModel)
    m.params = {"task_ids": [0], "tie_fwd_bkwd": True, "task_sample_ratios": {}, "batch_size": 8, "out_layer_dropout_keep_prob": 1.0,
                "graph_state_dropout_keep_prob": 1.0, "edge_weight_dropout_keep_prob": 1.0, "use_graph": True}
    m.params.update(cfg)
    m.num_edge_types, m.annotation_size = 4, len(mols[0]["node_features"][0])
    np.random.seed(13)
    tf_shim.CELL_RNG.seed(778)
    ro_w = build_model(m)                                         # make_model -> dense:68-91, 93-117, 119-129, unmodified
    f32 = lambda a: np.asarray(a, np.float64).astype(np.float32).astype(np.float64)
    m.weights["edge_weights"].value = f32(m.weights["edge_weights"].value)
    m.weights["edge_biases"].value = f32(np.random.RandomState(7).uniform(-0.1, 0.1, m.weights["edge_biases"].value.shape))
    data = m.process_raw_graphs(mols, is_training_data=False)
    feed = next(iter(m.make_minibatch_iterator(data, is_training=False)))
    h0 = np.asarray(feed[m.placeholders["initial_node_representation"]], dtype=np.float64)
    h0 = f32(h0 + np.random.RandomState(5).normal(0, 0.1, h0.shape))
    feed[m.placeholders["initial_node_representation"]] = h0
    evaluate_model(m, feed)                                       # first evaluation creates the cell's kernels
    for k in list(m.weights["node_gru"].vars):
        m.weights["node_gru"].vars[k] = f32(m.weights["node_gru"].vars[k])
    out_final, out_ro, loss, acc = evaluate_model(m, feed)
    out = {"params_json": np.asarray(json.dumps(cfg)), "h0": h0.astype(np.float32), "final": out_final, "readout": out_ro, "loss": np.float64(loss),
           "accuracy": np.float64(acc), "adj": np.asarray(feed[m.placeholders["adjacency_matrix"]], np.float32),
           "node_mask": np.asarray(feed[m.placeholders["node_mask"]], np.float32),
           "w_edge_weights": m.weights["edge_weights"].value.astype(np.float32), "w_edge_biases": m.weights["edge_biases"].value.astype(np.float32)}
    for k, v in m.weights["node_gru"].vars.items():
        out["w_" + k] = np.asarray(v, np.float32)
    for k, v in ro_w.items():
        out["ro_" + k] = v.value
    np.savez_compressed(os.path.join(HERE, "refgraph_dense_cfg3_shape.npz"), **out)
    print("dense cfg3 shape b,v =", h0.shape[:2], "final max %.3f readout[:3] %s" % (np.abs(out_final).max(), np.round(out_ro[:3], 4)))


def main():
    ref_sparse, ref_dense, ref_utils = import_reference()
    mols = synthetic.make_molecules(12, seed=321)
    mols[4]["targets"][0][0] = None                              # one unlabeled graph: the loss mask is exercised
    cases = {
        "true_default_shape": {"hidden_size": 12, "layer_timesteps": [2, 2, 1, 2, 1], "residual_connections": {"2": [0], "4": [0, 2]},
                               "use_edge_bias": False, "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU",
                               "graph_rnn_activation": "tanh"},
        "rnn_relu_bias_sum": {"hidden_size": 8, "layer_timesteps": [1, 2], "residual_connections": {"1": [0]},
                              "use_edge_bias": True, "use_edge_msg_avg_aggregation": False, "graph_rnn_cell": "RNN",
                              "graph_rnn_activation": "ReLU"},
        "attention_bias_avg": {"hidden_size": 12, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                               "use_edge_bias": True, "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU",
                               "graph_rnn_activation": "tanh", "use_propagation_attention": True},
        # sparse:105-108 -- tf.contrib.cudnn_rnn.CudnnCompatibleGRUCell (reset gate after the recurrent matmul)
        "cudnn_gru": {"hidden_size": 12, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                      "use_edge_bias": True, "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "CudnnCompatibleGRUCell",
                      "graph_rnn_activation": "tanh"},
        # more branch combinations of sparse:75-81,102-112,170-209 (oracle-level pins only; the engine is held to the cases above)
        "gru_relu_sum_nobias": {"hidden_size": 8, "layer_timesteps": [3], "residual_connections": {}, "use_edge_bias": False,
                                "use_edge_msg_avg_aggregation": False, "graph_rnn_cell": "gru", "graph_rnn_activation": "relu"},
        "rnn_tanh_avg_two_residuals": {"hidden_size": 8, "layer_timesteps": [1, 1, 2], "residual_connections": {"2": [0, 1]},
                                       "use_edge_bias": False, "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "RNN",
                                       "graph_rnn_activation": "tanh"},
        "attention_rnn_sum": {"hidden_size": 8, "layer_timesteps": [2], "residual_connections": {}, "use_edge_bias": False,
                              "use_edge_msg_avg_aggregation": False, "graph_rnn_cell": "RNN", "graph_rnn_activation": "tanh",
                              "use_propagation_attention": True},
    }
    for name, cfg in cases.items():
        sparse_case(ref_sparse, ref_utils, name, cfg, mols)
    dense_case(ref_dense, ref_utils, synthetic.make_molecules(40, seed=123))   # batch_size 4: several buckets fill (dense:150-164)
    # BASELINE widths (remove the oracle bridge between the D = 12 fixtures and the D = 100 / 256 checks)
    sparse_case_wide(ref_sparse, "cfg2_shape", {"hidden_size": 100, "layer_timesteps": [4], "residual_connections": {}, "use_edge_bias": False,
                                                "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"},
                     synthetic.make_molecules(24, seed=77), T=4, store_weights=True)
    sparse_case_wide(ref_sparse, "cfg4_shape", {"hidden_size": 256, "layer_timesteps": [2, 2, 2, 2], "residual_connections": {"2": [0]},
                                                "use_edge_bias": False, "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU",
                                                "graph_rnn_activation": "tanh"},
                     synthetic.make_molecules(12, seed=78, num_bond_types=8), T=8, store_weights=False)
    dense_case_wide(ref_dense, synthetic.make_molecules(32, seed=125))
    print("reference-graph fixtures written to", HERE)


if __name__ == "__main__":
    main()
