"""Pins the CPU oracle (the reference holds no tests/golden vectors for the hot path)."""
import json
import os

import numpy as np
import pytest

from gated_graph_neural_network_samples_b200 import packing, synthetic
from oracle import c_oracle as CO
from oracle import ggnn_oracle as O
from tests import _util as U

PARAM_SETS = {
    "default_true": {"hidden_size": 10, "layer_timesteps": [2, 2, 1, 2, 1], "residual_connections": {"2": [0], "4": [0, 2]},
                     "use_edge_bias": False, "use_edge_msg_avg_aggregation": True,
                     "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"},
    "bias_noavg": {"hidden_size": 9, "layer_timesteps": [3], "residual_connections": {},
                   "use_edge_bias": True, "use_edge_msg_avg_aggregation": False,
                   "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"},
    "rgcn": {"hidden_size": 8, "layer_timesteps": [1, 1], "residual_connections": {},
             "use_edge_bias": False, "use_edge_msg_avg_aggregation": True,
             "graph_rnn_cell": "RNN", "graph_rnn_activation": "ReLU"},
    "gru_relu_bias_avg": {"hidden_size": 7, "layer_timesteps": [2, 1], "residual_connections": {"1": [0, 1]},
                          "use_edge_bias": True, "use_edge_msg_avg_aggregation": True,
                          "graph_rnn_cell": "gru", "graph_rnn_activation": "relu"},
    # sparse:105-108: tf.contrib.cudnn_rnn.CudnnCompatibleGRUCell (reset gate applied after the recurrent matmul), residual input
    "cudnn_gru": {"hidden_size": 8, "layer_timesteps": [2, 2], "residual_connections": {"1": [0]},
                  "use_edge_bias": True, "use_edge_msg_avg_aggregation": True,
                  "graph_rnn_cell": "CudnnCompatibleGRUCell", "graph_rnn_activation": "tanh"},
}


def _batch(D, n=6, seed=3, T=4):
    mols = synthetic.make_molecules(n, seed=seed, num_bond_types=T)
    b = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), D, T)
    rng = np.random.default_rng(seed)
    b["initial_node_representation"] = b["initial_node_representation"] + rng.normal(0, 0.2, b["initial_node_representation"].shape).astype(np.float32)
    return mols, b


@pytest.mark.parametrize("name", sorted(PARAM_SETS))
def test_loops_vs_vectorised_vs_torch(name):
    p = PARAM_SETS[name]
    _, b = _batch(p["hidden_size"])
    w = O.init_sparse_weights(p, 4, np.random.default_rng(1))
    args = (b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"], w, p)
    ref = O.sparse_propagation_loops(*args)
    np.testing.assert_allclose(O.sparse_propagation_np(*args, dtype=np.float64), ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(O.sparse_propagation_np(*args, dtype=np.float32), ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(O.sparse_propagation_torch(*args).numpy(), ref, rtol=2e-5, atol=2e-6)
    np.testing.assert_allclose(CO.sparse_propagation_c(*args), ref, rtol=1e-12, atol=1e-12)   # the independent plain-C restatement


def test_sparse_equals_dense_cross_implementation():
    """sparse (bias x in-degree, no averaging, one layer) == dense on the real nodes."""
    D, T, steps = 8, 4, 3
    mols = synthetic.make_molecules(5, seed=11)
    sb = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), D, T)
    db = packing.pack_dense_batch(mols, 29, D, T)
    dw = O.init_dense_weights({"hidden_size": D}, T, np.random.default_rng(5))
    sw = [{"edge_weights": dw["edge_weights"], "edge_biases": dw["edge_biases"].reshape(T, D),
           "gate_kernel": dw["gate_kernel"], "gate_bias": dw["gate_bias"],
           "cand_kernel": dw["cand_kernel"], "cand_bias": dw["cand_bias"]}]
    sp = {"hidden_size": D, "layer_timesteps": [steps], "residual_connections": {}, "use_edge_bias": True,
          "use_edge_msg_avg_aggregation": False, "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}
    hs = O.sparse_propagation_loops(sb["initial_node_representation"], sb["adjacency_lists"],
                                    sb["num_incoming_edges_per_type"], sw, sp)
    hd = O.dense_propagation_loops(db["initial_node_representation"], db["adjacency_matrix"], dw,
                                   {"num_timesteps": steps, "use_edge_bias": True})
    real = db["node_mask"].astype(bool)
    np.testing.assert_allclose(hd[real], hs, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(O.dense_propagation_torch(db["initial_node_representation"], db["adjacency_matrix"], dw,
                                                         {"num_timesteps": steps}).numpy(), hd, rtol=2e-5, atol=2e-6)


def test_hand_derived_two_node_gru():
    """2 nodes, one edge each way, W = I, K_g = 0 (=> r = u = sigmoid(1)), K_c picks the message block:
    h'_v = u*h_v + (1-u)*tanh(h_other)."""
    D = 3
    p = {"hidden_size": D, "layer_timesteps": [1], "residual_connections": {}, "use_edge_bias": False,
         "use_edge_msg_avg_aggregation": False, "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}
    kc = np.zeros((2 * D, D)); kc[:D] = np.eye(D)
    w = [{"edge_weights": np.stack([np.eye(D)]), "gate_kernel": np.zeros((2 * D, 2 * D)), "gate_bias": np.ones(2 * D),
          "cand_kernel": kc, "cand_bias": np.zeros(D)}]
    h0 = np.array([[0.1, -0.2, 0.3], [0.5, 0.0, -1.0]])
    adj = [np.array([[0, 1], [1, 0]], np.int32)]
    out = O.sparse_propagation_loops(h0, adj, np.ones((2, 1)), w, p)
    u = 1 / (1 + np.exp(-1.0))
    np.testing.assert_allclose(out, u * h0 + (1 - u) * np.tanh(h0[::-1]), rtol=1e-14)


def test_hand_derived_three_node_rnn_mean_bias():
    """path 0-1-2, RNN/ReLU, W = 2I, bias b, mean aggregation: node 1 receives (2h0+2h2+2b)/(2+1e-7)."""
    D = 2
    p = {"hidden_size": D, "layer_timesteps": [1], "residual_connections": {}, "use_edge_bias": True,
         "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "RNN", "graph_rnn_activation": "relu"}
    k = np.concatenate([np.eye(D), -np.eye(D)])
    bias = np.array([[0.25, -0.5]])
    w = [{"edge_weights": np.stack([2 * np.eye(D)]), "edge_biases": bias, "rnn_kernel": k, "rnn_bias": np.zeros(D)}]
    h0 = np.array([[1.0, 2.0], [0.5, 0.5], [3.0, -1.0]])
    adj = [np.array([[0, 1], [1, 0], [1, 2], [2, 1]], np.int32)]
    indeg = np.array([[1.0], [2.0], [1.0]])
    out = O.sparse_propagation_loops(h0, adj, indeg, w, p)
    inc = np.stack([(2 * h0[1] + bias[0]) / (1 + 1e-7), (2 * h0[0] + 2 * h0[2] + 2 * bias[0]) / (2 + 1e-7),
                    (2 * h0[1] + bias[0]) / (1 + 1e-7)])
    np.testing.assert_allclose(out, np.maximum(inc - h0, 0), rtol=1e-14)


def test_invariants_permutation_isolated_empty_type_layer_split():
    p = dict(PARAM_SETS["bias_noavg"], use_edge_msg_avg_aggregation=True)
    D = p["hidden_size"]
    _, b = _batch(D, n=4, seed=8)
    w = O.init_sparse_weights(p, 4, np.random.default_rng(2))
    h0, adj, indeg = b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"]
    ref = O.sparse_propagation_loops(h0, adj, indeg, w, p)
    V = h0.shape[0]
    # permutation equivariance
    perm = np.random.default_rng(0).permutation(V); inv = np.argsort(perm)
    adj_p = [inv[a].astype(np.int32) if a.size else a for a in adj]
    out_p = O.sparse_propagation_loops(h0[perm], adj_p, indeg[perm], w, p)
    np.testing.assert_allclose(out_p, ref[perm], rtol=1e-12, atol=1e-13)
    # an isolated node (deg 0 -> 0/1e-7 = 0 incoming) leaves the others unchanged
    h_iso = np.concatenate([h0, np.full((1, D), 0.3, np.float32)])
    out_iso = O.sparse_propagation_loops(h_iso, adj, np.concatenate([indeg, np.zeros((1, 4), np.float32)]), w, p)
    np.testing.assert_allclose(out_iso[:V], ref, rtol=1e-13)
    assert np.all(np.isfinite(out_iso[V]))
    # an empty extra edge type is a no-op
    w5 = [dict(x, edge_weights=np.concatenate([x["edge_weights"], np.ones((1, D, D), np.float32)]),
               edge_biases=np.concatenate([x["edge_biases"], np.ones((1, D), np.float32)])) for x in w]
    out5 = O.sparse_propagation_loops(h0, adj + [np.zeros((0, 2), np.int32)],
                                      np.concatenate([indeg, np.zeros((V, 1), np.float32)], 1), w5, p)
    np.testing.assert_allclose(out5, ref, rtol=1e-13)
    # layer_timesteps [1,2] with copied weights == [3]
    out_split = O.sparse_propagation_loops(h0, adj, indeg, [w[0], w[0]], dict(p, layer_timesteps=[1, 2]))
    np.testing.assert_allclose(out_split, ref, rtol=1e-13)


def test_out_of_range_edge_raises():
    p = PARAM_SETS["bias_noavg"]
    w = O.init_sparse_weights(p, 1, np.random.default_rng(2))
    with pytest.raises(IndexError):
        O.sparse_propagation_np(np.zeros((2, 9), np.float32), [np.array([[0, 2]], np.int32)], np.zeros((2, 1)), w, p)
    with pytest.raises(IndexError):
        CO.sparse_propagation_c(np.zeros((2, 9), np.float32), [np.array([[0, 2]], np.int32)], np.zeros((2, 1)), w, p)


def test_stable_target_csr_is_message_order_within_target():
    _, b = _batch(6, n=5, seed=1)
    V = b["initial_node_representation"].shape[0]
    row_ptr, src, typ, order = O.stable_target_csr(b["adjacency_lists"], V)
    s, t, ty = O.message_arrays(b["adjacency_lists"])
    assert row_ptr[-1] == len(t) and np.all(np.diff(row_ptr) == b["num_incoming_edges_per_type"].sum(1))
    for v in range(V):
        seg = order[row_ptr[v]:row_ptr[v + 1]]
        assert np.all(t[seg] == v) and np.all(np.diff(seg) > 0)
    np.testing.assert_array_equal(src, s[order]); np.testing.assert_array_equal(typ, ty[order])


@pytest.mark.parametrize("name", ["gru_bias_avg_res", "gru_plain", "rgcn_relu"])
def test_golden_sparse_regression(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "prop_sparse_%s.npz" % name))
    p = json.loads(str(z["params_json"]))
    L = len(p["layer_timesteps"])
    w = [{k[len("w%d_" % li):]: z[k] for k in z.files if k.startswith("w%d_" % li)} for li in range(L)]
    adj = [z["adj%d" % e] for e in range(4)]
    states = O.sparse_propagation_loops(z["h0"], adj, z["indeg"], w, p, return_all_layers=True)
    for li, s in enumerate(states):
        np.testing.assert_allclose(s, z["state%d" % li], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(CO.sparse_propagation_c(z["h0"], adj, z["indeg"], w, p), z["final"], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(O.sparse_propagation_torch(z["h0"], adj, z["indeg"], w, p).numpy(), z["final"],
                               rtol=2e-5, atol=2e-6)


def test_golden_dense_regression(golden_dir):
    z = np.load(os.path.join(golden_dir, "prop_dense.npz"))
    p = json.loads(str(z["params_json"]))
    w = {k[2:]: z[k] for k in z.files if k.startswith("w_")}
    np.testing.assert_allclose(O.dense_propagation_loops(z["h0"], z["adj"], w, p), z["final"], rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(CO.dense_propagation_c(z["h0"], z["adj"], w, p), z["final"], rtol=1e-12, atol=1e-13)


def test_propagation_attention_three_statements_agree():
    """sparse:170-196 restated message by message (loops), vectorised (NumPy) and at TF op granularity (torch) -- one result;
    with all sources of a node equal the softmax is uniform and attention reduces to dividing by the in-degree."""
    import torch
    p = {"hidden_size": 12, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]}, "use_edge_bias": True,
         "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh", "use_propagation_attention": True}
    _, b = U.molecule_batch(6, 12, T=4, seed=3)
    w = O.init_sparse_weights(p, 4, np.random.default_rng(1), attention_scale=0.5)
    args = (b["initial_node_representation"] * 3, b["adjacency_lists"], b["num_incoming_edges_per_type"], w, p)
    a = O.sparse_propagation_loops(*args)
    np.testing.assert_allclose(O.sparse_propagation_np(*args, dtype=np.float64), a, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(O.sparse_propagation_torch(*args, dtype=torch.float64).numpy(), a, rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(CO.sparse_propagation_c(*args), a, rtol=1e-12, atol=1e-13)
    # uniform case: identical node states -> every message into a node has the same score
    p1 = dict(p, layer_timesteps=[1], residual_connections={}, use_edge_bias=False, use_edge_msg_avg_aggregation=False)
    w1 = O.init_sparse_weights(p1, 4, np.random.default_rng(2))
    h_same = np.tile(np.random.default_rng(3).normal(size=(1, 12)), (b["initial_node_representation"].shape[0], 1))
    att = O.sparse_propagation_np(h_same, args[1], args[2], w1, p1, dtype=np.float64)
    mean = O.sparse_propagation_np(h_same, args[1], args[2], w1, dict(p1, use_propagation_attention=False, use_edge_msg_avg_aggregation=True),
                                   dtype=np.float64)
    np.testing.assert_allclose(att, mean, rtol=1e-6, atol=1e-7)


REFGRAPH_SPARSE = ["true_default_shape", "rnn_relu_bias_sum", "attention_bias_avg", "cudnn_gru", "gru_relu_sum_nobias", "rnn_tanh_avg_two_residuals",
                   "attention_rnn_sum"]


def _load_refgraph_sparse(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "refgraph_sparse_%s.npz" % name))
    p = json.loads(str(z["params_json"]))
    w = [{k[len("w%d_" % l):]: z[k] for k in z.files if k.startswith("w%d_" % l)} for l in range(len(p["layer_timesteps"]))]
    return z, p, w, [z["adj%d" % e] for e in range(4)]


@pytest.mark.parametrize("name", REFGRAPH_SPARSE)
def test_oracle_reproduces_the_reference_graph_code_sparse(golden_dir, name):
    """refgraph_*.npz were computed by the reference's OWN prepare_specific_graph_model / compute_final_node_representations /
    gated_regression (imported unmodified, tf.* served by tests/golden/tf_shim.py in float64, batch from the reference's own packer):
    all three statements of the oracle and its readout reproduce them to rounding."""
    import torch
    z, p, w, adj = _load_refgraph_sparse(golden_dir, name)
    for got in (O.sparse_propagation_loops(z["h0"], adj, z["indeg"], w, p),
                O.sparse_propagation_np(z["h0"], adj, z["indeg"], w, p, dtype=np.float64),
                O.sparse_propagation_torch(z["h0"], adj, z["indeg"], w, p, dtype=torch.float64).numpy(),
                CO.sparse_propagation_c(z["h0"], adj, z["indeg"], w, p)):
        np.testing.assert_allclose(got, z["final"], rtol=1e-11, atol=1e-12)
    ro = O.gated_regression_torch(z["final"], z["h0"], z["ro_w_gate"], z["ro_b_gate"], z["ro_w_trans"], z["ro_b_trans"],
                                  graph_nodes_list=z["graph_nodes_list"], num_graphs=int(z["num_graphs"]), dtype=torch.float64).numpy()
    np.testing.assert_allclose(ro, z["readout"], rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("name", ["cfg2_shape", "cfg4_shape"])
def test_oracle_reproduces_the_reference_graph_code_at_baseline_widths(golden_dir, name):
    """The same at BASELINE widths: hidden 100 / [4] / 4 edge types (configs[1]) and hidden 256 / [2,2,2,2] + residual / 8 edge types
    (configs[3]), a few dozen molecules each, computed by the reference's unmodified graph code -- no oracle bridge between the D = 12
    fixtures and the widths the benchmarks run at."""
    import torch
    from tests import _util as U
    z, p, w, adj, T = U.load_refgraph_wide(golden_dir, name)
    h0, indeg = np.asarray(z["h0"], np.float64), np.asarray(z["indeg"], np.float64)
    for got in (O.sparse_propagation_np(h0, adj, indeg, w, p, dtype=np.float64),
                O.sparse_propagation_torch(h0, adj, indeg, w, p, dtype=torch.float64).numpy(),
                CO.sparse_propagation_c(h0, adj, indeg, w, p)):
        np.testing.assert_allclose(got, z["final"], rtol=1e-10, atol=1e-12)
    ro = O.gated_regression_torch(z["final"], h0, z["ro_w_gate"], z["ro_b_gate"], z["ro_w_trans"], z["ro_b_trans"],
                                  graph_nodes_list=z["graph_nodes_list"], num_graphs=int(z["num_graphs"]), dtype=torch.float64).numpy()
    np.testing.assert_allclose(ro, z["readout"], rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("fixture", ["refgraph_dense.npz", "refgraph_dense_cfg3_shape.npz"])   # hidden 12; BASELINE configs[2] width (hidden 100)
def test_oracle_reproduces_the_reference_graph_code_dense(golden_dir, fixture):
    import torch
    z = np.load(os.path.join(golden_dir, fixture))
    p = json.loads(str(z["params_json"]))
    w = {k[2:]: z[k] for k in z.files if k.startswith("w_")}
    for got in (O.dense_propagation_loops(z["h0"], z["adj"], w, p), O.dense_propagation_torch(z["h0"], z["adj"], w, p, dtype=torch.float64).numpy(),
                CO.dense_propagation_c(z["h0"], z["adj"], w, p)):
        np.testing.assert_allclose(got, z["final"], rtol=1e-10, atol=1e-12)
    ro = O.gated_regression_torch(z["final"], z["h0"], z["ro_w_gate"], z["ro_b_gate"], z["ro_w_trans"], z["ro_b_trans"],
                                  node_mask=z["node_mask"], dtype=torch.float64).numpy()
    np.testing.assert_allclose(ro, z["readout"], rtol=1e-10, atol=1e-12)
