"""GPU: ``graph_rnn_cell = CudnnCompatibleGRUCell`` (chem_tensorflow_sparse.py:105-108) through the C ABI.

tf.contrib.cudnn_rnn.CudnnCompatibleGRUCell applies the reset gate AFTER the recurrent product:
``c = tanh(x.K_in + b_in + r*(h.K_hid + b_hid))``.  The engine serves it on the fp32 kernel (whatever precision is requested, like
propagation attention); forward and every gradient are held to the float64 oracle, the forward also to the fixture the reference's own
graph code computed with this cell (tests/golden/refgraph_sparse_cudnn_gru.npz)."""
import json
import os
import pickle

import numpy as np
import pytest

from gated_graph_neural_network_samples_b200 import packing, synthetic
from oracle import ggnn_oracle as O
from tests import _util as U
from tests.test_gpu_backward import _autograd_reference, _cmp, _engine_grads

pytestmark = pytest.mark.gpu

CELL = "CudnnCompatibleGRUCell"
CASES = {
    # BASELINE configs[1] width, a residual input (Din = 2D in layer 1), edge bias + mean aggregation
    "d100_res_bias_avg": ({"hidden_size": 100, "layer_timesteps": [2, 2], "residual_connections": {"1": [0]}, "use_edge_bias": True,
                           "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": CELL, "graph_rnn_activation": "tanh"}, 4, 64),
    # the reference's default layer structure
    "default_shape": ({"hidden_size": 64, "layer_timesteps": [2, 2, 1, 2, 1], "residual_connections": {"2": [0], "4": [0, 2]},
                       "use_edge_bias": False, "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": CELL, "graph_rnn_activation": "tanh"}, 4, 48),
    # BASELINE configs[3] width / edge-type count
    "d256_t8": ({"hidden_size": 256, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]}, "use_edge_bias": False,
                 "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": CELL, "graph_rnn_activation": "tanh"}, 8, 40),
    "d20_sum": ({"hidden_size": 20, "layer_timesteps": [3], "residual_connections": {}, "use_edge_bias": True,
                 "use_edge_msg_avg_aggregation": False, "graph_rnn_cell": CELL, "graph_rnn_activation": "tanh"}, 4, 30),
}


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_forward_matches_float64_oracle(name, precision):
    p, T, n = CASES[name]
    _, b = U.molecule_batch(n, p["hidden_size"], T=T, seed=11)
    w = O.init_sparse_weights(p, T, np.random.default_rng(2))
    h0, adj, indeg = b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"]
    ref = O.sparse_propagation_np(h0, adj, indeg, w, p, dtype=np.float64, return_all_layers=True)
    got, eng = U.engine_sparse(p, T, w, adj, indeg, h0, precision=precision, return_engine=True)
    assert "cudnn-gru" in eng.plan and "fp32" in eng.plan, eng.plan     # served by the fp32 kernel whatever was asked for
    err = U.max_rel_err(got, ref[-1])
    print("cudnn-gru %-18s %-6s max rel err %.2e  [%s]" % (name, precision, err, eng.plan[:60]))
    assert np.all(np.isfinite(got)) and err < 1e-4
    for l in range(1, len(p["layer_timesteps"])):                       # every node_states_per_layer entry (residual sources)
        assert U.max_rel_err(eng.layer_state(l).cpu().numpy(), ref[l]) < 1e-4


def test_matches_the_reference_graph_code_fixture(golden_dir):
    import torch
    z = np.load(os.path.join(golden_dir, "refgraph_sparse_cudnn_gru.npz"))
    p = json.loads(str(z["params_json"]))
    w = [{k[len("w%d_" % l):]: z[k] for k in z.files if k.startswith("w%d_" % l)} for l in range(len(p["layer_timesteps"]))]
    adj = [z["adj%d" % e] for e in range(4)]
    got, eng = U.engine_sparse(p, 4, w, adj, z["indeg"].astype(np.float32), z["h0"].astype(np.float32), precision="bf16x3", return_engine=True)
    err = U.max_rel_err(got, z["final"])
    print("refgraph cudnn_gru propagation max rel err %.2e" % err)
    assert err < 1e-4
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    eng.readout_set_graphs(int(z["num_graphs"]), graph_nodes_list=z["graph_nodes_list"])
    ro = eng.readout_forward(f32(got), f32(z["h0"]), f32(z["ro_w_gate"]), f32(z["ro_b_gate"]), f32(z["ro_w_trans"]), f32(z["ro_b_trans"]))
    eng.sync_check()
    assert U.max_rel_err(ro.cpu().numpy(), z["readout"]) < 1e-4


def test_one_large_graph_runs_one_launch_per_step():
    """A component larger than a tile (GLOBAL plan of the fp32 kernel: one launch per timestep, states through L2)."""
    p = {"hidden_size": 64, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]}, "use_edge_bias": True,
         "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": CELL, "graph_rnn_activation": "tanh"}
    T, V = 3, 3000
    rng = np.random.default_rng(7)
    adj = []
    for t in range(T):
        src = rng.integers(0, V, 4000)
        tgt = (src + rng.integers(1, 50, 4000)) % V
        adj.append(np.stack([src, tgt], 1).astype(np.int32))
    adj[0] = np.concatenate([adj[0], np.stack([np.arange(V - 1), np.arange(1, V)], 1).astype(np.int32)])   # a path: one component
    indeg = np.zeros((V, T), np.float32)
    for t in range(T):
        np.add.at(indeg[:, t], adj[t][:, 1], 1.0)
    h0 = rng.normal(0, 0.5, (V, 64)).astype(np.float32)
    w = O.init_sparse_weights(p, T, np.random.default_rng(3))
    ref = O.sparse_propagation_np(h0, adj, indeg, w, p, dtype=np.float64)
    got, eng = U.engine_sparse(p, T, w, adj, indeg, h0, precision="fp32", return_engine=True)
    assert "GLOBAL" in eng.plan and "cudnn-gru" in eng.plan, eng.plan
    assert U.max_rel_err(got, ref) < 1e-4


@pytest.mark.parametrize("name", ["d100_res_bias_avg", "default_shape", "d20_sum"])
def test_gradients_match_float64_autograd(name):
    p, T, _ = CASES[name]
    _, b = U.molecule_batch(24, p["hidden_size"], T=T, seed=3)
    w = O.init_sparse_weights(p, T, np.random.default_rng(1))
    h0, adj, indeg = b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"]
    G = np.random.default_rng(5).normal(size=h0.shape).astype(np.float32)
    ref_out, ref_dh0, ref_gw = _autograd_reference(p, T, w, adj, indeg, h0, G)
    out, dh0, gw = _engine_grads(p, T, w, lambda e: e.set_graph_sparse(adj, indeg), h0, G, "bf16x3")
    _cmp(out, ref_out, "forward")
    _cmp(dh0, ref_dh0, "d h0")
    for l, (a, r) in enumerate(zip(gw, ref_gw)):
        assert set(r) <= set(a) and "cand_hidden_bias" in r
        for k in r:
            _cmp(a[k], r[k], "layer %d %s" % (l, k))


def test_gradients_with_state_dropout():
    p, T, _ = CASES["d20_sum"]
    _, b = U.molecule_batch(16, p["hidden_size"], T=T, seed=4)
    w = O.init_sparse_weights(p, T, np.random.default_rng(1))
    h0, adj, indeg = b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"]
    G = np.random.default_rng(6).normal(size=h0.shape).astype(np.float32)
    drop = (0.8, 1234)
    ref_out, ref_dh0, ref_gw = _autograd_reference(p, T, w, adj, indeg, h0, G, state_dropout=drop)
    out, dh0, gw = _engine_grads(p, T, w, lambda e: e.set_graph_sparse(adj, indeg), h0, G, "fp32", state_dropout=drop)
    _cmp(out, ref_out, "forward (dropout)")
    _cmp(dh0, ref_dh0, "d h0 (dropout)")
    for l, (a, r) in enumerate(zip(gw, ref_gw)):
        for k in r:
            _cmp(a[k], r[k], "dropout layer %d %s" % (l, k))


def test_relu_is_refused_like_the_reference_assert():
    from gated_graph_neural_network_samples_b200.engine import PropagationEngine
    p = dict(CASES["d20_sum"][0], graph_rnn_activation="ReLU")
    with pytest.raises(AssertionError):                                 # sparse:106: assert(activation_name == 'tanh')
        PropagationEngine(p, 4)


def test_chem_model_trains_and_checkpoints_with_the_cells_variable_names(tmp_path):
    from gated_graph_neural_network_samples_b200.chem_sparse import SparseGGNNChemModel
    mols = synthetic.make_molecules(96, seed=1)

    def model():
        return SparseGGNNChemModel({"--log_dir": str(tmp_path), "--train_data": mols[:64], "--valid_data": mols[64:],
                                    "--config": {"hidden_size": 32, "batch_size": 400, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                                                 "edge_weight_dropout_keep_prob": 1.0, "learning_rate": 0.01, "num_epochs": 1,
                                                 "graph_rnn_cell": CELL}})
    m = model()
    named = dict(m.trainable_variables())
    c = "graph_model/gnn_layer_1/timestep_0/cudnn_compatible_gru_cell/"
    assert tuple(named[c + "candidate/input_projection/kernel:0"].shape) == (64, 32)      # Din = 2D (one residual input)
    assert tuple(named[c + "candidate/hidden_projection/kernel:0"].shape) == (32, 32)
    assert tuple(named[c + "candidate/hidden_projection/bias:0"].shape) == (32,) and tuple(named[c + "gates/kernel:0"].shape) == (96, 64)
    l0 = m.run_epoch("valid0", m.valid_data, False)[0]
    for ep in range(6):
        m.run_epoch("train%d" % ep, m.train_data, True)
    l1 = m.run_epoch("valid1", m.valid_data, False)[0]
    print("cudnn-gru validation loss %.4f -> %.4f" % (l0, l1))
    assert np.isfinite(l1) and l1 < l0
    path = str(tmp_path / "ckpt.pickle")
    m.save_progress(path, 2, 1)
    assert c + "candidate/hidden_projection/kernel/Adam_1:0" in pickle.load(open(path, "rb"))["weights"]
    m2 = model()
    assert m2.restore_progress(path) == (2, 1)
    assert abs(m2.run_epoch("valid2", m2.valid_data, False)[0] - l1) < 1e-4 * max(1.0, abs(l1))
