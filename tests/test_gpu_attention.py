"""GPU: use_propagation_attention (sparse:94-96,147-149,170-196) -- the per-target softmax over incoming messages -- forward and
gradients through the C ABI against the float64 oracle."""
import numpy as np
import pytest

from oracle import ggnn_oracle as O
from tests import _util as U
from tests.test_gpu_backward import _autograd_reference, _cmp, _engine_grads

pytestmark = pytest.mark.gpu

ATT_CASES = {
    "gru_bias_avg_res": {"hidden_size": 20, "layer_timesteps": [2, 1, 2], "residual_connections": {"1": [0], "2": [0, 1]},
                         "use_edge_bias": True, "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh",
                         "use_propagation_attention": True},
    "gru_d100": {"hidden_size": 100, "layer_timesteps": [3], "residual_connections": {}, "use_edge_bias": False,
                 "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh", "use_propagation_attention": True},
    "rnn_relu": {"hidden_size": 32, "layer_timesteps": [1, 1, 1], "residual_connections": {}, "use_edge_bias": False,
                 "use_edge_msg_avg_aggregation": False, "graph_rnn_cell": "RNN", "graph_rnn_activation": "ReLU", "use_propagation_attention": True},
}


def _batch(D, T=4, n=24, seed=3):
    _, b = U.molecule_batch(n, D, T=T, seed=seed, noise=0.0)
    rng = np.random.default_rng(seed + 50)
    h0 = (b["initial_node_representation"] * 1.5 + rng.normal(0, 0.4, b["initial_node_representation"].shape)).astype(np.float32)
    return h0, b["adjacency_lists"], b["num_incoming_edges_per_type"]


@pytest.mark.parametrize("force_global", ["0", "1"])
@pytest.mark.parametrize("name", sorted(ATT_CASES))
def test_attention_forward_matches_oracle(name, force_global, monkeypatch):
    monkeypatch.setenv("GGNN_FORCE_GLOBAL", force_global)
    p = ATT_CASES[name]
    D, T = p["hidden_size"], 4
    h0, adj, indeg = _batch(D)
    w = O.init_sparse_weights(p, T, np.random.default_rng(1), attention_scale=0.6)
    ref = O.sparse_propagation_np(h0, adj, indeg, w, p, dtype=np.float64)
    plain = O.sparse_propagation_np(h0, adj, indeg, w, dict(p, use_propagation_attention=False), dtype=np.float64)
    assert np.max(np.abs(ref - plain)) > 1e-2          # the attention branch really changes the result on this input
    # the tensor-core precisions are accepted and served by the fp32 kernel (the plan says so)
    got, eng = U.engine_sparse(p, T, w, adj, indeg, h0, precision="bf16x3", return_engine=True)
    assert "fp32-ffma+attention" in eng.plan and ("GLOBAL" in eng.plan) == (force_global == "1"), eng.plan
    err = U.max_rel_err(got, ref)
    print("attention %-18s global=%s max rel err %.2e" % (name, force_global, err))
    assert err < 1e-4


@pytest.mark.parametrize("name", sorted(ATT_CASES))
def test_attention_gradients_match_float64_autograd(name):
    p = ATT_CASES[name]
    D, T = p["hidden_size"], 4
    h0, adj, indeg = _batch(D)
    w = O.init_sparse_weights(p, T, np.random.default_rng(1), attention_scale=0.6)
    G = np.random.default_rng(5).normal(size=h0.shape).astype(np.float32)
    ref_out, ref_dh0, ref_gw = _autograd_reference(p, T, w, adj, indeg, h0, G)
    out, dh0, gw = _engine_grads(p, T, w, lambda e: e.set_graph_sparse(adj, indeg), h0, G, "fp32")
    _cmp(out, ref_out, "forward")
    _cmp(dh0, ref_dh0, "d h0")
    for l, (a, r) in enumerate(zip(gw, ref_gw)):
        for k in r:
            _cmp(a[k], r[k], "layer %d %s" % (l, k))
        assert float(np.max(np.abs(r["edge_type_attention_weights"]))) > 0


def test_attention_isolated_nodes_and_single_message():
    """A node without incoming messages keeps incoming = 0; a node with exactly one message gets weight exp(0)/(1+1e-7)."""
    p = dict(ATT_CASES["gru_d100"], hidden_size=8, layer_timesteps=[2])
    T, V = 2, 5
    adj = [np.array([[0, 1], [2, 1], [3, 1]], np.int32), np.array([[1, 0]], np.int32)]
    indeg = np.zeros((V, T), np.float32)
    for t, a in enumerate(adj):
        for s, d in a:
            indeg[d, t] += 1
    h0 = np.random.default_rng(0).normal(0, 1, (V, 8)).astype(np.float32)
    w = O.init_sparse_weights(p, T, np.random.default_rng(1), attention_scale=0.3)
    ref = O.sparse_propagation_loops(h0, adj, indeg, w, p)
    got = U.engine_sparse(p, T, w, adj, indeg, h0)
    assert U.max_rel_err(got, ref) < 1e-4


def test_chem_model_trains_with_propagation_attention(tmp_path):
    from gated_graph_neural_network_samples_b200 import synthetic
    from gated_graph_neural_network_samples_b200.chem_sparse import SparseGGNNChemModel
    mols = synthetic.make_molecules(96, seed=1)
    args = {"--log_dir": str(tmp_path), "--train_data": mols[:64], "--valid_data": mols[64:],
            "--config": {"hidden_size": 32, "batch_size": 400, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                         "edge_weight_dropout_keep_prob": 1.0, "use_propagation_attention": True, "learning_rate": 0.01, "num_epochs": 1}}
    model = SparseGGNNChemModel(args)
    att0 = [a.detach().cpu().numpy().copy() for a in model.gnn_weights.edge_type_attention_weights]
    l0 = model.run_epoch("valid0", model.valid_data, False)[0]
    for ep in range(6):
        model.run_epoch("train%d" % ep, model.train_data, True)
    l1 = model.run_epoch("valid1", model.valid_data, False)[0]
    print("validation loss %.4f -> %.4f" % (l0, l1))
    assert np.isfinite(l1) and l1 < l0
    assert any(np.max(np.abs(a.detach().cpu().numpy() - b)) > 0 for a, b in zip(model.gnn_weights.edge_type_attention_weights, att0))
