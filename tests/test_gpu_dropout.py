"""GPU: DropoutWrapper(state_keep_prob) of sparse:113-114,216 / dense:89 inside the kernels -- forward and gradients against
the float64 oracle run with the same (restated) counter-based mask; TensorFlow's own random stream cannot be reproduced."""
import numpy as np
import pytest

from oracle import ggnn_oracle as O
from tests import _util as U
from tests.test_gpu_backward import CASES, _autograd_reference, _cmp, _engine_grads

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
@pytest.mark.parametrize("name", sorted(CASES))
def test_state_dropout_forward_and_gradients(name, precision):
    p = CASES[name]
    D, T = p["hidden_size"], 4
    drop = (0.75, 123456789)
    _, b = U.molecule_batch(24, D, T=T, seed=3)
    w = O.init_sparse_weights(p, T, np.random.default_rng(1))
    h0, adj, indeg = b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"]
    G = np.random.default_rng(5).normal(size=h0.shape).astype(np.float32)
    ref_out, ref_dh0, ref_gw = _autograd_reference(p, T, w, adj, indeg, h0, G, state_dropout=drop)
    out, dh0, gw = _engine_grads(p, T, w, lambda e: e.set_graph_sparse(adj, indeg), h0, G, precision, state_dropout=drop)
    if p["graph_rnn_activation"].lower() == "tanh":         # (ReLU has zeros of its own)
        frac_zero = float(np.mean(out == 0.0))
        assert 0.15 < frac_zero < 0.35, frac_zero            # the last timestep's mask is visible in the result
        np.testing.assert_array_equal(out == 0.0, ref_out == 0.0)
    _cmp(out, ref_out, "forward")
    _cmp(dh0, ref_dh0, "d h0")
    for l, (a, r) in enumerate(zip(gw, ref_gw)):
        for k in r:
            _cmp(a[k], r[k], "layer %d %s" % (l, k))


@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_state_dropout_global_mode_and_keep_one(precision, monkeypatch):
    """One launch per timestep (GLOBAL plan) draws the same per-step masks as the fused LOCAL plan; keep=1 is the identity."""
    import torch
    from gated_graph_neural_network_samples_b200.engine import GgnnError, PropagationEngine
    p = {"hidden_size": 40, "layer_timesteps": [2, 2], "residual_connections": {"1": [0]}, "use_edge_bias": True,
         "use_edge_msg_avg_aggregation": True, "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}
    T = 4
    _, b = U.molecule_batch(20, 40, T=T, seed=9)
    w = O.init_sparse_weights(p, T, np.random.default_rng(4))
    h0, adj, indeg = b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"]
    ref = O.sparse_propagation_torch(h0, adj, indeg, w, p, dtype=torch.float64, state_dropout=(0.9, 7)).numpy()
    outs = {}
    for mode in ("local", "global"):
        if mode == "global":
            monkeypatch.setenv("GGNN_FORCE_GLOBAL", "1")
        eng = PropagationEngine(p, T, precision=precision)
        eng.set_weights(U.to_cuda_weights(w))
        eng.set_state_dropout(0.9, 7)
        eng.set_graph_sparse(adj, indeg)
        th0 = torch.from_numpy(h0).cuda()
        outs[mode] = eng.forward(th0).cpu().numpy()
        assert ("GLOBAL" in eng.plan) == (mode == "global"), eng.plan
        _cmp(outs[mode], ref, "dropout forward " + mode)
        # the host restatement of the mask is the one the kernel used (last global step = 3)
        np.testing.assert_array_equal(outs[mode] != 0.0, eng.state_dropout_mask(3, 0.9, 7).astype(bool) & (ref != 0.0))
        eng.set_state_dropout(1.0)
        plain = eng.forward(th0).cpu().numpy()
        _cmp(plain, O.sparse_propagation_torch(h0, adj, indeg, w, p, dtype=torch.float64).numpy(), "keep=1 " + mode)
        with pytest.raises(GgnnError):
            eng.set_state_dropout(0.0)


def test_chem_model_trains_with_state_and_weight_dropout(tmp_path):
    """The sparse file's own defaults use edge-weight dropout 0.8 (sparse:57); with state dropout on top the training loop
    (ChemModel.run_epoch -> hooks -> engine forward/backward) still learns and evaluation (keep=1) stays deterministic."""
    from gated_graph_neural_network_samples_b200 import synthetic
    from gated_graph_neural_network_samples_b200.chem_sparse import SparseGGNNChemModel
    mols = synthetic.make_molecules(96, seed=1)
    args = {"--log_dir": str(tmp_path), "--train_data": mols[:64], "--valid_data": mols[64:],
            "--config": {"hidden_size": 32, "batch_size": 400, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                         "edge_weight_dropout_keep_prob": 0.8, "graph_state_dropout_keep_prob": 0.9,
                         "learning_rate": 0.01, "num_epochs": 1}}
    model = SparseGGNNChemModel(args)
    l0 = model.run_epoch("valid0", model.valid_data, False)[0]
    assert abs(l0 - model.run_epoch("valid0b", model.valid_data, False)[0]) < 1e-4 * max(1.0, abs(l0))   # no mask in evaluation
    for ep in range(8):
        model.run_epoch("train%d" % ep, model.train_data, True)
    l1 = model.run_epoch("valid1", model.valid_data, False)[0]
    print("validation loss %.4f -> %.4f" % (l0, l1))
    assert np.isfinite(l1) and l1 < l0
