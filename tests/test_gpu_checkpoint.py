"""GPU (the plugins need an engine): checkpoints keyed by the reference's TensorFlow variable names (chem_tensorflow.py:309-359) --
SURVEY 8(f4).  The names are restated from TF-1.3 conventions (no TF here to confirm them)."""
import pickle

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _sparse_model(tmp_path, **cfg):
    from gated_graph_neural_network_samples_b200 import synthetic
    from gated_graph_neural_network_samples_b200.chem_sparse import SparseGGNNChemModel
    mols = synthetic.make_molecules(48, seed=1)
    base = {"hidden_size": 16, "batch_size": 100000, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
            "edge_weight_dropout_keep_prob": 1.0, "learning_rate": 0.01, "num_epochs": 1, "use_edge_bias": True}
    base.update(cfg)
    return SparseGGNNChemModel({"--log_dir": str(tmp_path), "--train_data": mols[:32], "--valid_data": mols[32:], "--config": base})


def test_sparse_variable_names_and_shapes_follow_the_tf_graph(tmp_path):
    m = _sparse_model(tmp_path, use_propagation_attention=True)
    named = dict(m.trainable_variables())
    T, D = m.num_edge_types, 16
    expect = {
        "graph_model/gnn_layer_0/gnn_edge_weights_0:0": (T * D, D), "graph_model/gnn_layer_0/gnn_edge_biases_0:0": (T, D),
        "graph_model/gnn_layer_0/edge_type_attention_weights_0:0": (T,),
        "graph_model/gnn_layer_0/timestep_0/gru_cell/gates/kernel:0": (2 * D, 2 * D), "graph_model/gnn_layer_0/timestep_0/gru_cell/gates/bias:0": (2 * D,),
        "graph_model/gnn_layer_0/timestep_0/gru_cell/candidate/kernel:0": (2 * D, D), "graph_model/gnn_layer_0/timestep_0/gru_cell/candidate/bias:0": (D,),
        "graph_model/gnn_layer_1/gnn_edge_weights_1:0": (T * D, D),
        "graph_model/gnn_layer_1/timestep_0/gru_cell/gates/kernel:0": (3 * D, 2 * D),     # one residual input: Din = 2D
        "out_layer_task0/regression_gate/MLP_W_layer0:0": (2 * D, 1), "out_layer_task0/regression_gate/MLP_b_layer0:0": (1,),
        "out_layer_task0/regression/MLP_W_layer0:0": (D, 1), "out_layer_task0/regression/MLP_b_layer0:0": (1,),
    }
    for k, shp in expect.items():
        assert k in named, (k, sorted(named))
        assert tuple(named[k].shape) == shp, (k, tuple(named[k].shape))
    rnn = _sparse_model(tmp_path, graph_rnn_cell="RNN")
    assert "graph_model/gnn_layer_0/timestep_0/basic_rnn_cell/kernel:0" in dict(rnn.trainable_variables())


def test_save_restore_round_trip_including_adam_slots(tmp_path):
    a = _sparse_model(tmp_path)
    a.run_epoch("t0", a.train_data, True)
    a.run_epoch("t1", a.train_data, True)
    path = str(tmp_path / "ckpt.pickle")
    a.save_progress(path, 7, 3)
    saved = pickle.load(open(path, "rb"))
    assert "beta1_power:0" in saved["weights"] and "graph_model/gnn_layer_0/gnn_edge_weights_0/Adam_1:0" in saved["weights"]
    b = _sparse_model(tmp_path)
    assert b.restore_progress(path) == (7, 3)
    for (n, va), (_, vb) in zip(a.trainable_variables(), b.trainable_variables()):
        np.testing.assert_array_equal(va.detach().cpu().numpy(), vb.detach().cpu().numpy(), err_msg=n)
    # identical optimizer state: one more identical step keeps the two models identical (up to the tensor path's rounding noise)
    np.random.seed(5); a.run_epoch("t2", a.train_data, True)
    np.random.seed(5); b.run_epoch("t2", b.train_data, True)
    for (n, va), (_, vb) in zip(a.trainable_variables(), b.trainable_variables()):
        np.testing.assert_allclose(va.detach().cpu().numpy(), vb.detach().cpu().numpy(), rtol=2e-3, atol=2e-5, err_msg=n)


def test_reference_style_pickle_loads(tmp_path):
    """A pickle shaped like the reference writes it (TF names, TF shapes, extra optimizer entries) restores without remapping."""
    m = _sparse_model(tmp_path)
    rng = np.random.default_rng(0)
    weights = {n: rng.normal(size=tuple(v.shape)).astype(np.float32) for n, v in m.trainable_variables()}
    weights["some_other_graph/variable:0"] = np.zeros(3, np.float32)
    path = str(tmp_path / "ref.pickle")
    pickle.dump({"params": m.params, "weights": weights, "train_step": 11, "valid_step": 5}, open(path, "wb"))
    assert m.restore_progress(path) == (11, 5)
    for n, v in m.trainable_variables():
        np.testing.assert_array_equal(v.detach().cpu().numpy(), weights[n])
    loss = m.run_epoch("valid", m.valid_data, False)[0]
    assert np.isfinite(loss)


def test_dense_variable_names(tmp_path):
    from gated_graph_neural_network_samples_b200 import synthetic
    from gated_graph_neural_network_samples_b200.chem_dense import DenseGGNNChemModel
    mols = synthetic.make_molecules(24, seed=2)
    m = DenseGGNNChemModel({"--log_dir": str(tmp_path), "--train_data": mols[:16], "--valid_data": mols[16:],
                            "--config": {"hidden_size": 16, "batch_size": 8, "num_timesteps": 2, "num_epochs": 1}})
    names = dict(m.trainable_variables())
    for k in ("graph_model/Variable:0", "graph_model/Variable_1:0", "graph_model/gru_scope/gru_cell/gates/kernel:0",
              "graph_model/gru_scope/gru_cell/candidate/bias:0", "out_layer_task0/regression/MLP_W_layer0:0"):
        assert k in names, (k, sorted(names))
    assert tuple(names["graph_model/Variable_1:0"].shape) == (m.num_edge_types, 1, 16)
