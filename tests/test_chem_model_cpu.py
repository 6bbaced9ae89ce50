"""CPU: the host-side logic of the ChemModel mirror and its two plug-ins (data loading, the flattened packer, feed dicts, readout
fallback, loss, per-variable clipping + Adam, checkpoints keyed by TF variable names) driven end to end with a STAND-IN engine that
answers ``compute_final_node_representations`` from the oracle.  The real engine has no CPU path (see test_abi_cpu.py); this file only
makes sure the Python around it is exercised without a GPU."""
import pickle

import numpy as np
import pytest

from gated_graph_neural_network_samples_b200 import chem_dense, chem_sparse, synthetic
from oracle import ggnn_oracle as O


class StandInEngine:
    def __init__(self, params, num_edge_types, device=0, precision="fp32"):
        self.params, self.T, self.D = params, num_edge_types, int(params["hidden_size"])
        self.drop = None

    def set_save_for_backward(self, enable):
        pass

    def set_graph_sparse(self, adjacency_lists, indeg):
        self.adj, self.indeg = adjacency_lists, indeg

    def set_graph_dense(self, adjacency_matrix):
        self.adjm = np.asarray(adjacency_matrix)

    def set_state_dropout(self, keep, seed=0):
        self.drop = (keep, seed) if keep < 1.0 else None


class StandInPropagation:
    @staticmethod
    def apply(engine, layout, h0, *flat):
        weights = [{k: flat[i] for k, i in lay.items()} for lay in layout]
        if hasattr(engine, "adjm"):      # dense plug-in: one layer, [b*v, D] states
            b, _, v, _ = engine.adjm.shape
            w = dict(weights[0])
            p = {"num_timesteps": engine.params["layer_timesteps"][0], "use_edge_bias": "edge_biases" in w}
            return O.dense_propagation_torch(h0.reshape(b, v, engine.D), engine.adjm, w, p).reshape(b * v, engine.D)
        if engine.params.get("graph_rnn_cell", "GRU").lower() == "rnn":
            weights = [{{"cand_kernel": "rnn_kernel", "cand_bias": "rnn_bias"}.get(k, k): v for k, v in w.items()} for w in weights]
        return O.sparse_propagation_torch(h0, engine.adj, engine.indeg, weights, engine.params, state_dropout=engine.drop)


@pytest.fixture
def stand_in(monkeypatch):
    monkeypatch.setattr(chem_sparse, "PropagationEngine", StandInEngine)
    monkeypatch.setattr(chem_sparse, "_propagation_function", lambda: StandInPropagation)
    monkeypatch.setattr(chem_dense, "PropagationEngine", StandInEngine)
    monkeypatch.setattr(chem_dense, "_propagation_function", lambda: StandInPropagation)


def _args(tmp_path, mols, **cfg):
    base = {"hidden_size": 16, "batch_size": 300, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
            "edge_weight_dropout_keep_prob": 0.9, "graph_state_dropout_keep_prob": 0.9, "learning_rate": 0.01, "num_epochs": 2,
            "use_edge_bias": True}
    base.update(cfg)
    return {"--log_dir": str(tmp_path), "--device": "cpu", "--train_data": mols[:48], "--valid_data": mols[48:], "--config": base}


@pytest.mark.parametrize("cfg", [{}, {"use_propagation_attention": True}, {"graph_rnn_cell": "RNN", "graph_rnn_activation": "ReLU"},
                                 {"graph_rnn_cell": "CudnnCompatibleGRUCell"}])
def test_sparse_model_trains_saves_and_restores_on_the_host(tmp_path, stand_in, cfg):
    mols = synthetic.make_molecules(64, seed=1)
    m = chem_sparse.SparseGGNNChemModel(_args(tmp_path, mols, **cfg))
    l0 = m.run_epoch("valid0", m.valid_data, False)[0]
    for ep in range(5):
        train_loss, accs, errs, speed, steps = m.run_epoch("train%d" % ep, m.train_data, True)
        assert steps >= 3 and np.isfinite(train_loss)          # 48 molecules at <300 nodes per batch: several minibatches per epoch
    l1 = m.run_epoch("valid1", m.valid_data, False)[0]
    assert np.isfinite(l1) and l1 < l0
    path = str(tmp_path / "ckpt.pickle")
    m.save_progress(path, 3, 1)
    saved = pickle.load(open(path, "rb"))["weights"]
    assert "graph_model/gnn_layer_0/gnn_edge_weights_0:0" in saved and "out_layer_task0/regression/MLP_W_layer0:0" in saved
    assert "beta1_power:0" in saved and any(k.endswith("/Adam_1:0") for k in saved)
    if cfg.get("graph_rnn_cell") == "CudnnCompatibleGRUCell":   # sparse:105-108: the cell's own variable scopes, TF's shapes (Din = 2D in layer 1)
        c = "graph_model/gnn_layer_1/timestep_0/cudnn_compatible_gru_cell/"
        assert saved[c + "candidate/input_projection/kernel:0"].shape == (32, 16) and saved[c + "candidate/hidden_projection/kernel:0"].shape == (16, 16)
        assert saved[c + "candidate/hidden_projection/bias:0"].shape == (16,) and saved[c + "gates/kernel:0"].shape == (48, 32)
        assert np.abs(saved[c + "candidate/hidden_projection/bias:0"]).max() > 0          # it is trained
    m2 = chem_sparse.SparseGGNNChemModel(_args(tmp_path, mols, **cfg))
    assert m2.restore_progress(path) == (3, 1)
    for (n, a), (_, b) in zip(m.trainable_variables(), m2.trainable_variables()):
        np.testing.assert_array_equal(a.detach().numpy(), b.detach().numpy(), err_msg=n)
    assert abs(m2.run_epoch("valid2", m2.valid_data, False)[0] - l1) < 1e-5 * max(1.0, abs(l1))
    m.train()                                                   # the epoch loop with early stopping and best-model checkpointing
    assert pickle.load(open(m.best_model_file, "rb"))["params"]["hidden_size"] == 16


def test_dense_model_trains_on_the_host(tmp_path, stand_in):
    mols = synthetic.make_molecules(64, seed=2)
    args = {"--log_dir": str(tmp_path), "--device": "cpu", "--train_data": mols[:48], "--valid_data": mols[48:],
            "--config": {"hidden_size": 16, "batch_size": 4, "num_timesteps": 2, "learning_rate": 0.01, "num_epochs": 1}}
    m = chem_dense.DenseGGNNChemModel(args)
    l0 = m.run_epoch("valid0", m.valid_data, False)[0]
    for ep in range(4):
        m.run_epoch("train%d" % ep, m.train_data, True)
    l1 = m.run_epoch("valid1", m.valid_data, False)[0]
    assert np.isfinite(l1) and l1 < l0
    assert "graph_model/gru_scope/gru_cell/gates/kernel:0" in dict(m.trainable_variables())


def test_dense_task_sample_ratios_and_state_dropout_in_training(tmp_path, stand_in):
    """dense:153-158 (labels beyond the ratio are dropped per shuffled bucket) and dense:222 (the weight-dropout slot is fed with the
    state keep probability): a dense training run with both set must run, and the ratio must show up in the target masks."""
    mols = synthetic.make_molecules(80, seed=5)
    args = {"--log_dir": str(tmp_path), "--device": "cpu", "--train_data": mols[:64], "--valid_data": mols[64:],
            "--config": {"hidden_size": 16, "batch_size": 4, "num_timesteps": 2, "learning_rate": 0.01, "num_epochs": 1,
                         "task_sample_ratios": {"0": 0.5}, "graph_state_dropout_keep_prob": 0.9}}
    m = chem_dense.DenseGGNNChemModel(args)
    labelled = total = 0
    for feed in m.make_minibatch_iterator(m.train_data, is_training=True):
        labelled += float(np.sum(feed["target_mask"]))
        total += feed["num_graphs"]
        assert feed["edge_weight_dropout_keep_prob"] == 0.9 and feed["graph_state_keep_prob"] == 0.9
    assert total > 0 and 0.3 * total <= labelled <= 0.7 * total, (labelled, total)
    for feed in m.make_minibatch_iterator(m.valid_data, is_training=False):
        assert np.all(feed["target_mask"] == 1.0)               # validation data keeps every label (dense:149: training only)
    loss = m.run_epoch("train", m.train_data, True)[0]
    assert np.isfinite(loss)


@pytest.mark.parametrize("steps", [3, 2000, 150000])
def test_adam_step_survives_a_checkpoint_after_many_updates(tmp_path, stand_in, steps):
    """float32(0.9 ** (t + 1)) underflows to 0 after ~1000 updates (log -> -inf): the count is stored explicitly, and a TensorFlow pickle
    that only has the beta powers falls back to beta2_power, then to train_step."""
    mols = synthetic.make_molecules(64, seed=1)
    m = chem_sparse.SparseGGNNChemModel(_args(tmp_path, mols))
    m.run_epoch("train", m.train_data, True)
    for st in m.optimizer.state.values():
        st["step"].fill_(float(steps))
    path = str(tmp_path / "late.pickle")
    m.save_progress(path, steps, 0)
    m2 = chem_sparse.SparseGGNNChemModel(_args(tmp_path, mols))
    m2.restore_progress(path)
    assert all(int(st["step"]) == steps for st in m2.optimizer.state.values())
    # a pickle written by TensorFlow has no 'adam_step'
    data = pickle.load(open(path, "rb"))
    del data["weights"]["adam_step"]
    pickle.dump(data, open(path, "wb"))
    m3 = chem_sparse.SparseGGNNChemModel(_args(tmp_path, mols))
    m3.restore_progress(path)
    got = {int(st["step"]) for st in m3.optimizer.state.values()}
    assert len(got) == 1
    assert abs(got.pop() - steps) <= max(1, steps // 50)      # beta2_power = 0.999^(t+1) in float32 resolves t to well under 2 %


def test_the_real_engine_still_refuses_the_cpu(tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(Exception, match="CUDA"):
        chem_sparse.SparseGGNNChemModel(_args(tmp_path, synthetic.make_molecules(64, seed=1)))


@pytest.mark.parametrize("name", ["true_default_shape", "rnn_relu_bias_sum", "attention_bias_avg"])
def test_forward_batch_loss_matches_the_reference_make_model(tmp_path, stand_in, name):
    """refgraph_sparse_*.npz hold loss and MAE computed by the reference's OWN make_model (chem_tensorflow.py:133-170: hooks, readout
    MLPs, gated_regression, masked loss), run unmodified over tests/golden/tf_shim.py.  With the fixture's weights loaded, the mirror's
    forward_batch on the fixture's feed returns the same numbers (propagation answered by the stand-in engine = the oracle, fp32)."""
    import json
    import os
    import torch
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "refgraph_sparse_%s.npz" % name))
    cfg = json.loads(str(z["params_json"]))
    mols = synthetic.make_molecules(8, seed=1)                    # only sets num_edge_types / annotation size; the feed comes from the fixture
    args = {"--log_dir": str(tmp_path), "--device": "cpu", "--train_data": mols[:4], "--valid_data": mols[4:],
            "--config": dict(cfg, batch_size=100000, edge_weight_dropout_keep_prob=1.0)}
    m = chem_sparse.SparseGGNNChemModel(args)
    T, D = 4, cfg["hidden_size"]
    ren = {"rnn_kernel": "cand_kernel", "rnn_bias": "cand_bias"}
    with torch.no_grad():
        for l in range(len(cfg["layer_timesteps"])):
            m.gnn_weights.edge_weights[l].copy_(torch.from_numpy(z["w%d_edge_weights" % l].reshape(T * D, D).astype(np.float32)))
            if cfg["use_edge_bias"]:
                m.gnn_weights.edge_biases[l].copy_(torch.from_numpy(z["w%d_edge_biases" % l].astype(np.float32)))
            if cfg.get("use_propagation_attention"):
                m.gnn_weights.edge_type_attention_weights[l].copy_(torch.from_numpy(z["w%d_edge_type_attention_weights" % l].astype(np.float32)))
            for k in ("gate_kernel", "gate_bias", "cand_kernel", "cand_bias", "rnn_kernel", "rnn_bias"):
                if "w%d_%s" % (l, k) in z.files:
                    m.gnn_weights.rnn_cells[l][ren.get(k, k)].copy_(torch.from_numpy(z["w%d_%s" % (l, k)].astype(np.float32)))
        gate, trans = m.weights["regression_gate_task0"], m.weights["regression_transform_task0"]
        gate.weights[0].copy_(torch.from_numpy(z["ro_w_gate"].astype(np.float32))); gate.biases[0].copy_(torch.from_numpy(z["ro_b_gate"].astype(np.float32)))
        trans.weights[0].copy_(torch.from_numpy(z["ro_w_trans"].astype(np.float32))); trans.biases[0].copy_(torch.from_numpy(z["ro_b_trans"].astype(np.float32)))
    feed = {"initial_node_representation": z["h0"].astype(np.float32), "num_incoming_edges_per_type": z["indeg"].astype(np.float32),
            "graph_nodes_list": z["graph_nodes_list"], "num_graphs": int(z["num_graphs"]), "target_values": z["target_values"],
            "target_mask": z["target_mask"], "graph_state_keep_prob": 1.0, "edge_weight_dropout_keep_prob": 1.0,
            "out_layer_dropout_keep_prob": 1.0}
    for e in range(T):
        feed["adjacency_e%d" % e] = z["adj%d" % e]
    assert float(z["target_mask"].sum()) == z["target_mask"].size - 1            # one unlabeled graph in the fixture batch
    with torch.no_grad():
        loss, accs = m.forward_batch(feed)
    np.testing.assert_allclose(m.ops["final_node_representations"].numpy(), z["final"], rtol=1e-4, atol=1e-5 * float(np.abs(z["final"]).max()))
    np.testing.assert_allclose(m.output.numpy(), z["readout"], rtol=1e-4, atol=1e-5 * float(np.abs(z["readout"]).max()))
    assert abs(float(loss) - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    assert abs(float(accs[0]) - float(z["accuracy"])) < 1e-4 * abs(float(z["accuracy"]))


def test_threaded_iterator_delivers_items_and_re_raises_producer_errors():
    from gated_graph_neural_network_samples_b200.utils import ThreadedIterator
    assert list(ThreadedIterator(iter(range(7)), max_queue_size=2)) == list(range(7))

    def failing():
        yield 1
        yield 2
        raise ValueError("graph does not fit")

    got = []
    with pytest.raises(ValueError, match="does not fit"):
        for x in ThreadedIterator(failing(), max_queue_size=5):
            got.append(x)
    assert got == [1, 2]


def test_a_graph_larger_than_the_node_budget_raises_instead_of_hanging(tmp_path, stand_in):
    mols = synthetic.make_molecules(64, seed=1)
    m = chem_sparse.SparseGGNNChemModel(_args(tmp_path, mols, batch_size=5))   # every molecule has more than 5 nodes (sparse:297 loops forever)
    with pytest.raises(Exception, match="does not fit"):
        m.run_epoch("valid", m.valid_data, False)


@pytest.mark.parametrize("cfg", [{"hidden_size": 10}, {"hidden_size": 7, "graph_rnn_cell": "RNN", "graph_rnn_activation": "ReLU", "use_edge_bias": True},
                                 {"hidden_size": 9, "graph_rnn_cell": "CudnnCompatibleGRUCell"}, {"hidden_size": 6, "use_propagation_attention": True}])
def test_hidden_sizes_that_are_not_multiples_of_4_run_zero_padded(tmp_path, stand_in, cfg):
    """The reference accepts any hidden size; the engine wants multiples of 4 (16-byte row vectors).  The plug-in zero-pads states and
    weights at the engine boundary and slices the result: padded units stay exactly 0 through every cell, so the real units see the same
    sums.  Checked here through the stand-in engine (which receives the PADDED problem) against the oracle on the UNPADDED one."""
    import torch
    mols = synthetic.make_molecules(64, seed=3)
    m = chem_sparse.SparseGGNNChemModel(_args(tmp_path, mols, edge_weight_dropout_keep_prob=1.0, graph_state_dropout_keep_prob=1.0, **cfg))
    D = cfg["hidden_size"]
    assert m._padded_hidden == (D + 3) // 4 * 4 != D and m.engine.D == m._padded_hidden
    feed = next(iter(m.make_minibatch_iterator(m.valid_data, False)))
    m.feed = feed
    with torch.no_grad():
        got = m.compute_final_node_representations().numpy()
    assert got.shape[1] == D
    ren = {"cand_kernel": "rnn_kernel", "cand_bias": "rnn_bias"} if cfg.get("graph_rnn_cell") == "RNN" else {}
    weights = []
    for l in range(len(m.params["layer_timesteps"])):
        w = {"edge_weights": m.gnn_weights.edge_weights[l].detach().numpy().reshape(m.num_edge_types, D, D)}
        if m.params["use_edge_bias"]:
            w["edge_biases"] = m.gnn_weights.edge_biases[l].detach().numpy()
        if m.params["use_propagation_attention"]:
            w["edge_type_attention_weights"] = m.gnn_weights.edge_type_attention_weights[l].detach().numpy()
        cell = {k: v.detach().numpy() for k, v in m.gnn_weights.rnn_cells[l].items()}
        if "cand_input_kernel" in cell:
            cell["cand_kernel"] = np.concatenate([cell.pop("cand_input_kernel"), cell.pop("cand_hidden_kernel")], axis=0)
        w.update({ren.get(k, k): v for k, v in cell.items()})
        weights.append(w)
    ref = O.sparse_propagation_np(feed["initial_node_representation"], [feed[k] for k in m.placeholders["adjacency_lists"]],
                                  feed["num_incoming_edges_per_type"], weights, m.params, dtype=np.float64)
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6)
    l0 = m.run_epoch("valid0", m.valid_data, False)[0]           # and it trains: gradients flow back through pad / slice to the D-wide variables
    for ep in range(4):
        m.run_epoch("train%d" % ep, m.train_data, True)
    assert m.run_epoch("valid1", m.valid_data, False)[0] < l0
    assert tuple(m.gnn_weights.edge_weights[0].shape) == (m.num_edge_types * D, D)


def test_dense_hidden_size_that_is_not_a_multiple_of_4_runs_zero_padded(tmp_path, stand_in):
    import torch
    mols = synthetic.make_molecules(48, seed=4)
    args = {"--log_dir": str(tmp_path), "--device": "cpu", "--train_data": mols[:32], "--valid_data": mols[32:],
            "--config": {"hidden_size": 10, "batch_size": 4, "num_timesteps": 2, "learning_rate": 0.01, "num_epochs": 1}}
    m = chem_dense.DenseGGNNChemModel(args)
    assert m._padded_hidden == 12 and m.engine.D == 12
    feed = next(iter(m.make_minibatch_iterator(m.valid_data, False)))
    m.feed = feed
    with torch.no_grad():
        got = m.compute_final_node_representations().numpy()
    w = {"edge_weights": m.weights["edge_weights"].detach().numpy(), "edge_biases": m.weights["edge_biases"].detach().numpy()}
    w.update({k: v.detach().numpy() for k, v in m.weights["node_gru"].items()})
    ref = O.dense_propagation_loops(feed["initial_node_representation"], feed["adjacency_matrix"], w, {"num_timesteps": 2, "use_edge_bias": True})
    assert got.shape == ref.shape and got.shape[-1] == 10
    np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-6)
    l0 = m.run_epoch("valid0", m.valid_data, False)[0]
    for ep in range(4):
        m.run_epoch("train%d" % ep, m.train_data, True)
    assert m.run_epoch("valid1", m.valid_data, False)[0] < l0
