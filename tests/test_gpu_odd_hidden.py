"""GPU: hidden sizes that are not multiples of 4 (the reference accepts any; the engine moves rows as 16-byte vectors).  The sparse plug-in
zero-pads states and weights at the engine boundary and slices the result; padded units stay exactly 0, so the real units are unchanged."""
import numpy as np
import pytest

from oracle import ggnn_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("cfg", [{"hidden_size": 10}, {"hidden_size": 30, "graph_rnn_cell": "RNN", "graph_rnn_activation": "ReLU", "use_edge_bias": True},
                                 {"hidden_size": 9, "graph_rnn_cell": "CudnnCompatibleGRUCell"}, {"hidden_size": 101}])
def test_plugin_runs_any_hidden_size_and_matches_the_oracle(tmp_path, cfg):
    import torch
    from gated_graph_neural_network_samples_b200 import synthetic
    from gated_graph_neural_network_samples_b200.chem_sparse import SparseGGNNChemModel
    mols = synthetic.make_molecules(96, seed=3)
    base = {"batch_size": 400, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]}, "edge_weight_dropout_keep_prob": 1.0,
            "learning_rate": 0.01, "num_epochs": 1}
    base.update(cfg)
    m = SparseGGNNChemModel({"--log_dir": str(tmp_path), "--train_data": mols[:64], "--valid_data": mols[64:], "--config": base})
    D = cfg["hidden_size"]
    assert m.engine.D == (D + 3) // 4 * 4 != D
    feed = next(iter(m.make_minibatch_iterator(m.valid_data, False)))
    m.feed = feed
    with torch.no_grad():
        got = m.compute_final_node_representations().cpu().numpy()
    assert got.shape[1] == D
    ren = {"cand_kernel": "rnn_kernel", "cand_bias": "rnn_bias"} if cfg.get("graph_rnn_cell") == "RNN" else {}
    weights = []
    for l in range(2):
        w = {"edge_weights": m.gnn_weights.edge_weights[l].detach().cpu().numpy().reshape(m.num_edge_types, D, D)}
        if m.params["use_edge_bias"]:
            w["edge_biases"] = m.gnn_weights.edge_biases[l].detach().cpu().numpy()
        cell = {k: v.detach().cpu().numpy() for k, v in m.gnn_weights.rnn_cells[l].items()}
        if "cand_input_kernel" in cell:
            cell["cand_kernel"] = np.concatenate([cell.pop("cand_input_kernel"), cell.pop("cand_hidden_kernel")], axis=0)
        w.update({ren.get(k, k): v for k, v in cell.items()})
        weights.append(w)
    ref = O.sparse_propagation_np(feed["initial_node_representation"], [feed[k] for k in m.placeholders["adjacency_lists"]],
                                  feed["num_incoming_edges_per_type"], weights, m.params, dtype=np.float64)
    err = float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30))
    print("hidden %d (engine %d) max rel err %.2e  [%s]" % (D, m.engine.D, err, m.engine.plan[:50]))
    assert err < 1e-4
    l0 = m.run_epoch("valid0", m.valid_data, False)[0]
    for ep in range(5):
        m.run_epoch("train%d" % ep, m.train_data, True)
    assert m.run_epoch("valid1", m.valid_data, False)[0] < l0
