"""GPU: ``ggnn_prepare_graph_sparse`` (host half, producer thread) + ``ggnn_set_graph_prepared`` (device half) against the one-call
``ggnn_set_graph_sparse`` -- SURVEY 8 f3, the overlap of chem_tensorflow.py:225 / utils.py:16-36.  Same plan, same uploaded CSR (bit for
bit), same node states; a producer thread that runs ahead of the consumer; gradients through a prepared graph; the plug-in's own use."""
import queue
import threading

import numpy as np
import pytest

from oracle import ggnn_oracle as O
from tests import _util as U

pytestmark = pytest.mark.gpu

GRU = {"layer_timesteps": [2, 2], "residual_connections": {"1": [0]}, "use_edge_bias": True, "use_edge_msg_avg_aggregation": True,
       "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}


def _engine(p, T, precision, seed=1):
    from gated_graph_neural_network_samples_b200.engine import PropagationEngine
    w = O.init_sparse_weights(p, T, np.random.default_rng(seed))
    eng = PropagationEngine(p, T, precision=precision)
    eng.set_weights(U.to_cuda_weights(w))
    return eng, w


def _forward(eng, h0):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(h0, dtype=np.float32)).cuda()
    out = eng.forward(t)
    eng.sync_check()
    eng._keepalive = (t, out)
    return out.cpu().numpy()


@pytest.mark.parametrize("D,T,precision,fragment", [(100, 4, "bf16x3", "LOCAL"), (64, 4, "fp32", "LOCAL"), (256, 8, "bf16x3", "STREAM")])
def test_prepared_equals_direct(D, T, precision, fragment):
    p = dict(GRU, hidden_size=D)
    _, b = U.molecule_batch(96, D, T=T, seed=21)
    adj, indeg, h0 = b["adjacency_lists"], b["num_incoming_edges_per_type"], b["initial_node_representation"]
    eng, w = _engine(p, T, precision)
    eng.set_graph_sparse(adj, indeg)
    plan_direct, csr_direct = eng.plan, eng.csr()
    out_direct = _forward(eng, h0)
    g = eng.prepare_graph_sparse(adj, indeg)
    assert g.info()["plan"] == plan_direct and fragment in plan_direct
    eng2, _ = _engine(p, T, precision)
    eng2.set_graph_prepared(g)
    assert eng2.plan == plan_direct
    for a, r in zip(eng2.csr(), csr_direct):                       # what sits on the device: bit for bit
        np.testing.assert_array_equal(a, r)
    out_prep = _forward(eng2, h0)
    if fragment == "STREAM" or precision == "fp32":                # deterministic kernels: identical bits
        np.testing.assert_array_equal(out_prep, out_direct)
    else:                                                          # two MMA issuers: ~1e-6 run-to-run rounding noise
        np.testing.assert_allclose(out_prep, out_direct, rtol=1e-4, atol=1e-5)
    ref = O.sparse_propagation_np(h0, adj, indeg, w, p, dtype=np.float64)
    assert U.max_rel_err(out_prep, ref) < 1e-4


def test_producer_thread_runs_ahead_of_the_consumer():
    """The training-loop shape: a producer thread prepares batch i+1.. (rebuilding a small pool of prepared graphs in place) while the
    consumer uploads and runs batch i.  Every result equals the one-call path's."""
    D, T = 100, 4
    p = dict(GRU, hidden_size=D)
    batches = [U.molecule_batch(n, D, T=T, seed=s)[1] for n, s in ((64, 1), (20, 2), (128, 3), (40, 4), (90, 5), (64, 6), (10, 7), (77, 8))]
    eng, w = _engine(p, T, "bf16x3")
    direct = []
    for b in batches:
        eng.set_graph_sparse(b["adjacency_lists"], b["num_incoming_edges_per_type"])
        direct.append(_forward(eng, b["initial_node_representation"]))
    q, pool = queue.Queue(maxsize=3), []

    def producer():
        for b in batches:
            g = eng.prepare_graph_sparse(b["adjacency_lists"], b["num_incoming_edges_per_type"], save_for_backward=False,
                                         reuse=pool.pop() if pool else None)
            q.put((g, b))
        q.put(None)

    th = threading.Thread(target=producer, daemon=True)
    th.start()
    got = []
    while True:
        item = q.get()
        if item is None:
            break
        g, b = item
        eng.set_graph_prepared(g)
        pool.append(g)                                             # handed back right after the upload was ENQUEUED: the rebuild waits for it
        got.append(_forward(eng, b["initial_node_representation"]))
    th.join()
    assert len(got) == len(direct)
    for a, r in zip(got, direct):
        np.testing.assert_allclose(a, r, rtol=1e-4, atol=1e-5)


def test_gradients_through_a_prepared_graph_and_the_save_flag():
    import torch
    from gated_graph_neural_network_samples_b200.engine import GgnnError
    D, T = 32, 4
    p = dict(GRU, hidden_size=D)
    _, b = U.molecule_batch(24, D, T=T, seed=9)
    adj, indeg, h0 = b["adjacency_lists"], b["num_incoming_edges_per_type"], b["initial_node_representation"]
    G = np.random.default_rng(4).normal(size=h0.shape).astype(np.float32)

    def grads(prepared):
        eng, w = _engine(p, T, "fp32")
        dev_w = U.to_cuda_weights(w)
        eng.set_weights(dev_w)
        eng.set_save_for_backward(True)
        if prepared:
            eng.set_graph_prepared(eng.prepare_graph_sparse(adj, indeg, save_for_backward=True))
        else:
            eng.set_graph_sparse(adj, indeg)
        th0 = torch.from_numpy(h0).cuda()
        out = eng.forward(th0)
        gr = [{k: torch.zeros_like(v) for k, v in lw.items()} for lw in dev_w]
        d_h0 = torch.zeros_like(th0)
        eng.backward(torch.from_numpy(G).cuda(), gr, d_h0)
        eng.sync_check()
        return out.cpu().numpy(), d_h0.cpu().numpy(), [{k: v.cpu().numpy() for k, v in lw.items()} for lw in gr]

    o1, d1, g1 = grads(False)
    o2, d2, g2 = grads(True)
    np.testing.assert_array_equal(o1, o2)
    np.testing.assert_allclose(d1, d2, rtol=1e-5, atol=1e-6)       # weight gradients use float atomics: order noise only
    for a, r in zip(g2, g1):
        for k in r:
            np.testing.assert_allclose(a[k], r[k], rtol=1e-4, atol=1e-5, err_msg=k)
    # a graph prepared WITHOUT the source-keyed CSR cannot serve a training step: refused at adoption, not at ggnn_backward
    eng, _ = _engine(p, T, "fp32")
    eng.set_save_for_backward(True)
    with pytest.raises(GgnnError, match="save_for_backward"):
        eng.set_graph_prepared(eng.prepare_graph_sparse(adj, indeg, save_for_backward=False))
    # and one built for another configuration is refused too
    other, _ = _engine(dict(p, hidden_size=64), T, "fp32")
    with pytest.raises(GgnnError, match="different engine configuration"):
        other.set_graph_prepared(eng.prepare_graph_sparse(adj, indeg))


def test_plugin_epochs_use_prepared_graphs_and_match_the_one_call_path(tmp_path):
    """SparseGGNNChemModel.make_minibatch_iterator prepares every batch in the ThreadedIterator's producer thread; validation loss and a
    training epoch must equal a model whose engine has no prepare step (the one-call path)."""
    from gated_graph_neural_network_samples_b200 import synthetic
    from gated_graph_neural_network_samples_b200.chem_sparse import SparseGGNNChemModel
    mols = synthetic.make_molecules(96, seed=1)

    def model():
        np.random.seed(0)
        return SparseGGNNChemModel({"--log_dir": str(tmp_path), "--train_data": mols[:64], "--valid_data": mols[64:],
                                    "--config": {"hidden_size": 32, "batch_size": 300, "layer_timesteps": [2, 1], "residual_connections": {"1": [0]},
                                                 "edge_weight_dropout_keep_prob": 1.0, "learning_rate": 0.01, "num_epochs": 1, "random_seed": 3}})
    a, b = model(), model()
    b.prepare_graphs_in_producer = False                      # the one-call path: ggnn_set_graph_sparse inside hook 2
    feeds = list(a.make_minibatch_iterator(a.valid_data, False))
    assert all(f.get("_prepared_graph") is not None and not f["_prepared_graph"].for_training for f in feeds)
    assert all("_prepared_graph" not in f for f in b.make_minibatch_iterator(b.valid_data, False))
    for (_, va), (_, vb) in zip(a.trainable_variables(), b.trainable_variables()):
        np.testing.assert_array_equal(va.detach().cpu().numpy(), vb.detach().cpu().numpy())
    la, lb = a.run_epoch("valid", a.valid_data, False)[0], b.run_epoch("valid", b.valid_data, False)[0]
    assert abs(la - lb) < 1e-5 * max(1.0, abs(la))
    np.random.seed(5); ta = a.run_epoch("train", a.train_data, True)
    np.random.seed(5); tb = b.run_epoch("train", b.train_data, True)
    assert ta[4] == tb[4] >= 3 and abs(ta[0] - tb[0]) < 1e-3 * max(1.0, abs(ta[0]))
    assert len(a._prepared_pool) >= 1                          # prepared graphs went through hook 2 and back to the pool


def test_dense_prepared_equals_direct_and_weighted_matrices_keep_the_matrix_walk():
    import torch
    from gated_graph_neural_network_samples_b200 import packing, synthetic
    from gated_graph_neural_network_samples_b200.engine import GgnnError, PropagationEngine
    D, T, v = 100, 4, 29
    db = packing.pack_dense_batch(synthetic.make_molecules(48, seed=6), v, D, T)
    A = np.asarray(db["adjacency_matrix"], np.float32)
    b = A.shape[0]
    h0 = (db["initial_node_representation"] + np.random.default_rng(1).normal(0, 0.1, db["initial_node_representation"].shape)).astype(np.float32)
    dp = {"num_timesteps": 3, "use_edge_bias": True}
    dw = O.init_dense_weights({"hidden_size": D}, T, np.random.default_rng(5))
    params = U.dense_params_as_engine_params(dp, D)
    w = dict(dw, edge_biases=np.asarray(dw["edge_biases"]).reshape(T, D))
    ref = O.dense_propagation_loops(h0, A, dw, dp).reshape(b * v, D)

    def run(precision, prepared):
        eng = PropagationEngine(params, T, precision=precision)
        eng.set_weights(U.to_cuda_weights([w]))
        if prepared:
            g = eng.prepare_graph_dense(A)
            assert "binary dense adjacency -> CSR" in g.info()["plan"]
            eng.set_graph_prepared(g)
        else:
            eng.set_graph_dense(A)
        return _forward(eng, h0.reshape(b * v, D)), eng.plan

    for precision in ("fp32", "bf16x3"):
        (o1, p1), (o2, p2) = run(precision, False), run(precision, True)
        assert p1 == p2
        if precision == "fp32":
            np.testing.assert_array_equal(o1, o2)
        else:
            np.testing.assert_allclose(o1, o2, rtol=1e-4, atol=1e-5)
        assert U.max_rel_err(o2, ref) < 1e-4
    W = A.copy()
    W[A > 0] = 0.5                                                  # a weighted adjacency: the CSR shortcut does not apply ...
    eng = PropagationEngine(params, T, precision="fp32")
    eng.set_weights(U.to_cuda_weights([w]))
    with pytest.raises(GgnnError, match="not 0/1"):
        eng.prepare_graph_dense(W)
    eng.set_graph_dense(W)                                          # ... the one-call path walks the matrix
    got = _forward(eng, h0.reshape(b * v, D))
    assert U.max_rel_err(got, O.dense_propagation_loops(h0, W, dw, dp).reshape(b * v, D)) < 1e-4


def test_dense_plugin_epochs_use_prepared_graphs(tmp_path):
    from gated_graph_neural_network_samples_b200 import synthetic
    from gated_graph_neural_network_samples_b200.chem_dense import DenseGGNNChemModel
    mols = synthetic.make_molecules(64, seed=2)

    def model():
        np.random.seed(0)
        return DenseGGNNChemModel({"--log_dir": str(tmp_path), "--train_data": mols[:48], "--valid_data": mols[48:],
                                   "--config": {"hidden_size": 32, "batch_size": 8, "num_timesteps": 2, "learning_rate": 0.01, "num_epochs": 1,
                                                "random_seed": 3}})
    a, b = model(), model()
    b.prepare_graphs_in_producer = False
    assert all(f.get("_prepared_graph") is not None for f in a.make_minibatch_iterator(a.valid_data, False))
    assert all("_prepared_graph" not in f for f in b.make_minibatch_iterator(b.valid_data, False))
    la, lb = a.run_epoch("valid", a.valid_data, False)[0], b.run_epoch("valid", b.valid_data, False)[0]
    assert abs(la - lb) < 1e-5 * max(1.0, abs(la))
    np.random.seed(5); ta = a.run_epoch("train", a.train_data, True)
    np.random.seed(5); tb = b.run_epoch("train", b.train_data, True)
    assert ta[4] == tb[4] >= 3 and abs(ta[0] - tb[0]) < 1e-3 * max(1.0, abs(ta[0]))
    assert len(a._prepared_pool) >= 1
