"""GPU: fused gated-regression readout (ggnn_readout_* through the C ABI) against the oracle -- forward and gradients,
sparse (grouped and shuffled node lists) and dense (masked) layouts."""
import numpy as np
import pytest

from oracle import ggnn_oracle as O
from tests import _util as U

pytestmark = pytest.mark.gpu


def _case(V, D, G, seed):
    rng = np.random.default_rng(seed)
    last_h = rng.normal(0, 0.5, (V, D)).astype(np.float32)
    h0 = rng.normal(0, 0.5, (V, D)).astype(np.float32)
    w = dict(w_gate=rng.normal(0, 0.3, (2 * D, 1)).astype(np.float32), b_gate=rng.normal(0, 0.3, 1).astype(np.float32),
             w_trans=rng.normal(0, 0.3, (D, 1)).astype(np.float32), b_trans=rng.normal(0, 0.3, 1).astype(np.float32))
    Gw = rng.normal(size=G).astype(np.float32)
    return last_h, h0, w, Gw


def _engine(D):
    from gated_graph_neural_network_samples_b200.engine import PropagationEngine
    return PropagationEngine({"hidden_size": D, "layer_timesteps": [1], "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}, 2)


def _run(eng, last_h, h0, w, Gw):
    import torch
    from gated_graph_neural_network_samples_b200.readout import gated_readout_function
    f = gated_readout_function()
    th = torch.from_numpy(last_h).cuda().requires_grad_(True)
    tw = {k: torch.from_numpy(v).cuda().requires_grad_(True) for k, v in w.items()}
    out = f.apply(eng, th, torch.from_numpy(h0).cuda(), tw["w_gate"], tw["b_gate"], tw["w_trans"], tw["b_trans"])
    (out * torch.from_numpy(Gw).cuda()).sum().backward()
    eng.sync_check()
    return out.detach().cpu().numpy(), th.grad.cpu().numpy(), {k: v.grad.cpu().numpy() for k, v in tw.items()}


def _ref(last_h, h0, w, Gw, **kw):
    import torch
    th = torch.tensor(last_h, dtype=torch.float64, requires_grad=True)
    tw = {k: torch.tensor(v, dtype=torch.float64, requires_grad=True) for k, v in w.items()}
    out = O.gated_regression_torch(th, torch.tensor(h0, dtype=torch.float64), tw["w_gate"], tw["b_gate"], tw["w_trans"], tw["b_trans"],
                                   dtype=torch.float64, **kw)
    (out * torch.tensor(Gw, dtype=torch.float64)).sum().backward()
    return out.detach().numpy(), th.grad.numpy(), {k: v.grad.numpy() for k, v in tw.items()}


def _cmp(got, ref, tag):
    err = float(np.max(np.abs(got - ref))) / max(float(np.max(np.abs(ref))), 1e-12)
    print("readout %-24s %.2e" % (tag, err))
    assert err < 1e-4, tag


@pytest.mark.parametrize("D,shuffled", [(100, False), (100, True), (20, False), (256, True)])
def test_sparse_readout_forward_and_gradients(D, shuffled):
    G = 37
    sizes = np.random.default_rng(3).integers(1, 30, G)
    sizes[5] = 0 if G > 5 else sizes[5]                       # a graph without nodes sums to 0
    gnl = np.repeat(np.arange(G, dtype=np.int32), sizes)
    if shuffled:
        gnl = np.random.default_rng(4).permutation(gnl)        # not grouped -> the atomic variant
    V = gnl.shape[0]
    last_h, h0, w, Gw = _case(V, D, G, 11)
    eng = _engine(D)
    eng.readout_set_graphs(G, graph_nodes_list=gnl)
    out, dh, dw = _run(eng, last_h.reshape(V, D), h0, w, Gw)
    r_out, r_dh, r_dw = _ref(last_h, h0, w, Gw, graph_nodes_list=gnl, num_graphs=G)
    _cmp(out, r_out, "forward"); _cmp(dh, r_dh, "d h_last")
    for k in r_dw:
        _cmp(dw[k], r_dw[k], "d " + k)
    if not shuffled:                                           # grouped lists are summed in node order: run-to-run identical
        out2, _, _ = _run(eng, last_h, h0, w, Gw)
        np.testing.assert_array_equal(out, out2)


def test_dense_readout_masked():
    b, v, D = 9, 16, 24
    last_h, h0, w, Gw = _case(b * v, D, b, 5)
    mask = (np.random.default_rng(6).random((b, v)) < 0.7).astype(np.float32)
    eng = _engine(D)
    eng.readout_set_graphs(b, nodes_per_graph=v, node_mask=mask)
    out, dh, dw = _run(eng, last_h, h0, w, Gw)
    r_out, r_dh, r_dw = _ref(last_h.reshape(b, v, D), h0.reshape(b, v, D), w, Gw, node_mask=mask)
    _cmp(out, r_out, "dense forward"); _cmp(dh, r_dh.reshape(b * v, D), "dense d h_last")
    for k in r_dw:
        _cmp(dw[k], r_dw[k], "dense d " + k)


def test_readout_errors():
    from gated_graph_neural_network_samples_b200.engine import GgnnError
    eng = _engine(8)
    with pytest.raises(GgnnError, match="out of range"):
        eng.readout_set_graphs(2, graph_nodes_list=np.array([0, 1, 2], np.int32))
    with pytest.raises(GgnnError):
        eng.readout_set_graphs(3, nodes_per_graph=0)


@pytest.mark.parametrize("name", ["true_default_shape", "rnn_relu_bias_sum"])
def test_one_call_loss_and_accuracy_match_the_reference_make_model(golden_dir, name):
    """ggnn_run_sparse_host_readout = the fetches of sess.run([loss, accuracy_task0], feed_dict) (chem_tensorflow.py:231-235): the
    refgraph fixtures hold loss and MAE computed by the reference's OWN make_model (one graph of the batch is unlabeled, so the mask and
    the SMALL_NUMBER normalisation are exercised); only 2 floats come back over PCIe."""
    import json
    import os
    import torch
    from gated_graph_neural_network_samples_b200.engine import PropagationEngine
    z = np.load(os.path.join(golden_dir, "refgraph_sparse_%s.npz" % name))
    p = json.loads(str(z["params_json"]))
    w = [{k[len("w%d_" % l):]: z[k] for k in z.files if k.startswith("w%d_" % l)} for l in range(len(p["layer_timesteps"]))]
    adj = [z["adj%d" % e] for e in range(4)]
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    task = (f32(z["ro_w_gate"]), f32(z["ro_b_gate"]), f32(z["ro_w_trans"]), f32(z["ro_b_trans"]))
    for precision in ("fp32", "bf16x3"):
        eng = PropagationEngine(p, 4, precision=precision)
        eng.set_weights(U.to_cuda_weights(w))
        loss, acc = eng.run_sparse_host_readout(adj, z["indeg"].astype(np.float32), z["h0"].astype(np.float32), z["graph_nodes_list"],
                                                int(z["num_graphs"]), [task], z["target_values"], z["target_mask"])
        print("one-call %s %s: loss %.6f (ref %.6f) mae %.6f (ref %.6f)" % (name, precision, loss[0], float(z["loss"]), acc[0], float(z["accuracy"])))
        assert abs(loss[0] - float(z["loss"])) < 1e-4 * max(1.0, abs(float(z["loss"])))
        assert abs(acc[0] - float(z["accuracy"])) < 1e-4 * max(1.0, abs(float(z["accuracy"])))


def test_one_call_two_tasks_against_torch():
    import torch
    from gated_graph_neural_network_samples_b200.engine import PropagationEngine
    from oracle import ggnn_oracle as O
    from tests.test_gpu_parity import CFG2
    _, b = U.molecule_batch(40, 100, seed=3)
    w = O.init_sparse_weights(CFG2, 4, np.random.default_rng(1))
    eng = PropagationEngine(CFG2, 4, precision="bf16x3")
    eng.set_weights(U.to_cuda_weights(w))
    G = 40
    rng = np.random.default_rng(2)
    tasks = [tuple(torch.from_numpy(a.astype(np.float32)).cuda() for a in (rng.normal(0, 0.2, 200), rng.normal(0, 0.1, 1), rng.normal(0, 0.2, 100), rng.normal(0, 0.1, 1)))
             for _ in range(2)]
    tv = rng.normal(size=(2, G)).astype(np.float32)
    tm = (rng.random((2, G)) < 0.7).astype(np.float32)
    tm[1, :] = 0.0                                               # a task without any label in this batch: 0 / (0 + 1e-7) = 0
    loss, acc = eng.run_sparse_host_readout(b["adjacency_lists"], b["num_incoming_edges_per_type"], b["initial_node_representation"],
                                            b["graph_nodes_list"], G, tasks, tv, tm)
    h0 = torch.from_numpy(b["initial_node_representation"]).cuda()
    eng.set_graph_sparse(b["adjacency_lists"], b["num_incoming_edges_per_type"])
    out = eng.forward(h0)
    gnl = torch.from_numpy(np.asarray(b["graph_nodes_list"])).long().cuda()
    for t, (wg, bg, wt, bt) in enumerate(tasks):
        gated = torch.sigmoid(torch.cat([out, h0], -1) @ wg.view(-1, 1) + bg) * (out @ wt.view(-1, 1) + bt)
        pred = torch.zeros(G, 1, device="cuda").index_add_(0, gnl, gated).squeeze(-1)
        diff = (pred - torch.from_numpy(tv[t]).cuda()) * torch.from_numpy(tm[t]).cuda()
        num = float(tm[t].sum()) + 1e-7
        assert abs(loss[t] - float((0.5 * diff * diff).sum()) / num) < 1e-4 * max(1.0, abs(float(loss[t])))
        assert abs(acc[t] - float(diff.abs().sum()) / num) < 1e-4 * max(1.0, abs(float(acc[t])))
    assert loss[1] == 0.0 and acc[1] == 0.0
