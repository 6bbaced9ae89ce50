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
