"""GPU parity of the tcgen05 tensor-core path (precision "bf16x3": bf16 hi/lo split, 3 MMAs per product)
against the float64 oracle, at the north-star tolerance (1e-4 relative on node hidden states)."""
import json
import os

import numpy as np
import pytest

from gated_graph_neural_network_samples_b200 import packing, synthetic
from oracle import ggnn_oracle as O
from tests import _util as U
from tests.test_gpu_parity import CFG1_TRUE, CFG2, CFG5

pytestmark = pytest.mark.gpu
PREC = "bf16x3"


def _check(got, ref, tag=""):
    assert np.all(np.isfinite(got))
    err = U.max_rel_err(got, ref)
    print("[%s] %s max|err|/max|ref| = %.3e" % (PREC, tag, err))
    assert err < 1e-4
    np.testing.assert_allclose(got, ref, rtol=1e-4, atol=1e-4 * float(np.max(np.abs(ref))))


@pytest.mark.parametrize("name", ["gru_bias_avg_res", "gru_plain", "rgcn_relu"])
def test_golden_sparse(golden_dir, name):
    z, p, w, adj = U.load_golden_sparse(golden_dir, name)
    got, eng = U.engine_sparse(p, 4, w, adj, z["indeg"], z["h0"], precision=PREC, return_engine=True)
    assert "tcgen05" in eng.plan
    _check(got, z["final"], name)
    for li in range(len(p["layer_timesteps"]) + 1):
        _check(eng.layer_state(li).cpu().numpy(), z["state%d" % li], "%s layer %d" % (name, li))


def test_golden_dense(golden_dir):
    z = np.load(os.path.join(golden_dir, "prop_dense.npz"))
    p = json.loads(str(z["params_json"]))
    w = {k[2:]: z[k] for k in z.files if k.startswith("w_")}
    _check(U.engine_dense(p, 4, w, z["adj"], z["h0"], precision=PREC), z["final"], "dense golden")


@pytest.mark.parametrize("params,n,T", [(CFG2, 256, 4), (CFG1_TRUE, 256, 4),
                                        (dict(CFG2, use_edge_bias=True, graph_rnn_activation="relu", hidden_size=64), 64, 4),
                                        (dict(CFG1_TRUE, hidden_size=128, graph_rnn_cell="RNN"), 100, 4)])
def test_molecule_batches_vs_oracle(params, n, T):
    _, b = U.molecule_batch(n, params["hidden_size"], T=T, seed=5)
    w = O.init_sparse_weights(params, T, np.random.default_rng(1))
    ref = O.sparse_propagation_np(b["initial_node_representation"], b["adjacency_lists"],
                                  b["num_incoming_edges_per_type"], w, params, dtype=np.float64)
    got = U.engine_sparse(params, T, w, b["adjacency_lists"], b["num_incoming_edges_per_type"],
                          b["initial_node_representation"], precision=PREC)
    _check(got, ref, "molecules D=%d" % params["hidden_size"])


def test_global_mode_matches_local_mode(monkeypatch):
    params = dict(CFG1_TRUE, hidden_size=48, use_edge_bias=True)
    _, b = U.molecule_batch(60, 48, seed=9)
    w = O.init_sparse_weights(params, 4, np.random.default_rng(3))
    ref = O.sparse_propagation_np(b["initial_node_representation"], b["adjacency_lists"],
                                  b["num_incoming_edges_per_type"], w, params, dtype=np.float64)
    for fg in ("0", "1"):
        monkeypatch.setenv("GGNN_FORCE_GLOBAL", fg)
        got, eng = U.engine_sparse(params, 4, w, b["adjacency_lists"], b["num_incoming_edges_per_type"],
                                   b["initial_node_representation"], precision=PREC, return_engine=True)
        assert ("GLOBAL" in eng.plan) == (fg == "1")
        _check(got, ref, "force_global=%s" % fg)


def test_single_large_graph_rgcn_global_mode(monkeypatch):
    adj, indeg = synthetic.random_sparse_graph(10000, 40000, 4, seed=2)
    h0 = np.random.default_rng(4).normal(0, 0.1, (10000, 100)).astype(np.float32)
    w = O.init_sparse_weights(CFG5, 4, np.random.default_rng(1))
    ref = O.sparse_propagation_np(h0, adj, indeg, w, CFG5, dtype=np.float64)
    got, eng = U.engine_sparse(CFG5, 4, w, adj, indeg, h0, precision=PREC, return_engine=True)
    assert "STREAM" in eng.plan            # a 10 000-node component does not fit a tile: streaming plan (tests/test_gpu_stream.py)
    _check(got, ref, "cfg5")
    monkeypatch.setenv("GGNN_TC_STREAM", "0")   # ... and the one-launch-per-step form of the tile kernel still serves it
    got, eng = U.engine_sparse(CFG5, 4, w, adj, indeg, h0, precision=PREC, return_engine=True)
    assert "GLOBAL" in eng.plan
    _check(got, ref, "cfg5 global")


def test_dense_cfg3_shape():
    D, T, steps = 100, 4, 4
    mols = synthetic.make_molecules(64, seed=21)
    db = packing.pack_dense_batch(mols, 32, D, T)
    h0 = (db["initial_node_representation"] + np.random.default_rng(2).normal(0, 0.1, db["initial_node_representation"].shape)).astype(np.float32)
    dw = O.init_dense_weights({"hidden_size": D}, T, np.random.default_rng(5))
    dp = {"num_timesteps": steps, "use_edge_bias": True}
    _check(U.engine_dense(dp, T, dw, db["adjacency_matrix"], h0, precision=PREC),
           O.dense_propagation_loops(h0, db["adjacency_matrix"], dw, dp), "dense cfg3")


def test_edge_cases():
    params = dict(CFG2, hidden_size=8, layer_timesteps=[2], use_edge_bias=True)
    w = O.init_sparse_weights(params, 4, np.random.default_rng(0))
    rng = np.random.default_rng(1)
    h0 = rng.normal(size=(1, 8)).astype(np.float32)
    adj = [np.zeros((0, 2), np.int32)] * 4
    indeg = np.zeros((1, 4), np.float32)
    _check(U.engine_sparse(params, 4, w, adj, indeg, h0, precision=PREC), O.sparse_propagation_loops(h0, adj, indeg, w, params), "isolated")
    h0 = rng.normal(size=(7, 8)).astype(np.float32)
    a2 = np.array([[0, 1], [1, 0], [1, 0], [3, 3], [5, 6], [6, 5]], np.int32)
    adj = [np.zeros((0, 2), np.int32), np.zeros((0, 2), np.int32), a2, np.zeros((0, 2), np.int32)]
    indeg = np.zeros((7, 4), np.float32)
    np.add.at(indeg[:, 2], a2[:, 1], 1)
    _check(U.engine_sparse(params, 4, w, adj, indeg, h0, precision=PREC), O.sparse_propagation_loops(h0, adj, indeg, w, params), "one type")


def test_fast_single_bf16_mode_is_close_but_outside_the_bar():
    _, b = U.molecule_batch(64, 100, seed=5)
    w = O.init_sparse_weights(CFG2, 4, np.random.default_rng(1))
    ref = O.sparse_propagation_np(b["initial_node_representation"], b["adjacency_lists"],
                                  b["num_incoming_edges_per_type"], w, CFG2, dtype=np.float64)
    got = U.engine_sparse(CFG2, 4, w, b["adjacency_lists"], b["num_incoming_edges_per_type"],
                          b["initial_node_representation"], precision="bf16")
    err = U.max_rel_err(got, ref)
    print("[bf16] max|err|/max|ref| = %.3e" % err)
    assert err < 3e-2


def test_hidden_256_takes_the_streaming_tensor_core_plan():
    """D > 128 does not fit the tile-local fused kernel: the engine plans the streaming tcgen05 path (tests/test_gpu_stream.py)."""
    _, b = U.molecule_batch(8, 256, seed=5)
    p = dict(CFG2, hidden_size=256, layer_timesteps=[1])
    w = O.init_sparse_weights(p, 4, np.random.default_rng(1))
    ref = O.sparse_propagation_np(b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"], w, p, dtype=np.float64)
    got, eng = U.engine_sparse(p, 4, w, b["adjacency_lists"], b["num_incoming_edges_per_type"], b["initial_node_representation"],
                               precision=PREC, return_engine=True)
    assert "STREAM" in eng.plan
    _check(got, ref, "D=256")
