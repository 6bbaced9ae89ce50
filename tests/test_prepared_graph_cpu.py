"""CPU: the HOST half of ``ggnn_set_graph_sparse`` as its own object (``ggnn_prepared_graph``, SURVEY 8 f3: the CSR / tile-plan build that a
producer thread runs while the GPU works on the previous batch -- the ThreadedIterator overlap of chem_tensorflow.py:225 / utils.py:16-36).

``ggnn_set_graph_sparse`` IS ``ggnn_prepare_graph_sparse`` + ``ggnn_set_graph_prepared`` on an engine-owned prepared graph, and
``ggnn_host_prepare_graph_sparse`` runs the same builder without an engine or a GPU.  So everything the engine uploads is pinned here, bit
for bit, against the independent host functions and NumPy: the integer path (CSR = NumPy's stable sort by target), the tile plan, the
mean-aggregation denominators, the streaming plan's gather table."""
import ctypes as C
import threading

import numpy as np
import pytest

from gated_graph_neural_network_samples_b200 import _lib, packing, synthetic
from gated_graph_neural_network_samples_b200.engine import GgnnError, PreparedGraph
from oracle import ggnn_oracle as O

GRU = {"layer_timesteps": [2, 1], "residual_connections": {"1": [0]}, "use_edge_bias": False, "use_edge_msg_avg_aggregation": True,
       "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}
PREC = {"fp32": 0, "bf16x3": 1, "bf16": 2}


def _molecules(n, T, seed):
    mols = synthetic.make_molecules(n, seed=seed, num_bond_types=T)
    b = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), 8, T)
    return b["adjacency_lists"], b["num_incoming_edges_per_type"]


def _one_big_graph(V, T, E, seed):
    rng = np.random.default_rng(seed)
    adj = []
    for t in range(T):
        src = rng.integers(0, V, E)
        adj.append(np.stack([src, (src + rng.integers(1, 40, E)) % V], 1).astype(np.int32))
    adj[0] = np.concatenate([adj[0], np.stack([np.arange(V - 1), np.arange(1, V)], 1).astype(np.int32)])
    indeg = np.zeros((V, T), np.float32)
    for t in range(T):
        np.add.at(indeg[:, t], adj[t][:, 1], 1.0)
    return adj, indeg


def _ptrs(adjs):
    adjs = [np.ascontiguousarray(np.asarray(a, np.int32).reshape(-1, 2)) for a in adjs]
    return adjs, (C.c_void_p * len(adjs))(*[a.ctypes.data for a in adjs]), (C.c_int32 * len(adjs))(*[a.shape[0] for a in adjs])


def _host_csr(adjs, V):
    lib = _lib.load()
    adjs, ptrs, counts = _ptrs(adjs)
    T, M = len(adjs), sum(a.shape[0] for a in adjs)
    row_ptr, src, msg = np.empty(V * T + 1, np.int32), np.empty(max(M, 1), np.int32), np.empty(max(M, 1), np.int32)
    assert lib.ggnn_host_target_csr(V, T, ptrs, counts, row_ptr.ctypes.data, src.ctypes.data, msg.ctypes.data) == 0
    return row_ptr, src[:M], msg[:M]


def _host_plan(adjs, V, D, precision, sms):
    lib = _lib.load()
    adjs, ptrs, counts = _ptrs(adjs)
    ts, n, text = np.empty(V + 2, np.int32), C.c_int32(), C.create_string_buffer(512)
    assert lib.ggnn_host_tile_plan(D, len(adjs), PREC[precision], sms, V, ptrs, counts, ts.ctypes.data, V + 2, C.byref(n), text, 512) == 0
    return ts[:n.value + 1].copy(), text.value.decode()


CASES = [
    # (name, hidden, precision, batch builder, expected plan fragment)
    ("cfg2_like_tile_local", 100, "bf16x3", lambda: _molecules(64, 4, 5), "LOCAL"),
    ("fp32_tile_local", 64, "fp32", lambda: _molecules(40, 4, 6), "fp32-ffma LOCAL"),
    ("cfg4_like_streaming", 256, "bf16x3", lambda: _molecules(48, 8, 7), "STREAM"),
    ("one_big_graph_streaming", 100, "bf16x3", lambda: _one_big_graph(1500, 4, 2500, 8), "STREAM"),
    ("one_big_graph_fp32_global", 64, "fp32", lambda: _one_big_graph(700, 3, 900, 9), "GLOBAL"),
]


@pytest.mark.parametrize("name,D,precision,make,fragment", CASES, ids=[c[0] for c in CASES])
@pytest.mark.parametrize("save", [False, True])
def test_prepared_image_equals_the_independent_host_functions_and_numpy(name, D, precision, make, fragment, save):
    adj, indeg = make()
    T, V = len(adj), indeg.shape[0]
    p = dict(GRU, hidden_size=D)
    g = PreparedGraph.host_only(p, T, adj, indeg, precision=precision, num_sms=148, save_for_backward=save)
    info, arr = g.info(), g.arrays(T)
    assert info["num_nodes"] == V and info["num_messages"] == sum(len(a) for a in adj) and fragment in info["plan"], info
    # integer path: bit-exact against the stand-alone host function AND against NumPy's stable sort by target (reference message order)
    row_ptr, src, msg = _host_csr(adj, V)
    for k, ref in (("row_ptr", row_ptr), ("src", src), ("msg", msg)):
        np.testing.assert_array_equal(arr[k], ref, err_msg=k)
    o_row_ptr, o_src, o_typ, o_order = O.stable_target_csr(adj, V)
    np.testing.assert_array_equal(arr["msg"], o_order)
    np.testing.assert_array_equal(arr["src"], o_src)
    np.testing.assert_array_equal(arr["row_ptr"][::T], o_row_ptr)
    # tile plan: the one ggnn_host_tile_plan makes
    ts, text = _host_plan(adj, V, D, precision, 148)
    np.testing.assert_array_equal(arr["tile_start"], ts)
    assert info["plan"] == text and info["num_tiles"] == len(ts) - 1
    # mean aggregation denominator: fp32 sum over the type axis, + 1e-7 (sparse:207-209)
    ref_den = np.zeros(V, np.float32)
    for t in range(T):
        ref_den = (ref_den + indeg[:, t].astype(np.float32)).astype(np.float32)
    np.testing.assert_array_equal(arr["denom"], (ref_den + np.float32(1e-7)).astype(np.float32))
    if info["streaming"]:   # the gather table of the streaming kernels == ggnn_host_stream_tables
        lib = _lib.load()
        adjs, ptrs, counts = _ptrs(adj)
        nt = (V + 127) // 128
        pair, vptr, vsrc, tvp, nv = np.empty(nt * 128 * T, np.int32), np.empty(V * T + 2, np.int32), np.empty(info["num_messages"] + 1, np.int32), \
            np.empty(nt + 1, np.int32), C.c_int32()
        assert lib.ggnn_host_stream_tables(V, T, ptrs, counts, pair.ctypes.data, vptr.ctypes.data, vptr.size, vsrc.ctypes.data, vsrc.size,
                                           tvp.ctypes.data, C.byref(nv)) == 0
        np.testing.assert_array_equal(arr["pair_src"], pair)


def test_rebuild_in_place_and_from_a_producer_thread():
    """A prepared graph is reused batch after batch (its image keeps its allocation); building it from another thread gives the same
    bytes -- the builder shares no state with anything else."""
    p = dict(GRU, hidden_size=100)
    batches = [_molecules(n, 4, s) for n, s in ((50, 1), (20, 2), (70, 3))]
    fresh = [PreparedGraph.host_only(p, 4, a, d, precision="bf16x3").arrays(4) for a, d in batches]
    g = None
    out = []

    def producer():
        nonlocal g
        for a, d in batches:
            g = PreparedGraph.host_only(p, 4, a, d, precision="bf16x3", reuse=g)
            out.append(g.arrays(4))

    th = threading.Thread(target=producer)
    th.start(); th.join()
    assert len(out) == 3
    for got, ref in zip(out, fresh):
        for k in ref:
            np.testing.assert_array_equal(got[k], ref[k], err_msg=k)


def test_errors_are_reported_like_the_direct_call():
    p = dict(GRU, hidden_size=32)
    adj, indeg = _molecules(10, 4, 4)
    bad = [a.copy() for a in adj]
    bad[1] = np.concatenate([bad[1], np.array([[3, indeg.shape[0]]], np.int32)])   # target == V: TF-CPU gather raises on it
    with pytest.raises(GgnnError, match="out of range"):
        PreparedGraph.host_only(p, 4, bad, indeg)
    g = PreparedGraph.host_only(p, 4, adj, indeg)
    with pytest.raises(GgnnError):                                                  # a failed rebuild leaves the graph empty, not stale
        PreparedGraph.host_only(p, 4, bad, indeg, reuse=g)
    with pytest.raises(GgnnError):
        g.info()
    with pytest.raises(Exception, match="Unknown RNN cell type"):                   # sparse:112
        PreparedGraph.host_only(dict(p, graph_rnn_cell="lstm"), 4, adj, indeg)
    empty = PreparedGraph.host_only(p, 4, [np.zeros((0, 2), np.int32)] * 4, np.zeros((0, 4), np.float32))
    assert empty.info()["num_nodes"] == 0 and empty.info()["num_tiles"] == 0


@pytest.mark.parametrize("name,D,precision,make", [
    ("random_graph_streaming", 100, "bf16x3", lambda: _one_big_graph(6000, 4, 9000, 31)),        # many virtual rows (pairs with several messages)
    ("molecules_tile_local", 100, "bf16x3", lambda: _molecules(700, 4, 32)),
    ("molecules_streaming_8_types", 256, "bf16x3", lambda: _molecules(400, 8, 33)),
    ("molecules_fp32", 64, "fp32", lambda: _molecules(500, 4, 34)),
], ids=lambda x: x if isinstance(x, str) else None)
def test_image_is_identical_for_every_host_thread_count(monkeypatch, name, D, precision, make):
    """The builder splits every pass over host threads by target ranges (each thread walks the whole edge list in the reference's order and
    writes only its own rows): the packed image -- CSR, tile plan, masks, streaming tables, source-keyed CSR of the backward pass -- must
    not depend on the thread count, down to the last byte."""
    adj, indeg = make()
    p = dict(GRU, hidden_size=D)
    ref = None
    for nth in (1, 2, 3, 5, 8):
        monkeypatch.setenv("GGNN_HOST_THREADS", str(nth))
        g = PreparedGraph.host_only(p, len(adj), adj, indeg, precision=precision, save_for_backward=True)
        img, info = g.image(), g.info()
        if ref is None:
            ref, ref_info = img, info
        else:
            assert info == ref_info
            assert np.array_equal(img, ref), "image differs at %d host threads (first byte %d)" % (nth, int(np.argmax(img != ref)))


def test_dense_adjacency_prepares_like_its_edge_lists(monkeypatch):
    """ggnn_host_prepare_graph_dense (the host half of ggnn_set_graph_dense for the 0/1 adjacency the reference feeds, dense:30-36): the
    scan emits, per type, (source, target) pairs in the order (graph, target row, source column) with in-degree = row sums -- exactly
    NumPy's nonzero order -- and the image is the sparse builder's image of those lists, for every scan thread count."""
    from gated_graph_neural_network_samples_b200 import packing as P
    T, v, D = 4, 29, 100
    mols = synthetic.make_molecules(40, seed=17)
    db = P.pack_dense_batch(mols, v, D, T)
    A = np.asarray(db["adjacency_matrix"], np.float32)                      # [b, T, v, v], A[g, t, target, source]
    b = A.shape[0]
    lists, indeg = [], np.zeros((b * v, T), np.float32)
    for t in range(T):
        g, i, j = np.nonzero(A[:, t])                                       # row-major: graph, target row, source column
        lists.append(np.stack([g * v + j, g * v + i], 1).astype(np.int32))
        indeg[:, t] = A[:, t].sum(axis=2).reshape(-1)
    p = dict(GRU, hidden_size=D, layer_timesteps=[4], residual_connections={})
    ref = PreparedGraph.host_only(p, T, lists, indeg, precision="bf16x3", save_for_backward=True)
    for nth in (1, 2, 5):
        monkeypatch.setenv("GGNN_HOST_THREADS", str(nth))
        g = PreparedGraph.host_only_dense(p, T, A, precision="bf16x3", save_for_backward=True)
        assert g.info()["plan"] == ref.info()["plan"] + " [binary dense adjacency -> CSR]"
        assert g.info()["num_messages"] == int(A.sum()) and g.info()["num_nodes"] == b * v
        assert np.array_equal(g.image(), ref.image())
    monkeypatch.delenv("GGNN_HOST_THREADS")
    W = A.copy(); W[0, 1, 2, 3] = 0.5                                       # a weighted entry: not this path
    with pytest.raises(GgnnError, match="not 0/1"):
        PreparedGraph.host_only_dense(p, T, W)
