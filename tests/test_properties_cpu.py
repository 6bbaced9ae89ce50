"""CPU property tests (hypothesis) of the integer/host paths: random multigraphs with self-loops, duplicate edges, empty edge types and
isolated nodes through the CSR build the engine uses, the tile plan, and the flattened packer."""
import ctypes as C

import numpy as np
from hypothesis import given, settings, strategies as st

from gated_graph_neural_network_samples_b200 import _lib, packing
from oracle import ggnn_oracle as O


@st.composite
def multigraphs(draw):
    V = draw(st.integers(1, 60))
    T = draw(st.integers(1, 5))
    adjs = []
    for _ in range(T):
        n = draw(st.integers(0, 80))
        e = draw(st.lists(st.tuples(st.integers(0, V - 1), st.integers(0, V - 1)), min_size=n, max_size=n))
        adjs.append(np.asarray(e, np.int32).reshape(-1, 2))
    return V, adjs


def _ptrs(adjs):
    adjs = [np.ascontiguousarray(a) for a in adjs]
    T = len(adjs)
    return adjs, (C.c_void_p * T)(*[a.ctypes.data for a in adjs]), (C.c_int32 * T)(*[a.shape[0] for a in adjs])


@settings(max_examples=150, deadline=None)
@given(multigraphs())
def test_host_csr_equals_numpy_stable_sort(g):
    V, adjs = g
    lib = _lib.load()
    adjs, ptrs, counts = _ptrs(adjs)
    T, M = len(adjs), sum(a.shape[0] for a in adjs)
    row_ptr, src, msg = np.empty(V * T + 1, np.int32), np.empty(max(M, 1), np.int32), np.empty(max(M, 1), np.int32)
    assert lib.ggnn_host_target_csr(V, T, ptrs, counts, row_ptr.ctypes.data, src.ctypes.data, msg.ctypes.data) == 0
    ref_ptr, ref_src, ref_typ, ref_order = O.stable_target_csr(adjs, V)
    assert np.array_equal(row_ptr[::T], ref_ptr) and np.array_equal(src[:M], ref_src) and np.array_equal(msg[:M], ref_order)
    assert np.array_equal(np.repeat(np.tile(np.arange(T, dtype=np.int32), V), np.diff(row_ptr)), ref_typ)


@settings(max_examples=100, deadline=None)
@given(multigraphs(), st.sampled_from([0, 1]), st.sampled_from([8, 100, 128]))
def test_tile_plan_is_a_partition_that_respects_components(g, precision, D):
    V, adjs = g
    lib = _lib.load()
    adjs, ptrs, counts = _ptrs(adjs)
    ts, n, text = np.empty(V + 2, np.int32), C.c_int32(), C.create_string_buffer(512)
    assert lib.ggnn_host_tile_plan(D, len(adjs), precision, 148, V, ptrs, counts, ts.ctypes.data, V + 2, C.byref(n), text, 512) == 0
    ts = ts[:n.value + 1]
    assert ts[0] == 0 and ts[-1] == V and np.all(np.diff(ts) > 0)
    plan = text.value.decode()
    budget = int(plan.split("rows/tile<=")[1].split()[0])
    assert np.max(np.diff(ts)) <= budget
    if "LOCAL" in plan:
        tile_of = np.searchsorted(ts, np.arange(V), side="right") - 1
        for a in adjs:
            assert np.array_equal(tile_of[a[:, 0]], tile_of[a[:, 1]])


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2 ** 31 - 1), st.integers(1, 25))
def test_flat_packer_equals_per_graph_packer_on_random_subsets(seed, k):
    from gated_graph_neural_network_samples_b200 import synthetic
    rng = np.random.default_rng(seed)
    proc = test_flat_packer_equals_per_graph_packer_on_random_subsets.proc
    flat = test_flat_packer_equals_per_graph_packer_on_random_subsets.flat
    idx = rng.integers(0, len(proc), size=k)                 # with repetition: the same graph may appear twice in a batch
    a = packing.pack_sparse_batch([proc[i] for i in idx], 16, 4)
    b = flat.pack(idx, 16)
    for key in a:
        if key == "adjacency_lists":
            assert all(np.array_equal(x, y) and x.dtype == y.dtype for x, y in zip(a[key], b[key]))
        elif key == "num_graphs":
            assert a[key] == b[key]
        else:
            assert a[key].dtype == b[key].dtype and np.array_equal(a[key], b[key]), key


def _init_flat():
    from gated_graph_neural_network_samples_b200 import synthetic
    proc = packing.process_raw_graphs_sparse(synthetic.make_molecules(60, seed=17))
    test_flat_packer_equals_per_graph_packer_on_random_subsets.proc = proc
    test_flat_packer_equals_per_graph_packer_on_random_subsets.flat = packing.FlatSparseGraphs(proc, 4)


_init_flat()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 300), st.integers(1, 5), st.integers(0, 900), st.integers(0, 2 ** 31 - 1))
def test_stream_gather_tables_against_numpy(V, T, M, seed):
    """The streaming plan's (target, type) -> source table and virtual rows (ggnn_host_stream_tables, the code ggnn_set_graph_sparse uploads)
    against a NumPy restatement: bit-exact, including the message order inside every virtual row and the 128-row tile offsets."""
    import ctypes as C
    from gated_graph_neural_network_samples_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(seed)
    types = rng.integers(0, T, size=M)
    # skewed targets so that many pairs collect several messages
    tgt = np.minimum((rng.random(M) ** 2 * V).astype(np.int64), V - 1)
    src = rng.integers(0, V, size=M)
    adjs = [np.ascontiguousarray(np.stack([src[types == t], tgt[types == t]], 1).astype(np.int32).reshape(-1, 2)) for t in range(T)]
    ptrs = (C.c_void_p * T)(*[a.ctypes.data for a in adjs])
    counts = (C.c_int32 * T)(*[a.shape[0] for a in adjs])
    ntiles = (V + 127) // 128
    pair = np.empty(max(ntiles, 1) * 128 * T, np.int32)
    vptr = np.empty(V * T + 2, np.int32); vsrc = np.empty(M + 1, np.int32); tvp = np.empty(ntiles + 1, np.int32)
    nv = C.c_int32()
    assert lib.ggnn_host_stream_tables(V, T, ptrs, counts, pair.ctypes.data, vptr.ctypes.data, vptr.size, vsrc.ctypes.data, vsrc.size,
                                       tvp.ctypes.data, C.byref(nv)) == 0
    # NumPy restatement: messages in the reference's order (type-major, list order), grouped by (target, type) with a stable sort
    m_src = np.concatenate([a[:, 0] for a in adjs]) if M else np.zeros(0, np.int64)
    m_key = np.concatenate([a[:, 1].astype(np.int64) * T + t for t, a in enumerate(adjs)]) if M else np.zeros(0, np.int64)
    order = np.argsort(m_key, kind="stable")
    keys, starts, cnts = np.unique(m_key[order], return_index=True, return_counts=True)
    want = np.full(ntiles * 128 * T, -1, np.int64)
    want_lists, vid = [], 0
    for k, s0, c in zip(keys, starts, cnts):
        if c == 1:
            want[k] = m_src[order[s0]]
        else:
            want[k] = -(2 + vid); vid += 1
            want_lists.append(m_src[order[s0:s0 + c]])
    assert nv.value == vid
    np.testing.assert_array_equal(pair[:ntiles * 128 * T], want)
    for i, lst in enumerate(want_lists):
        np.testing.assert_array_equal(vsrc[vptr[i]:vptr[i + 1]], lst)
    multi_keys = keys[cnts >= 2]
    for i in range(ntiles + 1):
        assert tvp[i] == int(np.sum(multi_keys < min(i * 128, V) * T))


@settings(max_examples=120, deadline=None)
@given(multigraphs(), st.sampled_from([("fp32", 64), ("bf16x3", 100), ("bf16x3", 256)]), st.sampled_from([1, 2, 3, 7]), st.booleans())
def test_prepared_graph_image_on_random_multigraphs(g, prec_d, threads, save):
    """What the engine uploads for a batch (ggnn_host_prepare_graph_sparse = the host half of ggnn_set_graph_sparse), on multigraphs with
    self-loops, duplicate edges, empty types and isolated nodes, at any builder thread count: CSR == NumPy's stable sort by target in the
    reference's message order (bit for bit), tiles == ggnn_host_tile_plan, and the streaming gather table == its definition."""
    import os
    from gated_graph_neural_network_samples_b200.engine import PreparedGraph
    V, adjs = g
    precision, D = prec_d
    T = len(adjs)
    indeg = np.zeros((V, T), np.float32)
    for t, a in enumerate(adjs):
        np.add.at(indeg[:, t], a[:, 1], 1.0)
    p = {"hidden_size": D, "layer_timesteps": [1], "residual_connections": {}, "use_edge_bias": False, "use_edge_msg_avg_aggregation": True,
         "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}
    old = os.environ.get("GGNN_HOST_THREADS")
    os.environ["GGNN_HOST_THREADS"] = str(threads)
    try:
        pg = PreparedGraph.host_only(p, T, adjs, indeg, precision=precision, save_for_backward=save)
    finally:
        if old is None:
            os.environ.pop("GGNN_HOST_THREADS", None)
        else:
            os.environ["GGNN_HOST_THREADS"] = old
    info, arr = pg.info(), pg.arrays(T)
    M = sum(a.shape[0] for a in adjs)
    ref_ptr, ref_src, ref_typ, ref_order = O.stable_target_csr(adjs, V)
    assert info["num_messages"] == M
    assert np.array_equal(arr["row_ptr"][::T], ref_ptr) and np.array_equal(arr["src"], ref_src) and np.array_equal(arr["msg"], ref_order)
    assert np.array_equal(np.repeat(np.tile(np.arange(T, dtype=np.int32), V), np.diff(arr["row_ptr"])), ref_typ)
    lib = _lib.load()
    adjs_c, ptrs, counts = _ptrs(adjs)
    ts, n, text = np.empty(V + 2, np.int32), C.c_int32(), C.create_string_buffer(512)
    assert lib.ggnn_host_tile_plan(D, T, {"fp32": 0, "bf16x3": 1}[precision], 148, V, ptrs, counts, ts.ctypes.data, V + 2, C.byref(n), text, 512) == 0
    assert np.array_equal(arr["tile_start"], ts[:n.value + 1]) and info["plan"] == text.value.decode()
    assert np.array_equal(arr["denom"], (indeg.sum(axis=1, dtype=np.float32) + np.float32(1e-7)).astype(np.float32)) or T > 2   # fp32 sum order: checked exactly below
    den = np.zeros(V, np.float32)
    for t in range(T):
        den = (den + indeg[:, t]).astype(np.float32)
    assert np.array_equal(arr["denom"], den + np.float32(1e-7))
    if info["streaming"]:
        cnt = np.diff(arr["row_ptr"])
        pair = arr["pair_src"][:V * T]
        assert np.all(pair[cnt == 0] == -1) and np.array_equal(pair[cnt == 1], arr["src"][arr["row_ptr"][:-1][cnt == 1]])
        multi = np.flatnonzero(cnt >= 2)
        assert np.array_equal(pair[multi], -(2 + np.arange(multi.size)))          # virtual rows numbered in (target, type) order
        assert np.all(arr["pair_src"][V * T:] == -1)
