"""CPU-only: the C-ABI library builds for sm_100a, loads, and exports every symbol include/ggnn_b200.h declares
(no compute calls without a GPU); host-side error behaviour that does not need a device."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ggnn_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ggnn_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from gated_graph_neural_network_samples_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert name in _lib.SYMBOLS, "header symbol %s has no ctypes binding" % name
        assert getattr(lib, name) is not None
    assert sorted(_lib.SYMBOLS) == declared


def test_create_fails_loudly_without_a_gpu_or_with_bad_config():
    """No silent CPU fallback: without a CUDA device the engine refuses to exist."""
    import torch
    from gated_graph_neural_network_samples_b200.engine import GgnnError, PropagationEngine
    params = {"hidden_size": 8, "layer_timesteps": [1], "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}
    with pytest.raises(Exception, match="Unknown activation"):
        PropagationEngine(dict(params, graph_rnn_activation="swish"), 2)
    with pytest.raises(GgnnError, match="multiple of 4"):
        PropagationEngine(dict(params, hidden_size=6), 2)
    with pytest.raises(GgnnError, match="residual connection"):
        PropagationEngine(dict(params, residual_connections={"0": [1]}), 2)
    if not torch.cuda.is_available():
        with pytest.raises(GgnnError, match="CUDA device unavailable"):
            PropagationEngine(params, 2)


def test_workload_definitions_and_algorithmic_bytes():
    from gated_graph_neural_network_samples_b200 import workloads
    w = workloads.build("cfg2")
    assert w["V"] > 4000 and w["timesteps"] == 4 and w["num_edge_types"] == 4
    D, V, M, T = 100, w["V"], w["M"], 4
    step = 4 * D * (2 * V + M) + 8 * M + 4 * V * T + 4 * (T * D * D + 2 * D * 3 * D + 3 * D)
    assert workloads.algorithmic_bytes(w) == 4 * step
    assert workloads.algorithmic_flops(w) == 4 * (2 * M * D * D + 2 * V * 2 * D * 3 * D)
    w5 = workloads.build("cfg5_rgcn")
    assert w5["V"] == 10000 and w5["M"] == 80000 and "gate_kernel" not in w5["weights"][0]


def test_state_dropout_mask_host_restatement_matches_oracle():
    """ggnn_state_dropout_mask (host arithmetic only, no GPU) == the oracle's NumPy restatement of the counter hash."""
    import ctypes as C
    from gated_graph_neural_network_samples_b200 import _lib
    from oracle import ggnn_oracle as O
    lib = _lib.load()
    for V, D, step, keep, seed in [(37, 100, 0, 0.8, 0), (5, 8, 3, 0.5, 12345678901234), (64, 128, 11, 0.9, 2 ** 62 - 1), (0, 4, 0, 0.5, 1)]:
        m = np.empty((V, D), np.uint8)
        assert lib.ggnn_state_dropout_mask(V, D, step, C.c_float(keep), C.c_uint64(seed), m.ctypes.data) == 0
        np.testing.assert_array_equal(m.astype(bool), O.state_dropout_mask(seed, step, V, D, keep))
        if V * D > 1000:
            assert abs(m.mean() - keep) < 0.03


def test_host_csr_build_is_numpys_stable_sort_bit_for_bit():
    """Integer path without a GPU: ggnn_host_target_csr (the CSR fill ggnn_set_graph_sparse uses) == NumPy's stable argsort of the
    type-major message list by target -- row offsets, sources and original message ids; empty edge types, isolated nodes,
    multi-edges, self-loops, an empty batch; an out-of-range edge is refused like TF's CPU gather does."""
    import ctypes as C
    from gated_graph_neural_network_samples_b200 import _lib, packing, synthetic
    from oracle import ggnn_oracle as O
    lib = _lib.load()

    def host_csr(adjs, V):
        T = len(adjs)
        adjs = [np.ascontiguousarray(np.asarray(a, np.int32).reshape(-1, 2)) for a in adjs]
        M = sum(a.shape[0] for a in adjs)
        ptrs = (C.c_void_p * T)(*[a.ctypes.data for a in adjs])
        counts = (C.c_int32 * T)(*[a.shape[0] for a in adjs])
        row_ptr, src, msg = np.empty(V * T + 1, np.int32), np.empty(max(M, 1), np.int32), np.empty(max(M, 1), np.int32)
        rc = lib.ggnn_host_target_csr(V, T, ptrs, counts, row_ptr.ctypes.data, src.ctypes.data, msg.ctypes.data)
        return rc, row_ptr, src[:M], msg[:M]

    cases = []
    for seed, n, T in [(13, 50, 4), (2, 7, 8), (5, 300, 4)]:
        mols = synthetic.make_molecules(n, seed=seed, num_bond_types=T)
        b = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), 8, T)
        cases.append((b["adjacency_lists"], b["initial_node_representation"].shape[0]))
    cases.append(([np.array([[0, 1], [0, 1], [2, 2], [4, 1]], np.int32), np.zeros((0, 2), np.int32), np.array([[1, 0]], np.int32)], 6))
    cases.append(([np.zeros((0, 2), np.int32)], 0))
    for adjs, V in cases:
        T = len(adjs)
        rc, row_ptr, src, msg = host_csr(adjs, V)
        assert rc == 0
        ref_ptr, ref_src, ref_typ, ref_order = O.stable_target_csr(adjs, V)
        np.testing.assert_array_equal(row_ptr[::T], ref_ptr)
        np.testing.assert_array_equal(src, ref_src)
        np.testing.assert_array_equal(msg, ref_order)
        # rows are keyed target*T + type: the per-row type of every slot follows from row_ptr
        typ = np.repeat(np.tile(np.arange(T, dtype=np.int32), V), np.diff(row_ptr))
        np.testing.assert_array_equal(typ, ref_typ)
    rc, *_ = host_csr([np.array([[0, 3]], np.int32)], 3)
    assert rc == -5                                            # GGNN_ERANGE


def test_host_tile_plan_keeps_components_whole_and_fills_the_sms():
    """The tiling logic of ggnn_set_graph_sparse without a GPU (ggnn_host_tile_plan): tiles partition [0, V) in order, no edge crosses
    a tile in a LOCAL plan, no tile exceeds its row budget, a batch that cannot fill the chip is cut into <= num_sms smaller tiles (and
    <= 64-row tiles select the compact operand layout), a component larger than a tile switches to the GLOBAL plan."""
    import ctypes as C
    from gated_graph_neural_network_samples_b200 import _lib, packing, synthetic
    lib = _lib.load()

    def plan(adjs, V, D=100, precision=1, sms=148):
        T = len(adjs)
        adjs = [np.ascontiguousarray(np.asarray(a, np.int32).reshape(-1, 2)) for a in adjs]
        ptrs = (C.c_void_p * T)(*[a.ctypes.data for a in adjs])
        counts = (C.c_int32 * T)(*[a.shape[0] for a in adjs])
        ts = np.empty(V + 2, np.int32)
        n = C.c_int32()
        text = C.create_string_buffer(512)
        rc = lib.ggnn_host_tile_plan(D, T, precision, sms, V, ptrs, counts, ts.ctypes.data, V + 2, C.byref(n), text, 512)
        assert rc == 0
        return ts[:n.value + 1].copy(), text.value.decode()

    def check_partition(ts, V, adjs, local):
        assert ts[0] == 0 and ts[-1] == V and np.all(np.diff(ts) > 0)
        if local:
            tile_of = np.searchsorted(ts, np.arange(V), side="right") - 1
            for a in adjs:
                a = np.asarray(a).reshape(-1, 2)
                assert np.array_equal(tile_of[a[:, 0]], tile_of[a[:, 1]])       # every edge stays inside one tile

    for n_mols, expect_compact in [(256, True), (5500, False)]:
        mols = synthetic.make_molecules(n_mols, seed=0)
        b = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), 8, 4)
        V = b["initial_node_representation"].shape[0]
        for precision in (1, 0):                                               # bf16x3 (tcgen05 plan), fp32 (FFMA plan)
            ts, text = plan(b["adjacency_lists"], V, precision=precision)
            assert "LOCAL" in text, text
            check_partition(ts, V, b["adjacency_lists"], True)
            budget = int(text.split("rows/tile<=")[1].split()[0])
            assert np.max(np.diff(ts)) <= budget <= 128
            if precision == 1:
                assert ("compact" in text) == expect_compact, text
                if expect_compact:
                    assert len(ts) - 1 <= 148 and budget <= 64                 # cut small enough to use (almost) every SM
    # one 300-node ring: larger than any tile -> fixed 128-row tiles; tensor precisions take the streaming plan, fp32 one launch per step
    ring = np.stack([np.arange(300), (np.arange(300) + 1) % 300], 1).astype(np.int32)
    ts, text = plan([ring], 300)
    assert "STREAM" in text and list(ts) == [0, 128, 256, 300]
    ts, text = plan([ring], 300, precision=0)
    assert "GLOBAL" in text
    # hidden sizes above 128 stream whatever the component sizes, with N blocks of at least 128 columns
    mols = synthetic.make_molecules(64, seed=1, num_bond_types=8)
    b = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), 256, 8)
    V = b["initial_node_representation"].shape[0]
    ts, text = plan(b["adjacency_lists"], V, D=256)
    assert "STREAM" in text and list(ts[:-1]) == list(range(0, V, 128)) and ts[-1] == V
    assert "agg/cand=2x128" in text and "gate=4x128" in text, text          # 10 tiles: N is split so that more SMs get work
    ts, text = plan([np.zeros((0, 2), np.int32)], 0)
    assert list(ts) == [0]
