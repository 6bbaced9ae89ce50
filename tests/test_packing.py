"""packing.py against batches produced by the REFERENCE'S OWN packers (tests/golden/make_golden.py)."""
import json
import os

import numpy as np

import pytest

from gated_graph_neural_network_samples_b200 import packing, synthetic


def _mols(golden_dir):
    with open(os.path.join(golden_dir, "molecules_40.json")) as f:
        return json.load(f)


def test_sparse_batches_match_reference_packer(golden_dir):
    ref = np.load(os.path.join(golden_dir, "packing_sparse.npz"))
    data = packing.process_raw_graphs_sparse(_mols(golden_dir))
    batches = list(packing.iter_sparse_minibatches(data, 200, 8, 4))
    assert len(batches) == int(ref["num_batches"])
    for bi, b in enumerate(batches):
        np.testing.assert_array_equal(b["initial_node_representation"], ref["b%d_init" % bi])
        np.testing.assert_array_equal(b["num_incoming_edges_per_type"], ref["b%d_indeg" % bi])
        np.testing.assert_array_equal(b["graph_nodes_list"], ref["b%d_gnl" % bi])
        assert b["num_graphs"] == int(ref["b%d_num_graphs" % bi])
        np.testing.assert_allclose(b["target_values"], ref["b%d_targets" % bi])
        np.testing.assert_array_equal(b["target_mask"], ref["b%d_mask" % bi])
        for e in range(4):
            got = b["adjacency_lists"][e]
            assert got.dtype == np.int32 and got.shape[1] == 2
            np.testing.assert_array_equal(got, ref["b%d_adj%d" % (bi, e)])


def test_dense_batches_match_reference_packer(golden_dir):
    ref = np.load(os.path.join(golden_dir, "packing_dense.npz"))
    mols = _mols(golden_dir)
    for bi in range(int(ref["num_batches"])):
        idx = ref["b%d_mol_idx" % bi]
        v = int(ref["b%d_num_vertices" % bi])
        b = packing.pack_dense_batch([mols[i] for i in idx], v, 8, 4)
        np.testing.assert_array_equal(b["adjacency_matrix"], ref["b%d_adj" % bi])
        np.testing.assert_array_equal(b["initial_node_representation"], ref["b%d_init" % bi])
        np.testing.assert_array_equal(b["node_mask"], ref["b%d_mask" % bi])
        for i in idx:
            assert ref["bucket_sizes"][packing.choose_bucket(mols[i]["graph"])] == v


def test_empty_edge_type_and_indegree_counts():
    g = [[0, 1, 1], [1, 1, 2], [0, 3, 2]]  # no type-2 / type-4 bonds
    adj, indeg = packing.graph_to_adjacency_lists(g)
    assert sorted(adj) == [0, 2]
    np.testing.assert_array_equal(adj[0], [[0, 1], [1, 0], [1, 2], [2, 1]])
    assert indeg[0] == {0: 1, 1: 2, 2: 1}
    proc = [{"adjacency_lists": adj, "num_incoming_edge_per_type": indeg, "init": [[1, 0]] * 3, "labels": [0.5]}]
    b = packing.pack_sparse_batch(proc, 4, 4)
    assert b["adjacency_lists"][1].shape == (0, 2) and b["adjacency_lists"][3].shape == (0, 2)
    np.testing.assert_array_equal(b["num_incoming_edges_per_type"].sum(0), [4, 0, 2, 0])


def test_flat_packer_equals_the_per_graph_loop():
    """FlatSparseGraphs.pack / iter_minibatches return exactly what pack_sparse_batch / iter_sparse_minibatches return
    (which tests above pin to the reference's own packer): same keys, dtypes, shapes, values -- for shuffled subsets, missing labels,
    absent edge types and single-graph batches."""
    mols = synthetic.make_molecules(300, seed=4, num_bond_types=4)
    proc = packing.process_raw_graphs_sparse(mols, task_ids=(0,))
    proc[3]["labels"][0] = None
    proc[17]["labels"][0] = None
    flat = packing.FlatSparseGraphs(proc, 4)
    rng = np.random.default_rng(0)
    subsets = [rng.permutation(300)[:k] for k in (1, 2, 37, 300)] + [np.array([3, 17])]
    for idx in subsets:
        a = packing.pack_sparse_batch([proc[i] for i in idx], 100, 4)
        b = flat.pack(idx, 100)
        assert sorted(a) == sorted(b)
        for k in a:
            if k == "adjacency_lists":
                for x, y in zip(a[k], b[k]):
                    assert x.dtype == y.dtype and x.shape == y.shape and np.array_equal(x, y)
            elif k == "num_graphs":
                assert a[k] == b[k]
            else:
                assert a[k].dtype == b[k].dtype and a[k].shape == b[k].shape and np.array_equal(a[k], b[k]), k
    order = rng.permutation(300)
    A = list(packing.iter_sparse_minibatches([proc[i] for i in order], 700, 100, 4))
    B = list(flat.iter_minibatches(order, 700, 100))
    assert len(A) == len(B) > 3
    for x, y in zip(A, B):
        assert x["num_graphs"] == y["num_graphs"]
        np.testing.assert_array_equal(x["graph_nodes_list"], y["graph_nodes_list"])
        np.testing.assert_array_equal(x["adjacency_lists"][0], y["adjacency_lists"][0])
        np.testing.assert_array_equal(x["target_mask"], y["target_mask"])
    with pytest.raises(Exception, match="does not fit"):
        list(flat.iter_minibatches(order, 5, 100))


def test_sparse_plugin_iterator_equals_reference_style_packing_across_epochs():
    """SparseGGNNChemModel.make_minibatch_iterator (host logic only; built without an engine) over a list that is shuffled in place
    every epoch, copied, and shortened: every feed equals the per-graph packer's batch of the same graphs."""
    from gated_graph_neural_network_samples_b200.chem_sparse import SparseGGNNChemModel
    m = object.__new__(SparseGGNNChemModel)
    m.params = {"batch_size": 600, "hidden_size": 100, "graph_state_dropout_keep_prob": 1.0, "edge_weight_dropout_keep_prob": 0.8}
    m.num_edge_types = 4
    m.placeholders = {"adjacency_lists": ["adjacency_e%d" % e for e in range(4)]}
    data = packing.process_raw_graphs_sparse(synthetic.make_molecules(150, seed=9), task_ids=(0,))

    def check(lst, training):
        state = np.random.get_state()
        feeds = list(m.make_minibatch_iterator(lst, training))          # shuffles lst in place when training
        np.random.set_state(state)
        ref = list(packing.iter_sparse_minibatches(lst, 600, 100, 4))   # same (already shuffled) order
        assert len(feeds) == len(ref) > 2
        for f, r in zip(feeds, ref):
            np.testing.assert_array_equal(f["initial_node_representation"], r["initial_node_representation"])
            np.testing.assert_array_equal(f["graph_nodes_list"], r["graph_nodes_list"])
            np.testing.assert_array_equal(f["target_values"], r["target_values"])
            for e in range(4):
                np.testing.assert_array_equal(f["adjacency_e%d" % e], r["adjacency_lists"][e])
            assert f["edge_weight_dropout_keep_prob"] == (0.8 if training else 1.0)

    np.random.seed(3)
    for _ in range(3):
        check(data, True)          # epoch after epoch on the same list object (cache hit, new order)
    check(list(data), False)       # a copy of the list: rebuilt, the original's cache entry stays valid
    check(data, True)
    del data[10:40]                # membership changed: rebuilt
    check(data, False)
    assert len(m._flat_cache) <= 4
