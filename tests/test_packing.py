"""packing.py against batches produced by the REFERENCE'S OWN packers (tests/golden/make_golden.py)."""
import json
import os

import numpy as np

from gated_graph_neural_network_samples_b200 import packing


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
