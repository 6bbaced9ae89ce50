"""Host-side batch packing: raw molecule dicts -> the wire format the propagation engine consumes.

Mirrors (behaviour, not code) the reference packers that sit just before the hot path:

* sparse: ``chem_tensorflow_sparse.py:234-276`` (``process_raw_graphs`` / ``__graph_to_adjacency_lists``)
  and ``:278-350`` (``make_minibatch_iterator``): per edge type an ``[E_e, 2]`` int32 list of
  ``(source, target)`` sorted by (source, target), both directions when ``tie_fwd_bkwd``; per node and
  type the in-degree (multi-edges counted); graphs concatenated into one disconnected batch with node
  offsets until the *node* budget ``batch_size`` would be reached (strict ``<``, sparse:297).
* dense: ``chem_tensorflow_dense.py:30-36,132-228``: ``amat[e, dest, src] = 1`` (assignment, duplicates
  collapse), features and mask padded to the bucket size.

Outputs are keyed like the reference's ``self.placeholders`` feed-dict slots (sparse:331-348,
dense:214-224).  tests/test_packing.py checks them against batches produced by the reference's own
NumPy code (tests/golden/make_golden.py).
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Sequence

import numpy as np


# ------------------------------------------------------------------------------------------ sparse
def graph_to_adjacency_lists(graph: Sequence[Sequence[int]], tie_fwd_bkwd: bool = True):
    """One molecule's ``[[src, bond(1..), dst], ...]`` -> ({type: [E,2] int32 sorted}, {type: {node: indeg}}).

    Only ``tie_fwd_bkwd=True`` is supported: the reference's untied branch indexes past
    ``num_edge_types`` (sparse:271 vs chem_tensorflow.py:120; SURVEY 8a "latent bugs")."""
    if not tie_fwd_bkwd:
        raise Exception("tie_fwd_bkwd=False is broken in the reference sparse model and not supported")
    g = np.asarray(graph, dtype=np.int64).reshape(-1, 3)
    adj: Dict[int, np.ndarray] = {}
    indeg: Dict[int, Dict[int, int]] = {}
    for t in np.unique(g[:, 1]):
        rows = g[g[:, 1] == t]
        src = np.concatenate([rows[:, 0], rows[:, 2]])
        dst = np.concatenate([rows[:, 2], rows[:, 0]])
        order = np.lexsort((dst, src))  # == sorted(list of (src, dst) tuples), sparse:265
        e = int(t) - 1                  # sparse:258
        adj[e] = np.stack([src[order], dst[order]], axis=1).astype(np.int32)
        nodes, counts = np.unique(dst, return_counts=True)
        indeg[e] = {int(n): int(c) for n, c in zip(nodes, counts)}
    return adj, indeg


def process_raw_graphs_sparse(raw_data: Iterable[dict], task_ids=(0,), tie_fwd_bkwd: bool = True) -> List[dict]:
    """sparse:234-252 without the training-time shuffle / task sub-sampling (caller's business)."""
    out = []
    for d in raw_data:
        adj, indeg = graph_to_adjacency_lists(d["graph"], tie_fwd_bkwd)
        out.append({"adjacency_lists": adj, "num_incoming_edge_per_type": indeg, "init": d["node_features"],
                    "labels": [d["targets"][t][0] for t in task_ids]})
    return out


def pack_sparse_batch(graphs: Sequence[dict], hidden_size: int, num_edge_types: int) -> dict:
    """Concatenate processed graphs into one disconnected batch (sparse:288-350)."""
    feats, gnl, indeg_rows, tv, tm = [], [], [], [], []
    per_type: List[List[np.ndarray]] = [[] for _ in range(num_edge_types)]
    offset = 0
    for gi, g in enumerate(graphs):
        init = np.asarray(g["init"], dtype=np.float32)
        n, ann = init.shape
        padded = np.zeros((n, hidden_size), dtype=np.float32)                  # sparse:300-302
        padded[:, :ann] = init
        feats.append(padded)
        gnl.append(np.full(n, gi, dtype=np.int32))                             # sparse:304
        for e in range(num_edge_types):                                        # sparse:305-307
            a = g["adjacency_lists"].get(e)
            if a is not None:
                per_type[e].append(a + np.int32(offset))
        deg = np.zeros((n, num_edge_types), dtype=np.float32)                  # sparse:310-313
        for e, dct in g["num_incoming_edge_per_type"].items():
            for node, cnt in dct.items():
                deg[node, e] = cnt
        indeg_rows.append(deg)
        tv.append([0.0 if v is None else v for v in g["labels"]])              # sparse:316-326
        tm.append([0.0 if v is None else 1.0 for v in g["labels"]])
        offset += n
    adjacency_lists = [np.concatenate(l).astype(np.int32) if l else np.zeros((0, 2), np.int32)  # sparse:343-347
                       for l in per_type]
    return {
        "initial_node_representation": np.concatenate(feats, axis=0) if feats else np.zeros((0, hidden_size), np.float32),
        "adjacency_lists": adjacency_lists,
        "num_incoming_edges_per_type": np.concatenate(indeg_rows, axis=0) if indeg_rows
        else np.zeros((0, num_edge_types), np.float32),
        "graph_nodes_list": np.concatenate(gnl) if gnl else np.zeros(0, np.int32),
        "target_values": np.asarray(tv, dtype=np.float32).T.reshape(-1, len(graphs)),
        "target_mask": np.asarray(tm, dtype=np.float32).T.reshape(-1, len(graphs)),
        "num_graphs": len(graphs),
    }


def iter_sparse_minibatches(data: Sequence[dict], batch_size_nodes: int, hidden_size: int, num_edge_types: int):
    """sparse:286-350: greedy packing while ``node_offset + len(graph) < batch_size`` (strict)."""
    i = 0
    while i < len(data):
        start, nodes = i, 0
        while i < len(data) and nodes + len(data[i]["init"]) < batch_size_nodes:
            nodes += len(data[i]["init"])
            i += 1
        if i == start:
            raise Exception("graph %d has %d nodes and does not fit batch_size=%d"
                            % (i, len(data[i]["init"]), batch_size_nodes))  # the reference loops forever here
        yield pack_sparse_batch(data[start:i], hidden_size, num_edge_types)


# ------------------------------------------------------------------------------------------- dense
DEFAULT_BUCKET_SIZES = np.array(list(range(4, 28, 2)) + [29])  # dense:134


def graph_to_adj_mat(graph, max_n_vertices: int, num_edge_types: int, tie_fwd_bkwd: bool = True) -> np.ndarray:
    """dense:30-36 -- [T, v, v] with amat[e-1, dest, src] = 1 and the tied reverse entry."""
    g = np.asarray(graph, dtype=np.int64).reshape(-1, 3)
    bwd = 0 if tie_fwd_bkwd else num_edge_types // 2
    amat = np.zeros((num_edge_types, max_n_vertices, max_n_vertices), dtype=np.float32)
    amat[g[:, 1] - 1, g[:, 2], g[:, 0]] = 1.0
    amat[g[:, 1] - 1 + bwd, g[:, 0], g[:, 2]] = 1.0
    return amat


def pack_dense_batch(raw_graphs: Sequence[dict], bucket_size: int, hidden_size: int, num_edge_types: int,
                     task_ids=(0,), tie_fwd_bkwd: bool = True) -> dict:
    """dense:142-148,172-224 for one bucket: [b,T,v,v] adjacency, [b,v,D] features, [b,v] mask."""
    b = len(raw_graphs)
    adj = np.zeros((b, num_edge_types, bucket_size, bucket_size), dtype=np.float32)
    init = np.zeros((b, bucket_size, hidden_size), dtype=np.float32)
    mask = np.zeros((b, bucket_size), dtype=np.float32)
    tv, tm = [], []
    for i, d in enumerate(raw_graphs):
        adj[i] = graph_to_adj_mat(d["graph"], bucket_size, num_edge_types, tie_fwd_bkwd)
        f = np.asarray(d["node_features"], dtype=np.float32)
        init[i, :f.shape[0], :f.shape[1]] = f
        mask[i, :f.shape[0]] = 1.0
        labels = [d["targets"][t][0] for t in task_ids]
        tv.append([0.0 if v is None else v for v in labels])
        tm.append([0.0 if v is None else 1.0 for v in labels])
    return {"initial_node_representation": init, "adjacency_matrix": adj, "node_mask": mask,
            "num_vertices": int(bucket_size), "num_graphs": b,
            "target_values": np.asarray(tv, np.float32).T.reshape(-1, b),
            "target_mask": np.asarray(tm, np.float32).T.reshape(-1, b)}


def choose_bucket(graph, bucket_sizes=DEFAULT_BUCKET_SIZES) -> int:
    """dense:138-140 -- first bucket strictly larger than the largest node id."""
    mx = max(v for e in graph for v in (e[0], e[2]))
    return int(np.argmax(np.asarray(bucket_sizes) > mx))
