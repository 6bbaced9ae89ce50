"""Seeded synthetic inputs shaped like the reference's data.

The reference's real data comes from ``get_data.py:63-89`` (QM9 via RDKit, needs network); it emits a
list of ``{"targets": [[y]], "graph": [[src, bond(1..4), dst], ...], "node_features": [[one-hot 5], ...]}``
(``get_data.py:82-86``).  ``make_molecules`` emits the same JSON schema from a seeded generator so the
packers (``packing.py``, mirroring ``chem_tensorflow_sparse.py:234-350`` / ``chem_tensorflow_dense.py:132-228``)
see authentic structure: ~18 atoms, a spanning tree plus a few ring closures (~n+0.8 bonds), skewed
bond-type and atom-type distributions.
"""
from __future__ import annotations

import numpy as np

QM9_BOND_PROBS = (0.88, 0.06, 0.02, 0.04)
EIGHT_TYPE_PROBS = (0.5, 0.2, 0.1, 0.06, 0.05, 0.04, 0.03, 0.02)
ATOM_PROBS = (0.51, 0.35, 0.06, 0.07, 0.01)


def make_molecule(rng, bond_probs=QM9_BOND_PROBS, min_atoms=4, max_atoms=29, mean_atoms=18.0, std_atoms=3.0):
    n = int(np.clip(np.rint(rng.normal(mean_atoms, std_atoms)), min_atoms, max_atoms))
    bonds = set()
    for i in range(1, n):  # random spanning tree
        j = int(rng.integers(0, i))
        bonds.add((j, i))
    for _ in range(int(rng.poisson(1.8))):  # ring closures, no duplicates / self loops
        a, b = (int(x) for x in rng.integers(0, n, size=2))
        if a == b:
            continue
        bonds.add((min(a, b), max(a, b)))
    bonds = sorted(bonds)
    types = rng.choice(len(bond_probs), size=len(bonds), p=np.asarray(bond_probs) / np.sum(bond_probs)) + 1
    graph = [[int(a), int(t), int(b)] for (a, b), t in zip(bonds, types)]
    atoms = rng.choice(len(ATOM_PROBS), size=n, p=ATOM_PROBS)
    feats = np.eye(len(ATOM_PROBS), dtype=np.int64)[atoms].tolist()
    return {"targets": [[float(rng.normal())]], "graph": graph, "node_features": feats}


def make_molecules(count, seed=0, num_bond_types=4, **kw):
    """``count`` molecule dicts in the reference JSON schema; bond types 1..num_bond_types."""
    rng = np.random.default_rng(seed)
    probs = QM9_BOND_PROBS if num_bond_types == 4 else EIGHT_TYPE_PROBS[:num_bond_types]
    mols = [make_molecule(rng, bond_probs=probs, **kw) for _ in range(count)]
    # make sure every bond type occurs so num_edge_types (chem_tensorflow.py:116-120) is as requested
    seen = {e[1] for m in mols for e in m["graph"]}
    for t in range(1, num_bond_types + 1):
        if t not in seen:
            mols[(t - 1) % len(mols)]["graph"][0][1] = t
    return mols


def random_sparse_graph(num_nodes, num_undirected_edges, num_edge_types, seed=0):
    """cfg5: ONE graph, uniform random endpoints, no self loops, uniform edge types, both directions
    (tie_fwd_bkwd=True).  Returns the reference wire format: per-type [E_e, 2] int32 lists sorted by
    (src, dst) (sparse:265) and the [V, T] in-degree table (sparse:310-313)."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, num_nodes, size=num_undirected_edges)
    b = rng.integers(0, num_nodes - 1, size=num_undirected_edges)
    b = np.where(b >= a, b + 1, b)  # b != a
    t = rng.integers(0, num_edge_types, size=num_undirected_edges)
    adjacency_lists = []
    indeg = np.zeros((num_nodes, num_edge_types), dtype=np.float32)
    for e in range(num_edge_types):
        m = t == e
        src = np.concatenate([a[m], b[m]])
        dst = np.concatenate([b[m], a[m]])
        order = np.lexsort((dst, src))
        adjacency_lists.append(np.stack([src[order], dst[order]], axis=1).astype(np.int32))
        np.add.at(indeg[:, e], dst, 1.0)
    return adjacency_lists, indeg
