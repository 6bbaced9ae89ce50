"""ctypes face of oracle/ggnn_oracle.c (the plain-C float64 restatement) -- TEST INFRASTRUCTURE ONLY.

``build()`` compiles it with gcc into ``oracle/libggnn_oracle_c.so`` (git-ignored, travels to the GPU box with the snapshot).
The weight dictionaries are the ones ``ggnn_oracle.init_sparse_weights`` / ``init_dense_weights`` produce."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "ggnn_oracle.c")
LIB = os.path.join(HERE, "libggnn_oracle_c.so")
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


class _Config(C.Structure):
    _fields_ = [("hidden_size", C.c_int32), ("num_edge_types", C.c_int32), ("num_layers", C.c_int32), ("layer_timesteps", _i32p),
                ("residual_offsets", _i32p), ("residual_layers", _i32p), ("use_edge_bias", C.c_int32),
                ("use_edge_msg_avg_aggregation", C.c_int32), ("use_propagation_attention", C.c_int32), ("cell_is_rnn", C.c_int32),
                ("act_is_relu", C.c_int32)]


class _Layer(C.Structure):
    _fields_ = [(n, _f64p) for n in ("edge_weights", "edge_biases", "edge_type_attention_weights", "gate_kernel", "gate_bias",
                                     "cand_kernel", "cand_bias", "cand_hidden_bias")]


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        tmp = LIB + ".tmp"
        subprocess.run(["gcc", "-O2", "-std=c99", "-shared", "-fPIC", "-o", tmp, SRC, "-lm"], check=True)
        os.replace(tmp, LIB)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.ggnn_oracle_sparse.restype = C.c_int
        _lib.ggnn_oracle_dense.restype = C.c_int
    return _lib


def _d(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _layer(w: dict, keep: list) -> _Layer:
    ren = {"rnn_kernel": "cand_kernel", "rnn_bias": "cand_bias"}
    lay = _Layer()
    for k, v in w.items():
        a = _d(v)
        keep.append(a)
        setattr(lay, ren.get(k, k), a.ctypes.data_as(_f64p))
    return lay


def sparse_propagation_c(h0, adjacency_lists, num_incoming_edges_per_type, weights, params) -> np.ndarray:
    """chem_tensorflow_sparse.py:117-218 through the C restatement (float64)."""
    lib = _load()
    keep = []
    h0 = _d(h0)
    V, D = h0.shape
    T = len(adjacency_lists)
    steps = np.ascontiguousarray(params["layer_timesteps"], dtype=np.int32)
    L = steps.shape[0]
    offs, flat = [0], []
    for l in range(L):
        flat += [int(x) for x in ((params.get("residual_connections") or {}).get(str(l)) or [])]
        offs.append(len(flat))
    offs = np.asarray(offs, np.int32)
    flat = np.asarray(flat + [0], np.int32)
    cfg = _Config(D, T, L, steps.ctypes.data_as(_i32p), offs.ctypes.data_as(_i32p), flat.ctypes.data_as(_i32p),
                  int(bool(params.get("use_edge_bias", False))), int(bool(params.get("use_edge_msg_avg_aggregation", False))),
                  int(bool(params.get("use_propagation_attention", False))),
                  {"gru": 0, "rnn": 1, "cudnncompatiblegrucell": 2}[params.get("graph_rnn_cell", "GRU").lower()], int(params.get("graph_rnn_activation", "tanh").lower() == "relu"))
    layers = (_Layer * L)(*[_layer(w, keep) for w in weights])
    adjs = [np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1, 2)) for a in adjacency_lists]
    ptrs = (_i32p * T)(*[a.ctypes.data_as(_i32p) for a in adjs])
    counts = np.asarray([a.shape[0] for a in adjs], np.int32)
    indeg = _d(num_incoming_edges_per_type)
    out = np.empty_like(h0)
    rc = lib.ggnn_oracle_sparse(C.byref(cfg), layers, C.c_int32(V), ptrs, counts.ctypes.data_as(_i32p), indeg.ctypes.data_as(_f64p),
                                h0.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p))
    if rc != 0:
        raise IndexError("edge index out of range (or allocation failure) in the C oracle")
    return out


def dense_propagation_c(h0, adjacency_matrix, weights, params) -> np.ndarray:
    """chem_tensorflow_dense.py:93-117 through the C restatement (float64).  h0 [b, v, D], adjacency [b, T, v, v]."""
    lib = _load()
    keep = []
    h0 = _d(h0)
    b, v, D = h0.shape
    A = _d(adjacency_matrix)
    T = A.shape[1]
    w = dict(weights)
    if "edge_biases" in w:
        w["edge_biases"] = np.asarray(w["edge_biases"]).reshape(T, D)
    lay = _layer(w, keep)
    out = np.empty_like(h0)
    rc = lib.ggnn_oracle_dense(C.c_int32(D), C.c_int32(T), C.c_int32(int(params["num_timesteps"])), C.c_int32(int(bool(params.get("use_edge_bias", True)))),
                               C.byref(lay), C.c_int32(b), C.c_int32(v), A.ctypes.data_as(_f64p), h0.ctypes.data_as(_f64p), out.ctypes.data_as(_f64p))
    if rc != 0:
        raise MemoryError("C oracle allocation failure")
    return out
