"""Helpers shared by the GPU parity tests: drive the C-ABI engine on inputs the oracle also sees."""
import json
import os

import numpy as np

from gated_graph_neural_network_samples_b200 import packing, synthetic

# north_star tolerance: node hidden states within 1e-4 relative (fp32)
RTOL = 1e-4
ATOL = 1e-5


def to_cuda_weights(weights):
    import torch
    out = []
    for w in weights:
        d = {}
        for k, v in w.items():
            key = {"rnn_kernel": "cand_kernel", "rnn_bias": "cand_bias"}.get(k, k)
            d[key] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32)).cuda().contiguous()
        out.append(d)
    return out


def engine_sparse(params, T, weights, adj, indeg, h0, precision="fp32", return_engine=False):
    import torch
    from gated_graph_neural_network_samples_b200.engine import PropagationEngine
    eng = PropagationEngine(params, T, precision=precision)
    eng.set_weights(to_cuda_weights(weights))
    eng.set_graph_sparse(adj, indeg)
    h0_dev = torch.from_numpy(np.ascontiguousarray(h0, dtype=np.float32)).cuda()
    out = eng.forward(h0_dev)
    eng.sync_check()
    res = out.cpu().numpy()
    eng._keepalive = (h0_dev, out)   # node_states_per_layer[0] / [L] are caller-owned buffers the engine only points to
    return (res, eng) if return_engine else res


def dense_params_as_engine_params(dparams, D):
    return {"hidden_size": D, "layer_timesteps": [int(dparams["num_timesteps"])], "residual_connections": {},
            "use_edge_bias": bool(dparams.get("use_edge_bias", True)), "use_edge_msg_avg_aggregation": False,
            "graph_rnn_cell": "GRU", "graph_rnn_activation": "tanh"}


def engine_dense(dparams, T, w, adjm, h0, precision="fp32"):
    import torch
    from gated_graph_neural_network_samples_b200.engine import PropagationEngine
    b, v, D = h0.shape
    params = dense_params_as_engine_params(dparams, D)
    eng = PropagationEngine(params, T, precision=precision)
    w = dict(w)
    if "edge_biases" in w:
        w["edge_biases"] = np.asarray(w["edge_biases"]).reshape(T, D)
    eng.set_weights(to_cuda_weights([w]))
    eng.set_graph_dense(adjm)
    out = eng.forward(torch.from_numpy(np.ascontiguousarray(h0.reshape(b * v, D), dtype=np.float32)).cuda())
    eng.sync_check()
    # the one-call host-buffer entry point (ggnn_run_dense_host) must give the same result
    one_call = eng.run_dense_host(adjm, np.ascontiguousarray(h0.reshape(b * v, D), dtype=np.float32))
    np.testing.assert_allclose(one_call, out.cpu().numpy(), rtol=1e-4, atol=1e-5)   # the tensor path's MMA issue order is not fixed: ~1e-6 noise
    return out.cpu().numpy().reshape(b, v, D)


def molecule_batch(n, D, T=4, seed=0, noise=0.1):
    mols = synthetic.make_molecules(n, seed=seed, num_bond_types=T)
    b = packing.pack_sparse_batch(packing.process_raw_graphs_sparse(mols), D, T)
    if noise:
        rng = np.random.default_rng(seed + 1000)
        b["initial_node_representation"] = (b["initial_node_representation"]
                                            + rng.normal(0, noise, b["initial_node_representation"].shape)).astype(np.float32)
    return mols, b


def load_golden_sparse(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "prop_sparse_%s.npz" % name))
    p = json.loads(str(z["params_json"]))
    L = len(p["layer_timesteps"])
    w = [{k[len("w%d_" % li):]: z[k] for k in z.files if k.startswith("w%d_" % li)} for li in range(L)]
    adj = [z["adj%d" % e] for e in range(4)]
    return z, p, w, adj


def max_rel_err(got, ref):
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref)) / max(np.max(np.abs(ref)), 1e-30))


def seeded_weight(index, shape):
    """tests/golden/make_reference_graph_golden.py::seeded_weight -- the weights of the fixtures too large to commit (D = 256)."""
    r = np.sqrt(6.0 / (shape[-2] + shape[-1]))
    return np.random.RandomState(9000 + index).uniform(-r, r, size=shape).astype(np.float32).astype(np.float64)


def load_refgraph_wide(golden_dir, name):
    """refgraph_sparse_{cfg2,cfg4}_shape.npz: BASELINE-width fixtures computed by the reference's own graph code.  Returns
    (npz, params, per-layer weight dicts (float64 values that are exactly representable in fp32), adjacency lists, T)."""
    from gated_graph_neural_network_samples_b200.engine import residual_inputs_of_layer
    z = np.load(os.path.join(golden_dir, "refgraph_sparse_%s.npz" % name))
    p = json.loads(str(z["params_json"]))
    T, D, L = int(z["num_edge_types"]), int(p["hidden_size"]), len(p["layer_timesteps"])
    if int(z["weights_stored"]):
        w = [{k[len("w%d_" % l):]: np.asarray(z[k], np.float64) for k in z.files if k.startswith("w%d_" % l)} for l in range(L)]
    else:   # the generator's recipe: per layer [edge [T*D, D], cand_bias, cand_kernel, gate_bias, gate_kernel], one index each
        w, idx = [], 0
        for l in range(L):
            din = D * (1 + len(residual_inputs_of_layer(p, l)))
            lw = {"edge_weights": seeded_weight(idx, (T * D, D)).reshape(T, D, D)}
            lw["cand_bias"] = np.asarray(z["w%d_cand_bias" % l], np.float64)
            lw["cand_kernel"] = seeded_weight(idx + 2, (din + D, D))
            lw["gate_bias"] = np.asarray(z["w%d_gate_bias" % l], np.float64)
            lw["gate_kernel"] = seeded_weight(idx + 4, (din + D, 2 * D))
            idx += 5
            w.append(lw)
    return z, p, w, [z["adj%d" % e] for e in range(T)], T
