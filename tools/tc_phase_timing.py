"""Prints the per-phase clock64 deltas of tile 0 of the tensor-core kernel (GGNN_TC_DEBUG_TIMING=1)."""
import ctypes as C
import os
import sys

import numpy as np

os.environ["GGNN_TC_DEBUG_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gated_graph_neural_network_samples_b200 import workloads
from gated_graph_neural_network_samples_b200.engine import PropagationEngine

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
w = workloads.build(cfg)
eng = PropagationEngine(w["engine_params"], w["num_edge_types"], precision=prec)
eng.set_weights([{k: torch.from_numpy(v).cuda() for k, v in lw.items()} for lw in w["weights"]])
if w["kind"] == "dense":
    eng.set_graph_dense(w["adjacency_matrix"])
else:
    eng.set_graph_sparse(w["adjacency_lists"], w["num_incoming_edges_per_type"])
h0 = torch.from_numpy(w["h0"]).cuda()
for _ in range(3):
    out = eng.forward(h0)
eng.sync_check()
ts = np.zeros(64, np.int64)
eng._check(eng.lib.ggnn_debug_timestamps(eng._h, ts.ctypes.data))
extra = ts[60:63].copy()
ks = ts[40:60].copy()
print("MMA warp: start-to-start cycles of the first 20 GEMM blocks (7 K-steps each):", (ks[1:] - ks[:-1]).tolist())
ts = ts[:40]
ts = ts[ts > 0]
print(cfg, prec, eng.plan, "stages env", os.environ.get("GGNN_TC_STAGES"))
print("MMA issuer 0: cycles in slow waits for operands %d, total %d" % (extra[1], extra[2]))
print("stamps:", len(ts), "total cycles:", int(ts[-1] - ts[0]))
print("deltas:", (ts[1:] - ts[:-1]).tolist())
