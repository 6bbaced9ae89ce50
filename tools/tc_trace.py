"""Merged timeline of worker thread 0 / MMA issuer 0 / weight producer 0 of tile 0 during the SECOND timestep of the tensor-core kernel
(GGNN_TC_DEBUG_TIMING=1; event codes are the ev()/iev()/pev() calls in csrc/ggnn_fwd_tc.cuh)."""
import os
import sys

import numpy as np

os.environ["GGNN_TC_DEBUG_TIMING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gated_graph_neural_network_samples_b200 import workloads
from gated_graph_neural_network_samples_b200.engine import PropagationEngine

NAMES = {1: "W wait enter", 2: "W poll done", 3: "W bar.sync done", 8: "W step start", 9: "W stamp (phase end)",
         10: "I gemm enter", 11: "I first weights ready", 12: "I gemm issued", 13: "I wait operand", 14: "I operand ready",
         20: "P push begin", 21: "P push end", 22: "P wait xa_free", 23: "P xa_free", 24: "P wait xh_free", 25: "P xh_free"}
cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = workloads.build(cfg)
eng = PropagationEngine(w["engine_params"], w["num_edge_types"], precision="bf16x3")
eng.set_weights([{k: torch.from_numpy(v).cuda() for k, v in lw.items()} for lw in w["weights"]])
eng.set_graph_sparse(w["adjacency_lists"], w["num_incoming_edges_per_type"])
h0 = torch.from_numpy(w["h0"]).cuda()
for _ in range(3):
    eng.forward(h0)
eng.sync_check()
buf = np.zeros(512, np.int64)
eng._check(eng.lib.ggnn_debug_trace(eng._h, buf.ctypes.data, 512))
events = []
for base in (64, 192, 320):
    for i in range(64):
        code, clk = int(buf[base + 2 * i]), int(buf[base + 2 * i + 1])
        if code:
            events.append((clk, code))
events.sort()
t0 = events[0][0] if events else 0
print(eng.plan)
for clk, code in events:
    print("%8d  %s" % (clk - t0, NAMES.get(code, str(code))))
