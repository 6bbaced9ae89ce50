"""One forward(save)+backward of a workload, for an ncu launch list of the training propagation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from gated_graph_neural_network_samples_b200 import workloads
from gated_graph_neural_network_samples_b200.engine import PropagationEngine

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = workloads.build(cfg, seed=0)
P = w["engine_params"]
eng = PropagationEngine(P, w["num_edge_types"], precision="bf16x3" if P["hidden_size"] <= 128 else "fp32")
dev_w = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in lw.items()} for lw in w["weights"]]
eng.set_weights(dev_w)
eng.set_save_for_backward(True)
if w["kind"] == "dense":
    eng.set_graph_dense(w["adjacency_matrix"])
else:
    eng.set_graph_sparse(w["adjacency_lists"], w["num_incoming_edges_per_type"])
h0 = torch.from_numpy(w["h0"]).cuda()
out = torch.empty_like(h0)
grads = [{k: torch.zeros_like(v) for k, v in lw.items()} for lw in dev_w]
d_out = torch.ones_like(h0); d_h0 = torch.empty_like(h0)
for _ in range(3):
    eng.forward(h0, out); eng.backward(d_out, grads, d_h0)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.forward(h0, out); eng.backward(d_out, grads, d_h0)
e1.record(); torch.cuda.synchronize()
print("train ms/step", e0.elapsed_time(e1) / 10, "launches/step", eng.last_launch_count)
