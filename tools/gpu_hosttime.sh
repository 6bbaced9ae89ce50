#!/bin/bash
OUT=gpurun_out/ht; mkdir -p $OUT
GGNN_HOST_TIMING=1 timeout 300 python - > $OUT/host_timing.txt 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from gated_graph_neural_network_samples_b200 import workloads
from gated_graph_neural_network_samples_b200.engine import PropagationEngine
for cfg in ("cfg2", "cfg4"):
    w = workloads.build(cfg)
    eng = PropagationEngine(w["engine_params"], w["num_edge_types"], precision="bf16x3")
    eng.set_weights([{k: torch.from_numpy(v).cuda() for k, v in lw.items()} for lw in w["weights"]])
    h0 = torch.from_numpy(w["h0"]).pin_memory(); out = torch.empty_like(h0).pin_memory()
    for i in range(6):
        t0 = time.perf_counter()
        eng.run_sparse_host(w["adjacency_lists"], w["num_incoming_edges_per_type"], h0.numpy(), out.numpy())
        print("== %s run_sparse_host call %d: %.1f us" % (cfg, i, (time.perf_counter() - t0) * 1e6), file=sys.stderr)
PY
tail -40 $OUT/host_timing.txt
