"""Host-side cost of one batch's graph preparation (ggnn_host_prepare_graph_sparse: validation, CSR, tile plan, streaming tables), per
BASELINE workload, on THIS machine's CPU -- no GPU needed.  GGNN_HOST_TIMING=1 prints the builder's own phase laps to stderr."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from gated_graph_neural_network_samples_b200 import workloads  # noqa: E402
from gated_graph_neural_network_samples_b200.engine import PreparedGraph  # noqa: E402

for cfg in sys.argv[1:] or ["cfg2", "cfg4", "cfg5_rgcn", "default_batch_100k_nodes"]:
    w = workloads.build(cfg)
    if w["kind"] not in ("sparse", "single_graph"):
        continue
    g = None
    ts = []
    for i in range(12):
        t0 = time.perf_counter()
        g = PreparedGraph.host_only(w["engine_params"], w["num_edge_types"], w["adjacency_lists"], w["num_incoming_edges_per_type"],
                                    precision="bf16x3", reuse=g)
        ts.append((time.perf_counter() - t0) * 1e6)
    print("%-28s V=%-7d M=%-7d prepare (incl. ctypes marshalling) median %.1f us  min %.1f us   [%s]"
          % (cfg, w["V"], w["M"], float(np.median(ts[2:])), min(ts), g.info()["plan"][:50]))
