"""Per-launch phase breakdown of the streaming tcgen05 kernels (GGNN_TS_DEBUG=1): for every launch of one forward, the median over
CTAs of  gather done / accumulator ready / CTA end  (cycles from CTA start), the producer's wait for free stages and the issuer's
waits for B (TMA) and gathered A."""
import os
import sys

import numpy as np

os.environ["GGNN_TS_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gated_graph_neural_network_samples_b200 import workloads
from gated_graph_neural_network_samples_b200.engine import PropagationEngine

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
shard = os.environ.get("GGNN_BENCH_SHARD")
w = workloads.build(cfg, shard=tuple(int(x) for x in shard.split(",")) if shard else None)
eng = PropagationEngine(w["engine_params"], w["num_edge_types"], precision="bf16x3")
eng.set_weights([{k: torch.from_numpy(v).cuda() for k, v in lw.items()} for lw in w["weights"]])
eng.set_graph_sparse(w["adjacency_lists"], w["num_incoming_edges_per_type"])
h0 = torch.from_numpy(w["h0"]).cuda()
for _ in range(3):
    eng.forward(h0)
eng.sync_check()
print(eng.plan)
plan = eng.plan
ntiles = int(plan.split("tiles=")[1].split()[0])
nb = [int(x.split("x")[0]) for x in (plan.split("agg/cand=")[1].split()[0], plan.split("gate=")[1].split()[0])]
slice_ = ntiles * max(nb) * 16
steps = sum(w["engine_params"]["layer_timesteps"])
gru = w["engine_params"].get("graph_rnn_cell", "GRU").lower() == "gru"
per_step = 3 if gru else 2
stride = slice_ + 2048
n = stride * per_step * steps
buf = np.zeros(n, np.int64)
eng._check(eng.lib.ggnn_debug_trace(eng._h, buf.ctypes.data, n))
names = ["edge", "gate", "cand"] if gru else ["edge", "cand"]
print("%-6s %5s %9s %9s %9s %9s | %9s %9s %9s  (median cycles over CTAs; start spread = max-min of CTA start clocks on one SM clock domain is not comparable)" %
      ("launch", "nk", "gather", "acc", "end", "epilogue", "P wait", "I wait B", "I wait A"))
for li in range(per_step * steps):
    kind = names[li % per_step]
    ctas = ntiles * (nb[1] if kind == "gate" else nb[0])
    if kind == "gate" and ntiles * nb[1] > 148 and nb[1] % 2 == 0:
        ctas //= 2          # two N blocks per CTA (npass = 2)
    d = buf[li * stride: li * stride + ctas * 16].reshape(ctas, 16)
    med = lambda c: int(np.median(d[:, c]))
    print("%-6s %5d %9d %9d %9d %9d | %9d %9d %9d | setup %6d load %7d wait %7d tail %7d" % (kind, med(7), med(1) if kind == "edge" else 0, med(2), med(3), med(3) - med(2), med(4), med(5), med(6), med(11), med(8), med(9), med(10)) + ("  [setup: pair table %d, virtual rows %d (nv %d), fence %d]" % (med(12), med(13), med(15), med(14)) if kind == "edge" else ""))

# K-step timeline of CTA (0,0) for the first launch of each kind (clocks relative to the CTA's first stamp)
for li in range(per_step):
    tl = buf[li * stride + slice_: li * stride + slice_ + 2048].reshape(256, 8)
    nz = tl[tl > 0]
    if nz.size == 0:
        continue
    t0 = nz.min()
    print("\n%s launch, CTA (0,0): k | issuer: B landed, A landed, issued | gather: start, stage free, copies issued | producer: stage free, TMA issued" % names[li])
    for k in list(range(0, 24)) + list(range(64, 72)):
        r = tl[k]
        if not r.any():
            continue
        print("%4d | %7d %7d %7d | %7d %7d %7d | %7d %7d" % tuple([k] + [int(x - t0) if x > 0 else -1 for x in r]))
