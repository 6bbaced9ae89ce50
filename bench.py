#!/usr/bin/env python
"""Benchmark of the GGNN propagation step (BASELINE.json metric: node-state-updates/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl ours|reference]

A "step" is one pass of ``compute_final_node_representations`` over one batch of synthetic molecules.
Default workload = BASELINE.json configs[1] ("cfg2": sparse GGNN, hidden=100, 4 edge types, 4 timesteps,
256 molecules, one B200).  Prints ONE JSON line (rank 0).

* ``value``      : node-state updates / s with graph + states + weights already resident in HBM, timed with
                   CUDA events around every step (L2 flushed before each step, flush not timed), max over ranks.
* ``e2e``        : the same metric through the public one-call host-buffer API (``run_sparse_host`` / ``run_dense_host``):
                   graph arrays and node states start in (pinned) HOST memory every step, result read back, serial.
* ``e2e_pipelined``: the same calls with two batches in flight (two engines, two streams) -- reported beside ``e2e``, not instead.
* ``train_propagation``: forward with saved states + backward of the propagation, device-resident (SURVEY 8d secondary metric).
* ``readout``    : the fused gated-regression readout against the same op sequence as torch kernels.
* ``roofline``   : algorithmic bytes of the dominant kernel / its CUDA-event duration vs the measured HBM peak.
* ``cpu_baseline``: the fp32 PyTorch-CPU restatement of the TF1 graph (oracle/; TF 1.3 is not installable)
                   on this box's host cores, bounded sample.
``--impl reference`` times that CPU restatement alone (the reference arm).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", default="cfg2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=os.environ.get("GGNN_PRECISION", "auto"),
                    help="auto = bf16x3 (tcgen05, fp32-accurate hi/lo split, within the 1e-4 parity bar)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    return ap.parse_args()


def oracle_weights(w):
    cell = w["engine_params"].get("graph_rnn_cell", "GRU").lower()
    out = []
    for lw in w["weights"]:
        d = dict(lw)
        if cell == "rnn":
            d["rnn_kernel"] = d.pop("cand_kernel")
            d["rnn_bias"] = d.pop("cand_bias")
        out.append(d)
    return out


def time_cpu_reference(w, budget_s=12.0, max_iters=200, threads=None):
    """node-updates/s of the fp32 torch-CPU restatement (oracle.ggnn_oracle.sparse_propagation_torch /
    dense_propagation_torch) on the host cores, bounded by ``budget_s`` seconds of work.  The TF graph's matmuls are
    small, so more threads are not always faster: 1 thread, 16 threads and all cores are each timed on a slice of the
    budget and the FASTEST setting is reported (``cores`` = the thread count that won)."""
    import torch
    from oracle import ggnn_oracle as O
    ow = oracle_weights(w)
    if w["kind"] == "dense":
        b, v = w["dense_shape"]
        h0 = torch.from_numpy(w["h0"].reshape(b, v, -1))
        adj = torch.from_numpy(w["adjacency_matrix"])
        dw = dict(ow[0])
        dp = {"num_timesteps": w["engine_params"]["layer_timesteps"][0], "use_edge_bias": w["engine_params"]["use_edge_bias"]}
        fn = lambda: O.dense_propagation_torch(h0, adj, dw, dp)
    else:
        h0 = torch.from_numpy(w["h0"])
        adj = [torch.from_numpy(a.astype(np.int64)) for a in w["adjacency_lists"]]
        indeg = torch.from_numpy(w["num_incoming_edges_per_type"])
        tw = [{k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in lw.items()} for lw in ow]
        fn = lambda: O.sparse_propagation_torch(h0, adj, indeg, tw, w["engine_params"])
    ncpu = os.cpu_count() or 1
    candidates = [threads] if threads else sorted({1, min(16, ncpu), ncpu})
    best, tried = None, []
    with torch.no_grad():
        for nt in candidates:
            torch.set_num_threads(nt)
            fn(); fn()
            times, t_start = [], time.perf_counter()
            while len(times) < max_iters and (time.perf_counter() - t_start) < budget_s / len(candidates):
                t0 = time.perf_counter(); fn(); times.append(time.perf_counter() - t0)
            med = statistics.median(times)
            tried.append("%d thr: %.2f ms" % (nt, med * 1e3))
            if best is None or med < best[0]:
                best = (med, nt, len(times), sum(times))
    med, nt, cnt, tot = best
    return {"value": w["node_updates"] / med, "unit": "node-updates/s", "cores": int(nt), "kind": "port", "ms_per_step": med * 1e3,
            "sample": "%d full forwards of %s (V=%d, M=%d) in %.1f s, median, best thread count of [%s] on a %d-core host; fp32 PyTorch-CPU "
                      "restatement of the TF1 graph (TF 1.3 not installable)" % (cnt, w["name"], w["V"], w["M"], tot, "; ".join(tried), ncpu)}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path.  TF 1.3 cannot be installed (no
    wheel, no network), so this is the oracle port (oracle/ggnn_oracle.py) on all host threads."""
    if rank != 0:
        return
    from gated_graph_neural_network_samples_b200 import workloads
    w = workloads.build(args.config, seed=0)
    # K steps + W warm-ups of full forwards, bounded to a few minutes
    res = time_cpu_reference(w, budget_s=min(120.0, 2.0 * max(args.steps, 1)), max_iters=max(args.steps, 3))
    line = {"impl": "reference", "metric": "GGNN node-state-updates/sec (propagation step)", "value": res["value"],
            "unit": "node-updates/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": res["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "V": w["V"], "M": w["M"], "hidden": w["engine_params"]["hidden_size"],
                       "edge_types": w["num_edge_types"], "layer_timesteps": w["engine_params"]["layer_timesteps"]},
            "cpu_baseline": {"value": res["value"], "unit": res["unit"], "cores": res["cores"], "kind": res["kind"], "sample": res["sample"]},
            "e2e": {"value": res["value"], "unit": "node-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    from gated_graph_neural_network_samples_b200 import workloads
    from gated_graph_neural_network_samples_b200.engine import PropagationEngine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # ---- workload: every rank owns one full configs[1]-sized shard of independent graphs (weak scaling,
    # no data-path collective: forward propagation never crosses graphs)
    w = workloads.build(args.config, seed=rank)
    P = w["engine_params"]
    if args.precision == "auto":
        args.precision = "bf16x3"   # tcgen05 on every config: tile-local fused kernel for D <= 128, streaming kernel above
    eng = PropagationEngine(P, w["num_edge_types"], device=local_rank, precision=args.precision)
    dev_w = [{k: torch.from_numpy(np.ascontiguousarray(v)).cuda() for k, v in lw.items()} for lw in w["weights"]]
    eng.set_weights(dev_w)
    dense = w["kind"] == "dense"
    if dense:
        eng.set_graph_dense(w["adjacency_matrix"])
    else:
        eng.set_graph_sparse(w["adjacency_lists"], w["num_incoming_edges_per_type"])
    h0 = torch.from_numpy(w["h0"]).cuda()
    out = torch.empty_like(h0)
    flush_buf = None if args.no_flush else torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")

    def flush():
        if flush_buf is not None:
            flush_buf.fill_(1)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing
    for _ in range(max(args.warmup, 3)):
        flush(); eng.forward(h0, out)
    sync_all()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    wall0 = time.perf_counter()
    for i in range(args.steps):
        flush()
        ev[i][0].record()
        eng.forward(h0, out)
        ev[i][1].record()
        launches += eng.last_launch_count
    sync_all()
    wall = time.perf_counter() - wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    dev_ms_total = float(sum(step_ms))
    # hot-L2 variant (no flush), for context: every BASELINE config is L2-resident by nature
    hot = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        hot[i][0].record(); eng.forward(h0, out); hot[i][1].record()
    sync_all()
    hot_ms = statistics.median([a.elapsed_time(b) for a, b in hot])
    clocks = sampler.stop() if rank == 0 else None

    # ---- end-to-end through the host-buffer API (pinned host inputs, H2D + D2H inside the timed region)
    h0_host = torch.from_numpy(w["h0"]).pin_memory()
    out_host = torch.empty_like(h0_host).pin_memory()
    h0_np, out_np = h0_host.numpy(), out_host.numpy()

    def e2e_step():   # one public call per batch, host buffers in and out (the shape of sess.run(fetch, feed_dict))
        if dense:
            eng.run_dense_host(w["adjacency_matrix"], h0_np, out_np)
        else:
            eng.run_sparse_host(w["adjacency_lists"], w["num_incoming_edges_per_type"], h0_np, out_np)

    for _ in range(max(args.warmup, 3)):
        e2e_step()
    sync_all()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    sync_all()
    e2e_ms_total = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)  # host work + copies + kernel, every step
    # same result either way (the tensor-core path's MMA issue order across issuer warps is not fixed -> fp32 rounding noise)
    np.testing.assert_allclose(out_np, out.cpu().numpy(), rtol=1e-4, atol=1e-5)
    # ---- same, two batches in flight (two engines on two streams; the reference overlaps batch preparation with
    # sess.run through ThreadedIterator, chem_tensorflow.py:225): reported beside the serial number, never instead of it
    engs = [eng, PropagationEngine(P, w["num_edge_types"], device=local_rank, precision=args.precision)]
    engs[1].set_weights(dev_w)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    outs = [out_host, torch.empty_like(h0_host).pin_memory()]

    def pipe_step(i):
        k = i & 1
        streams[k].synchronize()            # batch i-2 (same engine, same pinned result buffer) has landed
        with torch.cuda.stream(streams[k]):
            if dense:
                engs[k].set_graph_dense(w["adjacency_matrix"])
            else:
                engs[k].set_graph_sparse(w["adjacency_lists"], w["num_incoming_edges_per_type"])
            engs[k].forward_host(h0_np, outs[k].numpy(), sync=False)

    for i in range(4):
        pipe_step(i)
    sync_all()
    t0 = time.perf_counter()
    for i in range(args.steps):
        pipe_step(i)
    for k in (0, 1):
        streams[k].synchronize()
        with torch.cuda.stream(streams[k]):
            engs[k].sync_check()
    pipe_ms_total = (time.perf_counter() - t0) * 1e3
    sync_all()
    np.testing.assert_allclose(outs[1].numpy(), out.cpu().numpy(), rtol=1e-4, atol=1e-5)
    if dense:
        h2d = int(w["adjacency_matrix"].nbytes + w["V"] * w["num_edge_types"] * 4 + w["V"] * 4 + w["h0"].nbytes)
    else:
        h2d = int(4 * (w["V"] * w["num_edge_types"] + 1) + 8 * w["M"] + w["num_incoming_edges_per_type"].nbytes + 4 * w["V"] + w["h0"].nbytes)
    d2h = int(w["h0"].nbytes)

    # ---- secondary metric: training propagation = forward with saved states + backward, device-resident
    eng.set_save_for_backward(True)
    if dense:
        eng.set_graph_dense(w["adjacency_matrix"])
    else:
        eng.set_graph_sparse(w["adjacency_lists"], w["num_incoming_edges_per_type"])
    grads = [{k: torch.zeros_like(v) for k, v in lw.items()} for lw in dev_w]
    d_out = torch.ones_like(h0)
    d_h0 = torch.empty_like(h0)

    def train_step():
        eng.forward(h0, out)
        eng.backward(d_out, grads, d_h0)

    for _ in range(3):
        train_step()
    sync_all()
    tr = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for i in range(args.steps):
        flush(); tr[i][0].record(); train_step(); tr[i][1].record()
    sync_all()
    eng.sync_check()
    train_ms_total = float(sum(a.elapsed_time(b) for a, b in tr))
    eng.set_save_for_backward(False)

    # ---- the fused gated-regression readout that follows the propagation (one task), against the same op written
    # as the reference's TF op sequence in torch on the GPU (cat, 2 matmuls, sigmoid, mul, index_add / masked sum)
    readout = None
    if w["kind"] in ("sparse", "dense"):
        D = int(P["hidden_size"])
        rr = np.random.default_rng(3)
        wg = torch.from_numpy(rr.normal(0, 0.2, (2 * D, 1)).astype(np.float32)).cuda(); bg = torch.zeros(1, device="cuda")
        wt = torch.from_numpy(rr.normal(0, 0.2, (D, 1)).astype(np.float32)).cuda(); bt = torch.zeros(1, device="cuda")
        if dense:
            nb, nv = w["dense_shape"]
            eng.readout_set_graphs(nb, nodes_per_graph=nv, node_mask=w["node_mask"])
            mask_t = torch.from_numpy(np.ascontiguousarray(w["node_mask"], dtype=np.float32)).cuda()
        else:
            eng.readout_set_graphs(w["num_graphs"], graph_nodes_list=w["graph_nodes_list"])
            gnl_t = torch.from_numpy(np.asarray(w["graph_nodes_list"])).long().cuda()

        def torch_readout():
            gated = torch.sigmoid(torch.cat([out, h0], dim=-1) @ wg + bg) * (out @ wt + bt)
            if dense:
                return (gated.reshape(nb, nv) * mask_t).sum(dim=1)
            return torch.zeros(w["num_graphs"], 1, device="cuda").index_add_(0, gnl_t, gated).squeeze(-1)

        def timed(fn):
            for _ in range(3):
                fn()
            sync_all()
            evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
            for a, b in evs:
                flush(); a.record(); fn(); b.record()
            sync_all()
            return statistics.median([a.elapsed_time(b) for a, b in evs])

        fused_ms = timed(lambda: eng.readout_forward(out, h0, wg, bg, wt, bt))
        torch_ms = timed(torch_readout)
        np.testing.assert_allclose(eng.readout_forward(out, h0, wg, bg, wt, bt).cpu().numpy(), torch_readout().cpu().numpy(), rtol=1e-4, atol=1e-5)
        ro_bytes = 2 * w["V"] * D * 4 + w["V"] * 4 + w["num_graphs"] * 4 + 3 * D * 4   # read h_T and h_0 once, node->graph map, write [G]
        readout = {"fused_ms": fused_ms, "torch_ops_ms": torch_ms, "algorithmic_bytes": ro_bytes,
                   "achieved_gbs": ro_bytes / (fused_ms * 1e-3) / 1e9,
                   "what": "gated_regression (sparse:220-231 / dense:119-129) forward, one task, L2 flushed; fused kernel vs the TF op sequence in torch"}

    # ---- max over ranks
    t = torch.tensor([dev_ms_total, e2e_ms_total, hot_ms, pipe_ms_total, train_ms_total], dtype=torch.float64, device="cuda")
    units = torch.tensor([float(w["node_updates"])], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    dev_ms_total, e2e_ms_total, hot_ms, pipe_ms_total, train_ms_total = (float(x) for x in t.tolist())
    total_units_per_step = float(units.item())

    if rank == 0:
        ms_per_step = dev_ms_total / args.steps
        value = total_units_per_step / (ms_per_step * 1e-3)
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        alg_bytes = workloads.algorithmic_bytes(w)
        # dominant kernel = the fused propagation kernel; in LOCAL mode it IS the step (1 launch), in GLOBAL mode
        # the step is `launches/steps` launches of the same kernel: bytes and time are both per step
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.config, {}).get(args.precision)
            except Exception:
                traffic = None
        line = {
            "metric": "GGNN node-state-updates/sec (propagation step)", "value": value, "unit": "node-updates/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": {"fp32": "f32", "bf16x3": "bf16x3 (fp32 operands split hi+lo, 3 tensor-core MMAs, fp32 accumulate)", "bf16": "bf16"}[args.precision],
            "data": "synthetic",
            "config": {"workload": "%s: %s" % (w["name"], "BASELINE.json configs[1]" if w["name"] == "cfg2" else "see workloads.py"),
                       "V_per_gpu": w["V"], "M_per_gpu": w["M"], "graphs_per_gpu": w["num_graphs"],
                       "hidden": P["hidden_size"], "edge_types": w["num_edge_types"], "layer_timesteps": P["layer_timesteps"],
                       "residual_connections": P.get("residual_connections", {}), "cell": P["graph_rnn_cell"],
                       "use_edge_bias": P["use_edge_bias"], "use_edge_msg_avg_aggregation": P["use_edge_msg_avg_aggregation"],
                       "parallelism": "graphs sharded over %d GPU(s), no data-path collective" % world,
                       "l2": "hot" if args.no_flush else "flushed before every timed step (256 MiB write, untimed)",
                       "plan": eng.plan, "precision": args.precision},
            "value_hot_l2": total_units_per_step / (hot_ms * 1e-3), "ms_per_step_hot_l2": hot_ms,
            "wall_ms_per_step_incl_flush": wall * 1e3 / args.steps,
            "gpu_launches": launches,
            "e2e": {"value": total_units_per_step / (e2e_ms_total / args.steps * 1e-3), "unit": "node-updates/s",
                    "ms_per_step": e2e_ms_total / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "what": "run_{sparse,dense}_host per step: H2D h0 (pinned) | host CSR build + H2D graph, kernel, D2H result, sync; serial"},
            "e2e_pipelined": {"value": total_units_per_step / (pipe_ms_total / args.steps * 1e-3), "unit": "node-updates/s",
                              "ms_per_step": pipe_ms_total / args.steps,
                              "what": "same calls and bytes, two batches in flight (2 engines x 2 streams, forward_host_async); wall clock"},
            "train_propagation": {"value": total_units_per_step / (train_ms_total / args.steps * 1e-3), "unit": "node-updates/s",
                                  "ms_per_step": train_ms_total / args.steps,
                                  "what": "forward (states saved) + backward of the propagation (d weights, d h0), device-resident, fp32 backward"},
            "readout": readout,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": alg_bytes / max(launches / args.steps, 1),
                         "algorithmic_bytes_per_step": alg_bytes, "kernel": "ggnn_fwd_*_kernel", "peak_source": peak_src,
                         "algorithmic_gflop_per_step": workloads.algorithmic_flops(w) / 1e9},
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world >= 1:
            cb = time_cpu_reference(w, budget_s=12.0)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
