"""GPU: the engine against fixtures computed by the REFERENCE'S OWN graph code -- ``refgraph_*.npz`` come from the unmodified
``prepare_specific_graph_model`` / ``compute_final_node_representations`` / ``gated_regression`` of /root/reference evaluated in float64
over ``tests/golden/tf_shim.py`` (generator: tests/golden/make_reference_graph_golden.py).  North-star tolerance: 1e-4 relative."""
import json
import os

import numpy as np
import pytest

from tests import _util as U

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "refgraph_sparse_%s.npz" % name))
    p = json.loads(str(z["params_json"]))
    w = [{k[len("w%d_" % l):]: z[k] for k in z.files if k.startswith("w%d_" % l)} for l in range(len(p["layer_timesteps"]))]
    return z, p, w, [z["adj%d" % e] for e in range(4)]


@pytest.mark.parametrize("name,precision", [("true_default_shape", "fp32"), ("true_default_shape", "bf16x3"), ("rnn_relu_bias_sum", "fp32"),
                                            ("attention_bias_avg", "fp32")])
def test_sparse_propagation_and_readout_match_the_reference_graph_code(golden_dir, name, precision):
    import torch
    z, p, w, adj = _load(golden_dir, name)
    got, eng = U.engine_sparse(p, 4, w, adj, z["indeg"].astype(np.float32), z["h0"].astype(np.float32), precision=precision, return_engine=True)
    err = U.max_rel_err(got, z["final"])
    print("refgraph %-20s %-6s propagation max rel err %.2e" % (name, precision, err))
    assert np.all(np.isfinite(got)) and err < 1e-4
    # gated_regression (sparse:220-231) on the engine's own final states
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    eng.readout_set_graphs(int(z["num_graphs"]), graph_nodes_list=z["graph_nodes_list"])
    ro = eng.readout_forward(f32(got), f32(z["h0"]), f32(z["ro_w_gate"]), f32(z["ro_b_gate"]), f32(z["ro_w_trans"]), f32(z["ro_b_trans"]))
    eng.sync_check()
    err = U.max_rel_err(ro.cpu().numpy(), z["readout"])
    print("refgraph %-20s %-6s readout     max rel err %.2e" % (name, precision, err))
    assert err < 1e-4


@pytest.mark.parametrize("fixture", ["refgraph_dense.npz", "refgraph_dense_cfg3_shape.npz"])   # hidden 12; BASELINE configs[2] width (hidden 100)
@pytest.mark.parametrize("precision", ["fp32", "bf16x3"])
def test_dense_propagation_matches_the_reference_graph_code(golden_dir, precision, fixture):
    z = np.load(os.path.join(golden_dir, fixture))
    p = json.loads(str(z["params_json"]))
    w = {k[2:]: z[k] for k in z.files if k.startswith("w_")}
    got = U.engine_dense(p, 4, w, z["adj"].astype(np.float32), z["h0"].astype(np.float32), precision=precision)
    err = U.max_rel_err(got, z["final"])
    print("refgraph dense %-6s max rel err %.2e" % (precision, err))
    assert np.all(np.isfinite(got)) and err < 1e-4


@pytest.mark.parametrize("name,precision,plan", [("cfg2_shape", "bf16x3", "LOCAL"), ("cfg2_shape", "fp32", "LOCAL"),
                                                 ("cfg4_shape", "bf16x3", "STREAM"), ("cfg4_shape", "fp32", "")])
def test_baseline_width_fixtures_of_the_reference_graph_code(golden_dir, name, precision, plan):
    """BASELINE configs[1] / configs[3] shapes (hidden 100 / 256) computed by the reference's own graph code: the tile-local fused tcgen05
    kernel, the streaming tcgen05 kernels and the fp32 kernel against them directly -- no oracle in between."""
    import torch
    z, p, w, adj, T = U.load_refgraph_wide(golden_dir, name)
    got, eng = U.engine_sparse(p, T, w, adj, z["indeg"], z["h0"], precision=precision, return_engine=True)
    assert plan in eng.plan, eng.plan
    err = U.max_rel_err(got, z["final"])
    print("refgraph %-12s %-6s propagation max rel err %.2e  [%s]" % (name, precision, err, eng.plan[:40]))
    assert np.all(np.isfinite(got)) and err < 1e-4
    f32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()
    eng.readout_set_graphs(int(z["num_graphs"]), graph_nodes_list=z["graph_nodes_list"])
    ro = eng.readout_forward(f32(got), f32(z["h0"]), f32(z["ro_w_gate"]), f32(z["ro_b_gate"]), f32(z["ro_w_trans"]), f32(z["ro_b_trans"]))
    eng.sync_check()
    assert U.max_rel_err(ro.cpu().numpy(), z["readout"]) < 1e-4
