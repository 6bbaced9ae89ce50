import os, sys, ctypes
sys.path.insert(0, '/root/repo')
import numpy as np
if len(sys.argv) > 1:
    from gated_graph_neural_network_samples_b200 import _build
    _build.LIB_PATH = sys.argv[1]
    _build.is_stale = lambda: False
from tests import _util as U
from tests.test_gpu_parity import CFG2
from oracle import ggnn_oracle as O
NM = int(os.environ.get("NM", "64"))
_, b = U.molecule_batch(NM, 100, seed=5)
w = O.init_sparse_weights(CFG2, 4, np.random.default_rng(1))
ref = O.sparse_propagation_np(b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"], w, CFG2, dtype=np.float64)
got, eng = U.engine_sparse(CFG2, 4, w, b["adjacency_lists"], b["num_incoming_edges_per_type"], b["initial_node_representation"], precision="bf16x3", return_engine=True)
print(sys.argv[1:] or "default lib", eng.plan, "max rel err %.2e" % U.max_rel_err(got, ref))
