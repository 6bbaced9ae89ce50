import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import _util as U
from tests.test_gpu_parity import CFG1_TRUE, CFG2
from oracle import ggnn_oracle as O
which = sys.argv[1] if len(sys.argv) > 1 else "rnn_res"
params = {"rnn_res": dict(CFG1_TRUE, hidden_size=128, graph_rnn_cell="RNN"),
          "gru_res": dict(CFG1_TRUE, hidden_size=128),
          "gru_plain": dict(CFG2, hidden_size=128),
          "rnn_plain": dict(CFG2, hidden_size=128, graph_rnn_cell="RNN")}[which]
_, b = U.molecule_batch(100, 128, T=4, seed=5)
w = O.init_sparse_weights(params, 4, np.random.default_rng(1))
ref = O.sparse_propagation_np(b["initial_node_representation"], b["adjacency_lists"], b["num_incoming_edges_per_type"], w, params, dtype=np.float64)
got, eng = U.engine_sparse(params, 4, w, b["adjacency_lists"], b["num_incoming_edges_per_type"], b["initial_node_representation"], precision="bf16x3", return_engine=True)
print(which, eng.plan, "max rel err %.2e" % U.max_rel_err(got, ref))
