import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gated_graph_neural_network_samples_b200 import _build
_build.LIB_PATH = os.path.abspath(sys.argv[1]); _build.is_stale = lambda: False
sys.argv = ["tc_phase_timing.py"] + sys.argv[2:]
import runpy
runpy.run_path(os.path.join(os.path.dirname(os.path.abspath(__file__)), "tc_phase_timing.py"), run_name="__main__")
