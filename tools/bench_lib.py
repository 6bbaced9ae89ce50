"""bench.py with an alternative build of the library (experiments): python tools/bench_lib.py <lib.so> [bench args]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gated_graph_neural_network_samples_b200 import _build
_build.LIB_PATH = os.path.abspath(sys.argv[1]); _build.is_stale = lambda: False
sys.argv = ["bench.py"] + sys.argv[2:]
import runpy
runpy.run_path(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py"), run_name="__main__")
