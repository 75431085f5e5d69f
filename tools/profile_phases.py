"""Runs init + a few rounds of derivs/backward/check/forward once (phase API) for ncu captures."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ipddp_b200
from ipddp_b200 import _lib, instances
from ipddp_b200.batch import BatchSolver
wl = sys.argv[1] if len(sys.argv) > 1 else "cartpole"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
rounds = int(sys.argv[3]) if len(sys.argv) > 3 else 3
lib = _lib.load()
b = instances.make_batch(wl, B, 101)
s = BatchSolver(wl, B, 101, options=lib.default_options(optimality_tolerance=1e-7), lib=lib)
s.set_batch(b)
s.initialize()
for r in range(rounds):
    s.eval_derivatives(); s.backward_pass(); n = s.check()
    if n > 0: s.forward_pass()
print("done", s.results().k[:4])
s.close()
