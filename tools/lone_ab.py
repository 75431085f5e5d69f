import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np, ipddp_b200
from ipddp_b200 import _lib, instances
from ipddp_b200.batch import BatchSolver
lib = _lib.load()
for bw, fw in ((592, 148), (0, 148), (592, 0), (0, 0)):
    lib.L.ipddp_set_tuning(None, b"bw_spec_max", bw); lib.L.ipddp_set_tuning(None, b"fw_spec_max", fw)
    b = instances.make_batch("cartpole", 8, 101)
    s = BatchSolver("cartpole", 8, 101, options=lib.default_options(optimality_tolerance=1e-7), lib=lib)
    s.set_batch(b); s.solve(); s.solve(); st = s.stats()
    print(json.dumps(dict(bw_spec=bw, fw_spec=fw, ms_total=round(st.ms_total,1), ms_backward=round(st.ms_backward,1), ms_forward=round(st.ms_forward,1), rounds=st.iterations)))
    s.close()
