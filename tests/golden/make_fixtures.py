"""Copies the reference's committed known-answer tables into tests/golden/ (run in the build container,
where /root/reference exists; the GPU box only sees the committed copies).

results/*.txt : per instance seed, iterations, converged flag, objective (9 sig.), primal infeasibility,
                wall ms, solver ms   (reference experiments/ipddp2/results/*.txt, written by e.g.
                experiments/ipddp2/cartpole_friction.jl:151-161)
params/*.txt  : the per-seed random model parameters (reference experiments/ipddp2/params/*.txt, written
                by e.g. experiments/ipddp2/cartpole_friction.jl:164-168).  Needed because Julia's
                Xoshiro256++ draws cannot be regenerated without Julia.
These are data tables, not source code.
"""
import hashlib
import os
import shutil

REF = "/root/reference/experiments/ipddp2"
HERE = os.path.dirname(os.path.abspath(__file__))
RESULTS = ["cartpole_friction", "acrobot_contact", "concar", "concar_quad", "pushing_1_obs", "double_integrator"]
PARAMS = ["cartpole_friction", "acrobot_contact", "concar", "pushing_1_obs"]

if __name__ == "__main__":
    lines = []
    for kind, names in (("results", RESULTS), ("params", PARAMS)):
        os.makedirs(os.path.join(HERE, kind), exist_ok=True)
        for n in names:
            src = os.path.join(REF, kind, n + ".txt")
            dst = os.path.join(HERE, kind, n + ".txt")
            shutil.copyfile(src, dst)
            h = hashlib.sha256(open(dst, "rb").read()).hexdigest()[:16]
            lines.append(f"{kind}/{n}.txt sha256[:16]={h} from {src}")
    open(os.path.join(HERE, "PROVENANCE.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
