"""CPU study (no GPU): does another FIXED summation order of the BLAS-like contractions reproduce more rows of the
reference's results tables than the `dot4` order (4 interleaved FMA partial sums) the oracle and the kernels share?

The reference's own order is not reproducible (OpenBLAS kernels chosen per CPU at run time, hardware unknown), so the
question is statistical: per class, rows of experiments/ipddp2/results/*.txt with identical iteration count AND 9-digit
objective.  A scratch copy of oracle/ is compiled once per variant with -DORACLE_DOT_ORDER=<n>:
  0  dot4: 4 interleaved FMA partial sums, (s0+s1)+(s2+s3)          (shipped)
  1  sequential, one FMA accumulator
  2  sequential, separate multiply and add (no FMA)
  3  2 interleaved FMA partial sums, s0+s1
  4  8 interleaved FMA partial sums, ((s0+s1)+(s2+s3))+((s4+s5)+(s6+s7))
    python tests/tools/summation_order_study.py > profiles/r2_summation_order.json
"""
import ctypes as C
import json
import os
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import ipddp_b200  # noqa: E402,F401
from ipddp_b200 import instances  # noqa: E402

VARIANT = '''static inline double dot4(int n, const double* a, int sa, const double* b, int sb) {
#if ORACLE_DOT_ORDER == 1
  double s = 0.0;
  for (int i = 0; i < n; ++i) s = fma(a[i * sa], b[i * sb], s);
  return s;
#elif ORACLE_DOT_ORDER == 2
  double s = 0.0;
  for (int i = 0; i < n; ++i) { volatile double pr = a[i * sa] * b[i * sb]; s = s + pr; }
  return s;
#elif ORACLE_DOT_ORDER == 3
  double s0 = 0.0, s1 = 0.0;
  int i = 0;
  for (; i + 1 < n; i += 2) { s0 = fma(a[i * sa], b[i * sb], s0); s1 = fma(a[(i + 1) * sa], b[(i + 1) * sb], s1); }
  if (i < n) s0 = fma(a[i * sa], b[i * sb], s0);
  return s0 + s1;
#elif ORACLE_DOT_ORDER == 4
  double s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int i = 0; i < n; ++i) s[i & 7] = fma(a[i * sa], b[i * sb], s[i & 7]);
  return ((s[0] + s[1]) + (s[2] + s[3])) + ((s[4] + s[5]) + (s[6] + s[7]));
#else
'''


def build_variant(order, tmp):
    d = os.path.join(tmp, f"o{order}")
    shutil.copytree(os.path.join(ROOT, "oracle"), d, ignore=shutil.ignore_patterns("*.so", "*.hash", "__pycache__"))
    p = os.path.join(d, "ldlt.h")
    s = open(p).read()
    head = "static inline double dot4(int n, const double* a, int sa, const double* b, int sb) {\n"
    a = s.index(head)
    e = s.index("\n}\n", a) + 3
    body = s[a + len(head):e]
    s = s[:a] + VARIANT + body.rstrip()[:-1].rstrip() + "\n#endif\n}\n" + s[e:]
    open(p, "w").write(s)
    mk = open(os.path.join(d, "Makefile")).read().replace("CXXFLAGS = ", f"CXXFLAGS = -DORACLE_DOT_ORDER={order} ")
    open(os.path.join(d, "Makefile"), "w").write(mk)
    subprocess.check_call(["make", "-C", d, "-s", "-B"])
    return os.path.join(d, "libipddp_oracle.so")


def run(lib_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    oracle._lib = None
    oracle.LIB_PATH = lib_path
    oracle.build = lambda force=False: lib_path
    out = {}
    for wl in ("cartpole", "concar_quad", "acrobot", "concar", "double_integrator"):
        g = instances.load_golden_results(wl)
        n = len(g["seed"])
        b = instances.make_batch(wl, n, 101)
        res, _, _ = oracle.solve_batch(wl, 101, b.p, b.lower, b.upper, b.x1, b.ubar,
                                       options=oracle.default_options(optimality_tolerance=1e-7))
        k = np.array([r.k for r in res])
        obj = np.array([r.objective for r in res])
        same_it = k == g["iterations"]
        same_obj = np.abs(obj - g["objective"]) <= 1e-8 * np.maximum(1.0, np.abs(g["objective"]))
        out[wl] = dict(rows=n, same_iterations=int(same_it.sum()), same_both=int((same_it & same_obj).sum()),
                       mean_abs_iteration_gap=float(np.abs(k - g["iterations"]).mean()))
    return out


if __name__ == "__main__":
    names = {0: "dot4 (shipped)", 1: "sequential FMA", 2: "sequential multiply + add", 3: "2-way FMA", 4: "8-way FMA"}
    result = {}
    with tempfile.TemporaryDirectory() as tmp:
        for order in (0, 1, 2, 3, 4):
            result[names[order]] = run(build_variant(order, tmp))
            print(names[order], json.dumps(result[names[order]]), file=sys.stderr, flush=True)
    print(json.dumps(result, indent=1))
