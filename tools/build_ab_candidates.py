"""Builds the k_backward candidates that wait for their A/B (DESIGN.md "Next", profiles/r1_ab/summary.md) as
interiorpointddp.jl_b200/libipddp_b200_<name>.so and prints the gpurun command that measures them (10 s of GPU each).
The variant libraries are never loaded by the package; delete them afterwards (they travel with every gpurun snapshot)."""
import concurrent.futures as cf
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ipddp_b200 import build as b  # noqa: E402

CANDIDATES = [
    ("c_base", []),
]

if __name__ == "__main__":
    with cf.ThreadPoolExecutor(4) as ex:
        for lib in ex.map(lambda a: b.build_variant(*a), CANDIDATES):
            print(lib)
    names = " ".join(n for n, _ in CANDIDATES)
    print(f"\ngpurun --timeout 150 -- 'bash tools/ab_variants.sh {names}'")
