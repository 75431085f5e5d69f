"""The reference experiment harness' file formats.

  results tables  experiments/ipddp2/results/<class>.txt, written by every experiment script, e.g.
                  experiments/ipddp2/cartpole_friction.jl:151-161 (`@printf` formats restated below), read back by
                  `read_results` in experiments/utils.jl:4-64 (both regular expressions restated below);
  params tables   experiments/ipddp2/params/<class>.txt, cartpole_friction.jl:163-168 (one instance per line, values joined
                  by blanks).

`write_results` / `read_results` / `write_params` are what tools/run_experiments.py uses to regenerate the tables from a
batched GPU solve; tests/test_results_io.py proves that parsing the reference's own tables and writing them again reproduces
the files byte for byte.
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List

HEADER = " seed  iterations  status     objective           primal        wall (ms)   solver(ms)  \n"

# experiments/utils.jl:5-6 -- Julia's regex syntax is PCRE, as Python's: copied character for character (the unescaped `.`
# matches any character there too)
REGEX_RESULTS = re.compile(r"\s*(\d+)\s+(\d+)\s+(\w+)\s+(\d+.\d+e[+-]\d+)\s+(\d+.\d+e[+-]\d+)\s+(\d+.\d+)\s+(\d+.\d+)")
REGEX_NO_BM = re.compile(r"\s*(\d+)\s+(\d+)\s+(\w+)\s+([+-]?\d+.\d+e?[+-]?\d+)?\s+(\d+.\d+e?[+-]?\d+)")


@dataclass
class ResultsTable:
    """Column vectors of one results file, in the order `read_results` returns them (experiments/utils.jl:63)."""
    seeds: List[int] = field(default_factory=list)
    iters: List[int] = field(default_factory=list)
    status: List[bool] = field(default_factory=list)
    objs: List[float] = field(default_factory=list)
    constrs: List[float] = field(default_factory=list)
    walls: List[float] = field(default_factory=list)
    solvers: List[float] = field(default_factory=list)
    benchmark: bool = True

    def __len__(self):
        return len(self.seeds)


def _parse_bool(s: str) -> bool:
    """Julia's parse(Bool, s): "true" / "false" (or an integer literal)."""
    if s == "true":
        return True
    if s == "false":
        return False
    return bool(int(s))


def read_results(fname: str) -> ResultsTable:
    """experiments/utils.jl:4-64: lines that match the 7-column pattern, else the 5-column pattern (no timings: wall =
    solver = 0); every other line (the header) is skipped."""
    t = ResultsTable()
    any_bm = False
    with open(fname, "r") as fh:
        for line in fh.read().splitlines():
            m = REGEX_RESULTS.match(line)
            wall = solver = 0.0
            if m is not None:
                wall, solver = float(m.group(6)), float(m.group(7))
                any_bm = True
            else:
                m = REGEX_NO_BM.match(line)
                if m is None:
                    continue
            t.seeds.append(int(m.group(1)))
            t.iters.append(int(m.group(2)))
            t.status.append(_parse_bool(m.group(3)))
            t.objs.append(float(m.group(4)))
            t.constrs.append(float(m.group(5)))
            t.walls.append(wall)
            t.solvers.append(solver)
    t.benchmark = any_bm
    return t


def write_results(fname: str, seeds, iters, status_ok, objs, constrs, walls_ms=None, solvers_ms=None) -> None:
    """cartpole_friction.jl:151-161.  status_ok[i] is the printed Bool (`status == 0`); with walls_ms / solvers_ms the
    benchmark format (timings in ms), without them the 5-column format."""
    bench = walls_ms is not None and solvers_ms is not None
    with open(fname, "w") as fh:
        fh.write(HEADER)
        for i in range(len(seeds)):
            ok = "true" if status_ok[i] else "false"
            if bench:
                fh.write(" %2s     %5s      %5s    %.8e    %.8e     %5.1f        %5.1f  \n" % (
                    int(seeds[i]), int(iters[i]), ok, objs[i], constrs[i], walls_ms[i], solvers_ms[i]))
            else:
                fh.write(" %2s     %5s      %5s    %.8e    %.8e \n" % (int(seeds[i]), int(iters[i]), ok, objs[i], constrs[i]))


def _julia_float_string(x: float) -> str:
    """Julia's string(::Float64): shortest round-trip digits, always with a decimal point, exponent form `1.0e-5`
    outside [1e-5, 1e21)."""
    r = repr(float(x))
    if "e" in r or "E" in r:
        mant, exp = r.lower().split("e")
        if "." not in mant:
            mant += ".0"
        return f"{mant}e{int(exp)}"
    if "." not in r and "inf" not in r and "nan" not in r:
        r += ".0"
    return {"inf": "Inf", "-inf": "-Inf", "nan": "NaN"}.get(r, r)


def write_params(fname: str, rows) -> None:
    """cartpole_friction.jl:163-168: `println(io, join(string.(params[i]), " "))` per instance."""
    with open(fname, "w") as fh:
        for row in rows:
            fh.write(" ".join(_julia_float_string(v) for v in row) + "\n")


def read_params(fname: str):
    with open(fname, "r") as fh:
        return [[float(t) for t in line.split()] for line in fh.read().splitlines() if line.strip()]
