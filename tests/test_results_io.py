"""The experiment harness' file formats (reference experiments/utils.jl:4-64 parser, experiments/ipddp2/*.jl writers):
parse the reference's own committed tables, write them again, compare BYTES."""
import os

import numpy as np
import pytest

import ipddp_b200  # noqa: F401
from ipddp_b200 import results_io

HERE = os.path.dirname(os.path.abspath(__file__))
RESULTS = ["cartpole_friction", "acrobot_contact", "concar", "concar_quad", "pushing_1_obs", "double_integrator"]
PARAMS = ["cartpole_friction", "acrobot_contact", "concar", "pushing_1_obs"]


@pytest.mark.parametrize("name", RESULTS)
def test_results_table_round_trip(name, tmp_path):
    src = os.path.join(HERE, "golden", "results", name + ".txt")
    t = results_io.read_results(src)
    n = len(open(src).read().splitlines()) - 1
    assert len(t) == n and t.benchmark
    assert t.seeds == list(range(1, n + 1))
    # the regex columns against a plain whitespace split of the same lines
    for i, line in enumerate(open(src).read().splitlines()[1:]):
        c = line.split()
        assert (t.iters[i], t.status[i]) == (int(c[1]), c[2] == "true")
        assert (t.objs[i], t.constrs[i], t.walls[i], t.solvers[i]) == (float(c[3]), float(c[4]), float(c[5]), float(c[6]))
    dst = tmp_path / (name + ".txt")
    results_io.write_results(str(dst), t.seeds, t.iters, t.status, t.objs, t.constrs, t.walls, t.solvers)
    assert open(dst, "rb").read() == open(src, "rb").read()
    # what the writer produced parses to the same table
    t2 = results_io.read_results(str(dst))
    assert (t2.iters, t2.status, t2.objs, t2.constrs, t2.walls, t2.solvers) == (t.iters, t.status, t.objs, t.constrs, t.walls, t.solvers)


def test_results_without_benchmark_columns(tmp_path):
    """the 5-column format (benchmark = false, cartpole_friction.jl:158): second regex of read_results, zero timings"""
    dst = tmp_path / "r.txt"
    results_io.write_results(str(dst), [1, 2], [60, 1000], [True, False], [9.29397628e-01, 1.5e3], [4.5e-14, 2.0e-3])
    t = results_io.read_results(str(dst))
    assert not t.benchmark and t.iters == [60, 1000] and t.status == [True, False]
    assert t.objs == [9.29397628e-01, 1.5e3] and t.constrs == [4.5e-14, 2.0e-3] and t.walls == [0.0, 0.0]
    assert open(dst).read().splitlines()[1] == "  1        60       true    9.29397628e-01    4.50000000e-14 "


@pytest.mark.parametrize("name", PARAMS)
def test_params_table_round_trip(name, tmp_path):
    src = os.path.join(HERE, "golden", "params", name + ".txt")
    rows = results_io.read_params(src)
    assert len(rows) == 100
    dst = tmp_path / (name + ".txt")
    results_io.write_params(str(dst), rows)
    assert open(dst, "rb").read() == open(src, "rb").read()
    assert np.array_equal(np.array(rows), np.loadtxt(src, ndmin=2))


def test_julia_float_strings():
    f = results_io._julia_float_string
    assert [f(1.0), f(0.1), f(1e-5), f(1.5e-7), f(1e21), f(123456.0), f(-2.5)] == \
        ["1.0", "0.1", "1.0e-5", "1.5e-7", "1.0e21", "123456.0", "-2.5"]
