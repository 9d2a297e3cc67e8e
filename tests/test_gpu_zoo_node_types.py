"""GPU parity on tests/geometry_zoo.py: node classes the BASELINE examples do not use -- half-way bounce-back,
equilibrium-velocity, regularised-density, Zou-He, do-nothing, copy and Yu outflow, TMS and full-slip walls, obstacles
with unused cores, D3Q27 with pressure planes.  The oracle is pinned on these cases (tests/test_oracle_golden.py,
reference-expression goldens from tools/gen_golden_steps.py zoo).

Status: all cases passed on a B200 in round 2 (profiles/r02_gpu_tests_new_node_types.txt,
profiles/r02_gpu_tests_slip.txt)."""
import numpy as np
import pytest

import util
from test_gpu_parity import _errs, _gpu_run, _oracle_run

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', util.ZOO_STEP_CASES)
def test_zoo_short_run_within_1e6(oracle, case):
    c = util.load_case(case)
    g = _gpu_run(c, 4)
    r = _oracle_run(oracle, c, 4, np.float64)
    r32 = _oracle_run(oracle, c, 4, np.float32)
    e_dist, e_rho, e_v = _errs(g, r, c)
    o_dist, o_rho, o_v = _errs(r32, r, c)
    assert e_dist <= max(1e-6, 1.5 * o_dist), (e_dist, o_dist)
    assert e_rho <= max(1e-6, 1.5 * o_rho), (e_rho, o_rho)
    assert e_v <= max(1e-6, 2.0 * o_v), (e_v, o_v)


@pytest.mark.parametrize('case', util.ZOO_STEP_CASES)
def test_zoo_golden_run_within_fp32_envelope(oracle, case):
    c = util.load_case(case)
    steps = int(c['steps'])
    g = _gpu_run(c, steps)
    r = _oracle_run(oracle, c, steps, np.float64)
    r32 = _oracle_run(oracle, c, steps, np.float32)
    e, o = _errs(g, r, c), _errs(r32, r, c)
    for a, b in zip(e, o):
        assert a <= 2.0 * b + 1e-6, (e, o)
    gold = (c['dist'], c['rho'], list(c['v']))
    for a, b in zip(_errs(g, gold, c), o):
        assert a <= 2.0 * b + 1e-6
