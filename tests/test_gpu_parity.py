"""Parity of the CUDA path (through the C ABI) with the oracle on the golden cases.

Metric: ||gpu - ref||_inf / ||ref||_inf over real (non-ghost) nodes; rho and v over wet nodes.
The reference arithmetic is fp32; its own round-off against exact (fp64) arithmetic is measured
with the fp32 oracle and is ~1e-6 for rho and up to ~2e-5 for v after 20 steps (v is a small
difference of O(1) populations).  Bar:
  * after 4 steps: <= 1e-6 (rho, populations, v) against the fp64 oracle [north_star: 1e-6 fp32], or 1.5x (v: 2x)
    the fp32 oracle's own error where that is larger;
  * after the full golden step count: within 2x the fp32 oracle's own error (+1e-6).
"""
import numpy as np
import pytest

import util

pytestmark = pytest.mark.gpu

ALL = [n + '_' + ap for n in util.CASES for ap in ('AB', 'AA')]


def _gpu_run(c, steps):
    from sailfish_b200.runtime import DeviceLattice
    p = util.params_from_case(c)
    dev = DeviceLattice(p, c['map'], c['rho0'], c['v0'])
    dev.run(steps)
    rho, v = dev.fields_to_host()
    return dev.get_dist(), np.array(rho), [np.array(x) for x in v]


def _oracle_run(oracle, c, steps, dtype):
    p = util.params_from_case(c)
    sim = oracle.Simulation(p, c['map'], c['rho0'], c['v0'], dtype=dtype)
    sim.run(steps)
    shape = c['map'].shape
    return (sim.current_dist().reshape((-1,) + shape), sim.rho.reshape(shape),
            [x.reshape(shape) for x in sim.v])


def _errs(a, b, c):
    m = util.real_mask(c)
    wet = np.zeros(m.shape, bool)
    wet[m] = c['fluid_map'].reshape(-1)
    return (util.rel_linf(a[0][:, m], b[0][:, m]), util.rel_linf(a[1][wet], b[1][wet]),
            util.rel_linf(np.array([x[wet] for x in a[2]]), np.array([x[wet] for x in b[2]])))


@pytest.mark.parametrize('case', ALL)
def test_short_run_within_1e6(oracle, case):
    c = util.load_case(case)
    g = _gpu_run(c, 4)
    r = _oracle_run(oracle, c, 4, np.float64)
    r32 = _oracle_run(oracle, c, 4, np.float32)
    e_dist, e_rho, e_v = _errs(g, r, c)
    o_dist, o_rho, o_v = _errs(r32, r, c)
    # 1e-6 where the reference's own fp32 arithmetic achieves it, else no worse than 1.5x that
    assert e_dist <= max(1e-6, 1.5 * o_dist), (e_dist, o_dist)
    assert e_rho <= max(1e-6, 1.5 * o_rho), (e_rho, o_rho)
    assert e_v <= max(1e-6, 2.0 * o_v), (e_v, o_v)
    assert e_rho <= 2e-6 and e_dist <= 2e-6


@pytest.mark.parametrize('case', ALL)
def test_golden_run_within_fp32_envelope(oracle, case):
    c = util.load_case(case)
    steps = int(c['steps'])
    g = _gpu_run(c, steps)
    r64 = _oracle_run(oracle, c, steps, np.float64)
    r32 = _oracle_run(oracle, c, steps, np.float32)
    eg = _errs(g, r64, c)
    eo = _errs(r32, r64, c)
    for a, b in zip(eg, eo):
        assert a <= 2.0 * b + 1e-6, (eg, eo)
    # and against the reference-derived golden vectors themselves
    gold = (c['dist'], c['rho'], list(c['v']))
    egold = _errs(g, gold, c)
    for a, b in zip(egold, eo):
        assert a <= 2.0 * b + 1e-6, (egold, eo)


@pytest.mark.parametrize('name', util.CASES)
def test_aa_equals_ab_on_gpu(name):
    """tests/gpu/access_pattern.sh of the reference: AA and AB give the same fields."""
    ca, cb = util.load_case(name + '_AB'), util.load_case(name + '_AA')
    steps = int(ca['steps'])
    a = _gpu_run(ca, steps)
    b = _gpu_run(cb, steps)
    _, e_rho, e_v = _errs(a, b, ca)
    assert e_rho <= 2e-6 and e_v <= 5e-5, (e_rho, e_v)
