"""Binary Shan-Chen path (BASELINE config 5) of the oracle against goldens computed with the
reference's own expressions and geometry (tools/ref_stepper.py:RefShanChenStepper)."""
import numpy as np
import pytest

import util

ALL = [n + '_' + ap for n in util.SC_CASES for ap in ('AB', 'AA')]


def run_oracle(oracle, c, dtype, steps=None):
    p = util.params_from_case(c)
    sim = oracle.ShanChenSimulation(p, c['map'], c['rho0'], c['phi0'], c['v0'], dtype=dtype)
    sim.run(int(c['steps']) if steps is None else steps)
    return sim


def errors(sim, c):
    m = util.real_mask(c)
    wet = np.zeros(m.shape, bool)
    wet[m] = c['fluid_map'].reshape(-1)
    shp = c['dist1'].shape
    e = [util.rel_linf(sim.current(0).reshape(shp)[:, wet], c['dist1'][:, wet]),
         util.rel_linf(sim.current(1).reshape(shp)[:, wet], c['dist2'][:, wet]),
         util.rel_linf(sim.rho.reshape(m.shape)[wet], c['rho'][wet]),
         util.rel_linf(sim.phi.reshape(m.shape)[wet], c['phi'][wet])]
    vs = max(1e-30, np.abs(c['v'][:, wet]).max())
    e.append(float(np.abs(np.array([x.reshape(m.shape)[wet] for x in sim.v]) - c['v'][:, wet]).max() / vs))
    return e


@pytest.mark.parametrize('case', ALL)
def test_fp64_oracle_matches_reference_expressions(oracle, case):
    c = util.load_case(case)
    e = errors(run_oracle(oracle, c, np.float64), c)
    assert max(e[:4]) < 1e-12 and e[4] < 1e-9, e


@pytest.mark.parametrize('case', ALL)
def test_fp32_oracle_roundoff(oracle, case):
    c = util.load_case(case)
    e = errors(run_oracle(oracle, c, np.float32), c)
    assert max(e[:4]) < 5e-6, e


@pytest.mark.parametrize('name', ['sc3d', 'sc2d_classic'])   # wall-free: full-way BB walls hold mass in transit
def test_total_mass_of_each_component_is_conserved(oracle, name):
    c = util.load_case(name + '_AA')
    sim = run_oracle(oracle, c, np.float64)
    m = util.real_mask(c)
    wet = np.zeros(m.shape, bool)
    wet[m] = c['fluid_map'].reshape(-1)
    for field, key in ((sim.rho, 'rho0'), (sim.phi, 'phi0')):
        assert abs(field.reshape(m.shape)[wet].sum() - c[key][wet].sum()) < 1e-9 * wet.sum()
