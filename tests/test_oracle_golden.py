"""Pins the CPU oracle (oracle/lbm_oracle.c) to the reference:

  * whole steps vs tests/golden/steps_*.npz -- N steps computed by tools/ref_stepper.py,
    whose arithmetic is the reference's SymPy expressions and whose node map / parameters
    are the reference's own (Subdomain.reset + GeoEncoderConst);
  * invariants the reference itself tests: AA == AB (tests/gpu/access_pattern.sh,
    tests/gpu/do_nothing_node.py:37-61) and the exact moments of f^eq
    (tests/sym_equilibrium.py:93-121).

fp64 oracle must reproduce the golden to round-off; the fp32 oracle (the precision the
reference computes in) within 5e-6 of the fp64 golden, relative to ||ref||_inf.
"""
import numpy as np
import pytest

import util

ALL = [n + '_' + ap for n in util.CASES for ap in ('AB', 'AA')]


def _run(oracle, c, dtype):
    p = util.params_from_case(c)
    sim = oracle.Simulation(p, c['map'], c['rho0'], c['v0'], dtype=dtype)
    sim.run(int(c['steps']))
    return sim


def _errors(sim, c):
    m = util.real_mask(c)
    wet = np.zeros(m.shape, bool)
    wet[m] = c['fluid_map'].reshape(-1)
    d = sim.current_dist().reshape(c['dist'].shape)
    return (util.rel_linf(d[:, m], c['dist'][:, m]),
            util.rel_linf(sim.rho.reshape(m.shape)[wet], c['rho'][wet]),
            util.rel_linf(np.array([x.reshape(m.shape)[wet] for x in sim.v]), c['v'][:, wet]))


@pytest.mark.parametrize('case', ALL)
def test_fp64_oracle_reproduces_reference_stepper(oracle, case):
    c = util.load_case(case)
    e_dist, e_rho, e_v = _errors(_run(oracle, c, np.float64), c)
    assert e_dist < 1e-12 and e_rho < 1e-12 and e_v < 1e-11, (e_dist, e_rho, e_v)


@pytest.mark.parametrize('case', ALL)
def test_fp32_oracle_within_roundoff(oracle, case):
    c = util.load_case(case)
    e_dist, e_rho, e_v = _errors(_run(oracle, c, np.float32), c)
    assert e_dist < 5e-6 and e_rho < 5e-6, (e_dist, e_rho)
    # velocities are O(1e-2) sums of O(1) populations: allow their cancellation error
    assert e_v < 5e-5, e_v


@pytest.mark.parametrize('name', util.CASES)
def test_aa_equals_ab(oracle, name):
    """Reference invariant: the AA and AB access patterns give the same fields."""
    a = _run(oracle, util.load_case(name + '_AB'), np.float64)
    c = util.load_case(name + '_AA')
    b = _run(oracle, c, np.float64)
    m = util.real_mask(c)
    wet = np.zeros(m.shape, bool)
    wet[m] = c['fluid_map'].reshape(-1)
    assert np.allclose(a.rho.reshape(m.shape)[wet], b.rho.reshape(m.shape)[wet], rtol=0, atol=1e-13)
    for x, y in zip(a.v, b.v):
        assert np.allclose(x.reshape(m.shape)[wet], y.reshape(m.shape)[wet], rtol=0, atol=1e-13)


@pytest.mark.parametrize('name', ['box2d_bgk', 'box3d_bgk', 'box3d_q27'])
def test_periodic_box_equals_wraparound_streaming(oracle, name):
    """The ghost-layer PBC kernels must be equivalent to true periodic streaming."""
    from sailfish_b200 import lattice
    c = util.load_case(name + '_AA')
    L = lattice.get(str(c['grid']))
    D = L.dim
    lat = [int(x) for x in c['lat']]
    m = util.real_mask(c)
    shape = tuple(n - 2 for n in reversed(lat[:D]))
    rho = c['rho0'][m].reshape(shape)
    v = [c['v0'][d][m].reshape(shape) for d in range(D)]

    def feq(rho, v):
        vv = sum(x * x for x in v)
        out = []
        for i in range(L.Q):
            ev = sum(int(L.basis[i][d]) * v[d] for d in range(D))
            out.append(L.weights[i] * (rho + rho * (3 * ev + 4.5 * ev * ev - 1.5 * vv)))
        return np.array(out)
    f = feq(rho, v)
    tau = (6 * float(c['visc']) + 1) / 2
    steps = int(c['steps'])
    for s in range(steps):
        rho = f.sum(0)
        v = [sum(int(L.basis[i][d]) * f[i] for i in range(L.Q)) / rho for d in range(D)]
        f = f + (feq(rho, v) - f) / tau
        for i in range(L.Q):
            e = L.basis[i]
            f[i] = np.roll(f[i], tuple(int(e[D - 1 - k]) for k in range(D)), axis=tuple(range(D)))
    sim = _run(oracle, c, np.float64)
    assert steps % 2 == 0   # AA: populations are back in natural slots
    d = sim.current_dist().reshape(c['dist'].shape)
    assert np.abs(d[:, m] - f.reshape(L.Q, -1)).max() < 1e-13


@pytest.mark.parametrize('case', util.ZOO_STEP_CASES)
def test_zoo_fp64_oracle_reproduces_reference_stepper(oracle, case):
    """Node types the examples do not use (tests/geometry_zoo.py)."""
    c = util.load_case(case)
    e_dist, e_rho, e_v = _errors(_run(oracle, c, np.float64), c)
    assert e_dist < 1e-12 and e_rho < 1e-12 and e_v < 1e-11, (e_dist, e_rho, e_v)
    e_dist, e_rho, e_v = _errors(_run(oracle, c, np.float32), c)
    assert e_dist < 5e-6 and e_rho < 5e-6 and e_v < 1e-4, (e_dist, e_rho, e_v)


def _ab_only(name):
    import geometry_zoo
    return name[len('zoo_'):] in geometry_zoo.AB_ONLY


def _tms_along_periodic_axis(name):
    # TMS walls that reach a periodic face: in the AA pattern the wall nodes next to the face read, on odd steps, the
    # slot their bounced population went to AFTER the periodic copy has overwritten it with the (stale) value of the
    # opposite side's unused node -- AA and AB differ there by construction (tests/geometry_zoo.py AB_PINNED_ONLY)
    return name == 'zoo_tms_channel2d'


@pytest.mark.parametrize('name', sorted(set(c[:-3] for c in util.ZOO_STEP_CASES
                                            if not _ab_only(c[:-3]) and not _tms_along_periodic_axis(c[:-3]))))
def test_zoo_aa_equals_ab(oracle, name):
    ab = util.load_case(name + '_AB')
    a = _run(oracle, ab, np.float64)
    p = util.params_from_case(ab, access_pattern='AA')
    b = oracle.Simulation(p, ab['map'], ab['rho0'], ab['v0'], dtype=np.float64)
    b.run(int(ab['steps']))
    m = util.real_mask(ab)
    wet = np.zeros(m.shape, bool)
    wet[m] = ab['fluid_map'].reshape(-1)
    assert np.allclose(a.rho.reshape(m.shape)[wet], b.rho.reshape(m.shape)[wet], rtol=0, atol=1e-13)
    for x, y in zip(a.v, b.v):
        assert np.allclose(x.reshape(m.shape)[wet], y.reshape(m.shape)[wet], rtol=0, atol=1e-13)



@pytest.mark.parametrize('name', sorted(set(c[:-3] for c in util.ZOO_STEP_CASES if _ab_only(c[:-3]))))
def test_copy_and_extrapolation_outflow_are_ab_only(oracle, name):
    """The reference generates '#error NTYuOutflow, NTGradFreeflow and NTCopy are not supported with the AA
    access_pattern' (boundary.mako:626-628): parameter construction and the oracle both refuse."""
    from sailfish_b200 import params as P
    ab = util.load_case(name + '_AB')
    with pytest.raises(P.UnsupportedConfiguration):
        util.params_from_case(ab, access_pattern='AA')
    p = util.params_from_case(ab)
    p.access_pattern = P.ACCESS_IDS['AA']
    with pytest.raises(Exception):
        oracle.Simulation(p, ab['map'], ab['rho0'], ab['v0'], dtype=np.float64).run(1)
