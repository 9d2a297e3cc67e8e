"""Formula-level pinning of the oracle against tests/golden/formulas.npz -- values obtained by
lambdifying the reference's own SymPy expressions (tools/gen_golden.py): f^eq
(sym_equilibrium.py:90-120), velocity/density-node closures (sym.py:573-682), non-equilibrium
bounce-back (sym.py:750-766), regularisation (sym.py:882-891), Guo forcing (sym_force.py:121-160),
MRT moments / equilibria / rates (sym.py:98-149, 361-406).  fp64 oracle, tolerance 1e-13."""
import ctypes
import os

import numpy as np
import pytest

import util

G = np.load(os.path.join(util.GOLD, 'formulas.npz'))
LID = {'D2Q9': 0, 'D3Q19': 1, 'D3Q27': 2}
c_double_p = ctypes.POINTER(ctypes.c_double)


def P(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64).ctypes.data_as(ctypes.c_void_p)


@pytest.fixture(scope='module')
def lib(oracle):
    return oracle._lib(np.float64)


@pytest.mark.parametrize('name', sorted(LID))
def test_equilibrium(lib, name):
    rho, v, ref = G[name + '_rho'], G[name + '_v'], G[name + '_feq']
    out = np.zeros_like(ref)
    assert lib.orc_eval_feq(LID[name], rho.size, P(rho), P(v), out.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.abs(out - ref).max() < 1e-14
    # reference invariants (tests/sym_equilibrium.py:93-121): sum feq = rho, sum e feq = rho v
    basis = np.load(os.path.join(util.GOLD, 'lattices.npz'))[name + '_basis']
    assert np.abs(out.sum(0) - rho).max() < 1e-14
    assert np.abs(basis.T.astype(float) @ out - rho * v).max() < 1e-14


@pytest.mark.parametrize('name', sorted(LID))
def test_boundary_closures(lib, name):
    rho, v, f = G[name + '_rho'], G[name + '_v'], G[name + '_f']
    n = rho.size
    D = v.shape[0]
    for o in range(1, 2 * D + 1):
        out = np.zeros(n)
        assert lib.orc_eval_bc(LID[name], 0, o, n, None, P(rho), P(v), ctypes.c_double(0), out.ctypes.data_as(ctypes.c_void_p), None, None) == 0
        assert np.abs(out - G['%s_vbc_rho_%d' % (name, o)]).max() < 1e-13
        vout = np.zeros((D, n))
        assert lib.orc_eval_bc(LID[name], 1, o, n, None, P(rho), None, ctypes.c_double(1.05), None, vout.ctypes.data_as(ctypes.c_void_p), None) == 0
        assert np.abs(vout - G['%s_dbc_v_%d' % (name, o)]).max() < 1e-13
        fout = np.zeros_like(f)
        assert lib.orc_eval_bc(LID[name], 2, o, n, P(f), P(rho), P(v), ctypes.c_double(0), None, None, fout.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.abs(fout - G['%s_noneq_bb_%d' % (name, o)]).max() < 1e-13


@pytest.mark.parametrize('name', sorted(LID))
def test_regularisation(lib, name):
    rho, v, f = G[name + '_rho'], G[name + '_v'], G[name + '_f']
    fout = np.zeros_like(f)
    assert lib.orc_eval_bc(LID[name], 3, 1, rho.size, P(f), P(rho), P(v), ctypes.c_double(0), None, None, fout.ctypes.data_as(ctypes.c_void_p)) == 0
    want = np.maximum(1e-7, G[name + '_feq'] + G[name + '_reg'])
    assert np.abs(fout - want).max() < 1e-13


@pytest.mark.parametrize('name', sorted(LID))
def test_guo_forcing(lib, name):
    rho, v, a = G[name + '_rho'], G[name + '_v'], G[name + '_a']
    out = np.zeros_like(G[name + '_guo'])
    tau = float(G[name + '_guo_tau'])
    assert lib.orc_eval_guo(LID[name], rho.size, P(rho), P(v), P(a), ctypes.c_double(tau), out.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.abs(out - G[name + '_guo']).max() < 1e-15


@pytest.mark.parametrize('name', ['D2Q9', 'D3Q19'])
def test_mrt_moments_equilibria_rates(lib, name):
    f = G[name + '_f']
    Q, n = f.shape
    m, meq, rates = np.zeros((Q, n)), np.zeros((Q, n)), np.zeros(Q)
    assert lib.orc_eval_mrt(LID[name], n, P(f), ctypes.c_double(float(G[name + '_mrt_visc'])),
                            m.ctypes.data_as(ctypes.c_void_p), meq.ctypes.data_as(ctypes.c_void_p),
                            rates.ctypes.data_as(ctypes.c_void_p)) == 0
    assert np.abs(m - G[name + '_mrt_m']).max() < 1e-13
    assert np.abs(meq - G[name + '_mrt_eq']).max() < 1e-13
    assert np.abs(rates - G[name + '_mrt_rates']).max() < 1e-14
