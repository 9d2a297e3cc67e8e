"""Lattice tables of the product (sailfish_b200/lattice.py) and of the oracle against the
reference's own tables (tests/golden/lattices.npz <- sailfish/sym.py via tools/gen_golden.py),
plus the reference's tests/sym.py known answers for get_interblock_dists (tests/sym.py:6-66)."""
import os

import numpy as np
import pytest

import util
from sailfish_b200 import lattice

G = np.load(os.path.join(util.GOLD, 'lattices.npz'))


@pytest.mark.parametrize('name', lattice.KNOWN)
def test_tables_match_reference(name):
    L = lattice.get(name)
    assert (L.basis == G[name + '_basis']).all()
    assert (L.opposite == G[name + '_opposite']).all()
    assert list(L.idx_name) == [str(x) for x in G[name + '_names']]
    num, den = G[name + '_weights_num'], G[name + '_weights_den']
    assert [(w.numerator, w.denominator) for w in L.weights_frac] == list(zip(num.tolist(), den.tolist()))
    assert [L.dir2vecidx[d] for d in range(1, 2 * L.dim + 1)] == G[name + '_dir2vecidx'].tolist()
    assert L.bb_swap_pairs() == G[name + '_bb_swap'].tolist()
    for ax in range(L.dim):
        for d in (-1, 1):
            assert L.get_prop_dists(d, ax) == G['%s_prop_%d_%d' % (name, ax, d)].tolist()
            vec = [0] * L.dim
            vec[ax] = d
            assert L.get_interblock_dists(vec) == G['%s_interblock_%d_%d' % (name, ax, d)].tolist()
    for o in range(1, 2 * L.dim + 1):
        assert L.missing_dists(o) == G['%s_missing_%d' % (name, o)].tolist()
    if L.has_mrt:
        assert (L.mrt_matrix == G[name + '_mrt_matrix']).all()
        assert list(L.mrt_names) == [str(x) for x in G[name + '_mrt_names']]
        # rows are mutually orthogonal -> M^-1 = M^T diag(1/|row|^2)
        mm = L.mrt_matrix @ L.mrt_matrix.T
        assert (mm == np.diag(np.diag(mm))).all()


def test_d3q27_has_no_mrt():
    # sym.py:409-428 defines neither mrt_names nor _init_mrt_basis
    assert not lattice.get('D3Q27').has_mrt


def test_interblock_known_answers_d2q9():
    """Literal expectations of the reference's tests/sym.py:6-40 (names -> slots)."""
    L = lattice.get('D2Q9')
    name = {n: i for i, n in enumerate(L.idx_name)}

    def s(*names):
        return sorted(name[n] for n in names)
    assert sorted(L.get_interblock_dists([1, 0])) == s('fE', 'fNE', 'fSE')
    assert sorted(L.get_interblock_dists([-1, 0])) == s('fW', 'fNW', 'fSW')
    assert sorted(L.get_interblock_dists([0, 1])) == s('fN', 'fNE', 'fNW')
    assert sorted(L.get_interblock_dists([0, -1])) == s('fS', 'fSE', 'fSW')
    assert sorted(L.get_interblock_dists([1, 1])) == s('fNE')
    assert sorted(L.get_interblock_dists([-1, -1])) == s('fSW')


def test_interblock_known_answers_d3q19():
    """tests/sym.py:42-66."""
    L = lattice.get('D3Q19')
    name = {n: i for i, n in enumerate(L.idx_name)}

    def s(*names):
        return sorted(name[n] for n in names)
    assert sorted(L.get_interblock_dists([0, 0, 1])) == s('fT', 'fTN', 'fTS', 'fTE', 'fTW')
    assert sorted(L.get_interblock_dists([0, 0, -1])) == s('fB', 'fBN', 'fBS', 'fBE', 'fBW')
    assert sorted(L.get_interblock_dists([1, 1, 0])) == s('fNE')
    assert sorted(L.get_interblock_dists([0, -1, 1])) == s('fTS')


@pytest.mark.parametrize('lid,name', [(0, 'D2Q9'), (1, 'D3Q19'), (2, 'D3Q27')])
def test_oracle_tables_match_reference(oracle, lid, name):
    info = oracle.lattice_info(lid)
    assert info['Q'] == G[name + '_basis'].shape[0]
    assert (info['e'][:, :info['D']] == G[name + '_basis']).all()
    assert (info['opp'] == G[name + '_opposite']).all()
    w = G[name + '_weights_num'] / G[name + '_weights_den']
    assert np.allclose(info['w'], w, rtol=0, atol=1e-16)
    if name != 'D3Q27':
        assert (info['M'] == G[name + '_mrt_matrix']).all()


@pytest.mark.parametrize('name', lattice.KNOWN)
def test_generated_mirror_tables_match_reference_slip_pairs(name):
    """L::mirror(axis, i) of csrc/lattice_tables.cuh (tools/gen_lattice_header.py) against the reference's
    sym.slip_bb_swap_pairs (tests/golden/slip_pairs.npz <- tools/gen_golden_steps.py slip_pairs)."""
    import re
    gold = np.load(os.path.join(util.GOLD, 'slip_pairs.npz'))
    src = open(os.path.join(util.ROOT, 'sailfish_b200', 'csrc', 'lattice_tables.cuh')).read()
    body = src[src.index('struct %s {' % name):]
    m = re.search(r'mirror\(int a, int i\) \{ constexpr int t\[(\d+)\] = \{([^}]*)\}', body)
    L = lattice.get(name)
    table = np.array([int(x) for x in m.group(2).split(',')]).reshape(3, L.Q)
    assert int(m.group(1)) == 3 * L.Q
    for o in range(1, 2 * L.dim + 1):
        n = gold['%s_%d_normal' % (name, o)]
        (axis,) = np.nonzero(n)[0]
        pairs = sorted((i, int(table[axis, i])) for i in range(L.Q) if int(L.basis[i][axis]) * int(n[axis]) > 0)
        assert pairs == [tuple(p) for p in gold['%s_%d' % (name, o)].tolist()]
