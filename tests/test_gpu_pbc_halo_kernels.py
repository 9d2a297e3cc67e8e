"""Bit-exact checks of the data-movement kernels against the oracle on random lattices:
ApplyPeriodicBoundaryConditions[WithSwap], Collect/DistributeContinuousData[WithSwap] and the
peer-to-peer plane push (vs the oracle's collect + distribute), and the sparse index transfers."""
import ctypes

import numpy as np
import pytest

import util
from sailfish_b200 import lattice

pytestmark = pytest.mark.gpu


def _dev(case):
    from sailfish_b200.runtime import DeviceLattice
    c = util.load_case(case)
    p = util.params_from_case(c)
    return c, p, DeviceLattice(p, c['map'], c['rho0'], c['v0'])


@pytest.mark.parametrize('case', ['box2d_bgk_AA', 'box3d_bgk_AA', 'box3d_q27_AA', 'poi3d_force_bgk_AA'])
@pytest.mark.parametrize('swap', [0, 1])
def test_pbc_kernels_bit_exact(oracle, case, swap):
    c, p, dev = _dev(case)
    rng = np.random.RandomState(7)
    shape = (p.Q,) + c['map'].shape
    d0 = rng.rand(*shape).astype(np.float32)
    # sprinkle non-finite values: they must never be copied (kernel_utils.mako isfinite tests)
    d0[rng.rand(*shape) < 0.02] = np.inf
    for axis in range(p.dim):
        if not p.periodic[axis]:
            continue
        dev.set_dist(d0)
        dev.lib.call('slf_apply_periodic_boundary_conditions', dev.plan, int(dev.d_dist_a), axis, swap,
                     dev.stream.handle)
        g = dev.get_dist()
        o = d0.copy().reshape(-1)
        oracle.apply_pbc(p, o, axis, bool(swap))
        o = o.reshape(shape)
        same = (g == o) | (np.isnan(g) & np.isnan(o))
        assert same.all(), (axis, np.argwhere(~same)[:10].tolist())


@pytest.mark.parametrize('case', ['ldc2d_bgk_AA', 'ldc3d_bgk_AA', 'ldc3d_q27_AA'])
@pytest.mark.parametrize('swap', [0, 1])
def test_collect_distribute_bit_exact(oracle, case, swap):
    c, p, dev = _dev(case)
    rng = np.random.RandomState(11)
    shape = (p.Q,) + c['map'].shape
    d0 = rng.rand(*shape).astype(np.float32)
    lat = [int(x) for x in c['lat']]
    L = lattice.get(str(c['grid']))
    for face in range(2, 2 * p.dim):
        ndists = len(L.get_interblock_dists([(1 if face & 1 else -1) if a == face // 2 else 0 for a in range(p.dim)]))
        max_lx = lat[0] - 2
        max_other = 1 if p.dim == 2 else (lat[2] - 2 if face // 2 == 1 else lat[1] - 2)
        nbuf = ndists * max_other * max_lx
        dev.set_dist(d0)
        dbuf = dev.be.alloc_buf(size=nbuf * 4)
        dev.lib.call('slf_collect_continuous_data', dev.plan, int(dev.d_dist_a), face, 1, 1, max_lx, max_other,
                     int(dbuf), swap, dev.stream.handle)
        dev.sync()
        gbuf = np.zeros(nbuf, np.float32)
        dev.lib.call('slf_memcpy_d2h', gbuf.ctypes.data_as(ctypes.c_void_p), int(dbuf), nbuf * 4)
        obuf = np.zeros(nbuf, np.float32)
        oracle.collect(p, d0.reshape(-1), face, 1, 1, max_lx, max_other, obuf, bool(swap))
        assert np.array_equal(gbuf, obuf), face
        # distribute a fresh random buffer
        src = rng.rand(nbuf).astype(np.float32)
        dev.lib.call('slf_memcpy_h2d', int(dbuf), src.ctypes.data_as(ctypes.c_void_p), nbuf * 4)
        dev.lib.call('slf_distribute_continuous_data', dev.plan, int(dev.d_dist_a), face, 1, 1, max_lx, max_other,
                     int(dbuf), swap, dev.stream.handle)
        g = dev.get_dist()
        o = d0.copy().reshape(-1)
        oracle.distribute(p, o, face, 1, 1, max_lx, max_other, src, bool(swap))
        assert np.array_equal(g, o.reshape(shape)), face


@pytest.mark.parametrize('case', ['ldc2d_bgk_AA', 'ldc3d_bgk_AA'])
def test_macro_collect_distribute_bit_exact(oracle, case):
    """Collect/DistributeContinuousMacroData on a scalar field (the rho field of the lattice object)."""
    c, p, dev = _dev(case)
    rng = np.random.RandomState(13)
    shape = c['map'].shape
    lat = [int(x) for x in c['lat']]
    f0 = rng.rand(*shape).astype(np.float32)

    def put(a):
        dev.lib.call('slf_memcpy_h2d', int(dev.d_rho), a.ctypes.data_as(ctypes.c_void_p), a.nbytes)

    def get():
        out = np.zeros(shape, np.float32)
        dev.sync()
        dev.lib.call('slf_memcpy_d2h', out.ctypes.data_as(ctypes.c_void_p), int(dev.d_rho), out.nbytes)
        return out

    faces = [None] if p.dim == 2 else [2, 3, 4, 5]
    for face in faces:
        max_lx = lat[0] - 3
        if p.dim == 2:
            rows = [(1, 1), (lat[1] - 2, 1), (0, 1)]          # (gy, max_other)
        else:
            rows = [(1, (lat[2] if face // 2 == 1 else lat[1]) - 3)]
        for base_other, max_other in rows:
            nbuf = max_lx * max_other
            dbuf = dev.be.alloc_buf(size=nbuf * 4)
            put(f0)
            dev.lib.call('slf_collect_continuous_macro_data', dev.plan, int(dev.d_rho), face or 0, 2, base_other, max_lx,
                         max_other, int(dbuf), dev.stream.handle)
            dev.sync()
            gbuf = np.zeros(nbuf, np.float32)
            dev.lib.call('slf_memcpy_d2h', gbuf.ctypes.data_as(ctypes.c_void_p), int(dbuf), nbuf * 4)
            obuf = np.zeros(nbuf, np.float32)
            oracle.collect_macro(p, f0.reshape(-1), face or 0, 2, base_other, max_lx, max_other, obuf)
            assert np.array_equal(gbuf, obuf), (face, base_other)
            src = rng.rand(nbuf).astype(np.float32)
            dev.lib.call('slf_memcpy_h2d', int(dbuf), src.ctypes.data_as(ctypes.c_void_p), nbuf * 4)
            dev.lib.call('slf_distribute_continuous_macro_data', dev.plan, int(dev.d_rho), face or 0, 2, base_other,
                         max_lx, max_other, int(dbuf), dev.stream.handle)
            o = f0.copy().reshape(-1)
            oracle.distribute_macro(p, o, face or 0, 2, base_other, max_lx, max_other, src)
            assert np.array_equal(get(), o.reshape(shape)), (face, base_other)
    # out-of-range requests are refused, not clipped
    from sailfish_b200.backend_b200 import B200Error
    with pytest.raises(B200Error):
        dev.lib.call('slf_collect_continuous_macro_data', dev.plan, int(dev.d_rho), 2, 2, 1, lat[0], 1, int(dbuf),
                     dev.stream.handle)


@pytest.mark.parametrize('case', ['ldc2d_bgk_AA', 'ldc3d_bgk_AA', 'ldc3d_q27_AA'])
@pytest.mark.parametrize('swap', [0, 1])
@pytest.mark.parametrize('signal', [0, 1])
def test_halo_push_matches_oracle_collect_distribute(oracle, case, swap, signal):
    """slf_halo_push into a second lattice == the oracle's CollectContinuousData on the sender followed by
    DistributeContinuousData on the receiver (kernel_utils.mako:526-794 restated in oracle/lbm_oracle.c), over
    the whole lattice cross-section.  With `signal` the copy kernel's last block must have raised the
    semaphore word exactly once per launch."""
    c, p, dev = _dev(case)
    rng = np.random.RandomState(13)
    shape = (p.Q,) + c['map'].shape
    src = rng.rand(*shape).astype(np.float32)
    lat = [int(x) for x in c['lat']]
    nx = lat[0]
    n_other = 1 if p.dim == 2 else lat[1]
    outer = shape[1]
    L = lattice.get(str(c['grid']))
    peer = dev.be.alloc_buf(size=src.nbytes)
    sem = dev.be.alloc_buf(size=64)
    dev.lib.call('slf_memset', int(sem), 0, 64, 0)
    dev.lib.call('slf_sync')
    launches = 0
    for side in (0, 1):
        face = 2 * (p.dim - 1) + side
        ndists = len(L.get_interblock_dists([(1 if side else -1) if a == p.dim - 1 else 0 for a in range(p.dim)]))
        dst0 = rng.rand(*shape).astype(np.float32)
        dev.set_dist(src)
        dev.lib.call('slf_memcpy_h2d', int(peer), dst0.ctypes.data_as(ctypes.c_void_p), dst0.nbytes)
        for _ in range(3):                       # repeated launches: the block counter must re-arm itself
            dev.lib.call('slf_halo_push', dev.plan, int(dev.d_dist_a), face, int(peer), lat[p.dim - 1], outer, swap,
                         int(sem) if signal else 0, dev.stream.handle)
            launches += 1
        dev.sync()
        got = np.zeros_like(dst0)
        dev.lib.call('slf_memcpy_d2h', got.ctypes.data_as(ctypes.c_void_p), int(peer), got.nbytes)
        buf = np.zeros(ndists * n_other * nx, np.float32)
        oracle.collect(p, src.reshape(-1), face, 0, 0, nx, n_other, buf, bool(swap))
        want = dst0.copy().reshape(-1)
        oracle.distribute(p, want, face, 0, 0, nx, n_other, buf, bool(swap))
        want = want.reshape(shape)
        assert np.array_equal(got[..., :nx], want[..., :nx]), (face, swap)
        # nothing but the target planes changed (x padding of those planes travels too and is never read)
        changed = np.any((got != dst0)[..., :nx].reshape(p.Q, outer, -1), axis=2)
        assert changed.sum() == ndists, (face, swap, changed.sum())
    word = np.zeros(1, np.uint32)
    dev.lib.call('slf_memcpy_d2h', word.ctypes.data_as(ctypes.c_void_p), int(sem), 4)
    assert int(word[0]) == (launches if signal else 0)


@pytest.mark.parametrize('case', ['ldc2d_bgk_AA', 'ldc3d_bgk_AA', 'ldc3d_q27_AA'])
def test_sparse_collect_distribute_bit_exact(case):
    """Collect/DistributeSparseData (kernel_utils.mako:796-836): buffer[k] = dist[idx[k]] and back, for the
    index sets the reference builds for x faces, edges and corners (global population indices, some
    repeated, 0xffffffff = padding entry that must be skipped)."""
    c, p, dev = _dev(case)
    rng = np.random.RandomState(17)
    shape = (p.Q,) + c['map'].shape
    d0 = rng.rand(*shape).astype(np.float32)
    total = d0.size
    n = min(5000, total // 2)
    idx = rng.randint(0, total, size=n).astype(np.uint32)
    idx[rng.rand(n) < 0.05] = 0xffffffff
    didx = dev.be.alloc_buf(size=idx.nbytes)
    dbuf = dev.be.alloc_buf(size=n * 4)
    dev.lib.call('slf_memcpy_h2d', int(didx), idx.ctypes.data_as(ctypes.c_void_p), idx.nbytes)
    dev.set_dist(d0)
    sentinel = np.full(n, -7.0, np.float32)
    dev.lib.call('slf_memcpy_h2d', int(dbuf), sentinel.ctypes.data_as(ctypes.c_void_p), n * 4)
    dev.lib.call('slf_collect_sparse_data', dev.plan, int(didx), int(dev.d_dist_a), int(dbuf), n, dev.stream.handle)
    dev.sync()
    got = np.zeros(n, np.float32)
    dev.lib.call('slf_memcpy_d2h', got.ctypes.data_as(ctypes.c_void_p), int(dbuf), n * 4)
    valid = idx != 0xffffffff
    want = sentinel.copy()
    want[valid] = d0.reshape(-1)[idx[valid]]
    assert np.array_equal(got, want)
    # distribute: unique targets so that the result does not depend on the write order
    uidx = rng.permutation(total)[:n].astype(np.uint32)
    uidx[::17] = 0xffffffff
    src = rng.rand(n).astype(np.float32)
    dev.lib.call('slf_memcpy_h2d', int(didx), uidx.ctypes.data_as(ctypes.c_void_p), uidx.nbytes)
    dev.lib.call('slf_memcpy_h2d', int(dbuf), src.ctypes.data_as(ctypes.c_void_p), n * 4)
    dev.lib.call('slf_distribute_sparse_data', dev.plan, int(didx), int(dev.d_dist_a), int(dbuf), n, dev.stream.handle)
    g = dev.get_dist()
    o = d0.copy().reshape(-1)
    ok = uidx != 0xffffffff
    o[uidx[ok]] = src[ok]
    assert np.array_equal(g, o.reshape(shape))
