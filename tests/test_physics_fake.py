"""The physics regtest harness (tests/physics.py) on the CPU stand-in device at reduced size: checks the
harness itself (profile extraction, theory, reference table look-up) without a GPU."""
import pytest

import fake_device
import physics


@pytest.fixture
def fake(oracle, monkeypatch):
    return fake_device.install_fake(monkeypatch)


def test_poiseuille_harness_small(fake):
    r = physics.poiseuille3d_max_velocity_error(0.1, 'bgk', width=24, length=4, steps=400)
    ref = physics.reference_poiseuille_error(0.1, 'bgk')
    assert r['finite'] and abs(r['err']) < 1e-2 and abs(r['err'] - ref) < 8e-3, (r, ref)


def test_ldc_harness_small(fake):
    """20^3 for a few hundred steps: nowhere near steady, but the extraction and the golden look-up run."""
    r = physics.ldc3d_re400(20, 300)
    assert r['finite'] and r['res_vx'].shape == (20,) and r['res_vx'][-2] > 0.3
    assert 0 <= r['dev_vx'] < 1.5 and 0 <= r['dev_vz'] < 1.5
