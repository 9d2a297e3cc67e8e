"""The recorded plugin-mode call streams (tests/golden/trace_*.pkl.gz, tools/gen_plugin_traces.py) replayed on
the CPU stand-in device must reproduce what was recorded exactly: validates the record / replay machinery that
tests/test_gpu_plugin_replay.py uses against the real library."""
import glob
import os

import pytest

import abi_trace
import fake_device
import util

TRACES = sorted(glob.glob(os.path.join(util.GOLD, 'trace_*.pkl.gz')))


def test_fixtures_exist():
    assert len(TRACES) >= 5


@pytest.mark.parametrize('path', TRACES, ids=[os.path.basename(p)[6:-7] for p in TRACES])
def test_replay_on_stand_in_device_is_exact(oracle, monkeypatch, path):
    fake = fake_device.install_fake(monkeypatch)
    ops = abi_trace.load_fixture(path)
    res = abi_trace.replay(ops, fake)
    assert res, 'trace has no device -> host copy to compare'
    for n, name, err, mask_ok, nfin in res:
        assert mask_ok and err == 0.0, (n, name, err)
