"""The device-semaphore protocol has a failure mode other than hanging: a wait whose
neighbour never pushes gives up after SLF_SEM_TIMEOUT_MS (or when the host raises the abort word), the streams drain
and the next sync() raises.  Set up: two subdomains of a z-split cavity in ONE process, linked through the
cross-rank mechanism (device semaphores instead of CUDA events), and only one of them is ever stepped."""
import os
import subprocess
import sys
import textwrap

import pytest

import util

pytestmark = pytest.mark.gpu

SCRIPT = textwrap.dedent('''
    import sys, time
    sys.path.insert(0, %r)
    from sailfish_b200 import backend_b200 as bk
    from sailfish_b200.controller import LBSimulationController
    from sailfish_b200.examples.ldc_3d import LDCSim
    from sailfish_b200.geo import EqualSubdomainsGeometry3D
    from sailfish_b200.subdomain_runner import HaloLink

    mode = sys.argv[1]
    ctrl = LBSimulationController(LDCSim, EqualSubdomainsGeometry3D)
    ctrl.setup(['--lat_nx=20', '--lat_ny=18', '--lat_nz=24', '--subdomains=2', '--visc=0.04', '--access_pattern=AA', '--quiet'])
    a, b = ctrl.runners
    # replace the event-ordered links by semaphore-ordered ones (what two ranks would have)
    for me, other, face in ((a, b, 5), (b, a, 4)):
        me._links.clear(); me._xlinks.clear()
        g = dict(lat=list(reversed(other._lat_size)), arr=list(reversed(other._physical_size)))
        me.add_link(HaloLink(face, other.export_buffers(), g['lat'][2], g['arr'][2], peer_sems=int(other._sems),
                             my_sems=int(me._sems), peer_lat=g['lat'], peer_arr=g['arr']))
    assert a._uses_semaphores()
    a.step(False)                       # step 0 needs nothing from the neighbour
    a.step(False)                       # step 1 waits for b's push of step 0, which never comes
    t0 = time.time()
    if mode == 'abort':
        time.sleep(0.3)
        a.abort_waits()
    try:
        a.sync()
        print('NO ERROR')
    except bk.B200Error as e:
        print('B200Error after %%.2f s: %%s' %% (time.time() - t0, e))
''') % util.ROOT


@pytest.mark.parametrize('mode,env,needle', [('timeout', {'SLF_SEM_TIMEOUT_MS': '400'}, 'timed out'),
                                             ('abort', {'SLF_SEM_TIMEOUT_MS': '60000'}, 'aborted')])
def test_wait_gives_up_instead_of_hanging(mode, env, needle, tmp_path):
    script = tmp_path / 'sem.py'
    script.write_text(SCRIPT)
    res = subprocess.run([sys.executable, str(script), mode], env=dict(os.environ, **env), capture_output=True, text=True,
                         timeout=120)
    assert res.returncode == 0, res.stderr[-2000:]
    assert 'B200Error' in res.stdout and needle in res.stdout, (res.stdout, res.stderr[-1000:])
