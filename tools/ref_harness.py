"""Builds a reference (sailfish) SubdomainRunner for a simulation class without
any GPU, following the sequence of SubdomainRunner.run()
(/root/reference/sailfish/subdomain_runner.py:1537-1558) and the recipe of
tests/common.py in the reference.  Used ONLY in the build container (where
/root/reference exists) to generate golden fixtures; never at GPU-test time.
"""
import importlib.util
import os
import sys

REF = os.environ.get('SAILFISH_REFERENCE', '/root/reference')
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from sailfish_b200 import compat
compat.install(REF)

import numpy as np  # noqa: E402


class _Quit(object):
    def is_set(self):
        return False


def load_example(relpath, name=None):
    path = os.path.join(REF, relpath)
    name = name or ('refex_' + os.path.basename(relpath)[:-3])
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def make_runners(lb_class, lb_geo=None, args=(), defaults=None, backend=None):
    """Returns (config, [runner...]) with geometry + fields initialised."""
    from sailfish import io, util
    from sailfish.backend_dummy import DummyBackend
    from sailfish.controller import LBSimulationController, LBGeometryProcessor
    from sailfish.subdomain_runner import SubdomainRunner, NNSubdomainRunner

    dflt = {'quiet': True, 'mode': 'batch', 'output': '', 'log': ''}
    if defaults:
        dflt.update(defaults)
    ctrl = LBSimulationController(lb_class, lb_geo, default_config=dflt)
    cfg = ctrl._config_parser.parse(list(args), internal_defaults={'quiet': True})
    ctrl.config = cfg
    lb_class.modify_config(cfg)
    ctrl.geo = ctrl._lb_geo(cfg)
    specs = ctrl.geo.subdomains()
    ctrl._init_subdomain_envelope(lb_class, specs)
    proc = LBGeometryProcessor(specs, ctrl.dim, ctrl.geo.gsize)
    specs = proc.transform(cfg)
    cfg.logger = util.setup_logger(cfg) if hasattr(util, 'setup_logger') else None
    runners = []
    if backend is None:
        backend = DummyBackend()
        if not hasattr(backend, 'supports_printf'):
            backend.supports_printf = False
        if not hasattr(backend, 'name'):
            backend.name = 'dummy'
    for spec in specs:
        sim = lb_class(cfg)
        cls = lb_class.subdomain_runner or SubdomainRunner
        r = cls(sim, spec, output=io.LBOutput(cfg, spec.id),
                backend=backend, quit_event=_Quit())
        r._init_geometry()
        sim.init_fields(r)
        r._subdomain.init_fields(sim)
        runners.append(r)
    return cfg, runners


def context_of(runner):
    ctx = {}
    runner._sim.update_context(ctx)
    runner.update_context(ctx)
    return ctx


if __name__ == '__main__':
    ex = load_example('examples/ldc_3d.py')
    cfg, rs = make_runners(ex.LDCSim, args=['--lat_nx=8', '--lat_ny=8', '--lat_nz=8',
                                            '--access_pattern=AA'])
    r = rs[0]
    ctx = context_of(r)
    m = r._subdomain.encoded_map()
    print(m.shape, r._physical_size, r._lat_size)
    for k in ('type_id_remap', 'nt_misc_shift', 'nt_type_mask', 'nt_param_shift',
              'nt_scratch_shift', 'node_params', 'tau', 'visc', 'bnd_limits',
              'arr_nx', 'arr_ny', 'lat_nx', 'dist_size', 'access_pattern',
              'block_periodicity', 'boundary_size', 'grid_nx', 'envelope_size'):
        print(k, ctx.get(k))
    print(sorted(n.__name__ for n in ctx['node_types']))
    print(m[:, :, 1])
    print(np.unique(m))
