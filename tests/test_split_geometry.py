"""Subdomain decomposition along the outermost axis: the mirror's specs, encoded maps and
fluid masks of every subdomain must equal the reference's (EqualSubdomainsGeometry*,
geo.py:69-135; ghost handling at connected faces, subdomain.py:978-1005)."""
import os

import numpy as np
import pytest

import util
from sailfish_b200.controller import LBSimulationController
from sailfish_b200.geo import EqualSubdomainsGeometry2D, EqualSubdomainsGeometry3D
from sailfish_b200.examples import ldc_2d, ldc_3d, poiseuille_3d

CASES = {
    'split_ldc3d_z3': (ldc_3d.LDCSim, EqualSubdomainsGeometry3D,
                       ['--lat_nx=10', '--lat_ny=9', '--lat_nz=14', '--subdomains=3', '--access_pattern=AA']),
    'split_ldc2d_y2': (ldc_2d.LDCSim, EqualSubdomainsGeometry2D,
                       ['--lat_nx=12', '--lat_ny=11', '--subdomains=2', '--access_pattern=AA']),
    'split_poi3d_z2': (poiseuille_3d.PoiseuilleSim, EqualSubdomainsGeometry3D,
                       ['--lat_nx=10', '--lat_ny=10', '--lat_nz=12', '--flow_direction=z', '--drive=force',
                        '--subdomains=2', '--access_pattern=AA']),
}


def build_all(name):
    cls, geo, args = CASES[name]
    ctrl = LBSimulationController(cls, geo)
    cfg = ctrl._parser.parse(args)
    ctrl.config = cfg
    cls.modify_config(cfg)
    specs = ctrl._make_specs()
    out = []
    for spec in specs:
        sim = cls(cfg)
        gsize = (cfg.lat_ny, cfg.lat_nx) if ctrl.dim == 2 else (cfg.lat_nz, cfg.lat_ny, cfg.lat_nx)
        sd = cls.subdomain(gsize, spec, sim.grid, cfg)
        sd.reset()
        out.append((spec, sd))
    return out


@pytest.mark.parametrize('name', sorted(CASES))
def test_split_maps_match_reference(name):
    g = np.load(os.path.join(util.GOLD, 'geometry_%s.npz' % name))
    parts = build_all(name)
    assert len(parts) == int(g['n'])
    for i, (spec, sd) in enumerate(parts):
        assert tuple(spec.location) == tuple(int(x) for x in g['location_%d' % i])
        assert tuple(spec.size) == tuple(int(x) for x in g['size_%d' % i])
        assert sd.encoded_map().shape == g['map_%d' % i].shape
        assert (sd.encoded_map() == g['map_%d' % i]).all(), i
        assert (sd.fluid_map() == g['fluid_map_%d' % i]).all()
        ref_ids = {str(n): int(t) for n, t in zip(g['type_names_%d' % i], g['type_ids_%d' % i])}
        assert sd.type_ids_by_name() == ref_ids
        assert [bool(x) for x in g['periodic_%d' % i]][:spec.dim] == list(spec._periodicity)
