#!/usr/bin/env python
"""Copies the reference-held physics tables the north star's regtests compare with into one small fixture
(tests/golden/physics.npz); runs in the build container only (/root/reference does not exist on the GPU box).

  regtest/ldc_golden/re400_horiz, re400_vert   centre-line profiles of the Re = 400 cubic cavity digitised from
                                               Sheu & Tsai, Computers & Fluids 31 (2002) 911 (regtest/ldc_3d.py:84-88)
  regtest/results/poiseuille3d/D3Q19/{bgk,mrt}/force/single/fullbb.dat
                                               max(v_sim)/max(v_theory) - 1 of the reference itself over 30 viscosities
"""
import os

import numpy as np

REF = '/root/reference/regtest'
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'physics.npz')

data = {
    'ldc_re400_horiz': np.loadtxt(os.path.join(REF, 'ldc_golden', 're400_horiz'), skiprows=1),
    'ldc_re400_vert': np.loadtxt(os.path.join(REF, 'ldc_golden', 're400_vert'), skiprows=1),
    'poiseuille3d_bgk_force_fullbb': np.loadtxt(os.path.join(REF, 'results/poiseuille3d/D3Q19/bgk/force/single/fullbb.dat')),
    'poiseuille3d_mrt_force_fullbb': np.loadtxt(os.path.join(REF, 'results/poiseuille3d/D3Q19/mrt/force/single/fullbb.dat')),
}
np.savez(OUT, **data)
print(OUT, {k: v.shape for k, v in data.items()})
