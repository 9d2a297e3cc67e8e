"""Geometry set-ups shared by the golden generator (run with the REFERENCE's classes, tools/gen_golden_steps.py
zoo) and tests/test_host_geometry.py (run with this package's mirror): same user code, two implementations of
Subdomain.reset() + encoding.  The scenarios follow the reference's tests/subdomain.py (solid interior nodes,
duct, channels with periodic directions, link tagging) and add the node types the examples do not use."""
import numpy as np

# set-ups whose node types the reference only generates code for in the AB access pattern (boundary.mako:626-628)
AB_ONLY = ('copy_channel2d', 'yu_channel2d', 'copy_duct3d', 'yu_duct3d_mrt', 'copy_duct3d_q27')
# link-bounce walls (half-way bounce-back, TMS) that run along a periodic axis up to the periodic face: in the AA
# pattern the bounced populations of the wall nodes next to that face pass through ghost slots that the periodic
# copy kernels also write, and the reference's AA == AB invariant does not hold there.  Only the AB run is used as a
# pin.  (Without link tags such walls have no step goldens at all: tools/gen_golden_steps.py.)
AB_PINNED_ONLY = ('duct3d', 'tms_channel2d')


def make_cases(ns):
    """ns: namespace with Subdomain2D, Subdomain3D, LBFluidSim, LBForcedSim and the NT* node types."""
    S2, S3, Sim = ns['Subdomain2D'], ns['Subdomain3D'], ns['LBFluidSim']
    NTFullBBWall, NTHalfBBWall = ns['NTFullBBWall'], ns['NTHalfBBWall']
    NTEquilibriumVelocity, NTEquilibriumDensity = ns['NTEquilibriumVelocity'], ns['NTEquilibriumDensity']
    NTRegularizedVelocity, NTRegularizedDensity = ns['NTRegularizedVelocity'], ns['NTRegularizedDensity']
    NTDoNothing = ns['NTDoNothing']
    NTZouHeVelocity, NTZouHeDensity = ns['NTZouHeVelocity'], ns['NTZouHeDensity']
    NTCopy, NTYuOutflow = ns['NTCopy'], ns['NTYuOutflow']
    NTWallTMS = ns['NTWallTMS']
    NTSlip = ns['NTSlip']
    Forced = ns['LBForcedSim']

    def sim(sub, **defaults):
        class _Sim(Sim):
            subdomain = sub

            @classmethod
            def update_defaults(cls, d):
                d.update(defaults)
        return _Sim

    def forced_sim(sub, accel, **defaults):
        class _Forced(Sim, Forced):
            subdomain = sub

            def __init__(self, config):
                super(_Forced, self).__init__(config)
                self.add_body_force(accel)

            @classmethod
            def update_defaults(cls, d):
                d.update(defaults)
        return _Forced

    class TMSChannel2D(S2):
        """The reference's tests/gpu/tms.py set-up (TMS walls at both y ends, periodic along x), two nodes thick so
        that the layer next to the ghosts is unused, with a block on the lower wall: the walls have nodes with one,
        two and three links into the fluid."""
        def boundary_conditions(self, hx, hy):
            wall = (hy <= 1) | (hy >= self.gy - 2) | ((hx >= 6) & (hx <= 9) & (hy <= 4))
            self.set_node(wall, NTWallTMS)

        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0 + 0.004 * np.cos(2 * np.pi * hx / self.gx)
            sim.vx[:] = 0.05 * np.sin(np.pi * (hy - 1) / (self.gy - 3))
            sim.vy[:] = 0.004 * np.sin(2 * np.pi * hx / self.gx)

    class TMSBlocks2D(S2):
        """Two TMS blocks inside a periodic box (an L-shaped one: concave and convex corners), away from the
        periodic faces."""
        def boundary_conditions(self, hx, hy):
            wall = ((hx >= 4) & (hx <= 8) & (hy >= 3) & (hy <= 6)) | ((hx >= 7) & (hx <= 8) & (hy >= 7) & (hy <= 10))
            wall |= (hx >= 12) & (hx <= 14) & (hy >= 8) & (hy <= 11)
            self.set_node(wall, NTWallTMS)

        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0 + 0.004 * np.cos(2 * np.pi * hx / self.gx) * np.sin(2 * np.pi * hy / self.gy)
            sim.vx[:] = 0.04 * np.sin(2 * np.pi * hy / self.gy)
            sim.vy[:] = -0.02 * np.cos(2 * np.pi * hx / self.gx)

    class TMSBlock3D(S3):
        """A TMS block with a step inside a fully periodic box."""
        def boundary_conditions(self, hx, hy, hz):
            wall = (hx >= 4) & (hx <= 7) & (hy >= 3) & (hy <= 6) & (hz >= 3) & (hz <= 5)
            wall |= (hx >= 4) & (hx <= 5) & (hy >= 3) & (hy <= 4) & (hz >= 6) & (hz <= 7)
            self.set_node(wall, NTWallTMS)

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0 + 0.003 * np.cos(2 * np.pi * hx / self.gx) * np.sin(2 * np.pi * hz / self.gz)
            sim.vx[:] = 0.04 * np.sin(2 * np.pi * hy / self.gy)
            sim.vy[:] = 0.01 * np.cos(2 * np.pi * hz / self.gz)
            sim.vz[:] = 0.005 * np.sin(2 * np.pi * hx / self.gx)

    class SlipChannel2D(S2):
        """Full-slip walls at both y ends (the node type takes its orientation from the user), periodic along x."""
        def boundary_conditions(self, hx, hy):
            self.set_node(hy == 0, NTSlip(orientation=self.grid.vec_to_dir([0, 1])))
            self.set_node(hy == self.gy - 1, NTSlip(orientation=self.grid.vec_to_dir([0, -1])))

        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0 + 0.004 * np.cos(2 * np.pi * hx / self.gx)
            sim.vx[:] = 0.04
            sim.vy[:] = 0.02 * np.sin(2 * np.pi * hx / self.gx) * np.sin(np.pi * hy / (self.gy - 1))

    class SlipBox3D(S3):
        """Full-slip planes at both z ends and at y = 0, full-way walls on the other faces and along the edges where
        two planes meet; a jet towards the z = top plane."""
        def boundary_conditions(self, hx, hy, hz):
            wall = (hx == 0) | (hx == self.gx - 1) | (hy == self.gy - 1)
            wall |= (hy == 0) & ((hz == 0) | (hz == self.gz - 1))
            self.set_node(wall, NTFullBBWall)
            free = np.logical_not(wall)
            self.set_node((hz == 0) & free, NTSlip(orientation=self.grid.vec_to_dir([0, 0, 1])))
            self.set_node((hz == self.gz - 1) & free, NTSlip(orientation=self.grid.vec_to_dir([0, 0, -1])))
            self.set_node((hy == 0) & free, NTSlip(orientation=self.grid.vec_to_dir([0, 1, 0])))

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0 + 0.003 * np.cos(0.5 * hx) * np.sin(0.4 * hy)
            sim.vx[:] = 0.02 * np.sin(np.pi * hz / (self.gz - 1))
            sim.vz[:] = 0.03 * np.sin(np.pi * hx / (self.gx - 1)) * np.sin(np.pi * hy / (self.gy - 1))
            sim.vy[:] = -0.015 * np.sin(np.pi * hx / (self.gx - 1))

    class SolidBox2D(S2):
        """A half-way bounce-back block inside a periodic box: shell stays, inside becomes propagation-only /
        unused."""
        def boundary_conditions(self, hx, hy):
            self.set_node((hx >= 5) & (hx <= 10) & (hy >= 4) & (hy <= 9), NTHalfBBWall)

        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0 + 0.01 * np.cos(0.4 * hx) * np.sin(0.3 * hy)
            sim.vx[:] = 0.02 * np.sin(0.5 * hy)
            sim.vy[:] = -0.01 * np.cos(0.3 * hx)

    class Channel2D(S2):
        """Velocity inlet, do-nothing outlet, full-way walls, a step on the lower wall."""
        def boundary_conditions(self, hx, hy):
            wall = (hy == 0) | (hy == self.gy - 1) | ((hx >= 6) & (hx <= 9) & (hy <= 3))
            self.set_node(wall, NTFullBBWall)
            self.set_node((hx == 0) & np.logical_not(wall), NTEquilibriumVelocity((0.03, 0.0)))
            self.set_node((hx == self.gx - 1) & np.logical_not(wall), NTDoNothing)

        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0
            sim.vx[:] = 0.03

    class Duct3D(S3):
        """Half-way walls on the y and z faces, periodic along x (link tags, orientation on edges)."""
        def boundary_conditions(self, hx, hy, hz):
            self.set_node((hy == 0) | (hy == self.gy - 1) | (hz == 0) | (hz == self.gz - 1), NTHalfBBWall)

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0 + 0.005 * np.sin(2 * np.pi * hx / self.gx)
            sim.vx[:] = 0.03 * np.sin(np.pi * hy / (self.gy - 1)) * np.sin(np.pi * hz / (self.gz - 1))
            sim.vz[:] = 0.005 * np.cos(2 * np.pi * hx / self.gx)

    class PressureBox3D(S3):
        """Regularised density planes at both x ends, regularised velocity on the z = top plane, full-way walls
        elsewhere, a full-way obstacle in the middle."""
        def boundary_conditions(self, hx, hy, hz):
            wall = (hy == 0) | (hy == self.gy - 1) | (hz == 0)
            wall |= (abs(hx - self.gx // 2) <= 1) & (abs(hy - self.gy // 2) <= 1) & (hz <= 3)
            wall |= (hz == self.gz - 1) & ((hx == 0) | (hx == self.gx - 1))     # no open edge between lid and x planes
            lid = (hz == self.gz - 1) & np.logical_not(wall)
            self.set_node(wall, NTFullBBWall)
            self.set_node(lid, NTRegularizedVelocity((0.0, 0.02, 0.0)))
            free = np.logical_not(wall | lid)
            self.set_node((hx == 0) & free, NTRegularizedDensity(1.01))
            self.set_node((hx == self.gx - 1) & free, NTRegularizedDensity(0.99))

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0 + 0.01 * (self.gx / 2.0 - hx) / self.gx

    class SlabPeriodicYZ(S3):
        """Full-way walls at both x ends, periodic in y and z, equilibrium-density patch on one wall."""
        def boundary_conditions(self, hx, hy, hz):
            patch = (hx == 0) & (hy >= 3) & (hy <= 5) & (hz >= 2) & (hz <= 4)
            self.set_node(((hx == 0) | (hx == self.gx - 1)) & np.logical_not(patch), NTFullBBWall)
            self.set_node(patch, NTEquilibriumDensity(1.02))

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0

    class ZouHeChannel2D(S2):
        """Zou-He velocity inlet, Zou-He density outlet, full-way walls with a step."""
        def boundary_conditions(self, hx, hy):
            wall = (hy == 0) | (hy == self.gy - 1) | ((hx >= 7) & (hx <= 9) & (hy <= 3))
            self.set_node(wall, NTFullBBWall)
            self.set_node((hx == 0) & np.logical_not(wall), NTZouHeVelocity((0.03, 0.0)))
            self.set_node((hx == self.gx - 1) & np.logical_not(wall), NTZouHeDensity(1.0))

        def initial_conditions(self, sim, hx, hy):
            sim.rho[:] = 1.0
            sim.vx[:] = 0.03

    class ZouHeBox3D(S3):
        """Zou-He density planes at both x ends, Zou-He velocity on the z = top plane, full-way walls elsewhere."""
        def boundary_conditions(self, hx, hy, hz):
            wall = (hy == 0) | (hy == self.gy - 1) | (hz == 0)
            wall |= (hz == self.gz - 1) & ((hx == 0) | (hx == self.gx - 1))
            lid = (hz == self.gz - 1) & np.logical_not(wall)
            self.set_node(wall, NTFullBBWall)
            self.set_node(lid, NTZouHeVelocity((0.0, 0.02, 0.0)))
            free = np.logical_not(wall | lid)
            self.set_node((hx == 0) & free, NTZouHeDensity(1.01))
            self.set_node((hx == self.gx - 1) & free, NTZouHeDensity(0.99))

        def initial_conditions(self, sim, hx, hy, hz):
            sim.rho[:] = 1.0 + 0.01 * (self.gx / 2.0 - hx) / self.gx

    def outflow_channel2d(outlet):
        class OutflowChannel2D(S2):
            """Equilibrium-velocity inlet, copy / extrapolation outlet (AB pattern only), full-way walls, a step."""
            def boundary_conditions(self, hx, hy):
                wall = (hy == 0) | (hy == self.gy - 1) | ((hx >= 6) & (hx <= 8) & (hy <= 3))
                self.set_node(wall, NTFullBBWall)
                self.set_node((hx == 0) & np.logical_not(wall), NTEquilibriumVelocity((0.03, 0.0)))
                self.set_node((hx == self.gx - 1) & np.logical_not(wall), outlet)

            def initial_conditions(self, sim, hx, hy):
                sim.rho[:] = 1.0 + 0.002 * np.cos(0.5 * hx)
                sim.vx[:] = 0.03 * np.sin(np.pi * hy / (self.gy - 1))
        return OutflowChannel2D

    def outflow_duct3d(outlet):
        class OutflowDuct3D(S3):
            """Regularised-velocity inlet at x = 0, copy / extrapolation outlet at the far x end and on the z = top
            plane (two outlet orientations), full-way walls elsewhere."""
            def boundary_conditions(self, hx, hy, hz):
                wall = (hy == 0) | (hy == self.gy - 1) | (hz == 0)
                wall |= (hz == self.gz - 1) & ((hx == 0) | (hx == self.gx - 1))
                top = (hz == self.gz - 1) & np.logical_not(wall)
                self.set_node(wall, NTFullBBWall)
                self.set_node(top, outlet)
                free = np.logical_not(wall | top)
                self.set_node((hx == 0) & free, NTRegularizedVelocity((0.02, 0.0, 0.0)))
                self.set_node((hx == self.gx - 1) & free, outlet)

            def initial_conditions(self, sim, hx, hy, hz):
                sim.rho[:] = 1.0 + 0.003 * np.sin(0.6 * hx) * np.cos(0.4 * hy)
                sim.vx[:] = 0.02
                sim.vz[:] = 0.004 * np.sin(0.5 * hx)
        return OutflowDuct3D

    return {
        'slip_channel2d': (sim(SlipChannel2D, lat_nx=16, lat_ny=11, grid='D2Q9', periodic_x=True), []),
        'slip_box3d': (sim(SlipBox3D, lat_nx=11, lat_ny=10, lat_nz=9, grid='D3Q19'), []),
        'slip_box3d_mrt': (sim(SlipBox3D, lat_nx=11, lat_ny=10, lat_nz=9, grid='D3Q19'), ['--model=mrt']),
        'slip_box3d_q27': (sim(SlipBox3D, lat_nx=11, lat_ny=10, lat_nz=9, grid='D3Q27'), []),
        # walls along a periodic axis: pinned in the AB pattern only (see AB_PINNED_ONLY)
        'tms_channel2d': (sim(TMSChannel2D, lat_nx=16, lat_ny=14, grid='D2Q9', periodic_x=True), []),
        'tms_channel2d_notags': (sim(TMSChannel2D, lat_nx=16, lat_ny=14, grid='D2Q9', periodic_x=True),
                                 ['--nouse_link_tags']),
        'tms_blocks2d': (sim(TMSBlocks2D, lat_nx=18, lat_ny=15, grid='D2Q9', periodic_x=True, periodic_y=True), []),
        'tms_block3d': (sim(TMSBlock3D, lat_nx=12, lat_ny=10, lat_nz=11, grid='D3Q19', periodic_x=True,
                            periodic_y=True, periodic_z=True), []),
        'tms_block3d_mrt_forced': (forced_sim(TMSBlock3D, (2e-4, 0.0, -1e-4), lat_nx=12, lat_ny=10, lat_nz=11,
                                              grid='D3Q19', periodic_x=True, periodic_y=True, periodic_z=True),
                                   ['--model=mrt']),
        'tms_block3d_q27': (sim(TMSBlock3D, lat_nx=12, lat_ny=10, lat_nz=11, grid='D3Q27', periodic_x=True,
                                periodic_y=True, periodic_z=True), []),
        'copy_channel2d': (sim(outflow_channel2d(NTCopy), lat_nx=18, lat_ny=11, grid='D2Q9'), []),
        'yu_channel2d': (sim(outflow_channel2d(NTYuOutflow), lat_nx=18, lat_ny=11, grid='D2Q9'), []),
        'copy_duct3d': (sim(outflow_duct3d(NTCopy), lat_nx=12, lat_ny=9, lat_nz=8, grid='D3Q19'), []),
        'yu_duct3d_mrt': (sim(outflow_duct3d(NTYuOutflow), lat_nx=12, lat_ny=9, lat_nz=8, grid='D3Q19'),
                          ['--model=mrt']),
        'copy_duct3d_q27': (sim(outflow_duct3d(NTCopy), lat_nx=12, lat_ny=9, lat_nz=8, grid='D3Q27'), []),
        'zouhe_channel2d': (sim(ZouHeChannel2D, lat_nx=20, lat_ny=12, grid='D2Q9'), []),
        'zouhe_box3d': (sim(ZouHeBox3D, lat_nx=12, lat_ny=10, lat_nz=9, grid='D3Q19'), []),
        'zouhe_box3d_mrt': (sim(ZouHeBox3D, lat_nx=12, lat_ny=10, lat_nz=9, grid='D3Q19'), ['--model=mrt']),
        'zouhe_box3d_q27': (sim(ZouHeBox3D, lat_nx=12, lat_ny=10, lat_nz=9, grid='D3Q27'), []),
        'solid_box2d': (sim(SolidBox2D, lat_nx=16, lat_ny=14, grid='D2Q9', periodic_x=True, periodic_y=True), []),
        'channel2d': (sim(Channel2D, lat_nx=20, lat_ny=12, grid='D2Q9'), []),
        'duct3d': (sim(Duct3D, lat_nx=9, lat_ny=8, lat_nz=7, grid='D3Q19', periodic_x=True), []),
        'duct3d_notags': (sim(Duct3D, lat_nx=9, lat_ny=8, lat_nz=7, grid='D3Q19', periodic_x=True), ['--nouse_link_tags']),
        'pressure_box3d': (sim(PressureBox3D, lat_nx=12, lat_ny=10, lat_nz=9, grid='D3Q19'), []),
        'pressure_box3d_q27': (sim(PressureBox3D, lat_nx=12, lat_ny=10, lat_nz=9, grid='D3Q27'), []),
        'slab_periodic_yz': (sim(SlabPeriodicYZ, lat_nx=8, lat_ny=9, lat_nz=7, grid='D3Q19', periodic_y=True,
                                 periodic_z=True), []),
    }
