"""FreeIntegration plugin -- device-backed mirror of
demo_algorithms/free_integration.py:15-186 (same constructor, .input/.output/.batch,
.run(set_of_input)/.get_results()/.reset(), same run_times -> initial-state-set rule),
so it can be handed to the UNMODIFIED reference Sim as `algorithm=` as well as to
gnss_ins_sim_b200.sim.Sim.  Adds run_batch() for R runs in one launch.

The recurrence is the reference's forward-Euler ZYX-Euler-angle integrator (NOT a
quaternion integrator, SURVEY section 0 finding 1); it runs in csrc/mc_kernel.cuh (K2).
There is no NumPy fallback: without the CUDA library / a GPU, run() raises.
"""
import numpy as np

from . import engine


class FreeIntegration(object):
    '''
    Integrate gyro to get attitude, double integrate linear acceleration to get position.
    '''

    def __init__(self, ini_pos_vel_att, earth_rot=True, lanes_per_run=0):
        '''
        Args:
            ini_pos_vel_att: (9,) or (10,) array, or (9|10, S) for S sets of initial states:
                LLA position [rad, rad, m], body-frame velocity [m/s], ZYX Euler angles
                [yaw, pitch, roll] rad; optional 10th row: gravity [m/s^2] used instead of
                the WGS-84 normal gravity (free_integration.py:59-61).
            earth_rot: consider the Earth rotation (only used when ref_frame == 0).
            lanes_per_run: lane-group width of the CUDA kernel (0 = automatic).
        '''
        self.input = ['ref_frame', 'fs', 'gyro', 'accel']
        self.output = ['att_euler', 'pos', 'vel']
        self.earth_rot = earth_rot
        self.batch = True
        self.results = None
        self.ref_frame = 1
        self.dt = 1.0
        self.att = None
        self.pos = None
        self.vel = None
        self.lanes_per_run = lanes_per_run
        self.ini_sets = engine.ini_sets_from_plugin(ini_pos_vel_att)   # [S][rows]
        self.set_of_inis = self.ini_sets.shape[0]
        self.run_times = int(0)   # as in the reference: never reset; run k (1-based) uses
                                  # initial-state set k-1 while k <= set_of_inis, else set 0
        self._ini_dev = None

    # ---- reference plugin protocol ---------------------------------------
    def run(self, set_of_input):
        '''
        One simulation run.  set_of_input = [ref_frame, fs, gyro (n,3) rad/s, accel (n,3) m/s^2].
        '''
        if set_of_input[0] == 0:       # sticky, like free_integration.py:71-72
            self.ref_frame = 0
        gyro = np.ascontiguousarray(set_of_input[2], dtype=np.float64)
        accel = np.ascontiguousarray(set_of_input[3], dtype=np.float64)
        att, pos, vel = self.run_batch(self.ref_frame, set_of_input[1], gyro[None], accel[None])
        self.att, self.pos, self.vel = att[0], pos[0], vel[0]
        self.results = [self.att, self.pos, self.vel]

    def get_results(self):
        '''
        return algorithm results as specified in self.output
        '''
        return self.results

    def reset(self):
        '''
        Nothing to reset (the reference keeps run_times across runs too).
        '''
        pass

    # ---- batched entry -------------------------------------------------------
    def ini_device(self):
        if self._ini_dev is None:
            self._ini_dev = engine.to_device(self.ini_sets)
        return self._ini_dev

    def run_batch(self, ref_frame, fs, gyro, accel, to_host=True):
        '''
        R runs in one launch.  gyro, accel: [R, n, 3] (numpy or CUDA tensors).
        Returns att, pos, vel [R, n, 3] (numpy if to_host else CUDA tensors).
        Run r of the batch is simulation run run_times + r for the initial-state rule.
        '''
        ref_frame = 0 if ref_frame == 0 else 1
        self.dt = 1.0 / fs
        g = engine.to_device(gyro)
        a = engine.to_device(accel)
        if g.dim() != 3 or g.shape[2] != 3 or g.shape != a.shape:
            raise ValueError('gyro and accel must both be [R, n, 3]')
        att, pos, vel = engine.free_integration(ref_frame, fs, g, a, self.ini_device(),
                                                earth_rot=self.earth_rot,
                                                run_offset=self.run_times,
                                                lanes_per_run=self.lanes_per_run)
        self.run_times += g.shape[0]
        if to_host:
            return att.cpu().numpy(), pos.cpu().numpy(), vel.cpu().numpy()
        return att, pos, vel
