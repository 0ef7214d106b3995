"""Odometer-aided free integration plugin -- device-backed mirror of
demo_algorithms/free_integration_odo.py:15-172 (input ['ref_frame','fs','gyro','odo'], output
['att_euler','pos','vel']): the attitude recurrence of FreeIntegration with the body velocity
taken from the odometer, [odo, 0, 0] (free_integration_odo.py:106-108, :142-144).  This is
`algo1` of the reference's demo_free_integration.py (:61-71)."""
import numpy as np

from . import engine
from .free_integration import FreeIntegration as _Base


class FreeIntegration(_Base):
    '''
    Integrate gyro to get attitude, take the forward speed from the odometer.
    '''

    def __init__(self, ini_pos_vel_att, earth_rot=True, lanes_per_run=0):
        super().__init__(ini_pos_vel_att, earth_rot, lanes_per_run)
        self.input = ['ref_frame', 'fs', 'gyro', 'odo']

    def run(self, set_of_input):
        if set_of_input[0] == 0:
            self.ref_frame = 0
        gyro = np.ascontiguousarray(set_of_input[2], dtype=np.float64)
        odo = np.ascontiguousarray(set_of_input[3], dtype=np.float64).reshape(-1)
        att, pos, vel = self.run_batch(self.ref_frame, set_of_input[1], gyro[None], odo[None])
        self.att, self.pos, self.vel = att[0], pos[0], vel[0]
        self.results = [self.att, self.pos, self.vel]

    def run_batch(self, ref_frame, fs, gyro, odo, to_host=True):
        '''
        gyro [R, n, 3], odo [R, n] -> att, pos, vel [R, n, 3].
        '''
        ref_frame = 0 if ref_frame == 0 else 1
        self.dt = 1.0 / fs
        g = engine.to_device(gyro)
        o = engine.to_device(odo)
        if g.dim() != 3 or g.shape[2] != 3 or tuple(o.shape) != tuple(g.shape[:2]):
            raise ValueError('gyro must be [R, n, 3] and odo [R, n]')
        att, pos, vel = engine.free_integration_odo(ref_frame, fs, g, o, self.ini_device(),
                                                    earth_rot=self.earth_rot,
                                                    run_offset=self.run_times,
                                                    lanes_per_run=self.lanes_per_run)
        self.run_times += g.shape[0]
        if to_host:
            return att.cpu().numpy(), pos.cpu().numpy(), vel.cpu().numpy()
        return att, pos, vel
