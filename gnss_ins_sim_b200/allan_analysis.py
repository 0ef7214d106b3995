"""Allan plugin -- device-backed mirror of demo_algorithms/allan_analysis.py:15-61
(input ['fs','accel','gyro'], output ['algo_time','ad_accel','ad_gyro']); the variance
itself is csrc/allan_kernel.cuh (K4, allan.allan_var allan.py:18-59)."""
import numpy as np
import torch

from . import engine


class Allan(object):
    '''
    Allan deviation of the three accelerometer and three gyroscope channels.
    '''

    def __init__(self):
        self.input = ['fs', 'accel', 'gyro']
        self.output = ['algo_time', 'ad_accel', 'ad_gyro']
        self.batch = True
        self.results = None

    def run(self, set_of_input):
        '''
        set_of_input = [fs, accel (n,3), gyro (n,3)]
        '''
        fs = set_of_input[0]
        tau, ad_a, ad_g = self.run_batch(fs, np.asarray(set_of_input[1])[None],
                                         np.asarray(set_of_input[2])[None])
        self.results = [tau, ad_a[0], ad_g[0]]

    def run_batch(self, fs, accel, gyro, to_host=True, channel_major=False):
        '''
        accel, gyro: [R, n, 3] (the reference's per-run arrays) or, channel_major, [R, 3, n].
        Returns tau [ntau], ad_accel [R, ntau, 3], ad_gyro [R, ntau, 3]
        (Allan DEVIATION = sqrt(avar), allan_analysis.py:47-49).
        '''
        a = engine.to_device(accel)
        g = engine.to_device(gyro)
        out = []
        for x in (a, g):
            if channel_major:     # 3R contiguous series: the bulk-copy front end of K4
                R, _, n = x.shape
                avar, tau = engine.allan(fs, x, n, R * 3)
            else:                 # 3R interleaved series, read in place (no copy)
                R, n, _ = x.shape
                avar, tau = engine.allan(fs, x, n, R * 3, inner=3, outer_stride=3 * n, sample_stride=3)
            out.append(torch.sqrt(avar).reshape(R, 3, -1).permute(0, 2, 1).contiguous())
        if to_host:
            return tau.cpu().numpy(), out[0].cpu().numpy(), out[1].cpu().numpy()
        return tau, out[0], out[1]

    def get_results(self):
        return self.results

    def reset(self):
        pass
