"""`python -m giwaxsim_b200.detectormaker --config file.txt`: second half of the
reference's two-step command line (old_modules/detectormaker.py:15-171) on the
B200 path.

Reads `<iq_output_folder>/<gen_name>_{iq,qx,qy,qz}.npy` (written by
`giwaxsim_b200.voxelgridmaker` or the reference's own script), sums the
detector image over the psi x phi x theta orientation grid and writes
`<iq_output_folder>/<gen_name>_det_sum[<i>]/<gen_name>_{det_h,det_v,det_sum}.npy`
and the echoed configuration.  Unlike detectormaker_fitting the two-step
script floors the image at 1e-6 but does NOT rescale it by 1e-6 (:152-154),
and its `mirror` default is False (:39).  The two PNG previews are not produced.
"""
import argparse
import os
import time

import numpy as np
import torch

from . import engine, parallel
from .tools.comparison import detector_base_device
from .tools.utilities import parse_config_file, save_config_to_txt, str_to_bool


def main(config):
    iq_output_folder = config.get('iq_output_folder')
    gen_name = config.get('gen_name')
    max_q = float(config.get('max_q', 2.5))
    num_pixels = int(config.get('num_pixels', 500))
    angle_init_vals = tuple(float(config.get('angle_init_val%d' % k, 0)) for k in (1, 2, 3))
    angle_init_axs = tuple(config.get('angle_init_ax%d' % k, 'None') for k in (1, 2, 3))
    angles, weights = {}, {}
    for name in ('psi', 'phi', 'theta'):
        num = int(config.get(name + '_num'))
        angles[name] = np.linspace(float(config.get(name + '_start')), float(config.get(name + '_end')), num=num)
        path = config.get(name + '_weights_path', None)
        weights[name] = np.load(path) if path else np.ones_like(angles[name]) / num
    mirror = str_to_bool(config.get('mirror', 'False'))

    save_path = iq_output_folder
    if not os.path.exists(save_path):
        raise Exception(f'Path does not exist: {save_path}')
    iq = np.load(f'{save_path}/{gen_name}_iq.npy')
    qx = np.load(f'{save_path}/{gen_name}_qx.npy')
    qy = np.load(f'{save_path}/{gen_name}_qy.npy')
    qz = np.load(f'{save_path}/{gen_name}_qz.npy')

    rank, world = parallel.rank_world()
    det_sum_path = f'{save_path}/{gen_name}_det_sum'
    if rank == 0:
        i = 0
        while os.path.exists(det_sum_path):
            i += 1
            det_sum_path = f'{save_path}/{gen_name}_det_sum{i}'
        os.mkdir(det_sum_path)

    for name in ('psi', 'phi', 'theta'):
        assert len(angles[name]) == len(weights[name]), f'{name} weights length must equal {name}_num'
    for name in ('psi', 'phi', 'theta'):
        assert np.abs(1 - np.sum(weights[name])) < 0.01, f'{name} weights must sum to 1'

    dev = engine.resolve_device()
    gx, gy, gz, det_h, det_v = detector_base_device(num_pixels, max_q, angle_init_vals, angle_init_axs, dev)
    if rank == 0:
        np.save(f'{det_sum_path}/{gen_name}_det_h.npy', det_h)
        np.save(f'{det_sum_path}/{gen_name}_det_v.npy', det_v)
    det = engine.DetectorEngine(iq, qx, qy, qz, device=dev)
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), angles['psi'], weights['psi'],
                                     angles['phi'], weights['phi'], angles['theta'], weights['theta'])
    sel = parallel.shard(np.arange(len(w)), rank, world)
    with torch.cuda.device(dev):
        image = torch.zeros(num_pixels * num_pixels, dtype=torch.float64, device=dev)
    if len(sel):
        det.accumulate(gx, gy, gz, np.ascontiguousarray(R[sel]), np.ascontiguousarray(w[sel]), image=image)
    if world > 1:
        parallel.all_reduce_sum([image])
    out = engine.detector_epilogue(image, num_pixels, num_pixels, mirror, dev, finish=2)   # floor, no rescale
    with torch.cuda.device(dev):
        det_sum = engine.to_host_f64(out, replicated=world > 1)
    if rank == 0:
        np.save(f'{det_sum_path}/{gen_name}_det_sum.npy', det_sum)
        save_config_to_txt(config, f'{det_sum_path}/{gen_name}_config.txt')
    return det_sum, det_h.copy(), det_v.copy()


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="3-D I(q) voxel grid (.npy) -> 2-D detector image")
    ap.add_argument('--config', type=str, required=True, help='Path to the configuration file')
    args = ap.parse_args()
    parallel.init_from_env()
    start = time.time()
    main(parse_config_file(args.config))
    print(f'\nTotal Time: {str(time.time() - start)}')
