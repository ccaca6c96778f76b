"""`python -m giwaxsim_b200.simulate --config file.txt`: the reference's
simulate_GIWAXS.py driver (simulate_GIWAXS.py:18-233) on the B200 path.

Same key=value configuration keys and defaults (simulate_GIWAXS.py:20-82), same
outputs per input structure: `<save_folder>/<name>/det_h.npy, det_v.npy,
det_sum.npy` (the h >= 0, v >= 0 quadrant when mirror is on, :172-180) and the
configuration echoed to `<save_folder>/config.txt` (:217).  Plotting and the
experimental-comparison branch (img_path; fabio / matplotlib) are not part of
the hot path: with `img_path` set, run the reference's own script with the
three imports swapped as INTEGRATION.md shows.

Under torchrun (one process per GPU) call it the same way: every rank computes
its share of the rotations / orientations, rank 0 writes the files.
"""
import argparse
import glob
import os
import time

import numpy as np

from . import parallel
from .tools.comparison import detectormaker_fitting, slabmaker_fitting, voxelgridmaker_fitting
from .tools.utilities import load_pdb_cell_params, parse_config_file, save_config_to_txt, str_to_bool


def read_settings(config):
    """Typed settings from the string dictionary, with the reference's defaults."""
    g = config.get
    s = dict(
        input_folder=g('input_folder', None), input_path=g('input_filepath', None), filetype=g('filetype', None),
        x_size=float(g('x_size', 0)), y_size=float(g('y_size', 0)), z_size=float(g('z_size', 0)),
        a=float(g('a', 0)), b=float(g('b', 0)), c=float(g('c', 0)),
        alpha=float(g('alpha', 0)), beta=float(g('beta', 0)), gamma=float(g('gamma', 0)),
        r_voxel_size=float(g('r_voxel_size', 0.3)), q_voxel_size=float(g('q_voxel_size', 0.04)),
        max_q=float(g('max_q', 2.5)), energy=float(g('energy', 10000)),
        fill_bkg=str_to_bool(g('fill_bkg', 'False')), smooth=int(g('smooth', 0)),
        psi_weights_path=g('psi_weights_path', None), phi_weights_path=g('phi_weights_path', None),
        theta_weights_path=g('theta_weights_path', None), mirror=str_to_bool(g('mirror', 'False')),
        img_path=g('img_path', None), save_folder=g('save_folder', os.getcwd()))
    s['num_pixels'] = int(g('num_pixels', s['max_q'] / s['q_voxel_size']))
    s['angle_init_vals'] = tuple(float(g('angle_init_val%d' % k, 0)) for k in (1, 2, 3))
    s['angle_init_axs'] = tuple(g('angle_init_ax%d' % k, 'None') for k in (1, 2, 3))
    for name in ('psi', 'phi', 'theta'):
        s[name + 's'] = np.linspace(float(g(name + '_start')), float(g(name + '_end')), int(g(name + '_num')))
    return s


def main(config):
    s = read_settings(config)
    if s['img_path']:
        raise NotImplementedError("experimental comparison (img_path) is outside the hot path: run the "
                                  "reference's simulate_GIWAXS.py with the imports swapped (INTEGRATION.md)")
    if s['input_folder']:
        if not s['filetype']:
            raise Exception('filetype must be specified')
        input_paths = glob.glob(f"{s['input_folder']}/*{s['filetype']}")
    elif s['input_path']:
        input_paths = [s['input_path']]
    else:
        raise Exception('Either input_folder or input_path must be specified')
    rank, _ = parallel.rank_world()
    if rank == 0:
        os.makedirs(s['save_folder'], exist_ok=True)
    x_size, y_size, z_size = s['x_size'], s['y_size'], s['z_size']
    results = {}
    for path in input_paths:
        if path.lower().endswith('.xyz'):
            cell = (s['a'], s['b'], s['c'], s['alpha'], s['beta'], s['gamma'])
        elif path.lower().endswith('.pdb'):
            cell = load_pdb_cell_params(path)
        else:
            raise Exception('Files must be a .pdb or .xyz file')
        if x_size == 0 or y_size == 0 or z_size == 0:
            x_size, y_size, z_size = s['a'], s['b'], s['c']
            print('at least one slab size was not defined. Defaulting x,y,z slab dimensions to a, b, c')
        if any(val == 0 for val in cell):
            raise Exception('at least one unit cell parameter a, b, c, alpha, beta, gamma not defined')
        coords, elements = slabmaker_fitting(path, x_size, y_size, z_size, *cell)
        iq, qx, qy, qz = voxelgridmaker_fitting(coords, elements, s['r_voxel_size'], s['q_voxel_size'], s['max_q'],
                                                s['energy'], num_cpus=None, fill_bkg=s['fill_bkg'],
                                                smooth=s['smooth'])
        det_sum, det_h, det_v = detectormaker_fitting(iq, qx, qy, qz, s['num_pixels'], s['max_q'],
                                                      s['angle_init_vals'], s['angle_init_axs'], s['psis'],
                                                      s['psi_weights_path'], s['phis'], s['phi_weights_path'],
                                                      s['thetas'], s['theta_weights_path'], mirror=s['mirror'])
        if s['mirror']:
            keep_h, keep_v = np.where(det_h >= 0)[0], np.where(det_v >= 0)[0]
            det_h, det_v, det_sum = det_h[keep_h], det_v[keep_v], det_sum[np.ix_(keep_v, keep_h)]
        name = os.path.splitext(os.path.basename(path))[0]
        results[name] = (det_sum, det_h, det_v)
        if rank == 0:
            sub = f"{s['save_folder']}/{name}"
            os.makedirs(sub, exist_ok=True)
            np.save(f'{sub}/det_h.npy', det_h)
            np.save(f'{sub}/det_v.npy', det_v)
            np.save(f'{sub}/det_sum.npy', det_sum)
    if rank == 0:
        save_config_to_txt(config, f"{s['save_folder']}/config.txt")
    return results


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="GIWAXS forward simulation on the B200 path")
    ap.add_argument("--config", type=str, required=True, help="key=value configuration file")
    args = ap.parse_args()
    parallel.init_from_env()
    t0 = time.time()
    main(parse_config_file(args.config))
    print(f'\nTotal Time: {str(np.round(time.time() - t0, 1))}s')
