"""`python -m giwaxsim_b200.voxelgridmaker --config file.txt`: first half of the
reference's two-step command line (old_modules/voxelgridmaker.py:10-82) on the
B200 path.

Structure file(s) -> `<output_dir>/<gen_name>_output_files/<gen_name>_{iq,qx,qy,qz}.npy`
(float64, iq indexed [qy,qx,qz], cropped to |q| < max_q + dq) plus the echoed
configuration - the hand-off format `giwaxsim_b200.detectormaker` and the
reference's old_modules/detectormaker.py read.  With `input_folder` every file
matching `*<filetype>` is simulated and the cropped grids are averaged
(iq_sum / len, :33-59); the f0 weight uses the most common element of the
first file (:66-68).
"""
import argparse
import glob
import os
import time

import numpy as np

from . import parallel
from .tools.utilities import most_common_element, parse_config_file, save_config_to_txt, str_to_bool
from .tools.voxelgrids import add_f0_q_3d, downselect_voxelgrid, generate_voxel_grid_low_mem


def main(config):
    input_folder = config.get('input_folder', None)
    input_filepath = config.get('input_filepath', None)
    filetype = config.get('filetype', 'xyz')
    gen_name = config.get('gen_name')
    r_voxel_size = float(config.get('r_voxel_size', 0.3))
    q_voxel_size = float(config.get('q_voxel_size', 0.01))
    aff_num_qs = int(config.get('aff_num_qs', 1))
    energy = float(config.get('energy', 1))
    max_q = float(config.get('max_q', 2.5))
    output_dir = config.get('output_dir', os.getcwd())
    num_cpus = int(config.get('num_cpus', os.cpu_count()))
    scratch_folder = config.get('scratch_folder', os.getcwd())
    smooth = int(config.get('smooth', 0))
    fill_bkg = str_to_bool(config.get('fill_bkg', 'False'))

    if input_folder:
        input_paths = glob.glob(f'{input_folder}/*{filetype}')
    elif input_filepath:
        input_paths = [input_filepath]
    else:
        raise Exception('Either input_folder or input_path must be specified')

    iq_sum = None
    for input_path in input_paths:
        iq, qx, qy, qz = generate_voxel_grid_low_mem(input_path, r_voxel_size, q_voxel_size, max_q, aff_num_qs,
                                                     energy, gen_name, scratch_folder=scratch_folder,
                                                     num_cpus=num_cpus, fill_bkg=fill_bkg, smooth=smooth)
        iq_small, qx, qy, qz = downselect_voxelgrid(iq, qx, qy, qz, max_q)
        del iq
        if iq_sum is None:
            iq_sum = np.ascontiguousarray(iq_small)
        else:
            iq_sum += iq_small
    iq = iq_sum
    iq /= len(input_paths)
    if aff_num_qs == 1:
        iq = add_f0_q_3d(iq, qx, qy, qz, most_common_element(input_paths[0]))

    save_path = f'{output_dir}/{gen_name}_output_files'
    if parallel.rank_world()[0] == 0:
        if not os.path.exists(save_path):
            os.mkdir(save_path)
        np.save(f'{save_path}/{gen_name}_iq.npy', iq)
        np.save(f'{save_path}/{gen_name}_qx.npy', qx)
        np.save(f'{save_path}/{gen_name}_qy.npy', qy)
        np.save(f'{save_path}/{gen_name}_qz.npy', qz)
        save_config_to_txt(config, f'{save_path}/{gen_name}_config.txt')
    return iq, qx, qy, qz


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Structure file(s) -> 3-D I(q) voxel grid (.npy)")
    ap.add_argument('--config', type=str, required=True, help='Path to the configuration file')
    args = ap.parse_args()
    parallel.init_from_env()
    start = time.time()
    main(parse_config_file(args.config))
    print(f'\nTotal Time: {str(time.time() - start)}')
