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


# key -> (converter, default) of the reference's voxelgridmaker config (old_modules/voxelgridmaker.py:12-25);
# a default that is a callable is evaluated when the key is absent
SCHEMA = {
    "input_folder": (str, None), "input_filepath": (str, None), "filetype": (str, "xyz"), "gen_name": (str, None),
    "r_voxel_size": (float, 0.3), "q_voxel_size": (float, 0.01), "aff_num_qs": (int, 1), "energy": (float, 1),
    "max_q": (float, 2.5), "output_dir": (str, os.getcwd), "num_cpus": (int, os.cpu_count),
    "scratch_folder": (str, os.getcwd), "smooth": (int, 0), "fill_bkg": (str_to_bool, "False"),
}
GRID_FILES = ("iq", "qx", "qy", "qz")          # <gen_name>_<part>.npy: the hand-off both detectormakers read


def read_settings(config):
    """Typed settings from the key=value dictionary, reference defaults for missing keys."""
    out = {}
    for key, (convert, default) in SCHEMA.items():
        raw = config.get(key, default() if callable(default) else default)
        out[key] = raw if raw is None else convert(raw)
    return out


def structure_files(settings):
    if settings["input_folder"]:
        return glob.glob(f'{settings["input_folder"]}/*{settings["filetype"]}')
    if settings["input_filepath"]:
        return [settings["input_filepath"]]
    raise Exception('Either input_folder or input_path must be specified')


def averaged_cropped_grid(paths, s):
    """Mean over the structure files of their cropped grids (:33-59)."""
    total = axes = None
    for path in paths:
        grid = generate_voxel_grid_low_mem(path, s["r_voxel_size"], s["q_voxel_size"], s["max_q"], s["aff_num_qs"],
                                           s["energy"], s["gen_name"], scratch_folder=s["scratch_folder"],
                                           num_cpus=s["num_cpus"], fill_bkg=s["fill_bkg"], smooth=s["smooth"])
        cropped, *axes = downselect_voxelgrid(*grid, s["max_q"])
        total = np.ascontiguousarray(cropped) if total is None else total + cropped
    return total / len(paths), axes


def main(config):
    s = read_settings(config)
    paths = structure_files(s)
    iq, (qx, qy, qz) = averaged_cropped_grid(paths, s)
    if s["aff_num_qs"] == 1:
        iq = add_f0_q_3d(iq, qx, qy, qz, most_common_element(paths[0]))          # (:66-68)
    if parallel.rank_world()[0] == 0:
        folder = f'{s["output_dir"]}/{s["gen_name"]}_output_files'
        os.makedirs(folder, exist_ok=True)
        for part, array in zip(GRID_FILES, (iq, qx, qy, qz)):
            np.save(f'{folder}/{s["gen_name"]}_{part}.npy', array)
        save_config_to_txt(config, f'{folder}/{s["gen_name"]}_config.txt')
    return iq, qx, qy, qz


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description="Structure file(s) -> 3-D I(q) voxel grid (.npy)")
    ap.add_argument('--config', type=str, required=True, help='Path to the configuration file')
    args = ap.parse_args()
    parallel.init_from_env()
    start = time.time()
    main(parse_config_file(args.config))
    print(f'\nTotal Time: {str(time.time() - start)}')
