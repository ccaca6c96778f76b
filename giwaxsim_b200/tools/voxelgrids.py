"""Drop-in for the hot functions of the reference's tools/voxelgrids.py:
same names, argument order and side effects, work done by the sm_100a kernels.

    rotate_project_fft_coords(args16)   voxelgrids.py:311-416
    process_file2(...)                  voxelgrids.py:464-506
    downselect_voxelgrid(...)           voxelgrids.py:16-48
    add_f0_q_3d(...)                    voxelgrids.py:828-857
    generate_voxel_grid_low_mem(...)    voxelgrids.py:535-722  (file -> whole I(q) grid, two-step CLI)

The two accumulator arguments are names of device shared arrays created with
tools.utilities.create_shared_array (they replace the POSIX shm names).
"""
import numpy as np
import torch

from .. import engine
from .._lib import call, ptr
from .utilities import (ATOMIC_NUMBER, CROMER_MANN, get_element_f0_dict, get_element_f1_f2_dict, load_structure,
                        lookup_shared_array)

_engine_cache = {}
_ENGINE_CACHE_MAX = 2


def _slice_engine_for(coords, f_values, grid_size, r_voxel_size, avg_voxel_f, x_bound, y_bound,
                      fill_bkg, smooth, q_axis):
    """SliceEngine for this slab, cached on the identity of the input arrays so
    that the reference's one-call-per-phi loop uploads and sorts the atoms once."""
    key = (id(coords), id(f_values), coords.shape, int(grid_size), float(r_voxel_size), complex(avg_voxel_f),
           float(x_bound), float(y_bound), bool(fill_bkg), int(smooth or 0), id(q_axis), len(q_axis))
    hit = _engine_cache.get(key)
    if hit is not None:
        return hit[0]
    codes, uniq = engine.encode_values(np.asarray(f_values, dtype=complex))
    kw = dict(species=codes, table=uniq) if codes is not None else dict(f_values=f_values)
    eng = engine.SliceEngine(coords, r_voxel_size, q_axis, grid_size, avg_voxel_f, x_bound, y_bound,
                             fill_bkg, smooth, count3d=True,
                             accumulators=(None, None, None), **kw)
    while len(_engine_cache) >= _ENGINE_CACHE_MAX:
        _engine_cache.pop(next(iter(_engine_cache)))
    _engine_cache[key] = (eng, coords, f_values, q_axis)   # keep the arrays alive: ids stay valid
    return eng


def rotate_project_fft_coords(args):
    """One phi slice: rotate, project, FFT, bin into the two named accumulators."""
    (coords, f_values, phi, grid_size, r_voxel_size, avg_voxel_f, x_bound, y_bound, z_bound,
     fill_bkg, smooth, qx, qy, qz, voxel_grid_shm_name, voxel_grid_count_shm_name) = args
    eng = _slice_engine_for(coords, f_values, grid_size, r_voxel_size, avg_voxel_f, x_bound, y_bound,
                            fill_bkg, smooth, qx)
    q3 = eng.q_num ** 3
    vsum = lookup_shared_array(voxel_grid_shm_name)
    vcnt = lookup_shared_array(voxel_grid_count_shm_name)
    if int(np.prod(vsum.shape)) != q3 or int(np.prod(vcnt.shape)) != q3:
        raise ValueError("accumulator shape does not match len(qx)^3")
    eng.vsum = vsum.device_tensor(torch.float32, eng.device)
    eng.count3 = vcnt.device_tensor(torch.int32, eng.device)
    eng.count2 = None
    eng.run(np.array([phi], dtype=np.float64))


def process_file2(iq_2d, det_h_qx, det_h_qy, det_v_qz, qx, qy, qz, voxel_grid_shm_name,
                  voxel_grid_count_shm_name):
    """Bin one slice image into the [qy,qx,qz] accumulators."""
    dev = engine.resolve_device()
    q_num = len(qx)
    vsum = lookup_shared_array(voxel_grid_shm_name).device_tensor(torch.float32, dev)
    vcnt = lookup_shared_array(voxel_grid_count_shm_name).device_tensor(torch.int32, dev)
    iq_2d = np.asarray(iq_2d)
    rows, cols = iq_2d.shape
    qmin_x, qmax_x = float(np.min(qx)), float(np.max(qx))
    if (float(np.min(qy)), float(np.max(qy)), float(np.min(qz)), float(np.max(qz))) != (qmin_x, qmax_x) * 2:
        raise ValueError("qx, qy, qz must share one axis (the reference's voxels are cubes)")
    dq = float(np.diff(qz)[0])
    with torch.cuda.device(dev):
        st = engine._stream()
        d_iq = engine._dev(iq_2d, dev, torch.float32).contiguous()
        col = torch.empty(cols, dtype=torch.int32, device=dev)
        row = torch.empty(rows, dtype=torch.int32, device=dev)
        d_hx, d_hy, d_vz = (engine._dev(np.asarray(a, dtype=np.float64), dev)
                            for a in (det_h_qx, det_h_qy, det_v_qz))
        call("gx_axis_col_index", ptr(d_hx), ptr(d_hy), cols, qmin_x, qmax_x, dq, q_num, ptr(col), st)
        call("gx_axis_row_index", ptr(d_vz), rows, qmin_x, qmax_x, dq, q_num, ptr(row), st)
        call("gx_bin_slices", ptr(d_iq), 1, rows, cols, ptr(col), cols, ptr(row), q_num,
             ptr(vsum), ptr(vcnt), None, st)
        torch.cuda.current_stream().synchronize()


def downselect_voxelgrid(grid, x_axis, y_axis, z_axis, max_val):
    """Cube |q| < max_val + dq around the origin (pure slicing, no arithmetic)."""
    x0, x1 = engine.crop_range(np.asarray(x_axis), max_val)
    lim = max_val + np.abs(x_axis[1] - x_axis[0])
    yi = np.where(np.abs(y_axis) < lim)[0]
    zi = np.where(np.abs(z_axis) < lim)[0]
    y0, y1, z0, z1 = yi[0], yi[-1] + 1, zi[0], zi[-1] + 1
    return grid[y0:y1, x0:x1, z0:z1], x_axis[x0:x1], y_axis[y0:y1], z_axis[z0:z1]


def add_f0_q_3d(iq, qx_axis, qy_axis, qz_axis, element):
    """iq * (f0(|q|)/Z)^2 evaluated on the device (fp32 result returned as fp64)."""
    qx_axis, qy_axis, qz_axis = (np.asarray(a, dtype=np.float64) for a in (qx_axis, qy_axis, qz_axis))
    if not (np.array_equal(qx_axis, qy_axis) and np.array_equal(qx_axis, qz_axis)):
        raise ValueError("the device finaliser assumes one shared cubic axis")
    dev = engine.resolve_device()
    V = len(qx_axis)
    with torch.cuda.device(dev):
        d_sum = engine._dev(np.asarray(iq), dev, torch.float32).contiguous()
        ones = torch.ones(V ** 3, dtype=torch.int32, device=dev)
        out = torch.empty(V ** 3, dtype=torch.float32, device=dev)
        aff = np.asarray(CROMER_MANN[element], dtype=np.float64)
        d_axis = engine._dev(qx_axis, dev)
        call("gx_voxel_finalize", ptr(d_sum), ptr(ones), None, None, V, 0, V, ptr(d_axis),
             ptr(aff), float(ATOMIC_NUMBER[element]), 0, -1, ptr(out), engine._stream())
        return out.cpu().to(torch.float64).numpy().reshape(V, V, V)


def generate_voxel_grid_low_mem(input_path, r_voxel_size, q_voxel_size, max_q, aff_num_qs, energy, gen_name,
                                output_dir=None, scratch_folder=None, num_cpus=None, fill_bkg=False, smooth=0):
    """Whole (uncropped, unweighted) 3-D I(q) grid of a structure file by the projection-slice
    method (voxelgrids.py:535-722).  Returns (iq[qy,qx,qz], qx, qy, qz) as float64 arrays, or
    writes `<output_dir>/<gen_name>_output_files/<gen_name>_{iq,qx,qy,qz}.npy` and returns None.
    `scratch_folder` and `num_cpus` are accepted and unused (nothing is staged on disk here).

    aff_num_qs == 1: f = Z + f' + i f'' per atom (:603-609).
    aff_num_qs > 1: the reference loops over |q| shells with f = f0(q_shell) + f' + i f'', but
    every pass REPLACES the grid by that shell's whole sum/count (:699) before adding the masked
    copy (:706-707), so what it returns is the last shell's grid, doubled inside that shell.
    Earlier passes cannot influence the result and are not run; the returned array is the same.
    """
    from .. import parallel
    from .comparison import f_table, voxelgrid_device
    import os

    coords, elements = load_structure(input_path)
    if aff_num_qs == 1:
        shell = None

        def table_of(uniq):
            return f_table(uniq, energy)
    elif aff_num_qs > 1:
        max_q_diag = np.sqrt(2) * max_q
        max_q_diag = max_q_diag + max_q_diag % q_voxel_size          # voxelgrids.py:590
        step = (max_q_diag / (int(aff_num_qs)))
        q_val = 0.5 * step + (int(aff_num_qs) - 1) * step            # last pass of the loop :660-662
        shell = (q_val - step / 2, q_val + step / 2)

        def table_of(uniq):
            names = [str(e) for e in uniq]
            f0, f12 = get_element_f0_dict(q_val, names), get_element_f1_f2_dict(energy, names)
            return [complex(f0[e] + f12[e]) for e in names]
    else:
        raise Exception('Invalid aff_num_qs value. Must be non-negative integer')
    iq_dev, axis, _, world = voxelgrid_device(coords, elements, table_of, r_voxel_size, q_voxel_size, max_q,
                                              fill_bkg, smooth, crop=False, f0=False)
    if shell is not None:
        engine.scale_shell(iq_dev, axis, shell[0], shell[1], 2.0, iq_dev.device)
    with torch.cuda.device(iq_dev.device):
        iq = engine.to_host_f64(iq_dev, replicated=world > 1)
    if output_dir:
        save_path = f'{output_dir}/{gen_name}_output_files/'
        if parallel.rank_world()[0] == 0:
            if not os.path.exists(save_path):
                os.mkdir(save_path)
            np.save(f'{save_path}{gen_name}_iq.npy', iq)
            for name in ('qx', 'qy', 'qz'):
                np.save(f'{save_path}{gen_name}_{name}.npy', axis)
        return None
    return iq, axis.copy(), axis.copy(), axis.copy()
