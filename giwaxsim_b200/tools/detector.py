"""Drop-in for the reference's tools/detector.py (same names and signatures).

    make_detector                 detector.py:5-29
    rotate_about_normal/..        detector.py:33-162
    rotate_psi_phi_theta          detector.py:234-244
    intersect_detector            detector.py:194-232
    generate_detector_ints        detector.py:278-300
    mirror_vertical_horizontal    detector.py:246-275

Host NumPy is used only for O(1) scalars (axes, 3x3 matrices); every per-pixel
operation (rotation, floor-binning, gather, mirror) runs in the kernels.
"""
import numpy as np
import torch

from .. import engine
from .utilities import lookup_shared_array, rotation_matrix


def make_detector(h_max, num_pixels_h, v_max, num_pixels_v):
    """Detector plane at x = 0; rows are the vertical (z) axis, columns the
    horizontal (y) axis.  Returns (x, y, z grids, h axis, v axis)."""
    h_axis_vals = np.linspace(-h_max, h_max, num_pixels_h)
    v_axis_vals = np.linspace(-v_max, v_max, num_pixels_v)
    det_y_grid = np.broadcast_to(h_axis_vals[None, :], (num_pixels_v, num_pixels_h)).copy()
    det_z_grid = np.broadcast_to(v_axis_vals[:, None], (num_pixels_v, num_pixels_h)).copy()
    return np.zeros_like(det_y_grid), det_y_grid, det_z_grid, h_axis_vals, v_axis_vals


def _axis(det_x_grid, det_y_grid, det_z_grid, which):
    p = engine.grid_corners(det_x_grid, det_y_grid, det_z_grid)
    across, down = p[1] - p[0], p[2] - p[0]
    if which == "normal":
        u = np.cross(across, down)
    elif which == "vertical":
        u = down
    else:
        u = across
    u = u / np.linalg.norm(u)
    return u


def _rotate(det_x_grid, det_y_grid, det_z_grid, which, angle_deg):
    R = rotation_matrix(_axis(det_x_grid, det_y_grid, det_z_grid, which), np.radians(angle_deg))
    return engine.rotate_points(R, det_x_grid, det_y_grid, det_z_grid)


def rotate_about_normal(det_x_grid, det_y_grid, det_z_grid, psi):
    return _rotate(det_x_grid, det_y_grid, det_z_grid, "normal", psi)


def rotate_about_vertical(det_x_grid, det_y_grid, det_z_grid, phi):
    return _rotate(det_x_grid, det_y_grid, det_z_grid, "vertical", phi)


def rotate_about_horizontal(det_x_grid, det_y_grid, det_z_grid, theta):
    return _rotate(det_x_grid, det_y_grid, det_z_grid, "horizontal", theta)


def rotate_psi_phi_theta(det_x, det_y, det_z, psi, phi, theta):
    g = rotate_about_normal(det_x, det_y, det_z, psi)
    g = rotate_about_vertical(*g, phi)
    return rotate_about_horizontal(*g, theta)


_IDENTITY_STEP = np.tile(np.eye(3).reshape(1, 9), (1, 3, 1))


def intersect_detector(int_voxels, qx, qy, qz, det_x_grid, det_y_grid, det_z_grid):
    """Intensity of every detector pixel: floor-bin to the voxel grid, clamp to
    its bounds, gather."""
    det = engine.DetectorEngine(int_voxels, qx, qy, qz)
    image, _ = det.accumulate(det_x_grid, det_y_grid, det_z_grid, _IDENTITY_STEP, np.ones(1))
    return image.cpu().numpy().reshape(np.shape(det_x_grid))


_det_cache = {}


def _detector_engine_for(iq, qx, qy, qz):
    key = (id(iq), np.shape(iq), id(qz))
    hit = _det_cache.get(key)
    if hit is None:
        _det_cache.clear()
        hit = (engine.DetectorEngine(iq, qx, qy, qz), iq, qz)
        _det_cache[key] = hit
    return hit[0]


def generate_detector_ints(args):
    """One (psi, phi, theta) orientation added, weighted, to the named image."""
    (iq, qx, qy, qz, det_x, det_y, det_z, psi, psi_weight, phi, phi_weight, theta, theta_weight,
     det_ints_shm_name) = args
    det = _detector_engine_for(iq, qx, qy, qz)
    R, w = engine.orientation_tables(engine.grid_corners(det_x, det_y, det_z), [psi], [psi_weight],
                                     [phi], [phi_weight], [theta], [theta_weight])
    image = lookup_shared_array(det_ints_shm_name).device_tensor(torch.float64, det.device)
    det.accumulate(det_x, det_y, det_z, R, w, image=image)


def mirror_vertical_horizontal(qmap):
    """qmap + its three mirror images, with the odd-size centre rules."""
    qmap = np.asarray(qmap, dtype=np.float64)
    dev = engine.resolve_device()
    rows, cols = qmap.shape
    with torch.cuda.device(dev):
        out = engine.detector_epilogue(engine._dev(qmap, dev), rows, cols, True, dev, finish=False)
        return out.cpu().numpy()
