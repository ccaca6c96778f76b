"""ctypes binding of libgiwaxs_b200.so (the C ABI in include/giwaxs_b200.h).

There is no CPU fallback: if the shared object is missing this module raises
at import of the first symbol, and every compute call raises GxError when the
library reports an error (no device, bad size, CUDA failure).
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GIWAXS_B200_LIB: another build of the same library (kernel experiments, scripts/build_variant.py)
LIB_PATH = os.environ.get("GIWAXS_B200_LIB") or os.path.join(_HERE, "libgiwaxs_b200.so")

GX_OK = 0
GX_ERR_INVALID = -1
GX_ERR_CUDA = -2
GX_ERR_UNSUPPORTED = -3
GX_ERR_NO_DEVICE = -4
GX_MAX_SPECIES = 16
ABI_VERSION = 5          # GX_ABI_VERSION of include/giwaxs_b200.h this binding was written for


class GxError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("giwaxs_b200 error %d: %s" % (code, message))
        self.code = code


class Chord(ctypes.Structure):
    """gx_chord (include/giwaxs_b200.h)."""
    _fields_ = [(n, ctypes.c_double) for n in
                ("hor", "ver", "stop1", "stop2", "stop12", "mid", "vcos", "rise",
                 "tan_phi", "tan_theta", "cos_phi", "cos_theta")] + \
               [("mode", ctypes.c_int32), ("pad", ctypes.c_int32)]


class FusedArgs(ctypes.Structure):
    """gx_fused_args (include/giwaxs_b200.h)."""
    _fields_ = [(n, ctypes.c_void_p) for n in
                ("d_xs", "d_ys", "d_species", "d_f", "d_row_start", "d_table", "d_sin", "d_cos", "d_yrange",
                 "d_bbox", "d_dmy", "d_mz", "d_plan", "d_col", "d_colrange", "d_row_index",
                 "d_work", "d_sum", "d_count2", "d_dc")] + \
               [(n, ctypes.c_double) for n in ("r", "pedestal_re", "pedestal_im", "avg_f_re", "avg_f_im")] + \
               [(n, ctypes.c_int32) for n in ("n_species", "n_phi", "N", "KC", "q_num", "row_lo", "row_hi",
                                              "fill_bkg", "smooth_sigma", "phases", "max_row_atoms", "pad")] + \
               [(n, ctypes.c_double) for n in ("max_row_abs_re", "max_row_abs_im", "max_abs_f_re", "max_abs_f_im")] + \
               [("table_f64", ctypes.c_double * (2 * GX_MAX_SPECIES))]


class SlabArgs(ctypes.Structure):
    """gx_slab_args (include/giwaxs_b200.h)."""
    _fields_ = [("d_cell_xyz", ctypes.c_void_p), ("d_cell_species", ctypes.c_void_p), ("n_cell", ctypes.c_int64),
                ("nx", ctypes.c_int32), ("ny", ctypes.c_int32), ("nz", ctypes.c_int32), ("pad", ctypes.c_int32)] + \
               [(n, ctypes.c_double) for n in ("ax", "bx", "by", "cx", "cy", "cz")]


_p = ctypes.c_void_p
_i = ctypes.c_int
_i64 = ctypes.c_int64
_d = ctypes.c_double

# name -> (restype, argtypes); every int-returning entry is error-checked
_PROTOTYPES = {
    "gx_abi_version": (_i, []),
    "gx_last_error": (ctypes.c_char_p, []),
    "gx_device_check": (_i, [_i]),
    "gx_coords_minmax": (_i, [_p, _i64, _p, _p]),
    "gx_atoms_sort_rows": (_i, [_p, _i64, _d, _d, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gx_species_histogram": (_i, [_p, _i, _i64, _p, _p]),
    "gx_species_codes": (_i, [_p, _i, _i64, _p, _p, _p]),
    "gx_checksum64": (_i, [_p, _i64, _i, _p, _p]),
    "gx_row_abs_f_max": (_i, [_p, _p, _p, _i, _p, _i, _p, _p]),
    "gx_slice_yrange": (_i, [_p, _p, _i64, _p, _p, _i, _p, _p]),
    "gx_extreme_atoms": (_i, [_p, _p, _i64, _p, _p, _p, _i, _p, _p]),
    "gx_hull_filter": (_i, [_p, _p, _i64, _p, _i, _d, _p, _p, _p, _i, _p]),
    "gx_slice_bbox": (_i, [_p, _p, _p, _i, _d, _p, _p, _p, _i, _p, _p, _p]),
    "gx_atom_pixel_indices": (_i, [_p, _p, _p, _p, _i64, _i, _d, _d, _d, _d, _p, _p, _p]),
    "gx_slice_vectors": (_i, [_p, _p, _i, _i, _d, _d, _d, _d, _d, _d, _i, _i, _p, _i, _p, _p, _p, _p, _p]),
    "gx_project_slices": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _d,
                               _d, _d, _i, _i, _p, _p]),
    "gx_fft_plan_bytes": (_i64, [_i]),
    "gx_fft_plan_fill": (_i, [_i, _p]),
    "gx_fft2_abs2_shift": (_i, [_p, _p, _p, _i, _i, _p, _d, _d, _p]),
    "gx_slice_col_index": (_i, [_p, _p, _p, _p, _i, _i, _d, _d, _d, _i, _p, _p]),
    "gx_axis_col_index": (_i, [_p, _p, _i, _d, _d, _d, _i, _p, _p]),
    "gx_axis_row_index": (_i, [_p, _i, _d, _d, _d, _i, _p, _p]),
    "gx_bin_slices": (_i, [_p, _i, _i, _i, _p, _i, _p, _i, _p, _p, _p, _p]),
    "gx_row_histogram": (_i, [_p, _i, _i, _p, _p]),
    "gx_voxel_finalize": (_i, [_p, _p, _p, _p, _i, _i, _i, _p, _p, _d, _i64, _i64, _p, _p]),
    "gx_voxel_shell_scale": (_i, [_p, _i, _p, _d, _d, _d, _p]),
    "gx_slice_col_range": (_i, [_p, _i, _i, _p, _p]),
    "gx_window_indices": (_i, [_p, _i64, _i, _i, _i, _i, _p]),
    "gx_slices_fused": (_i, [_p, _p]),
    "gx_fused_wants_zeroed_work": (_i, [_i, _i]),
    "gx_fold_dc": (_i, [_p, _p, _p]),
    "gx_polar_warp": (_i, [_p, _i, _i, _d, _d, _d, _i, _i, _d, _p, _p]),
    "gx_polar_unwarp": (_i, [_p, _i, _i, _d, _d, _d, _i, _i, _d, _p, _p]),
    "gx_gather_columns": (_i, [_p, _i, _i, _p, _i, _p, _p, _p]),
    "gx_masked_fit_sums": (_i, [_p, _p, _p, _i64, _p, _p]),
    "gx_host_register": (_i, [_p, _i64]),
    "gx_host_unregister": (_i, [_p]),
    "gx_copy_to_host_async": (_i, [_p, _p, _i64, _p]),
    "gx_host_widen_f32_f64": (_i, [_p, _p, _i64, _i]),
    "gx_comm_unique_id": (_i, [_p]),
    "gx_comm_init": (_i, [_p, _i, _i, _p]),
    "gx_comm_destroy": (_i, [_p]),
    "gx_comm_all_reduce": (_i, [_p, _p, _i64, _i, _p]),
    "gx_comm_reduce_scatter_f32": (_i, [_p, _p, _i64, _p]),
    "gx_comm_all_gather_f32": (_i, [_p, _p, _i64, _p]),
    "gx_slab_tiles": (_i64, [_p]),
    "gx_slab_minmax": (_i, [_p, _p, _p]),
    "gx_slab_count": (_i, [_p, _p, _p, _p, _p, _p, _p, _p]),
    "gx_slab_write": (_i, [_p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "gx_rotate_points": (_i, [_p, _p, _p, _p, _i64, _p, _p, _p, _p]),
    "gx_detector_accumulate": (_i, [_p, _i, _i, _i, _d, _d, _d, _d, _p, _p, _p, _i64, _p, _p, _i,
                                    _p, _i, _p, _p]),
    "gx_detector_accumulate_fast": (_i, [_p, _i, _i, _i, _d, _d, _d, _d, _p, _p, _p, _i64, _p, _p, _i,
                                         _p, _i, _p, _p, _p]),
    "gx_grid_affine_fit": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "gx_affine_record_bytes": (_i, []),
    "gx_affine_plan_doubles": (_i, []),
    "gx_host_affine_orientations": (_i, [_p, _p, _i, _i, _p, _p, _i, _d, _d, _d, _d, _i, _i, _i, _p, _p]),
    "gx_detector_accumulate_affine": (_i, [_p, _i, _i, _i, _d, _d, _d, _d, _p, _p, _p, _i, _i, _p, _p, _p, _i,
                                           _p, _p, _i, _p, _p, _p]),
    "gx_detector_accumulate_affine_brick": (_i, [_p, _p, _i, _i, _i, _i, _d, _d, _d, _d, _p, _p, _p, _i, _i, _p, _p, _p,
                                                 _i, _p, _p, _p]),
    "gx_fast_record_bytes": (_i, []),
    "gx_host_fast_orientations": (_i, [_p, _p, _i, _d, _d, _d, _d, _p, _p]),
    "gx_host_orientation_matrices": (_i, [_p, _p, _i, _p]),
    "gx_detector_epilogue": (_i, [_p, _i, _i, _i, _i, _p, _p]),
}

_UNCHECKED = {"gx_fused_wants_zeroed_work", "gx_abi_version", "gx_last_error", "gx_fft_plan_bytes", "gx_fast_record_bytes",
              "gx_affine_record_bytes", "gx_affine_plan_doubles", "gx_slab_tiles"}

_cdll = None


def exported_symbols():
    """Names include/giwaxs_b200.h declares (used by the CPU-side load test)."""
    return sorted(_PROTOTYPES)


def cdll():
    global _cdll
    if _cdll is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "%s is missing: build it with `python -m giwaxsim_b200.build` "
                "(nvcc, sm_100a). giwaxsim_b200 has no CPU fallback." % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _PROTOTYPES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _cdll = lib
    return _cdll


def last_error():
    return cdll().gx_last_error().decode("utf-8", "replace")


# kernels launched per entry point (bench.py reports the total as gpu_launches)
_LAUNCHES = {
    "gx_coords_minmax": 3, "gx_atoms_sort_rows": 3, "gx_slice_bbox": 3, "gx_atom_pixel_indices": 1,
    "gx_slice_vectors": 1, "gx_project_slices": 1, "gx_fft2_abs2_shift": 2, "gx_slice_col_index": 1,
    "gx_axis_col_index": 1, "gx_axis_row_index": 1, "gx_bin_slices": 1, "gx_row_histogram": 1,
    "gx_voxel_finalize": 1, "gx_voxel_shell_scale": 1, "gx_rotate_points": 1, "gx_detector_accumulate": 1, "gx_detector_epilogue": 1,
    "gx_detector_accumulate_fast": 1, "gx_detector_accumulate_affine": 1, "gx_detector_accumulate_affine_brick": 1, "gx_grid_affine_fit": 1,
    "gx_slices_fused": 2, "gx_species_histogram": 1, "gx_species_codes": 1, "gx_checksum64": 1, "gx_row_abs_f_max": 3, "gx_fold_dc": 1, "gx_polar_warp": 1, "gx_polar_unwarp": 1, "gx_gather_columns": 1, "gx_masked_fit_sums": 1, "gx_window_indices": 1, "gx_slab_minmax": 3, "gx_slab_count": 3, "gx_slab_write": 1, "gx_slice_col_range": 1, "gx_extreme_atoms": 1, "gx_hull_filter": 1,
}
_launch_count = 0


def reset_launch_count():
    global _launch_count
    _launch_count = 0


def launch_count():
    return _launch_count


def call(name, *args):
    """Invoke an int-returning entry point and raise GxError on failure."""
    global _launch_count
    rc = getattr(cdll(), name)(*args)
    if name not in _UNCHECKED and rc != GX_OK:
        raise GxError(rc, last_error())
    if name == "gx_slice_yrange":
        _launch_count += 1 if int(args[2]) <= 65536 else 2 + (int(args[5]) + 255) // 256
    elif name == "gx_slices_fused":
        ph = getattr(getattr(args[0], "_obj", None), "phases", 0)
        _launch_count += 2 if ph in (0, 3) else 1
    else:
        _launch_count += _LAUNCHES.get(name, 0)
    return rc


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array / None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        return ctypes.c_void_p(t.data_ptr())
    return ctypes.c_void_p(t.ctypes.data)
