"""TEST INFRASTRUCTURE ONLY -- imports the *unmodified* GIWAXSim reference.

Only usable in the build container, where the read-only reference checkout is
mounted at /root/reference.  Nothing under giwaxsim_b200/ imports this file and
nothing in the `-m gpu` tests, smoke() or bench.py needs it at run time: it is
used (a) by oracle/make_golden.py to generate the committed fixtures under
tests/golden/ and (b) by `-m "not gpu"` tests that pin oracle/giwaxs_oracle.py
against the live reference when the checkout is present.

The reference cannot be imported as-is (matplotlib / fabio / xraydb are not
installed; scipy.signal.tukey moved), so permissive stub modules are inserted
first.  xraydb is replaced by the fixed f'/f'' table in oracle/ftable.py -- the
hot path takes f-values as an *input*, so both sides of every parity test see
the same numbers ("parity unpinned at the xraydb boundary").

The reference's ThreadPoolExecutor fan-out has a lost-update race on the
shared accumulators, so the drivers here call the same worker functions
*serially* (tools/comparison.py:749-762 and :836-853 restated as plain loops).
"""
import os
import sys
import types

import numpy as np

from . import ftable

REFERENCE_ROOT = os.environ.get("GIWAXSIM_REFERENCE", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "tools"))


class _Anything(types.ModuleType):
    """Module whose every attribute is a callable that returns another stub."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        stub = _Anything(self.__name__ + "." + name)
        setattr(self, name, stub)
        return stub

    def __call__(self, *a, **k):
        return _Anything(self.__name__ + "()")


_loaded = None


def load():
    """Return the reference's tools.* modules as a namespace."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker", "matplotlib.cm",
                 "matplotlib.colors", "mpl_toolkits", "mpl_toolkits.mplot3d", "fabio", "lmfit"):
        if name not in sys.modules:
            sys.modules[name] = _Anything(name)
    if "xraydb" not in sys.modules:
        x = types.ModuleType("xraydb")
        x.f1_chantler = lambda el, energy: ftable.f1_f2(el, energy)[0]
        x.f2_chantler = lambda el, energy: ftable.f1_f2(el, energy)[1]
        sys.modules["xraydb"] = x
    import scipy.signal
    import scipy.signal.windows
    if not hasattr(scipy.signal, "tukey"):
        scipy.signal.tukey = scipy.signal.windows.tukey
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import tools.utilities as utilities
        import tools.detector as detector
        import tools.voxelgrids as voxelgrids
        import tools.comparison as comparison
    _loaded = types.SimpleNamespace(utilities=utilities, detector=detector,
                                    voxelgrids=voxelgrids, comparison=comparison)
    return _loaded


# --------------------------------------------------------------------------
# serial drivers around the unmodified reference workers
# --------------------------------------------------------------------------
def stage_a_setup(coords, elements, r_voxel_size, q_voxel_size, max_q, energy):
    """The scalar set-up of voxelgridmaker_fitting (tools/comparison.py:705-744),
    evaluated with the reference's own expressions so every rounding is its."""
    ref = load()
    ptable = ref.comparison.ptable
    max_q_diag = np.sqrt(2) * max_q
    grid_size = int(np.ceil(2 * np.pi / (q_voxel_size * r_voxel_size)))
    x_bound = np.max(coords[:, 0]) - np.min(coords[:, 0])
    y_bound = np.max(coords[:, 1]) - np.min(coords[:, 1])
    z_bound = np.max(coords[:, 2]) - np.min(coords[:, 2])
    max_q_diag = max_q_diag + max_q_diag % q_voxel_size
    q_num = ((2 * max_q_diag / q_voxel_size) + 1).astype(int)
    if q_num % 2 == 0:
        q_num += 1
    q_axis = np.linspace(-max_q_diag, max_q_diag, q_num)
    delta_phi_rad = np.arctan(q_voxel_size / max_q_diag)
    phi_num = np.ceil(2 * np.pi / delta_phi_rad).astype(int)
    last_phi = 180 - (180 / phi_num)
    phis = np.linspace(0, last_phi, num=phi_num)
    f1_f2_dict = ref.utilities.get_element_f1_f2_dict(energy, elements)
    f_values = np.array([f1_f2_dict[e] for e in elements], dtype=complex)
    f_values += np.array([ptable[e] for e in elements])
    avg_voxel_f = (np.sum(f_values) / (x_bound * y_bound * z_bound)) * r_voxel_size ** 3
    return dict(grid_size=grid_size, x_bound=x_bound, y_bound=y_bound, z_bound=z_bound,
                q_num=int(q_num), q_axis=q_axis, phis=phis, f_values=f_values,
                avg_voxel_f=avg_voxel_f)


def run_slices_serial(coords, setup, r_voxel_size, fill_bkg, smooth, phis=None, capture=None):
    """Call the reference's rotate_project_fft_coords once per phi, serially, on
    two fresh shared-memory accumulators; return (sum, count) as float64 copies.

    capture: optional dict; when given, per-slice intermediates are recorded by
    wrapping the module globals `fftn` (pre-FFT grid) and `process_file2`
    (iq_2d and the three q axes), both called by bare name in
    tools/voxelgrids.py:388,413.
    """
    ref = load()
    vg = ref.voxelgrids
    q_num = setup["q_num"]
    q = setup["q_axis"]
    shm_sum = ref.utilities.create_shared_array((q_num, q_num, q_num))
    shm_cnt = ref.utilities.create_shared_array((q_num, q_num, q_num))
    orig_fftn, orig_pf2 = vg.fftn, vg.process_file2
    if capture is not None:
        capture.setdefault("grid", [])
        capture.setdefault("iq_2d", [])
        capture.setdefault("det_h_qx", [])
        capture.setdefault("det_h_qy", [])
        capture.setdefault("det_v_qz", [])

        def fftn_wrap(a, *args, **kw):
            capture["grid"].append(np.array(a, copy=True))
            return orig_fftn(a, *args, **kw)

        def pf2_wrap(iq_2d, hx, hy, vz, *rest):
            capture["iq_2d"].append(np.array(iq_2d, copy=True))
            capture["det_h_qx"].append(np.array(hx, copy=True))
            capture["det_h_qy"].append(np.array(hy, copy=True))
            capture["det_v_qz"].append(np.array(vz, copy=True))
            return orig_pf2(iq_2d, hx, hy, vz, *rest)

        vg.fftn, vg.process_file2 = fftn_wrap, pf2_wrap
    try:
        for phi in (setup["phis"] if phis is None else phis):
            vg.rotate_project_fft_coords(
                (coords, setup["f_values"], phi, setup["grid_size"], r_voxel_size,
                 setup["avg_voxel_f"], setup["x_bound"], setup["y_bound"], setup["z_bound"],
                 fill_bkg, smooth, q, q, q, shm_sum.name, shm_cnt.name))
        vsum = np.ndarray((q_num,) * 3, dtype=np.float64, buffer=shm_sum.buf).copy()
        vcnt = np.ndarray((q_num,) * 3, dtype=np.float64, buffer=shm_cnt.buf).copy()
    finally:
        vg.fftn, vg.process_file2 = orig_fftn, orig_pf2
        for s in (shm_sum, shm_cnt):
            s.close()
            s.unlink()
    return vsum, vcnt


def finalize_serial(vsum, vcnt, q_axis, max_q):
    """tools/comparison.py:769-786 on explicit arrays."""
    ref = load()
    iq = np.divide(vsum, vcnt, out=np.zeros_like(vsum), where=vcnt != 0)
    iq_s, qx_s, qy_s, qz_s = ref.voxelgrids.downselect_voxelgrid(iq, q_axis, q_axis, q_axis, max_q)
    iq_s = ref.voxelgrids.add_f0_q_3d(iq_s, qx_s, qy_s, qz_s, "C")
    return iq_s, qx_s, qy_s, qz_s


def voxelgridmaker_serial(coords, elements, r_voxel_size, q_voxel_size, max_q, energy,
                          fill_bkg=False, smooth=0, phis=None):
    setup = stage_a_setup(coords, elements, r_voxel_size, q_voxel_size, max_q, energy)
    vsum, vcnt = run_slices_serial(coords, setup, r_voxel_size, fill_bkg, smooth, phis=phis)
    return finalize_serial(vsum, vcnt, setup["q_axis"], max_q) + (vsum, vcnt, setup)


def detector_base(num_pixels, max_q, angle_init_vals, angle_init_axs):
    """make_detector + the three optional init rotations (comparison.py:798-818)."""
    ref = load()
    d = ref.detector
    det_x, det_y, det_z, det_h, det_v = d.make_detector(max_q, num_pixels, max_q, num_pixels)
    fn = {"psi": d.rotate_about_normal, "phi": d.rotate_about_vertical,
          "theta": d.rotate_about_horizontal}
    for val, ax in zip(angle_init_vals, angle_init_axs):
        if ax in fn:
            det_x, det_y, det_z = fn[ax](det_x, det_y, det_z, val)
    return det_x, det_y, det_z, det_h, det_v


def detectormaker_serial(iq, qx, qy, qz, num_pixels, max_q, angle_init_vals, angle_init_axs,
                         psis, psi_w, phis, phi_w, thetas, theta_w, mirror=True, raw=False):
    """detectormaker_fitting (comparison.py:790-870) with the orientation loop
    run serially through the reference's own rotate/intersect functions."""
    ref = load()
    d = ref.detector
    det_x, det_y, det_z, det_h, det_v = detector_base(num_pixels, max_q, angle_init_vals, angle_init_axs)
    acc = np.zeros((num_pixels, num_pixels))
    for psi, wp in zip(psis, psi_w):
        for phi, wf in zip(phis, phi_w):
            for theta, wt in zip(thetas, theta_w):
                x2, y2, z2 = d.rotate_psi_phi_theta(det_x, det_y, det_z, psi, phi, theta)
                det_int = d.intersect_detector(iq, qx, qy, qz, x2, y2, z2)
                det_int *= wp * wf * wt
                acc += det_int
    if raw:
        return acc, det_h, det_v
    det_sum = acc
    if mirror:
        det_sum = d.mirror_vertical_horizontal(det_sum)
    det_sum[det_sum != det_sum] = 1e-6
    det_sum[det_sum <= 0] = 1e-6
    det_sum *= 1e-6
    return det_sum, det_h, det_v


class _SerialExecutor:
    """Stand-in for ThreadPoolExecutor that runs every task at submit time: the reference's
    threaded accumulation loses updates, parity is defined against the serial order."""

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def submit(self, fn, *args, **kwargs):
        from concurrent.futures import Future
        fut = Future()
        fut.set_result(fn(*args, **kwargs))
        return fut


def voxel_grid_low_mem_serial(input_path, r_voxel_size, q_voxel_size, max_q, aff_num_qs, energy,
                              fill_bkg=False, smooth=0):
    """The reference's generate_voxel_grid_low_mem (tools/voxelgrids.py:535-722), unmodified,
    with its thread pool replaced by serial execution."""
    import contextlib
    import io
    ref = load()
    saved = ref.voxelgrids.ThreadPoolExecutor
    ref.voxelgrids.ThreadPoolExecutor = _SerialExecutor
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return ref.voxelgrids.generate_voxel_grid_low_mem(input_path, r_voxel_size, q_voxel_size, max_q,
                                                              aff_num_qs, energy, "shim", fill_bkg=fill_bkg,
                                                              smooth=smooth)
    finally:
        ref.voxelgrids.ThreadPoolExecutor = saved
