"""Parity of the CUDA path (through the C ABI) against the fixtures generated
from the unmodified reference and against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star): atom pixel indices, q-bin indices and
per-voxel counts BIT-EXACT; voxel-grid and detector intensities within 1e-4
of the reference maximum (fp32 on the device).
"""
import numpy as np
import pytest
import torch

from giwaxsim_b200 import engine
from giwaxsim_b200.tools import comparison, detector, utilities, voxelgrids
from oracle import giwaxs_oracle as ox

pytestmark = pytest.mark.gpu

TOL_GRID = 1e-5      # pre-FFT grid, relative to max |grid|
TOL_INT = 1e-4       # intensities, relative to the reference maximum
CASES = ["graphite262", "silicon256", "clipped128"]


def _engine_for(g, count3d=False, window=None):
    """SliceEngine on the fixture's slab with the reference-derived scalars."""
    coords = g["coords"]
    codes, uniq = engine.encode_values(g["f_values"])
    b = g["bounds"]
    return engine.SliceEngine(coords, float(g["r"]), g["q_axis"], int(g["grid_size"]), complex(g["avg_voxel_f"]),
                              b[0], b[1], bool(g["fill_bkg"]), int(g["smooth"]), species=codes, table=uniq,
                              count3d=count3d, window=window)


@pytest.mark.parametrize("name", CASES)
def test_T1_atom_pixel_indices_bit_exact(golden, name):
    g = golden(name + ".npz")
    eng = _engine_for(g)
    N = int(g["grid_size"])
    # bounds computed on the device equal NumPy's max - min
    assert np.array_equal(np.array(eng.atoms.bounds), g["bounds"])
    for i in g["probe"]:
        y, z, bbox = eng.atom_indices(g["phis"][i])
        ry, rz = g["y_idx_%d" % i], g["z_idx_%d" % i]
        assert np.array_equal(y, ry)
        assert np.array_equal(np.minimum(z, N), np.minimum(rz, N))
        assert np.array_equal(bbox, g["bbox_%d" % i])


@pytest.mark.parametrize("name", CASES)
def test_T2_T3_T4_slice_intermediates(golden, name):
    g = golden(name + ".npz")
    eng = _engine_for(g)
    N, q_num = int(g["grid_size"]), int(g["q_num"])
    probe = [int(i) for i in g["probe"]]
    cap = {"want_grids": True}
    eng.run(g["phis"][probe], capture=cap)
    grids, iqs, cols = cap["grid"][0], cap["iq_2d"][0], cap["col"][0]
    row = eng.row_index.cpu().numpy()
    for k, i in enumerate(probe):
        ref_grid = g["grid_%d" % i]
        err = np.abs(grids[k] - ref_grid).max() / float(g["grid_absmax_%d" % i])
        assert err <= TOL_GRID, "pre-FFT grid slice %d: %.3g" % (i, err)
        ref_iq = g["iq2d_%d" % i]
        err = np.abs(iqs[k].astype(np.float64) - ref_iq).max() / float(g["iq2d_max_%d" % i])
        assert err <= TOL_INT, "slice intensity %d: %.3g" % (i, err)
        # q-bin indices: bit-exact
        cm = g["colmask_%d" % i]
        assert np.array_equal(cols[k] >= 0, cm)
        assert np.array_equal(cols[k][cm], g["iy_%d" % i] * q_num + g["ix_%d" % i])
        rm = g["rowmask_%d" % i]
        assert np.array_equal(row >= 0, rm)
        assert np.array_equal(row[rm], g["iz_%d" % i])


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("count3d", [False, True])
def test_T5_accumulators(golden, name, count3d):
    g = golden(name + ".npz")
    eng = _engine_for(g, count3d=count3d)
    eng.run(g["phis"])
    assert np.array_equal(eng.counts(), g["vcnt"].astype(np.int64))          # bit-exact
    err = np.abs(eng.sums().astype(np.float64) - g["vsum"]).max() / float(g["vsum_max"])
    assert err <= TOL_INT, err


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("count3d", [False, True])
def test_T5_accumulators_crop_window(golden, name, count3d):
    """The production driver accumulates only the voxels downselect_voxelgrid keeps: counts and
    sums of that window equal the reference's full accumulators cropped to it."""
    g = golden(name + ".npz")
    lo, hi = engine.crop_range(g["q_axis"], float(g["max_q"]))
    assert hi - lo == len(g["q_crop"]) and hi - lo < len(g["q_axis"])
    eng = _engine_for(g, count3d=count3d, window=(lo, hi))
    eng.run(g["phis"])
    ref_cnt = g["vcnt"].astype(np.int64)[lo:hi, lo:hi, lo:hi]
    ref_sum = g["vsum"][lo:hi, lo:hi, lo:hi]
    assert np.array_equal(eng.counts(), ref_cnt)                               # bit-exact
    err = np.abs(eng.sums().astype(np.float64) - ref_sum).max() / float(g["vsum_max"])
    assert err <= TOL_INT, err
    assert eng.KC <= _engine_for(g).KC


@pytest.mark.parametrize("name", CASES)
def test_T6_voxelgridmaker_fitting(golden, name):
    g = golden(name + ".npz")
    iq, qx, qy, qz = comparison.voxelgridmaker_fitting(
        g["coords"], g["elements"], float(g["r"]), float(g["q"]), float(g["max_q"]), float(g["energy"]),
        fill_bkg=bool(g["fill_bkg"]), smooth=int(g["smooth"]))
    assert iq.dtype == np.float64 and iq.shape == g["iq"].shape
    for a in (qx, qy, qz):
        assert np.array_equal(a, g["q_crop"])                                 # axes bit-exact
    err = np.abs(iq - g["iq"]).max() / g["iq"].max()
    assert err <= TOL_INT, err
    assert np.array_equal(iq == 0, g["iq"] == 0)                              # empty voxels stay exactly 0


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_T7_detectormaker_fitting(golden, tag, tmp_path):
    g = golden("detector.npz")
    iq, q = g["iq"], g["q"]
    P = int(g[tag + "_P"])
    vals, axs = tuple(g[tag + "_vals"]), tuple(str(a) for a in g[tag + "_axs"])
    paths = []
    for k in ("_pw", "_fw", "_tw"):
        p = str(tmp_path / (tag + k + ".npy"))
        np.save(p, g[tag + k])
        paths.append(p)
    # base grids (make_detector + init rotations): bit-exact
    gx, gy, gz, h, v = comparison.detector_base_device(P, float(g["max_q"]), vals, axs, engine.resolve_device())
    for a, k in ((gx, "_gx"), (gy, "_gy"), (gz, "_gz")):
        assert np.array_equal(a.cpu().numpy(), g[tag + k])
    # per-orientation voxel indices: bit-exact
    det = engine.DetectorEngine(iq, q, q, q)
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), g[tag + "_psis"], g[tag + "_pw"],
                                     g[tag + "_phis"], g[tag + "_fw"], g[tag + "_thetas"], g[tag + "_tw"])
    for o in g[tag + "_probes"]:
        _, index = det.accumulate(gx, gy, gz, R, w, probe=int(o))
        assert np.array_equal(index.cpu().numpy(), g["%s_index_%d" % (tag, o)])
    raw, _ = det.accumulate(gx, gy, gz, R, w)
    ref_raw = g[tag + "_raw"]
    assert np.abs(raw.cpu().numpy().reshape(P, P) - ref_raw).max() <= TOL_INT * ref_raw.max()
    out, dh, dv = comparison.detectormaker_fitting(iq, q, q, q, P, float(g["max_q"]), vals, axs, g[tag + "_psis"],
                                                   paths[0], g[tag + "_phis"], paths[1], g[tag + "_thetas"],
                                                   paths[2], mirror=bool(g[tag + "_mirror"]))
    ref = g[tag + "_final"]
    assert out.dtype == np.float64 and out.shape == ref.shape
    assert np.abs(out - ref).max() <= TOL_INT * ref.max()
    assert np.array_equal(dh, np.linspace(-float(g["max_q"]), float(g["max_q"]), P))


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_detector_module_functions(golden, tag):
    """make_detector / rotate_about_* / intersect_detector / mirror as free functions."""
    g = golden("detector.npz")
    iq, q = g["iq"], g["q"]
    P = int(g[tag + "_P"])
    gx, gy, gz, h, v = detector.make_detector(float(g["max_q"]), P, float(g["max_q"]), P)
    ox_base = ox.make_detector(float(g["max_q"]), P, float(g["max_q"]), P)
    for a, b in zip((gx, gy, gz, h, v), ox_base):
        assert np.array_equal(a, b)
    rot = {"psi": detector.rotate_about_normal, "phi": detector.rotate_about_vertical,
           "theta": detector.rotate_about_horizontal}
    for val, ax in zip(g[tag + "_vals"], g[tag + "_axs"]):
        if str(ax) in rot:
            gx, gy, gz = rot[str(ax)](gx, gy, gz, float(val))
    for a, k in ((gx, "_gx"), (gy, "_gy"), (gz, "_gz")):
        assert np.array_equal(a, g[tag + k])
    got = detector.intersect_detector(iq, q, q, q, gx, gy, gz)
    ref = g[tag + "_intersect"]
    assert np.abs(got - ref).max() <= 1e-6 * ref.max()          # fp32 copy of the same voxel
    x2, y2, z2 = detector.rotate_psi_phi_theta(gx, gy, gz, 80.0, 33.0, 1.0)
    ex = ox.rotate_psi_phi_theta(gx, gy, gz, 80.0, 33.0, 1.0)
    assert all(np.array_equal(a, b) for a, b in zip((x2, y2, z2), ex))
    m = detector.mirror_vertical_horizontal(g[tag + "_raw"])
    assert np.abs(m - g[tag + "_mirrored"]).max() <= 1e-15 * g[tag + "_mirrored"].max()


def test_worker_api_rotate_project_fft_coords(golden):
    """The reference's per-slice worker signature with named accumulators."""
    g = golden("clipped128.npz")
    b = g["bounds"]
    q_axis, q_num = g["q_axis"], int(g["q_num"])
    shm_sum = utilities.create_shared_array((q_num,) * 3)
    shm_cnt = utilities.create_shared_array((q_num,) * 3)
    coords, f = g["coords"], g["f_values"]
    phis = g["phis"][::9]
    for phi in phis:
        voxelgrids.rotate_project_fft_coords(
            (coords, f, phi, int(g["grid_size"]), float(g["r"]), complex(g["avg_voxel_f"]), b[0], b[1], b[2],
             bool(g["fill_bkg"]), int(g["smooth"]), q_axis, q_axis, q_axis, shm_sum.name, shm_cnt.name))
    got_sum = np.ndarray((q_num,) * 3, dtype=np.float64, buffer=shm_sum.buf)
    got_cnt = np.ndarray((q_num,) * 3, dtype=np.float64, buffer=shm_cnt.buf)
    setup = ox.stage_a_setup(coords, f, float(g["r"]), float(g["q"]), float(g["max_q"]))
    vsum, vcnt = np.zeros((q_num,) * 3), np.zeros((q_num,) * 3)
    for phi in phis:
        ox.run_slice(vsum, vcnt, coords, setup, float(g["r"]), phi, bool(g["fill_bkg"]), int(g["smooth"]))
    assert np.array_equal(got_cnt, vcnt)
    assert np.abs(got_sum - vsum).max() <= TOL_INT * vsum.max()
    shm_sum.close(); shm_sum.unlink(); shm_cnt.close(); shm_cnt.unlink()


def test_worker_api_process_file2_and_finalisers():
    rng = np.random.default_rng(9)
    N, q_num = 96, 31
    q_axis = np.linspace(-1.55, 1.55, q_num)
    iq_2d = rng.random((N, N)) * 1e5
    hx = np.linspace(-1.1, 1.1, N)
    hy = np.linspace(-1.9, 1.9, N)
    vz = ox.fft_q_axis(N, 1.4)
    shm_sum = utilities.create_shared_array((q_num,) * 3)
    shm_cnt = utilities.create_shared_array((q_num,) * 3)
    for _ in range(2):
        voxelgrids.process_file2(iq_2d, hx, hy, vz, q_axis, q_axis, q_axis, shm_sum.name, shm_cnt.name)
    vsum, vcnt = np.zeros((q_num,) * 3), np.zeros((q_num,) * 3)
    for _ in range(2):
        ox.bin_slice(vsum, vcnt, iq_2d, hx, hy, vz, q_axis)
    assert np.array_equal(shm_cnt.to_numpy(), vcnt)
    assert np.abs(shm_sum.to_numpy() - vsum).max() <= 1e-6 * vsum.max()
    small, ax, ay, az = voxelgrids.downselect_voxelgrid(vsum, q_axis, q_axis, q_axis, 1.0)
    lo, hi = ox.crop_range(q_axis, 1.0)
    assert np.array_equal(small, vsum[lo:hi, lo:hi, lo:hi]) and np.array_equal(ax, q_axis[lo:hi])
    got = voxelgrids.add_f0_q_3d(small, ax, ay, az, "C")
    ref = small * ox.carbon_f0_factor(ax, ay, az)
    assert np.abs(got - ref).max() <= 1e-6 * ref.max()


def test_worker_api_generate_detector_ints(golden):
    g = golden("detector.npz")
    iq, q = g["iq"], g["q"]
    P = 40
    gx, gy, gz, _, _ = ox.detector_base(P, 2.0, (90.0, 90.0, 90.0), ("psi", "phi", "psi"))
    shm = utilities.create_shared_array((P, P))
    todo = [(80.0, 0.3, 10.0, 0.5, 0.0, 1.0), (90.0, 0.7, 100.0, 0.5, 1.0, 1.0)]
    ref = np.zeros((P, P))
    for psi, wp, phi, wf, theta, wt in todo:
        detector.generate_detector_ints((iq, q, q, q, gx, gy, gz, psi, wp, phi, wf, theta, wt, shm.name))
        part = ox.intersect_detector(iq, q, q, q, *ox.rotate_psi_phi_theta(gx, gy, gz, psi, phi, theta))
        ref += part * (wp * wf * wt)
    got = np.ndarray((P, P), dtype=np.float64, buffer=shm.buf)
    assert np.abs(got - ref).max() <= 1e-6 * ref.max()
    shm.unlink()


@pytest.mark.parametrize("N", [16, 64, 100, 256, 262, 1024, 1048, 2048, 2095])
def test_fft2_abs2_shift_against_numpy(N):
    """K2 alone on random complex grids with a pedestal, pow2 and Bluestein sizes."""
    from giwaxsim_b200._lib import call, ptr
    dev = engine.resolve_device()
    rng = np.random.default_rng(N)
    batch = 3
    x = (rng.normal(size=(batch, N, N)) + 1j * rng.normal(size=(batch, N, N)) + 5.0).astype(np.complex64)
    plan = engine.FftPlan.get(N, dev)
    d_x = torch.from_numpy(x.view(np.float32)).to(dev)
    work = torch.empty_like(d_x)
    out = torch.empty(batch * N * N, dtype=torch.float32, device=dev)
    call("gx_fft2_abs2_shift", ptr(d_x), ptr(work), ptr(out), batch, N, ptr(plan.table), 0.0, 0.0, None)
    torch.cuda.synchronize()
    got = out.cpu().numpy().reshape(batch, N, N).astype(np.float64)
    for b in range(batch):
        ref = np.abs(np.fft.fftshift(np.fft.fftn(x[b].astype(np.complex128)))) ** 2
        assert np.abs(got[b] - ref).max() <= TOL_INT * ref.max()
        # away from the DC peak the error budget is much tighter in practice
        off = ref < 1e-3 * ref.max()
        assert np.abs(got[b] - ref)[off].max() <= 1e-4 * ref[off].max()


def test_random_slab_against_oracle_many_species_and_generic_path():
    """Seeded random slab: counting path (5 species) and the generic per-atom-f
    path (> GX_MAX_SPECIES distinct values) against the oracle."""
    rng = np.random.default_rng(77)
    A = 5000
    coords = rng.random((A, 3)) * [40.0, 25.0, 30.0]
    r, q, max_q = 0.4, 2 * np.pi / (0.4 * 127.5), 1.2
    for generic in (False, True):
        if generic:
            f = (rng.integers(1, 30, A) + rng.random(A) * 0.1 + 1j * rng.random(A) * 0.05)
        else:
            f = rng.choice(np.array([6.0049 + 0.0023j, 1.0 + 0j, 16.17 + 0.25j, 8.016 + 0.009j, 9.024 + 0.014j]), A)
        iq, qx, qy, qz, vsum, vcnt, setup = ox.voxelgridmaker(coords, f, r, q, max_q, True, 4)
        codes, uniq = engine.encode_values(f)
        assert (codes is None) == generic
        kw = dict(f_values=f) if generic else dict(species=codes, table=uniq)
        eng = engine.SliceEngine(coords, r, setup["q_axis"], setup["grid_size"], setup["avg_voxel_f"],
                                 setup["x_bound"], setup["y_bound"], True, 4, **kw)
        eng.run(setup["phis"])
        assert np.array_equal(eng.counts(), vcnt.astype(np.int64))
        assert np.abs(eng.sums() - vsum).max() <= TOL_INT * vsum.max()


def test_T8_pm6_default_config_end_to_end(golden):
    """config_templates/simulate_GIWAXS_config.txt (PM6 slab, N=1048 -> Bluestein, 892 slices,
    2880 orientations x 500^2) through the two drop-in drivers, against the oracle image and the
    reference's own shipped golden det_sum.npy (which differs by the unpinned xraydb table)."""
    g = golden("pm6.npz")
    names = [str(n) for n in g["element_names"]]
    elements = np.array(names)[g["element_codes"]]
    iq, qx, qy, qz = comparison.voxelgridmaker_fitting(g["coords"], elements, float(g["r"]), float(g["q"]),
                                                       float(g["max_q"]), 12700.0, fill_bkg=True,
                                                       smooth=int(g["smooth"]))
    assert iq.shape == tuple(g["iq_shape"])
    plane = iq[:, :, iq.shape[2] // 2]
    assert np.abs(plane - g["iq_center_plane"]).max() <= TOL_INT * float(g["iq_max"])
    det, h, v = comparison.detectormaker_fitting(iq, qx, qy, qz, int(g["P"]), float(g["max_q"]), tuple(g["vals"]),
                                                 tuple(str(a) for a in g["axs"]), g["psis"], None, g["phis"], None,
                                                 g["thetas"], None, mirror=True)
    quad = det[np.ix_(np.where(v >= 0)[0], np.where(h >= 0)[0])]      # simulate_GIWAXS.py:172-177
    ref = g["det_quadrant_oracle"]
    assert np.array_equal(h[h >= 0], g["det_h"])
    assert np.abs(quad - ref).max() <= TOL_INT * ref.max()
    gold = golden("pm6_det_sum_ref.npy")
    assert np.abs(quad - gold).max() <= 0.02 * gold.max()
    assert np.corrcoef(np.log(quad).ravel(), np.log(gold).ravel())[0, 1] > 0.9999


def _write_cell_files(tmp_path):
    """A small triclinic-capable unit cell as .xyz and .pdb (same atoms)."""
    rng = np.random.default_rng(8)
    n = 37
    xyz = np.round(rng.random((n, 3)) * [6.1, 7.3, 5.2], 3)
    el = rng.choice(np.array(["C", "H", "S", "O", "Si"]), size=n)
    px, pp = str(tmp_path / "cell.xyz"), str(tmp_path / "cell.pdb")
    with open(px, "w") as fh:
        fh.write("%d\ncomment line\n" % n)
        for e, p in zip(el, xyz):
            fh.write("%s1 %.3f %.3f %.3f\n" % (e, p[0], p[1], p[2]))
        fh.write("bad line\n")
    with open(pp, "w") as fh:
        fh.write("CRYST1    6.100    7.300    5.200  90.00  90.00  90.00 P 1           1\n")
        for i, (e, p) in enumerate(zip(el, xyz)):
            fh.write("ATOM  %5d  %-3s MOL A   1    %8.3f%8.3f%8.3f  1.00  0.00          %2s\n"
                     % (i + 1, e, p[0], p[1], p[2], e))
        fh.write("END\n")
    return px, pp


@pytest.mark.parametrize("cell", [(6.1, 7.3, 5.2, 90.0, 90.0, 90.0), (6.1, 7.3, 5.2, 82.0, 96.0, 107.0)])
@pytest.mark.parametrize("kind", ["xyz", "pdb"])
def test_slabmaker_fitting_bit_exact_and_resident_handoff(tmp_path, cell, kind):
    """Next row N3: the device slab builder returns the reference's array (order and every
    bit of every coordinate), and the slab it leaves on the device gives the same voxel grid
    as re-uploading the returned arrays."""
    px, pp = _write_cell_files(tmp_path)
    path = px if kind == "xyz" else pp
    size = (33.0, 41.0, 28.0)
    c0, e0 = ox.read_structure(path)
    o_coords, o_el = ox.slabmaker(c0, e0, *size, *cell)
    coords, el = comparison.slabmaker_fitting(path, *size, *cell)
    assert coords.dtype == np.float64 and coords.shape == o_coords.shape
    assert np.array_equal(coords, o_coords) and np.array_equal(el, o_el)
    previous = utilities._f1f2_provider
    utilities.set_f1f2_provider(lambda e, en=None: (0.01 * len(e), 0.002))
    try:
        r, q, max_q = 0.3, 0.1, 1.5
        a = comparison.voxelgridmaker_fitting(coords, el, r, q, max_q, 12700.0, fill_bkg=True, smooth=2)
        b = comparison.voxelgridmaker_fitting(coords.copy(), el.copy(), r, q, max_q, 12700.0, fill_bkg=True, smooth=2)
        assert np.abs(a[0] - b[0]).max() <= 1e-6 * b[0].max() and np.array_equal(a[0] == 0, b[0] == 0)
        f = ox.f_values_for(el, table=lambda e, en=None: (0.01 * len(e), 0.002))
        o_iq = ox.voxelgridmaker(o_coords, f, r, q, max_q, True, 2)[0]
        assert np.abs(a[0] - o_iq).max() <= TOL_INT * o_iq.max()
        # an in-place edit of the returned array must not be served from the stale device copy
        coords2, el2 = comparison.slabmaker_fitting(path, *size, *cell)
        coords2 *= 1.01
        c = comparison.voxelgridmaker_fitting(coords2, el2, r, q, max_q, 12700.0, fill_bkg=True, smooth=2)
        d = comparison.voxelgridmaker_fitting(coords2.copy(), el2.copy(), r, q, max_q, 12700.0, fill_bkg=True, smooth=2)
        assert c[0].shape == d[0].shape and np.abs(c[0] - d[0]).max() <= 1e-6 * d[0].max()
        # partial edits - a few coordinates, a few element symbols - are seen too (whole-array checksums)
        coords3, el3 = comparison.slabmaker_fitting(path, *size, *cell)
        coords3[len(coords3) // 3:len(coords3) // 3 + 5, 1] += 0.8
        swap = np.flatnonzero(el3 != "S")[7:12]
        el3[swap] = "S"
        e3 = comparison.voxelgridmaker_fitting(coords3, el3, r, q, max_q, 12700.0, fill_bkg=True, smooth=2)
        g3 = comparison.voxelgridmaker_fitting(coords3.copy(), el3.copy(), r, q, max_q, 12700.0, fill_bkg=True, smooth=2)
        assert np.abs(e3[0] - g3[0]).max() <= 1e-6 * g3[0].max()
        assert np.abs(e3[0] - a[0]).max() > 1e-6 * a[0].max()           # and the edit changes the grid at all
        # the voxel grid handed to detectormaker_fitting: edited in place -> the host array is read
        det_args = (48, max_q, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), np.linspace(70, 90, 3), None,
                    np.linspace(0, 120, 4), None, np.array([0.0]), None)
        iq, qx, qy, qz = a
        img0 = comparison.detectormaker_fitting(iq, qx, qy, qz, *det_args)[0]          # resident copy
        img0b = comparison.detectormaker_fitting(iq.copy(), qx, qy, qz, *det_args)[0]  # uploaded
        assert np.abs(img0 - img0b).max() <= 1e-6 * img0b.max()
        iq[iq.shape[0] // 2 - 3:iq.shape[0] // 2 + 3] *= 0.25                             # a mask over a few planes
        img1 = comparison.detectormaker_fitting(iq, qx, qy, qz, *det_args)[0]
        img1b = comparison.detectormaker_fitting(iq.copy(), qx, qy, qz, *det_args)[0]
        assert np.abs(img1 - img1b).max() <= 1e-6 * img1b.max()
        assert np.abs(img1 - img0).max() > 1e-3 * img0.max()
    finally:
        utilities.set_f1f2_provider(previous)


def test_slabmaker_fitting_errors(tmp_path):
    px, _ = _write_cell_files(tmp_path)
    with pytest.raises(Exception, match="must be a .pdb or .xyz"):
        comparison.slabmaker_fitting(str(tmp_path / "cell.cif"), 10, 10, 10, 6.1, 7.3, 5.2, 90, 90, 90)
    # a slab size of zero keeps only the exact mid-plane: nothing survives -> the reference's np.min error
    with pytest.raises(ValueError, match="zero-size array"):
        comparison.slabmaker_fitting(px, 1e-9, 1e-9, 1e-9, 6.1, 7.3, 5.2, 90, 90, 90)


def test_simulate_config_driver_end_to_end(tmp_path):
    """The simulate_GIWAXS.py configuration interface on the B200 path: .pdb unit cell -> slab ->
    voxel grid -> detector quadrant on disk, against the oracle pipeline on the same inputs."""
    from giwaxsim_b200 import simulate
    _, pp = _write_cell_files(tmp_path)
    save = tmp_path / "out"
    cfg = tmp_path / "cfg.txt"
    cfg.write_text("input_filepath=%s\nx_size=30\ny_size=36\nz_size=26\nr_voxel_size=0.3\nq_voxel_size=0.1\nmax_q=1.5\n"
                   "energy=12700\nfill_bkg=True\nsmooth=2\nnum_pixels=60\nangle_init_val1=90\nangle_init_val2=90\n"
                   "angle_init_val3=90\nangle_init_ax1=psi\nangle_init_ax2=phi\nangle_init_ax3=psi\n"
                   "psi_start=70\npsi_end=90\npsi_num=5\nphi_start=0\nphi_end=170\nphi_num=6\ntheta_start=0\n"
                   "theta_end=0\ntheta_num=1\nmirror=True\nsave_folder=%s\n" % (pp, save))
    config = utilities.parse_config_file(str(cfg))
    out = simulate.main(config)
    det_sum, det_h, det_v = out["cell"]
    on_disk = np.load(str(save / "cell" / "det_sum.npy"))
    assert np.array_equal(on_disk, det_sum) and (save / "config.txt").exists()
    assert np.array_equal(np.load(str(save / "cell" / "det_h.npy")), det_h)
    # oracle: same steps (comparison.py:595-870, simulate_GIWAXS.py:172-177)
    cell = utilities.load_pdb_cell_params(pp)
    c0, e0 = ox.read_structure(pp)
    coords, el = ox.slabmaker(c0, e0, 30.0, 36.0, 26.0, *cell)
    from giwaxsim_b200 import synth
    f = ox.f_values_for(el, table=synth.fixed_f1f2)
    iq, qx, qy, qz = ox.voxelgridmaker(coords, f, 0.3, 0.1, 1.5, True, 2)[:4]
    psis, phis, thetas = np.linspace(70, 90, 5), np.linspace(0, 170, 6), np.linspace(0, 0, 1)
    ones = lambda a: np.ones_like(a) / len(a)
    o_det, o_h, o_v = ox.detectormaker(iq, qx, qy, qz, 60, 1.5, (90.0, 90.0, 90.0), ("psi", "phi", "psi"),
                                       psis, ones(psis), phis, ones(phis), thetas, ones(thetas))
    kh, kv = np.where(o_h >= 0)[0], np.where(o_v >= 0)[0]
    assert np.array_equal(det_h, o_h[kh]) and np.array_equal(det_v, o_v[kv])
    ref = o_det[np.ix_(kv, kh)]
    assert det_sum.shape == ref.shape
    assert np.abs(det_sum - ref).max() <= TOL_INT * ref.max()


# ---- two-step command line, .npy hand-off, aff_num_qs > 1 ----
def _write_xyz(path, coords, elements):
    with open(path, "w") as fh:
        fh.write("%d\ncluster\n" % len(coords))
        for el, (x, y, z) in zip(elements, coords):
            fh.write("%s %.17g %.17g %.17g\n" % (el, x, y, z))


def _twostep_files(golden, tmp_path):
    g = golden("twostep.npz")
    folder = tmp_path / "structures"
    folder.mkdir()
    for k in range(2):
        _write_xyz(str(folder / ("c%d.xyz" % k)), g["coords_%d" % k], g["elements_%d" % k])
    return g, folder


@pytest.mark.parametrize("aff_num_qs", [1, 3])
def test_generate_voxel_grid_low_mem(golden, tmp_path, aff_num_qs):
    g, folder = _twostep_files(golden, tmp_path)
    args = (str(folder / "c0.xyz"), float(g["r"]), float(g["q"]), float(g["max_q"]), aff_num_qs, float(g["energy"]), "gen")
    kw = dict(fill_bkg=bool(g["fill_bkg"]), smooth=int(g["smooth"]))
    iq, qx, qy, qz = voxelgrids.generate_voxel_grid_low_mem(*args, **kw)
    ref = g["iq_full_aff%d" % aff_num_qs]
    assert iq.dtype == np.float64 and iq.shape == ref.shape
    assert np.array_equal(qx, g["axis"]) and np.array_equal(qy, g["axis"]) and np.array_equal(qz, g["axis"])
    assert np.array_equal(iq == 0, ref == 0)                    # never-hit voxels are exactly zero in both
    assert np.abs(iq - ref).max() <= TOL_INT * ref.max()
    # output_dir: the .npy hand-off files instead of a return value (voxelgrids.py:712-720)
    out_dir = tmp_path / "out"
    out_dir.mkdir()
    assert voxelgrids.generate_voxel_grid_low_mem(*args, output_dir=str(out_dir), **kw) is None
    on_disk = np.load(str(out_dir / "gen_output_files" / "gen_iq.npy"))
    # a second run: fp32 atomic accumulation order differs, so equal only to fp32 round-off
    assert on_disk.dtype == np.float64 and np.abs(on_disk - iq).max() <= 1e-5 * iq.max()
    assert np.array_equal(np.load(str(out_dir / "gen_output_files" / "gen_qz.npy")), g["axis"])
    with pytest.raises(Exception, match="must be a .pdb or .xyz"):
        voxelgrids.generate_voxel_grid_low_mem(str(folder / "c0.cif"), *args[1:], **kw)
    with pytest.raises(Exception, match="Invalid aff_num_qs"):
        voxelgrids.generate_voxel_grid_low_mem(*args[:4], 0, *args[5:], **kw)


def test_shell_mask_bit_exact_against_numpy():
    """gx_voxel_shell_scale picks exactly the voxels of (qr <= upper) & (qr > lower) with NumPy's qr."""
    rng = np.random.default_rng(5)
    for V, lo_hi in [(43, (0.7136, 1.4272)), (31, (0.0, 0.9)), (57, (1.1, 1.1000001))]:
        axis = np.linspace(-2.1408, 2.1408, V)
        mx, my, mz = np.meshgrid(axis, axis, axis)
        qr = np.sqrt(mx ** 2 + my ** 2 + mz ** 2)
        # put the bounds ON values qr takes, so <= / > decide on the last bit
        lower, upper = np.sort(rng.choice(qr.reshape(-1), 2)) if lo_hi[0] == 0.0 else lo_hi
        mask = (qr <= upper) & (qr > lower)
        iq = torch.ones(V, V, V, dtype=torch.float32, device=engine.resolve_device())
        engine.scale_shell(iq, axis, lower, upper, 2.0, iq.device)
        assert np.array_equal(iq.cpu().numpy() == 2.0, mask)


def test_two_step_drivers_npy_handoff(golden, tmp_path):
    """voxelgridmaker (folder of structures -> averaged, cropped, f0-weighted .npy) then detectormaker
    (.npy -> detector image) against the reference fixture / the oracle (old_modules/*.py)."""
    from giwaxsim_b200 import detectormaker, voxelgridmaker
    g, folder = _twostep_files(golden, tmp_path)
    out_dir = tmp_path / "work"
    out_dir.mkdir()
    cfg = {"input_folder": str(folder), "filetype": "xyz", "gen_name": "mix", "r_voxel_size": str(float(g["r"])),
           "q_voxel_size": str(float(g["q"])), "aff_num_qs": "1", "energy": str(float(g["energy"])),
           "max_q": str(float(g["max_q"])), "output_dir": str(out_dir), "smooth": str(int(g["smooth"])),
           "fill_bkg": "True"}
    iq, qx, qy, qz = voxelgridmaker.main(cfg)
    files = out_dir / "mix_output_files"
    assert np.array_equal(np.load(str(files / "mix_iq.npy")), iq) and (files / "mix_config.txt").exists()
    assert np.array_equal(np.load(str(files / "mix_qx.npy")), g["two_step_axis"])
    ref = g["two_step_iq"]
    assert iq.shape == ref.shape and np.abs(iq - ref).max() <= TOL_INT * ref.max()

    # second step from the REFERENCE's grid on disk (either producer's files must work)
    np.save(str(files / "mix_iq.npy"), ref)
    dcfg = {"iq_output_folder": str(files), "gen_name": "mix", "max_q": "1.5", "num_pixels": "65",
            "angle_init_val1": "90", "angle_init_val2": "90", "angle_init_val3": "90", "angle_init_ax1": "psi",
            "angle_init_ax2": "phi", "angle_init_ax3": "psi", "psi_start": "60", "psi_end": "90", "psi_num": "4",
            "phi_start": "0", "phi_end": "150", "phi_num": "5", "theta_start": "0", "theta_end": "3",
            "theta_num": "2", "mirror": "True"}
    det_sum, det_h, det_v = detectormaker.main(dcfg)
    sub = files / "mix_det_sum"
    assert np.array_equal(np.load(str(sub / "mix_det_sum.npy")), det_sum) and (sub / "mix_config.txt").exists()
    assert np.array_equal(np.load(str(sub / "mix_det_h.npy")), det_h)
    psis, phis, thetas = np.linspace(60, 90, 4), np.linspace(0, 150, 5), np.linspace(0, 3, 2)
    ones = lambda a: np.ones_like(a) / len(a)
    acc, o_h, o_v = ox.detectormaker(ref, qx, qy, qz, 65, 1.5, (90.0, 90.0, 90.0), ("psi", "phi", "psi"),
                                     psis, ones(psis), phis, ones(phis), thetas, ones(thetas), raw=True)
    want = ox.detector_epilogue_two_step(acc, mirror=True)
    assert np.array_equal(det_h, o_h) and np.abs(det_sum - want).max() <= TOL_INT * want.max()
    # a second run must not overwrite the first result folder (old_modules/detectormaker.py:60-65)
    detectormaker.main(dict(dcfg, mirror="False"))
    assert (files / "mix_det_sum1" / "mix_det_sum.npy").exists()
    with pytest.raises(Exception, match="Path does not exist"):
        detectormaker.main(dict(dcfg, iq_output_folder=str(tmp_path / "nowhere")))
