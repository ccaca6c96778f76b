"""Live pin of the oracle against the unmodified reference (only where the
read-only checkout is mounted, i.e. in the build container)."""
import numpy as np
import pytest

from oracle import ref_shim, giwaxs_oracle as ox

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")


def _cluster(scale=(1.0, 1.0, 1.0), seed=3, n=300):
    rng = np.random.default_rng(seed)
    coords = rng.random((n, 3)) * np.array([14.0, 21.0, 11.0]) * np.array(scale)
    elements = rng.choice(np.array(["C", "H", "S", "O"]), size=n)
    return coords, elements


@pytest.mark.parametrize("fill_bkg,smooth,scale", [(True, 4, (1, 1, 1)), (False, 0, (1, 1, 1)),
                                                    (False, 2, (1, 1, 1)), (True, 3, (3.5, 2.0, 1.0))])
def test_stage_a_bit_exact(fill_bkg, smooth, scale):
    coords, elements = _cluster(scale)
    r, q, max_q = 0.3, 0.11, 1.5
    setup = ref_shim.stage_a_setup(coords, elements, r, q, max_q, 12700.0)
    cap = {}
    vsum, vcnt = ref_shim.run_slices_serial(coords, setup, r, fill_bkg, smooth, capture=cap)
    ref_out = ref_shim.finalize_serial(vsum, vcnt, setup["q_axis"], max_q)
    f = ox.f_values_for(elements)
    iq, qx, qy, qz, osum, ocnt, osetup = ox.voxelgridmaker(coords, f, r, q, max_q, fill_bkg, smooth)
    assert np.array_equal(vsum, osum) and np.array_equal(vcnt, ocnt)
    assert np.array_equal(iq, ref_out[0]) and np.array_equal(qx, ref_out[1])
    q3 = (osetup["q_num"],) * 3
    for i in range(0, len(osetup["phis"]), 17):
        out = {}
        ox.run_slice(np.zeros(q3), np.zeros(q3), coords, osetup, r, osetup["phis"][i], fill_bkg, smooth, out=out)
        assert np.array_equal(out["grid"], cap["grid"][i])
        assert np.array_equal(out["iq_2d"], cap["iq_2d"][i])
        assert np.array_equal(out["det_h_qx"], cap["det_h_qx"][i])
        assert np.array_equal(out["det_h_qy"], cap["det_h_qy"][i])


def test_chord_lengths_bit_exact():
    ref = ref_shim.load()
    x = np.arange(300) * 0.3
    for phi in [0.0, 0.5, 12.25, 45.0, 89.9, 90.0, 90.3, 133.0, 179.8]:
        for hor, ver in [(21.0, 14.0), (14.0, 21.0), (40.0, 3.0)]:
            a = ref.voxelgrids.rectangular_collapse_lengths(x, hor, ver, np.float64(phi))
            b = ox.chord_lengths(x, hor, ver, np.float64(phi))
            assert np.array_equal(np.asarray(a, dtype=np.float64), b), (phi, hor, ver)


@pytest.mark.parametrize("P,vals,axs,mirror", [(40, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), True),
                                               (41, (5.0, 0.0, 33.0), ("phi", "None", "theta"), True)])
def test_stage_b_bit_exact(P, vals, axs, mirror):
    rng = np.random.default_rng(11)
    V = 31
    iq = rng.random((V, V, V)) * 1e6
    q = np.linspace(-2.1, 2.1, V)
    psis, phis, thetas = np.linspace(70, 90, 3), np.linspace(0, 170, 4), np.linspace(0, 2, 2)
    w = [np.ones_like(a) / len(a) for a in (psis, phis, thetas)]
    a = ref_shim.detectormaker_serial(iq, q, q, q, P, 2.0, vals, axs, psis, w[0], phis, w[1], thetas, w[2], mirror=mirror)
    b = ox.detectormaker(iq, q, q, q, P, 2.0, vals, axs, psis, w[0], phis, w[1], thetas, w[2], mirror=mirror)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))


@pytest.mark.parametrize("name,size,cell", [
    ("PM6_sample.pdb", (60.0, 50.0, 45.0), (44.456, 45.726, 40.097, 90.0, 90.0, 90.0)),
    ("graphite_medium.xyz", (20.0, 30.0, 25.0), (14.0, 21.0, 15.0, 90.0, 90.0, 90.0)),
    ("graphite_medium.xyz", (18.0, 22.0, 16.0), (14.0, 21.0, 15.0, 80.0, 95.0, 110.0)),     # triclinic cell
])
def test_slabmaker_bit_exact(name, size, cell):
    import os
    ref = ref_shim.load()
    path = os.path.join(ref_shim.REFERENCE_ROOT, "test_input_files", name)
    r_coords, r_el = ref.comparison.slabmaker_fitting(path, *size, *cell)
    c0, e0 = ox.read_structure(path)
    loader = ref.utilities.load_pdb if name.endswith(".pdb") else ref.utilities.load_xyz
    rc0, re0 = loader(path)
    assert np.array_equal(c0, rc0) and np.array_equal(e0, re0)
    o_coords, o_el = ox.slabmaker(c0, e0, *size, *cell)
    assert np.array_equal(o_coords, r_coords) and np.array_equal(o_el, r_el)


# ---- two-step command line and the aff_num_qs > 1 branch ----
def _write_xyz(path, coords, elements):
    with open(path, "w") as fh:
        fh.write("%d\ncluster\n" % len(coords))
        for el, (x, y, z) in zip(elements, coords):
            fh.write("%s %.17g %.17g %.17g\n" % (el, x, y, z))


@pytest.mark.parametrize("aff_num_qs,fill_bkg,smooth", [(1, True, 3), (3, False, 0), (2, True, 2)])
def test_voxel_grid_low_mem_bit_exact(tmp_path, aff_num_qs, fill_bkg, smooth):
    coords, elements = _cluster(n=120)
    path = str(tmp_path / "cluster.xyz")
    _write_xyz(path, coords, elements)
    r, q, max_q = 0.3, 0.16, 1.5
    a = ref_shim.voxel_grid_low_mem_serial(path, r, q, max_q, aff_num_qs, 12700.0, fill_bkg, smooth)
    c, e = ox.read_structure(path)
    b = ox.voxel_grid_low_mem(c, e, r, q, max_q, aff_num_qs, 12700.0, fill_bkg, smooth)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    if aff_num_qs > 1:
        # what the reference returns is the LAST shell's grid doubled inside that shell: earlier
        # passes are overwritten (voxelgrids.py:699), so running only the last pass is equivalent
        max_q_diag = np.sqrt(2) * max_q
        max_q_diag = max_q_diag + max_q_diag % q
        step = max_q_diag / aff_num_qs
        q_val = 0.5 * step + (aff_num_qs - 1) * step
        f0 = ox.f0_values(q_val, e)
        f = np.array([f0[x] for x in e]) + np.array([complex(*ox.ftable.f1_f2(x)) for x in e])
        last, axis, _ = ox._whole_grid(c, f, r, q, max_q, fill_bkg, smooth)
        mx, my, mz = np.meshgrid(axis, axis, axis)
        qr = np.sqrt(mx ** 2 + my ** 2 + mz ** 2)
        shell = (qr <= q_val + step / 2) & (qr > q_val - step / 2)
        assert np.array_equal(a[0], np.where(shell, 2 * last, last))


def test_f0_tables_match_reference():
    ref = ref_shim.load()
    from giwaxsim_b200.tools import utilities as ours
    aff = ref.utilities.aff_dict
    for el, coeff in ox.AFF.items():
        assert tuple(aff[el]) == coeff
    for el, coeff in ours.CROMER_MANN.items():
        assert tuple(aff[el]) == coeff and ours.ATOMIC_NUMBER[el] == ref.utilities.ptable[el]
    els = ["C", "H", "S", "O", "F", "N"]
    for q_val in (0.0, 0.37, 2.9):
        a = ref.utilities.get_element_f0_dict(q_val, els)
        assert a == ox.f0_values(q_val, els) == ours.get_element_f0_dict(q_val, els)


def test_most_common_element_and_two_step_average(tmp_path):
    ref = ref_shim.load()
    from giwaxsim_b200.tools import utilities as ours
    paths = []
    for k, seed in enumerate((3, 4)):
        coords, elements = _cluster(seed=seed, n=90)
        paths.append(str(tmp_path / ("c%d.xyz" % k)))
        _write_xyz(paths[-1], coords, elements)
    for p in paths:
        assert ref.utilities.most_common_element(p) == ours.most_common_element(p) \
            == ox.most_common_element(ox.read_structure(p)[1])
    # the reference's driver loop (old_modules/voxelgridmaker.py:33-68) on its own functions
    r, q, max_q = 0.3, 0.16, 1.5
    total = None
    for i, p in enumerate(paths):
        iq, qx, qy, qz = ref_shim.voxel_grid_low_mem_serial(p, r, q, max_q, 1, 12700.0, True, 2)
        small, qxs, qys, qzs = ref.voxelgrids.downselect_voxelgrid(iq, qx, qy, qz, max_q)
        total = small if i == 0 else total + small
    total = total / len(paths)
    want = ref.voxelgrids.add_f0_q_3d(total, qxs, qys, qzs, ref.utilities.most_common_element(paths[0]))
    got = ox.two_step_voxelgrid([ox.read_structure(p) for p in paths], r, q, max_q, 1, 12700.0, True, 2)
    assert np.array_equal(got[0], want) and np.array_equal(got[1], qxs)


@pytest.mark.parametrize("seed", range(6))
def test_randomised_pipeline_bit_exact(seed):
    """Random voxel sizes, q ranges, smoothing widths, detector sizes and init rotations: the oracle's whole
    pipeline (stage A grid, crop, f0 weight, stage B image) equals the unmodified reference's bit for bit."""
    rng = np.random.default_rng(100 + seed)
    n = int(rng.integers(40, 400))
    box = rng.uniform(6.0, 30.0, size=3)
    coords = rng.random((n, 3)) * box
    elements = rng.choice(np.array(["C", "H", "S", "O", "N", "F"]), size=n)
    r = float(rng.choice([0.25, 0.3, 0.41]))
    max_q = float(rng.choice([1.0, 1.5, 2.0]))
    q = float(rng.uniform(0.12, 0.2))
    fill_bkg, smooth = bool(rng.integers(0, 2)), int(rng.integers(0, 5))
    if 2 * np.pi / q < np.min(coords.max(0) - coords.min(0)):
        q = 2 * np.pi / (1.5 * np.max(box))
    a = ref_shim.voxelgridmaker_serial(coords, elements, r, q, max_q, 12700.0, fill_bkg, smooth)
    b = ox.voxelgridmaker(coords, ox.f_values_for(elements), r, q, max_q, fill_bkg, smooth)[:4]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    P = int(rng.integers(17, 40))
    vals = tuple(float(v) for v in rng.choice([0.0, 90.0, 33.0, 7.5], size=3))
    axs = tuple(str(v) for v in rng.choice(["psi", "phi", "theta", "None"], size=3))
    psis, phis, thetas = np.linspace(0, 90, 3), np.linspace(0, 175, 4), np.linspace(0, 5, 2)
    w = [np.ones_like(x) / len(x) for x in (psis, phis, thetas)]
    mirror = bool(rng.integers(0, 2))
    da = ref_shim.detectormaker_serial(a[0], a[1], a[2], a[3], P, max_q, vals, axs, psis, w[0], phis, w[1], thetas, w[2],
                                       mirror=mirror)
    db = ox.detectormaker(b[0], b[1], b[2], b[3], P, max_q, vals, axs, psis, w[0], phis, w[1], thetas, w[2], mirror=mirror)
    assert all(np.array_equal(x, y) for x, y in zip(da, db))


def _compare_case(seed=4, P=60):
    rng = np.random.default_rng(seed)
    h = np.linspace(-2.0, 2.0, P)
    v = np.linspace(-2.0, 2.0, P)
    xx, yy = np.meshgrid(h, v)
    img = np.exp(-((np.hypot(xx, yy) - 1.1) / 0.15) ** 2) * (1 + 0.3 * np.cos(3 * np.arctan2(yy, xx))) + 0.05 * rng.random((P, P))
    exp_qxy, exp_qz = np.linspace(0.0, 1.8, 37), np.linspace(0.0, 1.7, 35)
    return img, h, v, exp_qxy, exp_qz


def test_post_hoc_transforms_bit_exact():
    """trim_sim_data, linear_polar, polar_linear, add_pad, shift_peak, optimize_scale_offset and the tail of
    evaluate_fit (comparison.py:161-191, 469-592, 873-912) restated in the oracle == the reference functions."""
    ref = ref_shim.load().comparison
    img, h, v, exp_qxy, exp_qz = _compare_case()
    a = ref.trim_sim_data(img, h, v, exp_qxy, exp_qz)
    b = ox.trim_sim_data(img, h, v, exp_qxy, exp_qz)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    trim, th, tv = b
    centre = (int(np.argmin(np.abs(tv))), int(np.argmin(np.abs(th))))
    radius = float(np.sqrt(trim.shape[0] ** 2 + trim.shape[1] ** 2))
    pol_ref = ref.linear_polar(trim, o=centre, r=radius, output=None, order=1, cont=0)
    pol = ox.linear_polar(trim, centre, radius)
    assert np.array_equal(pol, pol_ref)
    assert np.array_equal(ox.polar_linear(pol, centre, trim.shape), ref.polar_linear(pol_ref, o=centre, r=None, output=trim.shape))
    for pad_width, pad_range in [(0.05, (0.9, 1.4)), (0.0, (0.9, 1.4)), (0.12, (0.3, 1.6))]:
        assert np.array_equal(ox.shift_peak(trim.copy(), th, tv, pad_width, pad_range),
                              ref.shift_peak(trim.copy(), th, tv, pad_width, pad_range))
    rng = np.random.default_rng(8)
    target = 3.5 * ox.shift_peak(trim.copy(), th, tv, 0.05, (0.9, 1.4)) + 0.7 + 0.01 * rng.random(trim.shape)
    mask = (rng.random(trim.shape) < 0.2).astype(int)
    s0, o0 = ref.optimize_scale_offset(trim, target, mask)
    s1, o1 = ox.optimize_scale_offset(trim, target, mask)
    assert (s0, o0) == (s1, o1)
