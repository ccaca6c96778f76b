"""SURVEY 8(f) N4, second half: the post-hoc transforms evaluate_fit applies to the finished detector image
(tools/comparison.py:161-191, 469-592, 873-912) on the device, against the oracle restatement (which is
pinned bit for bit against the reference functions in tests/test_oracle_vs_reference.py).  fp64 throughout;
the device evaluates sin / cos / atan2 itself, so sample coordinates differ from NumPy's in the last ulp:
tolerance 1e-9 of the image maximum."""
import numpy as np
import pytest

from giwaxsim_b200 import synth
from giwaxsim_b200.tools import comparison
from oracle import giwaxs_oracle as ox

pytestmark = pytest.mark.gpu
TOL = 1e-9


def _case(seed=4, P=90):
    rng = np.random.default_rng(seed)
    h = np.linspace(-2.0, 2.0, P)
    v = np.linspace(-2.0, 2.0, P)
    xx, yy = np.meshgrid(h, v)
    img = np.exp(-((np.hypot(xx, yy) - 1.1) / 0.15) ** 2) * (1 + 0.3 * np.cos(3 * np.arctan2(yy, xx))) + 0.05 * rng.random((P, P))
    return img, h, v, np.linspace(0.0, 1.8, 37), np.linspace(0.0, 1.7, 35)


def test_trim_and_polar_warps_against_oracle():
    img, h, v, exp_qxy, exp_qz = _case()
    a = comparison.trim_sim_data(img, h, v, exp_qxy, exp_qz)
    b = ox.trim_sim_data(img, h, v, exp_qxy, exp_qz)
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    trim, th, tv = b
    centre = (int(np.argmin(np.abs(tv))), int(np.argmin(np.abs(th))))
    radius = float(np.sqrt(trim.shape[0] ** 2 + trim.shape[1] ** 2))
    pol = comparison.linear_polar(trim, o=centre, r=radius, output=None, order=1, cont=0)
    want = ox.linear_polar(trim, centre, radius)
    assert pol.shape == want.shape and np.abs(pol - want).max() <= TOL * np.abs(want).max()
    # defaults: origin at the image centre, radius = half diagonal, cval respected outside the image
    pol2 = comparison.linear_polar(img, cont=-3.0)
    o2 = np.array(img.shape) / 2 - 0.5
    r2 = np.sqrt((np.array(img.shape) ** 2).sum()) / 2
    want2 = ox.linear_polar(img, o2, r2, cont=-3.0)
    assert np.abs(pol2 - want2).max() <= TOL * np.abs(want2).max() and (pol2 == -3.0).any()
    back = comparison.polar_linear(want, o=centre, r=None, output=trim.shape)
    want_back = ox.polar_linear(want, centre, trim.shape)
    assert np.abs(back - want_back).max() <= TOL * np.abs(want_back).max()
    with pytest.raises(ValueError):
        comparison.linear_polar(img, order=3)


@pytest.mark.parametrize("pad_width,pad_range", [(0.05, (0.9, 1.4)), (0.0, (0.9, 1.4)), (0.12, (0.3, 1.6))])
def test_shift_peak_against_oracle(pad_width, pad_range):
    img, h, v, exp_qxy, exp_qz = _case(seed=6)
    trim, th, tv = ox.trim_sim_data(img, h, v, exp_qxy, exp_qz)
    got = comparison.shift_peak(trim.copy(), th, tv, pad_width, pad_range)
    want = ox.shift_peak(trim.copy(), th, tv, pad_width, pad_range)
    assert got.shape == want.shape and np.abs(got - want).max() <= TOL * np.abs(want).max()
    with pytest.raises(AssertionError, match="pad_range is too small"):
        comparison.shift_peak(trim.copy(), th, tv, 0.6, (1.0, 1.05))


def test_scale_offset_fit_against_oracle():
    rng = np.random.default_rng(2)
    sim = rng.random((70, 64)) * 40
    target = 2.75 * sim - 11.0 + rng.normal(size=sim.shape)
    mask = (rng.random(sim.shape) < 0.3).astype(int)
    s0, o0 = ox.optimize_scale_offset(sim, target, mask)
    s1, o1 = comparison.optimize_scale_offset(sim, target, mask)
    assert abs(s1 - s0) <= 1e-10 * abs(s0) and abs(o1 - o0) <= 1e-9 * abs(o0)


def test_evaluate_fit_end_to_end(tmp_path):
    """evaluate_fit (comparison.py:884-912): slab -> voxel grid -> detector -> trim -> shift_peak -> scale/offset
    through the drop-in against the oracle pipeline (intensities 1e-4 of the maximum, as everywhere)."""
    from giwaxsim_b200.tools import utilities
    rng = np.random.default_rng(12)
    n = 30
    xyz = rng.random((n, 3)) * np.array([6.1, 7.3, 5.2])
    el = rng.choice(np.array(["C", "H", "S"]), size=n)
    path = str(tmp_path / "cell.xyz")
    with open(path, "w") as fh:
        fh.write("%d\ncell\n" % n)
        for e, p in zip(el, xyz):
            fh.write("%s %.6f %.6f %.6f\n" % (e, p[0], p[1], p[2]))
    cell = (6.1, 7.3, 5.2, 90.0, 90.0, 90.0)
    sizes = (30.0, 36.0, 26.0)
    r, q, max_q, P = 0.3, 0.1, 1.5, 80
    psis, phis, thetas = np.linspace(75, 90, 4), np.linspace(0, 150, 4), np.array([0.0])
    vals, axs = (90.0, 90.0, 90.0), ("psi", "phi", "psi")
    exp_qxy, exp_qz = np.linspace(0.0, 1.4, 30), np.linspace(0.0, 1.3, 28)
    # oracle pipeline
    c0, e0 = ox.read_structure(path)
    coords, elements = ox.slabmaker(c0, e0, *sizes, *cell)
    f = ox.f_values_for(elements, table=synth.fixed_f1f2)
    iq, qx, qy, qz, *_ = ox.voxelgridmaker(coords, f, r, q, max_q, True, 3)
    ones = lambda a: np.ones_like(a) / len(a)
    det, dh, dv = ox.detectormaker(iq, qx, qy, qz, P, max_q, vals, axs, psis, ones(psis), phis, ones(phis), thetas, ones(thetas))
    trim, th, tv = ox.trim_sim_data(det, dh, dv, exp_qxy, exp_qz)
    rebin_map = 1.7e3 * ox.shift_peak(trim.copy(), th, tv, 0.04, (0.7, 1.2)) + 5e-9 + 1e-9 * rng.random(trim.shape)
    rebin_mask = (rng.random(trim.shape) < 0.15).astype(int)
    want_comp, want_diff, scale, offset = ox.compare_maps(det, dh, dv, rebin_map, rebin_mask, exp_qxy, exp_qz, 0.04, (0.7, 1.2))
    got_ref, got_comp, got_diff = comparison.evaluate_fit(
        sizes, (path,) + cell, (r, q, max_q, 12700.0, True, 3), (P, vals, axs, psis, None, phis, None, thetas, None),
        (rebin_map, rebin_mask, exp_qxy, exp_qz, 0.04, (0.7, 1.2)))
    assert got_ref is rebin_map or np.array_equal(got_ref, rebin_map)
    top = np.abs(want_comp).max()
    assert np.abs(got_comp - want_comp).max() <= 1e-4 * top
    assert np.abs(got_diff - want_diff).max() <= 1e-4 * top
    assert np.array_equal(got_comp == 0, want_comp == 0)
