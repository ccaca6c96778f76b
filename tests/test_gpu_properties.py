"""Size-independent properties and internal cross-checks of the CUDA path."""
import numpy as np
import pytest
import torch

from giwaxsim_b200 import engine, synth
from giwaxsim_b200._lib import call, ptr
from giwaxsim_b200.tools import comparison
from oracle import giwaxs_oracle as ox

pytestmark = pytest.mark.gpu


def _yrange(atoms, xs, ys, n_atoms, phis):
    dev = atoms.device
    rad = np.radians(phis)
    d_sn, d_cs = engine._dev(np.sin(rad), dev), engine._dev(np.cos(rad), dev)
    out = torch.empty(2 * len(phis), dtype=torch.float64, device=dev)
    call("gx_slice_yrange", ptr(xs), ptr(ys), n_atoms, ptr(d_sn), ptr(d_cs), len(phis), ptr(out), None)
    torch.cuda.synchronize()
    return out.cpu().numpy().reshape(-1, 2)


@pytest.mark.parametrize("kind", ["random", "lattice", "collinear", "tiny"])
def test_hull_candidates_give_identical_y_range(kind):
    """min/max of y' over the hull candidates == over all atoms, bit for bit, and
    both equal the oracle's fma-chain values."""
    rng = np.random.default_rng(3)
    if kind == "random":
        coords = rng.random((200_000, 3)) * [300.0, 120.0, 200.0]
    elif kind == "lattice":          # many atoms exactly on the boundary lines
        g = np.stack(np.meshgrid(np.arange(60) * 2.456, np.arange(40) * 4.254, np.arange(20) * 3.348,
                                 indexing="ij"), -1).reshape(-1, 3)
        coords = g
    elif kind == "collinear":
        t = rng.random(5000)
        coords = np.stack([t * 50, t * 20 + 3.0, rng.random(5000) * 30], 1)
    else:
        coords = rng.random((5, 3)) * 10
    dev = engine.resolve_device()
    atoms = engine.AtomSet(coords, 0.5, 1024, dev, species=np.zeros(len(coords), np.uint8), table=[6 + 0j])
    phis = np.linspace(0, 179.9, 371)
    cx, cy, cn = atoms.candidates()
    full = _yrange(atoms, atoms.xs, atoms.ys, atoms.A, phis)
    cand = _yrange(atoms, cx, cy, cn, phis)
    assert np.array_equal(full, cand)
    if kind in ("random", "lattice"):
        assert cn < atoms.A // 10, "candidate reduction did not reduce (%d of %d)" % (cn, atoms.A)
    for k in (0, 17, 185, 370):
        yr = ox.rotz_y(coords, phis[k])
        assert full[k, 0] == yr.min() and full[k, 1] == yr.max()


@pytest.mark.parametrize("inits,axs", [((90.0, 90.0, 90.0), ("psi", "phi", "psi")),
                                        ((0.0, 0.0, 0.0), ("None", "None", "None")),
                                        ((12.5, 40.0, 3.0), ("theta", "phi", "psi"))])
def test_filtered_detector_kernel_is_bit_identical_to_exact(inits, axs):
    """fp32-filtered gather == all-fp64 gather: same voxel index for every pixel of every
    orientation (so the images are bitwise equal), including grid-aligned orientations
    where every pixel sits on a voxel edge and pixels far outside the voxel box."""
    rng = np.random.default_rng(1)
    V = 101
    q = np.linspace(-2.0, 2.0, V)                     # dq = 0.04: detector pixels land on voxel edges
    iq = rng.random((V, V, V)).astype(np.float32)
    dev = engine.resolve_device()
    P = 320
    gx, gy, gz, _, _ = comparison.detector_base_device(P, 2.4, inits, axs, dev)    # extends beyond the box
    psis = np.array([0.0, 17.3, 45.0, 89.75, 90.0])
    phis = np.array([0.0, 33.3, 90.0, 179.0])
    thetas = np.array([0.0, 1.0])
    ones = lambda a: np.ones_like(a) / len(a)
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, ones(psis), phis, ones(phis),
                                     thetas, ones(thetas))
    det = engine.DetectorEngine(iq, q, q, q)
    img_exact, _ = det.accumulate(gx, gy, gz, R, w, exact_only=True)
    img_fast, _ = det.accumulate(gx, gy, gz, R, w, exact_only=False, count_slow=True)
    assert torch.equal(img_exact, img_fast)
    assert 0.0 <= det.last_slow_fraction <= 1.0
    for o in (0, 7, 19, 39):
        _, a = det.accumulate(gx, gy, gz, R, w, probe=o, exact_only=True)
        _, b = det.accumulate(gx, gy, gz, R, w, probe=o, exact_only=False)
        assert torch.equal(a, b)


def test_filtered_detector_kernel_rarely_falls_back_on_generic_orientations():
    rng = np.random.default_rng(2)
    V = 403
    q = np.linspace(-2.01, 2.01, V)
    iq = rng.random((V, V, V)).astype(np.float32)
    dev = engine.resolve_device()
    gx, gy, gz, _, _ = comparison.detector_base_device(512, 2.0, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), dev)
    psis = np.linspace(3.1, 88.3, 24)
    ones = np.ones(1)
    det = engine.DetectorEngine(iq, q, q, q)
    # tilted planes (phi = 7.3 deg): only pixels within the fp32 bound of a voxel edge fall back
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones(24) / 24, [7.3], ones, [0.4], ones)
    a, _ = det.accumulate(gx, gy, gz, R, w, exact_only=True)
    b, _ = det.accumulate(gx, gy, gz, R, w, exact_only=False, count_slow=True)
    assert torch.equal(a, b)
    assert det.last_slow_fraction < 0.02, det.last_slow_fraction
    # phi = theta = 0: the plane lies in q_z = 0, i.e. exactly on a voxel edge, for every pixel
    # (rounding noise picks the bin) -> every pixel needs the exact z index; still identical
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones(24) / 24, [0.0], ones, [0.0], ones)
    a, _ = det.accumulate(gx, gy, gz, R, w, exact_only=True)
    b, _ = det.accumulate(gx, gy, gz, R, w, exact_only=False, count_slow=True)
    assert torch.equal(a, b)
    assert det.last_slow_fraction > 0.9


def test_linearity_of_detector_accumulation():
    """image(w1) + image(w2) == image(w1 + w2): the gather is linear in the weights."""
    rng = np.random.default_rng(0)
    V = 41
    iq = rng.random((V, V, V)).astype(np.float32)
    q = np.linspace(-2.05, 2.05, V)
    dev = engine.resolve_device()
    gx, gy, gz, _, _ = comparison.detector_base_device(96, 2.0, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), dev)
    psis = np.linspace(60, 90, 9)
    ones = np.ones(1)
    det = engine.DetectorEngine(iq, q, q, q)
    w1, w2 = rng.random(9), rng.random(9)
    R, _ = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, w1, [0.0], ones, [0.0], ones)
    a, _ = det.accumulate(gx, gy, gz, R, w1)
    b, _ = det.accumulate(gx, gy, gz, R, w2)
    c, _ = det.accumulate(gx, gy, gz, R, w1 + w2)
    assert torch.allclose(a + b, c, rtol=1e-12, atol=0)


def test_counts_are_sum_over_slices_and_fused_equals_staged():
    """Counts are integers that add over phi shards (what the multi-GPU reduce relies on),
    and the fused kernels agree with the staged ones."""
    coords, el = synth.random_slab(30_000, (60.0, 35.0, 50.0), seed=4)
    r, max_q = 0.25, 1.5
    q = synth.pow2_q_voxel(r, 256)
    codes, uniq, table = comparison.species_table(el, 12700.0)
    dev = engine.resolve_device()
    atoms = engine.AtomSet(coords, r, 256, dev, species=codes, table=table)
    N, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
    avg = np.sum(np.bincount(codes, minlength=len(table)) * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3

    def run(sel, staged):
        e = engine.SliceEngine(None, r, q_axis, N, avg, atoms.bounds[0], atoms.bounds[1], True, 6, atoms=atoms)
        e.run(phis[sel], staged=staged)
        return e.counts(), e.sums().astype(np.float64)

    c_all, s_all = run(slice(None), False)
    c0, s0 = run(slice(0, None, 2), False)
    c1, s1 = run(slice(1, None, 2), False)
    assert np.array_equal(c0 + c1, c_all)
    assert np.abs(s0 + s1 - s_all).max() <= 1e-5 * s_all.max()
    c_st, s_st = run(slice(None), True)
    assert np.array_equal(c_st, c_all)
    assert np.abs(s_st - s_all).max() <= 1e-5 * s_all.max()
    # every kept sample is counted exactly once: total count = sum over slices of kept rows x kept cols
    assert c_all.sum() > 0
