"""BASELINE configs[4] sizes (N = 4096 real-space grid, q_num = 569, 2048^2 detector) through
size-independent properties: at these sizes the CPU oracle needs seconds per slice, so the
checks are internal identities of the path plus the oracle on a single slice / a few pixels."""
import numpy as np
import pytest
import torch

from giwaxsim_b200 import engine, synth
from giwaxsim_b200.tools import comparison
from oracle import giwaxs_oracle as ox

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def slab5():
    cfg = synth.config5()
    coords, el = synth.random_slab(1_500_000, cfg["box"], seed=11)       # config-5 box, 15 % of its atoms
    dev = engine.resolve_device()
    codes, uniq, counts = engine.encode_elements_device(el, dev)
    table = comparison.f_table(uniq, cfg["energy"])
    r, q, max_q = cfg["r_voxel_size"], cfg["q_voxel_size"], cfg["max_q"]
    atoms = engine.AtomSet(coords, r, cfg["grid_size"], dev, species=codes, table=table)
    N, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
    assert (N, q_num) == (4096, 569)
    avg = np.sum(counts * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3
    return dict(cfg=cfg, coords=coords, el=el, atoms=atoms, table=table, uniq=uniq, q_axis=q_axis, phis=phis,
                avg=avg, r=r, q=q, max_q=max_q, N=N)


def _engine(s, window=None, count3d=False):
    a = s["atoms"]
    return engine.SliceEngine(None, s["r"], s["q_axis"], s["N"], s["avg"], a.bounds[0], a.bounds[1], True, 25,
                              atoms=a, window=window, count3d=count3d)


def test_full_size_fused_equals_staged_and_window_equals_crop(slab5):
    s = slab5
    sel = s["phis"][[0, 1, 400, 891, 1337, 1782]]
    lo, hi = engine.crop_range(s["q_axis"], s["max_q"])
    full = _engine(s)
    full.run(sel)                                   # fused, full q_num^3 accumulators
    win = _engine(s, window=(lo, hi))
    win.run(sel)                                    # fused, crop window only
    staged = _engine(s)
    staged.run(sel, staged=True)                    # project -> fft2 -> bin kernels
    c_full, s_full = full.counts(), full.sums().astype(np.float64)
    assert np.array_equal(win.counts(), c_full[lo:hi, lo:hi, lo:hi])                  # bit-exact
    top = s_full.max()
    assert np.abs(win.sums() - s_full[lo:hi, lo:hi, lo:hi]).max() <= 2e-6 * top
    assert np.array_equal(staged.counts(), c_full)
    assert np.abs(staged.sums() - s_full).max() <= 1e-5 * top
    # every kept (row, col) sample of every slice is counted exactly once
    kept_rows = int((full.row_index >= 0).sum())
    t = full.prepare(sel)
    kept_cols = int((t["col"] >= 0).sum())
    assert int(c_full.sum()) == kept_rows * kept_cols


def test_full_size_single_slice_against_oracle(slab5):
    """One N = 4096 slice of the 1.5 M-atom slab: atom pixel indices and voxel counts bit-exact,
    accumulated intensities within 1e-4 of the maximum."""
    s = slab5
    phi = float(s["phis"][777])
    e = _engine(s)
    y_idx, z_idx, bbox = e.atom_indices(phi)
    oy, oz, _valid = ox.atom_pixel_indices(s["coords"], phi, s["N"], s["r"])[:3]
    assert np.array_equal(y_idx, oy) and np.array_equal(z_idx, oz)
    e.run(np.array([phi]))
    f = ox.f_values_for(s["el"], table=synth.fixed_f1f2)
    setup = ox.stage_a_setup(s["coords"], f, s["r"], s["q"], s["max_q"])
    q3 = (setup["q_num"],) * 3
    vsum, vcnt = np.zeros(q3), np.zeros(q3)
    ox.run_slice(vsum, vcnt, s["coords"], setup, s["r"], phi, True, 25)
    assert np.array_equal(e.counts(), vcnt.astype(np.int64))
    assert np.abs(e.sums() - vsum).max() <= 1e-4 * vsum.max()


def test_full_size_detector_kernels_agree(slab5):
    """2048^2 detector, bench geometry: the fixed-point kernel's voxel index equals the all-fp64
    kernel's for every pixel of the probed orientations, and the oracle's on a sample of pixels."""
    cfg = slab5["cfg"]
    rng = np.random.default_rng(3)
    V = 403
    q = np.linspace(-2.01, 2.01, V)
    iq = rng.random((V, V, V)).astype(np.float32)
    dev = engine.resolve_device()
    P = 2048
    gx, gy, gz, _, _ = comparison.detector_base_device(P, 2.0, cfg["angle_init_vals"], cfg["angle_init_axs"], dev)
    psis = np.linspace(0, 89.75, 360)
    for phis, thetas in (([0.0], [0.0]), ([11.0], [0.7])):
        R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones(360) / 360, phis, np.ones(1),
                                         thetas, np.ones(1))
        det = engine.DetectorEngine(iq, q, q, q)
        a, _ = det.accumulate(gx, gy, gz, R, w, kernel="exact")
        b, _ = det.accumulate(gx, gy, gz, R, w, kernel="affine", count_slow=True)
        assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max())
        assert det.last_slow_fraction < 1e-3
        for o in (0, 1, 47, 48, 179, 359):                  # both sides of the 48-orientation first chunk
            _, ia = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="exact")
            _, ib = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="affine")
            assert torch.equal(ia, ib), o
    # oracle on a strided sample of pixels of one tilted orientation
    hx, hy, hz, _, _ = ox.detector_base(P, 2.0, cfg["angle_init_vals"], cfg["angle_init_axs"])
    g = ox.rotate_psi_phi_theta(hx, hy, hz, psis[179], 11.0, 0.7)
    ix, iy, iz = ox.detector_voxel_indices((V, V, V), q, q, q, *g)
    flat = ((iy * V + ix) * V + iz).reshape(P, P)
    _, ib = det.accumulate(gx, gy, gz, R, w, probe=179, kernel="affine")
    assert np.array_equal(ib.cpu().numpy().reshape(P, P)[::7, ::5], flat[::7, ::5])


def test_config2_grid_size_2095_bluestein_slice_against_oracle():
    """BASELINE configs[1] geometry (r = 0.3, q = 0.01 -> N = 2095 = 5 x 419, q_num = 569): the largest
    Bluestein transform (8192-point convolution) through the fused kernels, one slice against the oracle
    and fused == staged on three."""
    rng = np.random.default_rng(21)
    coords = rng.random((60_000, 3)) * [210.0, 340.0, 160.0]
    el = rng.choice(np.array(["S", "O", "H"]), size=len(coords))
    r, q, max_q = 0.3, 0.01, 2.0
    dev = engine.resolve_device()
    codes, uniq, counts = engine.encode_elements_device(el, dev)
    table = comparison.f_table(uniq, 12700.0)
    atoms = engine.AtomSet(coords, r, 2095, dev, species=codes, table=table)
    N, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
    assert (N, q_num) == (2095, 569)
    avg = np.sum(counts * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3
    mk = lambda **kw: engine.SliceEngine(None, r, q_axis, N, avg, atoms.bounds[0], atoms.bounds[1], False, 7,
                                         atoms=atoms, **kw)
    sel = phis[[3, 600, 1500]]
    fused, staged = mk(), mk()
    fused.run(sel)
    staged.run(sel, staged=True)
    assert np.array_equal(fused.counts(), staged.counts())
    assert np.abs(fused.sums() - staged.sums()).max() <= 1e-5 * staged.sums().max()
    one = mk()
    one.run(sel[1:2])
    f = ox.f_values_for(el, table=synth.fixed_f1f2)
    setup = ox.stage_a_setup(coords, f, r, q, max_q)
    q3 = (setup["q_num"],) * 3
    vsum, vcnt = np.zeros(q3), np.zeros(q3)
    ox.run_slice(vsum, vcnt, coords, setup, r, float(sel[1]), False, 7)
    assert np.array_equal(one.counts(), vcnt.astype(np.int64))
    assert np.abs(one.sums() - vsum).max() <= 1e-4 * vsum.max()


def _graphite_crystal(nx, ny, nz):
    """AB-stacked graphite (a = 2.456 A, c = 6.696 A) as an orthorhombic 8-atom cell replicated
    nx x ny x nz times: coordinates on a lattice, like test_input_files/graphite_large.xyz, so that
    whole atom columns share a pixel and many (y' - min y') are exact multiples of lattice steps."""
    a, c = 2.456, 6.696
    b = a * np.sqrt(3.0)
    basis = np.array([[0, 0, 0], [0, 1 / 3, 0], [0.5, 0.5, 0], [0.5, 5 / 6, 0],
                      [0, 0, 0.5], [0, 2 / 3, 0.5], [0.5, 0.5, 0.5], [0.5, 1 / 6, 0.5]])
    cells = np.stack(np.meshgrid(np.arange(nx), np.arange(ny), np.arange(nz), indexing="ij"), -1).reshape(-1, 1, 3)
    frac = (cells + basis[None]).reshape(-1, 3)
    return frac * np.array([a, b, c])


def test_config3_graphite_crystal_1024_fill_bkg_smooth_against_oracle():
    """BASELINE configs[2] geometry (graphite crystal, fill_bkg, smooth = 25, N = 1024, q_num = 277):
    lattice coordinates put hundreds of atoms into one pixel at the symmetric rotations and make
    floor-divide arguments land on exact multiples; phi = 0 / 30 / 90 take the special chord branches."""
    coords = _graphite_crystal(28, 12, 8)                                    # 21504 atoms, 69 x 51 x 54 A
    el = np.array(["C"] * len(coords))
    r, max_q = 0.3, 2.0
    q = synth.pow2_q_voxel(r, 1024)
    dev = engine.resolve_device()
    codes, uniq, counts = engine.encode_elements_device(el, dev)
    table = comparison.f_table(uniq, 12700.0)
    atoms = engine.AtomSet(coords, r, 1024, dev, species=codes, table=table)
    N, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
    assert (N, q_num) == (1024, 277)
    avg = np.sum(counts * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3
    mk = lambda **kw: engine.SliceEngine(None, r, q_axis, N, avg, atoms.bounds[0], atoms.bounds[1], True, 25,
                                         atoms=atoms, **kw)
    sel = np.array([0.0, 30.0, 90.0, float(phis[217]), 120.0])
    fused, staged = mk(), mk()
    fused.run(sel)
    staged.run(sel, staged=True)
    assert np.array_equal(fused.counts(), staged.counts())
    assert np.abs(fused.sums() - staged.sums()).max() <= 1e-5 * staged.sums().max()
    f = ox.f_values_for(el, table=synth.fixed_f1f2)
    setup = ox.stage_a_setup(coords, f, r, q, max_q)
    q3 = (setup["q_num"],) * 3
    vsum, vcnt = np.zeros(q3), np.zeros(q3)
    for phi in sel:
        y_idx, z_idx, valid = ox.atom_pixel_indices(coords, phi, N, r)
        gy, gz, bbox = fused.atom_indices(phi)
        assert valid.all() and np.array_equal(gy, y_idx) and np.array_equal(gz, z_idx), phi
        ox.run_slice(vsum, vcnt, coords, setup, r, float(phi), True, 25)
    assert np.array_equal(fused.counts(), vcnt.astype(np.int64))
    assert np.abs(fused.sums() - vsum).max() <= 1e-4 * vsum.max()


def test_config4_pm6_2048_grid_and_psi_weight_file_against_oracle(golden, tmp_path):
    """BASELINE configs[3] geometry: the PM6 polymer slab on a 2048^2 grid (q_num = 555) - two slices
    against the oracle - and the experimental-comparison detector set-up: 91 psi orientations weighted
    from a .npy file (as test_experimental_data/resampled_ints_PM65CN_91.npy is used), max_q = 2."""
    g = golden("pm6.npz")
    coords = g["coords"]
    el = np.array([str(n) for n in g["element_names"]])[g["element_codes"]]
    r, max_q = 0.3, 2.0
    q = synth.pow2_q_voxel(r, 2048)
    dev = engine.resolve_device()
    codes, uniq, counts = engine.encode_elements_device(el, dev)
    table = comparison.f_table(uniq, 12700.0)
    atoms = engine.AtomSet(coords, r, 2048, dev, species=codes, table=table)
    N, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
    assert (N, q_num) == (2048, 555)
    avg = np.sum(counts * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3
    eng = engine.SliceEngine(None, r, q_axis, N, avg, atoms.bounds[0], atoms.bounds[1], True, 25, atoms=atoms)
    sel = phis[[5, 1100]]
    eng.run(sel)
    f = ox.f_values_for(el, table=synth.fixed_f1f2)
    setup = ox.stage_a_setup(coords, f, r, q, max_q)
    q3 = (setup["q_num"],) * 3
    vsum, vcnt = np.zeros(q3), np.zeros(q3)
    for phi in sel:
        ox.run_slice(vsum, vcnt, coords, setup, r, float(phi), True, 25)
    assert np.array_equal(eng.counts(), vcnt.astype(np.int64))
    assert np.abs(eng.sums() - vsum).max() <= 1e-4 * vsum.max()

    # stage B: psi weights from a file; a 120^3 stand-in grid keeps the oracle loop short
    rng = np.random.default_rng(4)
    V = 120
    iq = rng.random((V, V, V)) * 1e5
    ax = np.linspace(-2.03, 2.03, V)
    psis, phis_d, thetas = np.linspace(0, 90, 91), np.linspace(0, 179, 5), np.array([0.0])
    w = np.exp(-0.5 * ((psis - 20.0) / 9.0) ** 2) + 0.02
    w /= w.sum()
    wpath = str(tmp_path / "psi_weights_91.npy")
    np.save(wpath, w)
    args = (200, max_q, (90.0, 90.0, 90.0), ("psi", "phi", "psi"))
    det, h, v = comparison.detectormaker_fitting(iq, ax, ax, ax, *args, psis, wpath, phis_d, None, thetas, None)
    ones = lambda a: np.ones_like(a) / len(a)
    o_det, o_h, _ = ox.detectormaker(iq, ax, ax, ax, *args, psis, w, phis_d, ones(phis_d), thetas, ones(thetas))
    assert np.array_equal(h, o_h) and np.abs(det - o_det).max() <= 1e-4 * o_det.max()
    with pytest.raises(AssertionError, match="psi weights length"):
        comparison.detectormaker_fitting(iq, ax, ax, ax, *args, psis[:-1], wpath, phis_d, None, thetas, None)


@pytest.mark.parametrize("n_species", [1, 2, 3, 4, 6, 7])
def test_species_count_specialisations_1024_against_oracle(n_species):
    """The row kernel is compiled per species count for large transforms (1..6; 7 takes the generic
    flush): each one against the oracle on a 1024^2 grid, three rotations, counts bit-exact."""
    rng = np.random.default_rng(40 + n_species)
    names = np.array(["C", "H", "S", "O", "F", "N", "P"])[:n_species]
    coords = rng.random((6000, 3)) * [120.0, 90.0, 100.0]
    el = rng.choice(names, size=len(coords))
    r, max_q = 0.3, 2.0
    q = synth.pow2_q_voxel(r, 1024)
    phis = np.array([0.0, 41.3, 133.7])
    fill_bkg, smooth = bool(n_species % 2), 3 * (n_species % 3)
    iq, qx, qy, qz, eng = comparison.voxelgridmaker_fitting(coords, el, r, q, max_q, 12700.0, fill_bkg=fill_bkg,
                                                            smooth=smooth, phis=phis, return_state=True)
    assert eng.N == 1024 and eng.atoms.n_species == n_species
    f = ox.f_values_for(el, table=synth.fixed_f1f2)
    o_iq, o_qx, _, _, o_sum, o_cnt, _ = ox.voxelgridmaker(coords, f, r, q, max_q, fill_bkg, smooth, phis=phis)
    lo, hi = eng.window
    assert np.array_equal(eng.counts(), o_cnt.astype(np.int64)[lo:hi, lo:hi, lo:hi])
    assert np.array_equal(qx, o_qx)
    assert np.abs(iq - o_iq).max() <= 1e-4 * o_iq.max()


def test_wide_q_window_uses_the_full_last_pass():
    """A 4096^2 grid whose kept q-columns reach beyond +-512 of DC (max_q = 12 at r = 0.05): the
    band-limited last FFT pass does not apply and the kernels fall back to the full pass.  Fused ==
    staged (the staged kernels always run full transforms), counts bit-exact."""
    rng = np.random.default_rng(9)
    r, max_q = 0.05, 12.0
    coords = rng.random((20_000, 3)) * [150.0, 100.0, 150.0]
    el = rng.choice(np.array(["C", "H", "O"]), size=len(coords))
    q = synth.pow2_q_voxel(r, 4096)
    dev = engine.resolve_device()
    codes, uniq, counts = engine.encode_elements_device(el, dev)
    table = comparison.f_table(uniq, 12700.0)
    atoms = engine.AtomSet(coords, r, 4096, dev, species=codes, table=table)
    N, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
    assert N == 4096 and q_num > 1024
    avg = np.sum(counts * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3
    window = engine.crop_range(q_axis, max_q)
    mk = lambda: engine.SliceEngine(None, r, q_axis, N, avg, atoms.bounds[0], atoms.bounds[1], True, 5,
                                    atoms=atoms, window=window)
    fused, staged = mk(), mk()
    assert fused.KC > 1024                                  # kept columns beyond +-512
    sel = np.array([17.0, 93.0])
    fused.run(sel)
    staged.run(sel, staged=True)
    a, b = fused.count2.cpu().numpy(), staged.count2.cpu().numpy()
    assert np.array_equal(a, b) and a.sum() > 0
    fs, ss = fused.vsum, staged.vsum
    assert float((fs - ss).abs().max()) <= 1e-5 * float(ss.max())


NAMED_CASES = ["config1_graphite_medium", "config1b_graphite_medium_smooth3", "config2_silicon_medium",
               "config3_graphite_large"]


@pytest.mark.parametrize("name", NAMED_CASES)
def test_named_config_files_against_reference_fixture(golden, name):
    """BASELINE configs[0..2] on their NAMED input files at the configured sizes - graphite_medium.xyz with the
    template values (N = 1048, Bluestein), silicon_medium.xyz at q = 0.01 (N = 2095, q_num 569) with a 512^2
    stage B, graphite_large.xyz at N = 1024 with fill_bkg and smooth 25 - through the public drivers, against
    fixtures produced by the UNMODIFIED reference (oracle/make_golden.py `named`).  Counts bit-exact, sums,
    iq and detector image within 1e-4 of the reference maximum (and, stricter than the north-star bar, the
    voxels away from the DC column within 1e-4 of THEIR maximum)."""
    g = golden(name + ".npz")
    r, q, max_q = float(g["r"]), float(g["q"]), float(g["max_q"])
    iq, qx, qy, qz, eng = comparison.voxelgridmaker_fitting(
        g["coords"], g["elements"], r, q, max_q, float(g["energy"]), fill_bkg=bool(g["fill_bkg"]),
        smooth=int(g["smooth"]), phis=g["probe_phis"], return_state=True)
    assert eng.N == int(g["grid_size"]) and len(eng.q_axis) == int(g["q_num"])
    lo, hi = (int(v) for v in g["crop"])
    assert eng.window == (lo, hi)
    V = hi - lo
    H, m = g["H"].astype(np.int64), g["m"].astype(np.int64)
    assert np.array_equal(eng.count2.cpu().numpy().astype(np.int64).reshape(V, V), H[lo:hi, lo:hi])
    assert np.array_equal(eng.row_hist.cpu().numpy().astype(np.int64), m[lo:hi])
    p = g["pairs"]
    inside = np.all((p >= lo) & (p < hi), axis=1)
    pw = p[inside] - lo
    want = g["vsum_pairs"][inside][:, lo:hi].astype(np.float64)
    got = eng.sums()[pw[:, 0], pw[:, 1], :].astype(np.float64)
    assert np.abs(got - want).max() <= 1e-4 * float(g["vsum_max"])
    col_max = want.max(axis=1)
    off = col_max < 0.5 * col_max.max()                  # every stored column but the one through q = 0
    if off.any() and want[off].max() > 1e-12 * float(g["vsum_max"]):     # (config 1: a bare pedestal, off-DC = round-off)
        assert np.abs(got[off] - want[off]).max() <= 1e-4 * want[off].max()
    ip = g["iq_pairs"]
    assert np.array_equal(qx, g["q_axis"][lo:hi])
    assert np.abs(iq[ip[:, 0], ip[:, 1], :] - g["iq_values"]).max() <= 1e-4 * float(g["iq_max"])
    if "det" in g.files:
        psis, phis, thetas = g["det_psis"], g["det_phis"], g["det_thetas"]
        det, _, _ = comparison.detectormaker_fitting(iq, qx, qy, qz, int(g["det_P"]), max_q, (90.0, 90.0, 90.0),
                                                     ("psi", "phi", "psi"), psis, None, phis, None, thetas, None)
        assert np.abs(det - g["det"]).max() <= 1e-4 * float(g["det_max"])


def test_row_with_more_than_65535_atoms_takes_the_chunked_counters():
    """A thin-z crystalline slab: one z pixel row holds > 65535 atoms, so the 16-bit species counters of the
    fused row kernel are flushed in chunks (gx_fused.cu, `!single` branch).  Fused == staged == oracle."""
    rng = np.random.default_rng(3)
    n = 150_000
    coords = np.empty((n, 3))
    coords[:, 0] = rng.random(n) * 30.0
    coords[:, 1] = rng.random(n) * 24.0
    coords[:, 2] = rng.random(n) * 0.55                     # two pixel rows at r = 0.3: ~82 k and ~68 k atoms
    coords[0, 2], coords[1, 2] = 0.0, 36.0                  # z extent of the slab (a third, sparse row)
    elements = rng.choice(np.array(["C", "H", "S"]), size=n, p=[0.6, 0.3, 0.1])
    r, max_q = 0.3, 1.5
    q = synth.pow2_q_voxel(r, 256)
    f = ox.f_values_for(elements, table=synth.fixed_f1f2)
    setup = ox.stage_a_setup(coords, f, r, q, max_q)
    phis = setup["phis"][[0, 17, 50]]
    rows = np.bincount(((coords[:, 2] - coords[:, 2].min()) // r).astype(int))
    assert rows.max() > 65535
    iq, qx, qy, qz, eng = comparison.voxelgridmaker_fitting(coords, elements, r, q, max_q, 12700.0, fill_bkg=False,
                                                            smooth=0, phis=phis, return_state=True)
    o_iq, _, _, _, o_sum, o_cnt, _ = ox.voxelgridmaker(coords, f, r, q, max_q, False, 0, phis=phis)
    lo, hi = eng.window
    assert np.array_equal(eng.counts(), o_cnt.astype(np.int64)[lo:hi, lo:hi, lo:hi])
    assert np.abs(eng.sums() - o_sum[lo:hi, lo:hi, lo:hi]).max() <= 1e-4 * o_sum.max()
    assert np.abs(iq - o_iq).max() <= 1e-4 * o_iq.max()
    staged = engine.SliceEngine(None, r, eng.q_axis, eng.N, eng.avg_voxel_f, eng.x_bound, eng.y_bound, False, 0,
                                atoms=eng.atoms, window=eng.window)
    staged.run(phis, staged=True)
    assert np.array_equal(staged.counts(), eng.counts())
    assert np.abs(staged.sums() - eng.sums()).max() <= 1e-5 * o_sum.max()
    # with background + blend too (the chunked path applies (d, my) after the last chunk)
    iq2, *_ = comparison.voxelgridmaker_fitting(coords, elements, r, q, max_q, 12700.0, fill_bkg=True, smooth=3,
                                                phis=phis)
    o_iq2 = ox.voxelgridmaker(coords, f, r, q, max_q, True, 3, phis=phis)[0]
    assert np.abs(iq2 - o_iq2).max() <= 1e-4 * o_iq2.max()


def test_grid_size_4189_needs_the_16384_point_bluestein_against_oracle():
    """r = 0.15, q = 0.01 -> grid_size = ceil(2 pi / (q r)) = 4189 (SURVEY hard part 2: config 5 at r = 0.15):
    not a power of two and above 4096, so rows and columns go through the 16384-point chirp-z transform
    (round 1 refused every non-power-of-two size above 4096).  One slice against the oracle."""
    r, q, max_q = 0.15, 0.01, 2.0
    coords, el = synth.random_slab(40_000, (200.0, 120.0, 180.0), seed=9)
    f = ox.f_values_for(el, table=synth.fixed_f1f2)
    setup = ox.stage_a_setup(coords, f, r, q, max_q)
    assert setup["grid_size"] == 4189
    phi = np.array([setup["phis"][321]])
    iq, qx, qy, qz, eng = comparison.voxelgridmaker_fitting(coords, el, r, q, max_q, 12700.0, fill_bkg=True, smooth=7,
                                                            phis=phi, return_state=True)
    q3 = (setup["q_num"],) * 3
    vsum, vcnt = np.zeros(q3), np.zeros(q3)
    ox.run_slice(vsum, vcnt, coords, setup, r, float(phi[0]), True, 7)
    lo, hi = eng.window
    assert np.array_equal(eng.counts(), vcnt.astype(np.int64)[lo:hi, lo:hi, lo:hi])
    assert np.abs(eng.sums() - vsum[lo:hi, lo:hi, lo:hi]).max() <= 1e-4 * vsum.max()
    o_iq = ox.finalize_voxelgrid(vsum, vcnt, setup["q_axis"], max_q)[0]
    assert np.abs(iq - o_iq).max() <= 1e-4 * o_iq.max()
