"""Size-independent properties and internal cross-checks of the CUDA path."""
import numpy as np
import pytest
import torch

from giwaxsim_b200 import _lib, engine, synth
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
def test_fast_detector_kernels_are_bit_identical_to_exact(inits, axs):
    """fixed-point affine gather == fp32-filtered gather == all-fp64 gather: same voxel index for
    every pixel of every orientation, including grid-aligned orientations where every pixel sits
    on a voxel edge and pixels far outside the voxel box (clamped)."""
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
    img_exact, _ = det.accumulate(gx, gy, gz, R, w, kernel="exact")
    img_fast, _ = det.accumulate(gx, gy, gz, R, w, kernel="filtered", count_slow=True)
    assert torch.equal(img_exact, img_fast)
    assert 0.0 <= det.last_slow_fraction <= 1.0
    img_aff, _ = det.accumulate(gx, gy, gz, R, w, kernel="affine", count_slow=True)
    assert det.last_kernel == "affine"
    # fp32 partial sums over <= 64 orientations, then fp64: not bitwise, but close
    assert float((img_aff - img_exact).abs().max()) <= 2e-6 * float(img_exact.abs().max())
    for o in range(len(w)):
        _, a = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="exact")
        _, c = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="affine")
        assert torch.equal(a, c), o
        if o in (0, 7, 19, 39):
            _, b = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="filtered")
            assert torch.equal(a, b)
    # default dispatch picks the affine kernel for make_detector grids
    det.accumulate(gx, gy, gz, R, w)
    assert det.last_kernel == "affine"


def test_affine_detector_kernel_partial_tiles_rectangular_and_split():
    """rows != cols, sizes that are not multiples of the 32 x 16 tile, and the orientation-range
    split with atomic accumulation used for small images."""
    rng = np.random.default_rng(5)
    V = 203
    q = np.linspace(-2.02, 2.02, V)
    iq = rng.random((V, V, V)).astype(np.float32)
    dev = engine.resolve_device()
    gx, gy, gz, _, _ = comparison.detector_base_device(150, 2.0, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), dev)
    gx, gy, gz = (g[:77, 3:140].contiguous() for g in (gx, gy, gz))
    psis = np.linspace(0.0, 89.0, 40)
    phis = np.array([0.0, 3.0, 90.0, 135.0, 177.0])
    ones = lambda a: np.ones(len(a)) / len(a)
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, ones(psis), phis, ones(phis), [0.0],
                                     np.ones(1))
    det = engine.DetectorEngine(iq, q, q, q)
    a, _ = det.accumulate(gx, gy, gz, R, w, kernel="exact")
    b, _ = det.accumulate(gx, gy, gz, R, w, kernel="affine", count_slow=True)
    assert float((a - b).abs().max()) <= 2e-6 * float(a.abs().max())
    assert det.last_slow_fraction < 0.01
    for o in (0, 1, 57, 123, 199):
        _, ia = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="exact")
        _, ib = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="affine")
        assert torch.equal(ia, ib), o


def test_non_affine_grid_falls_back_to_generic_kernels():
    rng = np.random.default_rng(6)
    V = 64
    q = np.linspace(-2.0, 2.0, V)
    iq = rng.random((V, V, V)).astype(np.float32)
    dev = engine.resolve_device()
    gx, gy, gz, _, _ = comparison.detector_base_device(64, 2.0, (0.0, 0.0, 0.0), ("None", "None", "None"), dev)
    gx = gx + 0.05 * torch.sin(gy * 3.0)                 # a curved detector: not affine in (row, col)
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), [10.0, 20.0], np.ones(2) / 2, [5.0],
                                     np.ones(1), [0.0], np.ones(1))
    det = engine.DetectorEngine(iq, q, q, q)
    a, _ = det.accumulate(gx, gy, gz, R, w, kernel="exact")
    b, _ = det.accumulate(gx, gy, gz, R, w)
    assert det.last_kernel in ("filtered", "exact")
    assert torch.equal(a, b)
    with pytest.raises(_lib.GxError):
        det.accumulate(gx, gy, gz, R, w, kernel="affine")


def test_fast_detector_kernels_rarely_fall_back():
    rng = np.random.default_rng(2)
    V = 403
    q = np.linspace(-2.01, 2.01, V)
    iq = rng.random((V, V, V)).astype(np.float32)
    dev = engine.resolve_device()
    gx, gy, gz, _, _ = comparison.detector_base_device(512, 2.0, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), dev)
    psis = np.linspace(3.1, 88.3, 24)
    ones = np.ones(1)
    det = engine.DetectorEngine(iq, q, q, q)
    # tilted planes (phi = 7.3 deg): only pixels within the error bound of a voxel edge fall back
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones(24) / 24, [7.3], ones, [0.4], ones)
    a, _ = det.accumulate(gx, gy, gz, R, w, kernel="exact")
    b, _ = det.accumulate(gx, gy, gz, R, w, kernel="filtered", count_slow=True)
    assert torch.equal(a, b)
    assert det.last_slow_fraction < 0.02, det.last_slow_fraction
    c, _ = det.accumulate(gx, gy, gz, R, w, kernel="affine", count_slow=True)
    assert float((a - c).abs().max()) <= 2e-6 * float(a.abs().max())
    assert det.last_slow_fraction < 1e-3, det.last_slow_fraction
    # phi = theta = 0: the plane lies in q_z = 0, i.e. exactly on a voxel edge, for every pixel
    # (rounding noise picks the bin) -> the fp32 filter sends every pixel to the exact chain; the
    # affine kernel models the coordinate as constant / a single rounding step on the host
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones(24) / 24, [0.0], ones, [0.0], ones)
    a, _ = det.accumulate(gx, gy, gz, R, w, kernel="exact")
    b, _ = det.accumulate(gx, gy, gz, R, w, kernel="filtered", count_slow=True)
    assert torch.equal(a, b)
    assert det.last_slow_fraction > 0.9
    c, _ = det.accumulate(gx, gy, gz, R, w, kernel="affine", count_slow=True)
    assert float((a - c).abs().max()) <= 2e-6 * float(a.abs().max())
    assert det.last_slow_fraction < 1e-3, det.last_slow_fraction
    for o in (0, 11, 23):
        _, ia = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="exact")
        _, ic = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="affine")
        assert torch.equal(ia, ic)


def test_affine_detector_kernel_rounding_step_model():
    """Dyadic voxel axis: q = 0 is exactly a voxel edge AND a binade boundary of p - qmin, so the
    1e-16 rounding noise of the in-plane coordinate decides the voxel pixel by pixel."""
    rng = np.random.default_rng(3)
    q = -2.0 + np.arange(513) * 2.0 ** -7
    iq = rng.random((513, 513, 513)).astype(np.float32)
    dev = engine.resolve_device()
    gx, gy, gz, _, _ = comparison.detector_base_device(400, 2.0, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), dev)
    psis = np.linspace(0, 89.75, 16)
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones(16) / 16, [0.0], np.ones(1),
                                     [0.0], np.ones(1))
    det = engine.DetectorEngine(iq, q, q, q)
    for o in range(16):
        _, ia = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="exact")
        _, ic = det.accumulate(gx, gy, gz, R, w, probe=o, kernel="affine", count_slow=True)
        assert torch.equal(ia, ic), o
    assert det.last_plan[7] >= 8          # the step model was used
    assert det.last_slow_fraction < 0.05


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
    # fp32 partial sums inside the kernel (<= 64 orientations each), fp64 across them
    assert torch.allclose(a + b, c, rtol=2e-6, atol=0)


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


@pytest.mark.parametrize("dtype", ["<U1", "<U2"])
def test_device_species_coding_matches_host_coding(dtype):
    rng = np.random.default_rng(4)
    symbols = np.array(["C", "H", "S", "O", "F"] if dtype == "<U1" else ["C", "H", "Si", "O", "Cl", "Br"], dtype=dtype)
    el = symbols[rng.integers(0, len(symbols), 100_003)]
    dev = engine.resolve_device()
    codes, uniq, counts = engine.encode_elements_device(el, dev)
    h_codes, h_uniq = engine.encode_values(el)
    assert [str(u) for u in uniq] == [str(u) for u in h_uniq]
    assert np.array_equal(codes.cpu().numpy(), h_codes)
    assert np.array_equal(counts, np.bincount(h_codes, minlength=len(h_uniq)))
    # not element-symbol shaped input -> None (the caller uses the host coder)
    assert engine.encode_elements_device(np.arange(5).astype(complex), dev) is None
    assert engine.encode_elements_device(np.array(["C", "é"]), dev) is None
    many = np.array(["%c%c" % (65 + i // 26, 97 + i % 26) for i in range(40)])
    assert engine.encode_elements_device(many, dev) is None


def test_to_host_f64_roundtrip():
    dev = engine.resolve_device()
    for dt in (torch.float32, torch.float64):
        t = torch.randn(37, 41, 5, device=dev, dtype=dt)
        out = engine.to_host_f64(t)
        assert out.dtype == np.float64 and out.shape == (37, 41, 5)
        assert np.array_equal(out, t.cpu().double().numpy())
        out2 = engine.to_host_f64(t * 2)                 # the staging buffer is reused, results are not aliased
        assert np.array_equal(out, t.cpu().double().numpy()) and np.array_equal(out2, 2 * out)


def test_to_host_f64_pipelined_chunks(monkeypatch):
    """Large results are downloaded in chunks, each widened while the next is on the wire."""
    dev = engine.resolve_device()
    monkeypatch.setattr(engine, "PIPELINED_DOWNLOAD_MIN_BYTES", 1024)
    for chunks in (1, 3, 8, 1000):
        monkeypatch.setattr(engine, "PIPELINED_DOWNLOAD_CHUNKS", chunks)
        t = torch.randn(101, 67, 13, device=dev, dtype=torch.float32)
        out = engine.to_host_f64(t)
        assert out.dtype == np.float64 and np.array_equal(out, t.cpu().double().numpy())
        view = np.zeros((101, 67, 13))
        engine.to_host_f64(t, out=view)
        assert np.array_equal(view, out)


def test_two_engines_on_two_streams_do_not_corrupt_each_other():
    """gx_slices_fused stages per-rotation scalars through __constant__ tables that belong to one launch at a
    time: launches of one device arriving on DIFFERENT streams are ordered through an event behind the row kernel
    (round 1 only documented 'one stream per device').  Two engines with different atoms, interleaved batch by
    batch on two streams, must give what each gives alone."""
    dev = engine.resolve_device()
    r, max_q = 0.25, 1.5
    q = synth.pow2_q_voxel(r, 256)
    jobs = []
    for seed, box in ((4, (60.0, 35.0, 50.0)), (9, (40.0, 55.0, 45.0))):
        coords, el = synth.random_slab(25_000, box, seed=seed)
        codes, uniq, table = comparison.species_table(el, 12700.0)
        atoms = engine.AtomSet(coords, r, 256, dev, species=codes, table=table)
        N, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
        avg = np.sum(np.bincount(codes, minlength=len(table)) * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3
        make = lambda a=atoms, qa=q_axis, n=N, av=avg: engine.SliceEngine(None, r, qa, n, av, a.bounds[0], a.bounds[1],
                                                                        True, 4, atoms=a)
        jobs.append((make, phis[::3]))
    alone = []
    for make, phis in jobs:
        e = make()
        e.run(phis)
        alone.append((e.counts(), e.sums().astype(np.float64)))
    torch.cuda.synchronize()
    engines = [make() for make, _ in jobs]
    streams = [torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)]
    step = 7                                            # small batches: many interleaved launch pairs
    n = max(len(p) for _, p in jobs)
    for i0 in range(0, n, step):
        for k, (e, (_, phis)) in enumerate(zip(engines, jobs)):
            chunk = phis[i0:i0 + step]
            if len(chunk):
                with torch.cuda.stream(streams[k]):
                    e.run(chunk)
    torch.cuda.synchronize()
    for e, (cnt, sums) in zip(engines, alone):
        assert np.array_equal(e.counts(), cnt)
        assert np.abs(e.sums() - sums).max() <= 1e-5 * sums.max()


def test_tma_brick_detector_variant_is_identical_to_the_ldg_kernel():
    """The TMA-brick variant of the affine detector gather (8^3 voxel bricks streamed by cp.async.bulk.tensor.3d
    through an mbarrier ring; north_star kernel (4)) reads the same voxels in the same order as the L1-fed
    kernel: images must be identical, also for a partial-tile size and orientations that make tiles touch the
    box boundary (those tiles run the clamped path)."""
    dev = engine.resolve_device()
    rng = np.random.default_rng(3)
    V = 121
    iq = torch.from_numpy(rng.random((V, V, V)).astype(np.float32) * 1e5).to(dev)
    q = np.linspace(-2.02, 2.02, V)
    for P, max_q in ((900, 2.0), (517, 1.1)):      # fine enough pixels for a tile to span < 7 voxels
        gx, gy, gz, _, _ = comparison.detector_base_device(P, max_q, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), dev)
        psis, phis, thetas = np.linspace(0, 89, 9), np.linspace(0, 170, 5), np.array([0.0, 3.0])
        ones = lambda a: np.ones_like(a) / len(a)
        R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, ones(psis), phis, ones(phis), thetas, ones(thetas))
        det = engine.DetectorEngine(iq, q, q, q, device=dev)
        a, _ = det.accumulate(gx, gy, gz, R, w, kernel="affine")
        b, _ = det.accumulate(gx, gy, gz, R, w, kernel="affine_tma")
        assert det.last_kernel == "affine_tma"
        assert torch.equal(a, b)


@pytest.mark.parametrize("grid", [1024, 2048])
def test_tma_fed_column_kernel_equals_the_ldg_fed_one(monkeypatch, grid):
    """slice_cols_tma<L> (N = 1024 through GIWAXS_B200_TMA_MIN_L=10, N = 2048 by default) against slice_cols_fused
    (GIWAXS_B200_NO_TMA=1): identical counts, sums equal to fp32 accumulation order.  (N = 4096 is covered by the
    full-size tests against the oracle.)"""
    coords, el = synth.random_slab(200_000, (grid * 0.13, grid * 0.06, grid * 0.12), seed=21)
    r, max_q = 0.15, 2.0
    q = synth.pow2_q_voxel(r, grid)
    codes, uniq, table = comparison.species_table(el, 12700.0)
    dev = engine.resolve_device()
    atoms = engine.AtomSet(coords, r, grid, dev, species=codes, table=table)
    N, q_num, q_axis, phis = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
    assert N == grid
    avg = np.sum(np.bincount(codes, minlength=len(table)) * np.asarray(table)) / np.prod(atoms.bounds) * r ** 3
    window = engine.crop_range(q_axis, max_q)
    sel = phis[:: max(1, len(phis) // 40)]

    def run():
        e = engine.SliceEngine(None, r, q_axis, N, avg, atoms.bounds[0], atoms.bounds[1], True, 9, atoms=atoms,
                               window=window)
        e.run(sel)
        return e.counts(), e.sums().astype(np.float64)

    monkeypatch.setenv("GIWAXS_B200_TMA_MIN_L", "10")
    c_tma, s_tma = run()
    monkeypatch.setenv("GIWAXS_B200_NO_TMA", "1")
    c_ldg, s_ldg = run()
    assert c_tma.sum() > 0 and np.array_equal(c_tma, c_ldg)
    assert np.abs(s_tma - s_ldg).max() <= 2e-6 * s_ldg.max()
