"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the
UNMODIFIED reference (through oracle/ref_shim.py, serial drivers) in the build
container, where /root/reference is mounted.  The fixtures travel with the
repo; nothing reads /root/reference at test time on the GPU box.

    python -m oracle.make_golden            # all cases
    python -m oracle.make_golden pm6        # only the (slow) end-to-end case

Cases
  graphite262  graphite_medium.xyz cluster, N=262 (2*131 -> Bluestein), fill_bkg, smooth=5
  silicon256   silicon_medium.xyz cluster,  N=256 (pow2), no background, no smoothing
  clipped128   graphite cluster on a grid smaller than its y/x extent (valid-mask path)
  detector     stage B on the graphite262 voxel grid: 64^2 and 65^2 detectors
  twostep      two seeded clusters through the reference's generate_voxel_grid_low_mem (aff_num_qs 1
               and 3, N=210 -> Bluestein) and the old_modules/voxelgridmaker.py crop/average/f0 loop
  named        BASELINE.json configs[0..2] on their NAMED input files at the configured sizes, a few probe
               slices each through the unmodified reference (compact: the count grid as its exact rank-1
               factors, the sums / iq on a subset of the hit (qy, qx) columns):
                 config1_graphite_medium  r 0.3, q 0.02 (template values) -> N = 1048 (Bluestein), q_num 285
                 config2_silicon_medium   r 0.3, q 0.01 -> N = 2095 (largest Bluestein), q_num 569, + 512^2 stage B
                 config3_graphite_large   N = 1024, fill_bkg, smooth 25
  pm6          config_templates/simulate_GIWAXS_config.txt end to end (N=1048, 892 slices,
               2880 orientations x 500^2) + the reference's own golden det_sum.npy
"""
import os
import sys
import time

import numpy as np

from . import ref_shim

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
REF = ref_shim.REFERENCE_ROOT
PROBE_SLICES = {"graphite262": [0, 7, 100, 224], "silicon256": [0, 31, 109], "clipped128": [0, 5, 40]}


def _stage_a_case(name, coords, elements, r, q, max_q, fill_bkg, smooth, energy=12700.0):
    ref = ref_shim.load()
    setup = ref_shim.stage_a_setup(coords, elements, r, q, max_q, energy)
    cap = {}
    t0 = time.time()
    vsum, vcnt = ref_shim.run_slices_serial(coords, setup, r, fill_bkg, smooth, capture=cap)
    iq, qx, qy, qz = ref_shim.finalize_serial(vsum, vcnt, setup["q_axis"], max_q)
    N = setup["grid_size"]
    out = dict(coords=coords, elements=np.asarray(elements), r=r, q=q, max_q=max_q, fill_bkg=fill_bkg,
               smooth=smooth, energy=energy, grid_size=N, q_num=setup["q_num"], q_axis=setup["q_axis"],
               phis=setup["phis"], f_values=setup["f_values"], avg_voxel_f=setup["avg_voxel_f"],
               bounds=np.array([setup["x_bound"], setup["y_bound"], setup["z_bound"]]),
               vsum=vsum.astype(np.float32), vsum_max=vsum.max(), vcnt=vcnt.astype(np.uint16),
               iq=iq, q_crop=qx, probe=np.array(PROBE_SLICES[name]))
    assert vcnt.max() < 65536 and np.array_equal(vcnt, np.round(vcnt))
    for i in PROBE_SLICES[name]:
        phi = setup["phis"][i]
        # atom indices exactly as voxelgrids.py:319-335 computes them
        rot = ref.utilities.rotate_coords_z(coords, phi)
        rot[:, 1] -= np.min(rot[:, 1])
        rot[:, 2] -= np.min(rot[:, 2])
        y_idx = (rot[:, 1] // r).astype(int)
        z_idx = (rot[:, 2] // r).astype(int)
        valid = (y_idx >= 0) & (y_idx < N) & (z_idx >= 0) & (z_idx < N)
        out["y_idx_%d" % i] = y_idx
        out["z_idx_%d" % i] = z_idx
        out["bbox_%d" % i] = np.array([y_idx[valid].min(), y_idx[valid].max(),
                                       z_idx[valid].min(), z_idx[valid].max()])
        out["grid_%d" % i] = cap["grid"][i].astype(np.complex64)
        out["grid_absmax_%d" % i] = np.abs(cap["grid"][i]).max()
        out["iq2d_%d" % i] = cap["iq_2d"][i].astype(np.float32)
        out["iq2d_max_%d" % i] = cap["iq_2d"][i].max()
        hx, hy, vz = cap["det_h_qx"][i], cap["det_h_qy"][i], cap["det_v_qz"][i]
        qa = setup["q_axis"]
        # voxelgrids.py:475-499
        cmask = (hx <= np.max(qa)) & (hx >= np.min(qa)) & (hy <= np.max(qa)) & (hy >= np.min(qa))
        rmask = (vz <= np.max(qa)) & (vz >= np.min(qa))
        dq = np.diff(qa)[0]
        out["colmask_%d" % i] = cmask
        out["ix_%d" % i] = ((hx[cmask] - np.min(qa)) // dq).astype(int)
        out["iy_%d" % i] = ((hy[cmask] - np.min(qa)) // dq).astype(int)
        out["rowmask_%d" % i] = rmask
        out["iz_%d" % i] = ((vz[rmask] - np.min(qa)) // dq).astype(int)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%s: N=%d q_num=%d slices=%d  (%.1fs)" % (name, N, setup["q_num"], len(setup["phis"]), time.time() - t0))
    return iq, qx


def graphite262():
    ref = ref_shim.load()
    coords, elements = ref.utilities.load_xyz(os.path.join(REF, "test_input_files/graphite_medium.xyz"))
    return _stage_a_case("graphite262", coords, elements, 0.3, 0.08, 2.0, True, 5)


def silicon256():
    ref = ref_shim.load()
    coords, elements = ref.utilities.load_xyz(os.path.join(REF, "test_input_files/silicon_medium.xyz"))
    r = 0.3
    return _stage_a_case("silicon256", coords, elements, r, 2 * np.pi / (r * 255.5), 2.0, False, 0)


def clipped128():
    """Grid smaller than the slab's x/y extent but not its z extent: allowed by the
    reference's min-bound check (comparison.py:715-717), exercises the valid mask."""
    ref = ref_shim.load()
    coords, elements = ref.utilities.load_xyz(os.path.join(REF, "test_input_files/graphite_medium.xyz"))
    coords = coords * np.array([3.0, 2.5, 1.0])
    r = 0.3
    return _stage_a_case("clipped128", coords, elements, r, 2 * np.pi / (r * 127.5), 2.0, False, 3)


def detector(iq, q_crop):
    ref = ref_shim.load()
    out = dict(iq=iq, q=q_crop, max_q=2.0)
    cases = {"a": (64, (90.0, 90.0, 90.0), ("psi", "phi", "psi"), True),
             "b": (65, (10.0, 20.5, 0.0), ("theta", "phi", "None"), True),
             "c": (48, (0.0, 0.0, 0.0), ("None", "None", "None"), False)}
    rng = np.random.default_rng(5)
    for tag, (P, vals, axs, mirror) in cases.items():
        psis = np.linspace(75, 90, 5)
        phis = np.linspace(0, 179, 7)
        thetas = np.linspace(0, 1, 2)
        pw = rng.random(5); pw /= pw.sum()
        fw = np.ones(7) / 7
        tw = np.array([0.25, 0.75])
        raw, h, v = ref_shim.detectormaker_serial(iq, q_crop, q_crop, q_crop, P, 2.0, vals, axs, psis, pw,
                                                  phis, fw, thetas, tw, mirror=mirror, raw=True)
        fin, _, _ = ref_shim.detectormaker_serial(iq, q_crop, q_crop, q_crop, P, 2.0, vals, axs, psis, pw,
                                                  phis, fw, thetas, tw, mirror=mirror)
        gx, gy, gz, _, _ = ref_shim.detector_base(P, 2.0, vals, axs)
        d = ref.detector
        probes = [0, 33, 69]      # flat orientation numbers, psi outermost
        for o in probes:
            ip, rem = divmod(o, 14)
            jf, kt = divmod(rem, 2)
            x2, y2, z2 = d.rotate_psi_phi_theta(gx, gy, gz, psis[ip], phis[jf], thetas[kt])
            dq = np.diff(q_crop)[0]
            ix = np.clip(((x2.ravel() - q_crop.min()) // dq).astype(int), 0, iq.shape[1] - 1)
            iy = np.clip(((y2.ravel() - q_crop.min()) // dq).astype(int), 0, iq.shape[0] - 1)
            iz = np.clip(((z2.ravel() - q_crop.min()) // dq).astype(int), 0, iq.shape[2] - 1)
            out["%s_index_%d" % (tag, o)] = ((iy * iq.shape[1] + ix) * iq.shape[2] + iz).astype(np.int64)
        out.update({tag + "_P": P, tag + "_vals": np.array(vals), tag + "_axs": np.array(axs),
                    tag + "_mirror": mirror, tag + "_psis": psis, tag + "_phis": phis, tag + "_thetas": thetas,
                    tag + "_pw": pw, tag + "_fw": fw, tag + "_tw": tw, tag + "_raw": raw, tag + "_final": fin,
                    tag + "_gx": gx, tag + "_gy": gy, tag + "_gz": gz, tag + "_probes": np.array(probes)})
        # the pure per-grid functions
        out[tag + "_intersect"] = d.intersect_detector(iq, q_crop, q_crop, q_crop, gx, gy, gz)
        out[tag + "_mirrored"] = d.mirror_vertical_horizontal(raw)
    np.savez_compressed(os.path.join(OUT, "detector.npz"), **out)
    print("detector: done")


def pm6():
    """Default config end to end.  Stage A/B run through the oracle port (bit-identical
    to the serial reference on every smaller case, and ~8x faster with threads) and the
    result is checked here against the reference's shipped golden image."""
    from . import giwaxs_oracle as ox
    ref = ref_shim.load()
    cfg = ref.utilities.parse_config_file(os.path.join(REF, "config_templates/simulate_GIWAXS_config.txt"))
    path = os.path.join(REF, "test_input_files/PM6_sample.pdb")
    cell = ref.utilities.load_pdb_cell_params(path)
    t0 = time.time()
    coords, elements = ref.comparison.slabmaker_fitting(path, float(cfg["x_size"]), float(cfg["y_size"]),
                                                        float(cfg["z_size"]), *cell)
    r, q, max_q = float(cfg["r_voxel_size"]), float(cfg["q_voxel_size"]), float(cfg["max_q"])
    f = ox.f_values_for(elements)
    iq, qx, qy, qz, vsum, vcnt, setup = ox.voxelgridmaker(coords, f, r, q, max_q, True, int(cfg["smooth"]),
                                                          threads=os.cpu_count())
    print("pm6 stage A %.0fs: atoms=%d N=%d slices=%d iq=%s" % (time.time() - t0, len(coords),
          setup["grid_size"], len(setup["phis"]), iq.shape))
    P = int(cfg["num_pixels"])
    psis = np.linspace(float(cfg["psi_start"]), float(cfg["psi_end"]), int(cfg["psi_num"]))
    phis = np.linspace(float(cfg["phi_start"]), float(cfg["phi_end"]), int(cfg["phi_num"]))
    thetas = np.linspace(float(cfg["theta_start"]), float(cfg["theta_end"]), int(cfg["theta_num"]))
    vals = tuple(float(cfg["angle_init_val%d" % k]) for k in (1, 2, 3))
    axs = tuple(cfg["angle_init_ax%d" % k] for k in (1, 2, 3))
    t0 = time.time()
    det, h, v = ox.detectormaker(iq, qx, qy, qz, P, max_q, vals, axs, psis, np.ones_like(psis) / len(psis),
                                 phis, np.ones_like(phis) / len(phis), thetas, np.ones_like(thetas) / len(thetas),
                                 mirror=True, threads=os.cpu_count())
    print("pm6 stage B %.0fs" % (time.time() - t0))
    kh, kv = np.where(h >= 0)[0], np.where(v >= 0)[0]          # simulate_GIWAXS.py:172-177
    quad = det[np.ix_(kv, kh)]
    gold = np.load(os.path.join(REF, "output_data/PM6_sample/det_sum.npy"))
    rel = np.abs(quad - gold).max() / gold.max()
    corr = np.corrcoef(np.log(quad).ravel(), np.log(gold).ravel())[0, 1]
    print("pm6: oracle vs reference golden det_sum: max|diff|/max = %.4g, corr(log) = %.6f" % (rel, corr))
    codes = {e: i for i, e in enumerate(sorted(set(elements)))}
    np.savez_compressed(os.path.join(OUT, "pm6.npz"), coords=coords,
                        element_codes=np.array([codes[e] for e in elements], dtype=np.uint8),
                        element_names=np.array(sorted(codes)), r=r, q=q, max_q=max_q, smooth=int(cfg["smooth"]),
                        P=P, psis=psis, phis=phis, thetas=thetas, vals=np.array(vals), axs=np.array(axs),
                        det_quadrant_oracle=quad, det_h=h[kh], det_v=v[kv], golden_rel=rel, golden_corr=corr,
                        iq_max=iq.max(), iq_shape=np.array(iq.shape),
                        iq_center_plane=iq[:, :, iq.shape[2] // 2].astype(np.float32))
    np.save(os.path.join(OUT, "pm6_det_sum_ref.npy"), gold)


def _named_case(name, xyz, r, q, max_q, fill_bkg, smooth, probes, detector_pixels=0, keep_every=4, energy=12700.0):
    """A few slices of a named input file at its configured grid size through the unmodified reference
    (rotate_project_fft_coords -> process_file2 on fresh accumulators, comparison.py:769-786 finalise)."""
    ref = ref_shim.load()
    coords, elements = ref.utilities.load_xyz(os.path.join(REF, "test_input_files", xyz))
    setup = ref_shim.stage_a_setup(coords, elements, r, q, max_q, energy)
    phis = setup["phis"][np.asarray(probes)]
    t0 = time.time()
    vsum, vcnt = ref_shim.run_slices_serial(coords, setup, r, fill_bkg, smooth, phis=phis)
    iq, qx, qy, qz = ref_shim.finalize_serial(vsum, vcnt, setup["q_axis"], max_q)
    q_num = setup["q_num"]
    # the count grid is rank-1 (every slice shares the row table): store its exact factors, validated
    # against the reference's full count grid right below
    from . import giwaxs_oracle as ox
    hx, hy, vz = ox.slice_q_axes(float(phis[0]), setup["grid_size"], r)
    iz = ox.bin_indices(hx, hy, vz, setup["q_axis"])[4]
    m = np.bincount(iz, minlength=q_num).astype(np.int64)
    iz0 = int(np.argmax(m))
    H = np.round(vcnt[:, :, iz0] / m[iz0]).astype(np.int64)
    assert np.array_equal(vcnt, (H[:, :, None] * m[None, None, :]).astype(np.float64)), "count grid is not rank-1"
    pairs = np.argwhere(H > 0)[::keep_every]
    lo, hi = ref_shim_crop(setup["q_axis"], max_q)
    inside = np.all((pairs >= lo) & (pairs < hi), axis=1)
    out = dict(xyz=np.array(xyz), coords=coords, elements=np.asarray(elements), r=r, q=q, max_q=max_q, fill_bkg=fill_bkg, smooth=smooth, energy=energy,
               grid_size=setup["grid_size"], q_num=q_num, q_axis=setup["q_axis"], probe_phis=phis,
               n_phis_reference=len(setup["phis"]), H=H.astype(np.uint16), m=m.astype(np.uint16),
               pairs=pairs.astype(np.int32), vsum_pairs=vsum[pairs[:, 0], pairs[:, 1], :].astype(np.float32),
               vsum_max=vsum.max(), crop=np.array([lo, hi]), iq_pairs=(pairs[inside] - lo).astype(np.int32),
               iq_values=iq[pairs[inside, 0] - lo, pairs[inside, 1] - lo, :].astype(np.float32), iq_max=iq.max())
    if detector_pixels:
        P = detector_pixels
        psis, dphis, thetas = np.linspace(75, 90, 3), np.linspace(0, 179, 4), np.array([0.0])
        det, h, v = ref_shim.detectormaker_serial(iq, qx, qy, qz, P, max_q, (90.0, 90.0, 90.0), ("psi", "phi", "psi"),
                                                  psis, np.ones(3) / 3, dphis, np.ones(4) / 4, thetas, np.ones(1))
        out.update(det_P=P, det_psis=psis, det_phis=dphis, det_thetas=thetas, det=det.astype(np.float32),
                   det_max=det.max())
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print("%s: N=%d q_num=%d probes=%s pairs=%d (%.1fs)" % (name, setup["grid_size"], q_num, list(probes),
                                                              len(pairs), time.time() - t0))


def ref_shim_crop(axis, max_q):
    """Index range downselect_voxelgrid keeps (voxelgrids.py:36-46), from the reference function itself."""
    ref = ref_shim.load()
    _, kept, _, _ = ref.voxelgrids.downselect_voxelgrid(_Zero3(len(axis)), axis, axis, axis, max_q)
    lo = int(np.where(axis == kept[0])[0][0])
    return lo, lo + len(kept)


class _Zero3:
    """Stand-in for a q_num^3 grid that downselect_voxelgrid only slices (saves 1.5 GB)."""

    def __init__(self, n):
        self.shape = (n, n, n)

    def __getitem__(self, key):
        return self


def named():
    r = 0.3
    _named_case("config1_graphite_medium", "graphite_medium.xyz", r, 0.02, 2.0, True, 25, [0, 223, 446, 700],
                keep_every=3)
    # the template's smooth = 25 pixels is wider than this 45-pixel cluster (the blend mask is all zero and the
    # slice is the bare pedestal): the same file and grid with smooth = 3 lets the atoms through
    _named_case("config1b_graphite_medium_smooth3", "graphite_medium.xyz", r, 0.02, 2.0, True, 3, [0, 223, 446, 700],
                keep_every=3)
    _named_case("config2_silicon_medium", "silicon_medium.xyz", r, 0.01, 2.0, False, 0, [0, 600, 1337],
                detector_pixels=512, keep_every=8)
    _named_case("config3_graphite_large", "graphite_large.xyz", r, 2 * np.pi / (r * 1023.5), 2.0, True, 25,
                [0, 145, 435, 869], keep_every=3)


def write_xyz(path, coords, elements):
    """XYZ text that load_xyz parses back to the very same doubles."""
    with open(path, "w") as fh:
        fh.write("%d\ncluster\n" % len(coords))
        for el, (x, y, z) in zip(elements, coords):
            fh.write("%s %.17g %.17g %.17g\n" % (el, x, y, z))


def twostep():
    import tempfile
    ref = ref_shim.load()
    r, q, max_q, energy, fill_bkg, smooth = 0.3, 0.1, 1.5, 12700.0, True, 3
    out = dict(r=r, q=q, max_q=max_q, energy=energy, fill_bkg=fill_bkg, smooth=smooth)
    tmp = tempfile.mkdtemp()
    paths = []
    for k, seed in enumerate((21, 22)):
        rng = np.random.default_rng(seed)
        coords = rng.random((400, 3)) * np.array([30.0, 22.0, 26.0])
        elements = rng.choice(np.array(["C", "H", "S", "O", "F"]), size=400, p=[0.5, 0.3, 0.08, 0.07, 0.05])
        out["coords_%d" % k], out["elements_%d" % k] = coords, elements
        paths.append(os.path.join(tmp, "c%d.xyz" % k))
        write_xyz(paths[-1], coords, elements)
    t0 = time.time()
    total = None
    for k, path in enumerate(paths):
        iq, qx, qy, qz = ref_shim.voxel_grid_low_mem_serial(path, r, q, max_q, 1, energy, fill_bkg, smooth)
        if k == 0:
            out["iq_full_aff1"], out["axis"] = iq.copy(), qx.copy()
        small, qxs, qys, qzs = ref.voxelgrids.downselect_voxelgrid(iq, qx, qy, qz, max_q)
        total = small.copy() if k == 0 else total + small
    total /= len(paths)
    element = ref.utilities.most_common_element(paths[0])
    out["two_step_iq"] = ref.voxelgrids.add_f0_q_3d(total, qxs, qys, qzs, element)
    out["two_step_axis"], out["element"] = qxs.copy(), np.array(element)
    out["iq_full_aff3"] = ref_shim.voxel_grid_low_mem_serial(paths[0], r, q, max_q, 3, energy, fill_bkg, smooth)[0]
    print("twostep: q_num %d, crop %d, %.1f s" % (len(out["axis"]), len(qxs), time.time() - t0))
    np.savez_compressed(os.path.join(OUT, "twostep.npz"), **out)


def main(argv):
    os.makedirs(OUT, exist_ok=True)
    todo = argv or ["graphite262", "silicon256", "clipped128", "detector", "twostep", "named", "pm6"]
    if "twostep" in todo:
        twostep()
    iq = q = None
    if "graphite262" in todo:
        iq, q = graphite262()
    if "silicon256" in todo:
        silicon256()
    if "clipped128" in todo:
        clipped128()
    if "detector" in todo:
        if iq is None:
            g = np.load(os.path.join(OUT, "graphite262.npz"))
            iq, q = g["iq"], g["q_crop"]
        detector(iq, q)
    if "named" in todo:
        named()
    if "pm6" in todo:
        pm6()


if __name__ == "__main__":
    main(sys.argv[1:])
