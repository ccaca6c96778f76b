"""CPU-side checks: the C-ABI library loads and exports what the header
declares, the product refuses to run without a GPU, the host-evaluated scalars
match the oracle, and the device FFT engine (compiled for the host) matches
numpy.fft.  No kernel is launched here."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest
import torch

from giwaxsim_b200 import _lib, engine
from oracle import giwaxs_oracle as ox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "giwaxs_b200.h")).read()
    declared = set(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(", header))
    declared.discard("gx_float2")
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "libgiwaxs_b200.so does not export %s" % name
    assert declared == set(_lib.exported_symbols())
    assert _lib.cdll().gx_abi_version() == _lib.ABI_VERSION


@pytest.mark.skipif(torch.cuda.is_available(), reason="a GPU is present")
def test_product_fails_loudly_without_gpu():
    with pytest.raises(_lib.GxError) as e:
        engine.resolve_device()
    assert e.value.code == _lib.GX_ERR_NO_DEVICE
    from giwaxsim_b200.tools.comparison import voxelgridmaker_fitting
    with pytest.raises(_lib.GxError):
        voxelgridmaker_fitting(np.random.rand(10, 3) * 10, np.array(["C"] * 10), 0.3, 0.1, 1.0, 12700.0)


def test_unsupported_fft_size_is_an_error():
    assert _lib.cdll().gx_fft_plan_bytes(8193) == _lib.GX_ERR_UNSUPPORTED      # non-pow2 above 8192
    assert _lib.cdll().gx_fft_plan_bytes(32768) == _lib.GX_ERR_UNSUPPORTED
    assert "unsupported" in _lib.last_error()
    assert _lib.cdll().gx_fft_plan_bytes(8) == _lib.GX_ERR_UNSUPPORTED


@pytest.fixture(scope="module")
def fft_emul(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("emul") / "libfft_emul.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so,
                           os.path.join(ROOT, "tests", "host_emul", "fft_emul.cpp")])
    return ctypes.CDLL(so)


@pytest.mark.parametrize("N", [16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 17, 131, 262, 524, 1048, 2095, 4095,
                               4189, 8192 - 3])
def test_fft_engine_host_build_matches_numpy(fft_emul, N):
    nbytes = _lib.cdll().gx_fft_plan_bytes(N)
    assert nbytes >= 0          # N = 16 is a single radix-16 pass: empty table
    plan = np.zeros(max(nbytes // 4, 2), np.float32)
    _lib.call("gx_fft_plan_fill", N, _lib.ptr(plan))
    rng = np.random.default_rng(N)
    x = (rng.normal(size=N) + 1j * rng.normal(size=N) + 8.0).astype(np.complex64)
    out = np.zeros(N, np.complex64)
    assert fft_emul.emul_dft(N, _lib.ptr(plan), _lib.ptr(x), _lib.ptr(out)) == 0
    ref = np.fft.fft(x.astype(np.complex128))
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.parametrize("N,klo,khi,full", [(4096, -288, 288, 0), (4096, -512, 512, 0), (4096, -200, 203, 0),
                                            (4096, -2048, 2048, 1), (4096, -257, 300, 0), (2048, -256, 256, 0),
                                            (2048, -198, 199, 0), (2048, -1024, 1024, 1), (1024, -128, 128, 0),
                                            (1024, -70, 75, 0), (1024, -512, 512, 1)])
def test_split_transform_of_the_tma_column_kernel_matches_numpy(fft_emul, N, klo, khi, full):
    """N = 16 boxes x N/16 (gx_split_*; N = 1024, 2048, 4096): rows permuted as the row kernel writes them, two
    in-box DIF passes, one DIT pass across the boxes with only the kept outputs - equals numpy.fft on the band."""
    plan = np.zeros(_lib.cdll().gx_fft_plan_bytes(N) // 4, np.float32)
    _lib.call("gx_fft_plan_fill", N, _lib.ptr(plan))
    rng = np.random.default_rng(abs(klo) * 7 + khi + N)
    x = (rng.normal(size=N) + 1j * rng.normal(size=N) + 3.0).astype(np.complex64)
    out = np.zeros(khi - klo, np.complex64)
    assert fft_emul.emul_dft_split(N, _lib.ptr(plan), _lib.ptr(x), _lib.ptr(out), klo, khi, full) == 0
    ref = np.fft.fft(x.astype(np.complex128))
    want = np.array([ref[k % N] for k in range(klo, khi)])
    assert np.abs(out - want).max() <= 3e-6 * np.abs(ref).max()


@pytest.mark.parametrize("klo,khi,nthreads", [(-288, 288, 256), (-200, 203, 256), (-512, 512, 256), (-1, 1, 64),
                                              (-300, 17, 1), (-256, 256, 256), (-257, 257, 512)])
def test_fft_lowband_last_pass_matches_numpy(fft_emul, klo, khi, nthreads):
    """Band-limited last pass of the 4096-point transform (only X[0], X[15] (+ X[1], X[14]) of every
    last-pass butterfly): every coefficient of the band equals numpy's."""
    N = 4096
    plan = np.zeros(_lib.cdll().gx_fft_plan_bytes(N) // 4, np.float32)
    _lib.call("gx_fft_plan_fill", N, _lib.ptr(plan))
    rng = np.random.default_rng(khi - klo)
    x = (rng.normal(size=N) + 1j * rng.normal(size=N) + 3.0).astype(np.complex64)
    out = np.zeros(khi - klo, np.complex64)
    assert fft_emul.emul_dft_lowband(_lib.ptr(plan), _lib.ptr(x), _lib.ptr(out), klo, khi, nthreads) == 0
    ref = np.fft.fft(x.astype(np.complex128))[np.arange(klo, khi) % N]
    assert np.abs(out - ref).max() <= 2e-6 * np.abs(ref).max()
    assert fft_emul.emul_dft_lowband(_lib.ptr(plan), _lib.ptr(x), _lib.ptr(out), -513, 10, 256) == -1


def test_orientation_matrices_match_oracle_chain():
    gx, gy, gz, _, _ = ox.detector_base(33, 2.0, (90.0, 90.0, 90.0), ("psi", "phi", "psi"))
    psis, phis, thetas = np.linspace(75, 90, 4), np.linspace(0, 179, 5), np.linspace(0, 1, 2)
    w = [np.ones_like(a) / len(a) for a in (psis, phis, thetas)]
    R, wt = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, w[0], phis, w[1], thetas, w[2])
    todo = ox.orientation_list(psis, w[0], phis, w[1], thetas, w[2])
    assert len(todo) == R.shape[0]
    for o, (psi, phi, theta, weight) in enumerate(todo):
        g = (gx, gy, gz)
        for step, (which, ang) in enumerate((("psi", psi), ("phi", phi), ("theta", theta))):
            M = ox.axis_angle_matrix(ox.detector_axis(*g, which), np.radians(ang))
            assert np.array_equal(M.ravel(), R[o, step]), (o, step)
            g = ox.matvec3(M, *g)
        assert wt[o] == weight


def test_encode_values():
    el = np.array(["C", "H", "C", "S", "H", "C"])
    codes, uniq = engine.encode_values(el)
    assert [str(u) for u in uniq] == ["C", "H", "S"] and codes.tolist() == [0, 1, 0, 2, 1, 0]
    many = np.arange(40).astype(complex)
    assert engine.encode_values(many) == (None, None)


def test_stage_a_geometry_matches_oracle():
    rng = np.random.default_rng(0)
    coords = rng.random((50, 3)) * [30, 20, 25]
    f = np.full(50, 6.0 + 0j)
    for r, q, mq in [(0.3, 0.02, 2.0), (0.3, 0.01, 2.0), (0.25, 0.05, 1.0)]:
        s = ox.stage_a_setup(coords, f, r, q, mq)
        bounds = (s["x_bound"], s["y_bound"], s["z_bound"])
        N, q_num, q_axis, phis = engine.stage_a_geometry(bounds, r, q, mq)
        assert (N, q_num) == (s["grid_size"], s["q_num"])
        assert np.array_equal(q_axis, s["q_axis"]) and np.array_equal(phis, s["phis"])
    with pytest.raises(Exception, match="non-physical"):
        engine.stage_a_geometry((1, 1, 1), 3.0, 0.1, 2.0)
    with pytest.raises(Exception, match="smaller than simulation"):
        engine.stage_a_geometry((500, 500, 500), 0.3, 0.1, 2.0)


def _chord_from_constants(c, x):
    """NumPy transcription of the device formula (gx_project.cu: chord_length)."""
    if c.mode == 0:
        return np.where(x < c.hor, c.ver, 0.0)
    if c.mode == 1:
        return np.where(x < c.ver, c.hor, 0.0)
    with np.errstate(all="ignore"):
        first = (c.rise + x * c.tan_phi) - (c.vcos - x) * c.tan_theta
        last = (c.hor - ((x - c.vcos) / c.cos_phi)) / c.cos_theta
    out = np.zeros_like(x)
    m0 = x == 0
    m1 = ~m0 & (x < c.stop1)
    m2 = ~m0 & ~m1 & (x <= c.stop2)
    m3 = ~m0 & ~m1 & ~m2 & (x < c.stop12)
    out[m1] = first[m1]
    out[m2] = c.mid
    out[m3] = last[m3]
    return out


def test_chord_constants_reproduce_oracle_lengths():
    x = np.arange(400) * 0.3
    phis = np.array([0.0, 0.2, 30.0, 45.0, 77.7, 90.0, 90.2, 120.0, 179.8])
    import types
    arr = engine.chord_constants(phis, 21.0, 14.0)
    assert arr.dtype.itemsize == ctypes.sizeof(_lib.Chord)
    for i, phi in enumerate(phis):
        c = types.SimpleNamespace(**{k: arr[i][k] for k in arr.dtype.names})
        assert np.array_equal(_chord_from_constants(c, x), ox.chord_lengths(x, 21.0, 14.0, phi)), phi


def test_gaussian_weights_match_scipy():
    from scipy.ndimage import gaussian_filter1d
    for sigma in (1, 5, 25):
        w, radius = engine.gaussian_weights(sigma)
        n = 4 * radius + 3
        delta = np.zeros(n)
        delta[n // 2] = 1.0
        k = gaussian_filter1d(delta, sigma=sigma, mode="wrap")
        assert np.array_equal(k[n // 2 - radius:n // 2 + radius + 1], w)
        assert radius == int(4 * sigma + 0.5)


def test_fft_pass_address_identity(fft_emul):
    """gx_phys(base + S*n) == gx_phys(base) + gx_phys(S*n) for every butterfly of every
    pass of every schedule (the kernels rely on it for immediate-offset addressing)."""
    assert fft_emul.emul_offsets_ok() == 1


def test_structure_loaders_and_cell_vectors_match_oracle(tmp_path):
    from giwaxsim_b200.tools import utilities
    xyz = tmp_path / "m.xyz"
    xyz.write_text("3\ncomment\nC1 0.0 1.5 2.25\nSi12 -1.0 2.0 3.0 extra\nbroken line\nH 1e-3 2 x\nO 4 5 6\n")
    pdb = tmp_path / "m.pdb"
    pdb.write_text("HEADER\n"
                   "ATOM      1  C1  MOL A   1      11.104   6.134  -6.504  1.00  0.00           C\n"
                   "HETATM    2 CL   MOL A   1       1.000  -2.000   3.500  1.00  0.00          CL\n"
                   "TER\nEND\n")
    for path, fn in ((str(xyz), utilities.load_xyz), (str(pdb), utilities.load_pdb)):
        c, e = fn(path)
        oc, oe = ox.read_structure(path)
        assert np.array_equal(c, oc) and np.array_equal(e, oe) and len(e) > 0
    for cell in [(4.0, 5.0, 6.0, 90.0, 90.0, 90.0), (2.456, 4.254, 6.696, 90.0, 90.0, 120.0), (7.0, 8.0, 9.0, 75.0, 85.0, 95.0)]:
        for u, v in zip(utilities.calc_real_space_abc(*cell), ox.cell_vectors(*cell)):
            assert np.array_equal(u, v)


def test_config_parser_and_defaults(tmp_path):
    from giwaxsim_b200 import simulate
    from giwaxsim_b200.tools import utilities
    cfg = tmp_path / "c.txt"
    cfg.write_text("# a comment without the sign\ninput_filepath=cell.pdb\nmax_q=2\nq_voxel_size=0.02\nfill_bkg=True\n"
                   "smooth=25\nmirror=yes\npsi_start=0\npsi_end=90\npsi_num=16\nphi_start=0\nphi_end=179\nphi_num=180\n"
                   "theta_start=0\ntheta_end=0\ntheta_num=1\nangle_init_ax1=psi\nangle_init_val1=90\nnote=a=b\n")
    config = utilities.parse_config_file(str(cfg))
    assert config["note"] == "a=b" and "# a comment without the sign" not in config
    s = simulate.read_settings(config)
    assert s["num_pixels"] == 100 and s["r_voxel_size"] == 0.3 and s["energy"] == 10000.0     # reference defaults
    assert s["fill_bkg"] is True and s["mirror"] is True and s["smooth"] == 25
    assert s["angle_init_vals"] == (90.0, 0.0, 0.0) and s["angle_init_axs"] == ("psi", "None", "None")
    assert np.array_equal(s["psis"], np.linspace(0.0, 90.0, 16)) and len(s["phis"]) == 180 and len(s["thetas"]) == 1
    assert utilities.str_to_bool(" ON ") is True and utilities.str_to_bool("nope") is False
    with pytest.raises(ValueError):
        utilities.str_to_bool("nope", default=None)
    out = tmp_path / "echo.txt"
    utilities.save_config_to_txt(config, str(out))
    assert utilities.parse_config_file(str(out)) == config
    pdb = tmp_path / "cell.pdb"
    pdb.write_text("CRYST1   44.456   45.726   40.097  90.00  90.00  90.00 P 1           1\n")
    assert utilities.load_pdb_cell_params(str(pdb)) == (44.456, 45.726, 40.097, 90.0, 90.0, 90.0)
    with pytest.raises(Exception, match="Either input_folder or input_path"):
        simulate.main({"psi_start": "0", "psi_end": "1", "psi_num": "1", "phi_start": "0", "phi_end": "1",
                       "phi_num": "1", "theta_start": "0", "theta_end": "1", "theta_num": "1"})


def test_ctypes_prototypes_have_the_headers_arity():
    """Every binding in _lib._PROTOTYPES passes exactly as many arguments as the header declares."""
    header = open(os.path.join(ROOT, "include", "giwaxs_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", " ", header, flags=re.S)
    decls = dict(re.findall(r"\b(gx_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", header, flags=re.S))
    assert set(decls) == set(_lib._PROTOTYPES)
    for name, params in decls.items():
        params = params.strip()
        n = 0 if params in ("", "void") else params.count(",") + 1
        assert n == len(_lib._PROTOTYPES[name][1]), (name, n, len(_lib._PROTOTYPES[name][1]))
    # the two argument blocks mirror the C structs field for field
    for cname, pyname in (("gx_fused_args", _lib.FusedArgs), ("gx_slab_args", _lib.SlabArgs), ("gx_chord", _lib.Chord)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, flags=re.S).group(1)
        c_fields = []
        for stmt in body.split(";"):
            stmt = stmt.strip()
            if not stmt:
                continue
            names = re.sub(r"^(const\s+)?(unsigned\s+)?[A-Za-z_0-9]+\s+", "", stmt)      # drop the type
            c_fields += [re.sub(r"[\s\*]|\[.*?\]", "", x) for x in names.split(",")]
        assert c_fields == [f[0] for f in pyname._fields_], cname


def test_two_step_host_helpers(tmp_path):
    """Host pieces of the two-step command line: structure dispatch by extension, most common element
    with Counter's tie rule (first seen wins), f0(q) table and its registration hook."""
    from collections import Counter
    from giwaxsim_b200.tools import utilities
    path = str(tmp_path / "tie.xyz")
    with open(path, "w") as fh:
        fh.write("6\ncomment\nS 0 0 0\nC1 1 0 0\nS 0 1 0\nC2 0 0 1\nO 1 1 1\nH 2 2 2\n")
    coords, el = utilities.load_structure(path)
    assert coords.shape == (6, 3) and list(el) == ["S", "C", "S", "C", "O", "H"]
    assert utilities.most_common_element(path) == Counter(list(el)).most_common(1)[0][0] == "S"
    with pytest.raises(Exception, match="must be a .pdb or .xyz"):
        utilities.load_structure(str(tmp_path / "x.cif"))
    f0 = utilities.get_element_f0_dict(0.0, ["C", "H", "C"])
    assert set(f0) == {"C", "H"}
    assert f0["C"] == pytest.approx(sum(utilities.CROMER_MANN["C"][0:8:2]) + utilities.CROMER_MANN["C"][8])
    assert abs(f0["C"] - 6.0) < 0.01 and abs(f0["H"] - 1.0) < 0.01          # f0(0) = Z
    a = utilities.get_element_f0_dict(1.3, ["S"])["S"]
    c = utilities.CROMER_MANN["S"]
    # the reference's exponent uses q, not q^2 (utilities.py:331-335)
    want = sum(c[2 * i] * np.exp(-c[2 * i + 1] * 1.3 / (16 * np.pi ** 2)) for i in range(4)) + c[8]
    assert a == pytest.approx(want, rel=1e-15)
    with pytest.raises(KeyError):
        utilities.get_element_f0_dict(0.5, ["Xx"])
    utilities.register_cromer_mann("Xx", range(9))
    try:
        assert utilities.get_element_f0_dict(0.0, ["Xx"])["Xx"] == 0 + 2 + 4 + 6 + 8
        with pytest.raises(ValueError):
            utilities.register_cromer_mann("Yy", (1, 2, 3))
    finally:
        del utilities.CROMER_MANN["Xx"]


def test_two_step_config_errors_need_no_gpu(tmp_path):
    """Argument errors of the two-step drivers are raised before any device work."""
    from giwaxsim_b200 import detectormaker, voxelgridmaker
    with pytest.raises(Exception, match="Either input_folder or input_path"):
        voxelgridmaker.main({"gen_name": "x"})
    with pytest.raises(Exception, match="Path does not exist"):
        detectormaker.main({"iq_output_folder": str(tmp_path / "missing"), "gen_name": "x", "psi_start": "0",
                            "psi_end": "1", "psi_num": "2", "phi_start": "0", "phi_end": "1", "phi_num": "2",
                            "theta_start": "0", "theta_end": "0", "theta_num": "1"})


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): stdout is exactly one JSON
    line with the contract's keys; everything else goes to stderr.  Tiny sample of the real workload."""
    import json
    import sys
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-slices", "1", "--atoms", "20000", "--pixels", "128",
                        "--orientations", "2"], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = p.stdout.splitlines()
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "slices/s" and d["value"] > 0 and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "slices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("BASELINE configs[4]") and d["gpu_launches"] == 0


def test_host_widening_copy_matches_numpy():
    """gx_host_widen_f32_f64 (the fp32 -> float64 pass of the result download, run on the library's host thread
    pool with streaming stores): every element equals NumPy's cast, whatever the alignment, the length and the
    number of threads; neighbours of the destination range are untouched; errors follow the C-ABI convention."""
    import torch
    from giwaxsim_b200 import _lib
    rng = np.random.default_rng(3)
    src = rng.standard_normal(3_000_017).astype(np.float32)
    src[:4] = [0.0, -0.0, np.float32(1e-45), np.float32(3.4e38)]
    for off, n, threads in ((0, len(src), 8), (1, 1_000_003, 3), (3, 9, 4), (2, 300_000, 1), (5, 0, 2)):
        dst = np.full(n + 3, 7.0)
        _lib.call("gx_host_widen_f32_f64", src.ctypes.data + 4 * off, dst.ctypes.data + 8, n, threads)
        assert np.array_equal(dst[1:1 + n], src[off:off + n].astype(np.float64))
        assert np.array_equal(np.signbit(dst[1:1 + n]), np.signbit(src[off:off + n]))
        assert dst[0] == 7.0 and dst[n + 1] == 7.0 and dst[n + 2] == 7.0
    with pytest.raises(_lib.GxError, match="NULL pointer"):
        _lib.call("gx_host_widen_f32_f64", 0, src.ctypes.data, 4, 1)
    with pytest.raises(_lib.GxError, match="negative length"):
        _lib.call("gx_host_widen_f32_f64", src.ctypes.data, src.ctypes.data, -1, 1)


def _bench_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_binding_bound_is_the_larger_on_chip_fraction():
    """`roofline.binding_bound` names the on-chip resource with the larger measured fraction (shared-memory
    data pipe or instruction issue) and lists the other one beside it; without ncu counts there is none."""
    b = _bench_module()
    k = {"lsu_pipe_achieved": 2.5e11, "lsu_pipe_frac": 0.86, "issue_achieved": 8.1e11, "issue_frac": 0.70}
    out = b.binding_bound(k, 1.163e12, 2.908e11)
    assert out["bound"] == "shared-memory data pipe" and out["frac"] == 0.86 and out["unit"] == "wavefronts/s"
    assert out["others"] == {"issue": 0.70} and out["frac"] <= 1.0
    out = b.binding_bound({"issue_achieved": 8.1e11, "issue_frac": 0.70}, 1.163e12, 2.908e11)
    assert out["bound"] == "issue" and out["others"] == {}
    assert b.binding_bound({}, 1.163e12, 2.908e11) is None


def test_bench_clock_sampler_windows():
    """Rows are kept only inside [mark_begin, mark_end]; nvidia-smi rows win, the in-process NVML rows fill in
    when nvidia-smi produced none inside the timed region, and throttle reasons are reported by name."""
    b = _bench_module()
    row = lambda sm, cap="Not Active": ["0", str(sm), "1965", "600.0", "Not Active", "Not Active", "Not Active", cap]
    s = b.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    s.t0, s.t1 = 10.0, 20.0
    s.rows = [(5.0, row(300)), (12.0, row(1965)), (15.0, row(1950, "Active")), (25.0, row(400))]
    out = s.stop()
    assert out["sm_mhz"] == 1957.5 and out["samples"] == 2 and out["reasons"] == ["sw_power_cap"]
    assert out["window"].startswith("timed region (nvidia-smi")
    s = b.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None})()
    s.t0, s.t1 = 10.0, 20.0
    s.rows = [(5.0, row(300))]
    s.nvml_rows = [(11.0, row(1965)), (30.0, row(500))]
    out = s.stop()
    assert out["sm_mhz"] == 1965.0 and out["samples"] == 1 and "NVML" in out["window"] and out["reasons"] == []
    s = b.ClockSampler(0)
    assert s.stop()["reasons"] == ["nvidia-smi unavailable"]


def test_result_mapping_tracks_in_place_edits():
    """Large results are handed out as private copy-on-write mappings of a pooled memory file
    (parallel.shared_result_f64(single=True)): reading leaves them 'unmodified', any write - also through a
    view - is seen by the kernel's page tracking (parallel.result_unmodified), arrays the pool did not hand out
    are 'unknown' (None), and a segment is recycled only when the previous array on it is gone."""
    from giwaxsim_b200 import parallel
    import gc

    def widen(dev_slice, view):
        view[:] = dev_slice.numpy().astype(np.float64)

    t = torch.arange(300_000, dtype=torch.float32).reshape(50, 60, 100)
    a = parallel.shared_result_f64(t, widen, min_bytes=0, single=True)
    assert a.dtype == np.float64 and a.shape == (50, 60, 100) and a.flags.writeable
    assert np.array_equal(a, t.numpy().astype(np.float64))
    if parallel.result_unmodified(a) is None:
        pytest.skip("/proc/self/pagemap not readable here")
    assert parallel.result_unmodified(a) is True
    assert float(a.sum()) == float(t.double().sum())                 # reading does not count
    assert parallel.result_unmodified(a) is True
    b = parallel.shared_result_f64(t * 2, widen, min_bytes=0, single=True)       # `a` alive -> another segment
    assert len(parallel._pool[a.nbytes]) == 2 and a[1, 2, 3] == t[1, 2, 3].item()
    a[49, 59, 99] += 1.0
    assert parallel.result_unmodified(a) is False and parallel.result_unmodified(b) is True
    view = b[10:12]
    view[0, 0, 0] = -5.0                                                          # a write through a view
    assert parallel.result_unmodified(b) is False
    assert parallel.result_unmodified(np.zeros(4)) is None and parallel.result_unmodified(a[1:]) is None
    del a, b, view
    gc.collect()
    c = parallel.shared_result_f64(t * 3, widen, min_bytes=0, single=True)       # a freed segment is reused ...
    assert len(parallel._pool[c.nbytes]) == 2
    assert parallel.result_unmodified(c) is True and c[2, 2, 2] == 3 * t[2, 2, 2].item()   # ... with no stale private pages
