"""Benchmark of the reciprocal-space hot path on BASELINE.json configs[4]
(synthetic ~10 M-atom slab, 4096^2 real-space grid, 1800 phi slices, 2048^2
detector over 360 psi orientations).

    python bench.py --gpus 1 --steps 3 --warmup 3            # our CUDA path
    python bench.py --impl reference ...                      # CPU port of the reference (oracle/)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

One step = the whole job: stage A (all phi slices -> I(q) voxel grid) followed
by stage B (all detector orientations -> image).  With N ranks the slices and
orientations are sharded round-robin and the partial grids / images are
all-reduced (NCCL), so total work is fixed: "scaling": "strong".

Printed JSON (rank 0, one line):
  value        phi-slices/s of stage A, inputs resident in HBM (CUDA events, max over ranks)
  detector     orientations/s of stage B, same rules
  e2e          the same metrics through the public drop-in calls
               (voxelgridmaker_fitting / detectormaker_fitting) from HOST arrays,
               host<->device copies and host-side preparation inside the timed region
  roofline     dominant kernel: algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json
  cpu_baseline the oracle port on the box's host cores on a bounded sample (N=1 only)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "phi-slices/sec (FFT+3D bin)"
UNIT = "slices/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--atoms", type=float, default=10_000_000)
    ap.add_argument("--phis", type=int, default=1800)
    ap.add_argument("--orientations", type=int, default=360)
    ap.add_argument("--pixels", type=int, default=2048)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-slices", type=int, default=0, help="CPU sample size (default: host cores)")
    return ap.parse_args()


def workload(args):
    from giwaxsim_b200 import synth
    cfg = synth.config5()
    cfg["n_atoms"] = int(args.atoms)
    cfg["n_phi"] = int(args.phis)
    cfg["num_pixels"] = int(args.pixels)
    cfg["psis"] = np.linspace(0, 89.75, int(args.orientations))
    cfg["phi_list"] = np.linspace(0, 180 - 180 / cfg["n_phi"], cfg["n_phi"])
    coords, elements = synth.random_slab(cfg["n_atoms"], cfg["box"])
    return cfg, coords, elements


def config_dict(cfg, world):
    return {"workload": "BASELINE configs[4]: synthetic random-atom slab, fill_bkg=True, smooth=25",
            "atoms": cfg["n_atoms"], "grid": cfg["grid_size"], "phi_slices": cfg["n_phi"], "q_num": 569,
            "detector_pixels": cfg["num_pixels"], "orientations": int(len(cfg["psis"])),
            "parallelism": "phi/psi sharded over %d rank(s), all-reduce of partial grids" % world,
            "l2": "inputs larger than L2 (atoms 170 MB, each slice grid 134 MB); no flush needed"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def start(self):
        """Launch nvidia-smi early (it needs ~0.1 s to produce its first row); rows are time-stamped
        and only those inside [mark_begin, mark_end] are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        good = [(t, r) for t, r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        rows = [r for t, r in good if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e300)]
        window = "timed region"
        if not rows:
            # a timed region shorter than the sampling period: report the rows of warm-up + timed region
            rows, window = [r for t, r in good], "warm-up + timed region (timed region shorter than one sample)"
        sm = [float(r[1]) for r in rows]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[4 + k] == "Active" for r in rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": reasons, "window": window}


# --------------------------------------------------------------------------- CPU arm
def cpu_stage_rates(cfg, coords, elements, n_slices, n_orient, threads):
    """Oracle port (kind "port") on a bounded sample: slices/s and orientations/s."""
    from giwaxsim_b200 import synth
    from oracle import giwaxs_oracle as ox
    f = ox.f_values_for(elements, table=synth.fixed_f1f2)
    phis = cfg["phi_list"][:: max(1, cfg["n_phi"] // n_slices)][:n_slices]
    t0 = time.perf_counter()
    iq, qx, qy, qz, *_ = ox.voxelgridmaker(coords, f, cfg["r_voxel_size"], cfg["q_voxel_size"], cfg["max_q"],
                                           cfg["fill_bkg"], cfg["smooth"], phis=phis, threads=threads)
    ta = time.perf_counter() - t0
    psis = cfg["psis"][:: max(1, len(cfg["psis"]) // n_orient)][:n_orient]
    t0 = time.perf_counter()
    ox.detectormaker(iq, qx, qy, qz, cfg["num_pixels"], cfg["max_q"], cfg["angle_init_vals"], cfg["angle_init_axs"],
                     psis, np.ones_like(psis) / len(psis), cfg["phis"], np.ones(1), cfg["thetas"], np.ones(1),
                     threads=threads)
    tb = time.perf_counter() - t0
    return len(phis) / ta, len(psis) / tb, len(phis), len(psis)


def cpu_threads():
    """Host threads of the CPU arm: all cores, capped at 32 (the reference's own pool is
    min(32, cores + 4) workers, comparison.py:759; one N = 4096 slice needs ~1.5 GB of scratch)."""
    return max(1, min(32, os.cpu_count() or 1))


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, coords, elements = workload(args)
    threads = cpu_threads()
    n = args.cpu_slices or threads
    rates = []
    for _ in range(max(1, args.warmup > 0) + args.steps):     # one warm-up pass is enough on the CPU
        rates.append(cpu_stage_rates(cfg, coords, elements, n, n, threads))
    rates = rates[1:] if len(rates) > 1 else rates
    sa = float(np.mean([r[0] for r in rates]))
    sb = float(np.mean([r[1] for r in rates]))
    sample = "%d phi slices and %d orientations of the full-size workload per step" % (rates[0][2], rates[0][3])
    line = {"impl": "reference", "metric": METRIC, "value": sa, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * rates[0][2] / sa,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(cfg, 1),
            "detector": {"metric": "detector orientations/sec", "value": sb, "unit": "orientations/s"},
            "cpu_baseline": {"value": sa, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                             "note": "oracle/giwaxs_oracle.py: NumPy restatement of the reference, bit-identical "
                                     "to it; the reference itself is Python and /root/reference is absent here"},
            "e2e": {"value": sa, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from giwaxsim_b200 import _lib, engine, parallel, synth
    from giwaxsim_b200.tools import comparison, utilities

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = engine.resolve_device()
    utilities.set_f1f2_provider(synth.fixed_f1f2)
    cfg, coords, elements = workload(args)
    r, q, max_q = cfg["r_voxel_size"], cfg["q_voxel_size"], cfg["max_q"]
    phis_all = cfg["phi_list"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident inputs (not timed): atoms sorted on the device, plans, base detector grids
    codes, uniq, table = comparison.species_table(elements, cfg["energy"])
    atoms = engine.AtomSet(coords, r, cfg["grid_size"], dev, species=codes, table=table)
    N, q_num, q_axis, _ = engine.stage_a_geometry(atoms.bounds, r, q, max_q)
    sum_f = np.sum(np.bincount(codes, minlength=len(table)) * np.asarray(table))
    avg_f = (sum_f / np.prod(atoms.bounds)) * r ** 3
    window = engine.crop_range(q_axis, max_q)          # as voxelgridmaker_fitting: accumulate the kept voxels only
    eng = engine.SliceEngine(None, r, q_axis, N, avg_f, atoms.bounds[0], atoms.bounds[1], cfg["fill_bkg"],
                             cfg["smooth"], device=dev, atoms=atoms, window=window)
    my_phis = parallel.shard(phis_all, rank, world)
    P = cfg["num_pixels"]
    gx, gy, gz, det_h, det_v = comparison.detector_base_device(P, max_q, cfg["angle_init_vals"],
                                                               cfg["angle_init_axs"], dev)
    psis = cfg["psis"]
    R, w = engine.orientation_tables(engine.grid_corners(gx, gy, gz), psis, np.ones_like(psis) / len(psis),
                                     cfg["phis"], np.ones(1), cfg["thetas"], np.ones(1))
    sel = parallel.shard(np.arange(len(w)), rank, world)
    R_my, w_my = np.ascontiguousarray(R[sel]), np.ascontiguousarray(w[sel])
    image = torch.zeros(P * P, dtype=torch.float64, device=dev)
    state = {}

    def stage_a():
        eng.vsum.zero_()
        eng.count2.zero_()
        eng.run(my_phis)
        if world > 1:
            parallel.all_reduce_sum([eng.vsum, eng.count2])
        state["iq"], state["axis"] = engine.finalize_voxels(eng.vsum, None, eng.count2, eng.row_hist, q_axis, max_q,
                                                            dev, window=window)

    def stage_b():
        image.zero_()
        det = engine.DetectorEngine(state["iq"], state["axis"], state["axis"], state["axis"], device=dev)
        if len(w_my):
            det.accumulate(gx, gy, gz, R_my, w_my, image=image)
        if world > 1:
            parallel.all_reduce_sum([image])
        state["det"] = engine.detector_epilogue(image, P, P, True, dev, finish=True)

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(args.warmup):
        stage_a()
        stage_b()
    barrier()
    eng.timers = {}
    _lib.reset_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    barrier()
    sampler.mark_begin()
    ev[0].record()
    for s in range(args.steps):
        stage_a()
        ev[2 * s + 1].record()
        stage_b()
        ev[2 * s + 2].record()
    barrier()
    sampler.mark_end()
    launches = _lib.launch_count()
    clocks = sampler.stop()
    ms_a = sum(ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps)) / args.steps
    ms_b = sum(ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps)) / args.steps
    kernel_ms = eng.collect_timers()
    eng.timers = None
    t = torch.tensor([ms_a, ms_b], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_a, ms_b = float(t[0]), float(t[1])

    # ---- end to end through the public drop-in calls, host arrays in, host arrays out
    e2e = None
    if not args.no_e2e:
        ones = lambda a: None
        ta = tb = 0.0
        n_e2e = max(1, min(args.steps, 2))
        n_warm = 2          # steady state: pinned staging, FFT plans and the result segments exist after two calls
        for it in range(n_warm + n_e2e):
            barrier()
            t0 = time.perf_counter()
            iq, qx, qy, qz = comparison.voxelgridmaker_fitting(coords, elements, r, q, max_q, cfg["energy"],
                                                               fill_bkg=cfg["fill_bkg"], smooth=cfg["smooth"],
                                                               phis=phis_all)
            barrier()
            t1 = time.perf_counter()
            det_sum, _, _ = comparison.detectormaker_fitting(iq, qx, qy, qz, P, max_q, cfg["angle_init_vals"],
                                                             cfg["angle_init_axs"], psis, None, cfg["phis"], None,
                                                             cfg["thetas"], None, mirror=True)
            barrier()
            t2 = time.perf_counter()
            if it >= n_warm:
                ta += t1 - t0
                tb += t2 - t1
        tt = torch.tensor([ta / n_e2e, tb / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ta, tb = float(tt[0]), float(tt[1])
        h2d = coords.nbytes + len(elements) + 27 * 8 * len(w)
        d2h = iq.size * 4 + det_sum.size * 8
        e2e = {"value": len(phis_all) / ta, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "seconds_stage_a": ta, "seconds_stage_b": tb,
               "detector_value": len(w) / tb, "detector_unit": "orientations/s",
               "api": "tools.comparison.voxelgridmaker_fitting + detectormaker_fitting, host NumPy in/out"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant stage-A kernel (SURVEY 8(d) algorithmic bytes per slice)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    A = cfg["n_atoms"]
    kr, kc = 567, 636
    alg = {"prepare": 16.0 * A / 256, "project": 28.0 * A + 8.0 * N * N, "fft2": 12.0 * N * N,
           "bin": 20.0 * kr * kc,
           # fused = F1 + F2: the whole slice: atoms + grid write + FFT read/write + binning
           "fused": 28.0 * A + 8.0 * N * N + 12.0 * N * N + 20.0 * kr * kc}
    n_my = len(my_phis)
    top = max(kernel_ms, key=kernel_ms.get) if kernel_ms else None
    roofline = None
    if top is not None:
        per_slice_ms = kernel_ms[top] / (args.steps * n_my)
        ach = alg.get(top, 0.0) / (per_slice_ms * 1e-3) / 1e9
        roofline = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": None,
                    "peak_source": "MEASURED_PEAKS.json (burst copy)" if peaks else "fallback 6650",
                    "algorithmic_bytes_per_slice": alg.get(top),
                    "kernel_ms_per_slice": {k: v / (args.steps * n_my) for k, v in kernel_ms.items()},
                    "whole_slice": {"algorithmic_bytes": 28.0 * A + 20.0 * N * N + 20.0 * kr * kc,
                                    "achieved": (28.0 * A + 20.0 * N * N + 20.0 * kr * kc) * len(phis_all)
                                    / (ms_a * 1e-3) / 1e9 / world}}
        roofline["whole_slice"]["frac"] = roofline["whole_slice"]["achieved"] / peak
        # achieved / traffic are per launch: one fused launch pair covers a batch of rotations
        per_launch = eng.fused_batch_size() if top == "fused" else eng.batch_size()
        roofline["slices_per_launch"] = per_launch
        roofline["algorithmic_bytes_per_launch"] = alg.get(top, 0.0) * per_launch
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r03_traffic.json"))).get(top)
            if tr:
                roofline["traffic"] = tr["dram_bytes_per_slice"] * per_launch
                roofline["traffic_source"] = tr["source"]
                # what the memory system really carries: DRAM bytes (ncu) over the measured kernel time
                roofline["dram_achieved"] = tr["dram_bytes_per_slice"] / (per_slice_ms * 1e-3) / 1e9
                roofline["dram_frac"] = roofline["dram_achieved"] / peak
        except Exception:
            pass
        roofline["note"] = ("fused = slice_rows_fused + slice_cols_fused; algorithmic bytes are those of the scatter, "
                            "2-D FFT and binning kernels they replace (SURVEY 8(d)), so frac > 1 means the fused pair "
                            "is faster than ANY implementation that moves those bytes through HBM; measured DRAM "
                            "traffic is ~21x lower (dram_frac) because the N x N grid and image never reach HBM - "
                            "the kernels are bound by instruction issue (profiles/r03_summary.md)")
    det_bytes = 4.0 * P * P
    det_ach = det_bytes * len(w) / world / (ms_b * 1e-3) / 1e9
    det_traffic = None
    try:
        det_traffic = json.load(open(os.path.join(ROOT, "profiles", "r03_traffic.json")))["detector_affine"]
    except Exception:
        pass
    line = {"metric": METRIC, "value": len(phis_all) / (ms_a * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_a + ms_b,
            "ms_per_step_stage_a": ms_a, "ms_per_step_stage_b": ms_b,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (indices f64)",
            "data": "synthetic", "config": config_dict(cfg, world),
            "detector": {"metric": "detector orientations/sec", "value": len(w) / (ms_b * 1e-3),
                         "unit": "orientations/s",
                         "roofline": {"bound": "hbm", "achieved": det_ach, "peak": peak, "unit": "GB/s",
                                      "frac": det_ach / peak,
                                      "traffic": (det_traffic["dram_bytes_per_orientation"] * len(w_my)
                                                  if det_traffic else None),
                                      "algorithmic_bytes_per_orientation": det_bytes,
                                      "note": "whole stage-B step (host model, gather kernel, mirror epilogue) per "
                                              "rank; the gather kernel alone: profiles/r03_summary.md"}},
            "roofline": roofline, "clocks": clocks, "gpu_launches": launches}
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        threads = cpu_threads()
        n = args.cpu_slices or threads
        sa, sb, na, nb = cpu_stage_rates(cfg, coords, elements, n, n, threads)
        line["cpu_baseline"] = {"value": sa, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d phi slices and %d orientations of the same workload" % (na, nb),
                                "detector_value": sb, "detector_unit": "orientations/s"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else was diverted to stderr."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


# Libraries write banners to fd 1 (NCCL prints "NCCL version ..." on the first communicator): keep
# the process's stdout for the single JSON line and send every other byte to stderr.
sys.stdout.flush()
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)

if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
