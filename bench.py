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
  roofline     per kernel: HBM bytes the fused design must move / CUDA-event time vs MEASURED_PEAKS.json,
               and the bound that binds (instruction issue) from the ncu counts in profiles/
  cpu_baseline the oracle port on the box's host cores on a bounded sample (N=1 only); only the
               per-slice / per-orientation loops are timed, one-off phases reported separately
  check        the run of record checks itself: count grid of the whole run == oracle bin indices,
               GPU == CPU on the sample the CPU arm computed (same 10 M-atom slab), N ranks == 1 rank
  digest       rank-count-invariant fingerprint of the result (identical at 1/2/4/8 GPUs)
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


def combine_text(world):
    from giwaxsim_b200 import parallel
    if world == 1:
        return "no collective"
    if parallel.COMBINE == "scatter":
        return "partial sums reduce-scattered, finalise sharded, iq all-gathered"
    return "partial sums all-reduced in place, finalise replicated"


def config_dict(cfg, world):
    return {"workload": "BASELINE configs[4]: synthetic random-atom slab, fill_bkg=True, smooth=25",
            "atoms": cfg["n_atoms"], "grid": cfg["grid_size"], "phi_slices": cfg["n_phi"], "q_num": 569,
            "detector_pixels": cfg["num_pixels"], "orientations": int(len(cfg["psis"])),
            "parallelism": "phi/psi sharded over %d rank(s); %s (NCCL behind the C ABI)" % (world, combine_text(world)),
            "l2": "inputs larger than L2 (atoms 170 MB, each slice grid 134 MB); no flush needed"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None
        self.nvml_rows, self._run = [], True

    def _nvml(self):
        """Second source, same counters read in-process through NVML (what nvidia-smi itself calls):
        a fresh box can need more than a second for nvidia-smi's first row, longer than a short run."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = [(nv.nvmlClocksEventReasonHwSlowdown, "Active"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "Active"),
                    (nv.nvmlClocksEventReasonSwThermalSlowdown, "Active"), (nv.nvmlClocksEventReasonSwPowerCap, "Active")]
            while self._run:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                row = [str(self.index), str(sm), str(mx), "0"] + [a if mask & b else "Not Active" for b, a in bits]
                self.nvml_rows.append((time.perf_counter(), row))
                time.sleep(0.01)
        except Exception:
            pass

    def start(self):
        """Launch nvidia-smi early (it needs ~0.1 s to produce its first row); rows are time-stamped
        and only those inside [mark_begin, mark_end] are reported."""
        threading.Thread(target=self._nvml, daemon=True).start()
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

    def _inside(self, stamped):
        good = [(t, r) for t, r in stamped if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        return good, [r for t, r in good if self.t0 is not None and self.t0 <= t <= (self.t1 or 1e300)]

    def stop(self):
        self._run = False
        if self.proc is None and not self.nvml_rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.proc is not None:
            self.proc.terminate()
        good, rows = self._inside(self.rows)
        window = "timed region (nvidia-smi -lms 20)"
        if not rows:
            good, rows = self._inside(self.nvml_rows)
            window = "timed region (NVML in-process, 10 ms; nvidia-smi produced no row inside it)"
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
def cpu_sample(cfg, coords, elements, n_slices, n_orient, threads, keep=False):
    """Oracle port (kind "port") on a bounded sample of the workload.  Only the per-unit loops
    are charged to the rates: a whole run pays the set-up (two zeroed q_num^3 grids) and the
    finalise (sum/count, crop, f0 on the whole grid) once per 1800 slices, a 16-slice sample must
    not pay them per 16 (they are reported separately).  keep=True also returns what the GPU
    result of the same sample is checked against."""
    from giwaxsim_b200 import synth
    from oracle import giwaxs_oracle as ox
    f = ox.f_values_for(elements, table=synth.fixed_f1f2)
    phis = cfg["phi_list"][:: max(1, cfg["n_phi"] // n_slices)][:n_slices]
    ta, tb = {}, {}
    iq, qx, qy, qz, vsum, vcnt, setup = ox.voxelgridmaker(
        coords, f, cfg["r_voxel_size"], cfg["q_voxel_size"], cfg["max_q"], cfg["fill_bkg"], cfg["smooth"],
        phis=phis, threads=threads, timing=ta)
    psis = cfg["psis"][:: max(1, len(cfg["psis"]) // n_orient)][:n_orient]
    w = np.ones_like(psis) / len(psis)
    det, _, _ = ox.detectormaker(iq, qx, qy, qz, cfg["num_pixels"], cfg["max_q"], cfg["angle_init_vals"],
                                 cfg["angle_init_axs"], psis, w, cfg["phis"], np.ones(1), cfg["thetas"],
                                 np.ones(1), threads=threads, timing=tb)
    out = {"slices_per_s": len(phis) / ta["slices_s"], "orient_per_s": len(psis) / tb["orientations_s"],
           "n_slices": len(phis), "n_orient": len(psis), "setup_s": ta["setup_s"], "finalize_s": ta["finalize_s"],
           "slices_s": ta["slices_s"], "orientations_s": tb["orientations_s"],
           "detector_base_s": tb["base_s"], "detector_epilogue_s": tb["epilogue_s"]}
    if keep:
        lo, hi = ox.crop_range(setup["q_axis"], cfg["max_q"])
        out.update(phis=phis, psis=psis, w=w, iq=iq, det=det,
                   vsum=vsum[lo:hi, lo:hi, lo:hi].copy(), vcnt=vcnt[lo:hi, lo:hi, lo:hi].astype(np.int64))
    return out


def cpu_threads():
    """Host threads of the CPU arm: all cores, capped at 32 (the reference's own pool is
    min(32, cores + 4) workers, comparison.py:759; one N = 4096 slice needs ~1.5 GB of scratch)."""
    return max(1, min(32, os.cpu_count() or 1))


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_baseline_dict(c, threads, n_phi):
    return {"value": c["slices_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "cpu": cpu_model(),
            "sample": "%d phi slices and %d orientations of the same workload; only the per-slice / "
                      "per-orientation loops are timed" % (c["n_slices"], c["n_orient"]),
            "detector_value": c["orient_per_s"], "detector_unit": "orientations/s",
            "setup_s": c["setup_s"], "finalize_s": c["finalize_s"],
            "whole_run_extrapolated_s": n_phi / c["slices_per_s"] + c["setup_s"] + c["finalize_s"],
            "note": "setup_s / finalize_s are the one-off phases of a whole run (two q_num^3 float64 grids, "
                    "sum/count + crop + f0); a 1800-slice run costs 1800/value + setup_s + finalize_s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg, coords, elements = workload(args)
    threads = cpu_threads()
    n = args.cpu_slices or threads
    runs = []
    for _ in range(max(1, args.warmup > 0) + args.steps):     # one warm-up pass is enough on the CPU
        runs.append(cpu_sample(cfg, coords, elements, n, n, threads))
    runs = runs[1:] if len(runs) > 1 else runs
    sa = float(np.mean([r["slices_per_s"] for r in runs]))
    sb = float(np.mean([r["orient_per_s"] for r in runs]))
    c = dict(runs[-1], slices_per_s=sa, orient_per_s=sb)
    base = cpu_baseline_dict(c, threads, cfg["n_phi"])
    base["note"] += ("; oracle/giwaxs_oracle.py: NumPy restatement of the reference, bit-identical to it; the "
                     "reference itself is Python and /root/reference is absent on the GPU box")
    line = {"impl": "reference", "metric": METRIC, "value": sa, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            # one step = the slice loop of the sample: ms_per_step / sample size = cost of one slice
            "ms_per_step": 1e3 * runs[0]["n_slices"] / sa,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_dict(cfg, 1),
            "detector": {"metric": "detector orientations/sec", "value": sb, "unit": "orientations/s"},
            "cpu_baseline": base,
            "e2e": {"value": sa, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# --------------------------------------------------------------------------- checks of the run of record
def expected_count2(cfg, q_axis, window, N, r):
    """Per-(iy, ix) number of kept slice columns over ALL rotations of the run, from the oracle's
    bin indices (process_file2, voxelgrids.py:475-499) - vectors only, so the whole 1800-slice run
    is covered in about a second on the host.  With the phi-invariant row histogram this IS the
    reference's count grid: count[iy, ix, iz] = H[iy, ix] * m[iz]."""
    from oracle import giwaxs_oracle as ox
    lo, hi = window
    V = hi - lo
    H = np.zeros((V, V), dtype=np.int64)
    for phi in cfg["phi_list"]:
        hx, hy, vz = ox.slice_q_axes(phi, N, r)
        col_mask, ix, iy, row_mask, iz = ox.bin_indices(hx, hy, vz, q_axis)
        keep = (ix >= lo) & (ix < hi) & (iy >= lo) & (iy < hi)
        np.add.at(H, (iy[keep] - lo, ix[keep] - lo), 1)
    m = np.bincount(iz[(iz >= lo) & (iz < hi)] - lo, minlength=V).astype(np.int64)
    return H, m


def digest(count2, row_hist, vsum_l1, iq, image):
    """Rank-count-invariant fingerprint of a finished run: integer parts are exact under any
    sharding, float norms agree to summation order."""
    import zlib
    c = count2.cpu().numpy().astype(np.int64)
    return {"count2_sum": int(c.sum()), "count2_crc32": int(zlib.crc32(c.tobytes())),
            "row_hist_sum": int(row_hist.sum().item()), "vsum_l1": float(vsum_l1),
            "iq_l1": float(iq.double().abs().sum().item()), "image_l1": float(image.double().abs().sum().item())}


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

    def slice_engine():
        return engine.SliceEngine(None, r, q_axis, N, avg_f, atoms.bounds[0], atoms.bounds[1], cfg["fill_bkg"],
                                  cfg["smooth"], device=dev, atoms=atoms, window=window)

    eng = slice_engine()
    my_phis = parallel.shard(phis_all, rank, world)
    P = cfg["num_pixels"]
    gx, gy, gz, det_h, det_v = comparison.detector_base_device(P, max_q, cfg["angle_init_vals"],
                                                               cfg["angle_init_axs"], dev)
    psis = cfg["psis"]
    corners = engine.grid_corners(gx, gy, gz)

    def tables(psi_list, weights):
        return engine.orientation_tables(corners, psi_list, weights, cfg["phis"], np.ones(1), cfg["thetas"],
                                         np.ones(1))

    R, w = tables(psis, np.ones_like(psis) / len(psis))
    sel = parallel.shard(np.arange(len(w)), rank, world)
    R_my, w_my = np.ascontiguousarray(R[sel]), np.ascontiguousarray(w[sel])
    image = torch.zeros(P * P, dtype=torch.float64, device=dev)
    state = {}

    def stage_a():
        eng.vsum.zero_()
        eng.count2.zero_()
        eng.vsum_is_partial = False
        eng.run(my_phis)
        state["iq"], state["axis"] = parallel.combine_and_finalize(eng, q_axis, max_q, dev, window=window)

    def detector_image(iq, axis, R_sel, w_sel, img, reduce):
        img.zero_()
        det = engine.DetectorEngine(iq, axis, axis, axis, device=dev)
        if len(w_sel):
            det.accumulate(gx, gy, gz, R_sel, w_sel, image=img)
        if reduce and world > 1:
            parallel.all_reduce_image(img, dev)
        return engine.detector_epilogue(img, P, P, True, dev, finish=True)

    def stage_b():
        state["det"] = detector_image(state["iq"], state["axis"], R_my, w_my, image, True)

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

    # ---- the run of record checks itself (not timed)
    check = {}
    if world > 1 and getattr(eng, "vsum_is_partial", False):
        # after the reduce-scatter only this rank's slab of columns holds totals
        cpr, _ = parallel.padded_columns(eng.q_out, world)
        vsum_l1 = eng.vsum_store[rank * cpr * eng.q_out:(rank + 1) * cpr * eng.q_out].double().abs().sum()
        dist.all_reduce(vsum_l1)
    else:
        vsum_l1 = eng.vsum.double().abs().sum()
    full_digest = digest(eng.count2, eng.row_hist, float(vsum_l1.item()), state["iq"], state["det"])
    if rank == 0:
        # (1) the count grid of the WHOLE run against the oracle's bin indices, bit for bit
        H, m = expected_count2(cfg, q_axis, window, N, r)
        V = window[1] - window[0]
        got_H = eng.count2.cpu().numpy().astype(np.int64).reshape(V, V)
        check["full_run_counts_equal_oracle"] = bool(np.array_equal(got_H, H) and
                                                     np.array_equal(eng.row_hist.cpu().numpy().astype(np.int64), m))
        check["full_run_voxel_samples"] = int(H.sum() * m.sum())
    if world > 1 and rank == 0:
        # (2) SURVEY T9 on hardware: the N-rank result against a 1-rank run of the same job on this GPU
        # (reference: one shared accumulator, voxelgrids.py:502-503, detector.py:298)
        solo = slice_engine()
        solo.run(phis_all)
        iq1, axis1 = engine.finalize_voxels(solo.vsum, None, solo.count2, solo.row_hist, q_axis, max_q, dev,
                                            window=window)
        img1 = torch.zeros_like(image)
        det1 = detector_image(iq1, axis1, R, w, img1, False)
        top_iq, top_det = float(iq1.abs().max().item()), float(det1.abs().max().item())
        check["multi_gpu_vs_1rank"] = {
            "ranks": world,
            "counts_equal": bool(torch.equal(solo.count2, eng.count2)),
            "iq_rel_err": float((state["iq"] - iq1).abs().max().item()) / top_iq,
            "image_rel_err": float((state["det"] - det1).abs().max().item()) / top_det,
            "digest_1rank": digest(solo.count2, solo.row_hist, float(solo.vsum.double().abs().sum().item()),
                                   iq1, det1)}
        del solo, iq1, img1, det1

    # ---- end to end through the public drop-in calls, host arrays in, host arrays out
    e2e = None
    if not args.no_e2e:
        ta = tb = 0.0
        per_call = []
        n_e2e = max(5, args.steps)
        n_warm = 3          # steady state: pinned staging, FFT plans and the pooled result segments exist
        import gc
        for it in range(n_warm + n_e2e):
            if it == n_warm:
                # like timeit: no cyclic garbage collection inside the timed calls (a full collection of this
                # process - torch, numpy and the 10 M-atom inputs loaded - is a sporadic 60 ms pause that has
                # nothing to do with the path; reference counting still frees every array at once)
                gc.collect()
                gc.disable()
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
            per_call.append((round(1e3 * (t1 - t0), 2), round(1e3 * (t2 - t1), 2)))
            if it >= n_warm:
                ta += t1 - t0
                tb += t2 - t1
        gc.enable()
        tt = torch.tensor([ta / n_e2e, tb / n_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ta, tb = float(tt[0]), float(tt[1])
        # bytes that cross PCIe per call, summed over ranks: coordinates (fp64) and element code points
        # (the '<U1'/'<U2' array as it is), orientation records; results come back in fp32
        h2d = coords.nbytes + np.asarray(elements).nbytes + 104 * len(w)
        d2h = iq.size * 4 + det_sum.size * 4
        if rank == 0:
            check["e2e_vs_resident"] = {
                "iq_rel_err": float(np.abs(iq - state["iq"].cpu().numpy()).max() / np.abs(iq).max()),
                "image_rel_err": float(np.abs(det_sum - state["det"].cpu().numpy()).max() / np.abs(det_sum).max())}
        e2e = {"value": len(phis_all) / ta, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "seconds_stage_a": ta, "seconds_stage_b": tb, "calls_averaged": n_e2e,
               "ms_per_call_incl_warmup": per_call,
               "median_ms_stage_a": float(np.median([c[0] for c in per_call[n_warm:]])),
               "gc": "cyclic GC disabled during the timed calls (as timeit does)",
               "detector_value": len(w) / tb, "detector_unit": "orientations/s",
               "api": "tools.comparison.voxelgridmaker_fitting + detectormaker_fitting, host NumPy in/out"}

    if rank != 0:
        if world > 1:
            parallel.shutdown()
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json (burst copy), of measured" if peaks else "6650 GB/s, of fallback"
    n_my = len(my_phis)
    roofline = roofline_block(cfg, eng, kernel_ms, args.steps * n_my, N, peak, peak_src, clocks)
    det_bytes = 4.0 * P * P
    det_ach = det_bytes * len(w) / world / (ms_b * 1e-3) / 1e9
    line = {"metric": METRIC, "value": len(phis_all) / (ms_a * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_a + ms_b,
            "ms_per_step_stage_a": ms_a, "ms_per_step_stage_b": ms_b,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32 (indices f64)",
            "data": "synthetic", "config": config_dict(cfg, world),
            "detector": {"metric": "detector orientations/sec", "value": len(w) / (ms_b * 1e-3),
                         "unit": "orientations/s",
                         "roofline": {"bound": "hbm", "achieved": det_ach, "peak": peak, "unit": "GB/s",
                                      "frac": det_ach / peak, "traffic": None,
                                      "algorithmic_bytes_per_orientation": det_bytes,
                                      "note": "whole stage-B step (orientation model, gather kernel, mirror "
                                              "epilogue) per rank; the gather is served by L1/L2, its DRAM "
                                              "traffic is in profiles/ (kernel table of the round summary)"}},
            "roofline": roofline, "clocks": clocks, "gpu_launches": launches, "digest": full_digest}
    if e2e is not None:
        line["e2e"] = e2e
    if world == 1 and not args.no_cpu:
        threads = cpu_threads()
        n = args.cpu_slices or threads
        c = cpu_sample(cfg, coords, elements, n, n, threads, keep=True)
        line["cpu_baseline"] = cpu_baseline_dict(c, threads, cfg["n_phi"])
        # (3) the GPU path on the very slices / orientations the CPU sample computed, same 10 M-atom slab
        probe = slice_engine()
        probe.run(c["phis"])
        iq_s, axis_s = engine.finalize_voxels(probe.vsum, None, probe.count2, probe.row_hist, q_axis, max_q, dev,
                                              window=window)
        Rs, ws = tables(c["psis"], c["w"])
        img_s = torch.zeros_like(image)
        det_s = detector_image(iq_s, axis_s, Rs, ws, img_s, False)
        check["sample_vs_cpu"] = {
            "slices": int(c["n_slices"]), "orientations": int(c["n_orient"]),
            "counts_equal": bool(np.array_equal(probe.counts(), c["vcnt"])),
            "sum_rel_err": float(np.abs(probe.sums() - c["vsum"]).max() / c["vsum"].max()),
            "iq_rel_err": float(np.abs(iq_s.cpu().numpy() - c["iq"]).max() / c["iq"].max()),
            "det_rel_err": float(np.abs(det_s.cpu().numpy() - c["det"]).max() / c["det"].max()),
            "tolerance": 1e-4}
    ok = check.get("full_run_counts_equal_oracle", True)
    if "sample_vs_cpu" in check:
        sc = check["sample_vs_cpu"]
        ok = ok and sc["counts_equal"] and max(sc["sum_rel_err"], sc["iq_rel_err"], sc["det_rel_err"]) <= 1e-4
    if "multi_gpu_vs_1rank" in check:
        mg = check["multi_gpu_vs_1rank"]
        # fp32 voxel sums: 1800 additions per voxel associate differently on N ranks (measured 1.4e-6 of the
        # maximum, at the DC voxel, between 2 ranks and 1); counts must be identical
        mg["tolerance"] = 5e-6
        ok = ok and mg["counts_equal"] and max(mg["iq_rel_err"], mg["image_rel_err"]) <= mg["tolerance"]
    check["ok"] = bool(ok)
    line["check"] = check
    emit(line)
    if world > 1:
        parallel.shutdown()
        dist.destroy_process_group()


SM_COUNT = 148


def roofline_block(cfg, eng, kernel_ms, n_slices_timed, N, peak, peak_src, clocks):
    """Per-kernel bounds of stage A.  The fused pair keeps the N x N grid and image on chip, so the
    HBM floor of the UNFUSED algorithm (622 MB per slice) no longer bounds it; each kernel is
    reported against (a) the HBM bytes the fused design must move (atoms once per batch of rotations,
    the kept-column intermediate written once and read once, voxel read-modify-write) and (b) the issue
    rate: warp instructions per launch (ncu, profiles/r04_kernel_metrics.json) over the live CUDA-event
    time against 4 schedulers x 148 SMs x the SM clock sampled during the run."""
    if not kernel_ms or "rows" not in kernel_ms:
        return None
    A = cfg["n_atoms"]
    B = eng.fused_batch_size()
    n_b = -(-len(cfg["phi_list"]) // B) if len(cfg["phi_list"]) > B else 1
    B = -(-len(cfg["phi_list"]) // n_b)
    rows_active = int(eng.atoms.bounds[2] / eng.r) + 1
    kc = float(getattr(eng, "mean_kept_columns", 0.0)) or 1.12 * (eng.window[1] - eng.window[0])
    kr = eng.row_hi - eng.row_lo
    inter = 8.0 * rows_active * kc                       # kept-column intermediate, complex64
    # the TMA-fed column kernel reads every row slot of its boxes (also the zero rows outside the atom band)
    inter_read = 8.0 * N * kc if N == 4096 else inter
    fused_bytes = {"rows": 17.0 * A / B + inter, "cols": inter_read + 8.0 * kr * kc}
    metrics = {}
    try:
        metrics = json.load(open(os.path.join(ROOT, "profiles", "r04_kernel_metrics.json")))
    except Exception:
        pass
    sm_mhz = clocks.get("sm_mhz") or 1965.0
    issue_peak = 4.0 * SM_COUNT * sm_mhz * 1e6            # warp instructions / s
    fp32_peak = 2.0 * 128 * SM_COUNT * sm_mhz * 1e6       # FLOP/s
    lsu_peak = 1.0 * SM_COUNT * sm_mhz * 1e6              # shared-memory / L1 wavefronts per s
    per = {}
    cols_kernel = "slice_cols_tma" if (N == 4096 and "slice_cols_tma" in metrics) else "slice_cols_fused"
    for name, kern in (("rows", "slice_rows_fused"), ("cols", cols_kernel)):
        us = 1e3 * kernel_ms[name] / n_slices_timed
        ach = fused_bytes[name] / (us * 1e-6) / 1e9
        k = {"kernel": kern, "us_per_slice": us, "us_per_launch": us * B,
             "hbm_bytes_fused_design_per_slice": fused_bytes[name], "hbm_achieved": ach, "hbm_frac": ach / peak}
        m = metrics.get(kern)
        if m:
            spl = float(m["slices_per_launch"])
            k["warp_inst_per_slice"] = m["inst_executed"] / spl
            k["issue_achieved"] = m["inst_executed"] / spl / (us * 1e-6)
            k["issue_frac"] = k["issue_achieved"] / issue_peak
            if m.get("flop"):
                k["fp32_tflops"] = m["flop"] / spl / (us * 1e-6) / 1e12
                k["fp32_frac"] = k["fp32_tflops"] * 1e12 / fp32_peak
            if m.get("lsu_wavefronts"):
                # L1 / shared-memory data pipe: one 128-byte wavefront per SM per cycle
                k["lsu_wavefronts_per_slice"] = m["lsu_wavefronts"] / spl
                k["lsu_pipe_achieved"] = m["lsu_wavefronts"] / spl / (us * 1e-6)
                k["lsu_pipe_frac"] = k["lsu_pipe_achieved"] / lsu_peak
            k["traffic_per_launch"] = m.get("dram_bytes")
            k["source"] = m.get("source")
        per[name] = k
    top = per["rows"]
    unfused = 28.0 * A + 20.0 * N * N + 20.0 * 567 * 636
    pair_us = top["us_per_slice"] + per["cols"]["us_per_slice"]
    return {"bound": "hbm", "kernel": top["kernel"], "achieved": top["hbm_achieved"], "peak": peak, "unit": "GB/s",
            "frac": top["hbm_frac"], "traffic": top.get("traffic_per_launch"), "peak_source": peak_src,
            "slices_per_launch": B, "algorithmic_bytes_per_launch": fused_bytes["rows"] * B,
            "binding_bound": binding_bound(top, issue_peak, lsu_peak),
            "per_kernel": per,
            "kernel_ms_per_slice": {k: v / n_slices_timed for k, v in kernel_ms.items()},
            "unfused_floor": {"algorithmic_bytes_per_slice": unfused,
                              "hbm_floor_us_per_slice": unfused / (peak * 1e9) * 1e6,
                              "fused_pair_us_per_slice": pair_us,
                              "note": "SURVEY 8(d) bytes of the scatter + 2-D FFT + binning kernels the fused "
                                      "pair replaces; the pair is faster than that floor because the N x N grid "
                                      "and image never reach HBM - it is not a fraction of this kernel's roofline"},
            "summary": roofline_summary(per, top),
            "note": "frac = HBM bytes the fused design must move / CUDA-event time / measured copy peak"}


def roofline_summary(per, top):
    """One sentence a reader can check against per_kernel."""
    r, c = per["rows"], per["cols"]
    share = r["us_per_slice"] / (r["us_per_slice"] + c["us_per_slice"])
    parts = ["row kernel = %.0f %% of the fused pair" % (100 * share), "HBM %.3f of peak" % r["hbm_frac"]]
    if r.get("lsu_pipe_frac"):
        parts.append("shared-memory data pipe %.2f of peak" % r["lsu_pipe_frac"])
    if r.get("issue_frac"):
        parts.append("issue slots %.2f" % r["issue_frac"])
    return "; ".join(parts) + "; column kernel: HBM %.2f of peak" % c["hbm_frac"]


def binding_bound(k, issue_peak, lsu_peak):
    """The on-chip resource the row kernel saturates first: its shared-memory data pipe (accumulator zero /
    read, two FFT exchanges, atomics with their bank conflicts) or instruction issue, whichever fraction is larger."""
    cands = [("shared-memory data pipe", k.get("lsu_pipe_achieved"), lsu_peak, "wavefronts/s", k.get("lsu_pipe_frac")),
             ("issue", k.get("issue_achieved"), issue_peak, "warp-inst/s", k.get("issue_frac"))]
    cands = [c for c in cands if c[4]]
    if not cands:
        return None
    b = max(cands, key=lambda c: c[4])
    return {"bound": b[0], "achieved": b[1], "peak": b[2], "unit": b[3], "frac": b[4],
            "others": {c[0]: c[4] for c in cands if c is not b},
            "note": "the kernel is bound on chip, not by DRAM: counts from the ncu capture of the same build "
                    "(profiles/r04_kernel_metrics.json) over the live CUDA-event time"}


_REAL_STDOUT = None


def divert_stdout():
    """Libraries write banners to fd 1 (NCCL prints "NCCL version ..." on the first communicator): keep
    the process's stdout for the single JSON line and send every other byte to stderr."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    """The ONE JSON line goes to the real stdout; everything else was diverted to stderr."""
    os.write(1 if _REAL_STDOUT is None else _REAL_STDOUT, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    divert_stdout()
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
