"""Per-kernel figures bench.py's roofline block needs, from `ncu --set full` captures:
warp instructions, fp32 flops (fadd + fmul + 2 ffma thread instructions; packed f32x2 forms count per lane),
DRAM bytes and duration per launch.
usage: python scripts/ncu_kernel_metrics.py out.json units_per_launch:report.ncu-rep [...]
(units = slices for the stage-A kernels, orientations for the detector kernel)"""
import csv, json, subprocess, sys

out = {}
for arg in sys.argv[2:]:
    units, rep = arg.split(":", 1)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, unit_row = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, unit_row))
        name = d["Kernel Name"].replace("void ", "").split("<")[0].split("(")[0]

        def val(k):
            v = float(d[k].replace(",", ""))
            mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e-3, "usecond": 1e-6,
                    "second": 1.0, "nsecond": 1e-9, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u[k], 1.0)
            return v * mult

        cyc = val("sm__cycles_elapsed.max")
        per_cycle = lambda op: val("smsp__sass_thread_inst_executed_op_%s_pred_on.sum.per_cycle_elapsed" % op)
        flop = (per_cycle("fadd") + per_cycle("fmul") + 2 * per_cycle("ffma")) * cyc
        out[name] = {
            "full_name": d["Kernel Name"][:120], "units_per_launch": float(units), "slices_per_launch": float(units),
            "duration_s": val("gpu__time_duration.sum"), "inst_executed": val("smsp__inst_executed.sum"),
            "flop": flop, "dram_bytes": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
            "dram_read": val("dram__bytes_read.sum"), "dram_write": val("dram__bytes_write.sum"),
            "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
            "registers": val("launch__registers_per_thread"),
            # L1 / shared-memory data pipe (one 128-byte wavefront per SM per cycle): the limit of a kernel that
            # lives in shared memory; wavefronts = pct x SM cycles, so bench.py can restate it for its own timing
            "lsu_data_pipe_pct": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "lsu_wavefronts": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed") / 100.0 *
                              val("sm__cycles_elapsed.avg") * val("launch__sm_count") if "launch__sm_count" in d else None,
            "shared_wavefronts": val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active"),
            "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": val("lts__t_sector_hit_rate.pct"),
            "stalls_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: float(d[k])
                                 for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and
                                 k.endswith("per_issue_active.ratio") and float(d[k] or 0) >= 0.3},
            "source": "profiles/%s (ncu --set full --clock-control none, one launch)" % rep.split("/")[-1].replace(".ncu-rep", ".csv")}
json.dump(out, open(sys.argv[1], "w"), indent=1)
for k, v in out.items():
    print("%-28s %8.1f us  %7.1f Minst  issue %4.1f%%  %6.2f GFLOP  DRAM %7.1f MB  stalls %s" % (
        k, v["duration_s"] * 1e6, v["inst_executed"] / 1e6, v["issue_active_pct"], v["flop"] / 1e9, v["dram_bytes"] / 1e6,
        {a: round(b, 2) for a, b in v["stalls_per_issue"].items()}))
