"""Summarise an .ncu-rep (ncu --set full) per launch: the counters the roofline discussion uses, as CSV on stdout and, with
--json KEY1,KEY2,..., as entries for profiles/roofline_captured.json (one key per captured launch, '-' skips a launch).

    python tools/ncu_summary.py gpurun_out/x.ncu-rep [--json k1,k2,...] [--source "capture name"]
"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}


def val(r, name, scale=1.0):
    if name not in col:
        return None
    try:
        v = float(r[col[name]].replace(",", ""))
    except Exception:
        return None
    u = units[col[name]]
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "msecond": 1e3, "usecond": 1.0, "second": 1e6, "nsecond": 1e-3,
            "ms": 1e3, "us": 1.0, "s": 1e6, "ns": 1e-3}.get(u, 1.0)
    return v * mult * scale


FIELDS = [("kernel", None), ("ncu_duration_us", "gpu__time_duration.sum"), ("registers", "launch__registers_per_thread"),
          ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"), ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
          ("warp_instructions", "smsp__inst_executed.sum"), ("threads_per_instruction", "smsp__thread_inst_executed_per_inst_executed.ratio"),
          ("dram_read_bytes", "dram__bytes_read.sum"), ("dram_write_bytes", "dram__bytes_write.sum"),
          ("lts_sectors", "lts__t_sectors.sum"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"), ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"),
          ("l1_wavefronts", "l1tex__data_pipe_lsu_wavefronts.sum"), ("lsu_pipe_pct", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active"),
          ("alu_pipe_pct", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"), ("fma_pipe_pct", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
          ("eligible_warps_per_cycle", "smsp__warps_eligible.avg.per_cycle_active"),
          ("stall_long_scoreboard", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
          ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
          ("stall_short_scoreboard", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
          ("stall_branch_resolving", "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio"),
          ("stall_not_selected", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"),
          ("stall_no_instruction", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"),
          ("stall_math_throttle", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
          ("sm_cycles_max", "sm__cycles_elapsed.max")]
out = []
for r in data:
    e = {}
    for key, metric in FIELDS:
        e[key] = r[col["Kernel Name"]][:60] if metric is None else val(r, metric)
    e["dram_bytes_per_launch"] = (e["dram_read_bytes"] or 0) + (e["dram_write_bytes"] or 0)
    e["lts_bytes_per_launch"] = (e["lts_sectors"] or 0) * 32.0
    out.append(e)
w = csv.writer(sys.stdout)
keys = [k for k, _ in FIELDS] + ["dram_bytes_per_launch", "lts_bytes_per_launch"]
w.writerow(keys)
for e in out:
    w.writerow([("%.6g" % e[k]) if isinstance(e[k], float) else e[k] for k in keys])
if "--json" in sys.argv:
    names = sys.argv[sys.argv.index("--json") + 1].split(",")
    src = sys.argv[sys.argv.index("--source") + 1] if "--source" in sys.argv else rep
    j = {}
    for n, e in zip(names, out):
        if n == "-":
            continue
        d = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in e.items() if v is not None}
        d["source"] = src
        d["bound"] = "l2/issue: instruction issue under divergence, scene served by L1/L2 (dram is the ray records only)"
        j[n] = d
    sys.stderr.write(json.dumps(j, indent=1) + "\n")
