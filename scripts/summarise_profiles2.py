"""Build-box helper (round 2): turn an ncu launch-list CSV and a full-capture .ncu-rep (brought back in gpurun_out/) into
the text summaries kept under profiles/.

    python scripts/summarise_profiles2.py <launches.csv | -> <capture.ncu-rep | -> <tag> "<command that was profiled>"
"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, tag, cmd = sys.argv[1:5]
P = os.path.join(ROOT, "profiles")

if launch_csv != "-":
    rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
    ix = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        unit = r[ix["Metric Unit"]]
        us = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
        a = agg.setdefault(r[ix["Kernel Name"]][:70], [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    shutil.copy(launch_csv, os.path.join(P, f"r2_launches_{tag}.csv"))
    with open(os.path.join(P, f"r2_launches_{tag}.txt"), "w") as f:
        f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised launches): {cmd}\n")
        f.write(f"{'kernel':70s} launches   total us   avg us  share\n")
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{k:70s} {a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:8.1f} {100 * a[1] / tot:5.1f}%\n")
    print(open(os.path.join(P, f"r2_launches_{tag}.txt")).read())

if rep != "-":
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, u = rows[0], rows[1]
    want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
            "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    seen, traffic = set(), {}
    with open(os.path.join(P, f"r2_{tag}_ncu_full_summary.txt"), "w") as f:
        f.write(f"ncu --set full --clock-control none --import-source on: {cmd}\n(first captured launch of every kernel)\n")
        for v in rows[2:]:
            name = v[h.index("Kernel Name")]
            if name in seen:
                continue
            seen.add(name)
            f.write(f"\n== {name}\n")
            d = {}
            for i, n in enumerate(h):
                d[n] = (v[i], u[i])
                if n in want or n.startswith("sm__pipe_tensor") or n.startswith("sm__inst_executed_pipe_tensor"):
                    f.write(f"{n} [{u[i]}] = {v[i]}\n")
            try:
                tr = float(d["dram__bytes_read.sum"][0].replace(",", "")) * scale[d["dram__bytes_read.sum"][1]]
                tw = float(d["dram__bytes_write.sum"][0].replace(",", "")) * scale[d["dram__bytes_write.sum"][1]]
                traffic[name] = {"dram_bytes_read": int(tr), "dram_bytes_write": int(tw), "dram_bytes_per_launch": int(tr + tw),
                                 "grid": d["launch__grid_size"][0], "duration_us_under_ncu": d["gpu__time_duration.sum"][0]}
            except Exception:
                pass
    json.dump({"source": f"profiles/r2_{tag}_ncu_full_summary.txt (ncu --set full, one launch per kernel)", "kernels": traffic},
              open(os.path.join(P, f"r2_{tag}_traffic.json"), "w"), indent=1)
    print(open(os.path.join(P, f"r2_{tag}_ncu_full_summary.txt")).read()[:3000])
