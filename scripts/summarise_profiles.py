"""Build box helper: turn an ncu launch-list CSV and a full-capture .ncu-rep (brought back in gpurun_out/) into the
text summaries kept under profiles/.   python scripts/summarise_profiles.py <launches.csv> <capture.ncu-rep> <tag> "<cmd>" """
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
launch_csv, rep, tag, cmd = sys.argv[1:5]
rows = [r for r in csv.reader(open(launch_csv)) if len(r) > 5]
ix = {h: i for i, h in enumerate(rows[0])}
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[ix["Metric Value"]])
    except ValueError:
        continue
    unit = r[ix["Metric Unit"]]
    us = v / 1000 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000)
    a = agg.setdefault(r[ix["Kernel Name"]][:60], [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in agg.values())
with open(os.path.join(ROOT, "profiles", f"r1_launches_{tag}.txt"), "w") as f:
    f.write(f"ncu --metrics gpu__time_duration.sum --clock-control none -c 400 (cold-cache, serialised): {cmd}\n")
    f.write(f"{'kernel':60s} launches   total us   avg us  share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:60s} {a[0]:8d} {a[1]:10.1f} {a[1] / a[0]:8.1f} {100 * a[1] / tot:5.1f}%\n")
    fit = [a for k, a in agg.items() if "k_icnn_fit_tc" in k]
    opt = [a for k, a in agg.items() if "k_reduce_opt_aug" in k]
    if fit and opt:
        s = fit[0][1] + opt[0][1]
        f.write(f"\nshare of the fit step (fit kernel + optimizer kernel): tc_fused {100 * fit[0][1] / s:.1f}%  "
                f"reduce_opt {100 * opt[0][1] / s:.1f}%\n")
print(open(os.path.join(ROOT, "profiles", f"r1_launches_{tag}.txt")).read())
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h, u, v = rows[0], rows[1], rows[2]
want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "launch__block_size", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "sm__cycles_elapsed.max", "smsp__cycles_active.avg", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active"]
with open(os.path.join(ROOT, "profiles", f"r1_{tag}_ncu_full_summary.txt"), "w") as f:
    f.write(f"ncu --set full --clock-control none --import-source on -k regex:k_icnn_fit_tc -s 5 -c 1: {cmd}\n")
    for i, n in enumerate(h):
        if n in want or n.startswith("sm__pipe_tensor"):
            f.write(f"{n} [{u[i]}] = {v[i]}\n")
d = {h[i]: (v[i], u[i]) for i in range(len(h))}
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tr = float(d["dram__bytes_read.sum"][0]) * scale[d["dram__bytes_read.sum"][1]]
tw = float(d["dram__bytes_write.sum"][0]) * scale[d["dram__bytes_write.sum"][1]]
json.dump({"kernel": "k_icnn_fit_tc<2,2>", "dram_bytes_read": int(tr), "dram_bytes_write": int(tw),
           "dram_bytes_per_launch": int(tr + tw), "grid": d["launch__grid_size"][0],
           "source": f"profiles/r1_{tag}_ncu_full_summary.txt (ncu --set full, one launch)"},
          open(os.path.join(ROOT, "profiles", f"r1_{tag}_traffic.json"), "w"), indent=1)
print(open(os.path.join(ROOT, "profiles", f"r1_{tag}_ncu_full_summary.txt")).read()[:1200])
