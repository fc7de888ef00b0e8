"""Turn an `ncu --set full` report into the committed summary csv and traffic.json.
usage: python profiles/scripts/summarize.py report.ncu-rep summary.csv traffic.json"""
import csv
import io
import json
import subprocess
import sys

rep, out_csv, out_json = sys.argv[1:4]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
cols = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"]
idx = [hdr.index(c) for c in cols]
scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
with open(out_csv, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[i] for i in idx])
    seen = set()
    for d in data:
        name = d[idx[0]].split("(")[0].split("::")[-1]
        if name in seen:  # one row per kernel: the first captured launch
            continue
        seen.add(name)
        w.writerow([name] + [d[i] for i in idx[1:]])
        rd = float(d[hdr.index("dram__bytes_read.sum")]) * scale[units[hdr.index("dram__bytes_read.sum")]]
        wr = float(d[hdr.index("dram__bytes_write.sum")]) * scale[units[hdr.index("dram__bytes_write.sum")]]
        key = name.replace("_kernel", "").replace("<0>", "_sizes").replace("<1>", "_emit")
        traffic[key] = rd + wr
json.dump(traffic, open(out_json, "w"), indent=1)
print(open(out_csv).read())
