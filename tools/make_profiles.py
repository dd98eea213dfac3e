"""Turn the ncu artefacts of a gpurun session (gpurun_out/) into the committed summaries under
profiles/.  Usage: python tools/make_profiles.py <round tag> <ncu-rep> <frames per launch>"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, rep, frames = sys.argv[1], sys.argv[2], int(sys.argv[3])
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


details = ncu("--page", "details")
with open(os.path.join(out_dir, f"{tag}_details.txt"), "w") as f:
    f.write(details)
raw = list(csv.reader(ncu("--page", "raw", "--csv").splitlines()))
hdr, units, vals = raw[0], raw[1], raw[2]
keep = ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__inst_executed.sum",
        "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")
metrics = {h: (v, u) for h, u, v in zip(hdr, units, vals) if h in keep or h.startswith("smsp__average_warps_issue_stalled")}
with open(os.path.join(out_dir, f"{tag}_metrics.csv"), "w") as f:
    f.write("metric,value,unit\n")
    for k in sorted(metrics):
        f.write(f"{k},{metrics[k][0]},{metrics[k][1]}\n")


def to_bytes(value, unit):
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
    return float(value.replace(",", "")) * scale


rd = to_bytes(*metrics["dram__bytes_read.sum"])
wr = to_bytes(*metrics["dram__bytes_write.sum"])
src = os.path.join(out_dir, f"{tag}_source.csv")
with open(src, "w") as f:
    f.write(ncu("--page", "source", "--csv"))
mix = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_opmix.py"), src, "24"],
                     capture_output=True, text=True).stdout
with open(os.path.join(out_dir, f"{tag}_opmix.txt"), "w") as f:
    f.write(mix)
os.remove(src)  # multi-megabyte; the op mix is the summary
traffic = {"kernel": tag, "frames_per_launch": frames, "dram_bytes_read": rd, "dram_bytes_write": wr,
           "dram_bytes_per_frame": (rd + wr) / frames, "algorithmic_bytes_per_frame": 804}
with open(os.path.join(out_dir, "traffic.json"), "w") as f:
    json.dump(traffic, f, indent=1)
print(json.dumps(traffic))
print(mix)
