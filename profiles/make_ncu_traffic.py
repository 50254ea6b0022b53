"""profiles/ncu_traffic.json from `ncu --page raw --csv` dumps: the DRAM bytes per launch that bench.py reports as
`roofline.traffic` / `spectral_step.traffic` (read, never hard-coded).  Usage:
    python profiles/make_ncu_traffic.py --ray RAW.csv --ray-label "<capture file, command>" [--flow RAW.csv --flow-label ...]
Every entry records its source capture and the commit the capture was taken at."""
import argparse
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rows(path):
    r = list(csv.reader(open(path)))
    hdr, units = r[0], r[1]
    out = []
    for line in r[2:]:
        d = {}
        for h, u, v in zip(hdr, units, line):
            d[h] = (v, u)
        out.append(d)
    return out


def to_bytes(vu):
    v, u = vu
    x = float(v.replace(",", ""))
    return x * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def pct(d, k):
    return round(float(d[k][0]), 1) if k in d else None


ap = argparse.ArgumentParser()
ap.add_argument("--ray"); ap.add_argument("--ray-label", default="")
ap.add_argument("--ray-packets", type=int, default=16777216); ap.add_argument("--nx", type=int, default=2048)
ap.add_argument("--flow"); ap.add_argument("--flow-label", default="")
a = ap.parse_args()
sha = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip()
path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
out = json.load(open(path)) if os.path.exists(path) else {}
if a.ray:
    rr = [d for d in rows(a.ray) if "raytrace_rk4" in d["Kernel Name"][0]]
    d = rr[-1]
    name = d["Kernel Name"][0]
    kern = name[name.index("raytrace_rk4"):name.index("(", name.index("raytrace_rk4"))].replace("(int)", "")
    out["raytrace"] = {
        "kernel": kern, "nx": a.nx, "packets": a.ray_packets,
        "dram_bytes_per_launch": to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"]),
        "source": f"ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch; {a.ray_label}; commit {sha}",
        "binding_unit": {"name": "l1tex__data_pipe_lsu_wavefronts", "pct_of_peak": pct(d, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                         "fp64_pipe_pct": pct(d, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                         "issue_active_pct": pct(d, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                         "warps_active_pct": pct(d, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                         "registers": int(float(d["launch__registers_per_thread"][0]))}}
if a.flow:
    want = ("ypass_inv", "xpass_kernel", "ypass_fwd", "ifmab3_update")
    fr = [d for d in rows(a.flow) if any(w in d["Kernel Name"][0] for w in want) and "Snapshot" not in d["Kernel Name"][0] and "Psi" not in d["Kernel Name"][0]]
    per = {}
    for d in fr:                                   # last launch of each of the four kernels of a step
        key = next(w for w in want if w in d["Kernel Name"][0])
        per[key] = to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"])
    out["spectral_step"] = {"nx": a.nx, "dram_bytes_per_step": sum(per.values()), "per_kernel": per,
                            "source": f"ncu dram__bytes_read.sum + dram__bytes_write.sum of the four kernels of one RSW step; {a.flow_label}; commit {sha}"}
json.dump(out, open(path, "w"), indent=1)
print(json.dumps(out, indent=1))
