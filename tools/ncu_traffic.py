"""Fold one ncu --set full capture into profiles/r02_traffic.json (what bench.py's roofline record quotes):
    python tools/ncu_traffic.py <report.ncu-rep> <kernel class> "<which launch>" <algorithmic bytes per launch>"""
import csv, io, json, os, subprocess, sys
rep, cls, launch, alg = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, r = rows[0], rows[2]
def g(k):
    return float(r[hdr.index(k)].replace(",", "")) if k in hdr else None
def in_bytes(k):
    v, u = g(k), rows[1][hdr.index(k)].lower()
    return v * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
dur_us = g("gpu__time_duration.sum") * {"us": 1, "ns": 1e-3, "ms": 1e3, "usecond": 1, "msecond": 1e3, "nsecond": 1e-3}.get(
    rows[1][hdr.index("gpu__time_duration.sum")].lower(), 1)
rec = {"launch": launch, "kernel": r[hdr.index("Kernel Name")][:80],
       "dram_bytes_per_launch": in_bytes("dram__bytes_read.sum") + in_bytes("dram__bytes_write.sum"),
       "algorithmic_bytes_per_launch": alg, "duration_us": dur_us,
       "issue_active": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
       "pipe_fma": g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
       "pipe_xu": g("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
       "dram_pct_of_peak": g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
       "warps_active_pct": g("sm__warps_active.avg.pct_of_peak_sustained_active"), "report": os.path.basename(rep)}
path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r02_traffic.json")
d = json.load(open(path)) if os.path.exists(path) else {}
d[cls] = rec
json.dump(d, open(path, "w"), indent=1)
print(json.dumps(rec, indent=1))
