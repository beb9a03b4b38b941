"""Per-launch counters of the search kernels out of an ncu --set full capture -> profiles/r02_kernel_counters.json
(what bench.py's roofline reads).

  python tools/ncu_counters.py capture.ncu-rep step.json [out.json]

step.json = the line tools/profile_step.py printed in the SAME run (node counts of the captured launches)."""
import csv
import json
import subprocess
import sys

rep, step = sys.argv[1], json.load(open(sys.argv[2]))
out_path = sys.argv[3] if len(sys.argv) > 3 else "profiles/r02_kernel_counters.json"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]


def col(name):
    return h.index(name) if name in h else None


UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3,
        "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}


def num(row, name, scaled=False):
    """value of a metric; scaled: in bytes / milliseconds (the report prints each launch's metric in its own unit)"""
    i = col(name)
    if i is None:
        return None
    try:
        v = float(row[i].replace(",", ""))
    except Exception:
        return None
    return v * UNIT.get(rows[1][i], 1) if scaled else v


name_i = col("Kernel Name")
out = {}
for row in rows[2:]:
    kn = row[name_i]
    short = "text_pool_kernel" if "text_pool_kernel" in kn else "fm_items_kernel" if "fm_items_kernel" in kn else None
    if not short or short in out:
        continue
    dram = (num(row, "dram__bytes_read.sum", True) or 0) + (num(row, "dram__bytes_write.sum", True) or 0)
    miss = num(row, "lts__t_sectors_srcunit_tex_lookup_miss.sum")
    rec = {"inst_executed": num(row, "smsp__inst_executed.sum"),
           "lanes_per_inst": num(row, "smsp__thread_inst_executed_per_inst_executed.ratio"),
           "issue_active_pct": num(row, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "dram_bytes": dram,
           "l2_miss_sectors": miss,
           # (only when the report has a request counter: missed SECTORS also count streamed traffic — items, seed records —
           # and must not be held against the dependent-gather ceiling)
           "l2_miss_requests": num(row, "lts__t_requests_srcunit_tex_lookup_miss.sum"),
           "duration_under_ncu_ms": round(num(row, "gpu__time_duration.sum", True) or 0, 4),
           "registers_per_thread": num(row, "launch__registers_per_thread"),
           "warps_active_pct": num(row, "sm__warps_active.avg.pct_of_peak_sustained_active"),
           "nodes_per_launch": step["nodes_text"] if short == "text_pool_kernel" else step["nodes_fm"],
           "workload": step["workload"], "source": rep.split("/")[-1]}
    out[short] = rec
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
