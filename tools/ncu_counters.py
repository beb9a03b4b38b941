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


def num(row, name):
    i = col(name)
    if i is None:
        return None
    try:
        return float(row[i].replace(",", ""))
    except Exception:
        return None


name_i = col("Kernel Name")
out = {}
for row in rows[2:]:
    kn = row[name_i]
    short = "text_pool_kernel" if "text_pool_kernel" in kn else "fm_items_kernel" if "fm_items_kernel" in kn else None
    if not short or short in out:
        continue
    dram = (num(row, "dram__bytes_read.sum") or 0) + (num(row, "dram__bytes_write.sum") or 0)
    unit_i = col("dram__bytes_read.sum")
    unit = rows[1][unit_i] if unit_i is not None else "byte"
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1)
    rec = {"inst_executed": num(row, "smsp__inst_executed.sum"),
           "lanes_per_inst": num(row, "smsp__thread_inst_executed_per_inst_executed.ratio"),
           "issue_active_pct": num(row, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "dram_bytes": dram * mult,
           "l2_miss_sectors": num(row, "lts__t_sectors_srcunit_tex_lookup_miss.sum"),
           "l2_miss_requests": num(row, "lts__t_requests_srcunit_tex_lookup_miss.sum"),
           "duration_under_ncu_ms": round((num(row, "gpu__time_duration.sum") or 0) / 1e6, 4),
           "registers_per_thread": num(row, "launch__registers_per_thread"),
           "warps_active_pct": num(row, "sm__warps_active.avg.pct_of_peak_sustained_active"),
           "nodes_per_launch": step["nodes_text"] if short == "text_pool_kernel" else step["nodes_fm"],
           "workload": step["workload"], "source": rep.split("/")[-1]}
    out[short] = rec
json.dump(out, open(out_path, "w"), indent=1)
print(json.dumps(out, indent=1))
