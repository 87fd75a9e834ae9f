"""Turn ncu output brought back in gpurun_out/ into the small text/JSON summaries kept under profiles/.

    python scripts/summarise_profiles.py launches gpurun_out/launches_bench_condensed.csv CYCLES "command line" > profiles/....txt
    python scripts/summarise_profiles.py full gpurun_out/prof_condensed_apply_full.ncu-rep > profiles/....json

`launches` reads the CSV log of `ncu --metrics gpu__time_duration.sum --clock-control none --csv`
(one row per launch) and prints per-kernel totals, for the whole process and for the kernels of
the F-cycle alone (everything that is not per-Newton-step setup).  `full` reads a `--set full`
report through `ncu -i ... --page raw --csv` and keeps the metrics the roofline argument uses.
"""
import collections
import csv
import json
import re
import subprocess
import sys

SETUP = re.compile(r"patch_factor|condense_blocks|cutlass|trsm|getrf|ipiv|laswp|create_pivot|xxtrf|copy_info|"
                   r"set_identity|bsr_to_dense|at::")
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.pct_of_peak_sustained_elapsed", "dram__bytes.sum.per_second",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
        "launch__occupancy_limit_registers", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = name.replace("<unnamed>::", "")
    return re.sub(r"\(.*$", "", name)


def launches(path, cycles, command):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    h = rows[0]
    ki, vi, ui = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for r in rows[1:]:
        v = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[r[ui]]
        tot[short(r[ki])] += v
        cnt[short(r[ki])] += 1
    print(command)
    print("times are cold-cache and serialised (ncu): compare shares, not absolutes\n")

    def table(items):
        T = sum(v for _, v in items)
        for k, v in sorted(items, key=lambda kv: -kv[1]):
            print("%-70s n=%6d %12.3f ms %6.2f%%" % (k[:70], cnt[k], v, 100 * v / T))
        print("%-70s n=%6d %12.3f ms" % ("total", sum(cnt[k] for k, _ in items), T))
    print("whole process (%d launches):" % sum(cnt.values()))
    table(list(tot.items()))
    cyc = [(k, v) for k, v in tot.items() if not SETUP.search(k)]
    print("\ncycle kernels only (%d F-cycles; per cycle = total / %d):" % (cycles, cycles))
    table(cyc)


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"Kernel Name": short(r[h.index("Kernel Name")])}
        for m in KEEP:
            if m in h:
                i = h.index(m)
                d[m] = ("%s %s" % (r[i], units[i])).strip()
        stalls = {}
        for i, m in enumerate(h):
            if m.startswith("smsp__average_warps_issue_stalled") and m.endswith("per_issue_active.ratio") and "not_issued" not in m:
                try:
                    stalls[m.split("stalled_")[1].split("_per_")[0]] = float(r[i])
                except ValueError:
                    pass
        d["top_stalls_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:4])
        res.append(d)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]), sys.argv[4])
    else:
        full(sys.argv[2])
