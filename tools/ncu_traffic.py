#!/usr/bin/env python
"""profiles/r02_traffic.json: the counters of the dominant kernels that bench.py quotes beside its live numbers, extracted from a
committed `ncu --set full` capture (launch-duration-weighted means over the captured launches of each kernel).
usage: tools/ncu_traffic.py <tag>[:<key prefix>[:<bench args of the capture>]] ...
(reads gpurun_out/<tag>.ncu-rep, merges into profiles/r02_traffic.json under <key prefix><kernel name>)"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles", "r02_traffic.json")
M = {"ms": "gpu__time_duration.sum", "issue_slots_busy_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
     "active_threads_per_warp": "smsp__thread_inst_executed_per_inst_executed.ratio",
     "l1tex_throughput_pct": "l1tex__throughput.avg.pct_of_peak_sustained_active", "l2_throughput_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "achieved_occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "dram_read": "dram__bytes_read.sum", "dram_write": "dram__bytes_write.sum", "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
     "registers": "launch__registers_per_thread", "warp_instructions": "smsp__inst_executed.sum"}
UNIT = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0, "msecond": 1.0, "usecond": 1e-3, "nsecond": 1e-6, "second": 1e3, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}


def main():
    data = json.load(open(OUT)) if os.path.exists(OUT) else {}
    for spec in sys.argv[1:]:
        tag, prefix, cmd = (spec.split(":", 2) + ["", ""])[:3]
        cmd = cmd or "--steps 1 --warmup 1 --spp 2 --no-cpu-baseline"
        rep = os.path.join(ROOT, "gpurun_out", tag + ".ncu-rep")
        rows = list(csv.reader(subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout.splitlines()))
        h, units = rows[0], rows[1]
        col = {k: h.index(v) for k, v in M.items() if v in h}
        per = collections.OrderedDict()
        for r in rows[2:]:
            name = r[h.index("Kernel Name")].split("(")[0].replace("void ", "").replace("crt::", "").split("<")[0].strip()
            vals = {k: float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0) for k, i in col.items()}
            per.setdefault(name, []).append(vals)
        for name, launches in per.items():
            key = prefix + name
            tot = sum(l["ms"] for l in launches)
            mean = {k: sum(l[k] * l["ms"] for l in launches) / tot for k in launches[0] if k not in ("ms", "dram_read", "dram_write", "registers", "warp_instructions")}
            e = {k: round(v, 2) for k, v in mean.items()}
            e.update(launches=len(launches), mean_launch_ms=round(tot / len(launches), 4), dram_bytes_per_launch=int(sum(l["dram_read"] + l["dram_write"] for l in launches) / len(launches)),
                     registers=int(launches[0]["registers"]), warp_instructions_per_launch=int(sum(l["warp_instructions"] for l in launches) / len(launches)),
                     **{"from": f"profiles/{tag}_details.txt (ncu --set full --clock-control none, `bench.py {cmd}`, "
                                f"{len(launches)} launches of {name} incl. all its template instances; duration-weighted means)"})
            e["issue_fraction"] = round(e["issue_slots_busy_pct"] / 100 * e["active_threads_per_warp"] / 32, 4)
            data[key] = e
    json.dump(data, open(OUT, "w"), indent=1)
    for k, v in data.items():
        print(k, {x: v[x] for x in ("issue_slots_busy_pct", "active_threads_per_warp", "issue_fraction", "l1tex_throughput_pct", "dram_throughput_pct", "mean_launch_ms")})


if __name__ == "__main__":
    main()
