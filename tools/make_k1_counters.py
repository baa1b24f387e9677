#!/usr/bin/env python
"""profiles/k1_counters.json from ncu metric passes over one K1 step (tools/prof_k1.py <cfg> <reactions> lockstep):

    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum,\\
gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none --csv \\
        --log-file counts_cfg2.csv python tools/prof_k1.py cfg2 4000000 lockstep

usage: make_k1_counters.py out.json cfg2=counts_cfg2.csv:4000000 [cfg1=...:N cfg3=...:N]

Every kernel of the sonics library and of the CUB sort counts (k_sort_keys, DeviceRadixSort*, k_pcr_ls); torch's own
fill / copy kernels of the measuring script do not.  The file carries the hash of the kernel sources it was captured
from; bench.py refuses it when the sources have changed."""
import csv
import json
import os
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_sha)

OURS = ("sonics::", "cub::", "DeviceRadixSort")


def read(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0] != "ID" and r[0].isdigit()]
    agg = {}
    for r in rows:
        name, metric, unit, val = r[4], r[-3], r[-2], float(r[-1].replace(",", ""))
        if not any(k in name for k in OURS) or "k_table_prep" in name:
            continue
        short = "k_pcr_ls" if "k_pcr_ls" in name else "k_sort_keys" if "k_sort_keys" in name else "radix_sort"
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "second": 1e9}.get(unit, 1.0)
        d = agg.setdefault(short, {})
        if "pct_of_peak" in metric:
            d.setdefault(metric, []).append(val)
        else:
            d[metric] = d.get(metric, 0.0) + val * scale
    return agg


def main():
    out = sys.argv[1]
    res = {"kernel_sha": bench.kernel_sha(), "per_config": {}}
    for spec in sys.argv[2:]:
        name, rest = spec.split("=")
        path, n = rest.rsplit(":", 1)
        n = float(n)
        agg = read(path)
        wi = sum(k.get("smsp__inst_executed.sum", 0.0) for k in agg.values())
        ti = sum(k.get("smsp__thread_inst_executed.sum", 0.0) for k in agg.values())
        dram = sum(k.get("dram__bytes_read.sum", 0.0) + k.get("dram__bytes_write.sum", 0.0) for k in agg.values())
        ns = sum(k.get("gpu__time_duration.sum", 0.0) for k in agg.values())
        pcr = agg["k_pcr_ls"]
        cfg = {"reactions_in_step": n, "warp_inst_per_reaction": wi / n, "thread_inst_per_reaction": ti / n,
               "avg_active_threads": ti / wi, "dram_bytes_per_reaction": dram / n,
               "k_pcr_ls_share_of_step_time": pcr["gpu__time_duration.sum"] / ns,
               "k_pcr_ls_issue_active_pct": pcr["smsp__issue_active.avg.pct_of_peak_sustained_active"][0],
               "kernels": {k: {"warp_inst_per_reaction": v.get("smsp__inst_executed.sum", 0.0) / n,
                               "dram_bytes_per_reaction": (v.get("dram__bytes_read.sum", 0.0) +
                                                           v.get("dram__bytes_write.sum", 0.0)) / n,
                               "ns_per_reaction_under_ncu": v.get("gpu__time_duration.sum", 0.0) / n}
                           for k, v in agg.items()},
               "source": os.path.basename(path)}
        res["per_config"][name] = cfg
        if name == "cfg2":
            for k in ("warp_inst_per_reaction", "thread_inst_per_reaction", "avg_active_threads",
                      "dram_bytes_per_reaction"):
                res[k] = cfg[k]
            res["issue_active_pct"] = cfg["k_pcr_ls_issue_active_pct"]
            res["source"] = "profiles/%s (ncu metric pass over every kernel of one cfg2 step of %d reactions)" % (
                os.path.basename(path), int(n))
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
