#!/usr/bin/env python3
"""Refresh profiles/playout_kernel_traffic.json from an ncu summary of ONE bench.py launch of the playout kernel.

usage: update_playout_traffic.py <summary.json from profiles/summarize_ncu.py> <bench line .json> [launch index]

bench.py reads the file for roofline.traffic and for the executed-instruction count per ply, and only trusts
it while `source_sha256` equals the hash of the kernel sources it is built from -- so this must be re-run
(after a fresh `ncu --set full` capture) whenever kernels.cuh or sq_device.cuh change.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    summ = json.load(open(sys.argv[1]))
    line = json.load(open(sys.argv[2]))
    idx = int(sys.argv[3]) if len(sys.argv) > 3 else -1
    launches = [l for l in summ["launches"] if "playout_kernel" in l["kernel"]]
    m = launches[idx]["metrics"]

    def val(k):
        v = float(m[k]["value"])
        u = m[k].get("unit", "")
        return v * {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(u, 1)
    out = {
        "dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
        "warp_instructions_per_launch": val("smsp__inst_executed.sum"),
        "plies_per_launch": line["roofline"]["plies_per_launch"],
        "threads_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": val("sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
        "workload": "bench.py configs[1], one launch (4096 leaves x 1e6 playouts)",
        "source": "%s (ncu --set full --clock-control none)" % os.path.relpath(sys.argv[1], ROOT),
        "source_sha256": bench.source_hash(),
    }
    with open(os.path.join(ROOT, "profiles", "playout_kernel_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
