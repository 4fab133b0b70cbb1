#!/usr/bin/env python3
"""Condenses a NegaScout knob sweep (tools/gpu_session_[s-w].sh output: SQ_NS_TRACE lines + tools/bench_negascout.py
lines) into one line per case.  No GPU needed."""
import re
import sys

split, first = None, None
for line in open(sys.argv[1]):
    line = line.rstrip()
    if line.startswith("=="):
        print(line)
        continue
    if line.startswith("ns one warp"):
        m = re.search(r"([\d.]+) s$", line)
        first = m.group(1) if m else None
        continue
    if line.startswith("ns split"):
        m = re.search(r"(?:([\d.]+) s, )?(\d+) positions.*?(\d+) (?:flights|rounds), (\d+) items \((\d+) dropped in flight, (\d+) given up\)", line)
        split = m.groups() if m else None
        continue
    m = re.match(r"(.*?) \{'nodes': (\d+), 'gpu_s': ([\d.]+)", line)
    if m:
        extra = ""
        if first:
            extra += f"first pass {first} s; "
        if split:
            extra += (f"split {split[0]} s: " if split[0] else "split: ") + \
                     f"{split[1]} positions, {split[2]} flights, {split[3]} items, {split[4]} dropped, {split[5]} given up"
        print(f"   {m.group(1):38s} {float(m.group(3)):8.4f} s   {extra}")
        split, first = None, None
