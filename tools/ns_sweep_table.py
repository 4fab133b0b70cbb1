#!/usr/bin/env python3
"""Condenses a NegaScout knob sweep (tools/gpu_session_[s-u].sh output) into one line per case.  No GPU needed."""
import re
import sys

cur = None
for line in open(sys.argv[1]):
    line = line.rstrip()
    if line.startswith("=="):
        print(line)
        continue
    if line.startswith("ns split"):
        m = re.search(r"(\d+) positions.*?(\d+) (?:flights|rounds), (\d+) items \((\d+) dropped in flight, (\d+) given up\)", line)
        cur = m.groups() if m else None
        continue
    m = re.match(r"(.*?) \{'nodes': (\d+), 'gpu_s': ([\d.]+)", line)
    if m:
        extra = "" if not cur else f"split: {cur[0]} positions, {cur[1]} flights, {cur[2]} items, {cur[3]} dropped, {cur[4]} given up"
        print(f"   {m.group(1):38s} {float(m.group(3)):8.4f} s   {extra}")
        cur = None
