"""Prints the headline numbers of one bench.py JSON line: python tools/r02_line.py file.json [label]"""
import json
import sys

try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print("%s: %.1f G/s kernel_ms %.3f frac %.3f" % (sys.argv[2] if len(sys.argv) > 2 else "", d["value"] / 1e9, r["kernel_ms"], r["frac"]))
except Exception as e:  # noqa: BLE001
    print(sys.argv[1], "FAILED", e)
