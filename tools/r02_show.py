"""Prints the gradient legs of a tools/bench_extra.py JSON: python tools/r02_show.py file.json"""
import json
import sys

d = json.load(open(sys.argv[1]))
for k in d:
    if "grad" in k or k == "forward_cost_lw_batched":
        print(k, {a: b for a, b in d[k].items() if a != "config"})
