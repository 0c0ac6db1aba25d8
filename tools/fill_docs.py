#!/usr/bin/env python3
"""Re-generates the measured tables/numbers in README.md, DESIGN.md and profiles/README.md from the recorded bench
lines (between <!-- r02-table --> markers), so the prose never drifts from the files under profiles/."""
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
table = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "profile_table.py"), "r02"], capture_output=True, text=True).stdout.strip()
block = f"<!-- r02-table -->\n{table}\n<!-- /r02-table -->"


def load(name):
    return json.loads([l for l in open(os.path.join(ROOT, "profiles", name)) if l.startswith("{")][-1])


n8 = load("r02_bench_n8.json")
subs = {"@@N8@@": f"{n8['value']:.3f}", "@@BAR8@@": f"{n8['barrier_us']:.0f}"}
for name in ("README.md", "DESIGN.md", os.path.join("profiles", "README.md")):
    p = os.path.join(ROOT, name)
    s = open(p).read()
    if "@@TABLE@@" in s:
        s = s.replace("@@TABLE@@", block)
    else:
        s = re.sub(r"<!-- r02-table -->.*?<!-- /r02-table -->", lambda m: block, s, flags=re.S)
    for k, v in subs.items():
        s = s.replace(k, v)
    open(p, "w").write(s)
    print("updated", name)
