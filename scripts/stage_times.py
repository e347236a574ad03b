#!/usr/bin/env python
"""Print the per-stage device times of one small bench run (profiling helper)."""
import json
import os
import subprocess
import sys

args = sys.argv[1:] or ["--records", "16000000", "--batch-records", "2000000", "--steps", "3"]
out = subprocess.run([sys.executable, "bench.py", "--no-cpu-baseline", "--no-e2e"] + args, capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    st = d["roofline"]["stage_ms_per_launch"]
    print(os.environ.get("TAG", ""), {k: round(v, 3) for k, v in st.items()}, "value %.1f M/s" % (d["value"] / 1e6))
except Exception as e:  # noqa: BLE001
    print("FAILED", e, out.stderr[-1500:])
