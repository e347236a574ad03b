#!/bin/bash
# GPU session: suite, bench (with the compressed-input leg), the CLI file to file (configs[1], device-side file path vs --host_io)
set -u
cd "$(dirname "$0")/.."
tag=${1:-r2i}
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > "$out/box.txt" 2>&1
timeout 1200 python -m pytest tests -q -m gpu --maxfail=6 > "$out/tests.log" 2>&1; tail -5 "$out/tests.log"
timeout 900 python bench.py --e2e-bgzf > "$out/bench_n1.json" 2> "$out/bench_n1.err"; cat "$out/bench_n1.json"
timeout 600 python scripts/run_cfg2_cli.py --host-io --runs 3 > "$out/cfg2_cli.json" 2> "$out/cfg2_cli.err"; grep "stage seconds\|ms on the GPU" "$out/cfg2_cli.err"; cat "$out/cfg2_cli.json"; echo
