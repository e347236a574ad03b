#!/bin/bash
# GPU session: file-path + CLI tests, the CLI file to file (configs[1]) on 5 M and 25 M records
set -u
cd "$(dirname "$0")/.."
tag=${1:-r2l}
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > "$out/box.txt" 2>&1
timeout 900 python -m pytest tests/test_gpu_y_filefast.py tests/test_gpu_cli.py -q -m gpu --maxfail=6 > "$out/tests_file.log" 2>&1; tail -5 "$out/tests_file.log"
timeout 600 python scripts/run_cfg2_cli.py --host-io --runs 3 > "$out/cfg2_cli.json" 2> "$out/cfg2_cli.err"; grep "stage seconds\|ms on the GPU\|start-up" "$out/cfg2_cli.err"; cat "$out/cfg2_cli.json"; echo
timeout 900 python scripts/run_cfg2_cli.py --records 25000000 --runs 2 > "$out/cfg2_cli_25m.json" 2> "$out/cfg2_cli_25m.err"; grep "stage seconds\|ms on the GPU\|start-up" "$out/cfg2_cli_25m.err"; cat "$out/cfg2_cli_25m.json"; echo
