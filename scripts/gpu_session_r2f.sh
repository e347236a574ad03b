#!/bin/bash
# GPU session: suite + the CLI file to file (configs[1], device-side file path vs --host_io)
set -u
cd "$(dirname "$0")/.."
tag=${1:-r2f}
out=gpurun_out/$tag
mkdir -p "$out"
timeout 900 python -m pytest tests -q -m gpu --maxfail=6 > "$out/tests.log" 2>&1; tail -3 "$out/tests.log"
timeout 900 python -m pytest tests/test_gpu_y_filefast.py tests/test_gpu_parity.py -q -m gpu -s > "$out/tests2.log" 2>&1; grep -v Warn "$out/tests2.log" | grep "passed\|failed\|identical\|Error" | head -20
timeout 600 python scripts/run_cfg2_cli.py --host-io > "$out/cfg2_cli.json" 2> "$out/cfg2_cli.err"; grep "stage seconds" "$out/cfg2_cli.err"; cat "$out/cfg2_cli.json"; echo
