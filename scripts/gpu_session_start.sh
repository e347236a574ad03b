#!/bin/bash
# First GPU call of a session: everything that was prepared on the CPU and still waits for a device.
#   make -C variantcalling_b200/csrc all variants    (before the call; built .so files travel)
#   gpurun --timeout 1500 -- 'bash scripts/gpu_session_start.sh'
# Results land in gpurun_out/session/ (merged back by gpurun).
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/session
mkdir -p "$out"
# 1. the whole GPU suite, without -x: every file reports (the x_ files have only run on the host emulation so far)
timeout 600 python -m pytest tests -q -m gpu > "$out/tests.log" 2>&1; tail -3 "$out/tests.log"
# 2. the headline bench line
timeout 400 python bench.py > "$out/bench_n1.json" 2> "$out/bench_n1.err"; tail -c 600 "$out/bench_n1.json"; echo
# 2b. end to end with BGZF-compressed host buffers inflated on the device (first GPU timing of csrc/inflate.cuh)
timeout 600 python bench.py --e2e-bgzf --no-cpu-baseline > "$out/bench_bgzf.json" 2> "$out/bench_bgzf.err"; grep e2e_bgzf "$out/bench_bgzf.err"; python -c "
import json; d=json.load(open('$out/bench_bgzf.json')); print('e2e', d['e2e']['value']/1e6, 'M/s  e2e_bgzf', d.get('e2e_bgzf'))"
# 3. A/B of the experimental K1 variants (NEGFAST, INLINE_DICT1, both, SPLIT)
timeout 900 bash scripts/bench_variants.sh 2>&1 | tee "$out/variants.txt"
# 4. configs[1] file to file through the CLI
timeout 200 python scripts/run_cfg2_cli.py > "$out/cfg2_cli.json" 2> "$out/cfg2_cli.err"; grep "stage seconds" "$out/cfg2_cli.err"; cat "$out/cfg2_cli.json"; echo
