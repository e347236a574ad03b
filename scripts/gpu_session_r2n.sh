#!/bin/bash
# GPU session: the token form of the K1 tile kernel -- parity, race / memory check, stage times against the queue form,
# K3 shape A/B, source-level capture
set -u
cd "$(dirname "$0")/.."
tag=${1:-r2n}
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > "$out/box.txt" 2>&1
timeout 600 python -m pytest ${TESTS:-tests/test_gpu_y_tiletok.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_edges.py tests/test_gpu_y_lanes.py} -q -m gpu --maxfail=8 > "$out/tests.log" 2>&1; tail -4 "$out/tests.log"
if [ "${SKIP_SANITIZER:-0}" != 1 ]; then
  MODE=learn timeout 300 compute-sanitizer --tool racecheck --racecheck-report all python scripts/k1_tiers_check.py > "$out/racecheck.log" 2>&1; grep -c "Race reported\|hazard" "$out/racecheck.log"; tail -3 "$out/racecheck.log"
  grep -A2 "hazard detected" "$out/racecheck.log" | grep "Thread" | sed -E 's/Thread \([0-9]+,0,0\)//; s/\+0x[0-9a-f]+//' | sort | uniq -c | sort -rn | head -12
fi
run_bench() {  # name, env...
  local name=$1; shift
  env "$@" timeout 200 python bench.py --records 16000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > "$out/b_$name.json" 2> "$out/b_$name.err"
  python - "$out/b_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); st = d["roofline"]["stage_ms_per_launch"]
    print(f"{sys.argv[2]:28s} value {d['value']/1e6:8.1f} M/s ms/step {d['ms_per_step']:7.2f} k1 {st['k1_field_parse']:.3f} k3 {st['k3_inference']:.3f} slow {d['config'].get('k1_slow_records_last_batch')} counts {d.get('counts_last_steps')}")
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run_bench tok A=1
for k in ${NCU_KERNELS:-k1_tok}; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o "$out/prof_$k" \
      python bench.py --records 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > "$out/ncu_$k.log" 2>&1
  ncu -i "$out/prof_$k.ncu-rep" --page raw --csv > "$out/prof_${k}_raw.csv" 2>/dev/null
  python scripts/ncu_summary.py "$out/prof_${k}_raw.csv" | tee "$out/prof_${k}_summary.txt"
done
