#!/bin/bash
# Round-2 GPU session: suite, bench (both K1 tiers vs the generic parser alone), launch list, full captures.
#   gpurun --timeout 1700 -- 'bash scripts/gpu_session_r2.sh <tag>'
set -u
cd "$(dirname "$0")/.."
tag=${1:-r2a}
out=gpurun_out/$tag
mkdir -p "$out"
bash scripts/gpu_probe.sh > /dev/null 2>&1; cp gpurun_out/box.txt "$out/box.txt"
if [ "${SKIP_TESTS:-0}" != 1 ]; then
  timeout 900 python -m pytest tests -q -m gpu --maxfail=6 > "$out/tests.log" 2>&1; tail -3 "$out/tests.log"
fi
timeout 800 python bench.py ${BENCH_ARGS:-} > "$out/bench_n1.json" 2> "$out/bench_n1.err"; tail -c 1500 "$out/bench_n1.json"; echo
if [ "${SKIP_LEGACY:-0}" != 1 ]; then
  UGVC_K1_LEGACY=1 timeout 300 python bench.py --no-e2e --no-cpu-baseline --steps 3 > "$out/bench_legacy.json" 2> "$out/bench_legacy.err"
  python - <<PY
import json
for n in ("bench_n1", "bench_legacy"):
    try:
        d = json.load(open("$out/%s.json" % n)); print(n, "value %.1f M/s" % (d["value"] / 1e6), d["roofline"]["stage_ms_per_launch"])
    except Exception as e: print(n, "failed", e)
PY
fi
if [ "${SKIP_NCU:-0}" != 1 ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file "$out/launches.csv" \
      python bench.py --records 8000000 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > "$out/ncu_launch.log" 2>&1
  for k in ${NCU_KERNELS:-k1_fast k3_heap}; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o "$out/prof_$k" \
        python bench.py --records 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > "$out/ncu_$k.log" 2>&1
    ncu -i "$out/prof_$k.ncu-rep" --page raw --csv > "$out/prof_${k}_raw.csv" 2>/dev/null
    python scripts/ncu_summary.py "$out/prof_${k}_raw.csv" | tee "$out/prof_${k}_summary.txt"
  done
fi
