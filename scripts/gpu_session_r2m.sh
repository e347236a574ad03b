#!/bin/bash
# GPU session: source-level captures of K1 (k1_fast) and K3 (k3_heap), then the tile-shape A/B of K1
set -u
cd "$(dirname "$0")/.."
tag=${1:-r2m}
out=gpurun_out/$tag
mkdir -p "$out"
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > "$out/box.txt" 2>&1
for k in ${NCU_KERNELS:-k1_fast k3_heap}; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o "$out/prof_$k" \
      python bench.py --records 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > "$out/ncu_$k.log" 2>&1
  ncu -i "$out/prof_$k.ncu-rep" --page raw --csv > "$out/prof_${k}_raw.csv" 2>/dev/null
  python scripts/ncu_summary.py "$out/prof_${k}_raw.csv" | tee "$out/prof_${k}_summary.txt"
done
for lib in default variantcalling_b200/variants/*.so; do
    name=$(basename "$lib" .so)
    if [ "$lib" = default ]; then unset UGVC_LIB_PATH; else export UGVC_LIB_PATH="$PWD/$lib"; fi
    timeout 200 python bench.py --records 16000000 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline \
        > "$out/var_$name.json" 2> "$out/var_$name.err"
    python - "$out/var_$name.json" "$name" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); st = d["roofline"]["stage_ms_per_launch"]
    print(f"{sys.argv[2]:28s} value {d['value']/1e6:8.1f} M/s ms/step {d['ms_per_step']:7.2f} k1 {st['k1_field_parse']:.3f} k3 {st['k3_inference']:.3f} slow {d['config'].get('k1_slow_records_last_batch')}")
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
done
