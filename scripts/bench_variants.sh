#!/bin/bash
# A/B of the experimental kernel variants on one GPU (run under gpurun):
#   make -C variantcalling_b200/csrc variants      (here, before the call: the .so files travel)
#   gpurun -- 'bash scripts/bench_variants.sh'
# For the default build and every variantcalling_b200/variants/*.so: the quick parity files (must pass),
# then bench.py without the e2e / CPU legs; one JSON line each in gpurun_out/variants/.
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/variants
for lib in default variantcalling_b200/variants/*.so; do
    name=$(basename "$lib" .so)
    if [ "$lib" = default ]; then unset UGVC_LIB_PATH; else export UGVC_LIB_PATH="$PWD/$lib"; fi
    if timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu \
         > "gpurun_out/variants/$name.tests.log" 2>&1; then
        timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline \
            > "gpurun_out/variants/$name.json" 2> "gpurun_out/variants/$name.err"
        python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/variants/{name}.json"))
    st = d["roofline"]["stage_ms_per_launch"]
    print(f"{name:32s} value {d['value'] / 1e6:8.1f} M/s  ms/step {d['ms_per_step']:7.2f}  k1 {st['k1_field_parse']:.3f} ms  k3 {st['k3_inference']:.3f} ms")
except Exception as e:  # noqa: BLE001
    print(f"{name:32s} bench failed: {e}")
PY
    else
        echo "$name: parity tests FAILED (see gpurun_out/variants/$name.tests.log)"
    fi
done
