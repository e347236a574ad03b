#!/bin/bash
# N-GPU session: bench.py at N ranks (both arms are the driver's job; here ours) and the TOOL itself under torchrun
# with NCCL, its output compared with the single-process output.
#   gpurun --gpus N --timeout 1500 -- 'bash scripts/gpu_session_multi.sh N tag'
set -u
cd "$(dirname "$0")/.."
N=${1:-2}
tag=${2:-r2_n$N}
out=gpurun_out/$tag
mkdir -p "$out"
W=/tmp/ugvc_multi
mkdir -p $W
nvidia-smi topo -m > "$out/topo.txt" 2>&1
if [ -z "${SKIP_BENCH:-}" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > "$out/bench_n$N.json" 2> "$out/bench_n$N.err"
fi
python - <<PY
import json
try:
    d = json.load(open("$out/bench_n$N.json")); print("N=$N value %.1f M/s e2e %.1f M/s" % (d["value"] / 1e6, d["e2e"]["value"] / 1e6), d["config"].get("numa_binding"))
except Exception as e: print("bench failed", e)
PY
# the tool: single process first (also builds the input), then N ranks
timeout 600 python scripts/run_cfg2_cli.py --records ${CLI_RECORDS:-8000000} --workdir $W --runs 1 > "$out/cli_single.json" 2> "$out/cli_single.err"
customs=$(python -c "from variantcalling_b200 import synth; print(' '.join('--custom_annotations ' + c for c in synth.custom_annotation_names(5)))")
t_begin=$(date +%s.%N)
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    ugvc/__main__.py filter_variants_pipeline --input_file $W/in.vcf.gz --model_file $W/model.pkl --output_file $W/out_multi.vcf.gz $customs \
    > "$out/cli_multi.log" 2>&1
echo "{\"n_ranks\": $N, \"records\": ${CLI_RECORDS:-8000000}, \"torchrun_wall_s\": $(python -c "import time; print(round(time.time() - $t_begin, 3))")}" | tee "$out/cli_multi_wall.json"
tail -3 "$out/cli_multi.log"
python scripts/check_same_vcf.py $W/out.vcf.gz $W/out_multi.vcf.gz | tee "$out/cli_multi_check.json"
grep -h "records written\|stage seconds\|NUMA\|device file path" "$out/cli_multi.log" | head -12
python -c "import json; d=json.load(open('$out/cli_single.json')); print('single', d['cli_wall_s'], d['variants_per_s_file_to_file'], d['checks'])"
