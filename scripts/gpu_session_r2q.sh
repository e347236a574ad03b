#!/bin/bash
# Round-2 profiling session (about 3 minutes): where K1's instructions go, measured.
#   gpurun --timeout 280 -- 'bash scripts/gpu_session_r2q.sh'
# 1. ncu --set full --import-source on of one k1_tok launch -> per-function / per-line instruction and stall shares
# 2. launch list (gpu__time_duration) of the --treat_multiallelics kernels on one dense contig
# 3. the same source-level capture of k3_heap, if the budget lasts
set -u
cd "$(dirname "$0")/.."
out=gpurun_out/r2q
mkdir -p "$out"
capture() {  # kernel regex, timeout
  local k=$1
  timeout "$2" ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o "$out/prof_$k" \
      python bench.py --records 4000000 --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > "$out/ncu_$k.log" 2>&1
  echo "capture $k rc=$?"
  ncu -i "$out/prof_$k.ncu-rep" --page raw --csv > "$out/prof_${k}_raw.csv" 2>/dev/null
  ncu -i "$out/prof_$k.ncu-rep" --page source --csv --print-source cuda,sass > "$out/prof_${k}_source.csv" 2>/dev/null
  python scripts/ncu_summary.py "$out/prof_${k}_raw.csv" > "$out/prof_${k}_summary.txt" 2>&1
  python scripts/ncu_src_funcs.py "$out/prof_${k}_source.csv" 60 > "$out/prof_${k}_functions.txt" 2>&1
  head -30 "$out/prof_${k}_functions.txt"
  ls -la "$out/prof_$k.ncu-rep"
  gzip -f "$out/prof_${k}_source.csv"
}
capture k1_tok 170
timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^ma_ --csv --log-file "$out/ma_launches.csv" \
    python scripts/bench_multiallelic.py 400000 > "$out/ma_under_ncu.json" 2> "$out/ma_under_ncu.err"
echo "ma list rc=$?"; tail -20 "$out/ma_launches.csv" | cut -c1-200
capture k3_heap 110
