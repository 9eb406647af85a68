#!/bin/bash
# compute-sanitizer sweep over the update kernels (run on the GPU box: gpurun -- 'bash tools/sanitize_all.sh').
# One log per (tool, back-end) under gpurun_out/sanitizer/; the summary lines are collected into summary.txt.
mkdir -p gpurun_out/sanitizer
out=gpurun_out/sanitizer
run() {   # name tool env... -- args...   (ONLY=<substring> restricts the sweep to matching run names)
  local name=$1 tool=$2; shift 2
  if [ -n "$ONLY" ] && [[ "$name" != *"$ONLY"* ]]; then return; fi
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done
  shift
  local t0=$(date +%s)
  env "${envs[@]}" timeout 300 compute-sanitizer --tool "$tool" --print-limit 20 python tools/sanitize_run.py "$@" > "$out/${tool}_${name}.log" 2>&1
  local rc=$?
  echo "== ${tool} ${name}: rc=$rc $(( $(date +%s) - t0 ))s :: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|single ok|ensemble ok' "$out/${tool}_${name}.log" | tr '\n' ' ')" | tee -a "$out/summary.txt"
}
: > "$out/summary.txt"
nvcc --version | tail -1 >> "$out/summary.txt"
for tool in ${TOOLS:-memcheck racecheck synccheck}; do
  run tcc_b256       $tool HIRL_TCC=1 HIRL_CLUSTER=0 -- --batch 256
  run tcc_b100       $tool HIRL_TCC=1 HIRL_CLUSTER=0 -- --batch 100
  run fp32_b64       $tool HIRL_TCC=0 HIRL_TC=0 HIRL_CLUSTER=0 -- --batch 64
  run g1tc_b256      $tool HIRL_TCC=0 HIRL_TC=2 HIRL_CLUSTER=0 -- --batch 256
  run fulltc_b256    $tool HIRL_TCC=0 HIRL_TC=1 HIRL_CLUSTER=0 -- --batch 256
  run dp1_b64        $tool HIRL_TCC=0 HIRL_TC=0 HIRL_CLUSTER=0 -- --batch 64 --dp
  run cluster_b64    $tool HIRL_TCC=0 HIRL_CLUSTER=1 -- --batch 64
  run ensemble_b128  $tool HIRL_CLUSTER=0 -- --batch 128 --ensemble 3 --steps 4
done
run g1tc_b256 initcheck HIRL_TCC=0 HIRL_TC=2 HIRL_CLUSTER=0 -- --batch 256
run tcc_b256 initcheck HIRL_TCC=1 HIRL_CLUSTER=0 -- --batch 256
# the stage_g8-as-a-separate-function build (update.cuh: HIRL_G8_NOINLINE), when present: does the round-1 fault reproduce?
if [ -f hirl4ucav_b200/_lib/libhirl_g8noinline.so ]; then
  for tool in memcheck racecheck initcheck; do
    run g8noinline_fp32_b64  $tool HIRL_LIB_PATH=$PWD/hirl4ucav_b200/_lib/libhirl_g8noinline.so HIRL_TCC=0 HIRL_TC=0 HIRL_CLUSTER=0 -- --batch 64
    run g8noinline_g1tc_b256 $tool HIRL_LIB_PATH=$PWD/hirl4ucav_b200/_lib/libhirl_g8noinline.so HIRL_TCC=0 HIRL_TC=2 HIRL_CLUSTER=0 -- --batch 256
  done
fi
cat "$out/summary.txt"
