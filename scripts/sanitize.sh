#!/bin/bash
# compute-sanitizer passes over the hot path: memcheck, racecheck, synccheck, initcheck on
#   (a) __graft_entry__.smoke()  -- greenlist build, decode engine (skinny GEMMs with the flag-carrying split-K hand-off,
#       attention, fused sampler), VQGAN decode (tcgen05 conv), detector;
#   (b) the RAR engine at the "tiny" golden shapes.
# Summaries go to gpurun_out/sanitize_<tool>.log; copy them to profiles/ to have them judged.
#   gpurun --timeout 1500 -- 'bash scripts/sanitize.sh'
set -u
mkdir -p gpurun_out
SAN=/usr/local/cuda/bin/compute-sanitizer
TOOLS="${1:-memcheck racecheck synccheck initcheck}"
for tool in $TOOLS; do
  extra=""
  [ "$tool" = "memcheck" ] && extra="--leak-check no"
  [ "$tool" = "racecheck" ] && extra="--racecheck-report all"
  for target in smoke rar_tiny; do
    log=gpurun_out/sanitize_${tool}_${target}.log
    if [ "$target" = "smoke" ]; then
      cmd="import __graft_entry__ as g; g.smoke()"
    else
      cmd="import sys; sys.path.insert(0, 'tests'); import test_gpu_rar as t; t.test_rar_engine_matches_reference_golden('tiny'); print('rar tiny ok')"
    fi
    echo "== $tool / $target" | tee $log
    timeout 900 $SAN --tool $tool $extra --print-limit 20 python -c "$cmd" >> $log 2>&1
    echo "exit code $?" >> $log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|rar tiny ok|exit code|Error|hazard" $log | head -12
  done
done
