#!/bin/bash
# compute-sanitizer passes over the hot path: memcheck, racecheck, synccheck, initcheck on
#   (a) __graft_entry__.smoke()  -- greenlist build, decode engine (skinny GEMMs with the flag-carrying split-K hand-off,
#       attention, fused sampler), VQGAN decode (tcgen05 conv), detector;
#   (b) the RAR engine at the "tiny" golden shapes;
#   (c) round 2: a small Taming VQGAN in bf16x3 mode (persistent tcgen05 conv incl. stride 2, conv_in3 / conv_out3) and
#       the two-lane wrapper test (TMA weight ring in the GEMMs, two engines on two streams).
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
  for target in ${TARGETS:-smoke rar_tiny vqgan_bf16x3 lanes}; do
    log=gpurun_out/sanitize_${tool}_${target}.log
    if [ "$target" = "smoke" ]; then
      cmd="import __graft_entry__ as g; g.smoke()"
    elif [ "$target" = "vqgan_bf16x3" ]; then
      # round 2: persistent bf16x3 tcgen05 conv (stride 1 and 2), conv_in3 / conv_out3 kernels, GroupNorm, on a small Taming VQGAN
      cmd="import sys; sys.path.insert(0, 'tests'); import torch, test_gpu_vqgan as t; from wmar_b200.models.vqgan_engine import VQGANEngine; ov, ocfg, ecfg, w = t._taming(dict(ch=128, ch_mult=(1, 2, 2), resolution=64, attn_resolutions=(16,), n_embed=1024), 3); e = VQGANEngine(w, ecfg, max_batch=2, precision='bf16x3'); c = torch.randint(0, 1024, (2, 256)).cuda(); img = e.decode(c); back = e.encode(img); torch.cuda.synchronize(); print('vqgan bf16x3 ok', tuple(img.shape), float((back == c).float().mean()))"
    elif [ "$target" = "lanes" ]; then
      cmd="import sys; sys.path.insert(0, 'tests'); import test_gpu_watermark as t; t.test_wrapper_lanes_equal_sequential_chunk_loop(); print('lanes ok')"
    else
      cmd="import sys; sys.path.insert(0, 'tests'); import test_gpu_rar as t; t.test_rar_engine_matches_reference_golden('tiny'); print('rar tiny ok')"
    fi
    echo "== $tool / $target" | tee $log
    timeout 900 $SAN --tool $tool $extra --print-limit 20 python -c "$cmd" >> $log 2>&1
    echo "exit code $?" >> $log
    grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|rar tiny ok|vqgan bf16x3 ok|lanes ok|exit code|Error|hazard" $log | head -12
  done
done
