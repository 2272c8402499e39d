#!/bin/bash
# The GPU-side recipes of this repository, one per argument (run under `gpurun -- 'bash tools/gpu_scripts/run.sh <what>'`).
#   suite   full GPU test suite
#   bench   full bench line (extras included) -> gpurun_out/bench.json
#   cfgs    the other BASELINE configs at full size + the published workload
#   phases  layer-tail phase timelines (cfg-2, 16 rows, GPT-XXL 64 rows)
#   vq      VQ index search: default + variants
#   micro   microbenchmarks behind DESIGN.md 4.1 (build them first: see the header of each .cu)
#   prof    ncu evidence: launch list, full captures of the step kernels and the VQ score kernel
#   pair    cfg-2 with and without CTA pairs in the layer tail (same box), phase timeline
#   switches  decode parity tests under every A/B switch
#   diff    diffusion-head sampler: tests, timing per row count, whole generate_video_diff loop, launch list, full ncu capture
set -u
mkdir -p gpurun_out
for what in "$@"; do
case "$what" in
suite)
  timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | cut -c1-250 ;;
bench)
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err ;;
cfgs)
  python tools/published_cfg.py 2>&1 | tail -4
  python tools/cfg_sweep.py 2>&1 | grep cfg- ;;
phases)
  python tools/lt_phases.py 288 2>&1 | tail -24 | tee gpurun_out/lt_phases.txt
  VLG_PH_BATCH=8 python tools/lt_phases.py 288 2>&1 | tail -22 | tee gpurun_out/lt_phases_b8.txt
  VLG_PH_MODEL=GPT-XXL VLG_PH_BATCH=32 python tools/lt_phases.py 288 2>&1 | tail -22 | tee gpurun_out/lt_phases_xxl_b32.txt ;;
vq)
  timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "vq or VQ or codebook" 2>&1 | tail -3 | cut -c1-200
  for v in 0 5 6 7; do VLG_VQ_VARIANT=$v python tools/vq_bench.py 2>&1 | tail -1 | cut -c90-290 | sed "s/^/variant $v: /"; done ;;
micro)
  for b in ingest_bench ingest_ldg_bench mma_issue_bench; do
    [ -x tools/micro/bin/$b ] && timeout 120 tools/micro/bin/$b > gpurun_out/$b.txt 2>&1 && tail -40 gpurun_out/$b.txt
  done ;;
prof)
  CMD="python bench.py --steps 1 --warmup 0 --new-tokens 320 --skip-extras"
  export VLG_PROFILE_STEP=286
  $CMD > gpurun_out/prof_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_step287.csv $CMD > gpurun_out/prof_ncu1.log 2>&1
  echo "launch list rc=$?"
  $CMD > gpurun_out/prof_plain2.log 2>&1 && ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"layer_tail_kernel|attn_decode_tma" -c 6 -f -o gpurun_out/ncu_full_step287 $CMD > gpurun_out/prof_ncu2.log 2>&1
  echo "full capture rc=$?"
  unset VLG_PROFILE_STEP
  python tools/vq_bench.py > gpurun_out/vq_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:vq_score -c 1 -f -o gpurun_out/vq_score python tools/vq_bench.py > gpurun_out/vq_ncu.log 2>&1
  echo "vq capture rc=$?" ;;
pair)
  timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras 2>/dev/null | cut -c1-150
  VLG_LT_PAIR=0 timeout 300 python bench.py --steps 3 --warmup 2 --skip-extras 2>/dev/null | cut -c1-150 | sed 's/^/PAIR=0 /'
  timeout 200 python tools/lt_phases.py 288 2>&1 | tail -24 | tee gpurun_out/lt_phases.txt ;;
switches)
  for sw in VLG_LT=0 VLG_LT_PAIR=0 VLG_ATTN=generic VLG_PDL=0 VLG_STEPPED_PREFILL=1 VLG_NO_GRAPH=1 VLG_FUSED_NORM=1; do
    env $sw timeout 600 python -m pytest tests/test_gpu_decode.py -q -x 2>&1 | tail -1 | cut -c1-120 | sed "s/^/$sw: /"
  done ;;
diff)
  timeout 420 python -m pytest tests/test_gpu_diffusion.py -q 2>&1 | tail -2 | cut -c1-200
  timeout 200 python tools/diff_bench.py 1 2 4 8 16 2>&1 | tail -5 | tee gpurun_out/diff_bench.txt
  timeout 300 python tools/diff_generate_bench.py 2>&1 | tail -1 | tee gpurun_out/diff_generate_bench.txt
  export DIFF_BENCH_ITERS=1 DIFF_BENCH_WARM=1 DIFF_BENCH_NO_EAGER=1
  python tools/diff_bench.py 1 8 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/diff_head_launches.csv python tools/diff_bench.py 1 8 > gpurun_out/diff_ncu1.log 2>&1
  echo "launch list rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:dh_sample -c 1 -f -o gpurun_out/diff_head_r1 python tools/diff_bench.py 1 > gpurun_out/diff_ncu2.log 2>&1
  echo "full capture rc=$?"
  unset DIFF_BENCH_ITERS DIFF_BENCH_WARM DIFF_BENCH_NO_EAGER ;;
*) echo "unknown recipe $what" ;;
esac
done
