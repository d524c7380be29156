#!/bin/bash
# One GPU call that evaluates the RecAvg variants which are compiled in but off by default (written after round 1's GPU
# budget was spent): correctness first (env-gated A/B tests), then the HBM sweep with each switch, then cfg1 bench lines.
#   gpurun --timeout 600 -- 'bash tools/ab_recavg_experimental.sh'     (about 5 minutes of box time)
# Outputs under gpurun_out/ab_recavg_*.  Switches (read per call by csrc/recavg.cu):
#   IMMTSF_RECAVG_MASKBIT=1  dropout keep flags in the mantissa LSB of the saved E_raw (backward skips Philox)
#   IMMTSF_RECAVG_SKIPQ=1    one-launch backward skips the Q accumulators of half passes whose c_nt are all zero
#   IMMTSF_RECAVG_FWD_PERSIST=1  persistent forward, segments double-buffered (two-stage ring of bulk copies)
#   IMMTSF_RECAVG_BWD_PIPE=1  warp-specialised one-launch backward (row warps / note warps, two-stage dS ring)
#   IMMTSF_RECAVG_FUSED_BWD=0|4|8  two-kernel backward | one launch with 4 / 8 notes per pass (default 8)
set -u
mkdir -p gpurun_out
# one pytest process per variant: a hung kernel (mbarrier deadlock) only costs its own 60 s
for t in keep_flags_in_e_raw_lsb skips_zero_sensitivity_passes persistent_forward_is_bit_identical bwd_pipe_equals_default; do
  IMMTSF_EXPERIMENTAL=1 timeout -s KILL 60 python -m pytest tests/test_gpu_experimental.py -m gpu -q --tb=short -x -k $t > gpurun_out/ab_recavg_tests_$t.log 2>&1
  echo "== tests $t rc=$? (137 = killed: hang)"; tail -2 gpurun_out/ab_recavg_tests_$t.log
done
# sweeps below: a variant whose tests failed or hung above is meaningless (and may hang again: each run has its own timeout)
for cfg in "base" "MASKBIT=1" "SKIPQ=1" "FWD_PERSIST=1" "MASKBIT=1 SKIPQ=1 FWD_PERSIST=1" "MASKBIT=1 SKIPQ=1 FUSED_BWD=4" "BWD_PIPE=1" "BWD_PIPE=1 MASKBIT=1 SKIPQ=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  envs=""
  for kv in $cfg; do [ "$kv" != "base" ] && envs="$envs IMMTSF_RECAVG_$kv"; done
  env $envs timeout -s KILL 60 python tools/sweep_hbm.py --out gpurun_out/ab_recavg_sweep_$tag.json > gpurun_out/ab_recavg_sweep_$tag.log 2>&1
  echo "== $cfg (rc=$?)"; grep "recavg_pool" gpurun_out/ab_recavg_sweep_$tag.log | cut -c1-170
done
for cfg in "base" "MASKBIT=1 SKIPQ=1 FWD_PERSIST=1"; do
  tag=$(echo "$cfg" | tr ' =' '__')
  envs=""
  for kv in $cfg; do [ "$kv" != "base" ] && envs="$envs IMMTSF_RECAVG_$kv"; done
  env $envs timeout -s KILL 90 python bench.py --workload cfg1 --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ab_recavg_bench_cfg1_$tag.json 2> gpurun_out/ab_recavg_bench_cfg1_$tag.err
  echo "== cfg1 $cfg (rc=$?)"; cut -c1-140 gpurun_out/ab_recavg_bench_cfg1_$tag.json
done
