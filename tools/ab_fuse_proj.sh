for v in 0 1; do
  IMMTSF_FUSE_PROJ=$v python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>/dev/null > gpurun_out/s4_ab_$v.json
done
