#!/bin/bash
# run_scaling.sh <N> <tag> [workloads...]: bench.py at N GPUs of one box (torchrun for N > 1), one JSON line per workload under gpurun_out/.
N=$1; tag=$2; shift; shift
WL=${@:-cfg2}
mkdir -p gpurun_out
for w in $WL; do
  extra=""
  case $w in *:strong) extra="--scaling strong"; w=${w%:strong};; esac
  out=gpurun_out/${tag}_${w}${extra:+_strong}_n${N}
  if [ "$N" = "1" ]; then
    python bench.py --workload $w --gpus 1 --no-cpu-baseline --no-gpu-baseline --no-hbm $extra > $out.json 2> $out.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --workload $w --gpus $N --no-cpu-baseline --no-gpu-baseline --no-hbm $extra > $out.json 2> $out.err
  fi
  echo "== $w $extra N=$N rc=$?"; python - <<PY
import json
try:
    d=json.loads(open("$out.json").read().strip().splitlines()[-1])
    print("value %.0f  ms/step %.4f  e2e %.0f  rounds %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], [round(x,4) for x in d["config"]["round_ms_per_step"]]))
except Exception as e:
    print("no line:", e); print(open("$out.err").read()[-800:])
PY
done
