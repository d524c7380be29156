#!/bin/bash
# A/B of the one-launch RecAvg backward variants on one B200: IMMTSF_RECAVG_FUSED_BWD=8 (note phase on CUDA cores) against
# =16 (note phase on mma.sync 3xTF32, padded dS rows, conflict-free rows phase, L2 prefetch); then the whole GPU suite with =16.
mkdir -p gpurun_out
echo "== A/B test (modes 0 / 8 / 4 / 16 in one process)"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "recavg_bwd_fused_equals" 2>&1 | tail -4
for m in 8 16; do
  IMMTSF_RECAVG_FUSED_BWD=$m timeout 200 python tools/sweep_hbm.py --only-recavg --out gpurun_out/r2_ab_mma2_sweep_$m.json > /dev/null 2>&1
  echo "== sweep FUSED_BWD=$m rc=$?"
done
IMMTSF_RECAVG_FUSED_BWD=16 IMMTSF_RECAVG_MMA_NOPF=1 timeout 200 python tools/sweep_hbm.py --only-recavg --out gpurun_out/r2_ab_mma2_sweep_16_nopf.json > /dev/null 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_ab_mma2_sweep_*.json")):
    d = json.load(open(f))
    print(f.split("/")[-1], " | ".join("%s B%d N%d T%d%s %.1f us %.3f" % (r["kernel"][-3:], r["B"], r["N_max"], r["T"], "", r["ms"] * 1e3, r["frac"]) for r in d["rows"] if "recavg" in r["kernel"]))
PY
echo "== whole GPU suite with IMMTSF_RECAVG_FUSED_BWD=16"
IMMTSF_RECAVG_FUSED_BWD=16 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
echo "== ncu of the mma backward"
IMMTSF_RECAVG_FUSED_BWD=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:recavg_bwd_mma --launch-skip 2 -c 1 -o gpurun_out/r2_recavg_bwd_mma2 -f python tools/prof_recavg.py > gpurun_out/r2_recavg_bwd_mma2.log 2>&1
echo "ncu rc=$?"
