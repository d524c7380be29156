#!/bin/bash
# Round 2, fourth RecAvg A/B call: forward VAR 0 vs 3, backward Philox key variants, tensor-core path with producer-written lo
# operands (forced for every large cell vs off), new parity tests, whole GPU suite.
mkdir -p gpurun_out
echo "== new / changed tests"
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity_full.py -q -x -k "recavg_bwd_fused_equals or tensor_core_path or long_segments" 2>&1 | tail -3
for v in 0 3; do
  IMMTSF_RECAVG_FWD_VAR=$v IMMTSF_RECAVG_BWD_PH=$((v/3)) timeout 200 python tools/sweep_hbm.py --only-recavg --out gpurun_out/r2d_sweep_fwdvar${v}_bwdph$((v/3)).json > /dev/null 2>&1
  echo "== sweep FWD_VAR=$v BWD_PH=$((v/3)) rc=$?"
done
for tc in 0 1; do
  IMMTSF_RECAVG_TC=$tc timeout 300 python tools/sweep_hbm.py --only-large --out gpurun_out/r2d_sweep_large_tc$tc.json > gpurun_out/r2d_sweep_large_tc$tc.log 2>&1
  echo "== large sweep TC=$tc rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_sweep_*.json")):
    d = json.load(open(f))
    print(f.split("/")[-1], " | ".join("%s B%d N%d T%d %.1f us %.3f %.1fTF" % (r["kernel"][-3:], r["B"], r["N_max"], r["T"], r["ms"] * 1e3, r["frac"], r.get("pool_gflops", 0) / 1e3) for r in d["rows"] if "recavg" in r["kernel"]))
PY
echo "== whole GPU suite (defaults)"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
