#!/bin/bash
# Round-end check on one B200: whole GPU suite, smoke(), the default bench line (cfg2, all secondary records), the cfg1 line, and the
# ncu launch list of one eager cfg2 / cfg1 step (shares only: per-launch times under ncu are cold-cache and serialised).
tag=${1:-r2final}
mkdir -p gpurun_out
echo "== GPU suite"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
echo "== bench cfg2 (default line)"; timeout 900 python bench.py > gpurun_out/${tag}_bench_cfg2.json 2> gpurun_out/${tag}_bench_cfg2.err; echo rc=$?
echo "== bench cfg1"; timeout 600 python bench.py --workload cfg1 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_bench_cfg1.json 2> gpurun_out/${tag}_bench_cfg1.err; echo rc=$?
python - <<PY
import json
for w in ("cfg2", "cfg1"):
    try:
        d = json.loads(open("gpurun_out/${tag}_bench_%s.json" % w).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(w, "value %.0f ms/step %.4f e2e %.0f | roofline %s frac %.3f share %.2f" % (d["value"], d["ms_per_step"], d["e2e"]["value"], r["bound"], r["frac"], r.get("kernel_share_of_step", 0)),
              "| 3xtf32 frac", r.get("frac_of_3xtf32_ceiling"), "| api varying", d["config"].get("public_api_varying_shapes_samples_per_s"))
        for h in (d.get("roofline_hbm") or {}).get("records", []):
            if h["B"] == 2048 and "recavg" in h["kernel"]:
                print("   hbm", h["kernel"], "%.1f us frac %.3f" % (h["us"], h["frac"]))
        print("   gpu_eager", (d.get("gpu_eager_baseline") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(w, "no line:", e)
PY
for w in cfg2 cfg1; do
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches_$w.csv python tools/prof_step.py $w > /dev/null 2>&1
  n=$(grep -c "gpu__time_duration.sum" gpurun_out/${tag}_launches_$w.csv)
  python tools/launch_shares.py gpurun_out/${tag}_launches_$w.csv $((n*2/3)) > gpurun_out/${tag}_launch_shares_$w.txt 2>&1
  head -12 gpurun_out/${tag}_launch_shares_$w.txt
done
