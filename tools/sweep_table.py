"""Render a tools/sweep_ragged.py JSON document as the N_max x T tables of profiles/r*_sweep_ragged_table*.txt."""
import json, sys
doc = json.load(open(sys.argv[1]))
rows = doc["rows"]
Ns = sorted({r["N_max"] for r in rows}); Ts = sorted({r["T"] for r in rows})
print(doc["method"]); print(f"d {doc['d']}, C {doc['C']}, GPUs {doc.get('n_gpus', 1)}; samples/s (whole job); per-GPU B per cell:",
      {f"N{r['N_max']}xT{r['T']}": r["B_per_gpu"] for r in rows if r["ttf"] == rows[0]["ttf"] and (r["N_max"] in (Ns[0], Ns[-1]))})
for ttf in dict.fromkeys(r["ttf"] for r in rows):
    print("\n" + ttf)
    print("  N_max \\ T" + "".join(f"{t:>10d}" for t in Ts))
    for n in Ns:
        line = f"  {n:9d}"
        for t in Ts:
            m = [r for r in rows if r["ttf"] == ttf and r["N_max"] == n and r["T"] == t]
            line += f"{m[0]['samples_per_s']:10.0f}" if m and "samples_per_s" in m[0] else f"{'err' if m else '-':>10}"
        print(line)
