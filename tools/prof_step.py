"""Three eager cfg2 training steps (FusionModel forward + loss + backward) for ncu captures:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python tools/prof_step.py [workload]
    ncu --set full --clock-control none --import-source on -k regex:<kernels> -c <n> -o gpurun_out/prof python tools/prof_step.py [workload]
The LAST step's launches are the ones to read (the first two pay first-touch costs)."""
import contextlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200")]
import torch
import bench
from immtsf import synth

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = dict(ttf=w["ttf"], mmf=w["mmf"], d_txt=w["d_txt"], C=w["C"], H=w["H"], kappa=w["kappa"])
with contextlib.redirect_stdout(sys.stderr):
    fm = synth.build_model(cfg, w["d_model"], dropout=bench.DROPOUT, seed=1)
fm.train()
d_in = [t.cuda() for t in bench.make_batch(w, 1234)[:4]]
for i in range(steps):
    for p in fm.parameters():
        p.grad = None
    out = fm(d_in[0], d_in[1], d_in[2], d_in[3].detach().requires_grad_(True))
    out.square().mean().backward()
    torch.cuda.synchronize()
    print("step", i, "done", file=sys.stderr)
