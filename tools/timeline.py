"""Kernel timeline of ONE replay of the captured training step (runtime.GraphedStep) from CUPTI activity records
(torch.profiler; there is no nsys in this image): per kernel its stream, start and duration, the idle gaps of the main
stream and the per-stream busy time.  Answers "what is on the critical path of the step" -- ncu's launch list cannot
(it serialises the launches).

    python tools/timeline.py [--workload cfg2] [--out gpurun_out/timeline_cfg2.txt]
    torchrun --nproc-per-node 8 tools/timeline.py --workload cfg2     # N > 1: rank 0 writes its own timeline (NCCL kernels included)
"""
import argparse
import contextlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200")]
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--out", default=None)
    ap.add_argument("--batch", type=int, default=0)
    args = ap.parse_args()
    from immtsf import runtime, synth
    import torch.distributed as dist

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    w = dict(bench.WORKLOADS[args.workload])
    if args.batch:
        w["B"] = args.batch
    cfg = dict(ttf=w["ttf"], mmf=w["mmf"], d_txt=w["d_txt"], C=w["C"], H=w["H"], kappa=w["kappa"])
    with contextlib.redirect_stdout(sys.stderr):
        fm = synth.build_model(cfg, w["d_model"], dropout=bench.DROPOUT, seed=1)
    fm.train()
    d_in = [t.to(dev) for t in bench.make_batch(w, 1234 + rank)[:4]]
    step = runtime.GraphedStep(fm, example=d_in, warmup=2, allreduce_group=True if world > 1 else None)
    for _ in range(5):
        step(*d_in)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            step(*d_in)
        torch.cuda.synchronize()
    if rank == 0:
        path = f"/tmp/immtsf_trace_{os.getpid()}.json"
        prof.export_chrome_trace(path)
        ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
        ev.sort(key=lambda e: e["ts"])
        # split into replays: the largest gaps between consecutive events separate them; take the LAST replay
        gaps = sorted(((ev[i + 1]["ts"] - (ev[i]["ts"] + ev[i]["dur"]), i) for i in range(len(ev) - 1)), reverse=True)[:2]
        cut = max(i for _, i in gaps) + 1
        one = ev[cut:]
        t0 = one[0]["ts"]
        t_end = max(e["ts"] + e["dur"] for e in one)
        streams = sorted({e["args"].get("stream", -1) for e in one})
        lines = [f"# {w['name']}  world {world}  one graph replay: {len(one)} device activities, span {t_end - t0:.1f} us, "
                 f"summed durations {sum(e['dur'] for e in one):.1f} us, streams {streams}"]
        busy = {s: 0.0 for s in streams}
        for e in one:
            busy[e["args"].get("stream", -1)] += e["dur"]
        lines.append("# busy us per stream: " + ", ".join(f"{s}: {b:.1f}" for s, b in busy.items()))
        main_s = max(busy, key=busy.get)
        last_end = {s: t0 for s in streams}
        lines.append("#  start_us   dur_us  gap_before_us(same stream)  stream  name")
        for e in one:
            s = e["args"].get("stream", -1)
            gap = e["ts"] - last_end[s]
            last_end[s] = e["ts"] + e["dur"]
            lines.append(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} {gap:8.1f}  {'*' if s == main_s else ' '}{s:<4} {e['name'][:110]}")
        # union of busy intervals over all streams = time the GPU had at least one kernel resident
        iv = sorted((e["ts"], e["ts"] + e["dur"]) for e in one)
        cov, cur_s, cur_e = 0.0, iv[0][0], iv[0][1]
        for a, b in iv[1:]:
            if a > cur_e:
                cov += cur_e - cur_s
                cur_s, cur_e = a, b
            else:
                cur_e = max(cur_e, b)
        cov += cur_e - cur_s
        lines.append(f"# GPU non-idle (union over streams) {cov:.1f} us of {t_end - t0:.1f} us span; idle {t_end - t0 - cov:.1f} us")
        txt = "\n".join(lines)
        out = args.out or os.path.join(ROOT, "gpurun_out", f"timeline_{args.workload}_n{world}.txt")
        os.makedirs(os.path.dirname(out), exist_ok=True)
        open(out, "w").write(txt + "\n")
        print("\n".join(lines[:4]))
        print(lines[-1])
    if world > 1:
        step = None
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
