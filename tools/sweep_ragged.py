#!/usr/bin/env python
"""Ragged-length sweep (BASELINE.json configs[3]): N_max in {1..1024} texts per sample x T in {1..4096} query times,
RecAvg vs the single-query T2V attention (active reference module) vs the per-(note, query) T2V attention
(TTF_T2V_XAttn_old semantics), each composed with MMF_GR_Add, forward + backward, dropout 0.1, ragged N_i ~ U{1..N_max}.
d = 768, C = 8; per-GPU B = the largest power of two <= 256 with B*T <= 65536 and B*N_max <= 65536.

Every cell is one runtime.GraphedStep (the whole step replayed as one CUDA graph: round 1 ran the sweep eagerly and its
small cells measured the host, ~1.5 ms per step).  Under torchrun (N GPUs of one box) every rank runs its own batch of
the cell (weak scaling) with the data-parallel all-reduce captured in the graph; times are the max over ranks.

    python tools/sweep_ragged.py --out profiles/r2_sweep_ragged_n1.json
    torchrun --nproc-per-node 8 tools/sweep_ragged.py --Ns 16,256 --Ts 16,256 --out profiles/r2_sweep_ragged_n8.json
"""
import argparse, contextlib, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200")]
import torch
import torch.distributed as dist


def pick_B(N, T):
    B = 256
    while B > 1 and (B * T > 65536 or B * N > 65536):
        B //= 2
    return B


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--Ns", default="1,4,16,64,256,1024")
    ap.add_argument("--Ts", default="1,16,64,256,1024,4096")
    ap.add_argument("--ttfs", default="TTF_RecAvg,TTF_T2V_XAttn,TTF_T2V_XAttn_old")
    a = ap.parse_args()
    from immtsf import runtime, synth

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d, C = 768, 8
    rows = []
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    for N in [int(x) for x in a.Ns.split(",")]:
        for T in [int(x) for x in a.Ts.split(",")]:
            B = pick_B(N, T)
            notes, tau, t_hat, Y, _ = synth.synth_batch(B, N, T, d, C, seed=N * 7 + T + 1000 * rank)
            batch = [t.to(dev) for t in (notes, tau / 7.0, t_hat, Y)]
            for ttf in a.ttfs.split(","):
                cfg = dict(ttf=ttf, mmf="MMF_GR_Add", d_txt=d, C=C, H=1, kappa=0.5)
                row = dict(ttf=ttf, N_max=N, T=T, B_per_gpu=B, n_gpus=world)
                step = None
                try:
                    with contextlib.redirect_stdout(sys.stderr):
                        fm = synth.build_model(cfg, d, dropout=0.1, seed=1)
                    fm.train()
                    step = runtime.GraphedStep(fm, example=batch, warmup=2, allreduce_group=True if world > 1 else None)
                    for _ in range(2):
                        step(*step.static_in)
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    ms = []
                    for _ in range(a.iters):
                        flush.fill_(1.0)
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record(); step(*step.static_in); e1.record(); e1.synchronize()
                        ms.append(e0.elapsed_time(e1))
                    med = statistics.median(ms)
                    if world > 1:
                        t = torch.tensor([med], dtype=torch.float64, device=dev)
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                        med = float(t.item())
                    row.update(ms_per_step=med, samples_per_s=B * world / med * 1e3)
                except Exception as e:  # unsupported shape: recorded, not hidden
                    row.update(error=f"{type(e).__name__}: {str(e)[:160]}")
                    torch.cuda.synchronize()
                finally:
                    if step is not None:
                        step.close()
                    step = fm = None
                rows.append(row)
                if rank == 0:
                    if a.out:  # rewritten after every cell: a cut-off run keeps what it measured
                        open(a.out, "w").write(json.dumps(dict(
                            method="runtime.GraphedStep replay of FusionModel fwd+bwd (TTF + MMF_GR_Add), dropout 0.1, resident inputs, CUDA "
                                   "events per replay with a 256 MiB L2 flush before each, median of %d, max over ranks" % a.iters,
                            d=d, C=C, n_gpus=world, rows=rows), indent=1))
                    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()}, flush=True)
            del batch
            torch.cuda.empty_cache()
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
