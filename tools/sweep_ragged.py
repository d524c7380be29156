#!/usr/bin/env python
"""Ragged-length sweep (BASELINE.json configs[3]): N_max in {1..1024} texts per sample x T in {1..4096} query times,
RecAvg vs the single-query T2V attention (active reference module) vs the per-(note, query) T2V attention
(TTF_T2V_XAttn_old semantics), each composed with MMF_GR_Add, forward + backward, dropout 0.1, ragged N_i ~ U{1..N_max}.
d = 768, C = 8; B = the largest power of two <= 256 with B*T <= 65536 and B*N_max <= 65536.  CUDA events, eager launches,
median of `--iters`; inputs resident.  Writes a JSON document of rows (samples/s and ms per step) to --out."""
import argparse, json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch
import gpu_common as G


def pick_B(N, T):
    B = 256
    while B > 1 and (B * T > 65536 or B * N > 65536):
        B //= 2
    return B


def time_step(fm, batch, iters):
    notes, tau, t_hat, Y = batch
    def step():
        fm.zero_grad(set_to_none=True)
        Yc = Y.clone().requires_grad_(True)
        fm(notes, tau, t_hat, Yc).square().mean().backward()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    ms = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); step(); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return statistics.median(ms)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--Ns", default="1,4,16,64,256,1024")
    ap.add_argument("--Ts", default="1,16,64,256,1024,4096")
    a = ap.parse_args()
    d, C = 768, 8
    rows = []
    os.environ.setdefault("IMMTSF_NAN_CHECK", "0")
    for N in [int(x) for x in a.Ns.split(",")]:
        for T in [int(x) for x in a.Ts.split(",")]:
            B = pick_B(N, T)
            notes, tau, t_hat, Y, _ = G.synth_batch(B, N, T, d, C, seed=N * 7 + T)
            batch = [t.cuda() for t in (notes, tau / 7.0, t_hat, Y)]
            for ttf in ("TTF_RecAvg", "TTF_T2V_XAttn", "TTF_T2V_XAttn_old"):
                cfg = dict(ttf=ttf, mmf="MMF_GR_Add", d_txt=d, C=C, H=1, kappa=0.5)
                row = dict(ttf=ttf, N_max=N, T=T, B=B)
                try:
                    fm = G.build_model(cfg, d, dropout=0.1, seed=1)
                    fm.train()
                    ms = time_step(fm, batch, a.iters)
                    row.update(ms_per_step=ms, samples_per_s=B / ms * 1e3)
                    del fm
                except Exception as e:  # unsupported shape: recorded, not hidden
                    row.update(error=f"{type(e).__name__}: {str(e)[:160]}")
                    torch.cuda.synchronize()
                rows.append(row)
                if a.out:  # rewritten after every cell: a cut-off run keeps what it measured
                    open(a.out, "w").write(json.dumps(dict(method="eager FusionModel fwd+bwd (TTF + MMF_GR_Add), dropout 0.1, CUDA events, median",
                                                           d=d, C=C, rows=rows), indent=1))
                print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in row.items()}, flush=True)
            del batch
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
