#!/usr/bin/env python
"""HBM-roofline sweep of the streaming kernels of the path (RecAvg and GR_Add are reported by
achieved GB/s): each kernel is timed alone with CUDA events, a 256 MiB L2 flush between iterations, on ragged
Time-IMM-shaped inputs; achieved = algorithmic bytes / time, against MEASURED_PEAKS.json's copy bandwidth.
Writes one JSON document (list of rows) to stdout / --out."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")]
import torch
from immtsf import ops


def peak_hbm():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "MEASURED_PEAKS.json"
    except Exception:
        return 6650.0, "fallback"


def timeit(fn, flush, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ms = []
    for _ in range(iters):
        flush.fill_(1.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    ms.sort()
    return ms[len(ms) // 2]


def ragged(B, N, d, full, seed=0):
    g = torch.Generator().manual_seed(seed)
    counts = torch.full((B,), N) if full else torch.randint(1, N + 1, (B,), generator=g)
    counts[0] = N
    notes = torch.zeros(B, N, d)
    mask = torch.arange(N)[None, :] < counts[:, None]
    notes[mask] = torch.randn(int(mask.sum()), d, generator=g)
    tau = torch.rand(B, N, generator=g) * 7.0 * mask
    return notes.cuda(), tau.cuda(), int(mask.sum())


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    ap.add_argument("--only-recavg", action="store_true", help="the three Time-IMM-sized RecAvg rows only (A/B runs of kernel variants)")
    ap.add_argument("--only-large", action="store_true", help="the long-segment x long-window RecAvg rows only (tensor-core path A/B)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    peak, src = peak_hbm()
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    rows = []
    d = 768
    shapes = [(256, 16, 24, False), (2048, 16, 24, False), (2048, 16, 24, True), (1024, 64, 28, False),
              (512, 256, 64, False), (64, 1024, 256, False), (4096, 4, 16, False)]
    if args.only_recavg:
        shapes = [(256, 16, 24, False), (2048, 16, 24, False), (2048, 16, 24, True), (4096, 4, 16, False)]
    if args.only_large:
        shapes = [(512, 256, 64, False), (128, 1024, 64, False), (128, 256, 256, False), (64, 1024, 256, False)]
        args.only_recavg = True
    for (B, N, T, full) in shapes:
        notes, tau, sumN = ragged(B, N, d, full)
        r = ops.csr_build(notes, tau)
        t_hat = (0.5 + 0.5 * torch.rand(B, T)).sort(dim=1)[0].cuda()
        ls = torch.tensor(0.0, device=dev)
        gamma, beta = torch.ones(d, device=dev), torch.zeros(d, device=dev)
        Vp = r.emb_flat
        thr, seed = ops.drop_thr(0.1), 1234
        ms_csr = timeit(lambda: ops.csr_build(notes, tau), flush)
        b_csr = 4.0 * (B * N * d + 2 * sumN * d)
        rows.append(dict(kernel="csr_build (mask+scan+gather, 3 launches)", B=B, N_max=N, T=T, sumN=sumN, ms=ms_csr,
                         alg_bytes=b_csr, GBps=b_csr / ms_csr / 1e6, frac=b_csr / ms_csr / 1e6 / peak))
        out = {}
        def fwd():
            out["v"] = ops.recavg_pool_fwd(Vp, r, t_hat, ls, gamma, beta, T, d, thr, seed, True)
        ms_f = timeit(fwd, flush)
        b_f = 4.0 * (sumN * d + sumN + B * T + 2 * B * T * d)  # V' in, E_drop + E_raw out (training)
        rows.append(dict(kernel="recavg_pool_fwd", B=B, N_max=N, T=T, sumN=sumN, ms=ms_f, alg_bytes=b_f,
                         GBps=b_f / ms_f / 1e6, frac=b_f / ms_f / 1e6 / peak,
                         pool_gflops=2.0 * T * sumN * d / ms_f / 1e6))
        E_drop, E_raw, mean, rstd, wsum = out["v"]
        dE = torch.randn_like(E_drop)
        ms_b = timeit(lambda: ops.recavg_pool_bwd(dE, E_raw, mean, rstd, wsum, Vp, r, t_hat, ls, gamma, T, d, thr, seed), flush)
        b_b = 4.0 * (2 * sumN * d + 2 * B * T * d)  # V' in, dV' out, dE_drop + E_raw in
        rows.append(dict(kernel="recavg_pool_bwd", B=B, N_max=N, T=T, sumN=sumN, ms=ms_b, alg_bytes=b_b,
                         GBps=b_b / ms_b / 1e6, frac=b_b / ms_b / 1e6 / peak,
                         pool_gflops=4.0 * T * sumN * d / ms_b / 1e6))
        del notes, tau, r, Vp, E_drop, E_raw, dE, out
        torch.cuda.empty_cache()
    # skinny streaming kernels at the cfg2 / large-batch row counts
    for M in (() if args.only_recavg else (6144, 49152)):
        X = torch.randn(M, d, device=dev)
        W4 = torch.randn(4, d, device=dev)
        Y4 = torch.randn(M, 4, device=dev)
        o4 = torch.empty(M, 4, device=dev)
        od = torch.empty(M, d, device=dev)
        for name, fn, byt in [
            ("colsum (bias gradient)", lambda: ops.colsum(X), 4.0 * M * d),
            ("gemm_smalln (residual_head, N=4)", lambda: ops.gemm(X, W4, o4, transB=True), 4.0 * M * d),
            ("gemm_smallk (proj_q, K=4)", lambda: ops.gemm(Y4, W4, od), 4.0 * M * d),
            ("gemm_tallt (dW of a C-wide projection)", lambda: ops.gemm(Y4, X, torch.empty(4, d, device=dev), transA=True), 4.0 * M * d),
        ]:
            ms = timeit(fn, flush)
            rows.append(dict(kernel=name, rows=M, d=d, ms=ms, alg_bytes=byt, GBps=byt / ms / 1e6, frac=byt / ms / 1e6 / peak))
    # GR_Add scan + tail (latency-bound T-step recurrence)
    for (B, T, C) in ([] if args.only_recavg else [(256, 24, 4), (2048, 24, 4), (256, 192, 96)]):
        G4 = torch.randn(B * T, 4 * C, device=dev)
        w_hh, b_hh = torch.randn(3 * C, C, device=dev) * 0.1, torch.zeros(3 * C, device=dev)
        ms = timeit(lambda: ops.gru_scan_fwd(G4, w_hh, b_hh, B, T, C), flush)  # C > 32: wide recurrence
        byt = 4.0 * B * T * (3 * C + 2 * C)
        rows.append(dict(kernel="gru_scan_fwd", B=B, T=T, C=C, ms=ms, alg_bytes=byt, GBps=byt / ms / 1e6, frac=byt / ms / 1e6 / peak,
                         us_per_scan_step=ms * 1e3 / T))
    doc = dict(peak_GBps=peak, peak_source=src, method="CUDA events per launch, median of 10, 256 MiB L2 flush between launches", rows=rows)
    txt = json.dumps(doc, indent=1)
    if args.out:
        open(args.out, "w").write(txt)
    for r in rows:
        print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()})


if __name__ == "__main__":
    main()
