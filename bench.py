#!/usr/bin/env python
"""Benchmark of the IMM-TSF text->time-series fusion hot path (BASELINE.json metric:
fused TTF+MMF samples/s, forward+backward, with roofline fraction and the host-CPU
reference beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg1|cfg3]
    python bench.py --impl reference ...     # CPU arm: the oracle port on the host cores

One "step" = FusionModel forward + backward (every parameter gradient and dY_ts) over one
synthetic Time-IMM-shaped batch; train mode, dropout 0.1, loss = mean(Y_out^2).  The
optimizer is not part of the fusion path and is not timed.

  value     samples/s with the batch already resident in HBM; per-step CUDA-event times,
            L2 flushed (256 MiB write) before every timed step, max over ranks.
  e2e       same step through the public FusionModel API fed from pinned HOST buffers:
            H2D of notes/tau/t_hat/Y_ts and D2H of the loss inside the timed region.
  roofline  the dominant kernel family of the workload (dense projections: immtsf_gemm),
            algorithmic FLOPs / summed CUDA-event launch durations vs the measured bf16 peak.
  cpu_baseline  the oracle port (oracle/immtsf_oracle.py, follows the reference line by line
            incl. its T_f-fold K/V expansion) on the host cores, bounded sample.
N > 1: launched by torchrun, one rank per GPU; every rank runs its own B-sized shard (weak
scaling) and the step ends with one NCCL all-reduce of the flat gradient bucket.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "imm-tsf_b200")):  # (tests/ is NOT on the path: the GPU arm never imports the oracle)
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(ttf="TTF_T2V_XAttn", mmf="MMF_XAttn_Add", B=256, N=16, T=24, d_model=768, d_txt=768, C=4, H=1, kappa=0.5,
                 history=7.0, pred=7.0, cpu_sample_B=256,
                 name="cfg2: T2V_XAttn+XAttn_Add, B256 N<=16 T24 d768 C4 H1 (BASELINE.json configs[1])"),
    # configs[0]: the reference's own CPU-runnable case
    "cfg1": dict(ttf="TTF_RecAvg", mmf="MMF_GR_Add", B=32, N=16, T=24, d_model=768, d_txt=768, C=4, H=1, kappa=0.5,
                 history=7.0, pred=7.0, cpu_sample_B=32,
                 name="cfg1: RecAvg+GR_Add, B32 N<=16 T24 d768 C4 (BASELINE.json configs[0])"),
    # configs[2]: LLaMA-width embeddings, GDELT-shaped
    "cfg3": dict(ttf="TTF_T2V_XAttn", mmf="MMF_GR_Add", B=256, N=64, T=28, d_model=4096, d_txt=768, C=5, H=1, kappa=0.5,
                 history=14.0, pred=14.0, cpu_sample_B=16,
                 name="cfg3: T2V_XAttn+GR_Add, B256 N<=64 T28 d_model4096->768 C5 (BASELINE.json configs[2])"),
    # configs[4]: MIMIC-shaped (many variables, long irregular histories), per-GPU batch of the 8-GPU run
    "cfg5": dict(ttf="TTF_T2V_XAttn", mmf="MMF_XAttn_Add", B=256, N=32, T=192, d_model=768, d_txt=768, C=96, H=1, kappa=0.5,
                 history=24.0, pred=24.0, cpu_sample_B=4,
                 name="cfg5: T2V_XAttn+XAttn_Add, B256 N<=32 T192 d768 C96 (BASELINE.json configs[4], per-GPU shard)"),
    # the per-(note, query) Time2Vec attention (fusions/TTF_T2V_XAttn_old.py semantics) on the cfg2 shape
    "cfg2q": dict(ttf="TTF_T2V_XAttn_old", mmf="MMF_XAttn_Add", B=256, N=16, T=24, d_model=768, d_txt=768, C=4, H=1, kappa=0.5,
                  history=7.0, pred=7.0, cpu_sample_B=8,
                  name="cfg2q: per-(note,query) T2V_XAttn(_old)+XAttn_Add, B256 N<=16 T24 d768 C4 H1 (SURVEY 8f row f3)"),
    "cfg5g": dict(ttf="TTF_T2V_XAttn", mmf="MMF_GR_Add", B=256, N=32, T=192, d_model=768, d_txt=768, C=96, H=1, kappa=0.5,
                  history=24.0, pred=24.0, cpu_sample_B=4,
                  name="cfg5g: T2V_XAttn+GR_Add, B256 N<=32 T192 d768 C96 (MIMIC-shaped, GRU fusion)"),
}
DROPOUT = 0.1
# dram__bytes_read.sum + dram__bytes_write.sum of one gemm_tc_kernel<.,.,256> launch (M6144 N768 K768) from the
# ncu --set full capture in profiles/r1_ncu_gemm_tc_bn256_summary.txt; algorithmic operand bytes of that launch:
# 42.5 MB (A, A_lo, B, B_lo; the 18.9 MB output stays in the 126 MB L2)
TRAFFIC_NCU = 36.26e6
TRAFFIC_NOTE = ("bytes per launch of the longest GEMM of the cfg2 step, dX = dKVp Wkv (gemm_tc_kernel<0,1,128>, 2134 live of "
                "4096 rows, N768, K1536), cold L2: dram read 36.23 MB + write 0.03 MB, profiles/r1_ncu_step_gemms_summary.txt "
                "(algorithmic operand bytes 35.6e6: A and A_lo 26.2 MB, B and B_lo 9.4 MB; the output stays in L2)")
# recavg_pool_fwd_s_kernel + recavg_bwd_mma_kernel at B 2048, N<=16, T 24, d 768 (ncu --set full, profiles/r1_ncu_recavg_v4_summary.txt and
# profiles/r1_ncu_recavg_fused_bwd_summary.txt): dram read + write per launch, summed over the two kernels
TRAFFIC_RECAVG_NCU = 295.7e6 + 425.0e6
TRAFFIC_RECAVG_NOTE = ("dram__bytes_read.sum + dram__bytes_write.sum of the two pooling kernels at B 2048, N<=16, T 24, d 768 (forward 51.8 + 243.9 MB, "
                       "one-launch backward recavg_bwd_mma_kernel 383.6 + 41.4 MB, profiles/r2_ncu_recavg_bwd_mma_summary.txt; algorithmic 353.7 + 404.9 MB): "
                       "no re-reads; per launch at that batch, not at this line's batch")
METRIC = "fused TTF+MMF fwd+bwd throughput"
UNIT = "samples/s"


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return dict(hbm=float(pk["hbm_gbs"]), tensor=float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])),
                    src="MEASURED_PEAKS.json (bf16 sustained)")
    except Exception:
        return dict(hbm=6650.0, tensor=1400.0, src="fallback (B200_PROFILING.md)")


def make_batch(w, seed, B=None):
    from immtsf import synth

    return synth.synth_batch(B or w["B"], w["N"], w["T"], w["d_model"], w["C"], seed, history=w["history"], pred=w["pred"])


def algorithmic_flops(w, sumN, per_query=True):
    """SURVEY.md 8(d) 'Algorithmic FLOPs per batch' (deduplicated), forward; backward = 2x except input_proj (1x)."""
    B, T, d_m, d, C = w["B"], w["T"], w["d_model"], w["d_txt"], w["C"]
    ST = B * T
    f_in = 2 * sumN * d_m * d
    if w["ttf"] == "TTF_RecAvg":
        ttf = 2 * sumN * T * d + 2 * ST * d * d  # pool (upper bound sum_i T_i N_i d) + proj
    elif w["ttf"] == "TTF_T2V_XAttn_old":
        # once per note: W_a; per (note, query): score (d_tau) + pooling of A and phi; per (sample, query) row: W_phi, W_v, W_o, proj_out
        ttf = 2 * sumN * d * d + 2 * sumN * T * (d // 2) + 2 * sumN * T * (d + d // 2) + ST * (2 * (d // 2) * d + 3 * 2 * d * d)
    else:
        R = ST if per_query else B
        ttf = 2 * sumN * (d + d // 2) * d + 2 * sumN * d * 2 * d + 2 * d * d + 2 * sumN * d + 2 * sumN * T * d + 2 * R * d * d * 2
    if w["mmf"] == "MMF_GR_Add":
        mmf = 2 * ST * (C + d) * 4 * C + 2 * ST * 3 * C * C + 2 * ST * C * C
    else:
        mmf = 2 * ST * C * d + 2 * 2 * ST * d * d + 3 * 2 * ST * d * d + 4 * B * T * T * d + 2 * ST * d * d + 2 * ST * d * C
    return f_in + ttf + mmf, f_in + 3 * (ttf + mmf) + f_in


# ----------------------------------------------------------------------------- reference-schedule arms (oracle port)
def oracle_step_fn(w, sample_B, device="cpu"):
    """One fwd+bwd of the oracle port (oracle/immtsf_oracle.py: the reference's modules restated line by line, run with the
    reference's own T_f-fold K/V expansion) on `sample_B` samples of the workload, on `device`.  This is the ONLY place
    bench.py touches oracle/: the cpu_baseline / --impl reference legs (host cores) and gpu_eager_baseline (the same
    schedule on torch eager CUDA = cuBLAS fp32 + ATen on the same B200).  Never inside the timed region of the GPU arm."""
    from immtsf import synth
    from oracle import immtsf_oracle as O

    cfg = dict(ttf=w["ttf"], mmf=w["mmf"], d_txt=w["d_txt"], C=w["C"], H=w["H"], kappa=w["kappa"])
    shapes = O.param_shapes(w["ttf"], w["mmf"], w["d_model"], w["d_txt"], w["C"])
    g = torch.Generator().manual_seed(7)
    P = {}
    for k, s in shapes.items():
        if k.endswith("log_recency_sigma"):
            P[k] = torch.tensor(0.0)
        elif k.endswith("layer_norm.weight"):
            P[k] = torch.ones(s)
        else:
            fan = s[-1] if len(s) >= 2 else 1
            P[k] = (torch.rand(s, generator=g) * 2 - 1) / max(fan, 1) ** 0.5
    P = {k: v.to(device).requires_grad_(True) for k, v in P.items()}
    notes, tau, t_hat, Y, _ = synth.synth_batch(sample_B, w["N"], w["T"], w["d_model"], w["C"], 1234, history=w["history"], pred=w["pred"])
    notes, tau, t_hat, Y = (t.to(device) for t in (notes, tau, t_hat, Y))
    B, N, T, C, H, d = sample_B, w["N"], w["T"], w["C"], w["H"], (w["d_txt"] or w["d_model"])
    keep = lambda *s: (torch.rand(*s, generator=g) >= DROPOUT).float().to(device)
    masks = {"ttf.dropout": keep(B, T, d), "mmf.dropout": keep(B, T, C)}
    if w["ttf"].startswith("TTF_T2V_XAttn"):
        masks["ttf.attn_dropout"] = keep(B, T, H, N)
    if w["mmf"] == "MMF_XAttn_Add":
        masks["mmf.attn_dropout"] = keep(B, H, T, T)

    def step():
        for v in P.values():
            v.grad = None
        Yr = Y.clone().requires_grad_(True)
        out = O.fusion_forward(P, cfg["ttf"], cfg["mmf"], notes, tau, t_hat, Yr, n_heads=H, kappa=cfg["kappa"], p=DROPOUT,
                               masks=masks, faithful_expand=True)
        loss = out.square().mean()
        loss.backward()
        return loss

    return step


def time_cpu(w, steps, warmup, threads, budget_s=150.0):
    """Median seconds per step of the CPU arm.  The full batch is used whenever steps+warmup fit the time budget (cfg1 / cfg2
    do); otherwise a power-of-two sample of it, stated in the result."""
    torch.set_num_threads(threads)
    nB = min(w["B"], w.get("cpu_sample_B", w["B"]))
    step = oracle_step_fn(w, nB)
    t0 = time.perf_counter()
    step()
    t1 = time.perf_counter() - t0
    if t1 * (steps + warmup) > budget_s and nB > 1:
        while nB > 1 and t1 * (steps + warmup) > budget_s:
            nB //= 2
            t1 /= 2
        step = oracle_step_fn(w, nB)
        step()
    for _ in range(max(warmup - 1, 0)):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return nB / med, med, nB


def time_gpu_eager_reference_schedule(w, dev, steps=5, warmup=2):
    """The reference schedule (oracle port, T_f-fold K/V expansion, one ATen/cuBLAS launch per op, the reference's three
    isnan().any() host syncs) on the same B200: the honest same-box GPU baseline."""
    torch.backends.cuda.matmul.allow_tf32 = False  # the reference's default: true-fp32 SGEMM
    try:
        step = oracle_step_fn(w, w["B"], device=dev)
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        ts = []
        for _ in range(steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            step()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        med = statistics.median(ts)
        peak_mb = torch.cuda.max_memory_allocated(dev) / 2**20
        return {"value": w["B"] / (med / 1e3), "unit": UNIT, "ms_per_step": med, "steps": steps, "warmup": warmup,
                "kind": "port: oracle restatement of the reference modules run with the reference's schedule (T_f-fold K/V expansion, "
                        "per-op launches, 3 isnan host syncs) on torch eager CUDA (cuBLAS fp32, allow_tf32=False, ATen) on the same B200",
                "batch": w["B"], "peak_mem_mib": round(peak_mb)}
    except Exception as e:  # e.g. out of memory on the expansion at sweep sizes
        return {"unavailable": f"{type(e).__name__}: {str(e)[:160]}"}
    finally:
        torch.cuda.empty_cache()


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps, warm = max(1, args.steps), max(1, args.warmup)
    sps, med, nB = time_cpu(w, steps, warm, threads)
    sample = (f"{nB} of {w['B']} samples per step" + (" (the whole batch)" if nB == w["B"] else " (bounded sample, ms_per_step scaled to the batch)")
              + f", fwd+bwd, dropout {DROPOUT}, {steps} timed steps (median), {warm} warm-up, {threads} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": sps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps, "warmup": warm,
        "ms_per_step": med * 1e3 * (w["B"] / nB), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": {"workload": w["name"], "parallelism": "host CPU threads"},
        "cpu_baseline": {"value": sps, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": sps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        self.dev, self.proc, self.path = dev, None, f"/tmp/immtsf_clocks_{os.getpid()}.csv"
        self.nv = None

    def _start_nvml(self):
        """NVML in a sampling thread (same counters nvidia-smi prints, but a 5 ms period: the timed region of a
        sub-millisecond step is shorter than nvidia-smi's start-up)."""
        import threading

        import pynvml as N

        N.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[self.dev]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else self.dev
        h = N.nvmlDeviceGetHandleByIndex(idx)
        bits = {"hw_slowdown": getattr(N, "nvmlClocksEventReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(N, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(N, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        reasons_fn = getattr(N, "nvmlDeviceGetCurrentClocksEventReasons", None) or N.nvmlDeviceGetCurrentClocksThrottleReasons
        st = {"sm": [], "mx": float(N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM)), "reasons": set(), "stop": False, "from": 0}

        def loop():
            while not st["stop"]:
                try:
                    st["sm"].append(float(N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM)))
                    r = reasons_fn(h)
                    for nm, b in bits.items():
                        if r & b:
                            st["reasons"].add(nm)
                except Exception:
                    pass
                time.sleep(0.005)

        st["thread"] = threading.Thread(target=loop, daemon=True)
        st["thread"].start()
        self.nv = st

    def mark(self):
        """Forget what was sampled so far (warm-up)."""
        if self.nv is not None:
            self.nv["from"] = len(self.nv["sm"])
            self.nv["reasons"].clear()

    def start(self):
        try:
            self._start_nvml()
            return
        except Exception:
            self.nv = None
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "25",
                                          "-i", str(self.dev)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.nv is not None:
            self.nv["stop"] = True
            self.nv["thread"].join(timeout=2)
            sm = self.nv["sm"][self.nv["from"]:]
            return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.nv["mx"],
                    "reasons": sorted(self.nv["reasons"]), "samples": len(sm), "source": "NVML, 5 ms period"}
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in open(self.path):
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:  # the timed region was shorter than one sampling period: one synchronous reading right after it
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.dev)],
                                     capture_output=True, text=True, timeout=20).stdout
                c = [x.strip() for x in out.strip().splitlines()[0].split(",")]
                sm.append(float(c[1])); mx.append(float(c[2]))
                for nm, v in zip(names, c[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


HBM_ENTRIES = ("immtsf_recavg_pool_fwd", "immtsf_recavg_pool_bwd", "immtsf_gru_scan_fwd", "immtsf_gru_scan_bwd",
               "immtsf_gr_tail_fwd", "immtsf_gr_tail_bwd")


def hbm_algorithmic_bytes(entry, B, T, d, C, sumN):
    """Algorithmic bytes of one launch of the HBM-bound kernels (SURVEY.md 8d per-unit figures x units)."""
    BT = B * T
    return {
        # V' in, tau, t_hat; E_drop + E_raw out (training: E_raw is the saved LayerNorm input), mean / rstd / wsum
        "immtsf_recavg_pool_fwd": 4.0 * (sumN * d + sumN + BT + 2 * BT * d + 3 * BT),
        # dE_drop + E_raw in, V' in, dV' out
        "immtsf_recavg_pool_bwd": 4.0 * (2 * BT * d + 2 * sumN * d + 3 * BT),
        "immtsf_gru_scan_fwd": 4.0 * BT * (3 * C + 2 * C),  # gate pre-activations in, h_all + h_prev out
        "immtsf_gru_scan_bwd": 4.0 * BT * (4 * C + C + C + 3 * C + 3 * C),  # G4, h_prev, dh_out in; dG4 (GRU part) + dGh out
        "immtsf_gr_tail_fwd": 4.0 * BT * (C + 4 * C + C + C),  # Y, G4, h_all in; Y_out out
        "immtsf_gr_tail_bwd": 4.0 * BT * (C + 4 * C + C + C + C + C),  # dY_out, G4, h_all in; dG4 (gate part), d_delta, dh_out out
    }[entry]


def hbm_records(w, B, dev, flush, peak_hbm, nprof=5):
    """In-step CUDA-event brackets (on the launching stream) around every RecAvg / GR_Add streaming kernel of an EAGER
    RecAvg+GR_Add training step at batch B: per-kernel mean launch duration -> algorithmic GB/s and fraction of the measured
    copy bandwidth."""
    from immtsf import _lib, synth

    cfg = dict(ttf="TTF_RecAvg", mmf="MMF_GR_Add", d_txt=w["d_txt"], C=w["C"], H=1, kappa=0.5)
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):
        fm = synth.build_model(cfg, w["d_model"], dropout=DROPOUT, seed=1)
    fm.train()
    notes, tau, t_hat, Y, _ = synth.synth_batch(B, w["N"], w["T"], w["d_model"], w["C"], 1234, history=w["history"], pred=w["pred"])
    sumN = int((notes.abs().sum(2) > 0).sum())
    d_in = [t.to(dev) for t in (notes, tau, t_hat, Y)]
    params = list(fm.parameters())

    def step():
        for p in params:
            p.grad = None
        out = fm(d_in[0], d_in[1], d_in[2], d_in[3].detach().requires_grad_(True))
        out.square().mean().backward()

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    _lib.PROFILE = {k: [] for k in HBM_ENTRIES}
    try:
        for _ in range(nprof):
            flush.fill_(1.0)
            step()
        torch.cuda.synchronize()
        recs = {k: [e0.elapsed_time(e1) for e0, e1 in v] for k, v in _lib.PROFILE.items()}
    finally:
        _lib.PROFILE = None
    d = w["d_txt"] or w["d_model"]
    out = []
    for k in HBM_ENTRIES:
        if not recs[k]:
            continue
        ms = statistics.mean(recs[k])
        byt = hbm_algorithmic_bytes(k, B, w["T"], d, w["C"], sumN)
        out.append({"kernel": k, "B": B, "N_max": w["N"], "T": w["T"], "d": d, "C": w["C"], "sumN": sumN, "us": ms * 1e3,
                    "algorithmic_bytes": byt, "achieved": byt / ms / 1e6, "peak": peak_hbm, "unit": "GB/s", "frac": byt / ms / 1e6 / peak_hbm,
                    "launches_timed": len(recs[k])})
    del fm
    torch.cuda.empty_cache()
    return out


def run_gpu_arm(args, w):
    import contextlib

    import torch.distributed as dist
    from immtsf import _lib, dp, ops, runtime, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # The reference's NaN guards (FusionModel.py:103-112) stay ON: their flag kernels run inside every timed region.  In graph
    # replay the single device->host read of the flags is GraphedStep.check_nan(); its cost is reported as nan_guard / e2e.checked.
    strong = args.scaling == "strong"
    Bl = w["B"] // world if strong else w["B"]  # per-GPU batch
    assert Bl >= 1

    cfg = dict(ttf=w["ttf"], mmf=w["mmf"], d_txt=w["d_txt"], C=w["C"], H=w["H"], kappa=w["kappa"])
    with contextlib.redirect_stdout(sys.stderr):  # FusionModel prints its module names like the reference does
        fm = synth.build_model(cfg, w["d_model"], dropout=DROPOUT, seed=1)  # same init on every rank
    fm.train()
    params = [p for p in fm.parameters()]
    notes, tau, t_hat, Y, _ = make_batch(w, 1234 + rank, B=Bl)
    sumN = int((notes.abs().sum(2) > 0).sum())
    h_in = [t.pin_memory() for t in (notes, tau, t_hat, Y)]
    d_in = [t.to(dev) for t in h_in]
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step(inp):
        for p in params:
            p.grad = None
        Yr = inp[3].detach().requires_grad_(True)
        out = fm(inp[0], inp[1], inp[2], Yr)
        loss = out.square().mean()
        loss.backward()
        if world > 1:
            dp.allreduce_grads(params)
        return loss

    def timed(n, resident, graphed=None, check=False):
        """n steps, each bracketed by its own CUDA-event pair (the 256 MiB L2 flush sits between the pairs).  The host
        enqueues all n steps and synchronises once at the end, as a training loop does: a host sync per step would add
        the host's launch latency to every step and, on several GPUs, let the ranks drift apart between steps.
        check: read the NaN flags after every step (GraphedStep.check_nan: one device->host sync per step)."""
        pairs = []
        for _ in range(n):
            flush.fill_(1.0)  # L2 flush, outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            src = d_in if resident else h_in
            if graphed is not None:
                if resident:
                    src = graphed.static_in  # resident = the batch already sits in the step's input buffers: no copy at all
                loss = graphed(*src)  # (host source: H2D into the static buffers) + one replay
                if world > 1 and graphed.group is None:
                    dp.allreduce_grads(params, flat=graphed.flat_grads)  # (in-graph mode: the replay already reduced)
                if check:
                    graphed.check_nan()
            else:
                inp = src if resident else [t.to(dev, non_blocking=True) for t in src]
                loss = step(inp)
            if not resident:
                loss_host.copy_(loss.detach(), non_blocking=True)
            e1.record()
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        return sum(e0.elapsed_time(e1) for e0, e1 in pairs)

    def timed_prefetched(n, graphed):
        """e2e with the double-buffered input pipeline (GraphedStep.prefetch): inside step i's timed region the H2D of
        step i+1's inputs runs on the copy stream next to step i's kernels, and step i waits for its own inputs, copied
        during step i-1.  Every timed region thus contains one full H2D and the D2H of its loss.  `hidden` counts
        the regions whose H2D had NOT landed when the region closed (it would then have run in the untimed flush gap)."""
        graphed.prefetch(*h_in)
        torch.cuda.synchronize()
        pairs, staged = [], []
        for _ in range(n):
            flush.fill_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            loss = graphed.step_prefetched()
            graphed.prefetch(*h_in)  # next step's inputs, overlapping this step
            staged.append(graphed._staged)
            if world > 1 and graphed.group is None:
                dp.allreduce_grads(params, flat=graphed.flat_grads)
            loss_host.copy_(loss.detach(), non_blocking=True)
            e1.record()
            pairs.append((e0, e1))
        torch.cuda.synchronize()
        hidden = sum(1 for (e0, e1), st in zip(pairs, staged) if st.elapsed_time(e1) < 0.0)
        return sum(e0.elapsed_time(e1) for e0, e1 in pairs), hidden

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def rounds(fn, R):
        """R rounds of `fn` (K timed steps each, barrier + synchronize on both sides, max over ranks): the median round is
        reported -- K steps of a sub-millisecond step are ~15 ms of timed region, too short for one round to be stable."""
        out = []
        for _ in range(R):
            barrier()
            out.append(max_over_ranks(fn()))
            barrier()
        return out

    W, K, R = max(args.warmup, 3), args.steps, max(args.rounds, 1)
    # ---- eager path: what the UNMODIFIED caller gets (lib/evaluation.py:95-100 + loss.backward(), main.py:1097): one Python
    # call per kernel launch, the NaN guard's host sync in every forward
    timed(W, True)
    barrier()
    l0 = _lib.launch_count()
    ms_eager = max_over_ranks(timed(K, True))
    launches_per_step = (_lib.launch_count() - l0) // max(K, 1)
    barrier()
    # ---- public API for a fixed-shape training loop: the whole step captured in a CUDA graph (runtime.GraphedStep)
    graphed = None
    dp_mode = "none"
    if not args.eager:
        for p in params:
            p.grad = None
        if world > 1 and args.dp_mode == "ingraph":
            try:
                graphed = runtime.GraphedStep(fm, example=d_in, warmup=2, allreduce_group=True)
                dp_mode = ("NCCL captured in the step graph: the autograd Functions all-reduce the sufficient statistics of their parameter "
                           "gradients (packed weight-gradient outputs before the weight-space un-folds) on the backward lanes; "
                           "%d collectives and %.2f MB per step per GPU for %d gradient floats (%.2f MB); gradients of %d floats are born reduced"
                           % (graphed.dp_calls_per_step, graphed.dp_floats_per_step * 4 / 1e6, graphed.n_total, graphed.n_total * 4 / 1e6,
                              graphed.n_first))
                dp_mode += ("; %d of the collectives are the hand-written NVSwitch multimem all-reduce (csrc/nvls.cu), the rest NCCL"
                            % graphed.dp_nvls_calls_per_step)
            except Exception as e:  # capture of NCCL refused: fall back to one all-reduce after the replay
                print(f"[bench] in-graph all-reduce unavailable ({type(e).__name__}: {e}); reducing after the replay", file=sys.stderr)
                torch.cuda.synchronize()
                graphed = None
        if graphed is None:
            for p in params:
                p.grad = None
            graphed = runtime.GraphedStep(fm, example=d_in, warmup=2, flat_grads=world > 1 or args.flat_grads)
            if world > 1:
                dp_mode = "one NCCL all-reduce of the flat gradient bucket after the graph replay"
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()  # before the warm-up: NVML's first queries are slow and would perturb the first timed steps
    timed(W, True, graphed)
    barrier()
    if rank == 0:
        clocks.mark()  # only samples from here on count
    res_rounds = rounds(lambda: timed(K, True, graphed), R)
    ms_res = statistics.median(res_rounds)
    launches = launches_per_step * K
    timed(2, False, graphed)
    serial_rounds = rounds(lambda: timed(K, False, graphed), min(R, 3))
    ms_e2e_serial = statistics.median(serial_rounds)
    ms_e2e, h2d_hidden, e2e_rounds, ms_checked = ms_e2e_serial, None, serial_rounds, None
    if graphed is not None:
        timed_prefetched(2, graphed)
        hid = []

        def pf():
            ms, h = timed_prefetched(K, graphed)
            hid.append(h)
            return ms

        e2e_rounds = rounds(pf, R)
        ms_e2e, h2d_hidden = statistics.median(e2e_rounds), max(hid)
        # the same step with the reference's ValueError guard honoured after EVERY step: one device->host read of the flags
        ms_checked = statistics.median(rounds(lambda: timed(K, False, graphed, check=True), min(R, 3)))
    clk = clocks.stop() if rank == 0 else None

    # ---- the UNMODIFIED caller's loop (lib/evaluation.py:95-100 + loss.backward(), main.py:1097) through the public API with
    # the transparent graph cache (immtsf/autograph.py): padded shapes change from batch to batch (N_max of the batch)
    autograph_sps = autograph_graphs = None
    if world == 1 and not args.eager:
        variants = []
        for i, nmax in enumerate((w["N"], max(w["N"] - 3, 1), max(w["N"] // 2, 1), w["N"])):
            nb = synth.synth_batch(Bl, nmax, w["T"], w["d_model"], w["C"], 4000 + i, history=w["history"], pred=w["pred"])
            variants.append([t.to(dev) for t in nb[:4]])
        fm.enable_graphs(True)

        def api_steps(n):
            pairs = []
            for i in range(n):
                flush.fill_(1.0)
                v = variants[i % len(variants)]
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for p in params:
                    p.grad = None
                out = fm(v[0], v[1], v[2], v[3].detach().requires_grad_(True))
                out.square().mean().backward()
                e1.record()
                pairs.append((e0, e1))
            torch.cuda.synchronize()
            return sum(e0.elapsed_time(e1) for e0, e1 in pairs)

        api_steps(2 * len(variants))
        ms_api = statistics.median([api_steps(K) for _ in range(min(R, 3))])
        autograph_sps = Bl * K / (ms_api / 1e3)
        autograph_graphs = fm._autograph.captures
        fm.enable_graphs(False)
        fm.clear_graphs()

    # instrumented pass (eager): CUDA-event pairs, on the launching stream, around every launch of the workload's dominant
    # kernel family -- the tcgen05 GEMMs (recorded inside the library: immtsf_profile_begin/end; ragged launches are issued
    # over M_alloc rows but only sumN are live: live work only is counted) and the HBM-bound RecAvg / GR_Add kernels.
    nprof = min(K, 5)
    # Timed ALONE: the pass runs the step on one stream (IMMTSF_SIDE_STREAM=0), so a launch's events bracket that kernel only.
    # In the measured step the lanes run several of these kernels at once and each then takes longer (they share the SMs):
    # `achieved_concurrent` below is the same figure with the lanes on.
    def instrumented():
        _lib.PROFILE = {k: [] for k in HBM_ENTRIES}
        try:
            with _lib.profile_gemm_tc() as rr:
                timed(nprof, True)
                torch.cuda.synchronize()
            return rr, {k: [e0.elapsed_time(e1) for e0, e1 in v] for k, v in _lib.PROFILE.items()}
        finally:
            _lib.PROFILE = None

    recs_conc, _ = instrumented()
    side_env = os.environ.get("IMMTSF_SIDE_STREAM")
    os.environ["IMMTSF_SIDE_STREAM"] = "0"
    try:
        timed(2, True)
        recs, hbm_ms = instrumented()
    finally:
        if side_env is None:
            del os.environ["IMMTSF_SIDE_STREAM"]
        else:
            os.environ["IMMTSF_SIDE_STREAM"] = side_env
    gemm_ms_conc = sum(r[4] for r in recs_conc)
    gemm_ms = sum(r[4] for r in recs)
    gemm_flops_live = 0.0
    for (m, n, k, rd, _ms) in recs:
        if rd == 1:
            m = min(m, sumN)
        elif rd == 2:
            k = min(k, sumN)
        gemm_flops_live += 2.0 * m * n * k
    n_gemm = len(recs)

    if rank != 0:
        if world > 1:
            graphed = None
            torch.cuda.synchronize()
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    samples = Bl * world * K
    value = samples / (ms_res / 1e3)
    e2e = samples / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in h_in)
    fwd_f, fb_f = algorithmic_flops(dict(w, B=Bl), sumN)
    d = w["d_txt"] or w["d_model"]
    step_ms = ms_res / K
    if w["ttf"] == "TTF_RecAvg" and w["mmf"] == "MMF_GR_Add":
        # RecAvg + GR_Add: the HBM roofline (BASELINE.json target, first clause).  Dominant kernels = the two pooling kernels.
        ks = ("immtsf_recavg_pool_fwd", "immtsf_recavg_pool_bwd")
        t_ms = sum(sum(hbm_ms[k]) for k in ks)
        byt = sum(hbm_algorithmic_bytes(k, Bl, w["T"], d, w["C"], sumN) * len(hbm_ms[k]) for k in ks)
        achieved = byt / t_ms / 1e6 if t_ms > 0 else 0.0
        roofline = {"bound": "hbm", "kernel": "recavg_pool_fwd_s_kernel + recavg_bwd_mma_kernel (immtsf_recavg_pool_fwd / _bwd: TMA-staged "
                                              "segments, in-register recency weights, fused LayerNorm + dropout; 2 launches/step)",
                    "achieved": achieved, "peak": peaks["hbm"], "unit": "GB/s", "frac": achieved / peaks["hbm"],
                    "traffic": TRAFFIC_RECAVG_NCU, "traffic_note": TRAFFIC_RECAVG_NOTE, "peak_source": peaks["src"],
                    "note": "achieved = algorithmic bytes (V' in; E_drop, E_raw out / dE_drop, E_raw, V' in; dV' out) of the %d-sample batch / "
                            "summed per-launch CUDA-event time inside the eager step; at this batch the launch is latency-bound (one CTA per "
                            "sample on 148 SMs): roofline_hbm carries the same kernels at B 2048" % Bl,
                    "kernel_share_of_step": (t_ms / nprof) / step_ms if ms_res > 0 else None}
    else:
        achieved = gemm_flops_live / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tc_kernel / gemm_tc_group_kernel (tcgen05 3xTF32 dense projections: CTA pairs, "
                                                 "single-CTA tiles, grouped folds; %d launches/step)" % (n_gemm // max(nprof, 1)),
                    "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor"],
                    "traffic": TRAFFIC_NCU, "traffic_note": TRAFFIC_NOTE, "peak_source": peaks["src"],
                    "note": "achieved = fp32-exact (algorithmic) FLOPs of the live rows / summed per-launch CUDA-event time; every "
                            "product costs 3 TF32 MMAs, so the ceiling of this kernel is peak/6 (TF32 = bf16/2, 3 passes)",
                    "frac_of_3xtf32_ceiling": achieved / (peaks["tensor"] / 6.0),
                    "tensor_pipe_frac_executed": 3.0 * achieved / (peaks["tensor"] / 2.0),
                    "kernel_share_of_step": (gemm_ms / nprof) / step_ms if ms_res > 0 else None,
                    "gemm_gflop_live_per_step": gemm_flops_live / nprof / 1e9,
                    "achieved_concurrent": gemm_flops_live / (gemm_ms_conc / 1e3) / 1e12 if gemm_ms_conc > 0 else None,
                    "timing": "CUDA events around every launch, the step's kernels serialised on one stream (each launch alone); "
                              "achieved_concurrent = the same with the side lanes on, as in the timed step",
                    "launches": sorted(([m, n, k, rd, round(ms * 1e3, 1), round(2.0 * (min(m, sumN) if rd == 1 else m) * n * (min(k, sumN) if rd == 2 else k) / ms / 1e9, 1)]
                                        for (m, n, k, rd, ms) in recs[:len(recs) // max(nprof, 1)]), key=lambda x: -x[4]),
                    "step_tensor_floor_ms": gemm_flops_live / nprof / (peaks["tensor"] / 6.0 * 1e12) * 1e3}
    threads = os.cpu_count() or 1
    cpu = gpu_eager = hbm = None
    if world == 1:
        graphed = None  # free the graph's pool before the secondary measurements
        torch.cuda.empty_cache()
        if not args.no_hbm:
            # secondary record: the HBM-bound kernels of RecAvg + GR_Add bracketed in-step, at the reference batch and at a
            # batch that fills the machine (BASELINE.json target: >= 60 % of the HBM roofline for RecAvg / GR_Add)
            w1 = WORKLOADS["cfg1"]
            hbm = {"peak_source": peaks["src"], "method": "CUDA-event pairs on the launching stream around each entry point inside an eager "
                   "RecAvg+GR_Add training step (d 768, C 4, N<=16, T 24, dropout 0.1), 256 MiB L2 flush between steps, mean of 5 steps",
                   "records": hbm_records(w1, 32, dev, flush, peaks["hbm"]) + hbm_records(w1, 2048, dev, flush, peaks["hbm"])}
        if not args.no_gpu_baseline:
            gpu_eager = time_gpu_eager_reference_schedule(w, dev)
        if not args.no_cpu_baseline:
            sps, med, nB = time_cpu(w, 3, 1, threads, budget_s=30.0)
            cpu = {"value": sps, "unit": UNIT, "cores": threads, "kind": "port",
                   "sample": f"{nB} of {w['B']} samples per step" + (" (the whole batch)" if nB == w["B"] else "")
                             + f", fwd+bwd, dropout {DROPOUT}, 3 timed steps (median), 1 warm-up; {med * 1e3:.0f} ms/step"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "step": "FusionModel forward+backward (all param grads + dY_ts), train mode, dropout 0.1, NaN guard flag kernels on",
                   "per_gpu_batch": Bl, "global_batch": Bl * world, "sum_notes_rank0": sumN,
                   "parallelism": f"dp{world}" if world > 1 else "single", "dp_allreduce": dp_mode if not args.eager else "after backward",
                   "l2": "256 MiB flush write before every timed step", "gemm_backend": os.environ.get("IMMTSF_GEMM", "auto"),
                   "launch": "eager (one host call per kernel)" if args.eager else "runtime.GraphedStep (whole step replayed as one CUDA graph)",
                   "rounds": R, "round_ms_per_step": [x / K for x in res_rounds], "value_is": "median round",
                   "eager_samples_per_s": samples / (ms_eager / 1e3), "kernels_per_step": launches_per_step,
                   "public_api_varying_shapes_samples_per_s": autograph_sps,
                   "public_api_note": "fusion(notes, tau, t_hat, Y_ts); loss.backward() through FusionModel with enable_graphs(): one captured "
                                      "forward/backward graph pair per padded shape (%s pairs for 4 batch shapes, N_max bucketed to 8), NaN "
                                      "ValueError check (one host sync) after every forward" % autograph_graphs,
                   "algorithmic_gflop_fwd_bwd_reference_schedule": fb_f / 1e9},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K, "round_ms_per_step": [x / K for x in e2e_rounds],
                "input_pipeline": "eager: H2D then step" if args.eager else
                                  "double-buffered (GraphedStep.prefetch): the H2D of step i+1 runs inside step i's timed region",
                "h2d_not_landed_at_region_end": h2d_hidden,
                "serial_value": samples / (ms_e2e_serial / 1e3),
                "checked_value": samples / (ms_checked / 1e3) if ms_checked else None,
                "checked_note": "serial H2D + replay + GraphedStep.check_nan() (the reference's ValueError guard: one device->host read of the "
                                "flags) after every step"},
        "nan_guard": {"flag_kernels_in_timed_region": runtime.nan_flags_enabled(),
                      "check_nan_ms_per_step": (ms_checked - ms_e2e_serial) / K if ms_checked else None},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline,
        "roofline_hbm": hbm,
        "gpu_eager_baseline": gpu_eager,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        graphed = None  # graphs that captured NCCL work must go before the communicator
        torch.cuda.synchronize()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--rounds", type=int, default=5, help="rounds of --steps timed steps; the median round is reported")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = the workload's batch per GPU; strong = the workload's batch split across the GPUs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true", help="skip gpu_eager_baseline (the reference schedule on torch eager CUDA)")
    ap.add_argument("--no-hbm", action="store_true", help="skip the roofline_hbm secondary records")
    ap.add_argument("--eager", action="store_true", help="time the eager path only (no CUDA-graph replay)")
    ap.add_argument("--flat-grads", action="store_true", help="N = 1: keep the gradients in the flat bucket of the N > 1 runs (diagnostic)")
    ap.add_argument("--dp-mode", default="ingraph", choices=["ingraph", "post"],
                    help="N > 1: one gradient all-reduce after the graph replay (default), or NCCL captured inside the step "
                         "graph with the rank-form gradients reduced through their small upstream tensors (DESIGN.md 6)")
    args = ap.parse_args()
    w = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference_arm(args, w)
    else:
        run_gpu_arm(args, w)


if __name__ == "__main__":
    main()
