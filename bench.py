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
for p in (ROOT, os.path.join(ROOT, "imm-tsf_b200"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on
    "cfg2": dict(ttf="TTF_T2V_XAttn", mmf="MMF_XAttn_Add", B=256, N=16, T=24, d_model=768, d_txt=768, C=4, H=1, kappa=0.5,
                 history=7.0, pred=7.0, cpu_sample_B=32,
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


def make_batch(w, seed):
    import gpu_common as G

    return G.synth_batch(w["B"], w["N"], w["T"], w["d_model"], w["C"], seed, history=w["history"], pred=w["pred"])


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


# ----------------------------------------------------------------------------- CPU arm
def cpu_oracle_step_fn(w, sample_B, threads):
    """Returns (fn, samples): one fwd+bwd of the oracle port on `sample_B` samples of the workload."""
    import gpu_common as G
    from oracle import immtsf_oracle as O

    torch.set_num_threads(threads)
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
    for v in P.values():
        v.requires_grad_(True)
    notes, tau, t_hat, Y, _ = G.synth_batch(sample_B, w["N"], w["T"], w["d_model"], w["C"], 1234, history=w["history"], pred=w["pred"])
    B, N, T, C, H, d = sample_B, w["N"], w["T"], w["C"], w["H"], w["d_txt"]
    keep = lambda *s: (torch.rand(*s, generator=g) >= DROPOUT).float()
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
        out.square().mean().backward()
        return float(out.detach()[0, 0, 0])

    return step, sample_B


def time_cpu(w, steps, warmup, threads):
    step, nB = cpu_oracle_step_fn(w, w["cpu_sample_B"], threads)
    for _ in range(warmup):
        step()
    ts = []
    for _ in range(steps):
        t0 = time.perf_counter()
        step()
        ts.append(time.perf_counter() - t0)
    med = statistics.median(ts)
    return nB / med, med, nB


def run_reference_arm(args, w):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 5))
    warm = max(1, min(args.warmup, 2))
    sps, med, nB = time_cpu(w, steps, warm, threads)
    sample = f"{nB} of {w['B']} samples per step, fwd+bwd, dropout {DROPOUT}, {steps} timed steps (median), {warm} warm-up"
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


def run_gpu_arm(args, w):
    import torch.distributed as dist
    import gpu_common as G
    from immtsf import _lib, dp, ops, runtime

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    os.environ.setdefault("IMMTSF_NAN_CHECK", "0")  # the guard's single host sync is measured separately in e2e_checked

    cfg = dict(ttf=w["ttf"], mmf=w["mmf"], d_txt=w["d_txt"], C=w["C"], H=w["H"], kappa=w["kappa"])
    import contextlib
    with contextlib.redirect_stdout(sys.stderr):  # FusionModel prints its module names like the reference does
        fm = G.build_model(cfg, w["d_model"], dropout=DROPOUT, seed=1)  # same init on every rank
    fm.train()
    params = [p for p in fm.parameters()]
    notes, tau, t_hat, Y, _ = make_batch(w, 1234 + rank)
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

    def timed(n, resident, graphed=None):
        """n steps, each bracketed by its own CUDA-event pair (the 256 MiB L2 flush sits between the pairs).  The host
        enqueues all n steps and synchronises once at the end, as a training loop does: a host sync per step would add
        the host's launch latency to every step and, on several GPUs, let the ranks drift apart between steps."""
        pairs = []
        for _ in range(n):
            flush.fill_(1.0)  # L2 flush, outside the timed events
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            src = d_in if resident else h_in
            if graphed is not None:
                loss = graphed(*src)  # copies into the static buffers (H2D when src is pinned host memory) + one replay
                if world > 1 and graphed.group is None:
                    dp.allreduce_grads(params, flat=graphed.flat_grads)  # (in-graph mode: the replay already reduced)
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

    W, K = max(args.warmup, 3), args.steps
    # ---- eager path (one Python call per kernel launch)
    timed(W, True)
    barrier()
    l0 = _lib.launch_count()
    ms_eager = max_over_ranks(timed(K, True))
    launches_per_step = (_lib.launch_count() - l0) // max(K, 1)
    barrier()
    # ---- public API for a fixed-shape training loop: the whole step captured in a CUDA graph (runtime.GraphedStep)
    graphed = None
    if not args.eager:
        for p in params:
            p.grad = None
        dp_mode = "none"
        if world > 1 and args.dp_mode == "ingraph":
            try:
                graphed = runtime.GraphedStep(fm, example=d_in, warmup=2, allreduce_group=True)
                dp_mode = ("NCCL captured in the step graph: rank-form upstream gradients (KBs) reduced mid-backward on the side stream, "
                           "one %s of the remaining gradients at the end; %d of %d gradient floats are never communicated"
                           % ("coalesced all-reduce (no flat bucket)" if graphed.coalesced else "all-reduce of the flat bucket", graphed.n_first, graphed.n_total))
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
    ms_res = timed(K, True, graphed)
    launches = launches_per_step * K
    barrier()
    ms_res = max_over_ranks(ms_res)
    timed(2, False, graphed)
    barrier()
    ms_e2e_serial = max_over_ranks(timed(K, False, graphed))
    barrier()
    ms_e2e, h2d_hidden = ms_e2e_serial, None
    if graphed is not None:
        timed_prefetched(2, graphed)
        barrier()
        ms_pf, h2d_hidden = timed_prefetched(K, graphed)
        ms_e2e = max_over_ranks(ms_pf)
        barrier()
    clk = clocks.stop() if rank == 0 else None

    # instrumented pass: a CUDA-event pair around every gemm_tc_kernel launch, recorded inside the library on the
    # launch stream (immtsf_profile_begin/end).  Ragged launches are issued over M_alloc rows but only sumN are
    # live: count live work only.
    nprof = min(K, 5)
    with _lib.profile_gemm_tc() as recs:
        timed(nprof, True)
        torch.cuda.synchronize()
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
    samples = w["B"] * world * K
    value = samples / (ms_res / 1e3)
    e2e = samples / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in h_in)
    fwd_f, fb_f = algorithmic_flops(w, sumN)
    achieved = gemm_flops_live / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    threads = os.cpu_count() or 1
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sps, med, nB = time_cpu(w, 2, 1, threads)
        cpu = {"value": sps, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{nB} of {w['B']} samples per step, fwd+bwd, dropout {DROPOUT}, 2 timed steps (median), 1 warm-up; "
                         f"{med * 1e3:.0f} ms/step"}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_res / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": w["name"], "step": "FusionModel forward+backward (all param grads + dY_ts), train mode, dropout 0.1",
                   "per_gpu_batch": w["B"], "global_batch": w["B"] * world, "sum_notes_rank0": sumN,
                   "parallelism": f"dp{world}" if world > 1 else "single", "dp_allreduce": dp_mode if not args.eager else "after backward",
                   "l2": "256 MiB flush write before every timed step", "gemm_backend": os.environ.get("IMMTSF_GEMM", "auto"),
                   "launch": "eager (one host call per kernel)" if args.eager else "runtime.GraphedStep (whole step replayed as one CUDA graph)",
                   "eager_samples_per_s": samples / (ms_eager / 1e3), "kernels_per_step": launches_per_step,
                   "algorithmic_gflop_fwd_bwd": fb_f / 1e9},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / K,
                "input_pipeline": "eager: H2D then step" if graphed is None else
                                  "double-buffered (GraphedStep.prefetch): the H2D of step i+1 runs inside step i's timed region",
                "h2d_not_landed_at_region_end": h2d_hidden,
                "serial_value": samples / (ms_e2e_serial / 1e3)},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "tensor", "kernel": "gemm_tc2_kernel / gemm_tc_kernel / gemm_tc_group_kernel (tcgen05 3xTF32 dense projections: CTA pairs, "
                               "single-CTA tiles, grouped folds; %d launches/step)" % (n_gemm // max(nprof, 1)),
                     "achieved": achieved, "peak": peaks["tensor"], "unit": "TFLOP/s", "frac": achieved / peaks["tensor"],
                     "traffic": TRAFFIC_NCU, "traffic_note": TRAFFIC_NOTE, "peak_source": peaks["src"],
                     "note": "achieved = fp32-exact (algorithmic) FLOPs of the live rows / summed per-launch CUDA-event time; every "
                             "product costs 3 TF32 MMAs, so the ceiling of this kernel is peak/6 (TF32 = bf16/2, 3 passes)",
                     "frac_of_3xtf32_ceiling": achieved / (peaks["tensor"] / 6.0),
                     "tensor_pipe_frac_executed": 3.0 * achieved / (peaks["tensor"] / 2.0),
                     "kernel_share_of_step": (gemm_ms / nprof) / (ms_res / K) if ms_res > 0 else None,
                     "step_tflops_algorithmic_reference_schedule": fb_f / (ms_res / K / 1e3) / 1e12},
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
