"""autograd.Function wrappers: one per reference module, forward and backward
composed from the immtsf CUDA kernels (imm-tsf_b200/csrc).  No torch math op
sits on the compute path; torch provides memory, views and the autograd tape.

Reference semantics (file:line):
  RecAvgFn     fusions/TTF_RecAvg.py:54-112
  T2VXAttnFn   fusions/TTF_T2V_XAttn.py:93-184 (+ nn.MultiheadAttention internals)
  T2VPerQueryFn fusions/TTF_T2V_XAttn_old.py:82-161 (per-(note, query) Time2Vec)
  GRAddFn      fusions/MMF_GR_Add.py:31-61
  XAttnAddFn   fusions/MMF_XAttn_Add.py:56-103
"""
from __future__ import annotations

import math
import os

import torch

from . import ops
from .ops import RaggedNotes

_f32 = torch.float32


def _need_save(*params) -> bool:
    return torch.is_grad_enabled() and any(p is not None and p.requires_grad for p in params)


# ============================================================== TTF_RecAvg
class RecAvgFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, r: RaggedNotes, t_hat, T, thr, seed, save, defer, log_sigma, W_in, b_in, gamma, beta, W_p, b_p):
        """defer: return dropout(LN(E_raw)) WITHOUT the final `proj` (TTF_RecAvg.py:109) -- the consumer (the rank form of
        MMF_XAttn_Add) folds W_p into its own skinny operand, and E_txt is never materialised."""
        B = r.B
        d = W_p.shape[0]
        step = ops.step_ctx()
        lo = step.lo
        ops.weight_los(lo, [(w, []) for w in ((W_in,) if defer else (W_in, W_p)) if w is not None])
        Vp = ops.linear_fwd(r.emb_flat, W_in, b_in, ragged=r.m_dev, lo=lo) if W_in is not None else r.emb_flat
        E_drop, E_raw, mean, rstd, wsum = ops.recavg_pool_fwd(Vp, r, t_hat, log_sigma, gamma, beta, T, d, thr, seed, save)
        if defer:
            E_txt = E_drop.view(B, T, d)
        else:
            E_txt = ops.linear_fwd(E_drop.view(B * T, d), W_p, b_p, out=step.final_out(B * T, d), lo=lo, emit_lo=step.e_txt_feeds_tc).view(B, T, d)
        if save:
            ctx.r, ctx.T, ctx.thr, ctx.seed, ctx.has_in, ctx.lo, ctx.defer = r, T, thr, seed, W_in is not None, lo, defer
            ctx.save_for_backward(t_hat, log_sigma, W_in, gamma, W_p, Vp, E_drop, E_raw, mean, rstd, wsum)
        return E_txt

    @staticmethod
    def backward(ctx, dE_txt):
        t_hat, log_sigma, W_in, gamma, W_p, Vp, E_drop, E_raw, mean, rstd, wsum = ctx.saved_tensors
        r, T, lo = ctx.r, ctx.T, ctx.lo
        ctx.lo = None
        B, d = r.B, W_p.shape[0]
        dE = dE_txt.contiguous().view(B * T, d)
        fk = ops.Fork(dE.device, name="recavg_bwd")  # weight / bias gradients beside the data-gradient chain
        res = {"dW_p": None, "db_p": None, "dW_in": None, "db_in": None}
        if ctx.defer:
            dE_drop = dE  # (the consumer returns dW_p, db_p)
        else:
            if lo is not None and ops.gemm_backend() != ops.BACKEND_FFMA and B * T >= 64:
                lo.lo_for(dE, None)  # read on both streams: split before the fork

            def params_p():
                res["dW_p"] = ops.linear_wgrad(dE, E_drop.view(B * T, d), lo=lo)
                res["db_p"] = ops.colsum(dE)

            fk.run(params_p, dE, E_drop)
            dE_drop = ops.linear_dgrad(dE, W_p, lo=lo)
        dVp, dgamma, dbeta, dls = ops.recavg_pool_bwd(dE_drop, E_raw, mean, rstd, wsum, Vp, r, t_hat, log_sigma, gamma, T, d,
                                                      ctx.thr, ctx.seed)
        if ctx.has_in:
            res["dW_in"] = ops.linear_wgrad(dVp, r.emb_flat, ragged=r.m_dev, lo=lo)
            res["db_in"] = ops.colsum(dVp, ragged=r.m_dev)
        fk.join(res["dW_p"], res["db_p"])
        dW_p, db_p, dW_in, db_in = res["dW_p"], res["db_p"], res["dW_in"], res["db_in"]
        return None, None, None, None, None, None, None, dls, dW_in, db_in, dgamma, dbeta, dW_p, db_p


# ============================================================== TTF_T2V_XAttn
class T2VXAttnFn(torch.autograd.Function):
    """TTF_T2V_XAttn.py:93-184, general schedule (any head count; eval mode).  K/V projections run once per note.  In eval /
    dropout-0 mode the T_f output rows of a sample are identical: they are computed once and returned as a broadcast view.
    (One head in train mode takes T2VXAttnFoldFn below.)"""

    @staticmethod
    def forward(ctx, r: RaggedNotes, T, H, thr, seed, save, defer, Qp, W_in, b_in, w_lin, b_lin, w_per, b_per, W_kv, b_kv,
                in_w, in_b, out_w, out_b, gamma, beta, W_po, b_po):
        """defer (train mode with attention dropout only): return dropout(LN(attn + Q)) WITHOUT proj_out (:182); the rank
        form of MMF_XAttn_Add folds W_po into its skinny operand and returns its gradient."""
        B = r.B
        d = W_po.shape[0]
        dt = d // 2
        dev = Qp.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        step = ops.step_ctx()
        lo = step.lo
        per_query = thr != 0  # attention dropout makes every (sample, query) row distinct
        # [V' ; phi] concat buffer (TTF_T2V_XAttn.py:139) and its tcgen05 lo operand, both written in place by the
        # two producers (input_proj epilogue / Time2Vec kernel): no separate split pass
        Xcat, Xcat_lo = new(r.M_alloc, d + dt), new(r.M_alloc, ops.round_up(d + dt, 4))
        extra = [(r.emb_flat, Xcat[:, :d], Xcat_lo[:, :d])] if W_in is None else []
        ws = [(W_in, [])] if W_in is not None else []
        ws += [(W_kv, []), (in_w, [slice(d, None)]), (out_w, [])] + ([] if defer else [(W_po, [])])
        ops.weight_los(lo, ws, extra)  # every weight matrix's lo in one launch
        if W_in is not None:
            ops.gemm(r.emb_flat, W_in, Xcat[:, :d], transB=True, bias=b_in, ragged=r.m_dev, ragged_dim=1, lo=lo,
                     emit_lo=Xcat_lo[:, :d])
        ops.time2vec_fwd(r, w_lin, b_lin, w_per, b_per, dt, Xcat[:, d:], Xcat_lo[:, d:])
        lo.put(Xcat, Xcat_lo)
        X = ops.linear_fwd(Xcat, W_kv, b_kv, ragged=r.m_dev, lo=lo, emit_lo=True)  # :140
        KVp = ops.linear_fwd(X, in_w[d:], in_b[d:], ragged=r.m_dev, lo=lo)  # MHA k/v in-projection, once per note
        scale = math.sqrt(1.0 / float(d // H))
        q0 = ops.linear_fwd(Qp.view(1, d), in_w[:d], in_b[:d])
        q = ops.axpby(q0, scale, torch.empty_like(q0), False)
        attn_cat, probs = ops.segattn_fwd(q, KVp, r, T, H, d, per_query, thr, seed, save)
        attn_out = ops.linear_fwd(attn_cat, out_w, out_b, lo=lo)
        rps = T if per_query else 1
        y, mean, rstd = ops.ln_fwd(attn_out, Qp.view(d), r.m_txt, rps, gamma, beta, thr, seed, ops.SITE_TTF_DROPOUT, save)
        assert not defer or per_query, "defer needs per-(sample, query) rows"
        E = y if defer else ops.linear_fwd(y, W_po, b_po, lo=lo, emit_lo=step.e_txt_feeds_tc)
        if save:
            ctx.r, ctx.T, ctx.H, ctx.thr, ctx.seed, ctx.lo, ctx.defer = r, T, H, thr, seed, lo, defer
            ctx.has_in, ctx.per_query, ctx.scale = W_in is not None, per_query, scale
            ctx.save_for_backward(Qp, w_per, b_per, W_kv, in_w, in_b, out_w, gamma, W_po, Xcat, X, KVp, q, attn_cat, probs,
                                  attn_out, y, mean, rstd)
        return E.view(B, T, d) if per_query else E.view(B, 1, d).expand(B, T, d)

    @staticmethod
    def backward(ctx, dE_txt):
        (Qp, w_per, b_per, W_kv, in_w, in_b, out_w, gamma, W_po, Xcat, X, KVp, q, attn_cat, probs, attn_out, y, mean,
         rstd) = ctx.saved_tensors
        r, T, H, thr, seed, lo = ctx.r, ctx.T, ctx.H, ctx.thr, ctx.seed, ctx.lo
        ctx.lo = None
        B, d = r.B, W_po.shape[0]
        dt = d // 2
        dev = Qp.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        per_query = ctx.per_query
        dEc = dE_txt.contiguous()
        dE = dEc.view(B * T, d) if per_query else ops.group_sum_rows(dEc.view(B * T, d), B, T, d)
        if ctx.defer:
            dW_po = db_po = None  # the consumer returns these
            dy = dE
        else:
            dW_po = ops.linear_wgrad(dE, y, lo=lo)
            db_po = ops.colsum(dE)
            dy = ops.linear_dgrad(dE, W_po, lo=lo)
        rps = T if per_query else 1
        dx, dres, dgamma, dbeta = ops.ln_bwd(dy, attn_out, Qp.view(d), r.m_txt, rps, gamma, mean, rstd, thr, seed,
                                             ops.SITE_TTF_DROPOUT)
        dW_o = ops.linear_wgrad(dx, attn_cat, lo=lo)
        db_o = ops.colsum(dx)
        d_attn_cat = ops.linear_dgrad(dx, out_w, lo=lo)
        dKVp, dq_partial = ops.segattn_bwd(d_attn_cat, q, KVp, probs, r, T, H, d, per_query, thr, seed)
        d_in_w = torch.empty_like(in_w)
        d_in_b = new(3 * d)
        # query path: q = (Qp W_q^T + b_q) * scale
        dq = ops.colsum(dq_partial)
        dq_pre = ops.axpby(dq, ctx.scale, torch.empty_like(dq), False).view(1, d)
        ops.linear_wgrad(dq_pre, Qp.view(1, d), out=d_in_w[:d])
        ops.axpby(dq_pre, 1.0, d_in_b[:d], False)
        dQp = ops.linear_dgrad(dq_pre, in_w[:d])  # [1,d]
        ops.axpby(dres, 1.0, dQp.view(d), True)
        # key/value path, once per note
        ops.linear_wgrad(dKVp, X, out=d_in_w[d:], ragged=r.m_dev, lo=lo)
        ops.colsum(dKVp, out=d_in_b[d:], ragged=r.m_dev)
        dX = ops.linear_dgrad(dKVp, in_w[d:], ragged=r.m_dev, lo=lo, emit_lo=True)
        dW_kv = ops.linear_wgrad(dX, Xcat, ragged=r.m_dev, lo=lo)
        db_kv = ops.colsum(dX, ragged=r.m_dev)
        dXcat_lo = new(r.M_alloc, ops.round_up(d + dt, 4)) if ctx.has_in else False
        dXcat = ops.linear_dgrad(dX, W_kv, ragged=r.m_dev, lo=lo, emit_lo=dXcat_lo)
        if ctx.has_in:
            lo.put(dXcat[:, :d], dXcat_lo[:, :d])
        dwl, dbl, dwp, dbp = ops.time2vec_bwd(dXcat[:, d:], r, w_per, b_per, dt)
        dW_in = db_in = None
        if ctx.has_in:
            dW_in = ops.linear_wgrad(dXcat[:, :d], r.emb_flat, ragged=r.m_dev, lo=lo)
            db_in = ops.colsum(dXcat[:, :d], ragged=r.m_dev)
        return (None, None, None, None, None, None, None, dQp.view(1, 1, d), dW_in, db_in, dwl, dbl, dwp, dbp, dW_kv, db_kv,
                d_in_w, d_in_b, dW_o, db_o, dgamma, dbeta, dW_po, db_po)


# ============================================================== TTF_T2V_XAttn, one head, train mode: collapsed schedule
class T2VXAttnFoldFn(torch.autograd.Function):
    """TTF_T2V_XAttn.py:93-184 for ONE head in train mode (attention dropout makes every (sample, query) row distinct).
    Exact reassociations of the reference's linear maps -- the per-note work is two dense products instead of four:

      * input_proj (:121) and the note half of KV_proj (:140) are consecutive linear maps: X = [emb ; phi] [W_a W_in | W_phi]^T
        + (W_a b_in + b_kv), with KV_proj.weight = [W_a | W_phi]; the d x d_model fold W_a W_in is a weight-space product;
      * ONE learned query: score_n = q . (W_k X_n + b_k) = (W_k^T q) . X_n + const, and the constant cancels in the softmax
        (d b_k = 0 exactly), so the key projection is one VECTOR u = W_k^T q and never a per-note product; dW_k = q (x) du;
      * the MHA out-projection is folded into the value projection (sum_n p_n (v_n W_o^T) = (sum_n p_n v_n) W_o^T).

    Per note: X (K = d_model + d_tau) and V' = X (W_o W_v)^T; backward: dX = dV' (W_o W_v) + ds (x) u, d phi = dX W_phi and the
    two weight gradients.  Everything that depends on parameters only runs on a side lane beside the data chain, and every
    weight gradient on one of three lanes beside the data-gradient chain (ops.Fork).  X lives in the K half and V' in the V half
    of the packed [rows, 2d] buffer the segment-attention kernels read (csrc/t2v_segattn.cu), with u in the place of q."""

    @staticmethod
    def forward(ctx, r: RaggedNotes, T, H, thr, seed, save, defer, Qp, W_in, b_in, w_lin, b_lin, w_per, b_per, W_kv, b_kv,
                in_w, in_b, out_w, out_b, gamma, beta, W_po, b_po):
        assert H == 1 and thr != 0
        B, d = r.B, W_po.shape[0]
        dt, dm = d // 2, r.d_m
        has_in = W_in is not None
        Kx = dm + dt
        dev = Qp.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        step = ops.step_ctx()
        lo = step.lo
        W_a, W_phi = W_kv[:, :d], W_kv[:, d:]
        W_q, W_k, W_v = in_w[:d], in_w[d:2 * d], in_w[2 * d:]
        # [emb ; phi] and its tcgen05 lo operand: the pad -> CSR gather wrote the notes (and their lo) into the left columns
        # already when the caller asked for the wide layout (FusionModel does); otherwise they are copied in here
        prefilled = r.emb_wide is not None and r.emb_wide.shape[1] == Kx
        if prefilled:
            Ecat, Ecat_lo = r.emb_wide, r.emb_wide_lo
        else:
            Ecat, Ecat_lo = new(r.M_alloc, Kx), new(r.M_alloc, ops.round_up(Kx, 4))
        KVp = new(r.M_alloc, 2 * d)  # [X | V']
        Wvf, Wvf_lo, bvf = new(d, d), new(d, d), new(d)
        if has_in:
            Wx, Wx_lo, bX = new(d, Kx), new(d, ops.round_up(Kx, 4)), new(d)
        else:
            Wx, bX = W_kv, b_kv
        scale = math.sqrt(1.0 / float(d))
        box = {}
        # parameter-only work on three lanes beside the data chain; the data chain waits for the operand it needs next
        fkw = ops.Fork(dev, lanes=3, name="t2v_fwd")

        def weights_x():  # lane 0: the operand of the first product
            if has_in:
                ops.weight_los(lo, [(W_in, []), (W_kv, [(slice(None), slice(0, d))])], [(W_phi, Wx[:, dm:], Wx_lo[:, dm:])])
                ops.gemm(W_a, W_in, Wx[:, :dm], lo=lo, emit_lo=Wx_lo[:, :dm])  # W_a W_in
                lo.put(Wx, Wx_lo)
                lo.put(Wx[:, dm:], Wx_lo[:, dm:])
            else:
                ops.weight_los(lo, [(W_kv, [(slice(None), slice(d, None))])])
            box["ev_x"] = fkw.mark(0)
            if save:
                lo.transposed(Wx[:, dm:])  # W_phi^T for d phi = dX W_phi (backward), off the critical path here

        def weights_v():  # lane 1: the operand of the second product
            ops.weight_los(lo, [(out_w, []), (in_w, [slice(2 * d, None)])] + ([] if defer else [(W_po, [])]))
            ops.gemm(out_w, W_v, Wvf, lo=lo, emit_lo=Wvf_lo)  # W_o W_v
            ops.gemm(in_b[2 * d:].view(1, d), out_w, bvf.view(1, d), transB=True)  # W_o b_v
            box["ev_v"] = fkw.mark(1)
            if save:
                lo.transposed(Wvf)  # (W_o W_v)^T for dX = dV' (W_o W_v) (backward)

        def vectors():  # lane 2: biases and the query side
            if has_in:
                ops.gemm(b_in.view(1, d), W_a, bX.view(1, d), transB=True, bias=b_kv)  # W_a b_in + b_kv
            box["ev_bx"] = fkw.mark(2)
            q0 = ops.linear_fwd(Qp.view(1, d), W_q, in_b[:d])
            q = ops.axpby(q0, scale, torch.empty_like(q0), False)
            box["q"], box["u"] = q, ops.gemm(q, W_k, new(1, d))  # u = W_k^T q
            box["ev_q"] = fkw.mark(2)

        fkw.run(weights_x, lane=0)
        fkw.run(weights_v, lane=1)
        fkw.run(vectors, lane=2)
        # data chain: [emb ; phi], X, V'
        if not prefilled:
            ops.multi_split([(r.emb_flat, Ecat[:, :dm], Ecat_lo[:, :dm])])
        ops.time2vec_fwd(r, w_lin, b_lin, w_per, b_per, dt, Ecat[:, dm:], Ecat_lo[:, dm:])
        lo.put(Ecat, Ecat_lo)
        u = box["u"]
        fkw.wait(box["ev_x"], Wx)
        fkw.wait(box["ev_bx"], bX)
        X = ops.gemm(Ecat, Wx, KVp[:, :d], transB=True, bias=bX, ragged=r.m_dev, ragged_dim=1, lo=lo, emit_lo=True)
        fkw.wait(box["ev_v"], Wvf, Wvf_lo, bvf)
        ops.gemm(X, Wvf, KVp[:, d:], transB=True, bias=bvf, ragged=r.m_dev, ragged_dim=1, lo=lo)
        fkw.wait(box["ev_q"], u, box["q"])
        fkw.join()
        if ops.segattn_ln_ok(d, r.N):
            y, attn_cat, probs, mean, rstd = ops.segattn_ln_fwd(u, KVp, r, T, d, thr, seed, out_b, Qp.view(d), gamma, beta, save)
        else:
            attn_cat, probs = ops.segattn_fwd(u, KVp, r, T, 1, d, True, thr, seed, save)
            y, mean, rstd = ops.ln_fwd(attn_cat, Qp.view(d), r.m_txt, T, gamma, beta, thr, seed, ops.SITE_TTF_DROPOUT, save, xbias=out_b)
        E = y if defer else ops.linear_fwd(y, W_po, b_po, out=step.final_out(B * T, d), lo=lo, emit_lo=step.e_txt_feeds_tc)
        if save:
            ctx.r, ctx.T, ctx.thr, ctx.seed, ctx.lo, ctx.defer, ctx.has_in, ctx.scale = r, T, thr, seed, lo, defer, has_in, scale
            ctx.save_for_backward(Qp, w_per, b_per, W_in, b_in, W_kv, in_w, in_b, out_w, out_b, gamma, W_po, Ecat, KVp, Wx, Wvf,
                                  box["q"], u, attn_cat, probs, y, mean, rstd)
        return E.view(B, T, d)

    @staticmethod
    def backward(ctx, dE_txt):
        """Data parallel (ops.DP_GROUP set by runtime.GraphedStep): every parameter gradient of this module is a linear function
        (with replicated parameters) of five packed buffers -- [dW_vf | db_vf], [dW_x | db_x], [dres | dgamma | dbeta | du | db_o],
        the Time2Vec gradients and, when proj_out is not deferred, [dW_po | db_po] -- so THOSE are all-reduced, each on the lane
        that produces it and before the weight-space un-folds: 1.5 M floats instead of the module's 3.8 M at cfg2, overlapped
        with the other lanes, and nothing is left to reduce when backward ends."""
        (Qp, w_per, b_per, W_in, b_in, W_kv, in_w, in_b, out_w, out_b, gamma, W_po, Ecat, KVp, Wx, Wvf, q, u, attn_cat, probs, y,
         mean, rstd) = ctx.saved_tensors
        r, T, thr, seed, lo, has_in = ctx.r, ctx.T, ctx.thr, ctx.seed, ctx.lo, ctx.has_in
        ctx.lo = None
        B, d = r.B, W_po.shape[0]
        dt, dm = d // 2, r.d_m
        Kx = dm + dt
        dev = Qp.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        W_a = W_kv[:, :d]
        W_q, W_k, W_v = in_w[:d], in_w[d:2 * d], in_w[2 * d:]
        tc = ops.gemm_backend() != ops.BACKEND_FFMA
        dp = ops.DP_GROUP is not None
        # single GPU with an input projection: a fourth lane takes the X side's small products (everything that needs db_x only)
        # and the Time2Vec gradients, so that the grouped un-fold starts as soon as the big weight gradient dW_x is reduced --
        # the timeline of the step (profiles/r2_timeline_cfg2_n1_final.txt) had 40 us of tiny dependent kernels between the two
        # at the very end of backward.  IMMTSF_T2V_TAIL_LANE=0 keeps the three-lane order (A/B runs); the data-parallel schedule
        # is unchanged (its lane order carries the all-reduces).
        lane4 = os.environ.get("IMMTSF_T2V_TAIL_LANE", "1") != "0"
        tail_lane = (not dp) and has_in and lane4
        # data parallel: the same fourth lane takes db_x and the Time2Vec gradients (both land in pack_x BEFORE its all-reduce), so
        # lane 2's all-reduce follows the big weight gradient directly; the collectives, their order and their buffers are unchanged
        dp_lane = dp and lane4
        fk = ops.Fork(dev, lanes=4 if (tail_lane or dp_lane) else 3)
        dE = dE_txt.contiguous().view(B * T, d)
        if ctx.defer:
            dW_po = db_po = None  # the consumer returns these
            dy = dE
        else:
            pack_po = ops.dp_pack(d * d + d, dev)
            dW_po, db_po = pack_po[:d * d].view(d, d), pack_po[d * d:]
            ops.linear_wgrad(dE, y, out=dW_po, lo=lo)
            ops.colsum(dE, out=db_po)
            if dp:
                fk.run(lambda: ops.dp_allreduce(pack_po), pack_po, lane=0)
            dy = ops.linear_dgrad(dE, W_po, lo=lo)
        # two packed buffers hold every gradient statistic of the module (data parallel: ONE all-reduce each, issued on a lane
        # as soon as the last contribution has landed; the lanes wait for each other through events, the data chain never waits)
        pack_qv = ops.dp_pack(5 * d + d * d + d, dev)  # dres | dgamma | dbeta (accumulated by ln_bwd) | du | db_o | dW_vf | db_vf
        pack_qv[:3 * d].zero_()
        pack_x = ops.dp_pack(d * Kx + d + 2 * dt, dev)  # dW_x | db_x | Time2Vec gradients (accumulated)
        pack_x[d * Kx + d:].zero_()
        dx, dres, dgamma, dbeta = ops.ln_bwd(dy, attn_cat, Qp.view(d), r.m_txt, T, gamma, mean, rstd, thr, seed, ops.SITE_TTF_DROPOUT,
                                             xbias=out_b, acc=pack_qv[:3 * d])
        # K half of dKVp: ds_n u (the score path's contribution to dX); V half: dV'; dq_partial: sum_n ds_n X_n = du per sample
        dKVp, du_partial = ops.segattn_bwd(dx, u, KVp, probs, r, T, 1, d, True, thr, seed)
        X, dXk, dVf = KVp[:, :d], dKVp[:, :d], dKVp[:, d:]
        if tc:
            lo.lo_for(dVf, r.m_dev)  # read on two streams: split before the fork
        d_in_w, d_in_b = torch.empty_like(in_w), new(3 * d)
        res, ev = {}, {}

        def sums_query():
            res["du"] = ops.colsum(du_partial, out=pack_qv[3 * d:4 * d]).view(1, d)
            res["db_o"] = ops.colsum(dx, out=pack_qv[4 * d:5 * d])
            ev["q"] = fk.mark(1)

        dWvf, dbvf = pack_qv[5 * d:5 * d + d * d].view(d, d), pack_qv[5 * d + d * d:]
        dW_o = new(d, d)

        def params_value():  # V' = X (W_o W_v)^T + W_o b_v: the statistics; the un-fold joins the X side's (one grouped launch)
            ops.linear_wgrad(dVf, X, out=dWvf, ragged=r.m_dev, lo=lo, emit_lo=False if dp else new(d, d))
            ops.colsum(dVf, out=dbvf, ragged=r.m_dev)
            if dp:
                fk.lane_wait(0, ev["q"])
                ops.dp_allreduce(pack_qv)
                ev["qv"] = fk.mark(0)
            ops.gemm(dbvf.view(1, d), out_w, d_in_b[2 * d:].view(1, d))
            if tail_lane or dp_lane:
                ops.gemm(dbvf.view(d, 1), in_b[2 * d:].view(1, d), dW_o)  # W_o b_v also depends on W_o (the un-fold accumulates onto it)
            ev["v"] = fk.mark(0)
            res["dW_o"] = dW_o

        def params_query():  # u = W_k^T q,  q = (Qp W_q^T + b_q) scale
            if dp:
                fk.lane_wait(1, ev["qv"])
            du = res["du"]
            ops.gemm(q.view(d, 1), du, d_in_w[d:2 * d])  # dW_k = q (x) du
            ops.axpby(du, 0.0, d_in_b[d:2 * d], False)  # d b_k = 0: q . b_k is constant over a segment
            dq = ops.gemm(du, W_k, new(1, d), transB=True)
            dq_pre = ops.axpby(dq, ctx.scale, torch.empty_like(dq), False)
            ops.linear_wgrad(dq_pre, Qp.view(1, d), out=d_in_w[:d])
            ops.axpby(dq_pre, 1.0, d_in_b[:d], False)
            dQp = ops.linear_dgrad(dq_pre, W_q)
            ops.axpby(dres, 1.0, dQp.view(d), True)
            res["dQp"] = dQp

        fk.run(sums_query, dx, du_partial, pack_qv, lane=1)
        fk.run(params_value, dKVp, pack_qv, d_in_w, d_in_b, lane=0)
        fk.run(params_query, d_in_w, d_in_b, lane=1, after_current=False)
        dX = ops.linear_dgrad(dVf, Wvf, out=dXk, beta=1.0, ragged=r.m_dev, lo=lo, emit_lo=True)  # + ds (x) u, in place
        dWx, dbX = pack_x[:d * Kx].view(d, Kx), pack_x[d * Kx:d * Kx + d]
        dWx_lo = new(d, ops.round_up(Kx, 4)) if (has_in and not dp) else False

        def sums_x():  # X = [emb ; phi] [W_a W_in | W_phi]^T + (W_a b_in + b_kv)
            ops.linear_wgrad(dX, Ecat, out=dWx, ragged=r.m_dev, lo=lo, emit_lo=dWx_lo)
            if not (tail_lane or dp_lane):
                ops.colsum(dX, out=dbX, ragged=r.m_dev)

        if tail_lane:
            dW_kv_t, dW_in_t, db_in_t = new(d, d + dt), new(d, dm), new(d)

        def smalls_x():  # lane 3: db_x and the two products that need nothing else, beside the big weight gradient of lane 2
            ops.colsum(dX, out=dbX, ragged=r.m_dev)
            ops.gemm(dbX.view(d, 1), b_in.view(1, d), dW_kv_t[:, :d])  # W_a b_in also depends on W_a (the un-fold accumulates onto it)
            ops.gemm(dbX.view(1, d), W_a, db_in_t.view(1, d))
            ev["xs"] = fk.mark(3)

        fk.run(sums_x, dX, pack_x, *([lo.lo_for(dX, r.m_dev)] if fk.side is not None and tc else []), lane=2)
        if tail_lane:
            fk.run(smalls_x, dX, pack_x, dW_kv_t, db_in_t, lane=3)
        elif dp_lane:
            fk.run(lambda: ops.colsum(dX, out=dbX, ragged=r.m_dev), dX, pack_x, lane=3)
        dphi = ops.linear_dgrad(dX, Wx[:, dm:], out=new(r.M_alloc, dt), ragged=r.m_dev, lo=lo)

        def params_t2v():
            res["t2v"] = ops.time2vec_bwd(dphi, r, w_per, b_per, dt, buf=pack_x[d * Kx + d:])

        def params_x():
            if dp:
                if dp_lane:
                    fk.lane_wait(2, ev["x3"])  # db_x and the Time2Vec gradients (lane 3) are in pack_x
                ops.dp_allreduce(pack_x)
            # the rank-1 bias terms first (K = 1 products, they overwrite), the grouped d x d products then ACCUMULATE onto them:
            # nothing is left to do after the big launch
            unfold_v = [dict(A=dWvf, B=W_v, C=dW_o, transB=True, beta=1.0), dict(A=out_w, B=dWvf, C=d_in_w[2 * d:], transA=True)]
            if not has_in:
                fk.lane_wait(2, ev["v"])
                if not dp_lane:
                    ops.gemm(dbvf.view(d, 1), in_b[2 * d:].view(1, d), dW_o)  # W_o b_v also depends on W_o
                ops.gemm_group(unfold_v, lo)
                res["dW_kv"], res["db_kv"], res["dW_in"], res["db_in"] = dWx, dbX, None, None
                return
            dP1 = dWx[:, :dm]
            if dWx_lo is not False:
                lo.put(dP1, dWx_lo[:, :dm])
            if tail_lane:  # the small products are lane 3's: copy the phi columns, wait for both sides' rank-1 terms, un-fold
                dW_kv, dW_in, db_in = dW_kv_t, dW_in_t, db_in_t
                ops.multi_split([(dWx[:, dm:], dW_kv[:, d:], None)])
                fk.lane_wait(2, ev["xs"])
                fk.lane_wait(2, ev["v"])
                ops.gemm_group(unfold_v + [dict(A=dP1, B=W_in, C=dW_kv[:, :d], transB=True, beta=1.0), dict(A=W_a, B=dP1, C=dW_in, transA=True)], lo)
                res["dW_kv"], res["db_kv"], res["dW_in"], res["db_in"] = dW_kv, dbX, dW_in, db_in
                return
            dW_kv, dW_in, db_in = new(d, d + dt), new(d, dm), new(d)
            ops.gemm(dbX.view(d, 1), b_in.view(1, d), dW_kv[:, :d])  # W_a b_in also depends on W_a
            ops.gemm(dbX.view(1, d), W_a, db_in.view(1, d))
            ops.multi_split([(dWx[:, dm:], dW_kv[:, d:], None)])
            fk.lane_wait(2, ev["v"])  # the value side's statistics (lane 0) are final
            if not dp_lane:
                ops.gemm(dbvf.view(d, 1), in_b[2 * d:].view(1, d), dW_o)  # W_o b_v also depends on W_o
            # both weight-space un-folds (four d x d products) in ONE grouped launch
            ops.gemm_group(unfold_v + [dict(A=dP1, B=W_in, C=dW_kv[:, :d], transB=True, beta=1.0), dict(A=W_a, B=dP1, C=dW_in, transA=True)], lo)
            res["dW_kv"], res["db_kv"], res["dW_in"], res["db_in"] = dW_kv, dbX, dW_in, db_in

        if dp_lane:
            fk.run(params_t2v, dphi, pack_x, lane=3)
            ev["x3"] = fk.mark(3)
            fk.run(params_x, lane=2, after_current=False)
        elif dp:
            fk.run(params_t2v, dphi, pack_x, lane=2)  # same lane as dW_x: the lane's all-reduce follows both
            fk.run(params_x, lane=2, after_current=False)
        elif tail_lane:
            fk.run(params_x, dW_kv_t, dW_in_t, db_in_t, lane=2, after_current=False)  # the un-fold does not wait for d phi
            fk.run(params_t2v, dphi, pack_x, lane=3)  # (after lane 3's small products; not behind the query path of lane 1)
        else:
            fk.run(params_x, lane=2, after_current=False)  # the un-fold does not wait for d phi
            fk.run(params_t2v, dphi, pack_x, lane=1)
        db_o = res["db_o"]
        dwl, dbl, dwp, dbp = res["t2v"]
        dQp, dW_kv, db_kv, dW_in, db_in, dW_o = (res[k] for k in ("dQp", "dW_kv", "db_kv", "dW_in", "db_in", "dW_o"))
        fk.join(dQp, dW_in, db_in, dW_kv, d_in_w, d_in_b, dW_o, pack_qv, pack_x, dW_po, db_po)
        return (None, None, None, None, None, None, None, dQp.view(1, 1, d), dW_in, db_in, dwl, dbl, dwp, dbp, dW_kv, db_kv,
                d_in_w, d_in_b, dW_o, db_o, dgamma, dbeta, dW_po, db_po)


# ============================================================== TTF_T2V_XAttn, per-(note, query) variant
class T2VPerQueryFn(torch.autograd.Function):
    """fusions/TTF_T2V_XAttn_old.py:82-161: Time2Vec of the clamped lag of every (note, query) pair enters keys and
    values.  The reference projects B*T*N concatenated rows twice (KV_proj, then the MHA in-projection); here, with
    X_nt = A_n + W_phi phi_nt, the projections run once per note (A) and once per (sample, query, head) row (Z, Phi), and
    the fused kernel between them never forms a per-pair vector of width d (csrc/t2v_perquery.cu).  The composition is
    stated in plain torch in oracle/perquery_schedule.py and checked there against the reference semantics."""

    @staticmethod
    def forward(ctx, r: RaggedNotes, t_hat, T, H, thr, seed, save, Qp, W_in, b_in, w_lin, b_lin, w_per, b_per, W_kv, b_kv,
                in_w, in_b, out_w, out_b, gamma, beta, W_po, b_po):
        B = r.B
        d = W_po.shape[0]
        dt, hd = d // 2, d // H
        dev = Qp.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        step = ops.step_ctx()
        lo = step.lo
        W_a, W_phi = W_kv[:, :d], W_kv[:, d:]
        W_q, W_k, W_v, b_v = in_w[:d], in_w[d:2 * d], in_w[2 * d:], in_b[2 * d:]
        # tcgen05 lo operands of every weight matrix (and of the views used as operands on their own) in one launch
        heads = [slice(2 * d + h * hd, 2 * d + (h + 1) * hd) for h in range(H)] if H > 1 else []
        ws = ([(W_in, [])] if W_in is not None else []) + [
            (W_kv, [(slice(None), slice(0, d)), (slice(None), slice(d, None))]),
            (in_w, [slice(d, 2 * d), slice(2 * d, None)] + heads), (out_w, []), (W_po, [])]
        ops.weight_los(lo, ws)
        Vp = ops.linear_fwd(r.emb_flat, W_in, b_in, ragged=r.m_dev, lo=lo) if W_in is not None else r.emb_flat
        A = ops.linear_fwd(Vp, W_a, b_kv, ragged=r.m_dev, lo=lo)  # the note part of KV_proj (:129), once per note
        # query side: u_h = W_k[h]^T q_h (block-diagonal q), g_h = W_phi^T u_h
        scale = math.sqrt(1.0 / float(hd))
        q0 = ops.linear_fwd(Qp.view(1, d), W_q, in_b[:d])
        q = ops.axpby(q0, scale, torch.empty_like(q0), False)
        Qblk = torch.zeros(H, d, dtype=_f32, device=dev)
        for h in range(H):
            Qblk[h, h * hd:(h + 1) * hd].copy_(q[0, h * hd:(h + 1) * hd])
        U = ops.gemm(Qblk, W_k, new(H, d))
        a_sc = ops.linear_fwd(A, U, None, ragged=r.m_dev)  # [M_alloc, H]
        g = ops.gemm(U, W_phi, new(H, dt))
        t2v = (w_lin, b_lin, w_per, b_per)
        Z, Phi, sp, probs = ops.t2vq_attn_fwd(A, a_sc, g, r, t_hat, t2v, T, H, d, dt, thr, seed, save)
        XZ = ops.gemm(Phi, W_phi, Z, transB=True, beta=1.0, lo=lo)  # Z + Phi W_phi^T, in place
        XZv, spv = XZ.view(B * T, H * d), sp.view(B * T, H)
        O = new(B * T, d)
        for h in range(H):
            hs = slice(h * hd, (h + 1) * hd)
            ops.gemm(XZv[:, h * d:(h + 1) * d], W_v[hs], O[:, hs], transB=True, lo=lo)
            ops.gemm(spv[:, h:h + 1], b_v[hs].view(1, hd), O[:, hs], beta=1.0)  # sum_n P~ b_v (P~ does not sum to 1 under dropout)
        attn_out = ops.linear_fwd(O, out_w, out_b, lo=lo)
        y, mean, rstd = ops.ln_fwd(attn_out, Qp.view(d), r.m_txt, T, gamma, beta, thr, seed, ops.SITE_TTF_DROPOUT, save)
        E = ops.linear_fwd(y, W_po, b_po, out=step.final_out(B * T, d), lo=lo, emit_lo=step.e_txt_feeds_tc)
        if save:
            ctx.r, ctx.T, ctx.H, ctx.thr, ctx.seed, ctx.lo, ctx.scale = r, T, H, thr, seed, lo, scale
            ctx.has_in = W_in is not None
            ctx.save_for_backward(Qp, w_lin, b_lin, w_per, b_per, W_kv, in_w, in_b, out_w, gamma, W_po, t_hat, Vp, A, Qblk, U, g, Phi, sp, probs,
                                  XZ, O, attn_out, y, mean, rstd)
        return E.view(B, T, d)

    @staticmethod
    def backward(ctx, dE_txt):
        (Qp, w_lin, b_lin, w_per, b_per, W_kv, in_w, in_b, out_w, gamma, W_po, t_hat, Vp, A, Qblk, U, g, Phi, sp, probs, XZ, O,
         attn_out, y, mean, rstd) = ctx.saved_tensors
        r, T, H, thr, seed, lo, scale = ctx.r, ctx.T, ctx.H, ctx.thr, ctx.seed, ctx.lo, ctx.scale
        ctx.lo = None
        B, d = r.B, W_po.shape[0]
        dt, hd = d // 2, d // H
        dev = Qp.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        W_a, W_phi = W_kv[:, :d], W_kv[:, d:]
        W_q, W_k, W_v, b_v = in_w[:d], in_w[d:2 * d], in_w[2 * d:], in_b[2 * d:]
        dE = dE_txt.contiguous().view(B * T, d)
        dW_po = ops.linear_wgrad(dE, y, lo=lo)
        db_po = ops.colsum(dE)
        dy = ops.linear_dgrad(dE, W_po, lo=lo)
        dx, dres, dgamma, dbeta = ops.ln_bwd(dy, attn_out, Qp.view(d), r.m_txt, T, gamma, mean, rstd, thr, seed, ops.SITE_TTF_DROPOUT)
        dW_o = ops.linear_wgrad(dx, O, lo=lo)
        db_o = ops.colsum(dx)
        dO = ops.linear_dgrad(dx, out_w, lo=lo)
        d_in_w, d_in_b = new(3 * d, d), new(3 * d)
        dXZ, dsp = new(B * T * H, d), new(B * T * H)
        XZv, spv, dXZv, dspv = XZ.view(B * T, H * d), sp.view(B * T, H), dXZ.view(B * T, H * d), dsp.view(B * T, H)
        for h in range(H):
            hs = slice(h * hd, (h + 1) * hd)
            vs = slice(2 * d + h * hd, 2 * d + (h + 1) * hd)
            dO_h = dO[:, hs]
            ops.gemm(dO_h, W_v[hs], dXZv[:, h * d:(h + 1) * d], lo=lo)  # [BT,hd] x [hd,d]
            ops.gemm(dO_h, XZv[:, h * d:(h + 1) * d], d_in_w[vs], transA=True, lo=lo)  # dW_v[h]
            ops.gemm(dO_h, spv[:, h:h + 1], d_in_b[vs].view(hd, 1), transA=True)  # db_v[h] = dO_h^T sp_h
            ops.gemm(dO_h, b_v[hs].view(hd, 1), dspv[:, h:h + 1])  # dsp_h = dO_h b_v[h]
        dPhi = ops.gemm(dXZ, W_phi, new(B * T * H, dt), lo=lo)
        dW_kv = new(d, d + dt)
        ops.gemm(dXZ, Phi, dW_kv[:, d:], transA=True, lo=lo)  # dW_phi = dXZ^T Phi
        t2v = (w_lin, b_lin, w_per, b_per)
        dA, da, dpart = ops.t2vq_attn_bwd(dXZ, dPhi, dsp, A, g, probs, r, t_hat, t2v, T, H, d, dt, thr, seed)
        tg = ops.colsum(dpart)  # [(2+H)*dt]: d w, d b, dg
        dg = tg[2 * dt:].view(H, dt)
        dU = ops.linear_wgrad(da, A, ragged=r.m_dev)  # [H,d] = da^T A
        ops.gemm(dg, W_phi, dU, transB=True, beta=1.0)  # + dg W_phi^T
        ops.gemm(da, U, dA, beta=1.0, ragged=r.m_dev, ragged_dim=1)  # dA += da U
        ops.gemm(U, dg, dW_kv[:, d:], transA=True, beta=1.0)  # dW_phi += U^T dg
        ops.gemm(Qblk, dU, d_in_w[d:2 * d], transA=True)  # dW_k: row j of head h is q_j dU_h
        d_in_b[d:2 * d].zero_()  # q_h . b_k[h] is constant over the notes of a segment: it cancels in the softmax
        dQblk = ops.gemm(dU, W_k, new(H, d), transB=True)
        dq = new(1, d)
        for h in range(H):
            dq[0, h * hd:(h + 1) * hd].copy_(dQblk[h, h * hd:(h + 1) * hd])
        dq_pre = ops.axpby(dq, scale, torch.empty_like(dq), False)
        ops.linear_wgrad(dq_pre, Qp.view(1, d), out=d_in_w[:d])
        d_in_b[:d].copy_(dq_pre.view(d))
        dQp = ops.linear_dgrad(dq_pre, W_q)
        ops.axpby(dres, 1.0, dQp.view(d), True)
        ops.linear_wgrad(dA, Vp, out=dW_kv[:, :d], ragged=r.m_dev, lo=lo)
        db_kv = ops.colsum(dA, ragged=r.m_dev)
        dW_in = db_in = None
        if ctx.has_in:
            dVp = ops.linear_dgrad(dA, W_a, ragged=r.m_dev, lo=lo)
            dW_in = ops.linear_wgrad(dVp, r.emb_flat, ragged=r.m_dev, lo=lo)
            db_in = ops.colsum(dVp, ragged=r.m_dev)
        dwl, dbl = tg[:1].view(1, 1), tg[dt:dt + 1]
        dwp, dbp = tg[1:dt].view(dt - 1, 1), tg[dt + 1:2 * dt]
        return (None, None, None, None, None, None, None, dQp.view(1, 1, d), dW_in, db_in, dwl, dbl, dwp, dbp, dW_kv, db_kv,
                d_in_w, d_in_b, dW_o, db_o, dgamma, dbeta, dW_po, db_po)


# ============================================================== MMF_GR_Add
class GRAddFn(torch.autograd.Function):
    """MMF_GR_Add.py:31-61.  x = [E ; Y] (the reference's [Y ; E] with the two blocks swapped, so that the wide block starts
    16-byte aligned and the TTF's final projection can write E_txt straight into it -- ops.StepCtx.x_cat; the columns of
    [W_ih ; W_g] are permuted the same way): the GRU input map and the gate net read x once, as ONE skinny product.
    No ATen kernel on the path: Y, the permuted weights and the biases are placed by one immtsf_multi_split launch."""

    @staticmethod
    def forward(ctx, Y, E, m_txt, thr, seed, save, flags, W_ih, W_hh, b_ih, b_hh, W_r, b_r, W_g, b_g, gamma, beta):
        B, T, C = Y.shape
        d = E.shape[2]
        dev = Y.device
        Yc = Y.contiguous()
        Kp = ops.round_up(C + d, 4)  # row length padded to 16 bytes
        step = ops.step_ctx()
        X = step.x_cat
        in_place = (X is not None and tuple(X.shape) == (B * T, Kp) and E.data_ptr() == X.data_ptr() and E.stride(-1) == 1
                    and E.stride(-2) == Kp and (T == 1 or E.stride(0) == T * Kp))
        tasks = []
        if not in_place:
            X = torch.empty(B * T, Kp, dtype=_f32, device=dev)
            E2 = E.reshape(B * T, d)  # (a broadcast view -- T2V eval mode -- is materialised here)
            tasks.append((E2, X[:, :d], None))
        Wcat = torch.empty(4 * C, Kp, dtype=_f32, device=dev)
        bcat = torch.empty(4 * C, dtype=_f32, device=dev)
        tasks += [(Yc.view(B * T, C), X[:, d:d + C], None),
                  (W_ih[:, C:], Wcat[:3 * C, :d], None), (W_ih[:, :C], Wcat[:3 * C, d:d + C], None),
                  (W_g[:, C:], Wcat[3 * C:, :d], None), (W_g[:, :C], Wcat[3 * C:, d:d + C], None),
                  (b_ih, bcat[:3 * C], None), (b_g, bcat[3 * C:], None)]
        if Kp > C + d:
            z = ops.zeros_cached(dev, max(B * T, 4 * C), Kp - C - d)
            tasks += [(z[:B * T], X[:, C + d:], None), (z[:4 * C], Wcat[:, C + d:], None)]
        ops.multi_split(tasks)
        G4 = ops.linear_fwd(X, Wcat, bcat)
        W_hh_c, b_hh_c, W_r_c = W_hh.contiguous(), b_hh.contiguous(), W_r.contiguous()
        h_all, h_prev, gates = ops.gru_scan_fwd(G4, W_hh_c, b_hh_c, B, T, C)
        Y_out = ops.gr_tail_fwd(Yc, G4, h_all, W_r_c, b_r, gamma, beta, m_txt, B, T, C, thr, seed, flags)
        if save:
            ctx.thr, ctx.seed, ctx.dims = thr, seed, (B, T, C, d)
            ctx.save_for_backward(m_txt, X, Wcat, G4, h_all, h_prev, W_hh_c, b_hh_c, W_r_c, b_r, gamma, beta, gates)
        return Y_out

    @staticmethod
    def backward(ctx, dY_out):
        m_txt, X, Wcat, G4, h_all, h_prev, W_hh, b_hh, W_r, b_r, gamma, beta, gates = ctx.saved_tensors
        B, T, C, d = ctx.dims
        dev = X.device
        dY_out = dY_out.contiguous()
        dG4 = torch.empty_like(G4)
        d_delta, dh_out, dgamma, dbeta = ops.gr_tail_bwd(dY_out, G4, h_all, W_r, b_r, gamma, beta, m_txt, B, T, C, ctx.thr,
                                                         ctx.seed, dG4)
        dGh = ops.gru_scan_bwd(G4, h_prev, W_hh, b_hh, dh_out, B, T, C, dG4, gates)
        fk = ops.Fork(dev, name="gr_bwd")  # every weight / bias gradient beside the data-gradient products (dY, dE)
        res = {}

        def params():
            res["dW_r"] = ops.linear_wgrad(d_delta, h_all)
            res["db_r"] = ops.colsum(d_delta)
            res["dW_hh"] = ops.linear_wgrad(dGh, h_prev)
            res["db_hh"] = ops.colsum(dGh)
            dWcat = ops.linear_wgrad(dG4, X)
            res["dbcat"] = ops.colsum(dG4)
            # back to the reference's column order [Y block | E block]
            dW_ih, dW_g = torch.empty(3 * C, C + d, dtype=_f32, device=dev), torch.empty(C, C + d, dtype=_f32, device=dev)
            ops.multi_split([(dWcat[:3 * C, d:d + C], dW_ih[:, :C], None), (dWcat[:3 * C, :d], dW_ih[:, C:], None),
                             (dWcat[3 * C:, d:d + C], dW_g[:, :C], None), (dWcat[3 * C:, :d], dW_g[:, C:], None)])
            res["dW_ih"], res["dW_g"] = dW_ih, dW_g

        fk.run(params, d_delta, h_all, dGh, h_prev, dG4, X)
        # dY = dY_out (the blend passes Y straight through) + dG4 Wcat[:, Y block] ;  dE = dG4 Wcat[:, E block]
        dY = ops.gemm(dG4, Wcat[:, d:d + C], torch.empty(B * T, C, dtype=_f32, device=dev))
        ops.axpby(dY_out.view(B * T, C), 1.0, dY, True)
        dE = ops.gemm(dG4, Wcat[:, :d], torch.empty(B * T, d, dtype=_f32, device=dev))
        dW_r, db_r, dW_hh, db_hh, dbcat, dW_ih, dW_g = (res[k] for k in ("dW_r", "db_r", "dW_hh", "db_hh", "dbcat", "dW_ih", "dW_g"))
        fk.join(dW_r, db_r, dW_hh, db_hh, dbcat, dW_ih, dW_g)
        return (dY.view(B, T, C), dE.view(B, T, d), None, None, None, None, None, dW_ih, dW_hh, dbcat[: 3 * C],
                db_hh, dW_r, db_r, dW_g, dbcat[3 * C:], dgamma, dbeta)


# ============================================================== MMF_XAttn_Add
class XAttnAddFn(torch.autograd.Function):
    """MMF_XAttn_Add.py:56-103.  The reference chains two bias-free projections into each MHA in-projection
    (proj_q -> in_proj_q, proj_k -> in_proj_k, proj_v -> in_proj_v, :68-76) and the MHA out-projection into
    residual_head (:76,:83).  Consecutive linear maps are folded in WEIGHT space -- (E W_K^T) W_k^T = E (W_k W_K)^T --
    so the per-row work is one [B*T, d_txt] x [d_txt, 2d] projection instead of six d x d ones; the folds are d^3
    (or rank-C) products, and backward un-folds the weight gradients the same way (the chain rule of the fold).

    Rank-(C+1) query path (T <= 32, the Time-IMM windows): q_i = Wq_f y_i + b_q is a projection of the C-channel
    series, so q_i . k_j = [y_i ; 1] . kq_j with kq_j = [Wq_f^T k_j ; b_q . k_j].  kq [B*T, H*(C+1)] is one skinny product
    per head; q and dq ([B*T, d] each) are never formed (csrc/xattn_small.cu, include/immtsf.h)."""

    @staticmethod
    def forward(ctx, Y, E, m_txt, H, kappa, thr, seed, save, flags, W_Q, W_K, W_V, in_w, in_b, out_w, out_b, W_r, b_r,
                gamma, beta):
        B, T, C = Y.shape
        d, de = W_Q.shape[0], E.shape[2]
        hd, C1 = d // H, C + 1
        dev = Y.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        Y2 = Y.contiguous().view(B * T, C)
        E2 = E.contiguous().view(B * T, de)
        lo = ops.step_ctx().lo
        kv = new(B * T, 2 * d)
        lowrank = ops.xattn_lowrank_ok(T, H, d, C, kv[:, d:])
        # ---- weight-space folds (their epilogues also write the folded operand's lo)
        Wq_aug = new(d, C1)  # [Wq_f | b_q]
        extra = [(in_b[:d].view(d, 1), Wq_aug[:, C:], None)] if lowrank else []
        ops.weight_los(lo, [(W_K, []), (W_V, []), (in_w, [slice(d, 2 * d), slice(2 * d, None)])], extra)
        Wq_f = ops.gemm(in_w[:d], W_Q, Wq_aug[:, :C])  # [d, C]
        Wkv_f = new(2 * d, de)
        Wkv_f_lo = new(2 * d, ops.round_up(de, 4))
        ops.gemm_group([dict(A=in_w[d:2 * d], B=W_K, C=Wkv_f[:d], emit_lo=Wkv_f_lo[:d]),
                        dict(A=in_w[2 * d:], B=W_V, C=Wkv_f[d:], emit_lo=Wkv_f_lo[d:])], lo)  # one launch
        lo.put(Wkv_f, Wkv_f_lo)
        Wo_f = ops.gemm(W_r, out_w, new(C, d))  # [C, d]
        bo_f = ops.gemm(out_b.view(1, d), W_r, new(1, C), transB=True, bias=b_r).view(C)
        # ---- per-row work
        ops.linear_fwd(E2, Wkv_f, in_b[d:], out=kv, lo=lo)  # :69-70 + in_proj_k / in_proj_v, E read once
        if lowrank:
            q = None
            kq = new(B * T, H * C1)
            for h in range(H):
                ops.gemm(kv[:, h * hd:(h + 1) * hd], Wq_aug[h * hd:(h + 1) * hd], kq[:, h * C1:(h + 1) * C1])
            o, probs = ops.xattn_lowrank_fwd(Y2, kq, kv[:, d:], m_txt, B, T, H, d, C, thr, seed, save)
        else:
            kq = None
            q = ops.linear_fwd(Y2, Wq_f, in_b[:d])  # :68 + in_proj_q
            o, probs = ops.xattn_core_fwd(q, kv[:, :d], kv[:, d:], m_txt, B, T, H, d, thr, seed, save, lo=lo)
        delta_y = ops.linear_fwd(o, Wo_f, bo_f)  # out_proj + residual_head (:83)
        Y_out = ops.xattn_tail_fwd(Y2, delta_y, gamma, beta, m_txt, B, T, C, kappa, thr, seed, flags)
        if save:
            ctx.H, ctx.kappa, ctx.thr, ctx.seed, ctx.dims, ctx.lo, ctx.lowrank = H, kappa, thr, seed, (B, T, C, d, de), lo, lowrank
            ctx.save_for_backward(m_txt, Y2, E2, W_Q, W_K, W_V, in_w, out_w, out_b, W_r, gamma, Wq_aug, Wkv_f, Wo_f, q, kq, kv, o,
                                  probs, delta_y)
        return Y_out

    @staticmethod
    def backward(ctx, dY_out):
        (m_txt, Y2, E2, W_Q, W_K, W_V, in_w, out_w, out_b, W_r, gamma, Wq_aug, Wkv_f, Wo_f, q, kq, kv, o, probs,
         delta_y) = ctx.saved_tensors
        B, T, C, d, de = ctx.dims
        H, kappa, thr, seed, lo, lowrank = ctx.H, ctx.kappa, ctx.thr, ctx.seed, ctx.lo, ctx.lowrank
        ctx.lo = None
        hd, C1 = d // H, C + 1
        dev = Y2.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        Wq_f = Wq_aug[:, :C]
        dY_out = dY_out.contiguous()
        d_delta, dgamma, dbeta = ops.xattn_tail_bwd(dY_out, delta_y, gamma, m_txt, B, T, C, kappa, thr, seed)
        # ---- folded out_proj + residual_head
        dWo_f = ops.linear_wgrad(d_delta, o)  # [C, d]
        db_r = ops.colsum(d_delta)  # d(bo_f) = d(b_r)
        do = ops.linear_dgrad(d_delta, Wo_f)
        dkv = new(B * T, 2 * d)
        d_in_w, d_in_b = torch.empty_like(in_w), new(3 * d)
        dY = ops.axpby(dY_out.view(B * T, C), 1.0 / (1.0 + kappa), new(B * T, C), False)
        if lowrank:
            z, dyh = ops.xattn_lowrank_bwd(do, Y2, kq, kv[:, d:], probs, m_txt, B, T, H, d, C, thr, seed, dkv[:, d:])
            dWq_aug = new(d, C1)
            for h in range(H):
                zh, ks = z[:, h * C1:(h + 1) * C1], slice(h * hd, (h + 1) * hd)
                ops.gemm(zh, Wq_aug[ks], dkv[:, ks], transB=True)  # dk = Z [Wq_f | b_q]^T
                ops.gemm(kv[:, ks], zh, dWq_aug[ks], transA=True)  # d[Wq_f | b_q] = k^T Z
                ops.axpby(dyh[h], 1.0, dY, True)
            dWq_f = dWq_aug[:, :C]
            d_in_b[:d].copy_(dWq_aug[:, C])
        else:
            dq = new(B * T, d)
            ops.xattn_core_bwd(do, q, kv[:, :d], kv[:, d:], probs, m_txt, B, T, H, d, thr, seed, dq, dkv[:, :d], dkv[:, d:], lo=lo)
            dWq_f = ops.linear_wgrad(dq, Y2)  # [d, C]
            ops.colsum(dq, out=d_in_b[:d])
            ops.linear_dgrad(dq, Wq_f, out=dY, beta=1.0)
        # ---- folded projections
        dWkv_f, dWkv_f_lo = new(2 * d, de), new(2 * d, ops.round_up(de, 4))
        ops.linear_wgrad(dkv, E2, out=dWkv_f, lo=lo, emit_lo=dWkv_f_lo)  # [2d, de]; its halves feed the un-fold products
        lo.put(dWkv_f[:d], dWkv_f_lo[:d])
        lo.put(dWkv_f[d:], dWkv_f_lo[d:])
        ops.colsum(dkv, out=d_in_b[d:])
        dE = ops.linear_dgrad(dkv, Wkv_f, lo=lo, emit_lo=True)  # handed to the TTF backward with its lo
        # ---- un-fold the weight gradients: W_f = W_a W_b  =>  dW_a = dW_f W_b^T, dW_b = W_a^T dW_f
        ops.gemm(dWq_f, W_Q, d_in_w[:d], transB=True)
        dW_Q = ops.gemm(in_w[:d], dWq_f, new(d, C), transA=True)
        dW_K, dW_V = new(d, de), new(d, de)
        ops.gemm_group([dict(A=dWkv_f[:d], B=W_K, C=d_in_w[d:2 * d], transB=True),
                        dict(A=in_w[d:2 * d], B=dWkv_f[:d], C=dW_K, transA=True),
                        dict(A=dWkv_f[d:], B=W_V, C=d_in_w[2 * d:], transB=True),
                        dict(A=in_w[2 * d:], B=dWkv_f[d:], C=dW_V, transA=True)], lo)  # the four un-folds, one launch
        dW_r = ops.gemm(dWo_f, out_w, new(C, d), transB=True)
        ops.gemm(db_r.view(C, 1), out_b.view(1, d), dW_r, beta=1.0)  # bo_f = W_r b_o + b_r also depends on W_r
        dW_o = ops.gemm(W_r, dWo_f, new(d, d), transA=True)
        db_o = ops.gemm(db_r.view(1, C), W_r, new(1, d)).view(d)
        return (dY.view(B, T, C), dE.view(B, T, de), None, None, None, None, None, None, None, dW_Q, dW_K, dW_V,
                d_in_w, d_in_b, dW_o, db_o, dW_r, db_r, dgamma, dbeta)


class XAttnRankWeightsFn(torch.autograd.Function):
    """Weight-space half of the rank-(2C+1) form of MMF_XAttn_Add (csrc/xattn_rank.cu): from the module's parameters to
    the skinny operand of the data pass,
        Wr = [A ; G] per head,  A_h = Wq_aug_h^T in_k_h W_K,  G_h = Wo_f_h in_v_h W_V,   br = [a0 ; g0],   bo_f = W_r b_o + b_r
    with Wq_aug = [in_q W_Q | b_q] and Wo_f = W_r W_o, optionally composed with the producer's deferred final projection
    (Wr_eff = Wr W_p, br_eff = Wr b_p + br).  Every product is one skinny immtsf_gemm; backward is the chain rule of each
    (C = A B  =>  dA = dC B^T, dB = A^T dC).  It depends on parameters only, so FusionModel runs it on a side stream next
    to the TTF forward, and autograd runs its backward on that stream next to the TTF backward."""

    @staticmethod
    def forward(ctx, H, C, W_Q, W_K, W_V, in_w, in_b, out_w, out_b, W_r, b_r, W_p, b_p, gamma=None, beta=None):
        """gamma, beta (the module's LayerNorm parameters) are handed through unchanged: their gradients then come back through
        THIS Function's backward, which runs on the side stream after the data-parallel all-reduce of the packed upstream
        gradients -- the data chain of the step never waits for that collective."""
        d, de = W_Q.shape[0], W_K.shape[1]
        hd, C1 = d // H, C + 1
        n1, nr = H * C1, H * (2 * C + 1)
        dev = W_Q.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        in_k, in_v, b_k, b_v = in_w[d:2 * d], in_w[2 * d:], in_b[d:2 * d], in_b[2 * d:]
        # Two independent chains of skinny products -- the key side (Wq_aug -> P1 -> A -> A W_p) and the value side (Wo_f -> P2 ->
        # G -> G W_p) -- on two lanes: depth 4 instead of one serial chain of ~16 launches (170 us on the cfg2 timeline).
        Wq_aug, Wo_f, bo_f = new(d, C1), new(C, d), new(C)
        Wr, br = new(nr, de), new(nr)
        P1, P2 = new(H, C1, d), new(H, C, d)
        fold = W_p is not None
        Wr_eff, br_eff = (new(nr, W_p.shape[1]), new(nr)) if fold else (Wr, br)
        fk = ops.Fork(dev, lanes=2, name="rankw_fwd")

        def fold_rows(rs):  # Wr_eff = Wr W_p, br_eff = Wr b_p + br, for a block of rows
            if fold:
                ops.gemm(Wr[rs], W_p, Wr_eff[rs])
                ops.gemm(Wr[rs], b_p.view(de, 1), br_eff[rs].view(-1, 1))  # one warp per row of Wr
                ops.axpby(br[rs], 1.0, br_eff[rs], True)

        def key_side():
            ops.multi_split([(in_b[:d].view(d, 1), Wq_aug[:, C:], None)])
            ops.gemm(in_w[:d], W_Q, Wq_aug[:, :C])
            for h in range(H):
                hs, ra = slice(h * hd, (h + 1) * hd), slice(h * C1, (h + 1) * C1)
                ops.gemm(Wq_aug[hs], in_k[hs], P1[h], transA=True)  # Wq_aug_h^T in_k_h            [C1, d]
                ops.gemm(P1[h], W_K, Wr[ra])  # A_h                                                  [C1, de]
                ops.gemm(b_k[hs].view(1, hd), Wq_aug[hs], br[ra].view(1, C1))  # a0_h^T = b_k_h^T Wq_aug_h
            fold_rows(slice(0, n1))

        def value_side():
            ops.gemm(W_r, out_w, Wo_f)
            ops.gemm(out_b.view(1, d), W_r, bo_f.view(1, C), transB=True, bias=b_r)
            for h in range(H):
                hs, rg = slice(h * hd, (h + 1) * hd), slice(n1 + h * C, n1 + (h + 1) * C)
                ops.gemm(Wo_f[:, hs], in_v[hs], P2[h])  # Wo_f_h in_v_h                              [C, d]
                ops.gemm(P2[h], W_V, Wr[rg])  # G_h                                                  [C, de]
                ops.gemm(Wo_f[:, hs], b_v[hs].view(hd, 1), br[rg].view(C, 1))  # g0_h
            fold_rows(slice(n1, nr))

        fk.run(key_side, lane=0)
        fk.run(value_side, lane=1)
        fk.join(Wq_aug, Wo_f, bo_f, Wr, br, P1, P2, Wr_eff, br_eff)
        ctx.H, ctx.C, ctx.dims, ctx.ln = H, C, (d, de), gamma is not None
        ctx.save_for_backward(W_Q, W_K, W_V, in_w, in_b, out_w, out_b, W_r, Wq_aug, Wo_f, P1, P2, Wr, W_p, b_p)
        if gamma is not None:
            return Wr_eff, br_eff, bo_f, gamma.view_as(gamma), beta.view_as(beta)
        return Wr_eff, br_eff, bo_f

    @staticmethod
    def backward(ctx, dWr_eff, dbr_eff, dbo_f, dgamma=None, dbeta=None):
        W_Q, W_K, W_V, in_w, in_b, out_w, out_b, W_r, Wq_aug, Wo_f, P1, P2, Wr, W_p, b_p = ctx.saved_tensors
        H, C = ctx.H, ctx.C
        d, de = ctx.dims
        hd, C1 = d // H, C + 1
        n1, nr = H * C1, H * (2 * C + 1)
        dev = W_Q.device
        new = lambda *s: torch.empty(*s, dtype=_f32, device=dev)
        in_k, in_v, b_k, b_v = in_w[d:2 * d], in_w[2 * d:], in_b[d:2 * d], in_b[2 * d:]
        dWr_eff, dbr, dbo_f = dWr_eff.contiguous(), dbr_eff.contiguous(), dbo_f.contiguous()
        if ops.DP_GROUP is not None:
            # data parallel: every gradient below is linear in these upstream tensors and the parameters are replicated, so
            # reducing THEM (H(2C+1) x (d+1) + 3C floats) gives all-reduced parameter gradients with no further traffic.
            # XAttnRankDataFn.backward wrote them into one buffer (ops.RANK_PACK): one collective, no gather copy.
            pack = ops.RANK_PACK
            ops.RANK_PACK = None
            if pack is not None and pack.data_ptr() == dWr_eff.data_ptr():
                pass  # already reduced on this (side) stream by XAttnRankDataFn.backward
            else:
                parts = [dWr_eff, dbr, dbo_f] + ([dgamma, dbeta] if dgamma is not None else [])
                pack = ops.dp_allreduce(torch.cat([t.reshape(-1) for t in parts]))
                outs, o = [], 0
                for t in parts:
                    outs.append(pack[o:o + t.numel()].view_as(t))
                    o += t.numel()
                dWr_eff, dbr, dbo_f = outs[:3]
                if dgamma is not None:
                    dgamma, dbeta = outs[3:]
        # the chain rule of the two forward chains, again on two lanes (key side / value side); a third lane takes the gradient of
        # the folded projection, which needs nothing from the chains
        dW_p = db_p = None
        fold = W_p is not None
        dWr = new(nr, de) if fold else dWr_eff
        d_in_w, d_in_b = torch.empty_like(in_w), new(3 * d)
        d_in_k, d_in_v, db_k, db_v = d_in_w[d:2 * d], d_in_w[2 * d:], d_in_b[d:2 * d], d_in_b[2 * d:]
        dWq_aug, dWo_f, dW_K, dW_V = new(d, C1), new(C, d), new(d, de), new(d, de)
        dW_Q, dW_r, dW_o, db_o = new(d, C), new(C, d), new(d, d), new(d)
        if fold:
            dW_p, db_p = new(de, W_p.shape[1]), new(de)
        fk = ops.Fork(dev, lanes=3, name="rankw_bwd")

        def unfold_rows(rs):  # Wr_eff = Wr W_p, br_eff = Wr b_p + br
            if fold:
                ops.gemm(dWr_eff[rs], W_p, dWr[rs], transB=True)
                ops.gemm(dbr[rs].view(-1, 1), b_p.view(1, de), dWr[rs], beta=1.0)

        def key_side():  # A_h = P1_h W_K ; P1_h = Wq_aug_h^T in_k_h ; a0_h = Wq_aug_h^T b_k_h ; Wq_aug = [in_q W_Q | b_q]
            unfold_rows(slice(0, n1))
            for h in range(H):
                hs, ra = slice(h * hd, (h + 1) * hd), slice(h * C1, (h + 1) * C1)
                dP1 = ops.gemm(dWr[ra], W_K, new(C1, d), transB=True)
                ops.gemm(in_k[hs], dP1, dWq_aug[hs], transB=True)
                ops.gemm(b_k[hs].view(hd, 1), dbr[ra].view(1, C1), dWq_aug[hs], beta=1.0)
                ops.gemm(P1[h], dWr[ra], dW_K, transA=True, beta=1.0 if h > 0 else 0.0)
                ops.gemm(Wq_aug[hs], dP1, d_in_k[hs])
                ops.gemm(Wq_aug[hs], dbr[ra].view(C1, 1), db_k[hs].view(hd, 1))
            dWq_f = dWq_aug[:, :C]
            ops.gemm(dWq_f, W_Q, d_in_w[:d], transB=True)
            ops.gemm(in_w[:d], dWq_f, dW_Q, transA=True)
            ops.multi_split([(dWq_aug[:, C:], d_in_b[:d].view(d, 1), None)])

        def value_side():  # G_h = P2_h W_V ; P2_h = Wo_f_h in_v_h ; g0_h = Wo_f_h b_v_h ; Wo_f = W_r W_o ; bo_f = W_r b_o + b_r
            unfold_rows(slice(n1, nr))
            for h in range(H):
                hs, rg = slice(h * hd, (h + 1) * hd), slice(n1 + h * C, n1 + (h + 1) * C)
                dP2 = ops.gemm(dWr[rg], W_V, new(C, d), transB=True)
                ops.gemm(dP2, in_v[hs], dWo_f[:, hs], transB=True)
                ops.gemm(dbr[rg].view(C, 1), b_v[hs].view(1, hd), dWo_f[:, hs], beta=1.0)
                ops.gemm(P2[h], dWr[rg], dW_V, transA=True, beta=1.0 if h > 0 else 0.0)
                ops.gemm(Wo_f[:, hs], dP2, d_in_v[hs], transA=True)
                ops.gemm(Wo_f[:, hs], dbr[rg].view(C, 1), db_v[hs].view(hd, 1), transA=True)
            ops.gemm(dWo_f, out_w, dW_r, transB=True)
            ops.gemm(dbo_f.view(C, 1), out_b.view(1, d), dW_r, beta=1.0)
            ops.gemm(W_r, dWo_f, dW_o, transA=True)
            ops.gemm(dbo_f.view(1, C), W_r, db_o.view(1, d))

        def folded_projection():
            ops.gemm(Wr, dWr_eff, dW_p, transA=True)
            ops.gemm(Wr, dbr.view(nr, 1), db_p.view(de, 1), transA=True)

        fk.run(key_side, lane=0)
        fk.run(value_side, lane=1)
        if fold:
            fk.run(folded_projection, lane=2)
        fk.join(d_in_w, d_in_b, dW_Q, dW_K, dW_V, dW_r, dW_o, db_o, dW_p, db_p)
        return None, None, dW_Q, dW_K, dW_V, d_in_w, d_in_b, dW_o, db_o, dW_r, dbo_f, dW_p, db_p, dgamma, dbeta


class XAttnRankDataFn(torch.autograd.Function):
    """Data half of the rank form: R = E Wr^T + br (the one skinny pass over the text-side tensor), the T x (2C+1)
    attention (immtsf_xattn_rank_*) and the LayerNorm / dropout / kappa-blend tail (MMF_XAttn_Add.py:83-102).  One launch
    per direction (+ a small ordered reduction of the weight-gradient partials) when csrc/xattn_rank_fused.cu applies
    (H (2C+1) <= 16, d <= 1024); otherwise the three-kernel forward / separate skinny products."""

    @staticmethod
    def forward(ctx, Y, E, m_txt, H, kappa, thr, seed, save, flags, d, Wr, br, bo_f, gamma, beta):
        B, T, C = Y.shape
        nr = H * (2 * C + 1)
        Y2 = Y.contiguous().view(B * T, C)
        E2 = E.contiguous().view(B * T, E.shape[2])
        Wr = Wr.contiguous()
        fused = ops.xattn_rank_fused_ok(T, H, d, C, E2, Wr)
        if fused:
            Y_out, R, delta_y, probs = ops.xattn_rank_fused_fwd(Y2, E2, Wr, br, bo_f, gamma, beta, m_txt, B, T, H, d, C, kappa, thr, seed,
                                                                 save, flags)
        else:
            nrp = ops.round_up(nr, 4)  # rows of R / dR are padded to 16 bytes (vector loads in the skinny kernels); pads unused
            if flags is not None:
                ops.nan_check(E2, flags, ops.FLAG_E)
            R = ops.gemm(E2, Wr, torch.empty(B * T, nrp, dtype=_f32, device=Y.device)[:, :nr], transB=True, bias=br)
            delta_y, probs = ops.xattn_rank_fwd(Y2, R, bo_f, m_txt, B, T, H, d, C, thr, seed, save)
            Y_out = ops.xattn_tail_fwd(Y2, delta_y, gamma, beta, m_txt, B, T, C, kappa, thr, seed, flags)
        if save:
            ctx.H, ctx.kappa, ctx.thr, ctx.seed, ctx.dims, ctx.fused = H, kappa, thr, seed, (B, T, C, d), fused
            ctx.save_for_backward(m_txt, Y2, E2, Wr, gamma, R, probs, delta_y)
        return Y_out

    @staticmethod
    def backward(ctx, dY_out):
        m_txt, Y2, E2, Wr, gamma, R, probs, delta_y = ctx.saved_tensors
        B, T, C, d = ctx.dims
        H, kappa, thr, seed = ctx.H, ctx.kappa, ctx.thr, ctx.seed
        nr, dk = Wr.shape
        dY_out = dY_out.contiguous()
        if ctx.fused:
            dE, dY, dWr, dbr, dbo_f, dgamma, dbeta, pack = ops.xattn_rank_fused_bwd(dY_out, delta_y, gamma, Y2, R, probs, m_txt, E2, Wr, B, T,
                                                                                    H, d, C, kappa, thr, seed)
            if ops.DP_GROUP is not None:
                # [dWr | dbr | d bo_f | dgamma | dbeta] are the sufficient statistics of every MMF gradient: all-reduced HERE (first in
                # the step's collective order) but on the SIDE stream, where XAttnRankWeightsFn.backward -- their only consumer, the
                # LayerNorm gradients included -- runs; the data chain does not wait for the collective
                cur = torch.cuda.current_stream()
                side = ops.side_stream(cur.device)
                side.wait_stream(cur)
                pack.record_stream(side)
                with torch.cuda.stream(side):
                    ops.dp_allreduce(pack)
                ops.RANK_PACK = pack
        else:
            d_delta, dgamma, dbeta = ops.xattn_tail_bwd(dY_out, delta_y, gamma, m_txt, B, T, C, kappa, thr, seed)
            dbo_f = ops.colsum(d_delta)  # = d(b_r)
            dR, dY = ops.xattn_rank_bwd(d_delta, Y2, R, probs, m_txt, B, T, H, d, C, thr, seed)
            ops.axpby(dY_out.view(B * T, C), 1.0 / (1.0 + kappa), dY, True)  # the blend passes Y straight through
            dWr = ops.gemm(dR, E2, torch.empty(nr, dk, dtype=_f32, device=dR.device), transA=True)
            dbr = ops.colsum(dR)
            dE = ops.gemm(dR, Wr, torch.empty(B * T, dk, dtype=_f32, device=dR.device))
        return dY.view(B, T, C), dE.view(B, T, dk), None, None, None, None, None, None, None, None, dWr, dbr, dbo_f, dgamma, dbeta
