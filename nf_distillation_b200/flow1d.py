"""1-D (tabular) FlowStep: host glue for the fused kernels in csrc/flow1d.cu (reference: is_1d branches of
models/flows.py:37-52,142-202 and models/layers.py:76,117,410-411)."""
from __future__ import annotations

import torch

from . import functional as Fn
from . import ops

F32 = torch.float32


def affine_rows(x4, Wf, bf, sl, y, ld_in, ld_out):
    """Stand-alone y = W x + b for [B, D] rows (or [B, C, 1, 1])."""
    B, C, H, W = x4.shape
    if H * W != 1:
        raise NotImplementedError("stand-alone 2-D ActNorm / InvertibleConv1x1 kernels exist for C in {12,24,48,96}")
    ops.affine_rows(x4, Wf, bf, sl, y, ld_in, ld_out, B, C, 1.0)


def _mlp_params(step):
    b = step.block
    ws = [b[i].weight for i in (0, 2, 4, 6, 8, 10)]
    bs = [b[i].bias for i in (0, 2, 4, 6, 8, 10)]
    return ws, bs


def _pack(step, reverse, with_bwd, an_bias, an_logs, inv, ws, bs):
    D, Cc, hid = step.in_channels, step.condition_features, step.hidden_channels
    Wf, bf, sl = Fn.build_affine(an_bias, an_logs, inv, D, reverse, True)
    tf, tb, tg, n_act, offs = ops.flow1d_sizes(D, Cc, hid)
    dev = Wf.device
    PF = torch.empty(tf, device=dev, dtype=F32)
    PB = torch.empty(tb, device=dev, dtype=F32) if with_bwd else None
    ops.flow1d_pack(Wf, bf, ws, bs, D, Cc, hid, PF, PB)
    return Wf, sl, PF, PB, (tg, n_act, offs)


def _consts(step, reverse):
    """No-grad path: packed weights cached until a parameter changes (frozen teacher, sampling)."""
    key = (reverse, tuple((p.data_ptr(), p._version) for p in step._all_params()))
    hit = step._cache.get(("1d", reverse))
    if hit is not None and hit[0] == key:
        return hit[1]
    ws, bs = _mlp_params(step)
    inv = tuple(None if t is None else t.detach() for t in step.invconv.lu_tensors())
    out = _pack(step, reverse, False, step.actnorm.bias.detach(), step.actnorm.logs.detach(), inv,
                [w.detach() for w in ws], [b.detach() for b in bs])
    step._cache[("1d", reverse)] = (key, out)
    return out


class FlowStep1dFn(torch.autograd.Function):
    """Differentiable 1-D FlowStep in either direction (the tabular configs train through the inverse pass:
    perceptual L1 on reverse-pass samples, conf/training/tabular.yaml:11-19)."""

    @staticmethod
    def forward(ctx, x, ld_in, cond, step, reverse, an_bias, an_logs, lower, upper, log_s, p, sign_s, *mlp):
        ws, bs = list(mlp[:6]), list(mlp[6:])
        D, Cc, hid = step.in_channels, step.condition_features, step.hidden_channels
        B = x.shape[0]
        x = x.contiguous()
        cond = None if cond is None else cond.contiguous()
        Wf, sl, PF, PB, (tg, n_act, offs) = _pack(step, reverse, True, an_bias, an_logs,
                                                  (lower, upper, log_s, p, sign_s, None), ws, bs)
        acts = torch.empty(B, n_act, device=x.device, dtype=F32)
        y = torch.empty_like(x)
        ld_out = torch.empty(B, device=x.device, dtype=F32)
        ops.flow1d_fwd(x, cond, PF, sl, y, ld_in.contiguous(), ld_out, acts, B, D, Cc, hid, reverse)
        ctx.meta = (D, Cc, hid, reverse, tg, offs, cond is not None)
        ctx.save_for_backward(x, acts, PF, PB, Wf, an_bias, an_logs, lower, upper, log_s, p, sign_s,
                              *([cond] if cond is not None else []), *ws, *bs)
        return y, ld_out

    @staticmethod
    def backward(ctx, g_y, g_ld):
        D, Cc, hid, reverse, tg, offs, has_cond = ctx.meta
        saved = ctx.saved_tensors
        x, acts, PF, PB, Wf, an_bias, an_logs, lower, upper, log_s, p, sign_s = saved[:12]
        rest = saved[12:]
        cond = rest[0] if has_cond else None
        ws = rest[1:7] if has_cond else rest[0:6]
        B = x.shape[0]
        dev = x.device
        g_y = torch.zeros_like(x) if g_y is None else g_y.contiguous()
        g_ld = torch.zeros(B, device=dev, dtype=F32) if g_ld is None else g_ld.contiguous()
        G = torch.zeros(tg, device=dev, dtype=F32)
        dx = torch.empty_like(x)
        ops.flow1d_bwd(x, cond, acts, PB, PF, g_y, g_ld, dx, G, B, D, Cc, hid, reverse)
        offG, offGB, ninp, _ = offs[0]
        dWf = G[offG:offG + D * ninp].view(D, ninp)[:, :D].contiguous()
        dbf = G[offGB:offGB + D].contiguous()
        d_bias, d_logs = torch.empty_like(an_bias), torch.empty_like(an_logs)
        d_lower, d_upper, d_log_s = torch.empty_like(lower), torch.empty_like(upper), torch.empty_like(log_s)
        ops.invconv_prep_bwd(an_bias, an_logs, lower, upper, log_s, p, sign_s, None, D, reverse, True, Wf, dWf, dbf,
                             g_ld, B, 1.0, d_bias, d_logs, d_lower, d_upper, d_log_s, None)
        dws, dbs = [], []
        for l in range(1, 7):
            offG, offGB, ninp, _ = offs[l]
            nout, nin = ws[l - 1].shape
            dws.append(G[offG:offG + nout * ninp].view(nout, ninp)[:, :nin])
            dbs.append(G[offGB:offGB + nout])
        return (dx, g_ld, None, None, None, d_bias, d_logs, d_lower, d_upper, d_log_s, None, None, *dws, *dbs)


def flowstep1d(step, input, y_onehot, logdet, reverse):
    from .models.layers import _as_logdet
    B = input.shape[0]
    ld = _as_logdet(logdet, B, input.device)
    want_ld = ld is not None
    if ld is None:
        ld = torch.zeros(B, device=input.device)
    cond = y_onehot if step.condition_features else None
    params = step._all_params()
    needs_grad = torch.is_grad_enabled() and (input.requires_grad or ld.requires_grad
                                              or any(p.requires_grad for p in params))
    if needs_grad:
        if not step.invconv.LU_decomposed:
            raise NotImplementedError("training with LU_decomposed=False is not built")
        iv = step.invconv
        ws, bs = _mlp_params(step)
        z, ld_out = FlowStep1dFn.apply(input, ld, cond, step, bool(reverse), step.actnorm.bias, step.actnorm.logs,
                                       iv.lower, iv.upper, iv.log_s, iv.p, iv.sign_s, *ws, *bs)
    else:
        Wf, sl, PF, _, _ = _consts(step, bool(reverse))
        x = input.contiguous()
        z = torch.empty_like(x)
        ld_out = torch.empty(B, device=x.device, dtype=F32)
        ops.flow1d_fwd(x, None if cond is None else cond.contiguous(), PF, sl, z, ld.contiguous(), ld_out, None, B,
                       step.in_channels, step.condition_features, step.hidden_channels, bool(reverse))
    return z, (ld_out if want_ld else None)
