"""torch.autograd.Function bindings over the C ABI (include/dvdgan_b200.h).

Every Function's forward/backward only enqueues this library's CUDA kernels on the current stream;
PyTorch supplies device memory (``torch.empty``), streams and the autograd tape.  No op has a CPU or
ATen fallback: a missing library or a CPU tensor raises.
"""
import ctypes
import os

import torch

from . import _C
from ._C import ConvDesc, call, ptr

F32 = torch.float32


def _new(shape, like):
    return torch.empty(shape, device=like.device, dtype=F32)


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _spatial(x):
    """(N, C, [D,] H, W) -> N, C, D, H, W"""
    if x.dim() == 4:
        n, c, h, w = x.shape
        return n, c, 1, h, w
    n, c, d, h, w = x.shape
    return n, c, d, h, w


def _ksize(w):
    """reference-layout weight (Co, Ci, [kD,] kH, kW) or Linear (Co, Ci) -> (kD, kH, kW)"""
    if w.dim() == 2:
        return 1, 1, 1
    if w.dim() == 4:
        return 1, w.shape[2], w.shape[3]
    return w.shape[2], w.shape[3], w.shape[4]


# --------------------------------------------------------------------------------------------------
# raw (non-autograd) helpers
# --------------------------------------------------------------------------------------------------

def sn_sigma(w_bar, u, v):
    """One power iteration (Normalization.py:19-31): updates u, v IN PLACE, returns sigma (1,)."""
    rows = w_bar.shape[0]
    cols = w_bar.numel() // rows
    sigma = _new((1,), w_bar)
    scratch = _new((rows + cols + 4,), w_bar)
    call("dvd_specnorm_fwd", ptr(w_bar), rows, cols, ptr(u), ptr(v), ptr(sigma), ptr(scratch))
    return sigma


def sn_backward(g_ref, w_bar, u, v, sigma):
    """dW_bar from G = dL/dW_sn (reference layout) using the CURRENT u, v (Q3 / Q17)."""
    rows = w_bar.shape[0]
    cols = w_bar.numel() // rows
    dw = _new(w_bar.shape, w_bar)
    scratch = torch.empty(2, device=w_bar.device, dtype=torch.float64)
    call("dvd_specnorm_bwd", ptr(g_ref), ptr(w_bar), ptr(u), ptr(v), ptr(sigma), rows, cols, ptr(dw), 0, ptr(scratch))
    return dw


def pack_weight(w, sigma=None, transpose=False):
    """reference layout (Co, Ci, k...) -> GEMM operand [taps][Ci][Co] (or flipped/transposed for dgrad)."""
    co, ci = w.shape[0], w.shape[1]
    taps = w.numel() // (co * ci)
    rows, ld = (co, ci) if transpose else (ci, co)
    dst = _new((taps, rows, ld), w)
    call("dvd_weight_pack", ptr(w), ci, taps, 0, co, 0, ci, ptr(sigma), 1 if transpose else 0, ptr(dst), rows, 0, ld, 0)
    return dst


def unpack_wgrad(dwp, like_w):
    co, ci = like_w.shape[0], like_w.shape[1]
    taps = like_w.numel() // (co * ci)
    g = _new(like_w.shape, like_w)
    call("dvd_weight_unpack", ptr(dwp), co, 0, ci, taps, 0, co, 0, ci, 0, ptr(g))
    return g


def _desc(N, Cin, Cout, D, H, W, k, *, in_relu=0, in_up=0, accumulate=0, out_act=0, res_up=0, x_kind=0):
    d = ConvDesc()
    d.x_kind = x_kind
    d.N1, d.N2, d.Cin, d.Cout = N, 1, Cin, Cout
    d.D, d.H, d.W = D, H, W
    d.kD, d.kH, d.kW = k
    hs, ws = H >> in_up, W >> in_up
    d.x_cs = D * hs * ws
    d.x_s1 = Cin * d.x_cs
    d.x_s2 = 0
    d.y_cs = D * H * W
    d.y_s1 = Cout * d.y_cs
    d.y_s2 = 0
    d.in_relu, d.in_up, d.accumulate, d.out_act, d.res_up = in_relu, in_up, accumulate, out_act, res_up
    d.r_cs = D * (H >> res_up) * (W >> res_up)
    d.r_s1 = Cout * d.r_cs
    d.r_s2 = 0
    return d


def conv_raw(x, wp, bias, Cout, k, *, in_relu=0, in_up=0, out_act=0, res=None, res_up=0, out=None, accumulate=0,
             x_kind=0):
    """x (N,Cin,[D,]Hs,Ws) contiguous -> y (N,Cout,[D,]H,W) with H = Hs << in_up."""
    N, Cin, D, Hs, Ws = _spatial(x)
    H, W = Hs << in_up, Ws << in_up
    d = _desc(N, Cin, Cout, D, H, W, k, in_relu=in_relu, in_up=in_up, accumulate=accumulate, out_act=out_act,
              res_up=res_up, x_kind=x_kind)
    if out is None:
        shape = (N, Cout, H, W) if x.dim() == 4 else (N, Cout, D, H, W)
        out = _new(shape, x)
    call("dvd_conv_fwd", ctypes.byref(d), ptr(x), ptr(wp), ptr(bias), ptr(res), ptr(out))
    return out


def wgrad_raw(x, dy, k, *, in_relu=0, in_up=0):
    """-> dwp [taps][Cin][Cout]"""
    N, Cin, D, Hs, Ws = _spatial(x)
    Cout = dy.shape[1]
    H, W = Hs << in_up, Ws << in_up
    d = _desc(N, Cin, Cout, D, H, W, k, in_relu=in_relu, in_up=in_up, x_kind=1)
    dwp = _new((k[0] * k[1] * k[2], Cin, Cout), x)
    call("dvd_conv_wgrad", ctypes.byref(d), ptr(x), ptr(dy), ptr(dwp))
    return dwp


def channel_sum(x, C):
    """x (N, C, P...) -> (C,) sum over N and P"""
    N = x.shape[0]
    Pn = x.numel() // (N * C)
    out = _new((C,), x)
    scratch = torch.empty(C, device=x.device, dtype=torch.float64)
    call("dvd_channel_sum", ptr(x), N, C, Pn, C * Pn, 0, ptr(out), ptr(scratch))
    return out


def gemm(a, b, *, ta=False, tb=False, bias=None, out=None, beta=0.0, alpha=1.0):
    """row-major 2-D GEMM: out = alpha * op(a) @ op(b) + beta * out (+ bias per column)"""
    M = a.shape[1] if ta else a.shape[0]
    K = a.shape[0] if ta else a.shape[1]
    N = b.shape[0] if tb else b.shape[1]
    if out is None:
        out = _new((M, N), a)
    call("dvd_bgemm", int(ta), int(tb), M, N, K, alpha, ptr(a), a.shape[1], 0, ptr(b), b.shape[1], 0, beta, ptr(out),
         N, 0, 1, ptr(bias))
    return out


def add(a, b):
    """a + b with the library's axpby kernel."""
    y = _new(a.shape, a)
    call("dvd_axpby", ptr(a), 1.0, 0.0, a.numel(), ptr(y))
    call("dvd_axpby", ptr(b), 1.0, 1.0, b.numel(), ptr(y))
    return y


# --------------------------------------------------------------------------------------------------
# autograd Functions
# --------------------------------------------------------------------------------------------------

class Fork(torch.autograd.Function):
    """x -> n aliases of x; the backward adds the branch gradients with this library's kernel instead of
    letting the autograd engine call ATen's add (residual fan-out: GResBlock.py:67-73, Discriminators.py:198-209)."""

    @staticmethod
    def forward(ctx, x, n):
        return tuple(x.view_as(x) for _ in range(n))

    @staticmethod
    def backward(ctx, *gs):
        gs = [_c(g) for g in gs if g is not None]
        if not gs:
            return None, None
        if len(gs) == 1:
            return gs[0], None
        acc = add(gs[0], gs[1])
        for g in gs[2:]:
            call("dvd_axpby", ptr(g), 1.0, 1.0, g.numel(), ptr(acc))
        return acc, None


def fork(x, n=2):
    """n aliases of x (plain aliases when no gradient is needed)."""
    if not (torch.is_grad_enabled() and x.requires_grad):
        return (x,) * n
    return Fork.apply(x, n)


def _conv_forward(x, w, bias, u, v, res, in_relu, in_up, out_act, res_up):
    """Shared forward of ConvFn / CBNConvFn: (optional) power iteration, weight packing with 1/sigma folded in, the
    conv itself.  Returns (y, sigma)."""
    sn = u is not None
    sigma = sn_sigma(w, u, v) if sn else None
    k = _ksize(w)
    wp = pack_weight(w, sigma)
    lin = x.dim() == 2
    xin = x.view(x.shape[0], x.shape[1], 1, 1) if lin else x
    rin = _c(res) if res is not None else None
    y = conv_raw(xin, wp, bias, w.shape[0], k, in_relu=in_relu, in_up=in_up, out_act=out_act, res=rin, res_up=res_up,
                 x_kind=1)
    if lin:
        y = y.view(y.shape[0], y.shape[1])
    return y, sigma


def _conv_backward(x, w, sigma, u, v, y, dy, cfg, need_x, need_w, need_b, need_res):
    """Shared backward: (dx, dw, db, dres) of a conv whose input was x (before the fused ReLU / upsample)."""
    in_relu, in_up, out_act, res_up, k, lin, has_bias, has_res = cfg
    dy = _c(dy)
    if out_act:
        d2 = _new(dy.shape, dy)
        call("dvd_act_bwd", ptr(y), ptr(dy), dy.numel(), out_act, ptr(d2))
        dy = d2
    xin = x.view(x.shape[0], x.shape[1], 1, 1) if lin else x
    dyin = dy.view(dy.shape[0], dy.shape[1], 1, 1) if lin else dy
    Co = w.shape[0]
    dx = dw = db = dres = None
    if need_w:
        g = unpack_wgrad(wgrad_raw(xin, dyin, k, in_relu=in_relu, in_up=in_up), w)
        dw = sn_backward(g, w, u, v, sigma) if sigma is not None else g
    if has_bias and need_b:
        db = channel_sum(dyin, Co)
    if need_x:
        wpt = pack_weight(w, sigma, transpose=True)
        if in_up:
            full = conv_raw(dyin, wpt, None, w.shape[1], k)
            N, Ci, D, H, W = _spatial(full)
            dx = _new(xin.shape, xin)
            call("dvd_avgpool_fwd", ptr(full), N * Ci * D, 1, H, W, 1, 2, 2, 4.0, 0, ptr(dx))
        else:
            dx = conv_raw(dyin, wpt, None, w.shape[1], k)
        if in_relu:
            call("dvd_act_bwd", ptr(xin), ptr(dx), dx.numel(), 1, ptr(dx))
        dx = dx.view(x.shape)
    if has_res and need_res:
        if res_up:
            N, C, D, H, W = _spatial(dyin)
            shape = (N, C, H // 2, W // 2) if dyin.dim() == 4 else (N, C, D, H // 2, W // 2)
            dres = _new(shape, dy)
            call("dvd_avgpool_fwd", ptr(dyin), N * C * D, 1, H, W, 1, 2, 2, 4.0, 0, ptr(dres))
        else:
            dres = dy
    return dx, dw, db, dres


class ConvFn(torch.autograd.Function):
    """Stride-1 same-padded conv (2-D/3-D) or Linear, optionally spectrally normalised, with fused
    input ReLU / nearest x2 upsample, bias, (upsampled) residual add and output activation.

    SN (Normalization.py:62-64): one power iteration updates u, v in place on every forward; the weight
    used is W_bar / sigma, folded into the operand packing; sigma is differentiable (Q3)."""

    @staticmethod
    def forward(ctx, x, w, bias, u, v, res, in_relu, in_up, out_act, res_up):
        _C.require_cuda(x, w)
        x = _c(x)
        y, sigma = _conv_forward(x, w, bias, u, v, res, in_relu, in_up, out_act, res_up)
        ctx.save_for_backward(x, w, sigma, u, v, y if out_act else None)
        ctx.cfg = (in_relu, in_up, out_act, res_up, _ksize(w), x.dim() == 2, bias is not None, res is not None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, sigma, u, v, y = ctx.saved_tensors
        dx, dw, db, dres = _conv_backward(x, w, sigma, u, v, y, dy, ctx.cfg, ctx.needs_input_grad[0],
                                          ctx.needs_input_grad[1], ctx.needs_input_grad[2], ctx.needs_input_grad[5])
        return dx, dw, db, None, None, dres, None, None, None, None


def conv(x, w, bias=None, u=None, v=None, res=None, in_relu=0, in_up=0, out_act=0, res_up=0):
    return ConvFn.apply(x, w, bias, u, v, res, in_relu, in_up, out_act, res_up)


class SNWeightFn(torch.autograd.Function):
    """W_bar / sigma as a tensor (for SpectralNorm around arbitrary modules, e.g. nn.Embedding)."""

    @staticmethod
    def forward(ctx, w, u, v):
        _C.require_cuda(w)
        sigma = sn_sigma(w, u, v)
        out = _new(w.shape, w)
        co = w.shape[0]
        cols = w.numel() // co
        # pack with taps = 1 in "transpose" orientation keeps the (rows, cols) layout and applies 1/sigma
        call("dvd_weight_pack", ptr(w), cols, 1, 0, co, 0, cols, ptr(sigma), 1, ptr(out), co, 0, cols, 0)
        ctx.save_for_backward(w, sigma, u, v)
        return out

    @staticmethod
    def backward(ctx, g):
        w, sigma, u, v = ctx.saved_tensors
        return sn_backward(_c(g), w, u, v, sigma), None, None


# Cross-replica BatchNorm statistics (opt-in; the reference leaves it as a TODO, Generator.py:57-58): with a process
# group set here, the per-channel sums of every ConditionalNorm are reduced over the ranks in the forward (sum, sum of
# squares) and in the backward (the two means of the batch-norm gradient), so that N data-parallel ranks normalise
# exactly like one process holding the global batch.  None = per-replica statistics, the reference's DataParallel.
SYNC_BN_GROUP = None


def set_sync_bn(group):
    """group: a torch.distributed process group (or True for the default group), None to switch it off."""
    global SYNC_BN_GROUP
    SYNC_BN_GROUP = group


def _sync_world():
    import torch.distributed as dist
    if SYNC_BN_GROUP is None or not (dist.is_available() and dist.is_initialized()):
        return None, 1
    g = None if SYNC_BN_GROUP is True else SYNC_BN_GROUP
    n = dist.get_world_size(g)
    return g, n


def _cbn_stats(x, running_mean, running_var, nbt, training, momentum, eps):
    N, C, H, W = x.shape
    mean, rstd = _new((C,), x), _new((C,), x)
    scratch = torch.empty(2 * C, device=x.device, dtype=torch.float64)
    group, world = _sync_world()
    if training and world > 1:
        import torch.distributed as dist
        args = (ptr(x), N, C, H * W, 1, momentum, eps, ptr(running_mean), ptr(running_var), ptr(nbt), ptr(mean),
                ptr(rstd), ptr(scratch))
        call("dvd_bn_stats_ex", *args, 1, 1)
        dist.all_reduce(scratch, op=dist.ReduceOp.SUM, group=group)
        call("dvd_bn_stats_ex", *args, 2, world)
        return mean, rstd
    call("dvd_bn_stats", ptr(x), N, C, H * W, int(training), momentum, eps, ptr(running_mean), ptr(running_var),
         ptr(nbt), ptr(mean), ptr(rstd), ptr(scratch))
    return mean, rstd


def _cbn_apply(x, gb, mean, rstd, relu, up):
    N, C, H, W = x.shape
    y = _new((N, C, H << up, W << up), x)
    call("dvd_cbn_apply", ptr(x), ptr(gb), gb.shape[0], ptr(mean), ptr(rstd), N, C, H, W, int(relu), up, ptr(y))
    return y


def _cbn_backward(x, gb, mean, rstd, dy, relu, up, training):
    N, C, H, W = x.shape
    dy = _c(dy)
    dx = _new(x.shape, x)
    dgb = _new(gb.shape, gb)
    scratch = _new((2 * C,), x)
    group, world = _sync_world()
    if training and world > 1:
        import torch.distributed as dist
        args = (ptr(x), ptr(gb), gb.shape[0], ptr(mean), ptr(rstd), ptr(dy), N, C, H, W, int(relu), up, 1, ptr(dx),
                ptr(dgb), ptr(scratch))
        call("dvd_cbn_bwd_ex", *args, 1)
        dist.all_reduce(scratch, op=dist.ReduceOp.SUM, group=group)
        call("dvd_axpby", ptr(scratch), 0.0, 1.0 / world, 2 * C, ptr(scratch))        # equal shards: mean of the means
        call("dvd_cbn_bwd_ex", *args, 2)
        return dx, dgb
    call("dvd_cbn_bwd", ptr(x), ptr(gb), gb.shape[0], ptr(mean), ptr(rstd), ptr(dy), N, C, H, W, int(relu), up,
         int(training), ptr(dx), ptr(dgb), ptr(scratch))
    return dx, dgb


class CBNFn(torch.autograd.Function):
    """BatchNorm2d(affine=False) + per-row (gamma|beta) affine + optional ReLU + optional nearest x2
    upsample (Normalization.py:78-88, GResBlock.py:49-55)."""

    @staticmethod
    def forward(ctx, x, gb, running_mean, running_var, nbt, relu, up, training, momentum, eps):
        _C.require_cuda(x, gb)
        x, gb = _c(x), _c(gb)
        mean, rstd = _cbn_stats(x, running_mean, running_var, nbt, training, momentum, eps)
        y = _cbn_apply(x, gb, mean, rstd, relu, up)
        ctx.save_for_backward(x, gb, mean, rstd)
        ctx.cfg = (relu, up, training)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gb, mean, rstd = ctx.saved_tensors
        relu, up, training = ctx.cfg
        dx, dgb = _cbn_backward(x, gb, mean, rstd, dy, relu, up, training)
        return dx, dgb, None, None, None, None, None, None, None, None


class CBNConvFn(torch.autograd.Function):
    """ConditionalNorm -> ReLU -> [nearest x2] -> SN-conv (+ residual) as ONE autograd node (GResBlock.py:49-57, 59-64).
    Same kernels as CBNFn followed by ConvFn; the difference is what is kept for the backward: only the block's
    pre-normalisation input x (and the per-channel statistics).  The normalised / rectified / upsampled activation --
    up to 4x the size of x and 61 % of what a GResBlock used to save -- is transient in the forward and recomputed by
    one more dvd_cbn_apply (an HBM-speed pass) at the start of the backward."""

    @staticmethod
    def forward(ctx, x, gb, running_mean, running_var, nbt, up, training, momentum, eps, w, bias, u, v, res, res_up):
        _C.require_cuda(x, gb, w)
        x, gb = _c(x), _c(gb)
        mean, rstd = _cbn_stats(x, running_mean, running_var, nbt, training, momentum, eps)
        a = _cbn_apply(x, gb, mean, rstd, True, up)
        y, sigma = _conv_forward(a, w, bias, u, v, res, 0, 0, 0, res_up)
        del a
        ctx.save_for_backward(x, gb, mean, rstd, w, sigma, u, v)
        ctx.cfg = (up, training, (0, 0, 0, res_up, _ksize(w), False, bias is not None, res is not None))
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gb, mean, rstd, w, sigma, u, v = ctx.saved_tensors
        up, training, ccfg = ctx.cfg
        a = _cbn_apply(x, gb, mean, rstd, True, up)                  # recompute the conv's input
        ni = ctx.needs_input_grad
        da, dw, db, dres = _conv_backward(a, w, sigma, u, v, None, dy, ccfg, ni[0] or ni[1], ni[9], ni[10], ni[13])
        del a
        dx = dgb = None
        if da is not None:
            dx, dgb = _cbn_backward(x, gb, mean, rstd, da, True, up, training)
        return dx, dgb, None, None, None, None, None, None, None, dw, db, None, None, dres, None


# ConvGRU BPTT state policy.  Full: keep the activated gates (3Ch) and r*h (Ch) of every frame next to h (Ch) -- 20
# bytes per hidden element, nothing recomputed.  Lean: keep h only (4 bytes per hidden element) and re-run the layer's
# forward sweep into transient buffers at the start of its backward (same kernels, same operands -> the same gate
# values); costs one extra forward of that layer per step.  The policy is per layer: GRU_LEAN is False (never), True
# (every layer) or a set of layer signatures (Cx, Ch, H, W, k) chosen by ``gru_lean_policy`` so that the layers that
# free the most bytes per recomputed FLOP go lean first -- which is what lets 128x128 clips at 32 per GPU fit 180 GB
# at +8 % step time instead of +25 %.
GRU_LEAN = False


def set_gru_lean(policy):
    """False / None: full state everywhere; True: lean everywhere; a set of (Cx, Ch, H, W, k): lean for those layers."""
    global GRU_LEAN
    GRU_LEAN = False if policy is None else (policy if isinstance(policy, bool) else frozenset(policy))


def gru_layers(ch, latent_dim):
    """(Cx, Ch, H, W, k) of the Generator's 12 ConvGRU layers in order (Generator.py:38-52)"""
    out = []
    for stage, side in enumerate((latent_dim, 2 * latent_dim, 4 * latent_dim, 8 * latent_dim)):
        hs, ks = ((4 * ch, 8 * ch, 4 * ch), (3, 5, 5)) if stage == 3 else ((8 * ch, 16 * ch, 8 * ch), (3, 5, 3))
        cx = hs[0]
        for h, k in zip(hs, ks):
            out.append((cx, h, side, side, k))
            cx = h
    return out


def gru_layer_state_bytes(B, T, ch, latent_dim):
    """full-state bytes (20 per hidden element) of those layers"""
    return [B * T * Ch * H * W * 20 for (_, Ch, H, W, _) in gru_layers(ch, latent_dim)]


def gru_state_bytes(B, T, ch, latent_dim, lean):
    """fp32 bytes of ConvGRU state the Generator keeps for the backward pass: all layers full or all lean"""
    return sum(gru_layer_state_bytes(B, T, ch, latent_dim)) // (5 if lean else 1)


def gru_lean_policy(B, T, ch, latent_dim, budget_bytes):
    """Which layers keep h only so that the kept ConvGRU state fits ``budget_bytes``: greedy by bytes freed (4/5 of the
    full state) per FLOP of the forward sweep that has to be repeated ((Cx + Ch) * Ch * k^2 per hidden-state pixel).
    -> (policy for set_gru_lean, bytes kept under it)."""
    layers = gru_layers(ch, latent_dim)
    sizes = gru_layer_state_bytes(B, T, ch, latent_dim)
    kept = sum(sizes)
    if kept <= budget_bytes:
        return False, kept
    order = sorted(range(len(layers)), key=lambda i: -(sizes[i] * 0.8) /
                   ((layers[i][0] + layers[i][1]) * layers[i][1] * layers[i][4] ** 2 * layers[i][2] * layers[i][3]))
    chosen = set()
    for i in order:
        chosen.add(layers[i])
        kept -= sizes[i] - sizes[i] // 5
        if kept <= budget_bytes:
            return chosen, kept
    return True, kept


class GRULayerFn(torch.autograd.Function):
    """One ConvGRU layer over a clip (ConvGRU.py:29-54 x Generator.py:87-97), BPTT in the backward.
    x: (B,T,Cx,H,W), or (B,Cx,H,W) fed to every frame (Generator.py:88-92, Q13) when T_bcast > 0."""

    @staticmethod
    def _run_fwd(x, h0, wu, wr, wo, bu, br, bo, cfg, h=None):
        B, T, Cx, Ch, H, W, k, x_bs, x_ts, _ = cfg
        gates = _new((B, T, 3 * Ch, H, W), x)
        if h is None:
            h = _new((B, T, Ch, H, W), x)
        rh = _new((B, T, Ch, H, W), x)
        nbytes = _C.lib().dvd_convgru_layer_workspace_bytes(B, T, Cx, Ch, H, W, k)
        ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
        call("dvd_convgru_layer_fwd", ptr(x), x_bs, x_ts, ptr(h0), ptr(wu), ptr(wr), ptr(wo), ptr(bu), ptr(br),
             ptr(bo), ptr(gates), ptr(h), ptr(rh), B, T, Cx, Ch, H, W, k, ptr(ws), nbytes)
        return gates, h, rh

    @staticmethod
    def forward(ctx, x, h0, wu, wr, wo, bu, br, bo, T_bcast):
        _C.require_cuda(x, wu)
        x = _c(x)
        if T_bcast:
            B, Cx, H, W = x.shape
            T = T_bcast
            x_bs, x_ts = Cx * H * W, 0
        else:
            B, T, Cx, H, W = x.shape
            x_bs, x_ts = T * Cx * H * W, Cx * H * W
        Ch, k = wu.shape[0], wu.shape[-1]
        if h0 is not None:
            h0 = _c(h0)
        cfg = (B, T, Cx, Ch, H, W, k, x_bs, x_ts, T_bcast)
        gates, h, rh = GRULayerFn._run_fwd(x, h0, wu, wr, wo, bu, br, bo, cfg)
        # (grad mode is always off inside Function.forward: do not test it here)
        ctx.lean = GRU_LEAN is True or (GRU_LEAN is not False and (Cx, Ch, H, W, k) in GRU_LEAN)
        if ctx.lean:
            ctx.save_for_backward(x, h0, wu, wr, wo, bu, br, bo, h)        # gates / rh are dropped here
        else:
            ctx.save_for_backward(x, h0, wu, wr, wo, gates, h, rh)
        ctx.cfg = cfg
        ctx.consumed = False
        return h

    @staticmethod
    def backward(ctx, dh):
        if ctx.consumed:
            raise RuntimeError("GRULayerFn: second backward through the same graph -- the saved gate buffer was "
                               "overwritten in place by the first one (the reference trainer backpropagates once)")
        ctx.consumed = True
        B, T, Cx, Ch, H, W, k, x_bs, x_ts, T_bcast = ctx.cfg
        if ctx.lean:
            x, h0, wu, wr, wo, bu, br, bo, h = ctx.saved_tensors
            # the recomputation rewrites the kept h with the values it already holds (same kernels, same operands)
            gates, _, rh = GRULayerFn._run_fwd(x, h0, wu, wr, wo, bu, br, bo, ctx.cfg, h=h)
        else:
            x, h0, wu, wr, wo, gates, h, rh = ctx.saved_tensors
        dh = _c(dh)
        dx = _new((B, T, Cx, H, W), x)
        dh0 = _new(h0.shape, x) if h0 is not None else None
        dwu, dwr, dwo = _new(wu.shape, x), _new(wr.shape, x), _new(wo.shape, x)
        dbu, dbr, dbo = _new((Ch,), x), _new((Ch,), x), _new((Ch,), x)
        nbytes = _C.lib().dvd_convgru_layer_workspace_bytes(B, T, Cx, Ch, H, W, k)
        ws = torch.empty(nbytes, device=x.device, dtype=torch.uint8)
        # NB: overwrites `gates` in place with the pre-activation gradients
        call("dvd_convgru_layer_bwd", ptr(x), x_bs, x_ts, ptr(h0), ptr(wu), ptr(wr), ptr(wo), ptr(gates), ptr(h),
             ptr(rh), ptr(dh), ptr(dx), ptr(dh0), ptr(dwu), ptr(dwr), ptr(dwo), ptr(dbu), ptr(dbr), ptr(dbo), B, T,
             Cx, Ch, H, W, k, ptr(ws), nbytes)
        del gates, rh
        if T_bcast:
            # sum over frames: dx_sum[b] = ones(1,T) @ dx[b] (T, Cx*H*W)
            ones = torch.ones(T, device=x.device, dtype=F32)
            dxs = _new((B, Cx, H, W), x)
            n = Cx * H * W
            call("dvd_bgemm", 0, 0, 1, n, T, 1.0, ptr(ones), T, 0, ptr(dx), n, T * n, 0.0, ptr(dxs), n, n, B, None)
            dx = dxs
        return dx, dh0, dwu, dwr, dwo, dbu, dbr, dbo, None


# ---- the layers of a ConvGRU stack as a wavefront (Generator.py:87-97: layer l+1 of frame t needs layer l of frame t only)
# GRULayerFn runs one layer over the whole clip before the next layer starts: a chain of 2T dependent GEMM launches that,
# on the small stages (4x4 / 8x8 / 16x16 frames, or few clips per GPU), are latency-bound -- a 200-k-block reduction on a
# handful of row tiles takes ~130 us however few rows there are, and most SMs idle.  GRUStackFn cuts the clip into chunks
# of frames and runs layer l on stream l: layer l works on chunk c while layer l-1 is already on chunk c+1, so up to
# n_layers independent chains are in flight (forward and BPTT).  Same kernels, same operands, same results.
# max_rows: only stages with at most this many output pixels per time step (B*H*W) -- above it one layer's GEMMs fill the
# machine for several waves, the batch chains of the library (option "gru_streams") fill the tails equally well
# (measured, profiles/r2) and the wavefront's extra transient memory (three layers' gradient planes at once) buys nothing
GRU_WAVEFRONT = {"enabled": 1, "chunk": 8, "max_rows": 32768}
for _kv in filter(None, os.environ.get("DVD_GRU_WAVEFRONT", "").split(",")):        # A/B runs: "enabled=0", "chunk=4"
    _k, _, _v = _kv.partition("=")
    GRU_WAVEFRONT[_k.strip()] = int(_v)
_WAVE_STREAMS = {}


def _wave_streams(dev, n):
    key = dev.index if dev.index is not None else torch.cuda.current_device()
    pool = _WAVE_STREAMS.setdefault(key, [])
    while len(pool) < n:
        pool.append(torch.cuda.Stream(device=dev))
    return pool[:n]


def _is_lean(sig):
    return GRU_LEAN is True or (GRU_LEAN is not False and sig in GRU_LEAN)


def gru_wavefront_chunk(B, T, sigs):
    """Frames per chunk for a stack of layers with signatures (Cx, Ch, H, W, k), or 0 when the stack runs layer by
    layer: one layer, one frame, a layer in lean mode (its recomputation is per layer), or a stage above ``max_rows``."""
    w = GRU_WAVEFRONT
    if not w["enabled"] or len(sigs) < 2 or T < 2 or any(_is_lean(sg) for sg in sigs):
        return 0
    if w["max_rows"] and B * sigs[0][2] * sigs[0][3] > w["max_rows"]:
        return 0
    return max(1, min(int(w["chunk"]), (T + 2) // 3))


class GRUStackFn(torch.autograd.Function):
    """ConvGRU.forward_sequence for a stack of layers with zero initial state: h of the LAST layer for every frame.
    apply(x, T_bcast, chunk, wu0, wr0, wo0, bu0, br0, bo0, wu1, ...)."""

    @staticmethod
    def _bounds(T, chunk):
        return [(t0, min(t0 + chunk, T)) for t0 in range(0, T, chunk)]

    @staticmethod
    def forward(ctx, x, T_bcast, chunk, *params):
        _C.require_cuda(x, *params)
        x = _c(x)
        if T_bcast:
            B, Cx, H, W = x.shape
            T = T_bcast
            x_bs, x_ts = Cx * H * W, 0
        else:
            B, T, Cx, H, W = x.shape
            x_bs, x_ts = T * Cx * H * W, Cx * H * W
        L = len(params) // 6
        lib = _C.lib()
        layers, inp = [], x
        for l in range(L):
            wu = params[6 * l]
            Ch, k = wu.shape[0], wu.shape[-1]
            nbytes = lib.dvd_convgru_layer_workspace_bytes(B, T, Cx, Ch, H, W, k)
            layers.append(dict(x=inp, x_bs=x_bs, x_ts=x_ts, Cx=Cx, Ch=Ch, k=k, gates=_new((B, T, 3 * Ch, H, W), x),
                               h=_new((B, T, Ch, H, W), x), rh=_new((B, T, Ch, H, W), x), nbytes=nbytes,
                               ws=torch.empty(nbytes, device=x.device, dtype=torch.uint8)))
            inp, Cx, x_bs, x_ts = layers[-1]["h"], Ch, T * Ch * H * W, Ch * H * W
        main = torch.cuda.current_stream(x.device)
        streams = _wave_streams(x.device, L)
        start = torch.cuda.Event()
        start.record(main)
        for st in streams:
            st.wait_event(start)
        prev = [None] * L          # event: layer l has finished the chunk being enqueued
        for t0, t1 in GRUStackFn._bounds(T, chunk):
            for l, ly in enumerate(layers):
                w = params[6 * l:6 * l + 6]
                with torch.cuda.stream(streams[l]):
                    if l > 0:
                        streams[l].wait_event(prev[l - 1])
                    call("dvd_convgru_layer_fwd_range", ptr(ly["x"]), ly["x_bs"], ly["x_ts"], None, ptr(w[0]), ptr(w[1]),
                         ptr(w[2]), ptr(w[3]), ptr(w[4]), ptr(w[5]), ptr(ly["gates"]), ptr(ly["h"]), ptr(ly["rh"]), B, T,
                         ly["Cx"], ly["Ch"], H, W, ly["k"], t0, t1, ptr(ly["ws"]), ly["nbytes"])
                    prev[l] = torch.cuda.Event()
                    prev[l].record(streams[l])
        for st in streams:          # the caller's stream continues after every layer (and only then are buffers released)
            e = torch.cuda.Event()
            e.record(st)
            main.wait_event(e)
        ctx.save_for_backward(x, *params, *[t for ly in layers for t in (ly["gates"], ly["h"], ly["rh"])])
        ctx.cfg = (B, T, H, W, T_bcast, chunk, [(ly["Cx"], ly["Ch"], ly["k"], ly["x_bs"], ly["x_ts"]) for ly in layers])
        ctx.consumed = False
        return layers[-1]["h"]

    @staticmethod
    def backward(ctx, dh_last):
        if ctx.consumed:
            raise RuntimeError("GRUStackFn: second backward through the same graph -- the saved gate buffers were "
                               "overwritten in place by the first one (the reference trainer backpropagates once)")
        ctx.consumed = True
        B, T, H, W, T_bcast, chunk, lcfg = ctx.cfg
        L = len(lcfg)
        saved = ctx.saved_tensors
        x, params, states = saved[0], saved[1:1 + 6 * L], saved[1 + 6 * L:]
        dh_last = _c(dh_last)
        lib = _C.lib()
        work = []
        for l, (Cx, Ch, k, x_bs, x_ts) in enumerate(lcfg):
            w = params[6 * l:6 * l + 6]
            nbytes = lib.dvd_convgru_layer_range_workspace_bytes(B, T, Cx, Ch, H, W, k)
            work.append(dict(dx=_new((B, T, Cx, H, W), x), dw=[_new(t.shape, x) for t in w], nbytes=nbytes,
                             ws=torch.empty(nbytes, device=x.device, dtype=torch.uint8)))
        main = torch.cuda.current_stream(x.device)
        streams = _wave_streams(x.device, L)
        start = torch.cuda.Event()
        start.record(main)
        for st in streams:
            st.wait_event(start)
        prev = [None] * L          # event: layer l has produced dx (= dh of layer l-1) of the chunk being enqueued
        for t0, t1 in reversed(GRUStackFn._bounds(T, chunk)):
            for l in reversed(range(L)):
                Cx, Ch, k, x_bs, x_ts = lcfg[l]
                w, wk = params[6 * l:6 * l + 6], work[l]
                gates, h, rh = states[3 * l:3 * l + 3]
                xin = x if l == 0 else states[3 * (l - 1) + 1]
                dh = dh_last if l == L - 1 else work[l + 1]["dx"]
                with torch.cuda.stream(streams[l]):
                    if l < L - 1:
                        streams[l].wait_event(prev[l + 1])
                    # NB: overwrites `gates` in place with the pre-activation gradients
                    call("dvd_convgru_layer_bwd_range", ptr(xin), x_bs, x_ts, None, ptr(w[0]), ptr(w[1]), ptr(w[2]),
                         ptr(gates), ptr(h), ptr(rh), ptr(dh), ptr(wk["dx"]), None, *[ptr(t) for t in wk["dw"]], B, T,
                         Cx, Ch, H, W, k, t0, t1, ptr(wk["ws"]), wk["nbytes"])
                    prev[l] = torch.cuda.Event()
                    prev[l].record(streams[l])
        for st in streams:
            e = torch.cuda.Event()
            e.record(st)
            main.wait_event(e)
        dx = work[0]["dx"]
        if T_bcast:          # sum over frames: dx_sum[b] = ones(1,T) @ dx[b] (T, Cx*H*W)
            Cx = lcfg[0][0]
            ones = torch.ones(T, device=x.device, dtype=F32)
            dxs = _new((B, Cx, H, W), x)
            n = Cx * H * W
            call("dvd_bgemm", 0, 0, 1, n, T, 1.0, ptr(ones), T, 0, ptr(dx), n, T * n, 0.0, ptr(dxs), n, n, B, None)
            dx = dxs
        return (dx, None, None) + tuple(g for wk in work for g in wk["dw"])


class AvgPoolFn(torch.autograd.Function):
    """F.avg_pool2d(x, 2) / F.avg_pool3d(x, 2) (window = stride)."""

    @staticmethod
    def forward(ctx, x, pd, ph, pw):
        _C.require_cuda(x)
        x = _c(x)
        N, C, D, H, W = _spatial(x)
        shape = (N, C, H // ph, W // pw) if x.dim() == 4 else (N, C, D // pd, H // ph, W // pw)
        y = _new(shape, x)
        call("dvd_avgpool_fwd", ptr(x), N * C, D, H, W, pd, ph, pw, 1.0, 0, ptr(y))
        ctx.cfg = (x.shape, N * C, D, H, W, pd, ph, pw)
        return y

    @staticmethod
    def backward(ctx, dy):
        shape, NC, D, H, W, pd, ph, pw = ctx.cfg
        dy = _c(dy)
        dx = _new(shape, dy)
        call("dvd_avgpool_bwd", ptr(dy), NC, D, H, W, pd, ph, pw, 0, ptr(dx))
        return dx, None, None, None


class MaxPoolFn(torch.autograd.Function):
    """nn.MaxPool3d with window = stride (Attention.py:48,138)."""

    @staticmethod
    def forward(ctx, x, pd, ph, pw):
        _C.require_cuda(x)
        x = _c(x)
        N, C, D, H, W = x.shape
        y = _new((N, C, D // pd, H // ph, W // pw), x)
        call("dvd_maxpool_fwd", ptr(x), N * C, D, H, W, pd, ph, pw, ptr(y))
        ctx.save_for_backward(x)
        ctx.cfg = (N * C, D, H, W, pd, ph, pw)
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        NC, D, H, W, pd, ph, pw = ctx.cfg
        dx = _new(x.shape, x)
        call("dvd_maxpool_bwd", ptr(x), ptr(_c(dy)), NC, D, H, W, pd, ph, pw, ptr(dx))
        return dx, None, None, None


class AttnCoreFn(torch.autograd.Function):
    """softmax(Q^T K) applied to V (Discriminators.py:108-114; Attention.py:92-101,165-176).
    q (B,dq,Nq) [or (B,Nq,dq) if q_token_major], k (B,dq,Nk), v (B,dv,Nk) -> (B,dv,Nq).

    Channel-major shapes the tensor-core kernel covers (dvd_attn_flash_supported) never materialise the (B,Nq,Nk) map:
    the forward keeps the log-sum-exp per query and the backward recomputes the probabilities from it.  The rest
    (SeparableAttnCell's raw-view token-major q, odd token counts) runs the materialised path."""

    @staticmethod
    def forward(ctx, q, k, v, q_token_major):
        _C.require_cuda(q, k, v)
        q, k, v = _c(q), _c(k), _c(v)
        B = q.shape[0]
        Nq, dq = (q.shape[1], q.shape[2]) if q_token_major else (q.shape[2], q.shape[1])
        dv, Nk = v.shape[1], v.shape[2]
        out = _new((B, dv, Nq), q)
        lib = _C.lib()
        ctx.cfg = (B, dq, dv, Nq, Nk, q_token_major)
        ctx.flash = (not q_token_major) and bool(lib.dvd_attn_flash_supported(B, dq, dv, Nq, Nk, 0))
        if ctx.flash:
            lse = _new((B, Nq), q)
            nbytes = lib.dvd_attn_flash_workspace_bytes(B, dq, dv, Nq, Nk, 0)
            ws = torch.empty(nbytes, device=q.device, dtype=torch.uint8)
            call("dvd_attn_flash_fwd", ptr(q), dq * Nq, ptr(k), dq * Nk, ptr(v), dv * Nk, ptr(out), dv * Nq, ptr(lse),
                 B, dq, dv, Nq, Nk, ptr(ws), nbytes)
            ctx.save_for_backward(q, k, v, out, lse)
            return out
        attn = _new((B, Nq, Nk), q)
        call("dvd_attn_fwd", ptr(q), dq * Nq, ptr(k), dq * Nk, ptr(v), dv * Nk, ptr(attn), ptr(out), dv * Nq, B, dq,
             dv, Nq, Nk, int(q_token_major))
        ctx.save_for_backward(q, k, v, attn)
        return out

    @staticmethod
    def backward(ctx, dout):
        B, dq, dv, Nq, Nk, qtm = ctx.cfg
        dout = _c(dout)
        lib = _C.lib()
        if ctx.flash and lib.dvd_attn_flash_supported(B, dq, dv, Nq, Nk, 1):
            q, k, v, out, lse = ctx.saved_tensors
            dq_, dk_, dv_ = _new(q.shape, q), _new(k.shape, q), _new(v.shape, q)
            nbytes = lib.dvd_attn_flash_workspace_bytes(B, dq, dv, Nq, Nk, 1)
            ws = torch.empty(nbytes, device=q.device, dtype=torch.uint8)
            call("dvd_attn_flash_bwd", ptr(q), dq * Nq, ptr(k), dq * Nk, ptr(v), dv * Nk, ptr(out), dv * Nq, ptr(dout),
                 dv * Nq, ptr(lse), ptr(dq_), dq * Nq, ptr(dk_), dq * Nk, ptr(dv_), dv * Nk, B, dq, dv, Nq, Nk,
                 ptr(ws), nbytes)
            return dq_, dk_, dv_, None
        if ctx.flash:       # forward on the tensor cores, backward shape outside their coverage (dv > 128): rebuild the map
            q, k, v, _, _ = ctx.saved_tensors
            attn = _new((B, Nq, Nk), q)
            scratch_out = _new((B, dv, Nq), q)
            call("dvd_attn_fwd", ptr(q), dq * Nq, ptr(k), dq * Nk, ptr(v), dv * Nk, ptr(attn), ptr(scratch_out), dv * Nq,
                 B, dq, dv, Nq, Nk, 0)
        else:
            q, k, v, attn = ctx.saved_tensors
        dattn = _new(attn.shape, attn)
        dq_, dk_, dv_ = _new(q.shape, q), _new(k.shape, q), _new(v.shape, q)
        call("dvd_attn_bwd", ptr(q), dq * Nq, ptr(k), dq * Nk, ptr(v), dv * Nk, ptr(attn), ptr(dattn), ptr(dout),
             dv * Nq, ptr(dq_), dq * Nq, ptr(dk_), dq * Nk, ptr(dv_), dv * Nk, B, dq, dv, Nq, Nk, int(qtm))
        return dq_, dk_, dv_, None


class ScaleResidualFn(torch.autograd.Function):
    """gamma * o + x (Discriminators.py:118, Attention.py:110,184)."""

    @staticmethod
    def forward(ctx, o, x, gamma):
        _C.require_cuda(o, x, gamma)
        o, x = _c(o), _c(x)
        y = _new(x.shape, x)
        call("dvd_scale_residual_fwd", ptr(o), ptr(x), ptr(gamma), x.numel(), ptr(y))
        ctx.save_for_backward(o, gamma)
        return y

    @staticmethod
    def backward(ctx, dy):
        o, gamma = ctx.saved_tensors
        dy = _c(dy)
        do = _new(o.shape, o)
        dg = _new(gamma.shape, o)
        scratch = torch.empty(1, device=o.device, dtype=torch.float64)
        call("dvd_scale_residual_bwd", ptr(o), ptr(dy), ptr(gamma), o.numel(), ptr(do), ptr(dg), ptr(scratch))
        return do, dy, dg


class Permute5Fn(torch.autograd.Function):
    """x.permute(perm).contiguous() for 5-D tensors."""

    @staticmethod
    def forward(ctx, x, perm):
        _C.require_cuda(x)
        x = _c(x)
        dims = (ctypes.c_int * 5)(*x.shape)
        pm = (ctypes.c_int * 5)(*perm)
        y = _new(tuple(x.shape[p] for p in perm), x)
        call("dvd_permute5", ptr(x), dims, pm, ptr(y))
        ctx.perm = perm
        return y

    @staticmethod
    def backward(ctx, dy):
        inv = [0] * 5
        for i, p in enumerate(ctx.perm):
            inv[p] = i
        dy = _c(dy)
        dims = (ctypes.c_int * 5)(*dy.shape)
        pm = (ctypes.c_int * 5)(*inv)
        dx = _new(tuple(dy.shape[p] for p in inv), dy)
        call("dvd_permute5", ptr(dy), dims, pm, ptr(dx))
        return dx, None


class EmbeddingFn(torch.autograd.Function):
    """nn.Embedding lookup (Generator.py:70)."""

    @staticmethod
    def forward(ctx, idx, w):
        _C.require_cuda(w)
        _C.require_index(idx)
        n, dim = idx.numel(), w.shape[1]
        y = _new((n, dim), w)
        call("dvd_embedding_fwd", ptr(w), ptr(idx), n, dim, w.shape[0], ptr(y))
        ctx.save_for_backward(idx)
        ctx.shape = w.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        dy = _c(dy)
        dw = torch.zeros(ctx.shape, device=dy.device, dtype=F32)
        call("dvd_embedding_bwd", ptr(dy), ptr(idx), idx.numel(), ctx.shape[1], ctx.shape[0], ptr(dw))
        return None, dw


class DHeadFn(torch.autograd.Function):
    """Projection head (Discriminators.py:264-291 / 421-447): ReLU, sum over HxW, SN-Linear + <feat, SN-Embedding>.
    Both spectral norms run their power iteration here; returns per-frame scores (N,)."""

    @staticmethod
    def forward(ctx, x, wl, bl, ul, vl, we, ue, ve, class_id, T):
        _C.require_cuda(x, wl, we)
        _C.require_index(class_id)
        x = _c(x)
        N, C = x.shape[0], x.shape[1]
        HW = x.numel() // (N * C)
        sl = sn_sigma(wl, ul, vl)
        se = sn_sigma(we, ue, ve)
        feat = _new((N, C), x)
        out = _new((N,), x)
        call("dvd_dhead_fwd", ptr(x), N, C, HW, T, we.shape[0], ptr(wl), ptr(sl), ptr(bl), ptr(we), ptr(se),
             ptr(class_id), ptr(feat), ptr(out))
        ctx.save_for_backward(x, feat, wl, sl, ul, vl, we, se, ue, ve, class_id)
        ctx.cfg = (N, C, HW, T)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, feat, wl, sl, ul, vl, we, se, ue, ve, class_id = ctx.saved_tensors
        N, C, HW, T = ctx.cfg
        dout = _c(dout)
        dx = _new(x.shape, x)
        gwl, db, gwe = _new(wl.shape, x), _new((1,), x), _new(we.shape, x)
        call("dvd_dhead_bwd", ptr(x), ptr(feat), ptr(dout), N, C, HW, T, we.shape[0], ptr(wl), ptr(sl), ptr(we),
             ptr(se), ptr(class_id), ptr(dx), ptr(gwl), ptr(db), ptr(gwe))
        dwl = sn_backward(gwl, wl, ul, vl, sl)
        dwe = sn_backward(gwe, we, ue, ve, se)
        return dx, dwl, db, None, None, dwe, None, None, None, None


class PhiFn(torch.autograd.Function):
    """vid_downsample (utils.py:77-83)."""

    @staticmethod
    def forward(ctx, x):
        _C.require_cuda(x)
        x = _c(x)
        B, T, C, H, W = x.shape
        y = _new((B, C, T, H // 2, W // 2), x)
        call("dvd_phi_fwd", ptr(x), B, T, C, H, W, ptr(y))
        ctx.shape = x.shape
        return y

    @staticmethod
    def backward(ctx, dy):
        B, T, C, H, W = ctx.shape
        dy = _c(dy)
        dx = _new(ctx.shape, dy)
        call("dvd_phi_bwd", ptr(dy), B, T, C, H, W, 0, ptr(dx))
        return dx


class GatherFramesFn(torch.autograd.Function):
    """data[:, idx] for sorted frame indices (utils.py:60-63)."""

    @staticmethod
    def forward(ctx, x, idx):
        _C.require_cuda(x)
        _C.require_index(idx)
        x = _c(x)
        B, T = x.shape[0], x.shape[1]
        k = idx.numel()
        fe = x.numel() // (B * T)
        y = _new((B, k) + tuple(x.shape[2:]), x)
        call("dvd_gather_frames_fwd", ptr(x), ptr(idx), B, T, k, fe, ptr(y))
        ctx.save_for_backward(idx)
        ctx.cfg = (x.shape, B, T, k, fe)
        return y

    @staticmethod
    def backward(ctx, dy):
        (idx,) = ctx.saved_tensors
        shape, B, T, k, fe = ctx.cfg
        dy = _c(dy)
        dx = _new(shape, dy)
        call("dvd_gather_frames_bwd", ptr(dy), ptr(idx), B, T, k, fe, 0, ptr(dx))
        return dx, None


class GanLossFn(torch.autograd.Function):
    """trainer.py:114-121 for one or two score vectors: sum_i mean(relu(1 + sign_i * x_i)) (hinge) or
    sum_i mean(sign_i * x_i) (wgan-gp without its penalty, Q6)."""

    @staticmethod
    def forward(ctx, hinge, sign_a, xa, sign_b, xb):
        _C.require_cuda(xa)
        xa = _c(xa)
        loss = _new((1,), xa)
        call("dvd_gan_loss_fwd", ptr(xa), xa.numel(), float(sign_a), int(hinge), 0, ptr(loss))
        if xb is not None:
            xb = _c(xb)
            call("dvd_gan_loss_fwd", ptr(xb), xb.numel(), float(sign_b), int(hinge), 1, ptr(loss))
        ctx.save_for_backward(xa, xb)
        ctx.cfg = (hinge, sign_a, sign_b)
        return loss.view(())

    @staticmethod
    def backward(ctx, g):
        xa, xb = ctx.saved_tensors
        hinge, sa, sb = ctx.cfg
        g = _c(g).view(1)
        da = _new(xa.shape, xa)
        call("dvd_gan_loss_bwd", ptr(xa), ptr(g), xa.numel(), float(sa), int(hinge), ptr(da))
        db = None
        if xb is not None:
            db = _new(xb.shape, xb)
            call("dvd_gan_loss_bwd", ptr(xb), ptr(g), xb.numel(), float(sb), int(hinge), ptr(db))
        return None, None, da, None, db


class ActFn(torch.autograd.Function):
    """relu (1) / tanh (2) as stand-alone ops."""

    @staticmethod
    def forward(ctx, x, act):
        _C.require_cuda(x)
        x = _c(x)
        y = _new(x.shape, x)
        call("dvd_act_fwd", ptr(x), x.numel(), act, ptr(y))
        ctx.save_for_backward(y)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (y,) = ctx.saved_tensors
        dy = _c(dy)
        dx = _new(y.shape, y)
        call("dvd_act_bwd", ptr(y), ptr(dy), y.numel(), ctx.act, ptr(dx))
        return dx, None


class AddFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b):
        _C.require_cuda(a, b)
        return add(_c(a), _c(b))

    @staticmethod
    def backward(ctx, g):
        return g, g


def adam_step(p, g, m, v, lr, beta1, beta2, eps, step, grad_scale=1.0):
    """Fused Adam over flat fp32 arenas (trainer.py:136-141)."""
    _C.require_cuda(p, g, m, v)
    call("dvd_adam_step", ptr(p), ptr(g), ptr(m), ptr(v), p.numel(), lr, beta1, beta2, eps, step, grad_scale)
