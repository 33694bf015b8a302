"""Fused operators beyond the reference's operator packages: what the completion models do
with torch glue around the hot path, as single exact CUDA searches.  Opt-in — see model_patches.py."""
import torch

from . import _lib


def knn_points(k, cloud, queries=None):
    """k nearest points of `cloud` (B, M, 3) for every point of `queries` (B, N, 3; default: the cloud itself,
    so every point's first neighbour is itself).  Returns (dist2 (B, N, k) float32 squared distances,
    idx (B, N, k) int32), ascending in (distance, index).  Not differentiable (indices; recompute distances
    from the gathered points where a gradient is needed — model_patches.knn_point does).
    Replaces the (B, N, M) matmul + torch.topk of completion/model_utils.py:242-259."""
    if queries is None:
        queries = cloud
    cloud = cloud.detach().contiguous()
    queries = queries.detach().contiguous()
    dev = _lib.require_cuda(queries, cloud, dtype=torch.float32, what="knn_points")
    if cloud.dim() != 3 or queries.dim() != 3 or cloud.size(2) != 3 or queries.size(2) != 3:
        raise ValueError(f"knn_points: expected (B, N, 3) tensors, got {tuple(queries.shape)} and {tuple(cloud.shape)}")
    B, N, _ = queries.shape
    M = cloud.size(1)
    if cloud.size(0) != B:
        raise ValueError("knn_points: batch sizes differ")
    k = int(k)
    if not 1 <= k <= min(M, 64):
        raise ValueError(f"knn_points: k={k} out of range 1..min(M, 64) (M={M})")
    dist2 = torch.empty(B, N, k, device=dev, dtype=torch.float32)
    idx = torch.empty(B, N, k, device=dev, dtype=torch.int32)
    with torch.cuda.device(dev):
        ws = _lib.workspace(_lib.lib.mvp_knn_points_workspace_bytes(B, N, M, k), dev)
        rc = _lib.lib.mvp_knn_points(B, N, M, k, _lib.ptr(queries), _lib.ptr(cloud), _lib.ptr(dist2), _lib.ptr(idx),
                                     _lib.ptr(ws), ws.numel(), _lib.stream_of(queries))
    _lib.check(rc, "mvp_knn_points")
    return dist2, idx


class _ChamferLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, dist1, dist2):
        dist1, dist2 = dist1.contiguous(), dist2.contiguous()
        dev = _lib.require_cuda(dist1, dist2, dtype=torch.float32, what="chamfer_loss")
        if dist1.dim() != 2 or dist2.dim() != 2 or dist1.size(0) != dist2.size(0):
            raise ValueError(f"chamfer_loss: expected (B, N) and (B, M), got {tuple(dist1.shape)} and {tuple(dist2.shape)}")
        B, n = dist1.shape
        m = dist2.size(1)
        cd_p = torch.empty(B, device=dev, dtype=torch.float32)
        cd_t = torch.empty(B, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            rc = _lib.lib.mvp_chamfer_loss(B, n, m, _lib.ptr(dist1), _lib.ptr(dist2), _lib.ptr(cd_p), _lib.ptr(cd_t),
                                           _lib.stream_of(dist1))
        _lib.check(rc, "mvp_chamfer_loss")
        ctx.save_for_backward(dist1, dist2)
        return cd_p, cd_t

    @staticmethod
    def backward(ctx, g_p, g_t):
        dist1, dist2 = ctx.saved_tensors
        B, n = dist1.shape
        m = dist2.size(1)
        g_p = (torch.zeros(B, device=dist1.device) if g_p is None else g_p).contiguous().float()
        g_t = (torch.zeros(B, device=dist1.device) if g_t is None else g_t).contiguous().float()
        gd1, gd2 = torch.empty_like(dist1), torch.empty_like(dist2)
        with torch.cuda.device(dist1.device):
            rc = _lib.lib.mvp_chamfer_loss_grad(B, n, m, _lib.ptr(dist1), _lib.ptr(dist2), _lib.ptr(g_p), _lib.ptr(g_t),
                                                _lib.ptr(gd1), _lib.ptr(gd2), _lib.stream_of(dist1))
        _lib.check(rc, "mvp_chamfer_loss_grad")
        return gd1, gd2


def chamfer_loss(dist1, dist2):
    """(cd_p, cd_t) per cloud from the Chamfer operator's dist1 (B, N) and dist2 (B, M) — the torch glue of
    completion/model_utils.py:71-72 as one reduction kernel (and one elementwise kernel backward):
    cd_p = (sqrt(dist1).mean(1) + sqrt(dist2).mean(1)) / 2,  cd_t = dist1.mean(1) + dist2.mean(1).  Differentiable."""
    return _ChamferLoss.apply(dist1, dist2)


def emd_loss(dist):
    """calc_emd's epilogue (completion/model_utils.py:84: `torch.sqrt(dist).mean(1)`) on the EMD operator's dist (B, N):
    the Chamfer epilogue's cd_p half with both arguments equal ((m + m) / 2 == m exactly).  Differentiable."""
    return _ChamferLoss.apply(dist, dist)[0]


def fscore(dist1, dist2, threshold=0.0001):
    """(fscore, precision_1, precision_2) per cloud — utils/metrics/CD/fscore.py:12-15 as ONE launch (mvp_fscore: a CTA
    per cloud counts both directions; torch's arithmetic, NaN -> 0).  Not differentiable (comparisons)."""
    d1, d2 = dist1.detach().contiguous(), dist2.detach().contiguous()
    dev = _lib.require_cuda(d1, d2, dtype=torch.float32, what="fscore")
    B, n = d1.shape
    m = d2.shape[1]
    f, p1, p2 = (torch.empty(B, device=dev, dtype=torch.float32) for _ in range(3))
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.mvp_fscore(B, n, m, _lib.ptr(d1), _lib.ptr(d2), float(threshold), _lib.ptr(f), _lib.ptr(p1),
                                       _lib.ptr(p2), _lib.stream_of(d1)), "mvp_fscore")
    return f, p1, p2


def three_nn_weights(target, source):
    """The three nearest points of `source` (B, M, 3) for every point of `target` (B, N, 3) and their normalised
    inverse-distance weights — completion/model_utils.py:286-293 (three_nn_upsampling: three_nn, clamp, reciprocal,
    sum, divide) as the grid search plus ONE elementwise kernel.  Returns (idx (B, N, 3) int32, weight (B, N, 3));
    not differentiable (the original's weights are not either: three_nn marks its outputs non-differentiable)."""
    target, source = target.detach().contiguous(), source.detach().contiguous()
    dev = _lib.require_cuda(target, source, dtype=torch.float32, what="three_nn_weights")
    B, N, _ = target.shape
    m = source.size(1)
    dist2 = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
    idx = torch.empty(B, N, 3, device=dev, dtype=torch.int32)
    weight = torch.empty(B, N, 3, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        ws = _lib.workspace(_lib.lib.mvp_three_nn_workspace_bytes(B, N, m), dev)
        # the kernel that finds a target's neighbours also writes its weights: no second pass over the distances
        _lib.check(_lib.lib.mvp_three_nn_weights_ws(B, N, m, _lib.ptr(target), _lib.ptr(source), _lib.ptr(dist2), _lib.ptr(idx),
                                                    _lib.ptr(weight), _lib.ptr(ws), ws.numel(), _lib.stream_of(target)),
                   "mvp_three_nn_weights_ws")
    return idx, weight


class _FpsGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, points, m, channels_first):
        pts = points.contiguous()
        dev = _lib.require_cuda(pts, dtype=torch.float32, what="fps_gather")
        B, N, _ = pts.shape
        idx = torch.empty(B, m, device=dev, dtype=torch.int32)
        out = torch.empty((B, 3, m) if channels_first else (B, m, 3), device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.mvp_furthest_point_sampling_gather(B, N, m, _lib.ptr(pts), None, _lib.ptr(idx), _lib.ptr(out),
                                                                   1 if channels_first else 0, _lib.stream_of(pts)),
                       "mvp_furthest_point_sampling_gather")
        ctx.mark_non_differentiable(idx)
        ctx.save_for_backward(idx)
        ctx.shape, ctx.cf = pts.shape, channels_first
        return idx, out

    @staticmethod
    def backward(ctx, _g_idx, g_out):
        (idx,) = ctx.saved_tensors
        if g_out is None:
            return None, None, None
        g = g_out.transpose(1, 2) if ctx.cf else g_out  # (B, m, 3)
        grad = torch.zeros(ctx.shape, device=g_out.device, dtype=g_out.dtype)
        grad.scatter_add_(1, idx.long().unsqueeze(-1).expand(-1, -1, 3), g.contiguous())  # a point may be picked twice
        return grad, None, None


def fps_gather(points, m, channels_first=False):
    """furthest_point_sample(points, m) and the sampled points themselves from ONE launch — the sampling idiom of the
    completion models (completion/model_utils.py:91-93 and :209-210: FPS, transpose, gather_points, transpose back;
    completion/models/vrcnet.py:451 with channels_first=True).  points (B, N, 3) -> (idx (B, m) int32,
    sampled (B, m, 3) or (B, 3, m)).  Same indices and coordinates bit for bit; differentiable in `points`."""
    return _FpsGather.apply(points, int(m), bool(channels_first))


def ball_query_group(min_radius, max_radius, nsample, xyz, new_xyz):
    """ball_query and the neighbours' coordinates from ONE launch (completion/model_utils.py:211-214: ball_query,
    grouping_operation on the transposed cloud, permute(0, 2, 3, 1).contiguous()).  xyz (B, N, 3), new_xyz (B, P, 3) ->
    (idx (B, P, nsample) int32, grouped (B, P, nsample, 3)).  Differentiable in `xyz` (get_uniform_loss differentiates
    through the grouped points): the gradient is scattered back by index, as grouping_operation's backward does."""
    return _BallQueryGroup.apply(float(min_radius), float(max_radius), int(nsample), xyz, new_xyz)


class _BallQueryGroup(torch.autograd.Function):
    @staticmethod
    def forward(ctx, min_radius, max_radius, nsample, xyz, new_xyz):
        x, c = xyz.contiguous(), new_xyz.detach().contiguous()
        dev = _lib.require_cuda(x, c, dtype=torch.float32, what="ball_query_group")
        B, N, _ = x.shape
        P = c.size(1)
        idx = torch.empty(B, P, nsample, device=dev, dtype=torch.int32)
        grouped = torch.empty(B, P, nsample, 3, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.mvp_ball_query_group(B, N, P, min_radius, max_radius, nsample, _lib.ptr(c), _lib.ptr(x),
                                                     _lib.ptr(idx), _lib.ptr(grouped), _lib.stream_of(x)),
                       "mvp_ball_query_group")
        ctx.mark_non_differentiable(idx)
        ctx.save_for_backward(idx)
        ctx.shape = x.shape
        return idx, grouped

    @staticmethod
    def backward(ctx, _g_idx, g_grouped):
        (idx,) = ctx.saved_tensors
        if g_grouped is None:
            return None, None, None, None, None
        B, N, _ = ctx.shape
        grad = torch.zeros(ctx.shape, device=g_grouped.device, dtype=g_grouped.dtype)
        grad.scatter_add_(1, idx.long().view(B, -1, 1).expand(-1, -1, 3), g_grouped.reshape(B, -1, 3))
        return None, None, None, grad, None


def _pointwise_conv_raw(x3, w2, bias, relu=False, mask=None, out_shape=None):
    """x3 (B, C, N), w2 (O, C), bias (O) or None -> (B, O, N) (or out_shape, same elements) through
    mvp_pointwise_conv(_masked); all contiguous fp32."""
    dev = _lib.require_cuda(x3, w2, dtype=torch.float32, what="pointwise_conv")
    B, C, N = x3.shape
    O = w2.shape[0]
    y = torch.empty(out_shape if out_shape is not None else (B, O, N), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        if mask is None:
            _lib.check(_lib.lib.mvp_pointwise_conv(B, C, O, N, _lib.ptr(x3), _lib.ptr(w2),
                                                   _lib.ptr(bias) if bias is not None else None, 1 if relu else 0,
                                                   _lib.ptr(y), _lib.stream_of(x3)), "mvp_pointwise_conv")
        else:
            _lib.check(_lib.lib.mvp_pointwise_conv_masked(B, C, O, N, _lib.ptr(x3), _lib.ptr(mask), _lib.ptr(w2),
                                                          _lib.ptr(y), _lib.stream_of(x3)), "mvp_pointwise_conv_masked")
    return y


def _pointwise_wgrad_raw(g3, x3, with_bias):
    """g3 (B, O, N), x3 (B, C, N) -> (grad_w (O, C), grad_bias (O) or None) through mvp_pointwise_wgrad."""
    dev = _lib.require_cuda(g3, x3, dtype=torch.float32, what="pointwise_wgrad")
    B, O, N = g3.shape
    C = x3.shape[1]
    gw = torch.empty(O, C, device=dev, dtype=torch.float32)
    gb = torch.empty(O, device=dev, dtype=torch.float32) if with_bias else None
    nbytes = int(_lib.lib.mvp_pointwise_wgrad_workspace_bytes(C, O, 1 if with_bias else 0))
    ws = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.mvp_pointwise_wgrad(B, C, O, N, _lib.ptr(g3), _lib.ptr(x3), _lib.ptr(gw),
                                                _lib.ptr(gb) if gb is not None else None, _lib.ptr(ws), nbytes,
                                                _lib.stream_of(g3)), "mvp_pointwise_wgrad")
    return gw, gb


kWgradMinIn, kWgradMinOut = 96, 128  # layers whose weight gradient goes through mvp_pointwise_wgrad ...
kWgradMinPositions = 400000          # ... or any layer over at least this many positions (B x N)
kPointwiseBmmWgrad = 8192  # in x out channels up to which the weight gradient is a batched fp32 matmul (see model_patches)


class _PointwiseConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, relu):
        shp = x.shape
        B, C = shp[0], shp[1]
        N = x.numel() // max(1, B * C) if B * C else 0
        x3 = x.reshape(B, C, N).contiguous()
        w2 = weight.reshape(weight.shape[0], -1).contiguous()
        if w2.shape[1] != C:
            raise _lib.MvpOpsError("pointwise_conv: weight must be (out_channels, in_channels[, 1[, 1]])")
        b1 = bias.contiguous() if bias is not None else None
        y = _pointwise_conv_raw(x3, w2, b1, relu, out_shape=(B, w2.shape[0], *shp[2:]))
        ctx.save_for_backward(x3, w2, y if relu else None)
        ctx.meta = (shp, weight.shape, bias is not None)
        return y

    @staticmethod
    def backward(ctx, g):
        x3, w2, y = ctx.saved_tensors
        shp, wshape, has_bias = ctx.meta
        B, C, N = x3.shape
        O = w2.shape[0]
        g3 = g.reshape(B, O, N).contiguous()
        y3 = y.view(B, O, N) if y is not None else None
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            # the input gradient is the same contraction with the weight transposed (and, behind a fused ReLU, with the
            # gradient masked by the sign of the output as it is staged)
            gx = _pointwise_conv_raw(g3, w2.t().contiguous(), None, False, mask=y3, out_shape=shp)
        if y3 is not None and (ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2])):
            g3 = g3 * (y3 > 0)
        want_w, want_b = ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        if want_w and C + (1 if want_b else 0) <= 256 and B * N > 0 and (
                (kWgradMinIn <= C and O >= kWgradMinOut) or B * N >= kWgradMinPositions):
            # both gradients from one pass over g and x (tcgen05); measured to pay for 128 -> 256 channels (0.098 ms against
            # cuDNN's 0.144 + the bias sum) and for layers over very many positions (ECG's 48 -> 24 over 32 x 49 152: the
            # batched fp32 matmul takes 1.2 ms), not for 64 -> 256 over 64 x 3072 (0.116 against 0.053 + 0.044)
            gw, gb = _pointwise_wgrad_raw(g3, x3, want_b)
            gw = gw.reshape(wshape) if want_w else None
            return gx, gw, gb, None
        if want_w:
            if O * C <= kPointwiseBmmWgrad:
                gw = torch.bmm(g3, x3.transpose(1, 2)).sum(0)
            else:  # a library TF32 weight gradient (cuDNN), as the layer's own backward would run
                gw = torch.ops.aten.convolution_backward(g3.unsqueeze(-1), x3.unsqueeze(-1), w2.view(O, C, 1, 1), None,
                                                         [1, 1], [0, 0], [1, 1], False, [0, 0], 1,
                                                         [False, True, False])[1]
            gw = gw.reshape(wshape)
        if want_b:
            gb = channel_sum(g3)
        return gx, gw, gb, None


def bias_add_(y, bias, relu=False):
    """y (B, C, ...) += bias (C) over the channel axis IN PLACE (then ReLU if asked), through mvp_bias_add.  No autograd:
    a building block of conv_bias / of a caller's own Function."""
    dev = _lib.require_cuda(y, bias, dtype=torch.float32, what="bias_add_")
    if not y.is_contiguous() or not bias.is_contiguous():
        raise _lib.MvpOpsError("bias_add_: contiguous tensors expected")
    B, C = y.shape[0], y.shape[1]
    if bias.numel() != C:
        raise _lib.MvpOpsError("bias_add_: bias must have one entry per channel")
    N = y.numel() // max(1, B * C)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.mvp_bias_add(B, C, N, _lib.ptr(y), _lib.ptr(bias), 1 if relu else 0, _lib.stream_of(y)),
                   "mvp_bias_add")
    return y


def channel_sum(g):
    """g (B, C, ...) -> (C,): the sum over clouds and points of every channel (a bias gradient), deterministic, through
    mvp_channel_sum.  No autograd."""
    g = g.contiguous()
    dev = _lib.require_cuda(g, dtype=torch.float32, what="channel_sum")
    B, C = g.shape[0], g.shape[1]
    N = g.numel() // max(1, B * C)
    out = torch.zeros(C, device=dev, dtype=torch.float32)
    if B == 0 or N == 0:
        return out
    nbytes = int(_lib.lib.mvp_channel_sum_workspace_bytes(B, C))
    ws = torch.empty(nbytes // 4, device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.mvp_channel_sum(B, C, N, _lib.ptr(g), _lib.ptr(out), _lib.ptr(ws), nbytes, _lib.stream_of(g)),
                   "mvp_channel_sum")
    return out


class _AddPerCloud(torch.autograd.Function):
    """y (B, C, N) + v (B, C) broadcast over the points (then ReLU if asked), IN PLACE on y: mvp_bias_add on the
    (1, B C, N) view; the gradient of v is mvp_channel_sum on the same view."""

    @staticmethod
    def forward(ctx, y, v, relu):
        if not y.is_contiguous():
            raise _lib.MvpOpsError("add_per_cloud: y must be contiguous")
        B, C = y.shape[0], y.shape[1]
        bias_add_(y.view(1, B * C, -1), v.contiguous().view(-1), relu)
        ctx.mark_dirty(y)
        ctx.relu = relu
        if relu:
            ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        if ctx.relu:
            (y,) = ctx.saved_tensors
            g = g * (y > 0)
        B, C = g.shape[0], g.shape[1]
        gv = channel_sum(g.view(1, B * C, -1)).view(B, C) if ctx.needs_input_grad[1] else None
        return g, gv, None


def add_per_cloud(y, v, relu=False):
    """y (B, C, N) += v (B, C)[:, :, None] in place (then ReLU): a per-cloud bias — what a 1x1 convolution contributes
    for input channels that are CONSTANT over the points (PCN's decoder concatenates the global feature to every
    point, completion/models/pcn.py:62-70).  y must be a fresh intermediate (it is overwritten)."""
    return _AddPerCloud.apply(y, v, relu)


class _ConvBias(torch.autograd.Function):
    """A wide 1x1 layer: the contraction on the library's TF32 GEMM (cuDNN, as nn.Conv would run it) WITHOUT its bias,
    the bias added in place by mvp_bias_add and its gradient summed by mvp_channel_sum."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        nd = x.dim() - 2
        if weight.dim() != x.dim() or any(k != 1 for k in weight.shape[2:]):
            raise _lib.MvpOpsError("conv_bias: weight must be (out_channels, in_channels, 1[, 1]) for a (B, C, N[, M]) input")
        y = torch.ops.aten.convolution(x, weight, None, [1] * nd, [0] * nd, [1] * nd, False, [0] * nd, 1)
        if y.numel() and y.is_contiguous():
            bias_add_(y, bias.contiguous())
        elif y.numel():  # a channels-last result: torch's own add
            y.add_(bias.view(1, -1, *([1] * nd)))
        ctx.save_for_backward(x, weight)
        return y

    @staticmethod
    def backward(ctx, g):
        x, weight = ctx.saved_tensors
        nd = x.dim() - 2
        g = g.contiguous()
        gx, gw, _ = torch.ops.aten.convolution_backward(g, x, weight, None, [1] * nd, [0] * nd, [1] * nd, False, [0] * nd, 1,
                                                        [ctx.needs_input_grad[0], ctx.needs_input_grad[1], False])
        gb = channel_sum(g) if ctx.needs_input_grad[2] else None
        return gx, gw, gb


def conv_bias(x, weight, bias):
    """A 1x1 convolution WITH bias whose contraction stays on the library (see _ConvBias): x (B, C, N[, M]),
    weight (O, C, 1[, 1]) (an nn.Conv1d's / nn.Conv2d's), bias (O) -> (B, O, N[, M])."""
    return _ConvBias.apply(x, weight, bias)


def pointwise_conv(x, weight, bias=None, relu=False):
    """A 1x1 convolution (nn.Conv1d / nn.Conv2d with kernel_size 1: completion/models/*.py, model_utils.py) as what it
    is over a point cloud — one (out, in) matrix applied to every point's feature vector — on the tensor cores through
    this repository's tcgen05 kernel (csrc/pointwise.cu): TF32 operands, fp32 accumulation, bias (and optionally ReLU)
    in the epilogue.  x (B, C, ...), weight (O, C[, 1[, 1]]), bias (O) or None -> (B, O, ...).  The input gradient runs
    through the same kernel (weight transposed); the weight gradient is a library matmul / cuDNN call."""
    return _PointwiseConv.apply(x, weight, bias, relu)


class _MaxLast(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        x_ = x.contiguous()
        dev = _lib.require_cuda(x_, dtype=torch.float32, what="max_last")
        k = x_.shape[-1]
        rows = x_.numel() // max(1, k)
        out = torch.empty(x_.shape[:-1], device=dev, dtype=torch.float32)
        arg = torch.empty(x_.shape[:-1], device=dev, dtype=torch.uint8)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.mvp_max_last(rows, k, _lib.ptr(x_), _lib.ptr(out), _lib.ptr(arg), _lib.stream_of(x_)),
                       "mvp_max_last")
        ctx.save_for_backward(arg)
        ctx.k = k
        ctx.mark_non_differentiable(arg)
        return out, arg

    @staticmethod
    def backward(ctx, g, _):
        (arg,) = ctx.saved_tensors
        g = g.contiguous()
        gx = torch.empty(*arg.shape, ctx.k, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib.mvp_max_last_grad(arg.numel(), ctx.k, _lib.ptr(g), _lib.ptr(arg), _lib.ptr(gx),
                                                  _lib.stream_of(g)), "mvp_max_last_grad")
        return gx


def max_last(x):
    """(values, arg) of the maximum over the LAST axis of x (..., k), k <= 255 — `torch.max(x, -1)` on the (B, C, N, k)
    neighbour tensors of the completion models (ecg.py:64, model_utils.py:53,104) as one bandwidth-bound launch; arg is
    uint8, the first position of the maximum; the gradient goes there.  (k > 255: torch.max itself, int64 indices.)"""
    if x.shape[-1] > 255 or x.shape[-1] == 0:
        v, i = torch.max(x, -1)
        return v, i
    return _MaxLast.apply(x)


def topk_rows(scores, k):
    """The k largest entries along the last axis of `scores` (..., cols), descending, equal scores in ascending index:
    (values, indices int64) like torch.topk(scores, k, dim=-1) — one warp per row, no multi-block radix select.  For
    the feature-space kNN of the completion models (completion/model_utils.py:242-247).  k <= 32; not differentiable
    through the values (the models only use the indices)."""
    s_ = scores.detach().contiguous()
    dev = _lib.require_cuda(s_, dtype=torch.float32, what="topk_rows")
    cols = s_.size(-1)
    rows = s_.numel() // cols
    vals = torch.empty(*s_.shape[:-1], k, device=dev, dtype=torch.float32)
    idx = torch.empty(*s_.shape[:-1], k, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.mvp_topk_rows(rows, cols, int(k), _lib.ptr(s_), _lib.ptr(vals), _lib.ptr(idx), None,
                                          _lib.stream_of(s_)), "mvp_topk_rows")
    return vals, idx


def topk_rows_sqdist(gram, sqnorm, k):
    """The feature-space kNN of completion/model_utils.py:242-247 from its two reductions, without its (B, N, N) score
    matrix: gram (B, N, N) = torch.matmul(x^T, x), sqnorm (B, N) = torch.sum(x ** 2, dim=1) -> (values, indices int64)
    of the k largest  -sqnorm[j] - (-2 gram[i, j]) - sqnorm[i]  per row — identical to
    topk_rows(-xx - (-2 * gram) - xx^T, k): the kernel performs torch's IEEE operations in torch's order on the fly."""
    g_, n_ = gram.detach().contiguous(), sqnorm.detach().contiguous()
    dev = _lib.require_cuda(g_, n_, dtype=torch.float32, what="topk_rows_sqdist")
    if g_.dim() != 3 or g_.shape[1] != g_.shape[2] or n_.shape != g_.shape[:2]:
        raise _lib.MvpOpsError("topk_rows_sqdist: gram must be (B, N, N) and sqnorm (B, N)")
    B, N = n_.shape
    vals = torch.empty(B, N, k, device=dev, dtype=torch.float32)
    idx = torch.empty(B, N, k, device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        _lib.check(_lib.lib.mvp_topk_rows_sqdist(B, N, int(k), _lib.ptr(g_), _lib.ptr(n_), _lib.ptr(vals), _lib.ptr(idx), None,
                                                 _lib.stream_of(g_)), "mvp_topk_rows_sqdist")
    return vals, idx


class _GatherMax(torch.autograd.Function):
    @staticmethod
    def forward(ctx, features, idx):
        f, i = features.contiguous(), idx.contiguous()
        dev = _lib.require_cuda(f, dtype=torch.float32, what="gather_max")
        if i.dtype != torch.int32 or i.device != f.device:
            raise _lib.MvpOpsError("gather_max: idx must be an int32 CUDA tensor on the features' device")
        B, C, N = f.shape
        M, K = i.shape[1], i.shape[2]
        out = torch.empty(B, C, M, device=dev, dtype=torch.float32)
        arg = torch.empty(B, C, M, device=dev, dtype=torch.int32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.mvp_gather_max(B, C, N, M, K, _lib.ptr(f), _lib.ptr(i), _lib.ptr(out), _lib.ptr(arg),
                                               _lib.stream_of(f)), "mvp_gather_max")
        ctx.save_for_backward(arg)
        ctx.n = N
        return out

    @staticmethod
    def backward(ctx, g):
        (arg,) = ctx.saved_tensors
        g = g.contiguous()
        B, C, M = g.shape
        grad = torch.empty(B, C, ctx.n, device=g.device, dtype=torch.float32)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib.mvp_gather_max_grad(B, C, ctx.n, M, _lib.ptr(g), _lib.ptr(arg), _lib.ptr(grad),
                                                    _lib.stream_of(g)), "mvp_gather_max_grad")
        return grad, None


def gather_max(features, idx):
    """max over the K neighbours of every point of their gathered features, without the gathered tensor:
    features (B, C, N), idx (B, M, K) int32 -> (B, C, M) — `gather_points(features, idx.view(B, M * K)).view(B, C, M, K)`
    followed by `torch.max(..., 3)[0]` (completion/model_utils.py:97-102) as ONE launch; the gradient goes to the
    neighbour that attained each maximum (the first among equals, as torch.max routes it).  N <= 16384."""
    return _GatherMax.apply(features, idx)


class _NeighborWeightedSum(torch.autograd.Function):
    @staticmethod
    def forward(ctx, y, idx, w):
        y_, i_, w_ = y.contiguous(), idx.contiguous(), w.contiguous()
        dev = _lib.require_cuda(y_, w_, dtype=torch.float32, what="neighbor_weighted_sum")
        if i_.dtype != torch.int32 or i_.device != y_.device:
            raise _lib.MvpOpsError("neighbor_weighted_sum: idx must be an int32 CUDA tensor on the features' device")
        B, C, N = y_.shape
        Cw, K = w_.shape[1], w_.shape[2]
        if i_.shape != (B, N, K) or w_.shape != (B, Cw, K, N) or C % Cw or C // Cw > 8:
            raise _lib.MvpOpsError("neighbor_weighted_sum: y (B,C,N), idx (B,N,K), w (B,Cw,K,N) with C = S*Cw, S <= 8")
        out = torch.empty(B, C, N, device=dev, dtype=torch.float32)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.mvp_neighbor_weighted_sum(B, C, Cw, N, K, _lib.ptr(y_), _lib.ptr(i_), _lib.ptr(w_),
                                                          _lib.ptr(out), _lib.stream_of(y_)), "mvp_neighbor_weighted_sum")
        ctx.save_for_backward(y_, i_, w_)
        return out

    @staticmethod
    def backward(ctx, g):
        y_, i_, w_ = ctx.saved_tensors
        g = g.contiguous()
        B, C, N = y_.shape
        Cw, K = w_.shape[1], w_.shape[2]
        gy, gw = torch.empty_like(y_), torch.empty_like(w_)
        with torch.cuda.device(g.device):
            _lib.check(_lib.lib.mvp_neighbor_weighted_sum_grad(B, C, Cw, N, K, _lib.ptr(y_), _lib.ptr(i_), _lib.ptr(w_),
                                                               _lib.ptr(g), _lib.ptr(gy), _lib.ptr(gw), _lib.stream_of(g)),
                       "mvp_neighbor_weighted_sum_grad")
        return gy, None, gw


def neighbor_weighted_sum(y, idx, w):
    """out[b,c,p] = sum_j w[b, c mod Cw, j, p] * y[b, c, idx[b,p,j]] — SA_module's aggregation of the k neighbours'
    features with attention weights shared by groups of channels (completion/models/vrcnet.py:49-52:
    `w.repeat(1, share_planes, 1, 1) * get_edge_features(y, idx)` summed over k) as ONE launch, without the (B, C, k, N)
    intermediates.  y (B, C, N), idx (B, N, K) int32, w (B, Cw, K, N) with C = S * Cw, S <= 8; differentiable in y and w."""
    return _NeighborWeightedSum.apply(y, idx, w)
