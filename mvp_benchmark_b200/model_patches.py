"""Opt-in replacements for the k-nearest-neighbour helpers of the reference's completion models
(completion/model_utils.py:242-271) — SURVEY.md §8(f) row 1 — and for its neighbour-feature gather
(get_edge_features, :113-124; §8(f) row 2: the grouping around the kNN) and the Chamfer loss epilogue (calc_cd,
:67-77; §8(f) row 3).  They are CALLER code, outside the drop-in
boundary, so nothing here is applied by default:

    import model_utils, models.vrcnet
    from mvp_benchmark_b200 import model_patches
    model_patches.apply(model_utils, models.vrcnet)      # rebinds knn / knn_point / knn_point_all

`from model_utils import *` copies the names into each model module, hence every module that uses them is
passed.  The originals build a (B, N, M) matrix -|x|^2 + 2 x.y - |y|^2 with a batched matmul and call
torch.topk (at VRCNet's sizes: 64 x 3072 x 3072 floats = 2.4 GB per call, ~45 % of the CUDA time of a training
step); the replacements run one exact grid search (fused.knn_points).  Differences a user can observe:
neighbours are ranked by the directly evaluated squared distance, ties by index, where the original ranks by
the matmul expansion (≈1e-7 apart; near-ties may swap) and leaves exact ties unspecified; returned distances
(knn_point) are recomputed from the gathered points, differentiable like the original's.
Point sets that are not 3-D (feature-space kNN) keep the original's score matrix and replace only its torch.topk by
fused.topk_rows (one warp per row)."""
import torch

from . import fused

_ORIGINAL = {}


def _neg_sqdist(queries, cloud, idx):
    """-(|q - c_idx|^2) for idx (B, N, k) — differentiable wrt both, what the originals' `dist` output holds."""
    B, N, k = idx.shape
    gathered = cloud[torch.arange(B, device=cloud.device).view(B, 1, 1), idx.long()]  # (B, N, k, 3)
    return -((queries.unsqueeze(2) - gathered) ** 2).sum(-1)


def knn(x, k):
    """model_utils.py:242-247: x (B, C, N) -> idx (B, N, k) int64 of the k nearest points (self included)."""
    if x.dim() == 3 and x.size(1) != 3 and x.is_cuda and x.dtype == torch.float32 and k <= min(x.size(2), 32):
        # feature-space neighbours (ECG's / EF_expansion's dynamic graph): the original's matmul and squared norms; its
        # three elementwise kernels over the (B, N, N) matrix (x -2, two broadcast subtractions: 2 ms at N = 3072) and
        # torch.topk's multi-block radix select are ONE pass of one warp per row over the Gram matrix
        # (fused.topk_rows_sqdist: the same IEEE operations in the same order, so the same neighbours)
        gram = torch.matmul(x.transpose(2, 1).contiguous(), x)
        return fused.topk_rows_sqdist(gram, torch.sum(x ** 2, dim=1), k)[1]
    if x.dim() != 3 or x.size(1) != 3 or not x.is_cuda or x.dtype != torch.float32 or k > min(x.size(2), 64):
        return _ORIGINAL["knn"](x, k)
    _, idx = fused.knn_points(k, x.transpose(1, 2))
    return idx.long()


def knn_point(pk, point_input, point_output):
    """model_utils.py:250-259: for every point of point_output (B, m, 3) its pk nearest points of point_input
    (B, n, 3) -> (dist (B, m, pk) = NEGATIVE squared distances, descending; idx (B, m, pk) int64)."""
    if (point_input.dim() != 3 or point_input.size(2) != 3 or not point_input.is_cuda
            or point_input.dtype != torch.float32 or pk > min(point_input.size(1), 64)
            or point_output.dim() != 3 or point_output.size(2) != 3 or not point_output.is_cuda
            or point_output.dtype != torch.float32 or point_output.size(0) != point_input.size(0)):
        return _ORIGINAL["knn_point"](pk, point_input, point_output)
    _, idx = fused.knn_points(pk, point_input, point_output)
    return _neg_sqdist(point_output, point_input, idx), idx.long()


def get_edge_features(x, idx):
    """model_utils.py:113-124: x (B, C, 1, N), idx (B, N, k) -> the neighbours' features (B, C, k, N).  The original
    transposes x, gathers rows with advanced indexing and returns a permuted (non-contiguous) view that the next
    convolution has to copy; this is the same gather as ONE grouping_operation call (the reference's own operator,
    utils/mm3d_pn2/ops/group_points/group_points.py:166-218) on the transposed index, contiguous in the layout the
    convolution wants, with the operator's scatter as backward.  Same values bit for bit in the forward pass."""
    if (x.dim() != 4 or x.size(2) != 1 or not x.is_cuda or x.dtype != torch.float32 or idx.dim() != 3
            or idx.size(1) != x.size(3)):
        return _ORIGINAL["get_edge_features"](x, idx)
    import mm3d_pn2
    return mm3d_pn2.grouping_operation(x.squeeze(2).contiguous(), idx.transpose(1, 2).int().contiguous())


def get_graph_feature(x, k=20, minus_center=True):
    """model_utils.py:156-178: x (B, C, N) -> (B, 2C, N, k): every point's features next to (its k nearest neighbours'
    features minus its own).  The original gathers rows of the transposed x with advanced indexing
    (`x.view(B * N, -1)[idx, :]`: an index kernel forward, a sort + index_put backward), concatenates on the LAST axis
    and returns a permuted, non-contiguous view that the next convolution copies.  Here: the same neighbours by ONE
    grouping_operation call (the reference's own operator; its scatter is the backward), the subtraction and the
    concatenation channel-first and contiguous.  Same values bit for bit in the forward pass."""
    if not (x.dim() == 3 and x.is_cuda and x.dtype == torch.float32 and k <= x.size(2)):
        return _ORIGINAL["get_graph_feature"](x, k, minus_center)
    import mm3d_pn2
    idx = knn(x, k)                                                           # (B, N, k) int64
    feature = mm3d_pn2.grouping_operation(x.contiguous(), idx.int().contiguous())   # (B, C, N, k)
    centre = x.unsqueeze(3).expand(-1, -1, -1, k)
    return torch.cat((centre, feature - centre if minus_center else feature), dim=1)


def three_nn_upsampling(target_points, source_points):
    """model_utils.py:286-293: (idx, weight) for three_interpolate — three_nn plus five torch kernels of glue as the
    grid search plus one elementwise kernel (fused.three_nn_weights).  Same values."""
    if not target_points.is_cuda or target_points.dtype != torch.float32:
        return _ORIGINAL["three_nn_upsampling"](target_points, source_points)
    return fused.three_nn_weights(target_points, source_points)


def edge_preserve_sampling(feature_input, point_input, num_samples, k=10):
    """model_utils.py:86-108 with its first two statements (furthest_point_sample, then transpose + gather_points +
    transpose back) as ONE launch (fused.fps_gather) and the neighbour-feature gather + torch.max
    as ONE launch (fused.gather_max: the (B, C, M, pk) tensor never exists); the rest is the original's sequence."""
    if not point_input.is_cuda or point_input.dtype != torch.float32 or point_input.dim() != 3 or point_input.size(2) != 3:
        return _ORIGINAL["edge_preserve_sampling"](feature_input, point_input, num_samples, k)
    import mm3d_pn2
    batch_size, feature_size, num_points = feature_input.size()
    p_idx, point_output = fused.fps_gather(point_input, num_samples)
    pk = int(min(k, num_points))
    _, pn_idx = knn_point(pk, point_input, point_output)
    pn_idx = pn_idx.detach().int()
    if feature_input.is_cuda and feature_input.dtype == torch.float32 and num_points <= 16384:
        neighbor_feature = fused.gather_max(feature_input, pn_idx)   # gather + max over the pk neighbours: ONE launch
    else:
        neighbor_feature = mm3d_pn2.gather_points(feature_input, pn_idx.view(batch_size, num_samples * pk)).view(
            batch_size, feature_size, num_samples, pk)
        neighbor_feature, _ = torch.max(neighbor_feature, 3)
    center_feature = mm3d_pn2.grouping_operation(feature_input, p_idx.unsqueeze(2)).view(batch_size, -1, num_samples)
    net = torch.cat((center_feature, neighbor_feature), 1)
    return net, p_idx, pn_idx, point_output


def get_uniform_loss(pcd, percentages=[0.004, 0.006, 0.008, 0.010, 0.012], radius=1.0):
    """model_utils.py:201-227 (ECG's uniformity loss) with FPS + gather as one launch (fused.fps_gather) and
    ball_query + grouping_operation + permute as one launch (fused.ball_query_group); the arithmetic after the grouping
    is the original's, statement for statement."""
    import math
    if not pcd.is_cuda or pcd.dtype != torch.float32:
        return _ORIGINAL["get_uniform_loss"](pcd, percentages, radius)
    B, N, C = pcd.size()
    npoint = int(N * 0.05)
    loss = 0
    for p in percentages:
        nsample = int(N * p)
        r = math.sqrt(p * radius)
        disk_area = math.pi * (radius ** 2) * p / nsample
        _, new_xyz = fused.fps_gather(pcd, npoint)
        _, grouped_pcd = fused.ball_query_group(0, r, nsample, pcd, new_xyz)
        expect_len = math.sqrt(disk_area)
        grouped_pcd = grouped_pcd.view(-1, nsample, 3)
        var, _ = knn_point(2, grouped_pcd, grouped_pcd)
        uniform_dis = -var[:, :, 1:]
        uniform_dis = torch.sqrt(torch.abs(uniform_dis + 1e-8))
        uniform_dis = torch.mean(uniform_dis, dim=-1)
        uniform_dis = ((uniform_dis - expect_len) ** 2 / (expect_len + 1e-8))
        mean = torch.mean(uniform_dis)
        mean = mean * math.pow(p * 100, 2)
        loss += mean
    return loss / len(percentages)


def sa_module_forward(self, input):
    """completion/models/vrcnet.py:43-57 (SA_module.forward) with the two convolutions of the NEIGHBOUR tensor moved
    in front of the gather.  The original gathers the k neighbours' features first — xn = get_edge_features(x, idx), a
    (B, C, k, N) tensor, 1 GB at the first level — and then applies conv2 and conv3, both 1x1, to it.  A 1x1 convolution
    acts on every (b, :, j, n) column independently, so it commutes with the gather:
        conv(get_edge_features(x, idx)) == get_edge_features(conv(x), idx)          (bias included)
    Here conv2 and conv3 run once per POINT on (B, C, 1, N) — k times fewer multiply-adds, no 1 GB tensor, and their
    weight / data gradients shrink by the same factor (cuDNN's weight-gradient kernel over (B, C, k, N) is the largest
    single kernel of the step otherwise) — and the gathers move C/16 and C/4 channels instead of C (SURVEY.md §8f row 4:
    the grouped-feature MLP; no tensor-core kernel is needed for work that is not done).  Same arithmetic per output
    element (one dot product over the C input channels); everything after the gathers is the original's code."""
    import sys
    x, idx = input
    if not x.is_cuda or x.dtype != torch.float32:
        return _ORIGINAL["SA_module.forward"](self, input)
    gef = sys.modules[type(self).__module__].get_edge_features
    batch_size, _, _, num_points = x.size()
    identity = x  # B C 1 N
    x = self.activation_fn(x)
    x1, y2, y3 = self.conv1(x), self.conv2(x), self.conv3(x)     # all three on (B, C, 1, N)
    x2 = gef(y2, idx)                                              # B C/16 K N
    x2 = x2.contiguous().view(batch_size, -1, 1, num_points)       # B kC 1 N
    w = self.conv_w(torch.cat([x1, x2], 1)).view(batch_size, -1, self.k, num_points)
    if self.share_planes <= 8 and num_points <= 6144 and idx.dim() == 3 and idx.size(2) == self.k:
        # sum_j w[c mod Cw, j] * y3[c, idx_j]: one launch instead of gather, repeat, multiply and sum over (B, C, K, N)
        out = fused.neighbor_weighted_sum(y3.squeeze(2), idx.int(), w).unsqueeze(2)
    else:
        x3 = gef(y3, idx)                                          # B C/4 K N
        w = w.repeat(1, self.share_planes, 1, 1)
        out = w * x3
        out = torch.sum(out, dim=2, keepdim=True)
    out = self.activation_fn(out)
    out = self.conv_out(out)  # B C 1 N
    out += identity
    return [out, idx]


def _dense_conv_forward_for(module):
    """completion/models/ecg.py:58-65 (Dense_conv.forward) with its last line — `y, _ = torch.max(y, 3)` over the k
    neighbours of the (B, C, N, k) tensor, which torch's generic reduce kernel runs at 0.36 TB/s (604 MB in 1.67 ms at
    the first level) — through fused.max_last (mvp_max_last: a thread per row of k values; the gradient is written
    whole, no memset + scatter).  Everything else is the original's code, against the names of `module`."""
    import torch.nn.functional as F

    def dense_conv_forward(self, x):
        y = module.get_graph_feature(x, k=self.k)
        y = F.relu(self.first_conv(y))
        y = torch.cat((y, x.unsqueeze(3).repeat(1, 1, 1, self.k)), 1)
        y = self.model(y)
        if y.is_cuda and y.dtype == torch.float32:
            y, _ = fused.max_last(y)
        else:
            y, _ = torch.max(y, 3)
        return y
    return dense_conv_forward


def pcn_decoder_forward(self, x):
    """completion/models/pcn.py:48-71 (PCN_decoder.forward) without the (B, 1029, num_fine) tensor.  The original
    repeats the 1024-channel global feature over all num_fine points, concatenates it with 2 grid and 3 point channels
    and runs conv1 (1029 -> 512) over the result: 552 GFLOP at B = 32, num_fine = 16 384, and — 1029 not being a multiple
    of 4 — on the library's unaligned legacy TF32 kernels (4.6 + 3.9 + 3.7 ms forward / input / weight gradient of a
    22 ms step).  A 1x1 convolution of channels that are CONSTANT over the points is one vector per cloud:
        conv1(cat(grid, point, global)) = W[:, :5] . cat(grid, point) + (W[:, 5:] . global)[:, :, None] + bias
    so conv1 becomes a 5-channel contraction (fused.pointwise_conv: the tcgen05 kernel, bandwidth-bound), a (B, 1024) x
    (1024, 512) matmul, and one in-place pass that adds the per-cloud vector and applies the ReLU (fused.add_per_cloud).
    Same function of the same parameters (state_dict unchanged); sums are taken in a different order."""
    import torch.nn.functional as F
    if not (x.is_cuda and x.dtype == torch.float32 and self.conv1.weight.shape[1] == x.shape[1] + 5):
        return type(self)._mvp_original_forward(self, x)
    batch_size = x.size()[0]
    coarse = F.relu(self.fc1(x))
    coarse = F.relu(self.fc2(coarse))
    coarse = self.fc3(coarse).view(-1, 3, self.num_coarse)
    grid_feat = self.grid.detach().unsqueeze(0).repeat(batch_size, 1, self.num_coarse).contiguous()
    point_feat = (coarse.transpose(1, 2).contiguous()).unsqueeze(2).repeat(1, 1, self.scale, 1).view(
        -1, self.num_fine, 3).transpose(1, 2).contiguous()          # also the original's `center`
    w = self.conv1.weight                                            # (512, 2 + 3 + 1024, 1)
    hidden = fused.pointwise_conv(torch.cat((grid_feat, point_feat), 1), w[:, :5], self.conv1.bias)
    hidden = fused.add_per_cloud(hidden, F.linear(x, w[:, 5:, 0]), relu=True)
    fine = self.conv3(F.relu(self.conv2(hidden))) + point_feat
    return coarse, fine


def _pointwise_conv_forward(self, x):
    """An nn.Conv1d / nn.Conv2d with a 1x1 kernel as what it is over a point cloud — one (out, in) matrix applied to every
    point's feature vector — routed by shape (measured on B200, tools/pointwise_probe.py, profiles/r2_pointwise.md):
      * up to 128 input and at least 64 output channels (HBM-bound: 64 -> 256 over 64 x 3072 points moves 252 MB for
        6.4 GFLOP): this repository's tcgen05 kernel, bias in its epilogue (fused.pointwise_conv; 0.058 ms against
        cuDNN + torch's bias add 0.187 ms), input gradient through the same kernel;
      * THIN layers ((64, 64, 1, 3072) -> 4, 16 or 64 channels; 68 -> 2): torch.baddbmm, a library fp32 GEMM — cuDNN's
        heuristics pick `wgrad2d_grouped_direct_kernel` for their weight gradient, 0.5-1.6 ms per call for 0.1-1.6 GFLOP:
        8.8 ms of a 36 ms VRCNet training step;
      * the wide, compute-bound layers with a bias (512 -> 1024 over 64 x 2048 points: 137 GFLOP): the library's TF32
        GEMM without the bias, the bias added in place and its gradient summed by this repository's kernels
        (fused.conv_bias: torch's broadcasting add and its sum run at 2.7 and 2.1 TB/s, these at 6.7 and 5.6);
      * anything else: the module's own forward."""
    shp = x.shape
    cin, cout = self.in_channels, self.out_channels
    if not (x.is_cuda and x.dtype == torch.float32 and x.dim() >= 3):
        return self._conv_forward(x, self.weight, self.bias)
    thin = cin * cout <= kPointwiseMaxWeights and 2.0 * x.numel() * cout <= kPointwiseMaxFlops
    if x.numel() >= kTensorCoreMinElems and (thin or (cin <= kTensorCoreMaxIn and cout >= kTensorCoreMinOut)):
        return fused.pointwise_conv(x, self.weight, self.bias)
    if thin:
        xs = x.reshape(shp[0], shp[1], -1)
        w = self.weight.view(1, cout, cin).expand(shp[0], -1, -1)
        if self.bias is not None:
            y = torch.baddbmm(self.bias.view(1, -1, 1), w, xs)
        else:
            y = torch.bmm(w, xs)
        return y.view(shp[0], cout, *shp[2:])
    if self.bias is not None and x.numel() // cin * cout >= kConvBiasMinElems:
        return fused.conv_bias(x, self.weight, self.bias)
    return self._conv_forward(x, self.weight, self.bias)   # cuDNN's TF32 kernels


kPointwiseMaxWeights = 16384  # in_channels * out_channels up to 64 x 256: above that cuDNN's TF32 kernels are the faster ones
kPointwiseMaxFlops = 4e9      # ... and so they are for a thin convolution over a very long tensor (decided per call)
kTensorCoreMaxIn = 128        # the tcgen05 kernel: weight tile resident, HBM-bound shapes
kTensorCoreMinOut = 64
kTensorCoreMinElems = 1 << 20  # input elements below which a launch is all there is to save
kConvBiasMinElems = 1 << 22    # output elements from which the separate bias kernels pay


def apply_pointwise_convs(model, max_weights=None):
    """Route the 1x1, stride-1, ungrouped nn.Conv1d / nn.Conv2d modules of an instantiated model through
    _pointwise_conv_forward (same parameters, same state_dict; opt-in like everything in this file; the route is chosen
    per call from the shapes).  max_weights: only modules with in_channels x out_channels up to it (None: all).
    Returns the number of modules switched."""
    import types
    count = 0
    for mod in model.modules():
        if (isinstance(mod, (torch.nn.Conv1d, torch.nn.Conv2d)) and all(k == 1 for k in mod.kernel_size)
                and all(s_ == 1 for s_ in mod.stride) and not isinstance(mod.padding, str)
                and all(p_ == 0 for p_ in mod.padding) and all(d == 1 for d in mod.dilation) and mod.groups == 1
                and (max_weights is None or mod.in_channels * mod.out_channels <= max_weights)):
            mod.forward = types.MethodType(_pointwise_conv_forward, mod)
            count += 1
    return count


def calc_cd(output, gt, calc_f1=False):
    """model_utils.py:67-77: the Chamfer operator, then its loss epilogue (two sqrt, four means, three elementwise
    torch kernels) as ONE reduction kernel (fused.chamfer_loss).  Same returns."""
    import metrics
    dist1, dist2, _, _ = metrics.cd()(gt, output)
    if not dist1.is_cuda:
        return _ORIGINAL["calc_cd"](output, gt, calc_f1)
    cd_p, cd_t = fused.chamfer_loss(dist1, dist2)
    if calc_f1:
        f1, _, _ = fused.fscore(dist1, dist2)
        return cd_p, cd_t, f1
    return cd_p, cd_t


def calc_emd(output, gt, eps=0.005, iterations=50):
    """model_utils.py:80-85: the EMD operator, then `torch.sqrt(dist).mean(1)` as the one reduction kernel of the
    Chamfer epilogue (fused.emd_loss).  Same return."""
    import metrics
    dist, _ = metrics.emd()(output, gt, eps, iterations)
    if not dist.is_cuda:
        return _ORIGINAL["calc_emd"](output, gt, eps, iterations)
    return fused.emd_loss(dist)


def apply(*modules):
    """Rebind knn / knn_point / knn_point_all / get_edge_features / get_graph_feature / calc_cd / calc_emd / three_nn_upsampling /
    edge_preserve_sampling / get_uniform_loss in the given (already imported) modules (and SA_module.forward,
    Dense_conv.forward where the module defines those classes).
    Returns the number of names replaced."""
    from . import install
    install()  # `mm3d_pn2` must resolve to this repository's package
    count = 0
    for mod in modules:
        for name, fn in (("knn", knn), ("knn_point", knn_point), ("knn_point_all", knn_point),
                         ("get_edge_features", get_edge_features), ("get_graph_feature", get_graph_feature),
                         ("calc_cd", calc_cd), ("calc_emd", calc_emd),
                         ("three_nn_upsampling", three_nn_upsampling),
                         ("edge_preserve_sampling", edge_preserve_sampling), ("get_uniform_loss", get_uniform_loss)):
            cur = getattr(mod, name, None)
            if cur is None or cur is fn:
                continue
            _ORIGINAL.setdefault("knn_point" if name == "knn_point_all" else name, cur)
            setattr(mod, name, fn)
            count += 1
        dc = getattr(mod, "Dense_conv", None)  # models.ecg
        if isinstance(dc, type) and hasattr(mod, "get_graph_feature") and dc.forward.__name__ != "dense_conv_forward":
            _ORIGINAL.setdefault("Dense_conv.forward", dc.forward)
            dc.forward = _dense_conv_forward_for(mod)
            count += 1
        pd = getattr(mod, "PCN_decoder", None)  # models.pcn
        if isinstance(pd, type) and pd.forward is not pcn_decoder_forward:
            pd._mvp_original_forward = pd.forward   # kept on the class: several modules may define a PCN_decoder
            pd.forward = pcn_decoder_forward
            count += 1
        cls = getattr(mod, "SA_module", None)  # models.vrcnet: the class's forward, not a module-level function
        if isinstance(cls, type) and cls.forward is not sa_module_forward and all(
                hasattr(cls, a) for a in ("forward",)):
            _ORIGINAL.setdefault("SA_module.forward", cls.forward)
            cls.forward = sa_module_forward
            count += 1
    return count
