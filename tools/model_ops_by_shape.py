"""Diagnostic: one patched training step (python tools/model_ops_by_shape.py [vrcnet|ecg|pcn]) under torch.profiler with record_shapes — device time per aten op and input
shape (how the thin 1x1 convolutions whose weight gradient cuDNN runs through wgrad2d_grouped_direct_kernel were found).
Run under gpurun: python tools/model_ops_by_shape.py"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
MODEL = sys.argv[1] if len(sys.argv) > 1 else "vrcnet"
sys.argv = ["model_step.py", "--model", MODEL, "--ops", "ours", "--patch-knn"]
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__))))
import model_step as ms
module, args = ms.load_model(MODEL, "ours")
import mvp_benchmark_b200.model_patches as mp
mp.apply(sys.modules["model_utils"], sys.modules["models." + MODEL])
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = module.Model(args).to(dev).train()
mp.apply_pointwise_convs(net)
for mod in net.modules():
    if isinstance(mod, torch.nn.ReLU):
        mod.inplace = False
x, gt = torch.rand(32, 3, 2048, device=dev), torch.rand(32, 2048, 3, device=dev)
for _ in range(2):
    net.zero_grad(); out = net(x, gt, alpha=0.01); out[2].mean().backward()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU], record_shapes=True) as prof:
    net.zero_grad(); out = net(x, gt, alpha=0.01); out[2].mean().backward(); torch.cuda.synchronize()
rows = []
for e in prof.key_averages(group_by_input_shape=True):
    if e.device_time_total > 150 and e.key.startswith("aten::") and not e.key.startswith("aten::_") :
        rows.append((e.device_time_total / 1e3, e.count, e.key, str(e.input_shapes)[:150]))
rows.sort(reverse=True)
for r in rows[:70]:
    print("%.3f ms x%d %s %s" % r)

# the same step by operator name only (no shape split, no threshold): where the long tail goes
tot = {}
nk = 0
for e in prof.key_averages():
    if e.device_time_total > 0 and not e.key.startswith("aten::") and not e.key.startswith("autograd::") and "Backward" not in e.key:
        nk += e.count
for e in prof.key_averages():
    if e.key.startswith("aten::") and e.self_device_time_total > 0:
        tot[e.key] = (e.self_device_time_total / 1e3, e.count)
print("-- self device time by aten op (ms, calls); %d device kernels/memsets in the step" % nk)
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][0])[:40]:
    print("%.3f ms x%d %s" % (v[0], v[1], k))
