#!/usr/bin/env python
"""End-to-end training step of the reference's UNMODIFIED completion models (VRCNet / PCN / ECG) on top of
the operators of this repository, beside the same step on the reference's own CUDA kernels rebuilt for
sm_100a on the same B200 — the north star's "VRCNet end-to-end training step vs the reference ops" figure.

The models are the callers of the hot path (out of scope, not rebuilt).  They are not
part of this repository: `oracle/build_ref.py:stage_models()` stages completion/model_utils.py,
completion/models/*.py and completion/cfgs/*.yaml byte for byte under the git-ignored baseline/_ref/completion/
in the container where /root/reference exists, and this tool imports them from there.  Which operator
library the model's `from metrics import …` / `from mm3d_pn2 import …` (completion/model_utils.py:19-21)
resolve to is decided by what is first on sys.path:

    --ops ours   mvp_benchmark_b200/utils          (libmvp_ops.so, the product)
    --ops ref    oracle/ref_packages               (oracle/_ref/libref_ops.so: the reference's kernels)

One step = completion/train.py:122-142: zero_grad, forward (prefix "train", alpha from the config), backward,
Adam step; synthetic uniform clouds of the config's shape, random-initialised weights, inputs resident on the
device (as after train.py:129-131).  Timed with CUDA events around K steps.  `--both` runs the two arms in two
subprocesses (the packages have the same names) and prints one JSON object with the ratio; `--profile` adds the
per-kernel CUDA time of one step (torch.profiler) so that the share of the operators is visible.
Under torchrun (WORLD_SIZE>1) every rank runs its own B (weak scaling) with the gradient all-reduce of
mvp_benchmark_b200.dist after backward; the time is the max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STAGED = os.path.join(ROOT, "baseline", "_ref", "completion")


class Cfg(dict):
    """Attribute access over the YAML dict (what munch.munchify gives completion/train.py:200)."""
    __getattr__ = dict.__getitem__


def load_model(model_name, ops):
    import yaml
    if not os.path.isdir(STAGED):
        raise SystemExit("baseline/_ref/completion missing: run `python oracle/build_ref.py` where /root/reference exists")
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    if ops == "ours":
        import mvp_benchmark_b200
        first = mvp_benchmark_b200.install()
    else:
        first = os.path.join(ROOT, "oracle", "ref_packages")
        sys.path.insert(0, first)
    sys.path.insert(1, STAGED)
    import importlib
    import metrics
    import mm3d_pn2
    assert os.path.abspath(metrics.__file__).startswith(first), metrics.__file__
    assert os.path.abspath(mm3d_pn2.__file__).startswith(first), mm3d_pn2.__file__
    args = Cfg(yaml.safe_load(open(os.path.join(STAGED, "cfgs", f"{model_name}.yaml"))))
    module = importlib.import_module(f"models.{model_name}")      # completion/train.py:48
    return module, args


def run_arm(a):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    module, args = load_model(a.model, a.ops)
    if a.num_points:
        args["num_points"] = a.num_points                        # BASELINE config C2: PCN 2048 -> 16384
    if a.patch_knn:
        import mvp_benchmark_b200.model_patches as mp
        mp.apply(sys.modules["model_utils"], sys.modules[f"models.{a.model}"])
    torch.manual_seed(0)
    net = module.Model(args).to(dev)
    net.train()
    if a.patch_knn and not a.no_pointwise:
        mp.apply_pointwise_convs(net)
    # torch >= 1.x refuses the in-place ReLU the reference applies to a view returned by chunk()
    # (vrcnet.py:460 -> :105, "Output 0 of SplitBackward0 is a view and is being modified inplace"): switch the
    # flag on the instantiated modules — same arithmetic, both arms alike, model source untouched.
    for mod in net.modules():
        if isinstance(mod, torch.nn.ReLU):
            mod.inplace = False
    betas = tuple(float(x) for x in str(args.betas).split(","))
    opt = torch.optim.Adam(net.parameters(), lr=args.lr, weight_decay=args.weight_decay, betas=betas)
    alpha = float(str(args.varying_constant).split(",")[0])       # train.py:100-107 at epoch 0
    B, n = a.batch, args.num_points
    g = torch.Generator(device=dev).manual_seed(1 + rank)
    sets = [(torch.rand(B, 3, 2048, device=dev, generator=g), torch.rand(B, n, 3, device=dev, generator=g))
            for _ in range(4)]
    if world > 1:
        from mvp_benchmark_b200 import dist as mdist
    params = list(net.parameters())
    fwd_net = net
    reducer = None
    if world > 1 and not a.ddp and not a.no_overlap:   # bucketed all-reduce overlapped with backward (dist.py)
        reducer = mdist.OverlappedGradientAllReduce(params, bucket_bytes=a.bucket_mb << 20)
    if world > 1 and a.ddp:   # comparison arm: torch's DistributedDataParallel (bucketed all-reduce overlapped with backward)
        fwd_net = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], bucket_cap_mb=a.bucket_mb,
                                                            find_unused_parameters=True)

    def step(i):
        x, gt = sets[i % len(sets)]
        opt.zero_grad()
        out2, loss2, net_loss = fwd_net(x, gt, alpha=alpha)       # train.py:134
        net_loss.backward()                                        # train.py:141 (one device per process)
        if reducer is not None:
            reducer.finish()                                       # the one collective of a step
        elif world > 1 and not a.ddp:
            mdist.allreduce_gradients(params, world, bucket_bytes=a.bucket_mb << 20)
        opt.step()                                                 # train.py:142
        return net_loss

    for i in range(a.warmup):
        loss = step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(a.steps + 1)]
    t0 = time.perf_counter()
    evs[0].record()
    for i in range(a.steps):
        loss = step(i)
        evs[i + 1].record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / a.steps
    per_step = sorted(evs[i].elapsed_time(evs[i + 1]) for i in range(a.steps))
    mean_ms = evs[0].elapsed_time(evs[-1]) / a.steps
    ms = per_step[len(per_step) // 2]  # median step: the host side of a step (Python, allocator) hiccups now and then
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    out = {"model": a.model, "num_points": n, "ops": a.ops, "patch_knn": bool(a.patch_knn), "ddp": bool(a.ddp), "overlapped_allreduce": reducer is not None, "batch_per_gpu": B, "n_gpus": world,
           "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms, "mean_ms_per_step": mean_ms,
           "wall_ms_per_step": wall,
           "samples_per_s": B * world / ms * 1e3, "loss": float(loss.item()),
           "params": sum(p.numel() for p in net.parameters())}
    if world > 1:
        # The one collective of the step on its own: the same buckets, all-reduced back to back with nothing else on
        # the device, timed with events on the stream the collective synchronises with (torch's NCCL process group
        # makes the current stream wait for its own).  bus bandwidth = 2 (N - 1) / N x bytes / time (ring convention).
        from mvp_benchmark_b200 import dist as mdist2
        nbytes = sum(p.numel() * p.element_size() for p in params if p.requires_grad)
        flats = []
        cur, size = [], 0
        for p in params:
            if not p.requires_grad:
                continue
            if cur and size + p.numel() * 4 > (a.bucket_mb << 20):
                flats.append(torch.zeros(sum(q.numel() for q in cur), device=dev))
                cur, size = [], 0
            cur.append(p)
            size += p.numel() * 4
        if cur:
            flats.append(torch.zeros(sum(q.numel() for q in cur), device=dev))
        for _ in range(3):
            for f in flats:
                dist.all_reduce(f)
        torch.cuda.synchronize()
        dist.barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10
        a0.record()
        for _ in range(reps):
            for f in flats:
                dist.all_reduce(f)
        a1.record()
        torch.cuda.synchronize()
        ar_ms = mdist2.max_over_ranks(a0.elapsed_time(a1) / reps, dev)
        out["allreduce"] = {"bytes": nbytes, "buckets": len(flats), "standalone_ms": ar_ms,
                            "bus_GBps": 2.0 * (world - 1) / world * nbytes / (ar_ms * 1e-3) / 1e9,
                            "note": "the step's gradient all-reduce alone (same buckets, nothing else on the device); "
                                    "inside the step it overlaps the backward pass"}
    if a.profile:
        from torch.profiler import profile, ProfilerActivity
        with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
            step(0)
            torch.cuda.synchronize()
        rows = [(e.key, e.device_time_total / 1e3, e.count) for e in prof.key_averages()
                if e.device_type.name == "CUDA" and e.device_time_total > 0]
        rows.sort(key=lambda r: -r[1])
        total = sum(r[1] for r in rows)
        out["profile"] = {"cuda_ms_total": total,
                          "top": [{"kernel": k[:110], "ms": round(m, 4), "calls": c} for k, m, c in rows[:a.top]]}
    if rank == 0:
        print("MODEL_STEP " + json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="vrcnet", choices=["vrcnet", "pcn", "ecg"])
    ap.add_argument("--ops", default="ours", choices=["ours", "ref"])
    ap.add_argument("--both", action="store_true")
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--num-points", type=int, default=0, help="override cfg num_points (output / gt size)")
    ap.add_argument("--top", type=int, default=25)
    ap.add_argument("--patch-knn", action="store_true",
                    help="opt-in: replace model_utils.knn_point / knn by the fused operators (SURVEY.md §8f row 1)")
    ap.add_argument("--no-pointwise", action="store_true", help="with --patch-knn: keep cuDNN for the 1x1 convolutions")
    ap.add_argument("--ddp", action="store_true", help="N>1: torch DDP instead of mvp_benchmark_b200.dist.allreduce_gradients")
    ap.add_argument("--bucket-mb", type=int, default=32)
    ap.add_argument("--no-overlap", action="store_true", help="N>1: all-reduce after backward instead of overlapped with it")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    if not a.both:
        return run_arm(a)
    res = {}
    for ops in ("ref", "ours"):
        cmd = [sys.executable, os.path.abspath(__file__), "--model", a.model, "--ops", ops, "--batch", str(a.batch),
               "--steps", str(a.steps), "--warmup", str(a.warmup), "--top", str(a.top)]
        if a.profile:
            cmd.append("--profile")
        if a.num_points:
            cmd += ["--num-points", str(a.num_points)]
        p = subprocess.run(cmd, capture_output=True, text=True)
        line = [l for l in p.stdout.splitlines() if l.startswith("MODEL_STEP ")]
        if p.returncode != 0 or not line:
            res[ops] = {"error": (p.stderr or p.stdout)[-2000:]}
        else:
            res[ops] = json.loads(line[-1][len("MODEL_STEP "):])
    if "ms_per_step" in res.get("ref", {}) and "ms_per_step" in res.get("ours", {}):
        res["speedup_vs_reference_kernels"] = res["ref"]["ms_per_step"] / res["ours"]["ms_per_step"]
    text = json.dumps(res, indent=1)
    print(text)
    if a.out:
        with open(a.out, "w") as f:
            f.write(text + "\n")


if __name__ == "__main__":
    main()
