import sys, os, torch, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mvp_benchmark_b200; mvp_benchmark_b200.install()
import metrics
from oracle import ref_cuda
def t(fn, it=3):
    fn(); torch.cuda.synchronize()
    e0,e1=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1)/it
for (b,n,eps,iters) in ((32,2048,0.004,3000),(32,2048,0.005,50),(64,8192,0.005,50),(8,8192,0.004,3000)):
    x1,x2=torch.rand(b,n,3,device='cuda'),torch.rand(b,n,3,device='cuda')
    ours=t(lambda: metrics.emd()(x1,x2,eps,iters))
    ref=t(lambda: ref_cuda.emd_forward(x1,x2,eps,iters))
    d,a=metrics.emd()(x1,x2,eps,iters)
    print(f"emd {b}x{n} eps {eps} iters {iters}: ours {ours:8.2f} ms  ref {ref:8.2f} ms  {ref/ours:5.2f}x  unique {a[0].unique().numel()}", flush=True)
