"""Probe: do the per-kernel CUDA events perturb / misreport the fused pass?  (debug helper, not a benchmark)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
from bench import model_params, make_data_device, DT
from multioutputihgp_b200 import MOIHGPSequences
p, L, N, T, d = 16, 8, 4096, 16384, 3
dev = torch.device("cuda", 0)
m = MOIHGPSequences(DT, p, L, "Matern52", threading=True, device=0)
params, Hmix = model_params(p, L, "Matern52", 1236)
m.update(params)
Y = make_data_device(torch, dev, Hmix, N, T, p, L, 1236, 0)
X = torch.empty((N, T, L, d), dtype=torch.float64, device=dev); Xs = torch.empty_like(X); nll = torch.empty(N, dtype=torch.float64, device=dev)
def loop(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k):
        m.filter_smoother_nll_device(Y, smoother_mode=1, X=X, Xs=Xs, nll=nll)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
for _ in range(3): loop(2)
for rep in range(3):
    a = loop(10)
    m.profile(True); b = loop(10); pr = m.profile_read(); m.profile(False)
    print("plain %.3f ms/step | with markers %.3f ms/step | marker sums %s" % (a, b, {k: round(v[0] / v[1], 3) for k, v in pr.items()}))
# filter only / smoother cost by difference
def loopf(k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(k):
        m.filter_smoother_nll_device(Y, smoother_mode=-1, X=X, nll=nll)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k
print("filter-only loop %.3f ms/step" % loopf(10))
