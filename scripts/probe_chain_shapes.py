"""Many-chains path vs chunked-scan path for every instantiated (p, L) shape: device-resident fused pass (filter + RTS
smoother + NLL) on synthetic data, CUDA events, ms per pass and latent-steps/s.  Run on a B200; prints one line per shape."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from bench import DT, model_params
from multioutputihgp_b200 import MOIHGPSequences

dev = torch.device("cuda:0")
# (32, 32), (64, 8), (64, 16) were instantiated and measured in round 2: the scan path is 2 - 4x faster there
shapes = [(4, 2), (4, 4), (8, 2), (8, 4), (8, 8), (16, 2), (16, 4), (16, 8), (16, 16), (32, 2), (32, 4), (32, 8), (32, 16), (32, 32)]
if len(sys.argv) > 1 and sys.argv[1] == "padded":      # p between the instantiated widths: the padded variant of k_filter_chain
    shapes = [(6, 2), (6, 4), (10, 4), (12, 4), (12, 8), (14, 8), (20, 4), (24, 8), (28, 16), (30, 2), (5, 2), (7, 4), (11, 8), (13, 4), (25, 8)]
kernels = ("Matern32", "Matern52")
if len(sys.argv) > 1 and sys.argv[1] == "one":         # one latent: Matern-3/2 only (chain.cu)
    shapes = [(2, 1), (4, 1), (7, 1), (8, 1), (16, 1), (32, 1)]
    kernels = ("Matern32",)
if len(sys.argv) > 1 and sys.argv[1] == "padL":        # L between the instantiated widths (Matern-5/2: even L only)
    shapes = [(8, 3), (12, 5), (8, 6), (16, 6), (16, 7), (16, 10), (16, 12), (24, 6), (30, 14)]
T = 4096
for kernel in kernels:
    for p, L in shapes:
        N = max(256, int(2.5e8 / (T * L)) // 32 * 32)
        if (L * (3 if kernel == "Matern52" else 2)) % 2 and L not in (1, 2, 4, 8, 16):
            continue
        params, Hmix = model_params(p, L, kernel, 4321)
        m = MOIHGPSequences(DT, p, L, kernel, threading=True, device=0)
        m.update(params)
        d = m.igp_dim
        Y = torch.randn((N, T, p), dtype=torch.float64, device=dev)
        X = torch.empty((N, T, L, d), dtype=torch.float64, device=dev)
        Xs = torch.empty_like(X)
        nll = torch.empty(N, dtype=torch.float64, device=dev)
        out = {}
        for path in ("chain", "scan"):
            m.set_path(path)
            for _ in range(2):
                m.filter_smoother_nll_device(Y, smoother_mode=1, X=X, Xs=Xs, nll=nll)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                m.filter_smoother_nll_device(Y, smoother_mode=1, X=X, Xs=Xs, nll=nll)
            e1.record()
            torch.cuda.synchronize()
            out[path] = e0.elapsed_time(e1) / 5
        balg = 8.0 * (p / L + 2 * d) * N * T * L
        print("%s p=%2d L=%2d N=%5d T=%d: chain %.3f ms (%.2f TB/s alg, %.0f%% of 6553 GB/s), scan %.3f ms -> chain is %.2fx"
              % (kernel, p, L, N, T, out["chain"], balg / out["chain"] / 1e9, 100 * balg / out["chain"] / 1e6 / 6553.3, out["scan"], out["scan"] / out["chain"]))
        del Y, X, Xs, m
        torch.cuda.empty_cache()
