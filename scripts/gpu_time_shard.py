"""2+ GPU check of the time-sharded objective (run under torchrun): one long sequence, blocks of time per rank, carries by
all-gather, [loss, grad] by NCCL all-reduce; compared with the whole sequence on one GPU.  Prints one summary line."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bench import model_params, DT
from multioutputihgp_b200 import MOIHGPSequences
from multioutputihgp_b200.parallel import TimeShardedDeviceObjective, TimeShardedObjective, time_block_bounds_aligned

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
# TS_BACKEND=gloo: the ranks share the visible GPUs (rank r on device r % device_count) and exchange through gloo - lets a
# ONE-GPU box run the whole protocol with two ranks (NCCL refuses two ranks on one device)
backend = os.environ.get("TS_BACKEND", "nccl")
local = local % torch.cuda.device_count() if backend == "gloo" else local
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if backend == "gloo":
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=dev)
p, L, T, kernel = int(os.environ.get("TS_P", 64)), int(os.environ.get("TS_L", 32)), int(os.environ.get("TS_T", 400000)), "Matern32"
params, Hmix = model_params(p, L, kernel, 4321)
m = MOIHGPSequences(DT, p, L, kernel, threading=True, device=local)
m.update(params)
rng = np.random.default_rng(99)                      # every rank builds the same sequence and keeps its block
t = np.arange(T) * DT
F = np.sin(t[:, None] * (1.0 + 3.0 * np.arange(L) / max(L - 1, 1))[None, :])
Y = F @ Hmix.T + 0.1 * (2 * rng.random((T, p)) - 1)
bounds = [time_block_bounds_aligned(T, world, r) for r in range(world)]
consts = [m.latent_consts(l) for l in range(L)]
t0_b, t1_b = bounds[rank]


d = m.igp_dim
Yb_dev = torch.from_numpy(np.ascontiguousarray(Y[t0_b:t1_b])).to(dev)[None].contiguous()     # this rank's block, resident in HBM
out = torch.zeros(2 + m.num_param, dtype=torch.float64, device=dev)
xT_dev = torch.zeros((1, L, d), dtype=torch.float64, device=dev)
dxT_dev = torch.zeros((1, L, 3, d), dtype=torch.float64, device=dev)


def evaluate_dev(_Yb, x0, dx0):
    """Device-resident evaluator: only the L*d*(1+K) carry doubles and [loss, grad] cross PCIe."""
    x0d = torch.from_numpy(np.ascontiguousarray(x0[None])).to(dev)
    dx0d = torch.from_numpy(np.ascontiguousarray(dx0[None])).to(dev)
    m.objective_device(Yb_dev, out[0:1], out[2:], x0=x0d, dx0=dx0d, xT=xT_dev, dxT=dxT_dev)
    torch.cuda.synchronize(dev)
    h = out.cpu().numpy()
    return float(h[0]), h[2:].copy(), xT_dev.cpu().numpy()[0], dxT_dev.cpu().numpy()[0]


obj = TimeShardedObjective(evaluate_dev, m.block_transition, [b[1] - b[0] for b in bounds], m.num_param, device=dev, num_latent=L, igp_dim=d)
obj(None)                                             # warm-up
dist.barrier()
tic = time.perf_counter()
loss, grad = obj(None)
dist.barrier()
t_shard = time.perf_counter() - tic
# one-pass variant: begin (projection + summaries + block end) -> all-gather -> finish (from the true carry-in)
obj1 = TimeShardedDeviceObjective(m, [b[1] - b[0] for b in bounds])
obj1(Yb_dev)
dist.barrier()
tic = time.perf_counter()
loss1, grad1 = obj1(Yb_dev)
dist.barrier()
t_one_pass = time.perf_counter() - tic
if rank == 0:
    Yd = torch.from_numpy(Y).to(dev)[None].contiguous()
    o1 = torch.zeros(2 + m.num_param, dtype=torch.float64, device=dev)
    m.objective_device(Yd, o1[0:1], o1[2:])
    torch.cuda.synchronize(dev)
    tic = time.perf_counter()
    m.objective_device(Yd, o1[0:1], o1[2:])
    torch.cuda.synchronize(dev)
    t_one = time.perf_counter() - tic
    h1 = o1.cpu().numpy()
    l1, g1 = float(h1[0]), h1[2:]
    err_l = abs(loss - l1) / abs(l1)
    err_g = float(np.max(np.abs(grad - g1)) / np.max(np.abs(g1)))
    err_l1 = abs(loss1 - l1) / abs(l1)
    err_g1 = float(np.max(np.abs(grad1 - g1)) / np.max(np.abs(g1)))
    print("time-sharded objective: world=%d p=%d L=%d T=%d  rel.err two-pass loss %.2e grad %.2e, one-pass loss %.2e grad %.2e | "
          "device-resident wall: two-pass %.2f ms, one-pass %.2f ms, one GPU %.2f ms"
          % (world, p, L, T, err_l, err_g, err_l1, err_g1, 1e3 * t_shard, 1e3 * t_one_pass, 1e3 * t_one))
    assert err_l < 1e-9 and err_g < 1e-9 and err_l1 < 1e-9 and err_g1 < 1e-9
dist.destroy_process_group()
