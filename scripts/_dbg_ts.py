import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from bench import model_params, DT
from multioutputihgp_b200 import MOIHGPSequences
from multioutputihgp_b200.parallel import time_block_bounds_aligned, TimeShardedDeviceObjective, stack_consts, carry_in_from_block_ends
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
p, L, T = 16, 8, 60000
params, Hmix = model_params(p, L, "Matern32", 4321)
m = MOIHGPSequences(DT, p, L, "Matern32", threading=True, device=local); m.update(params)
rng = np.random.default_rng(99); t = np.arange(T) * DT
Y = np.sin(t[:, None] * (1.0 + 3.0 * np.arange(L) / max(L - 1, 1))[None, :]) @ Hmix.T + 0.1 * (2 * rng.random((T, p)) - 1)
b = [time_block_bounds_aligned(T, world, r) for r in range(world)]
Yblk = np.ascontiguousarray(Y[b[rank][0]:b[rank][1]])
Yb = torch.from_numpy(Yblk).to(dev)[None].contiguous()
d = m.igp_dim
# true carry-in from the host-buffer API on the preceding part of the sequence
if rank == 0:
    x0, dx0 = np.zeros((1, L, d)), np.zeros((1, L, 3, d))
else:
    _, _, x0, dx0 = m.objective(Y[None, :b[rank][0]], want_state=True)
ref_loss, ref_grad = m.objective(Yblk[None], x0=x0, dx0=dx0)
buf = torch.zeros(2 + m.num_param, dtype=torch.float64, device=dev)
m.objective_begin_device(Yb, want_end=False)
m.objective_finish_device(Yb, buf[0:1], buf[2:], x0=torch.from_numpy(x0).to(dev), dx0=torch.from_numpy(dx0).to(dev))
torch.cuda.synchronize()
h = buf.cpu().numpy()
print("rank%d block [%d,%d): begin/finish loss %.10g  host-API loss %.10g  grad err %.2e" % (rank, b[rank][0], b[rank][1], h[0], ref_loss, np.abs(h[2:] - ref_grad).max() / np.abs(ref_grad).max()), flush=True)
obj1 = TimeShardedDeviceObjective(m, [bb[1] - bb[0] for bb in b])
end = m.objective_begin_device(Yb, want_end=(rank < world - 1))
flat = np.zeros(L * d * 4) if rank == world - 1 else np.concatenate([end[0][0].ravel(), end[1][0].ravel()])
tt = torch.from_numpy(flat).to(dev)
outl = [torch.empty_like(tt) for _ in range(world)]
dist.all_gather(outl, tt)
ends = [o.cpu().numpy() for o in outl]
consts = stack_consts([m.latent_consts(l) for l in range(L)])
xin, dxin = carry_in_from_block_ends(consts, [bb[1] - bb[0] for bb in b], [e[:L * d].reshape(L, d) for e in ends], [e[L * d:].reshape(L, 3, d) for e in ends], np.zeros((L, d)), np.zeros((L, 3, d)), rank)
print("rank%d carry err x %.2e dx %.2e | gathered end0 norm %.3g" % (rank, np.abs(xin - x0[0]).max(), np.abs(dxin - dx0[0]).max(), np.abs(ends[0]).max()), flush=True)
l1, g1 = obj1(Yb)
print("rank%d one-pass total loss %.10g (sum of blocks expected)" % (rank, l1), flush=True)
dist.barrier()
dist.destroy_process_group()
