"""Multi-GPU use of the hot path: one process per GPU, independent sequences sharded over ranks.

The reference has a single sequence and no communication layer; its objective is a plain sum
over observations (moihgp_regression.h:46-47), so N independent sequences shard over ranks with NO data-path
collective.  The only exchange is one fp64 all-reduce per evaluation of the fused buffer ``[loss, grad]``
(``1 + num_param`` doubles; objective) or of the summed NLL (filter pass) - SURVEY.md 8(e).  Every rank ends up
with the identical reduced result, which is what a replicated L-BFGS loop needs.

``torch.distributed`` is plumbing only: NCCL over NVLink on the GPU box, gloo in the CPU tests
(tests/test_parallel.py, world_size 2, with a stand-in evaluator).
"""
import numpy as np


def shard_bounds(num_sequences, world_size, rank):
    """Contiguous, balanced shard [lo, hi) of ``num_sequences`` for ``rank``: the first ``N % world`` ranks get one more.
    A rank may receive an empty shard (N < world)."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad rank %r / world_size %r" % (rank, world_size))
    base, extra = divmod(int(num_sequences), world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _dist():
    import torch.distributed as dist
    return dist


class ShardedObjective(object):
    """loss, grad = sum over ALL ranks' sequences of the objective; call with this rank's shard.

    ``evaluate(Y_local) -> (loss, grad)`` is the local evaluator: ``MOIHGPSequences.objective`` on a GPU box.
    An empty local shard contributes zeros.  The reduction is one all-reduce of a single fp64 buffer."""

    def __init__(self, evaluate, num_param, group=None, device=None):
        self.evaluate = evaluate
        self.num_param = int(num_param)
        self.group = group
        self.device = device

    def __call__(self, Y_local):
        import torch
        buf = np.zeros(1 + self.num_param)
        if Y_local is not None and len(Y_local) > 0:
            loss, grad = self.evaluate(Y_local)
            buf[0] = loss
            buf[1:] = grad
        dist = _dist()
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            t = torch.from_numpy(buf)
            if self.device is not None:
                t = t.to(self.device)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            buf = t.cpu().numpy()
        return float(buf[0]), buf[1:].copy()


class ShardedDeviceObjective(object):
    """Device-resident variant for the GPU box: the model writes ``[loss, pad, grad]`` into one torch CUDA buffer and
    NCCL all-reduces it in place on the same stream (no host round trip before the reduction)."""

    def __init__(self, model, group=None):
        import torch
        self.model = model
        self.group = group
        self.buf = torch.zeros(2 + model.num_param, dtype=torch.float64, device=torch.device("cuda", torch.cuda.current_device()))

    def __call__(self, Y_local_device):
        dist = _dist()
        self.buf.zero_()
        if Y_local_device is not None and Y_local_device.shape[0] > 0:
            self.model.objective_device(Y_local_device, self.buf[0:1], self.buf[2:])
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            dist.all_reduce(self.buf, op=dist.ReduceOp.SUM, group=self.group)
        host = self.buf.cpu().numpy()
        return float(host[0]), host[2:].copy()


def sharded_nll(local_nll, group=None, device=None):
    """Total NLL over all ranks' sequences from this rank's per-sequence values (filter + smoother + NLL pass)."""
    import torch
    t = torch.tensor([float(np.sum(local_nll)) if local_nll is not None and len(local_nll) else 0.0], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist = _dist()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return float(t.item())
