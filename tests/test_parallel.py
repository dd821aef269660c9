"""Host-side multi-GPU logic on CPU: two ranks under gloo, the oracle standing in for the CUDA evaluator.

(The oracle is only the stand-in evaluator and the checker here - test infrastructure, as allowed.)"""
import os
import socket

import numpy as np
import pytest

from conftest import rel_err
from multioutputihgp_b200.parallel import shard_bounds


def test_shard_bounds_cover_everything_once():
    for N in (0, 1, 2, 5, 8, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            cover = []
            for r in range(world):
                lo, hi = shard_bounds(N, world, r)
                assert 0 <= lo <= hi <= N
                cover += list(range(lo, hi))
            assert cover == list(range(N)), (N, world)
            sizes = [shard_bounds(N, world, r)[1] - shard_bounds(N, world, r)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, N, out_dir):
    import torch.distributed as dist
    from multioutputihgp_b200.parallel import ShardedObjective, sharded_nll
    from oracle.binding import OracleMOIHGP
    from oracle.gen_golden import make_data, make_params
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(77)
    p, L, T = 6, 3, 40
    params = make_params(rng, p, L, "Matern32")
    Y = np.stack([make_data(rng, p, L, T) for _ in range(N)]) if N else np.zeros((0, T, p))
    o = OracleMOIHGP(0.1, p, L, "Matern32", threading=True)
    o.update(params)
    lo, hi = shard_bounds(N, world, rank)
    obj = ShardedObjective(lambda Yl: o.objective(Yl)[:2], o.num_param)
    loss, grad = obj(Y[lo:hi])
    nll_local = o.filter_smoother_nll(Y[lo:hi], smoother_mode=-1, want_states=False)["nll"] if hi > lo else np.zeros(0)
    total = sharded_nll(nll_local)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), loss=loss, grad=grad, total=total, lo=lo, hi=hi)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("N", [5, 1])
def test_sharded_objective_two_ranks_gloo(tmp_path, N):
    import torch.multiprocessing as mp
    from oracle.binding import OracleMOIHGP
    from oracle.gen_golden import make_data, make_params
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), N, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(77)
    p, L, T = 6, 3, 40
    params = make_params(rng, p, L, "Matern32")
    Y = np.stack([make_data(rng, p, L, T) for _ in range(N)])
    o = OracleMOIHGP(0.1, p, L, "Matern32", threading=True)
    o.update(params)
    loss, grad = o.objective(Y)[:2]
    nll = o.filter_smoother_nll(Y, smoother_mode=-1, want_states=False)["nll"].sum()
    outs = [np.load(os.path.join(str(tmp_path), "rank%d.npz" % r)) for r in range(world)]
    for z in outs:                                   # every rank holds the identical reduced result
        assert abs(float(z["loss"]) - loss) <= 1e-12 * abs(loss)
        assert rel_err(z["grad"], grad) < 1e-12
        assert abs(float(z["total"]) - nll) <= 1e-12 * abs(nll)
    assert float(outs[0]["loss"]) == float(outs[1]["loss"]) and np.array_equal(outs[0]["grad"], outs[1]["grad"])
    assert int(outs[0]["lo"]) == 0 and int(outs[1]["hi"]) == N


def test_block_transition_matches_the_step_by_step_recursion():
    from multioutputihgp_b200.parallel import block_transition
    rng = np.random.default_rng(3)
    for d in (2, 3):
        M = 0.6 * rng.standard_normal((d, d)) / np.sqrt(d)
        dM = [rng.standard_normal((d, d)) for _ in range(3)]
        for n in (0, 1, 2, 7, 64, 1000):
            P, E = block_transition(M, dM, n)
            Pr, Er = np.eye(d), [np.zeros((d, d)) for _ in range(3)]
            for _ in range(n):
                Er = [M @ e + dm @ Pr for e, dm in zip(Er, dM)]
                Pr = M @ Pr
            assert np.allclose(P, Pr, rtol=1e-12, atol=1e-300)
            for a, b in zip(E, Er):
                assert np.allclose(a, b, rtol=1e-11, atol=1e-300)


def _time_worker(rank, world, port, T, out_dir):
    import torch.distributed as dist
    from multioutputihgp_b200.parallel import TimeShardedObjective, time_block_bounds
    from oracle.binding import OracleMOIHGP
    from oracle.gen_golden import make_data, make_params
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(78)
    p, L = 6, 3
    params = make_params(rng, p, L, "Matern52")
    Y = make_data(rng, p, L, T)
    o = OracleMOIHGP(0.1, p, L, "Matern52", threading=True)
    o.update(params)
    consts = [o.ihgp_consts(l) for l in range(L)]
    bounds = [time_block_bounds(T, world, r) for r in range(world)]

    def evaluate(Yb, x0, dx0):
        loss, grad, xT, dxT = o.objective(Yb[None], x0=x0[None], dx0=dx0[None])
        return loss, grad, xT[0], dxT[0]
    obj = TimeShardedObjective(evaluate, consts, [b[1] - b[0] for b in bounds], o.num_param)
    t0, t1 = bounds[rank]
    loss, grad = obj(Y[t0:t1])
    np.savez(os.path.join(out_dir, "trank%d.npz" % rank), loss=loss, grad=grad)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,T", [(2, 301), (3, 100)])
def test_time_sharded_objective_gloo(tmp_path, world, T):
    """One sequence split in time over the ranks: carries exchanged by all-gather, [loss, grad] all-reduced;
    equals the single-process evaluation of the whole sequence."""
    import torch.multiprocessing as mp
    from oracle.binding import OracleMOIHGP
    from oracle.gen_golden import make_data, make_params
    mp.spawn(_time_worker, args=(world, _free_port(), T, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(78)
    p, L = 6, 3
    params = make_params(rng, p, L, "Matern52")
    Y = make_data(rng, p, L, T)
    o = OracleMOIHGP(0.1, p, L, "Matern52", threading=True)
    o.update(params)
    loss, grad = o.objective(Y[None])[:2]
    for r in range(world):
        z = np.load(os.path.join(str(tmp_path), "trank%d.npz" % r))
        assert abs(float(z["loss"]) - loss) <= 1e-11 * abs(loss)
        assert rel_err(z["grad"], grad) < 1e-10
