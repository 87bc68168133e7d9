"""world_size-2 gloo test of the multi-GPU path: shard the batch, broadcast the scalar plan once,
no tensor traffic; concatenated shards == the single-process result, bitwise."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, outdir):
    for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")):
        sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dpm_solver_b200 import DPM_Solver, model_wrapper, ops, plan as P
    from dpm_solver_b200.distributed import broadcast_plan, shard_batch
    from helpers import product_schedule
    from cases import exact_net, seeded
    from oracle_backend import OracleBackend
    ops.set_backend(OracleBackend())
    ns = product_schedule("sd")
    # rank 1 starts from a deliberately different plan: after the broadcast it must hold rank 0's
    ts = torch.linspace(1., 1e-3, 11) if rank == 0 else torch.linspace(1., 2e-3, 11)
    mine = P.multistep_plan(ns, "dpmsolver++", "dpmsolver", ts, 3, True)
    got = broadcast_plan(mine)
    ref = P.multistep_plan(ns, "dpmsolver++", "dpmsolver", torch.linspace(1., 1e-3, 11), 3, True)
    assert [c.__dict__ for c in got] == [c.__dict__ for c in ref]
    x = seeded((6, 4, 8, 8), 5)
    s = DPM_Solver(model_wrapper(exact_net, ns), ns, plan_broadcast=True)
    y = s.sample(shard_batch(x).contiguous(), steps=10, order=3)
    np.save(os.path.join(outdir, f"y{rank}.npy"), y.numpy())
    s2 = DPM_Solver(model_wrapper(exact_net, ns), ns, algorithm_type="dpmsolver", plan_broadcast=True)
    y2 = s2.sample(shard_batch(x).contiguous(), steps=9, order=3, method="singlestep")
    np.save(os.path.join(outdir, f"z{rank}.npy"), y2.numpy())
    # adaptive: the error estimate is a max over the batch -> one all-reduce(max) per iteration
    s3 = DPM_Solver(model_wrapper(exact_net, ns), ns, algorithm_type="dpmsolver", plan_broadcast=True)
    y3 = s3.sample(shard_batch(x).contiguous(), order=2, method="adaptive", atol=0.05, rtol=0.1)
    np.save(os.path.join(outdir, f"a{rank}.npy"), y3.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shards_equal_single_process(tmp_path, oracle_backend):
    import socket
    with socket.socket() as so:
        so.bind(("127.0.0.1", 0))
        port = so.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    from dpm_solver_b200 import DPM_Solver, model_wrapper
    from helpers import product_schedule
    from cases import exact_net, seeded
    ns = product_schedule("sd")
    x = seeded((6, 4, 8, 8), 5)
    full = DPM_Solver(model_wrapper(exact_net, ns), ns).sample(x, steps=10, order=3).numpy()
    got = np.concatenate([np.load(tmp_path / "y0.npy"), np.load(tmp_path / "y1.npy")])
    np.testing.assert_array_equal(got, full)
    full = DPM_Solver(model_wrapper(exact_net, ns), ns, algorithm_type="dpmsolver").sample(x, steps=9, order=3, method="singlestep").numpy()
    got = np.concatenate([np.load(tmp_path / "z0.npy"), np.load(tmp_path / "z1.npy")])
    np.testing.assert_array_equal(got, full)
    full = DPM_Solver(model_wrapper(exact_net, ns), ns, algorithm_type="dpmsolver").sample(
        x, order=2, method="adaptive", atol=0.05, rtol=0.1).numpy()
    got = np.concatenate([np.load(tmp_path / "a0.npy"), np.load(tmp_path / "a1.npy")])
    np.testing.assert_array_equal(got, full)


def test_shard_bounds():
    from dpm_solver_b200.distributed import pack_plan, shard_bounds, unpack_plan
    from dpm_solver_b200.plan import Coeffs
    assert [shard_bounds(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert [shard_bounds(32768, r, 8) for r in range(8)][-1] == (28672, 32768)
    c = [Coeffs(5, 0.1, -0.2, 0.3, -0.4, 1.5, 2.5, 3.5, 4.5, 5.5, True, 3)]
    assert unpack_plan(pack_plan(c))[0].__dict__ == c[0].__dict__
