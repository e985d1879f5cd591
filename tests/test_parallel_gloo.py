"""world_size-2 gloo test of the multi-GPU host logic (table broadcast, voxel sharding, map gather) on CPU."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys, numpy as np, torch, torch.distributed as dist
    sys.path.insert(0, %r)
    from amico_b200 import parallel, synth
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cpu")
    P = parallel.broadcast_problem(1, 64, rank, world, dev)
    Q = synth.make_problem(1, n_vox=8)                       # what rank 0 built, rebuilt locally
    for k in ("D", "CSF"):
        assert np.array_equal(P.KERNELS[k], Q.KERNELS[k]) and P.KERNELS[k].dtype == Q.KERNELS[k].dtype, k
    assert np.array_equal(P.htable, Q.htable) and P.KERNELS["model"] == "FreeWater"
    # Fortran-ordered float64 tables survive the broadcast (SANDI 'signal')
    S = parallel.broadcast_problem(4, 16, rank, world, dev)
    S0 = synth.make_problem(4, n_vox=4)
    assert np.array_equal(S.KERNELS["signal"], S0.KERNELS["signal"]) and S.KERNELS["signal"].flags.f_contiguous
    # each rank draws different voxels (weak scaling); slabs of a common volume tile it exactly
    ys = [None] * world
    dist.all_gather_object(ys, float(P.y.sum()))
    assert len(set(ys)) == world
    i0, i1 = parallel.shard_bounds(1001, world, rank)
    est = torch.full((i1 - i0, 3), float(rank), dtype=torch.float64)
    est[:, 0] = torch.arange(i0, i1, dtype=torch.float64)
    maps = parallel.gather_maps(est, rank, world)
    if rank == 0:
        assert maps.shape == (1001, 3) and torch.equal(maps[:, 0], torch.arange(1001, dtype=torch.float64))
        assert float(maps[-1, 1]) == world - 1
    else:
        assert maps is None
    even = parallel.gather_maps(torch.full((10, 2), float(rank)), rank, world)
    if rank == 0:
        assert even.shape == (10 * world, 2) and float(even[10, 0]) == 1.0
    dist.barrier()
    dist.destroy_process_group()
    print("worker", rank, "ok")
""") % ROOT


def test_broadcast_shard_gather_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(script)]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2
