"""2-rank probe of the sharded cfg3 path: non-finite values per rank before the gather and on rank 0 after it."""
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from amico_b200 import synth, models as amx_models, parallel
from amico_b200.plan import Plan
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
cfg = 3
n_total = int(np.prod(synth.CONFIGS[cfg][1]))
i0, i1 = parallel.shard_bounds(n_total, world, rank)
P = parallel.broadcast_problem(cfg, 8, rank, world, dev)
mdl = amx_models.NODDI(); mdl.set_solver()
l1, l2 = mdl.solver_params["lambda1"], mdl.solver_params["lambda2"]
plan = Plan("NODDI", P.KERNELS, P.htable, P.params, dwi_idx=P.scheme.dwi_idx, device=dev.index)
y, d = synth.make_voxels_torch("NODDI", P.KERNELS, P.htable, i1 - i0, 20251017 + cfg + 1000 * rank, dev)
est = torch.full((i1 - i0, 3), float("nan"), dtype=torch.float64, device=dev)
for step in range(4):
    dist.barrier(); torch.cuda.synchronize()
    plan.fit(y, d, l1, l2, out=est)
    torch.cuda.synchronize()
    nb = int((~torch.isfinite(est)).sum().item())
    maps = parallel.gather_maps(est.to(torch.float32), rank, world)
    torch.cuda.synchronize()
    ng = int((~torch.isfinite(maps)).sum().item()) if rank == 0 else -1
    print(json.dumps({"rank": rank, "step": step, "nonfinite_before_gather": nb, "after_gather_rank0": ng, "y_nonfinite": int((~torch.isfinite(y)).sum().item()),
                      "counters": plan.last_counters()}), flush=True)
    if nb:
        bad = torch.nonzero(~torch.isfinite(est).all(dim=1)).flatten()
        i = int(bad[0]); print("rank", rank, "voxel", i, est[i].tolist(), "n_bad_vox", len(bad), "first", bad[:6].tolist(), "last", bad[-3:].tolist(), flush=True)
plan.close()
dist.destroy_process_group()
