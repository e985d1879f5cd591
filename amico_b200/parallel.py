"""Multi-GPU plumbing: one process per GPU, voxels sharded in contiguous slabs.

Every voxel is independent (no halo, no reduction), so the only communication is
  * one broadcast of the read-only tables (KERNELS, hash table) from rank 0, and
  * one gather of the output maps,
both through ``torch.distributed`` (NCCL over NVLink on the GPU box, gloo in the CPU tests).  The
reference's equivalent is the contiguous chunking of ``BaseModel.fit`` over host threads
(``amico/models.pyx:204-211``).
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n, world, rank):
    """Contiguous slab [i0, i1) of rank ``rank``: like the reference's chunks (``n // nthreads`` per chunk, the
    last one takes the remainder, amico/models.pyx:205-211) but also valid for n < world."""
    if world <= 1:
        return 0, n
    c = n // world
    i0 = rank * c
    i1 = n if rank == world - 1 else (rank + 1) * c
    return i0, i1


def _dist():
    import torch.distributed as dist
    return dist


def broadcast_arrays(arrays, src, rank, world, device):
    """Broadcast a dict of numpy arrays from ``src``; returns the dict on every rank."""
    if world <= 1:
        return arrays
    import torch
    dist = _dist()
    meta = [None]
    if rank == src:
        meta[0] = [(k, tuple(np.asarray(v).shape), np.asarray(v).dtype.str, np.asarray(v).flags.f_contiguous and np.asarray(v).ndim > 1)
                   for k, v in arrays.items()]
    dist.broadcast_object_list(meta, src=src)
    out = {}
    for k, shape, dt, fortran in meta[0]:
        if rank == src:
            a = np.asarray(arrays[k])
            a = np.asfortranarray(a) if fortran else np.ascontiguousarray(a)
            flat = np.frombuffer(a.tobytes(order="A"), dtype=np.uint8).copy()
        else:
            flat = np.empty(int(np.prod(shape)) * np.dtype(dt).itemsize, dtype=np.uint8)
        t = torch.from_numpy(flat).to(device)
        dist.broadcast(t, src=src)
        b = t.cpu().numpy().view(np.dtype(dt))
        out[k] = b.reshape(shape, order="F" if fortran else "C")
    return out


def broadcast_problem(cfg, n_vox, rank, world, device, model=None, ndirs=500):
    """Rank 0 synthesises the protocol tables (scheme, LUT directions, hash table, KERNELS) and broadcasts them;
    every rank then draws its own seeded voxels (weak scaling: ``n_vox`` per rank)."""
    from . import synth
    model = model or synth.CONFIGS[cfg][0]
    scheme = synth.make_scheme(cfg)
    if model == "SANDI":
        scheme = synth.directional_average_scheme(scheme)
    p = dict(synth.default_params(model))
    if rank == 0:
        lut = synth.lut_directions(ndirs)
        ht = synth.build_htable(lut)
        K, p = synth.make_kernels(model, scheme, lut)
        payload = {"__lut": lut, "__htable": ht}
        payload.update({k: v for k, v in K.items() if k != "model"})
    else:
        payload = None
    payload = broadcast_arrays(payload, 0, rank, world, device)
    lut, ht = payload.pop("__lut"), payload.pop("__htable")
    K = dict(payload)
    K["model"] = "FreeWater" if model.startswith("FreeWater") else model
    y, dirs = synth.make_voxels(model, K, ht, n_vox, 20251017 + cfg + 1000 * rank)
    return synth.Problem(cfg, model, scheme, lut, ht, K, p, y, dirs)


def gather_maps(est, rank, world, dst=0):
    """Gather per-rank map tensors (n_vox_r, n_maps) on ``dst`` in rank order; returns the concatenation there."""
    if world <= 1:
        return est
    import torch
    dist = _dist()
    on_cuda = est.is_cuda
    if on_cuda and dist.get_backend() == "gloo":  # gloo cannot gather CUDA tensors (CPU tests, two ranks on one GPU)
        out = gather_maps(est.cpu(), rank, world, dst)
        return None if out is None else out.to(est.device)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(est.shape[0]))
    if len(set(sizes)) == 1:
        out = torch.empty((world * sizes[0],) + tuple(est.shape[1:]), dtype=est.dtype, device=est.device) if rank == dst else None
        parts = list(out.split(sizes[0])) if rank == dst else None
        dist.gather(est.contiguous(), parts, dst=dst)
        return out
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(est.shape[1:]), dtype=est.dtype, device=est.device)
    pad[: est.shape[0]] = est
    parts = [torch.empty_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, parts, dst=dst)
    if rank != dst:
        return None
    return torch.cat([p[:s] for p, s in zip(parts, sizes)])


def fit_sharded(plan_factory, y, dirs, lambda1, lambda2, rank, world, **kw):
    """Fit rank ``rank``'s slab of a volume every rank holds (host arrays); returns (i0, i1, result dict)."""
    i0, i1 = shard_bounds(y.shape[0], world, rank)
    plan = plan_factory()
    d = None if dirs is None else np.ascontiguousarray(dirs[i0:i1], dtype=np.float64)
    res = plan.fit(np.ascontiguousarray(y[i0:i1]), d, lambda1, lambda2, **kw)
    return i0, i1, res
