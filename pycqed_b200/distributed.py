"""Multi-GPU sweeps: one process per GPU, sweep points sharded, results gathered once.

Sweep points are independent (the reference's loop carries no state between iterations,
pyscqed/numerical_system.py:944-965), so the flattened C-order grid (parameters.py:876-884) is cut
into contiguous blocks, one per rank, every rank runs its block through its own
:class:`~pycqed_b200.engine.Engine`, and a single ``all_gather`` (NCCL over NVLink on GPUs, gloo in
the CPU tests) reassembles eigenvalues / scalar outputs / strips.  Eigenvectors stay sharded unless
``gather_vectors=True`` (n = 3375, k = 10 is 540 KB per point).  There is no collective inside the
solve.
"""
from __future__ import annotations

import numpy as np


def shard_bounds(n_pts: int, world: int, rank: int):
    """[lo, hi) of rank's contiguous block: ceil(n/world) points per rank, last ranks may be short."""
    per = -(-n_pts // world)
    lo = min(n_pts, rank * per)
    return lo, min(n_pts, lo + per)


def _all_gather_rows(local, n_total, world, device):
    """Gather row-blocks of unequal length into one [n_total, ...] array (pads to the common size)."""
    import torch
    import torch.distributed as dist
    per = -(-n_total // world)
    t = torch.as_tensor(np.ascontiguousarray(local))
    pad = torch.zeros((per,) + tuple(t.shape[1:]), dtype=t.dtype)
    pad[: t.shape[0]] = t
    pad = pad.to(device)
    full = torch.empty((world * per,) + tuple(t.shape[1:]), dtype=t.dtype, device=device)
    dist.all_gather_into_tensor(full, pad)
    return full[:n_total].cpu().numpy()


def sweep_sharded(make_engine, grid, k, want_vectors=False, resonator=False, gather_vectors=False, device=None,
                  gather_strips=True, **sweep_kw):
    """Run ``grid`` [n_pts, n_swept] across all ranks of the default process group (one process per GPU, launched
    by torchrun; inside ONE process use :class:`pycqed_b200.engine.EnginePool` instead, which is what the drop-in
    ``paramSweep`` does).

    ``make_engine()`` returns this rank's engine, or the engine itself may be passed (anything with
    ``.sweep(params, k, want_vectors=, resonator=)``).  Every rank returns the full gathered eigenvalues, scalars
    and info; the resonator strips are gathered too unless ``gather_strips=False`` (then ``Erwa`` is the local
    shard, like ``V`` unless ``gather_vectors``).  ``sweep_kw`` goes to ``engine.sweep`` (``out=``,
    ``real_vectors=``, ``vectors_on_device=`` ...).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    n_pts = len(grid)
    lo, hi = shard_bounds(n_pts, world, rank)
    eng = make_engine() if callable(make_engine) else make_engine
    local = eng.sweep(grid[lo:hi], k, want_vectors=want_vectors, resonator=resonator, **sweep_kw) if hi > lo else None
    if world == 1:
        return {"E": local.E, "V": local.V, "scalars": local.scalars, "Erwa": local.Erwa, "info": local.info,
                "shard": (lo, hi), "local": local}
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    n_loc = hi - lo

    def rows(a, shape_tail, dtype):
        if a is None or n_loc == 0:
            a = np.zeros((0,) + shape_tail, dtype=dtype)
        return _all_gather_rows(a, n_pts, world, device)

    n = getattr(eng, "n", None)
    out = {"shard": (lo, hi), "local": local}
    # eigenvalues, scalar outputs and info travel as ONE [points, k + n_scalar + 1] block: one collective and one
    # device -> host copy per sweep (info values are small integers, exact in float64)
    names = list(eng.plan.program.scalar_names)
    width = k + len(names) + 1
    blk = np.zeros((n_loc, width), dtype=np.float64)
    if local is not None and n_loc:
        blk[:, :k] = local.E
        for i, nm in enumerate(names):
            blk[:, k + i] = local.scalars[nm]
        blk[:, width - 1] = local.info
    full = rows(blk, (width,), np.float64)
    out["E"] = np.ascontiguousarray(full[:, :k])
    out["scalars"] = {nm: np.ascontiguousarray(full[:, k + i]) for i, nm in enumerate(names)}
    out["info"] = full[:, width - 1].astype(np.int32)
    if resonator and not gather_strips:
        out["Erwa"] = local.Erwa if local else None
    elif resonator:
        ns = eng.plan.post["resonator"]["nmax"] + 1 + k
        out["Erwa"] = rows(local.Erwa if local else None, (ns, k), np.float64)
    else:
        out["Erwa"] = None
    if want_vectors and gather_vectors:
        v = local.V if local else None
        vr = rows(None if v is None else np.ascontiguousarray(v).view(np.float64).reshape(n_loc, k, 2 * n), (k, 2 * n), np.float64)
        out["V"] = vr.reshape(n_pts, k, n, 2).view(np.complex128).reshape(n_pts, k, n)
    else:
        out["V"] = local.V if local else None
    return out
