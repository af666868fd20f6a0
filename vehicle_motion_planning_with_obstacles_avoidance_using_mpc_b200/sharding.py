"""Multi-GPU sharding of an OBCA batch: contiguous instance ranges per rank, no exchange during the solve,
ONE collective (a gather of the packed result buffer) at the end.

The reference has no distributed code at all; instances are independent NLPs, so rank r simply owns
``[r*ceil(B/G), min(B, (r+1)*ceil(B/G)))``.  ``PackedOutputs`` lays the eight result arrays of
``obca_b200_solve`` out in one contiguous per-rank buffer (float64 words; status/iters ride in the tail as
int32 pairs) so that the gather is a single ``torch.distributed`` call - NCCL on GPUs, gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(B: int, rank: int, world: int):
    """Contiguous, equally padded split: every rank owns ``per = ceil(B/world)`` slots, the last ranks may be
    short (or empty).  -> (lo, hi, per)"""
    per = -(-B // world)
    lo = min(B, rank * per)
    return lo, min(B, lo + per), per


class PackedOutputs:
    """One contiguous result buffer for ``cap`` instances with typed views into it.

    Word layout (float64 words): x [cap,N+1,3] | u [cap,N,2] | lam [cap,N+1,R] | mu [cap,N+1,4*no] | T [cap] |
    obj [cap] | (status, iters) int32 [cap] each, padded to whole words."""
    FIELDS = ("x", "u", "lam", "mu", "T", "obj")

    def __init__(self, cap, N, rows, n_obs, device="cpu"):
        import torch
        self.cap, self.N, self.rows, self.n_obs = int(cap), int(N), int(rows), int(n_obs)
        self.shapes = dict(x=(cap, N + 1, 3), u=(cap, N, 2), lam=(cap, N + 1, rows), mu=(cap, N + 1, 4 * n_obs),
                           T=(cap,), obj=(cap,))
        self.offsets = {}
        o = 0
        for k in self.FIELDS:
            self.offsets[k] = o
            o += int(np.prod(self.shapes[k]))
        self.int_words = -(-cap // 2)            # cap int32 -> ceil(cap/2) float64 words
        self.offsets["status"] = o; o += self.int_words
        self.offsets["iters"] = o; o += self.int_words
        self.words = o
        self.buf = torch.zeros(self.words, dtype=torch.float64, device=device)
        self.views = self.views_of(self.buf)

    @property
    def nbytes(self):
        return self.words * 8

    def views_of(self, flat):
        """Typed views into a flat float64 buffer of ``self.words`` words (this rank's, or one gathered row)."""
        import torch
        v = {}
        for k in self.FIELDS:
            n = int(np.prod(self.shapes[k]))
            v[k] = flat[self.offsets[k]:self.offsets[k] + n].view(self.shapes[k])
        for k in ("status", "iters"):
            v[k] = flat[self.offsets[k]:self.offsets[k] + self.int_words].view(torch.int32)[:self.cap]
        return v


def gather_packed(packed: PackedOutputs, dst=0, group=None):
    """The single collective of the path: gather every rank's packed buffer to ``dst``.
    Returns a [world, words] tensor on ``dst`` and None elsewhere."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rank == dst:
        full = torch.empty((world, packed.words), dtype=torch.float64, device=packed.buf.device)
        dist.gather(packed.buf, list(full.unbind(0)), dst=dst, group=group)
        return full
    dist.gather(packed.buf, None, dst=dst, group=group)
    return None


def unpack_gathered(packed: PackedOutputs, full, B):
    """[world, words] -> dict of arrays for the first ``B`` instances in global order."""
    import torch
    parts = [packed.views_of(full[r]) for r in range(full.shape[0])]
    out = {}
    for k in PackedOutputs.FIELDS + ("status", "iters"):
        out[k] = torch.cat([p[k] for p in parts], dim=0)[:B]
    return out


def solve_sharded(solver, arrays, B, rank, world, device, dst=0, group=None):
    """Solve this rank's contiguous shard of ``arrays`` (ABI-level host arrays of the whole batch; obstacle
    rows shared) on ``device`` and gather the packed results to ``dst``.
    Returns (dict of torch tensors for the whole batch on dst | None, PackedOutputs of this rank)."""
    import torch
    lo, hi, per = shard_range(B, rank, world)
    p = solver.params
    packed = PackedOutputs(per, p.N, p.rows, p.n_obs, device=device)
    if hi > lo:
        t = lambda v: None if v is None else torch.as_tensor(np.ascontiguousarray(v[lo:hi]), dtype=torch.float64,
                                                             device=device)
        shared = lambda v: None if v is None else torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float64,
                                                                  device=device)
        n = hi - lo
        out = {k: (packed.views[k][:n]) for k in packed.views}
        solver.solve(t(arrays["x0"]), t(arrays["u0"]), t(arrays["xref"]), shared(arrays["A"]), shared(arrays["b0"]),
                     shared(arrays.get("db")), T_max=t(arrays.get("T_max")), term=t(arrays.get("term")), out=out)
    if world == 1:
        return unpack_gathered(packed, packed.buf[None], B), packed
    full = gather_packed(packed, dst=dst, group=group)
    return (unpack_gathered(packed, full, B) if full is not None else None), packed
