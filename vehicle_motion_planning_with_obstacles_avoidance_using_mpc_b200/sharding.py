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

    def __init__(self, cap, N, rows, n_obs, device="cpu", buf=None):
        """``buf``: a flat float64 tensor of ``words`` words to lay the views over instead of a fresh allocation (rank 0 of
        a peer-to-peer gather lets its solver write straight into its row of the gathered buffer)"""
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
        if buf is not None and (buf.numel() != self.words or buf.dtype != torch.float64 or not buf.is_contiguous()):
            raise ValueError("buf must be a contiguous float64 tensor of %d words" % self.words)
        self.buf = torch.zeros(self.words, dtype=torch.float64, device=device) if buf is None else buf
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


def gather_packed(packed: PackedOutputs, dst=0, group=None, out=None):
    """The single collective of the path: gather every rank's packed buffer to ``dst``.
    Returns a [world, words] tensor on ``dst`` (``out`` if given: callers that overlap the gather of one launch with
    the next launch keep two of them) and None elsewhere.  Runs on the current CUDA stream (NCCL)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if rank == dst:
        full = out if out is not None else torch.empty((world, packed.words), dtype=torch.float64, device=packed.buf.device)
        dist.gather(packed.buf, list(full.unbind(0)), dst=dst, group=group)
        return full
    dist.gather(packed.buf, None, dst=dst, group=group)
    return None


class PeerGather:
    """The gather of the packed results as peer-to-peer copies over NVLink: rank 0 owns the [nbuf, world, words] result
    buffer and shares it with the other ranks of the node (CUDA IPC); every rank copies its packed buffer straight into
    its row with one asynchronous device-to-device copy on a side stream.  The copy engines move the data, no SM is
    involved - an NCCL gather issued under the next solver launch has to wait for SMs the persistent solver blocks hold
    and, once running, keeps them spinning on its peers (measured on two B200: the solver launch 23.4 -> 27.7 ms).
    Completion is per rank (its own stream); `wait_all` (a barrier) tells rank 0 that every row has landed.
    Construction is collective and may raise (no peer access, IPC refused): callers fall back to `gather_packed`."""

    def __init__(self, words, nbuf, device, group=None):
        import torch
        import torch.distributed as dist
        from torch.multiprocessing.reductions import reduce_tensor
        self.rank, self.world, self.group = dist.get_rank(group), dist.get_world_size(group), group
        self.side = torch.cuda.Stream(device)
        ok = 1
        box = [None]
        try:
            if self.rank == 0:
                self.full = torch.empty((nbuf, self.world, words), dtype=torch.float64, device=device)
                box = [reduce_tensor(self.full)]
        except Exception:
            ok = 0
        dist.broadcast_object_list(box, src=0, group=group)
        try:
            if self.rank != 0:
                fn, args = box[0]
                self.full = fn(*args)                   # rank 0's buffer, mapped into this process
                self.full[0, self.rank, :1].copy_(torch.zeros(1, dtype=torch.float64, device=device))   # (peer access works?)
                torch.cuda.synchronize(device)
        except Exception:
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        if int(flag[0]) != 1:
            raise RuntimeError("peer-to-peer result buffer not available on this node")

    def push(self, packed: PackedOutputs, j, after_event):
        """row (j, rank) <- this rank's packed buffer, on the side stream, once `after_event` (the solve) has completed.
        Returns the event that marks the copy done (the packed buffer may be reused after it)."""
        import torch
        self.side.wait_event(after_event)
        with torch.cuda.stream(self.side):
            if packed.buf.data_ptr() != self.full[j, self.rank].data_ptr():      # (rank 0 may solve straight into its row)
                self.full[j, self.rank].copy_(packed.buf, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.side)
        return done

    def wait_all(self):
        """every rank's outstanding pushes have landed in rank 0's buffer"""
        import torch.distributed as dist
        self.side.synchronize()
        dist.barrier(group=self.group)


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


def closed_loop_sharded(make_driver, dyn, rank, world, dst=0, group=None, terminal_rule="shipped"):
    """cfg 4 across ranks: scenarios are independent, so rank r runs the closed loops of its contiguous
    shard of ``dyn`` with no per-step communication; the logs are packed into one buffer per rank and gathered once.

    ``make_driver(dyn_shard)`` builds a ``ClosedLoopBatch`` / ``ClosedLoopDevice`` for that shard.  Returns the logs of
    all scenarios in global order on ``dst`` (dict of NumPy arrays, same keys as ``ClosedLoopBatch.run``) and None
    elsewhere."""
    import torch
    import torch.distributed as dist
    dyn = np.asarray(dyn, float)
    B = dyn.shape[0]
    lo, hi, per = shard_range(B, rank, world)
    o = None
    if hi > lo:
        drv = make_driver(dyn[lo:hi])
        try:
            o = drv.run(terminal_rule)
        finally:
            drv.close()
    K = None if o is None else o["traj"].shape[1] - 1
    use_cuda = world > 1 and dist.get_backend(group) == "nccl"
    if world > 1:                      # every rank must know the log length even if its shard is empty
        k_t = torch.tensor([K if K is not None else -1], dtype=torch.int64, device="cuda" if use_cuda else "cpu")
        dist.all_reduce(k_t, op=dist.ReduceOp.MAX, group=group)
        K = int(k_t[0])
    # packed row per scenario: traj (K+1)*3 | mode K | x 3 | u 2 | Ts_opt | steps | failed | reached
    W = (K + 1) * 3 + K + 3 + 2 + 1 + 3
    buf = torch.zeros((per, W), dtype=torch.float64)
    if o is not None:
        n = hi - lo
        cols = [o["traj"].reshape(n, -1), o["mode"].astype(float), o["x"], o["u"], o["Ts_opt"][:, None],
                o["steps"][:, None].astype(float), o["failed"][:, None].astype(float), o["reached"][:, None].astype(float)]
        buf[:n] = torch.as_tensor(np.concatenate(cols, axis=1))
    counts = torch.tensor([0 if o is None else o["solves"], 0 if o is None else o["launches"]], dtype=torch.float64)
    if world == 1:
        full, cnt = buf[None], counts[None]
    else:
        send = torch.cat([buf.reshape(-1), counts])
        if use_cuda:
            send = send.cuda()
        if rank == dst:
            rows = [torch.empty_like(send) for _ in range(world)]
            dist.gather(send, rows, dst=dst, group=group)
            allr = torch.stack(rows).cpu()
            full = allr[:, :-2].reshape(world, per, W); cnt = allr[:, -2:]
        else:
            dist.gather(send, None, dst=dst, group=group)
            return None
    flat = full.reshape(-1, W)[:B].numpy()
    c = 0
    def take(n):
        nonlocal c
        v = flat[:, c:c + n]; c += n
        return v
    out = dict(traj=take((K + 1) * 3).reshape(B, K + 1, 3).copy(), mode=np.rint(take(K)).astype(int), x=take(3).copy(),
               u=take(2).copy(), Ts_opt=take(1)[:, 0].copy(), steps=np.rint(take(1)[:, 0]).astype(int),
               failed=take(1)[:, 0] > 0.5, reached=take(1)[:, 0] > 0.5)
    out["solves"] = int(cnt[:, 0].sum()); out["launches"] = int(cnt[:, 1].sum())
    return out
