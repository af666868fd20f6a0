"""Developer experiment (GPU): how much of the work-queue tail does a longest-first order recover?  The first-pass launch is
run over a device-side work list in (a) the given order, (b) random orders, (c) descending TRUE iteration count (upper
bound), (d) descending difficulty estimate computed from the inputs alone (reference poses in collision, heading jumps,
start heading error - full separating-axis clearance of the reference window)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import obca as om, scenario as sc
import torch
argv = sys.argv; sys.argv = ["audit_cfg5.py"]
import audit_cfg5 as au
sys.argv = argv
B = 8192
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 0
b = sc.make_batch(3, B, pose_seed=None if seed == 0 else seed)
prm, a = sc.batch_arrays(b)
s = om.BatchSolver(prm, a['edge_ptr'], B)
t = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.float64, device='cuda').contiguous()
dv = {k: t(a[k]) for k in ('x0', 'u0', 'xref', 'A', 'b0', 'db', 'T_max', 'term')}
out = s.alloc_outputs(B, 'cuda')
cnt = torch.tensor([B], dtype=torch.int32, device='cuda')
def run(order, tag, reps=4):
    idx = torch.as_tensor(np.ascontiguousarray(order, dtype=np.int32), device='cuda')
    ms = []
    for i in range(reps):
        s.solve(dv['x0'], dv['u0'], dv['xref'], dv['A'], dv['b0'], dv['db'], T_max=dv['T_max'], term=dv['term'], out=out, index=idx, count=cnt)
        torch.cuda.synchronize(); ms.append(s.last_kernel_ms())
    print('%-34s kernel ms %s -> best %.2f' % (tag, ' '.join('%.2f' % m for m in ms[1:]), min(ms[1:])), flush=True)
    return out['iters'].cpu().numpy().astype(float)
it = run(np.arange(B), 'given order')
st = out['status'].cpu().numpy()
print('status', {int(v): int((st == v).sum()) for v in np.unique(st)}, 'iters sum %d max %d, over 40: %s' % (it.sum(), it.max(), sorted(it[it > 40].astype(int).tolist())), flush=True)
rng = np.random.default_rng(1)
for r in range(3): run(rng.permutation(B), 'random order %d' % r)
run(np.argsort(-it, kind='stable'), 'true iterations, longest first')
# difficulty estimate from the inputs
xr = a['xref']; x0 = a['x0']; N = xr.shape[1] - 1
polys = [np.asarray(p[:4], float) for p in b.polygons]
def sep(P):
    C = au.corners(P, b.ego); th = P[:, 2]
    ax_e = np.stack([np.stack([np.cos(th), np.sin(th)], -1), np.stack([-np.sin(th), np.cos(th)], -1)], 1)
    best = np.full(len(P), np.inf)
    for Q in polys:
        s_ = np.full(len(P), -np.inf)
        e = np.roll(Q, -1, 0) - Q; nrm = np.stack([e[:, 1], -e[:, 0]], -1); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
        for ax in nrm:
            p = C @ ax; q = Q @ ax; s_ = np.maximum(s_, np.maximum(q.min() - p.max(1), p.min(1) - q.max()))
        for j in range(2):
            ax = ax_e[:, j]; p = np.einsum('nij,nj->ni', C, ax); q = Q @ ax.T
            s_ = np.maximum(s_, np.maximum(q.min(0) - p.max(1), p.min(1) - q.max(0)))
        best = np.minimum(best, s_)
    return best
S = np.stack([sep(xr[:, k]) for k in range(N + 1)], 1)
dth = np.abs(np.diff(xr[:, :, 2], axis=1))
F = np.stack([np.ones(B), S.min(1), (S < 0.5).sum(1), (S < 0.0).sum(1), (S < 1.5).sum(1), dth.sum(1), np.abs(x0[:, 2] - xr[:, 0, 2]), dth.max(1), x0[:, 0], np.minimum(S[:, 0], 3)], 1)
w = np.array([13.41, 0.37, -0.17, 2.29, 0.17, -0.38, 1.12, 5.34, 0.03, 0.35])
pred = F @ w
print('correlation of the estimate with the iteration count: %.2f' % np.corrcoef(pred, it)[0, 1])
run(np.argsort(-pred, kind='stable'), 'estimate (10 features), longest first')
w3 = np.array([15.71, 1.83, 5.1, 1.37]); pred3 = F[:, [0, 3, 7, 6]] @ w3
run(np.argsort(-pred3, kind='stable'), 'estimate (3 features), longest first')
