"""Developer experiment: does phase alignment of the co-resident blocks matter (instruction cache)?  A batch of 8,192
copies of ONE instance keeps the three blocks of an SM in lockstep for the whole launch; compare the time per
interior-point iteration with the ordinary batch."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import obca as om, scenario as sc
import torch
B = 8192
b = sc.make_batch(3, B)
prm, a = sc.batch_arrays(b)
s = om.BatchSolver(prm, a['edge_ptr'], B)
t = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.float64, device='cuda').contiguous()
def run(arr, tag):
    dv = {k: t(arr[k]) for k in ('x0', 'u0', 'xref', 'A', 'b0', 'db', 'T_max', 'term')}
    out = s.alloc_outputs(B, 'cuda')
    for i in range(3):
        s.solve(dv['x0'], dv['u0'], dv['xref'], dv['A'], dv['b0'], dv['db'], T_max=dv['T_max'], term=dv['term'], out=out)
        torch.cuda.synchronize()
    it = out['iters'].cpu().numpy()
    ms = s.last_kernel_ms()
    print('%-28s kernel %.2f ms, iterations %d (mean %.1f) -> %.1f ns per iteration' % (tag, ms, it.sum(), it.mean(), ms * 1e6 / it.sum()))
run(a, 'ordinary batch')
for j in (0, 1, 2, 3):
    rep = dict(a)
    for k in ('x0', 'u0', 'xref', 'T_max'):
        rep[k] = np.repeat(a[k][j:j + 1], B, 0)
    run(rep, 'copies of instance %d' % j)
