"""Developer experiment (GPU): kernel time of the cfg-3 launch over several pose seeds (the batches of different ranks)."""
import os, sys, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import obca as om, scenario as sc
import torch
B = 8192
t = lambda v: None if v is None else torch.as_tensor(v, dtype=torch.float64, device='cuda').contiguous()
res = []
for seed in [0] + [977 * (r + 1) for r in range(7)]:
    b = sc.make_batch(3, B, pose_seed=None if seed == 0 else seed)
    prm, a = sc.batch_arrays(b)
    s = om.BatchSolver(prm, a['edge_ptr'], B)
    dv = {k: t(a[k]) for k in ('x0', 'u0', 'xref', 'A', 'b0', 'db', 'T_max', 'term')}
    out = s.alloc_outputs(B, 'cuda')
    ms = []
    for i in range(5):
        s.solve(dv['x0'], dv['u0'], dv['xref'], dv['A'], dv['b0'], dv['db'], T_max=dv['T_max'], term=dv['term'], out=out)
        torch.cuda.synchronize(); ms.append(s.last_kernel_ms())
    it = out['iters'].cpu().numpy()
    res.append(min(ms[1:]))
    print('seed %5d kernel ms best %.2f median %.2f  iters sum %d max %d' % (seed, min(ms[1:]), float(np.median(ms[1:])), it.sum(), it.max()), flush=True)
    s.close()
print('mean %.2f max %.2f min %.2f' % (np.mean(res), np.max(res), np.min(res)))
