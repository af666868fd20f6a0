import os, sys, time, numpy as np
ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from vehicle_motion_planning_with_obstacles_avoidance_using_mpc_b200 import obca as om, scenario as sc
import torch
cfg=int(sys.argv[1]); B=int(sys.argv[2])
b=sc.make_batch(cfg,B)
prm,a=sc.batch_arrays(b, init=int(os.environ.get('OBCA_QUICK_INIT', '2')))   # 786 = WARM | RECOVER
s=om.BatchSolver(prm,a['edge_ptr'],B)
t=lambda v: None if v is None else torch.as_tensor(v,dtype=torch.float64,device='cuda').contiguous()
dv={k:t(a[k]) for k in ('x0','u0','xref','A','b0','db','T_max','term')}
out=s.alloc_outputs(B,'cuda')
for i in range(3):
    s.solve(dv['x0'],dv['u0'],dv['xref'],dv['A'],dv['b0'],dv['db'],T_max=dv['T_max'],term=dv['term'],out=out)
    torch.cuda.synchronize()
    print('cfg',cfg,'B',B,'kernel ms %.2f -> %.0f solves/s'%(s.last_kernel_ms(),B/s.last_kernel_ms()*1e3), 'scratch MB %.1f'%(s.scratch_bytes/1e6))
st=out['status'].cpu().numpy(); it=out['iters'].cpu().numpy()
import hashlib
print('status',{int(v):int((st==v).sum()) for v in np.unique(st)},'iters mean %.1f max %d sum %d'%(it.mean(),it.max(),it.sum()),
      'sha(x,u,obj)',hashlib.sha256(out['x'].cpu().numpy().tobytes()+out['u'].cpu().numpy().tobytes()+out['obj'].cpu().numpy().tobytes()).hexdigest()[:16])
