"""Development aid: thirty full-size launches against the C oracle (intermittent-hazard hunt).  python tools/repro_c4.py [keyframes]"""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from slam_experiments_b200 import _native as nat, synth
from oracle import c_oracle as co
nkf=int(sys.argv[1]) if len(sys.argv)>1 else 4096
q,t=synth.keyframe_database(nkf)
print('launch', nat.describe_launch(2000,t.shape[0],1,'f4'))
exp=co.knn2_keys(q,t)
td=torch.from_numpy(t).cuda(); qd=torch.from_numpy(q).cuda()
tp=nat.prepare(td,variant='f4'); qp=nat.prepare(qd,variant='f4')
for it in range(30):
    got=nat.knn2_keys_prepared(qp,2000,tp,t.shape[0],0,variant='f4').cpu().numpy().view(np.uint64)
    bad=np.flatnonzero((got!=exp).any(axis=1))
    if bad.size:
        r=bad[0]
        print(it,'bad rows',bad.size,'first',bad[:6],'got',[(int(k>>32),int(k&0xffffffff)) for k in got[r]],'exp',[(int(k>>32),int(k&0xffffffff)) for k in exp[r]], 'rows%256', (bad%256)[:8], 'blk', (bad//256)[:8])
print('done')
