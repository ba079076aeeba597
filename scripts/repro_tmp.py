import sys, numpy as np
sys.path.insert(0,'/root/repo')
from rchol_b200 import capi, problems, producer
A = problems.laplace_3d(64); f = producer.factor(*A, threads=8)
for mode, thr, bo, npr in ((1, 512, 2, 1), (1, 512, 2, 2), (1, 512, 2, 4), (1, 512, 0, 1), (1, 512, 0, 2), (1, 512, 0, 4), (0, 512, 0, 2), (0, 512, 0, 4)):
    s=capi.Solver(0, chain_threads=thr, chain_mode=mode, dbg=bo, producers=npr)
    s.set_factor(f.rowPtr,f.colIdx,f.val,f.part)
    gs = s.groups(0)
    out=[]
    for g in range(len(gs)):
        out.append('%d:%dx%d=%.3f' % (g, gs[g]['blocks'], gs[g]['max_rows'], s.time_group(0, g, 0, 3)))
    print('mode', mode, 'thr', thr, 'dbg', bo, 'np', npr, ' '.join(out))
    s.close()
