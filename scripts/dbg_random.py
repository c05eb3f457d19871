import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from test_gpu_random import random_batch, run, orc_endpos
from _util import oracle_reads_from_arrays
from oracle import oracle as orc
from pysamstats_b200 import _cabi, config
ctx=_cabi.get_context()
nb,ref=random_batch(2)
reads=oracle_reads_from_arrays(nb)
hi=int(nb['pos'].max())+900
kw=dict(stepper='samtools', min_baseq=13)
pos,F,_=orc.pileup_packed('coverage',reads,0,hi,truncate=True,**kw)
out,mask=run(ctx,nb,ref,'coverage',0,hi,**kw)
rows=np.flatnonzero(mask); got=out[:,rows].T
bad=np.argwhere(got!=F)
print('n bad', len(bad), bad[:10], [pos[b[0]] for b in bad[:10]])
# tweaked quals from the GPU
dev={k:(torch.from_numpy(np.ascontiguousarray(v).view(np.int16) if k=='flag' else (np.ascontiguousarray(v).view(np.int32) if k=='cigar' else np.ascontiguousarray(v))).cuda() if isinstance(v,np.ndarray) else v) for k,v in nb.items()}
rs=_cabi.make_reads_struct(dev)
qout=torch.empty_like(dev['qual']); pf=torch.empty(nb['n_reads'],dtype=torch.int32,device='cuda')
ctx.overlap_tweak(rs, 2, 8000, qout, pf); ctx.synchronize()
q2=qout.cpu().numpy(); pf=pf.cpu().numpy()
OPS='MIDNSHP=X'
for p in sorted(set(pos[b[0]] for b in bad))[:3]:
    print('--- pos', p, 'got', got[list(pos).index(p)], 'exp', F[list(pos).index(p)])
    for i in range(nb['n_reads']):
        e=orc_endpos(nb,i)
        if nb['pos'][i]<=p<e:
            c=nb['cigar'][nb['cigar_off'][i]:nb['cigar_off'][i+1]]
            cs=''.join('%d%s'%(w>>4,OPS[w&15]) for w in c)
            # qpos
            x=nb['pos'][i]; y=0; qp=None; isdel=False
            for w in c:
                op=w&15; l=int(w>>4)
                if op in (0,7,8):
                    if x<=p<x+l: qp=y+(p-x)
                    x+=l; y+=l
                elif op in (2,3):
                    if x<=p<x+l: isdel=True
                    x+=l
                elif op in (1,4): y+=l
            qo=nb['qual_off'][i]
            print(i, 'pos',nb['pos'][i],'flag',nb['flag'][i],cs,'lseq',nb['l_seq'][i],'tlen',nb['tlen'][i],'prev',nb['name_prev'][i],'pairfirst',pf[i],'qpos',qp,'del',isdel,
                  'rawq', nb['qual'][qo+qp] if qp is not None and qp<nb['l_seq'][i] else None, 'gpuq', q2[qo+qp] if qp is not None and qp<nb['l_seq'][i] else None)
