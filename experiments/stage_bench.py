import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from aligngraph2_b200 import synth
from aligngraph2_b200.mecat2ref import Mecat2RefDevice
n=int(sys.argv[1]); R=int(sys.argv[2])
d=synth.make_batch_torch(5,R,n,10000,device='cuda')
ref=d['ref'].cpu().numpy(); bases=d['bases'].cpu().numpy(); off=d['offsets'].cpu().numpy()
dev=Mecat2RefDevice(0)
t=time.time(); dev.load_reference(ref); torch.cuda.synchronize(); print('ref load',time.time()-t)
t=time.time(); dev.load_reads(bases=bases,offsets=off); print('reads load',time.time()-t)
for it in range(2):
    t=time.time(); dev.build_index(200,0.5,2.0); print('index build',time.time()-t)
for it in range(2):
    t=time.time(); c,nc=dev.seed_candidates(0,10); print('seed',time.time()-t, 'cands/read',nc.mean(), 'reads with 0',(nc==0).sum())
t=time.time(); rec,qa,sa=dev.extend_seed_candidates(10); print('extend from seeds',time.time()-t, len(rec), rec['ok'].sum(), dev.stats())
truth=d['start'].cpu().numpy()
ok=rec[rec['ok']==1]
print('aligned Gbp', (ok['qe']-ok['qb']).sum()/1e9, 'correct locus frac', np.mean(np.abs(ok['sb']-truth[ok['read']])<100))
