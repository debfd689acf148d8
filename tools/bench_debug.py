import sys
import numpy as np, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from spotify_recommender_b200 import synth
from spotify_recommender_b200.engine import Engine
from spotify_recommender_b200.sharded import ShardedRecommender
n=10_000_000; B=4096; K=10
eng=Engine(0); sh=ShardedRecommender(eng, n, device=torch.device("cuda",0))
f=synth.features(n); sh.load_shard(f)
qs=[(((np.arange(B,dtype=np.int64)+b*B)*7919+13)%n).astype(np.int32) for b in range(6)]
qd=[torch.from_numpy(q).cuda() for q in qs]
host=[sh.query_by_index(q,K) for q in qs]
# back-to-back device path
outs=[]
for b in range(6):
    oi,os_=sh.query_by_index_dev(qd[b],K)
    outs.append((oi.clone(), os_.clone()))
torch.cuda.synchronize()
for b in range(6):
    d=(outs[b][0].cpu().numpy()!=host[b][0]).any(axis=1)
    print("batch",b,"rows differing (device back-to-back vs host):", int(d.sum()), np.where(d)[0][:8])
# direct engine host API
for b in range(2):
    gi,gs=eng.query_by_index(qs[b],K)
    print("engine host api vs sharded host:", int((gi!=host[b][0]).any(axis=1).sum()))
b=0; d=np.where((outs[b][0].cpu().numpy()!=host[b][0]).any(axis=1))[0]
for r in d[:4]:
    print(r, qs[b][r], outs[b][0][r].cpu().numpy(), host[b][0][r], outs[b][1][r].cpu().numpy(), host[b][1][r])
from oracle_lib import Oracle
o=Oracle()
for b in (0,1):
    sel=np.array([0,1,2,1009,2000,4095])
    wi,ws=o.query_index(f, qs[b][sel], K, threads=o.max_threads)
    gi,gs=eng.query_by_index(qs[b],K)
    print("batch",b,"engine-host vs oracle rows bad:", int((gi[sel]!=wi).any(axis=1).sum()), " sharded-host vs oracle rows bad:", int((host[b][0][sel]!=wi).any(axis=1).sum()), " device vs oracle:", int((outs[b][0].cpu().numpy()[sel]!=wi).any(axis=1).sum()))
    print(qs[b][:3], gi[0], host[b][0][0], wi[0])
