import os, sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import gzip, shutil, tempfile
import lambda_b200, orc
from lambda_b200 import synth
from lambda_b200._abi import MATCH_DT
tmp = tempfile.mkdtemp()
with gzip.open('/root/repo/tests/golden/prot_flat/db.lba.gz','rb') as fi, open(tmp+'/db.lba','wb') as fo: shutil.copyfileobj(fi, fo)
path = tmp+'/db.lba'
o = orc.Oracle(path); ix = lambda_b200.Index.load(path)
rng = np.random.default_rng(7)
db, offs = synth.protein_db(500, seed=101)
lens = np.diff(offs)
res_list=[]; qo=[0]; wins=[]
for qi,L in enumerate([1,2,3,5,7]):
    sid = int(rng.integers(0,len(lens))); src = db[offs[sid]:offs[sid+1]]
    seq = src[:L].copy(); res_list.append(seq); qo.append(qo[-1]+L)
    for k in range(40):
        s2 = int(rng.integers(0,len(lens))); a = int(rng.integers(0,lens[s2])); b = min(int(lens[s2]), a+int(rng.integers(1,6)))
        wins.append((qi,s2,0,L,a,b))
res = lambda_b200.encode(np.concatenate(res_list),0); qo=np.array(qo,np.uint64); win=np.array(wins,dtype=MATCH_DT)
s = lambda_b200.Searcher(ix,"protein"); p=o.params(0)
_, hc = o.extend(p,res,qo,win,True)
hg,_ = s.extend_trace(res,qo,win)
F=["q_start","q_end","s_start","s_end","score","n_match","n_mismatch","n_gap_open","aln_len"]
nbad=0
for i in range(len(win)):
    if any(hg[f][i]!=hc[f][i] for f in F):
        nbad+=1
        if nbad<=8:
            w=win[i]; q=res[int(qo[w['qry_id']]):int(qo[w['qry_id']+1])]
            print('win',w,'query',q)
            print(' gpu',[int(hg[f][i]) for f in F]); print(' cpu',[int(hc[f][i]) for f in F])
print('bad',nbad,'of',len(win))
