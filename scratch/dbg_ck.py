import os, sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import gzip, shutil, tempfile
import lambda_b200, orc
from lambda_b200._abi import MATCH_DT
tmp = tempfile.mkdtemp()
g='/root/repo/tests/golden/prot_flat/'
with gzip.open(g+'db.lba.gz','rb') as fi, open(tmp+'/db.lba','wb') as fo: shutil.copyfileobj(fi, fo)
path = tmp+'/db.lba'
o = orc.Oracle(path); ix = lambda_b200.Index.load(path)
ids, data, offs = lambda_b200.read_fasta(g+'q.fasta')
res = lambda_b200.encode(data, 0)
s = lambda_b200.Searcher(ix,"protein"); p=o.params(0)
m,_ = o.seed(p,res,offs,2)
win,_ = o.merge(p,res,offs,m)
win = win.astype(MATCH_DT)
_, hc = o.extend(p,res,offs,win,True)
hg,_ = s.extend_trace(res,offs,win)
F=["q_start","q_end","s_start","s_end","score","n_match","n_mismatch","n_gap_open","n_gap_ext","aln_len"]
nbad=0
for i in range(len(win)):
    if any(hg[f][i]!=hc[f][i] for f in F):
        nbad+=1
        if nbad<=12:
            w=win[i]
            print('win',w, 'nq',w['qry_end']-w['qry_start'],'nt',w['subj_end']-w['subj_start'])
            print(' gpu',[int(hg[f][i]) for f in F]); print(' cpu',[int(hc[f][i]) for f in F])
print('bad',nbad,'of',len(win))
