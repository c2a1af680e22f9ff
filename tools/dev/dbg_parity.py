import sys, numpy as np
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import regex_fpga_b200 as R
from oracle import oracle_py as O
z=np.load('/root/repo/tests/golden/snort_16.npz'); E,n,lo,hi=z['entries'],int(z['n_states']),z['lo'],z['hi']
ctx=R.Context(0); nfa=ctx.nfa_from_entries(E)
for name,data in (('lo',lo),('hi',hi)):
    M=20000
    got=nfa.scan(data[:M].reshape(1,-1),1,n_steps=M-1,stride=M)
    want=O.b_scan(E,n,data,M-1)
    g=list(zip(got.records['pos'].tolist(),got.records['state'].tolist())); w=list(zip(want['recs']['pos'].tolist(),want['recs']['state'].tolist()))
    print(name,len(g),len(w), 'first missing', [x for x in w if x not in set(g)][:5], 'extra', [x for x in g if x not in set(w)][:5])
# small random nfa check
from nfa_gen import random_nfa, random_streams
rng=np.random.default_rng(7000)
(E2,n2),syms=random_nfa(rng,n_states=40,alphabet=6,p_sticky=0.1)
nf2=ctx.nfa_from_entries(E2,n2); d=random_streams(rng,syms,3,60)
got=nf2.scan(d,3,n_steps=60,stride=60); want=O.b_scan_many(E2,n2,d,3,60,60)
print('rand',got.n_matches,want['n_recs'], nf2.info)
