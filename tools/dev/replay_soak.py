"""CPU-only soak (dev tool): the execution image interpreted on the host exactly as the lane kernel does
(tests/image_replay.cpp) against oracle B over many random unanchored NFAs.  python tools/dev/replay_soak.py N"""
import sys, os, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import regex_fpga_b200 as R
from oracle import oracle_py as O
from nfa_gen import random_nfa, random_streams
CSRC=os.path.join(ROOT, 'regex_fpga_b200', 'csrc')
subprocess.run(['g++','-O2','-std=c++17','-I',CSRC,os.path.join(ROOT, 'tests', 'image_replay.cpp')]+[f'{CSRC}/{f}' for f in ('image.cpp','nfa.cpp','formats.cpp')]+['-o','/tmp/rfb_image_replay'],check=True)
bad=0; n_abs=0
for i in range(int(sys.argv[1])):
    rng=np.random.default_rng(900000+int(sys.argv[2] if len(sys.argv)>2 else 0)+i)
    (E,n),syms=random_nfa(rng,n_states=int(rng.integers(5,400)),alphabet=int(rng.integers(2,16)),p_sticky=float(rng.choice([0.05,0.2,0.4])),max_fanout=int(rng.integers(1,4)),unanchored=True)
    L=int(rng.integers(10,200)); ns=int(rng.integers(2,30))
    data=random_streams(rng,syms,ns,L,p_alpha=float(rng.choice([0.7,0.95])))
    R.coe_write('/tmp/rfb_rs.coe',E,style=0); np.ascontiguousarray(data).tofile('/tmp/rfb_rs.bin')
    p=subprocess.run(['/tmp/rfb_image_replay','/tmp/rfb_rs.coe','/tmp/rfb_rs.bin',str(ns),str(L),str(L),str(int(rng.choice([64,2000,32766])))],capture_output=True,text=True)
    if p.returncode!=0:
        continue   # image not ok (too large etc.)
    got=[tuple(map(int,l.split())) for l in p.stdout.splitlines()]
    w=O.b_scan_many(E,n,data,ns,L,L,cap=1<<20)['recs']
    want=list(zip(w['stream'].tolist(),w['pos'].tolist(),w['state'].tolist()))
    if got!=want:
        bad+=1; print('MISMATCH',i,flush=True)
print('cases',sys.argv[1],'mismatches',bad)
