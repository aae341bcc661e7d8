import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
from oracle import hmm_oracle as O
from test_properties import _random_model, _random_corpus, _pack
from hmm_training_b200 import engine
N,M,W,S,ltr,zf,seed=16,256,3,1,True,0.0,197
rng=np.random.default_rng(seed)
pi0,A0,B0=_random_model(rng,W,N,M,ltr,zf)
seqs,wos=_random_corpus(rng,W,S,1,60,M)
obs,off=_pack(seqs)
w=2
mine=[seqs[r] for r in range(len(seqs)) if wos[r]==w]
for it in range(1,5):
    with engine.BaumWelch(obs, off, wos, W, N, M) as bw:
        bw.set_params(pi0,A0,B0)
        bw.iterate(it, -1.0, it)
        pi,A,B=bw.params(); th=bw.thin_states(); dg=bw.diagnostics(); fam=bw.kernel_family()
    Ao,Bo,po=O.hmm_training(mine,N=N,M=M,epsilon=-1.0,max_iterations=it,init=(pi0[w],A0[w],B0[w]))
    bad=np.argwhere(~np.isclose(B[w],Bo,rtol=1e-9,atol=1e-30))
    print("it",it,fam,"thin",th,"diag",dg,"B mismatches",bad[:6].tolist(), [ (B[w][tuple(b)], Bo[tuple(b)]) for b in bad[:4]])
    badA=np.argwhere(~np.isclose(A[w],Ao,rtol=1e-9,atol=1e-30))
    print("     A mismatches", badA[:6].tolist())
