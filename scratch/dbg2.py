import numpy as np
from amico_b200 import synth
from oracle import oracle as orc
from scratch.proto_gram import lh_gram
P = synth.make_problem(2, n_vox=300)
K = P.KERNELS
lut = synth.lut_index_numpy(P.DIRs, P.htable)
cnt=0
for i in range(300):
    A = synth.dictionary_for_direction('NODDI', K, lut[i]); H=A.T@A
    y = P.y[i].astype(np.float64)
    xo,_ = orc.nnls(A,y)
    xg,Pg = lh_gram(H, A.T@y, 'solve')
    if ((xg>0)!=(xo>0)).any():
        cnt+=1
        if cnt>4: break
        so = np.nonzero(xo>0)[0]; sg=np.nonzero(xg>0)[0]
        print('vox',i,'oracle supp',so, xo[so]); print('   gram supp',sg, xg[sg])
        w = A.T@(y-A@xg); print('   gram dual max', w.max(), np.argmax(w), 'oracle dual max', (A.T@(y-A@xo)).max())
        # cond of oracle support
        print('   cond A_S oracle', np.linalg.cond(A[:,so]), 'sv', np.linalg.svd(A[:,so],compute_uv=False)[-3:])
