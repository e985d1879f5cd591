import numpy as np
from amico_b200 import synth
from oracle import oracle as orc
from scratch.proto_gram import lh_gram
P = synth.make_problem(2, n_vox=300)
K = P.KERNELS
lut = synth.lut_index_numpy(P.DIRs, P.htable)
for mode in ('inv','solve'):
    bad=0; diffs=[]; objd=[]
    for i in range(300):
        A = synth.dictionary_for_direction('NODDI', K, lut[i]); H=A.T@A
        y = P.y[i].astype(np.float64)
        xo,_ = orc.nnls(A,y)
        xg,_ = lh_gram(H, A.T@y, mode)
        diffs.append(np.abs(xg-xo).max()); bad += ((xg>0)!=(xo>0)).any()
        objd.append(np.linalg.norm(A@xg-y)-np.linalg.norm(A@xo-y))
    print(mode, bad, np.median(diffs), np.max(diffs), 'obj diff', np.min(objd), np.max(objd))
