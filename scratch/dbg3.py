import numpy as np, sys
sys.path.insert(0,'.')
import litiv_b200 as lv
from oracle import oracle as O
from litiv_b200.synth import SynthSequence
INT_STATE = ["rawmask","lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "unstable", "blinks", "lastraw", "lastrawblink", "dilinv"]
FLT_STATE = ["T", "R", "v", "Dlast", "DminLT", "DminST", "rawLT", "rawST", "finLT", "finST", "dsLT", "dsST"]
for trial in range(3):
    seq=SynthSequence(320,240,3,seed=1)
    g=lv.BackgroundSubtractorSuBSENSE(seed=0); g.initialize(seq.frame(0))
    o=O.Oracle(O.ALGO_SUBSENSE,mode=1,seed=0); o.initialize(seq.frame(0))
    for t in range(1,130):
        f=seq.frame(t); lr=1.0 if t<=50 else 0.0
        mg=g.apply(f,lr); mo=o.apply(f,lr)
        if (mg!=mo).any():
            print('trial',trial,'frame',t,'mask diff',(mg!=mo).sum(), np.argwhere(mg!=mo)[:5].tolist())
            for n in INT_STATE:
                a,b=g.state_get(n),o.state_get(n)
                if (a!=b).any(): print('  ',n,int((a!=b).sum()),np.flatnonzero(a!=b)[:4].tolist())
            for n in FLT_STATE:
                a,b=g.state_get(n),o.state_get(n)
                if (a!=b).any(): print('  ',n,int((a!=b).sum()))
            print(g.state_get('scalars')[:13].tolist()); print(o.state_get('scalars')[:13].tolist())
            break
    else: print('trial',trial,'all masks equal')
