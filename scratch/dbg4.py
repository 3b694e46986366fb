import numpy as np, sys
sys.path.insert(0,'.')
import litiv_b200 as lv
from oracle import oracle as O
from litiv_b200.synth import SynthSequence
INT_STATE = ["rawmask","lastfg", "lastcolor", "lastdesc", "lut", "bg_color", "bg_desc", "unstable", "blinks", "lastraw", "lastrawblink", "dilinv"]
FLT_STATE = ["T", "R", "v", "Dlast", "DminLT", "DminST", "rawLT", "rawST", "finLT", "finST", "dsLT", "dsST"]
seq=SynthSequence(320,240,3,seed=1)
frames=[seq.frame(t) for t in range(130)]
o=O.Oracle(O.ALGO_SUBSENSE,mode=1,seed=0); o.initialize(frames[0])
om=[None]+[o.apply(frames[t],1.0 if t<=50 else 0.0) for t in range(1,130)]
nbad=0
for trial in range(40):
    g=lv.BackgroundSubtractorSuBSENSE(seed=0); g.initialize(frames[0])
    for t in range(1,130):
        mg=g.apply(frames[t],1.0 if t<=50 else 0.0)
        if (mg!=om[t]).any():
            nbad+=1
            print('trial',trial,'frame',t,'mask diff',(mg!=om[t]).sum(), np.argwhere(mg!=om[t])[:6].tolist())
            o2=O.Oracle(O.ALGO_SUBSENSE,mode=1,seed=0); o2.initialize(frames[0])
            for k in range(1,t+1): o2.apply(frames[k],1.0 if k<=50 else 0.0)
            for n in INT_STATE:
                a,b=g.state_get(n),o2.state_get(n)
                if (a!=b).any(): print('  ',n,int((a!=b).sum()),[(i//320,i%320) for i in np.flatnonzero(a!=b)[:6].tolist()] if a.size==76800 else np.flatnonzero(a!=b)[:6].tolist())
            for n in FLT_STATE:
                a,b=g.state_get(n),o2.state_get(n)
                if (a!=b).any(): print('  ',n,int((a!=b).sum()))
            break
print('bad trials',nbad)
