import numpy as np, sys
sys.path.insert(0,'.')
import litiv_b200 as lv
from oracle import oracle as O
from litiv_b200.synth import SynthSequence
seq=SynthSequence(320,240,3,seed=11)
g=lv.BackgroundSubtractorSuBSENSE(seed=7); g.initialize(seq.frame(0))
o=O.Oracle(O.ALGO_SUBSENSE,mode=1,seed=7); o.initialize(seq.frame(0))
l0=g.state_get('lut').copy()
g.apply(seq.frame(1),1.0); o.apply(seq.frame(1),1.0)
l1=g.state_get('lut'); print((l0!=l1).sum(), g.state_get('scalars')[:13], l0[:10], l1[:10])
print(o.state_get('scalars')[:13], o.state_get('lut')[:10])
