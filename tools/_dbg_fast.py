import sys, numpy as np
sys.path.insert(0,'/root/repo')
from psl_slam_b200 import ORBextractor, synth
gray,_,_ = synth.sequence(4, 2)
ex = ORBextractor()
k,d = ex.extract(gray[0]) if hasattr(ex,'extract') else ex(gray[0])
print(len(k))
