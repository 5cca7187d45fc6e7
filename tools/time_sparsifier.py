import sys, time
import os; ROOT=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
import oracle_lib as O
from plinopt_b200 import capi
capi.set_device(0)
for name,q,c in [("2x2x2_7_DPS-smallrat-12.2034_L",0,4),("4x4x4_48_rational_L",2147483647,11),("4x4x4_48_rational_L",0,11),("3x4x7_63_rational_R",0,11),("4x4x4_48_rational_L",2147483647,40)]:
    M=O.dense_fractions(name)
    capi.sparsifier(M,q,4,c,True)
    reps=20
    t=time.perf_counter()
    for _ in range(reps): CoB,Res,ok,st=capi.sparsifier(M,q,4,c,True)
    dt=(time.perf_counter()-t)/reps
    t=time.perf_counter(); O.sparsifier(M,q,4,c,True); dto=time.perf_counter()-t
    print(name,q,c,'gpu driver %.6fs'%dt,'%.3g candidates/s'%(st['candidates']/dt),'oracle %.3fs'%dto,st, 'nnz', sum(1 for r in Res for v in r if v!=0))
