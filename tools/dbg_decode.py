import sys, numpy as np
import os; sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)),'..','tests')); sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)),'..'))
from _oracle import Oracle, frame_random, frame_lattice
from _gpu import decode
ORA=Oracle()
for (w,h) in [(4,2),(128,2),(256,4),(1920,8)]:
    y,c = frame_random(w,h,seed=709)
    for g in (0,3):
        got=decode(g,y,c); want=ORA.decode_nv12(g,y,c)
        bad=np.argwhere(got!=want)
        print(w,h,g,'mismatches',len(bad), bad[:6].tolist())
        if len(bad):
            r,cc,ch=bad[0]; print(' got',got[r,cc],'want',want[r,cc],'Y',y[r,cc],'C',c[r//2,(cc//2)*2:(cc//2)*2+2])
y,c = frame_lattice()
for g in (0,1,2,3):
    got=decode(g,y,c); want=ORA.decode_nv12(g,y,c)
    bad=np.argwhere(got!=want)
    print('lattice',g,'mismatches',len(bad))
    for r,cc,ch in bad[:8]:
        print('  at',r,cc,ch,'got',got[r,cc],'want',want[r,cc],'Y',y[r,cc],'CbCr',c[r//2,(cc//2)*2:(cc//2)*2+2])
    if len(bad):
        print('  channels', np.bincount(bad[:,2],minlength=4), 'rows%2', np.bincount(bad[:,0]%2), 'cols%4', np.bincount(bad[:,1]%4), 'col//128 %2', np.bincount((bad[:,1]//128)%2))
