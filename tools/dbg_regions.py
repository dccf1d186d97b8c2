import sys; sys.path.insert(0,'/root/repo')
import numpy as np, torch
from sabreur_b200 import demux, synth, _ffi
w=synth.CFG3; tab=synth.table(w); bcs=[bytes(r) for r in tab]
n=60000
text=synth.reads_host(w,tab,777,n).tobytes()
with demux.Demux(bcs,w.mismatch,batch_bytes=4<<20) as dm:
    for b in dm.stream(text[i:i+(1<<20)] for i in range(0,len(text),1<<20)):
        out=np.zeros((64,8),dtype=np.uint32)
        dm.lib.sbr_debug_regions(dm.h,b.seq%2,out.ctypes.data,64)
        print('batch',b.seq,'scan_path',b.scan_path,'n',b.n_records,'consumed',b.consumed,'bytes',b.text.size,'status',b.status)
        nreg=int((out[:,0]>0).sum())
        nl=np.cumsum(np.concatenate([[0],out[:,0]]))
        exp=[int(x)&3 for x in nl[:nreg]]
        got=out[:nreg,2].tolist()
        print(' nreg',nreg,'phase mismatch at',[i for i in range(nreg) if exp[i]!=got[i]],'flags',out[:nreg,3].tolist(),'nrec sum',out[:nreg,1].sum())
        print(out[:4])
