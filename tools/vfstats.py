import ctypes, os, sys, json, subprocess
os.environ["ADB_LIB_PATH"]="/root/repo/adapted_b200/csrc/libadb_stats.so"
sys.path.insert(0,"/root/repo")
import numpy as np, torch
from adapted_b200 import _lib
from adapted_b200.config import flatten_config, get_chemistry_specific_config
from adapted_b200.synth import make_reads_torch
import ctypes as C
chem=sys.argv[1]; n=20000
L=_lib.load(); ctx=_lib.Context(0)
spc=get_chemistry_specific_config(chem); flat=flatten_config(spc); m=flat["sig_preload_size"]
data=make_reads_torch(n,chem,m,seed=5,device="cuda")
cfg=_lib.fill_config(flat)
w=None
if flat["primary_method"]==1:
    from adapted_b200.detect import flatten_cnn_weights
    z=np.load("/root/repo/tests/golden/cnn_weights_rna004_130bps_v0.2.4.npz"); w=torch.from_numpy(flatten_cnn_weights({k:z[k] for k in z.files})).cuda()
rec=torch.zeros(n*512,dtype=torch.uint8,device="cuda"); st=torch.zeros(n//1000,dtype=torch.int32,device="cuda")
b=_lib.AdbBatch(signal=data["adc"].data_ptr(),sig_type=1,n_reads=n,m=m,batch_size=1000,offsets=data["offsets"].data_ptr(),full_lens=data["full_lens"].data_ptr(),calib_offset=data["calib_offset"].data_ptr(),calib_scale=data["calib_scale"].data_ptr())
out=(C.c_ulonglong*32)()
L.adb_vf_stats(out,1)
_lib.check(L.adb_detect_dev(ctx.handle,C.byref(b),C.byref(cfg),w.data_ptr() if w is not None else None,rec.data_ptr(),st.data_ptr(),None))
L.adb_vf_stats(out,1)
o=list(out); r=max(o[0],1)
names=["reads","passesA","passesB","series passes","stage","minmax/openpore/means","setup+bounds","sample brackets","round A","medians/lr","MAD setup+brackets","round B","mad_of","partition sums","checks+series"]
print(chem, "reads", o[0])
for i in range(1,4): print(f"  {names[i]}: {o[i]/r:.2f} per read")
tot=sum(o[4:15])
for i in range(4,15): print(f"  {names[i]}: {o[i]/r:.0f} cycles/read ({100*o[i]/tot:.1f}%)")
print("  total cycles/read", tot/r)
for rnd,base in (("A",16),("B",20)):
    p=max(o[1 if rnd=="A" else 2],1)
    print(f"  round {rnd} per pass: top-barrier wait {o[base]/p:.0f}, own counting {o[base+1]/p:.0f}, wait others {o[base+2]/p:.0f}, update+prepare {o[base+3]/p:.0f}")
