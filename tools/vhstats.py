"""Per-phase cycles of validate_hist_kernel from an instrumented build (-DADB_VH_STATS): python tools/vhstats.py <lib.so> <chem>"""
import os, sys
os.environ["ADB_LIB_PATH"] = os.path.abspath(sys.argv[1])
os.environ["ADB_HIST_VALIDATE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from adapted_b200 import _lib
from adapted_b200.config import flatten_config, get_chemistry_specific_config
from adapted_b200.synth import make_reads_torch
chem = sys.argv[2]; n = 20000
L = _lib.load(); ctx = _lib.Context(0)
spc = get_chemistry_specific_config(chem); flat = flatten_config(spc); m = flat["sig_preload_size"]
data = make_reads_torch(n, chem, m, seed=5, device="cuda")
cfg = _lib.fill_config(flat)
w = None
if flat["primary_method"] == 1:
    from adapted_b200.detect import flatten_cnn_weights
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/cnn_weights_rna004_130bps_v0.2.4.npz"))
    w = torch.from_numpy(flatten_cnn_weights({k: z[k] for k in z.files})).cuda()
rec = torch.zeros(n * 512, dtype=torch.uint8, device="cuda"); st = torch.zeros(n // 1000, dtype=torch.int32, device="cuda")
b = _lib.AdbBatch(signal=data["adc"].data_ptr(), sig_type=1, n_reads=n, m=m, batch_size=1000, offsets=data["offsets"].data_ptr(),
                  full_lens=data["full_lens"].data_ptr(), calib_offset=data["calib_offset"].data_ptr(), calib_scale=data["calib_scale"].data_ptr())
out = (C.c_ulonglong * 16)()
L.adb_vh_stats.argtypes = [C.c_void_p, C.c_int]
def step():
    _lib.check(L.adb_detect_dev(ctx.handle, C.byref(b), C.byref(cfg), w.data_ptr() if w is not None else None, rec.data_ptr(), st.data_ptr(), None))
step(); torch.cuda.synchronize(); L.adb_vh_stats(out, 1)
step(); torch.cuda.synchronize(); L.adb_vh_stats(out, 1)
o = list(out); r = max(o[0], 1)
names = ["reads", "open pores + means", "cuts + zero accumulators", "stream (tiles -> MMA)", "readout + prefix", "queries", "median / arena zero", "sums + checks + record"]
print(os.path.basename(sys.argv[1]), chem, "reads", o[0])
tot = sum(o[1:8])
for i in range(1, 8): print(f"  {names[i]}: {o[i]/r:.0f} cycles/read ({100*o[i]/tot:.1f}%)")
print("  total cycles/read", tot / r)
sub = []

