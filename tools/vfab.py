"""A/B timing of validate_fast_kernel variants: python tools/vfab.py <lib.so> <chem> [reads]"""
import os, sys
os.environ["ADB_LIB_PATH"] = os.path.abspath(sys.argv[1])
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np, torch
from adapted_b200 import _lib
from adapted_b200.config import flatten_config, get_chemistry_specific_config
from adapted_b200.synth import make_reads_torch
chem = sys.argv[2]; n = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
L = _lib.load(); ctx = _lib.Context(0)
for kv in os.environ.get("ADB_OPTS", "").split(","):
    if kv: _lib.check(L.adb_ctx_set_option(ctx.handle, kv.split("=")[0].encode(), int(kv.split("=")[1])))
spc = get_chemistry_specific_config(chem); flat = flatten_config(spc); m = flat["sig_preload_size"]
data = make_reads_torch(n, chem, m, seed=1234, device="cuda")
cfg = _lib.fill_config(flat)
w = None
if flat["primary_method"] == 1:
    from adapted_b200.detect import flatten_cnn_weights
    z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests/golden/cnn_weights_rna004_130bps_v0.2.4.npz"))
    w = torch.from_numpy(flatten_cnn_weights({k: z[k] for k in z.files})).cuda()
rec = torch.zeros(n * 512, dtype=torch.uint8, device="cuda"); st = torch.zeros((n + 999) // 1000, dtype=torch.int32, device="cuda")
b = _lib.AdbBatch(signal=data["adc"].data_ptr(), sig_type=1, n_reads=n, m=m, batch_size=1000, offsets=data["offsets"].data_ptr(),
                  full_lens=data["full_lens"].data_ptr(), calib_offset=data["calib_offset"].data_ptr(), calib_scale=data["calib_scale"].data_ptr())
def step():
    _lib.check(L.adb_detect_dev(ctx.handle, C.byref(b), C.byref(cfg), w.data_ptr() if w is not None else None, rec.data_ptr(), st.data_ptr(), None))
for _ in range(2): step()
torch.cuda.synchronize()
L.adb_ctx_set_timing(ctx.handle, 1)
for _ in range(3): step()
torch.cuda.synchronize()
tim = (C.c_double * 16)(); L.adb_ctx_get_timing(ctx.handle, tim)
names = ["gsb_pass", "gsb_small", "validate_fast", "llr_primary", "mvs_series", "cnn_conv", "cnn_prepost/sp", "handover"]
print(os.path.basename(sys.argv[1]), os.environ.get("ADB_OPTS", ""), chem, " ".join(f"{nm}={tim[2*i]/3:.2f}" for i, nm in enumerate(names) if tim[2*i+1] > 0), "checksum", int(rec.to(torch.int64).sum().item()))
