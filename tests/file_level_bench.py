import sys, time, os, tempfile
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from adapted_b200.config import get_chemistry_specific_config, flatten_config
from adapted_b200.synth import make_reads_torch
from adapted_b200.ingest import write_container, detect_files
sys.path.insert(0, "tests")
from tests.golden_io import load_cnn_weights
chem = sys.argv[1]; n = int(sys.argv[2])
spc = get_chemistry_specific_config(chem); flat = flatten_config(spc)
d = make_reads_torch(n, chem, flat["sig_preload_size"], seed=5, device="cuda")
h = {k: d[k].cpu().numpy() for k in ("adc", "offsets", "full_lens", "calib_offset", "calib_scale")}
ids = [f"{i:08x}-aaaa-bbbb-cccc-000000000000" for i in range(n)]
tmp = tempfile.mkdtemp(dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
t0 = time.perf_counter(); p = write_container(os.path.join(tmp, "reads"), h["adc"], h["offsets"], h["full_lens"], h["calib_offset"], h["calib_scale"], ids); t1 = time.perf_counter()
model = load_cnn_weights() if flat["primary_method"] == 1 else None
for rep in range(2):
    out = os.path.join(tmp, f"out{rep}")
    t2 = time.perf_counter(); st = detect_files([p], out, spc, model=model); t3 = time.perf_counter()
    print(chem, n, "container write", round(t1 - t0, 2), "s; detect_files", round(t3 - t2, 2), "s =", round(n / (t3 - t2)), "reads/s", st)
