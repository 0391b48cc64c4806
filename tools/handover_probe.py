import sys; sys.path.insert(0, '/root/repo')
import numpy as np, torch
from adapted_b200 import _lib
from adapted_b200.config import get_chemistry_specific_config, config_as_dict, config_from_dict
from adapted_b200.detect import detect_reads
from adapted_b200.synth import make_reads_torch
from tests.golden_io import load_cnn_weights
n = 20000
spc = get_chemistry_specific_config("rna004")
d = config_as_dict(spc); d["cnn_boundaries"]["fallback_to_llr_short_reads"] = False
spc_nohm = config_from_dict(d)
data = make_reads_torch(n, "rna004", spc.sig_preload_size, seed=77, device="cuda")
h = {k: data[k].cpu().numpy() for k in ("adc", "offsets", "full_lens", "calib_offset", "calib_scale")}
ctx = _lib.default_context(0)
w = load_cnn_weights()
for name, cfg, opt in (("default", spc, 0), ("no follow-up", spc, 1), ("no hail mary", spc_nohm, 0), ("no hail mary, no follow-up", spc_nohm, 1)):
    ctx.set_option("no_cand_followup", opt)
    recs, st = detect_reads(h["adc"], h["offsets"], h["full_lens"], h["calib_offset"], h["calib_scale"], cfg, model=w, return_records=True)
    print(name, "handovers", ctx.query("validate_handovers"), "of", n, "pass", int((recs["success"] != 0).sum()))
ctx.set_option("no_cand_followup", 0)
