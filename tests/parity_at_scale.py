#!/usr/bin/env python
"""Parity at scale: N synthetic reads through the CUDA path (int16 ingest) and through the CPU oracle, every
DetectResults field of every read compared.  Test infrastructure (imports oracle/); run by hand on a GPU box:

    python tests/parity_at_scale.py --chemistry rna004 --reads 100000 [--stress]

The same comparison runs as a -m gpu test (tests/test_gpu_parity_at_scale.py, sizes from ADB_PARITY_READS).

LLR path: every field must be identical (floats within 1e-5).  CNN path: the float32 convolutions are summed in a
different order than torch's CPU kernels, so a primary coordinate may move by one downscaled step (north_star: +-1
step); the script counts the reads that are field-for-field identical, those whose primaries moved by <= 1 step, and
anything else (must be 0).  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _oracle_minibatch(args):
    import warnings

    warnings.simplefilter("ignore")
    chem, adc, offs, lens, coff, cs, m = args
    from adapted_b200.config import get_chemistry_specific_config
    from adapted_b200.synth import calibrate
    from oracle import detect_ref

    spc = get_chemistry_specific_config(chem)
    n = lens.size
    x = np.full((n, m), np.nan, np.float32)
    for i in range(n):
        a = adc[offs[i] - offs[0]: offs[i + 1] - offs[0]]
        x[i, : a.size] = calibrate(a, coff[i], cs[i])
    if spc.primary_method == "cnn":
        with np.load(os.path.join(ROOT, "tests", "golden", "cnn_weights_rna004_130bps_v0.2.4.npz")) as z:
            w = {k: z[k] for k in z.files}
        res = detect_ref.detect_cnn(x, lens, w, spc)
        return res if isinstance(res, list) else [res]
    return detect_ref.detect_llr2(x, lens, spc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--chemistry", default="rna004")
    ap.add_argument("--reads", type=int, default=100000)
    ap.add_argument("--minibatch", type=int, default=1000)
    ap.add_argument("--stress", action="store_true")
    ap.add_argument("--seed", type=int, default=4242)
    ap.add_argument("--chunks", type=int, default=1, help="repeat with seeds seed, seed+1, ... (bounds host memory)")
    args = ap.parse_args()
    if args.chunks > 1:
        import subprocess

        tot = None
        for c in range(args.chunks):
            cmd = [sys.executable, os.path.abspath(__file__), "--chemistry", args.chemistry, "--reads", str(args.reads),
                   "--minibatch", str(args.minibatch), "--seed", str(args.seed + c)] + (["--stress"] if args.stress else [])
            out = subprocess.run(cmd, capture_output=True, text=True).stdout.strip().splitlines()
            d = json.loads(out[-1])
            if tot is None:
                tot = d
                tot["seeds"] = [args.seed]
            else:
                for k in ("reads", "lost_minibatches", "identical", "primary_moved_by_one_step", "other_differences",
                          "gpu_s_incl_host_conversion", "oracle_s"):
                    tot[k] += d[k]
                tot["examples"] = (tot["examples"] + d["examples"])[:5]
                tot["seeds"].append(args.seed + c)
        tot.pop("pass_fraction_oracle", None)
        print(json.dumps(tot))
        return 0 if tot["other_differences"] == 0 else 1

    d = run(args.chemistry, args.reads, args.minibatch, args.stress, args.seed)
    print(json.dumps(d))
    return 0 if d["other_differences"] == 0 else 1


def run(chemistry: str, reads: int, minibatch: int = 1000, stress: bool = False, seed: int = 4242) -> dict:
    """One comparison: `reads` synthetic reads through the CUDA path (int16 ingest) and through the CPU oracle on all
    host cores; returns the counts (see the module docstring)."""
    import argparse as _ap

    args = _ap.Namespace(chemistry=chemistry, reads=reads, minibatch=minibatch, stress=stress, seed=seed)
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from adapted_b200.config import flatten_config, get_chemistry_specific_config
    from adapted_b200.detect import detect_reads
    from adapted_b200.synth import make_reads_torch
    from tests.golden_io import load_cnn_weights
    from tests.helpers import as_dict, diff_results

    spc = get_chemistry_specific_config(args.chemistry)
    flat = flatten_config(spc)
    m, mbs, n = flat["sig_preload_size"], args.minibatch, args.reads
    cnn = flat["primary_method"] == 1
    kw = dict(stress=True, short_frac=0.1, short_min=50 if cnn else flat["min_obs_adapter"] + 200) if args.stress else {}
    data = make_reads_torch(n, args.chemistry, m, seed=args.seed, device="cuda", **kw)
    host = {k: data[k].cpu().numpy() for k in ("adc", "offsets", "full_lens", "calib_offset", "calib_scale")}
    t0 = time.perf_counter()
    got, status = detect_reads(host["adc"], host["offsets"], host["full_lens"], host["calib_offset"], host["calib_scale"], spc,
                               model=load_cnn_weights() if cnn else None, minibatch_size=mbs)
    t_gpu = time.perf_counter() - t0
    jobs = []
    offs = host["offsets"]
    for s in range(0, n, mbs):
        e = min(s + mbs, n)
        jobs.append((args.chemistry, host["adc"][offs[s]: offs[e]], offs[s: e + 1], host["full_lens"][s:e],
                     host["calib_offset"][s:e], host["calib_scale"][s:e], m))
    t0 = time.perf_counter()
    with ProcessPoolExecutor(max_workers=os.cpu_count(), mp_context=mp.get_context("spawn")) as ex:
        want = [r for mb in ex.map(_oracle_minibatch, jobs) for r in mb]
    t_cpu = time.perf_counter() - t0
    ds = flat["downscale_factor"]
    identical = moved = bad = 0
    examples = []
    prim = ("cnn_adapter_end", "cnn_polya_end")
    for i, (g, w) in enumerate(zip(got, want)):
        d = diff_results([g], [w])
        if not d:
            identical += 1
            continue
        gd, wd = as_dict(g), as_dict(w)
        ok = cnn and all((gd.get(k) is None and wd.get(k) is None) or
                         (gd.get(k) is not None and wd.get(k) is not None and abs(int(gd[k]) - int(wd[k])) <= ds) for k in prim)
        same_primary = all((gd.get(k) is None and wd.get(k) is None) or
                           (gd.get(k) is not None and wd.get(k) is not None and int(gd[k]) == int(wd[k])) for k in prim)
        cand_same = (gd.get("polya_candidates") is None and wd.get("polya_candidates") is None) or np.array_equal(
            gd.get("polya_candidates"), wd.get("polya_candidates"))
        if ok and not (same_primary and cand_same):
            moved += 1
        else:
            bad += 1
            if len(examples) < 5:
                examples.append({"read": i, "diff": d[:3]})
    return {"chemistry": args.chemistry, "stress": args.stress, "reads": n, "lost_minibatches": int((status != 0).sum()),
            "pass_fraction_oracle": float(np.mean([bool(as_dict(w)["success"]) for w in want])),
            "identical": identical, "primary_moved_by_one_step": moved, "other_differences": bad,
            "examples": examples, "gpu_s_incl_host_conversion": round(t_gpu, 2), "oracle_s": round(t_cpu, 2),
            "cores": os.cpu_count(), "seed": args.seed}




if __name__ == "__main__":
    sys.exit(main())
