#!/usr/bin/env python
"""bench.py -- throughput of the boundary-detection hot path on B200 (contract: see the task brief / DESIGN.md).

Workload (BASELINE.json configs[1]): RNA002 (rna002_70bps config), 100 000 synthetic reads per GPU in
minibatches of 1000, LLR primary path (global med/MAD -> LLR traces -> peak picking -> validation -> segment
statistics).  A "step" is one pass over all reads of the rank.

  value : reads/s with the int16 ADC reads already resident in HBM (CUDA events on the launching stream)
  e2e   : reads/s through the pipelined host ingest (pinned host buffers, H2D + kernels + D2H of the records)
  roofline : dominant kernel (per-read kernel) algorithmic bytes / its event-timed duration vs measured HBM peak
  cpu_baseline / --impl reference : the CPU oracle (port of the reference's path) on the host cores, bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "reads_per_sec"
WORKLOADS = {
    "rna002": "RNA002 rna002_70bps, LLR primary path, synthetic squiggles, minibatches of 1000 reads",
    "rna004": "RNA004 rna004_130bps, CNN primary path (+ hail-mary LLR fallback), synthetic squiggles, minibatches of 1000 reads",
}


def _env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference path) on the host cores
# ---------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    import warnings

    warnings.simplefilter("ignore")
    seed, n, chem = args
    from adapted_b200.config import get_chemistry_specific_config
    from adapted_b200.synth import make_reads
    from oracle import detect_ref

    spc = get_chemistry_specific_config(chem)
    b = make_reads(n, chem, spc.sig_preload_size, seed=seed)
    x = b.to_dense_pa()
    weights = _cnn_weights() if spc.primary_method == "cnn" else None
    t0 = time.perf_counter()
    if weights is not None:
        res = detect_ref.detect_cnn(x, b.full_lens, weights, spc)
        res = res if isinstance(res, list) else [res]
    else:
        res = detect_ref.detect_llr2(x, b.full_lens, spc)
    dt = time.perf_counter() - t0
    samples = int(np.minimum(b.full_lens, spc.sig_preload_size).sum())
    return n, samples, dt, sum(bool(r["success"]) for r in res)


def _cnn_weights():
    """the shipped model's weights (adapted/models/rna004_130bps@v0.2.4.pth) as committed fixture"""
    with np.load(os.path.join(ROOT, "tests", "golden", "cnn_weights_rna004_130bps_v0.2.4.npz")) as z:
        return {k: z[k] for k in z.files}


def cpu_arm(chem: str, steps: int, warmup: int, reads_per_worker: int, cores: int):
    """Times `steps` passes; each pass = `cores` independent minibatches of `reads_per_worker` reads, one per
    worker process (mirrors the reference's ProcessPoolExecutor, file_proc.py:738-784)."""
    from concurrent.futures import ProcessPoolExecutor

    from oracle._clib import build as build_oracle

    build_oracle()
    import multiprocessing as mp

    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("spawn")) as ex:
        for w in range(max(warmup, 1)):
            list(ex.map(_cpu_worker, [(1000 + i, 16, chem) for i in range(cores)]))
        reads = samples = 0
        dt = 0.0
        for s in range(steps):
            # synthetic inputs are generated inside the workers (untimed); a pass lasts as long as its slowest worker
            res = list(ex.map(_cpu_worker, [(2000 + s * cores + i, reads_per_worker, chem) for i in range(cores)]))
            reads += sum(r[0] for r in res)
            samples += sum(r[1] for r in res)
            dt += max(r[2] for r in res)
    return reads / dt, samples / dt, dt / steps, reads // steps


# ---------------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = False
        self.sm, self.smax, self.reasons = [], [], set()

    def run(self):
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.sm.append(float(f[0]))
                self.smax.append(float(f[1]))
                for nm, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        return {"sm_mhz": float(np.median(self.sm)), "sm_max_mhz": float(max(self.smax)), "reasons": sorted(self.reasons)}


def bind_to_gpu_numa_node(local_rank: int):
    """Run this rank on the CPUs of its GPU's NUMA node before any pinned host buffer is allocated (first touch puts the
    pages there): on a two-socket 8-GPU box half of the host -> device copies would otherwise cross the socket link.
    Best effort: returns the node or None."""
    try:
        import torch

        pr = torch.cuda.get_device_properties(local_rank)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{dev}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:  # noqa: BLE001 -- topology files missing (containers): keep the inherited affinity
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=100000, help="reads per GPU")
    ap.add_argument("--chemistry", default="rna002")
    ap.add_argument("--minibatch", type=int, default=1000)
    ap.add_argument("--chunk-batches", type=int, default=0,
                    help="minibatches per full-size H2D chunk of the pipelined ingest (default: 16, CNN path 32)")
    ap.add_argument("--cpu-reads-per-worker", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-steps-only", action="store_true",
                    help="for ncu launch lists: one untimed end-to-end call only, so that the list holds whole-step launches")
    ap.add_argument("--stress", action="store_true",
                    help="BASELINE config 4: poly(A) lengths up to the preload limit, 10 %% of the reads ending early")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank, world = _env_int("RANK", 0), _env_int("WORLD_SIZE", 1)
    local_rank = _env_int("LOCAL_RANK", 0)
    cores = os.cpu_count() or 1
    config = {"workload": WORKLOADS.get(args.chemistry.lower(), args.chemistry), "chemistry": args.chemistry, "reads_per_gpu": args.reads,
              "minibatch": args.minibatch, "preload_window": None, "l2": "inputs (>= 5 GB per GPU) exceed the 126 MB L2",
              "parallelism": f"minibatches sharded over {world} GPU(s), no collective"}

    # ---------------- reference arm: CPU only, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return 0
        from adapted_b200.config import flatten_config as _fc, get_chemistry_specific_config as _gc

        config["preload_window"] = _fc(_gc(args.chemistry))["sig_preload_size"]
        rpw = args.cpu_reads_per_worker or 128
        rps, sps, sec_per_step, reads_step = cpu_arm(args.chemistry, args.steps, args.warmup, rpw, cores)
        line = {"impl": "reference", "metric": METRIC, "value": rps, "unit": "reads/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec_per_step * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64/f32", "data": "synthetic",
                "config": config, "samples_per_sec": sps,
                "cpu_baseline": {"value": rps, "unit": "reads/s", "cores": cores, "kind": "port",
                                 "sample": f"{reads_step} reads per step = {cores} minibatches of {rpw} reads, one per worker process"},
                "e2e": {"value": rps, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    # ---------------- our arm ----------------
    import torch

    from adapted_b200 import _lib
    from adapted_b200.config import flatten_config, get_chemistry_specific_config
    from adapted_b200.synth import make_reads_torch

    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device (adapted_b200 has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    config["host_numa_node_rank0"] = numa_node
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL announces its version on stdout when the communicator is created; stdout carries the one JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if dist is not None:
            dist.barrier()

    spc = get_chemistry_specific_config(args.chemistry)
    flat = flatten_config(spc)
    m = flat["sig_preload_size"]
    config["preload_window"] = m
    n = args.reads
    gen_kw = {}
    if args.stress:
        gen_kw = dict(stress=True, short_frac=0.1, short_min=50 if flat["primary_method"] == 1 else flat["min_obs_adapter"] + 200)
        config["workload"] += " -- stress set (config 4: long poly(A) / truncated preload, short reads)"
    data = make_reads_torch(n, args.chemistry, m, seed=1234 + rank, device=dev, **gen_kw)
    torch.cuda.synchronize()
    samples = int(data["offsets"][-1].item())
    L = _lib.load()
    ctx = _lib.Context(local_rank)
    cfg = _lib.fill_config(flat)
    w_host = w_dev = None
    if flat["primary_method"] == 1:
        from adapted_b200.detect import flatten_cnn_weights

        w_host = torch.from_numpy(flatten_cnn_weights(_cnn_weights()))
        w_dev = w_host.to(dev)
    records = torch.zeros(n * 512, dtype=torch.uint8, device=dev)
    n_batches = (n + args.minibatch - 1) // args.minibatch
    status = torch.zeros(n_batches, dtype=torch.int32, device=dev)
    batch = _lib.AdbBatch(signal=data["adc"].data_ptr(), sig_type=_lib.SIG_I16, n_reads=n, m=m,
                          batch_size=args.minibatch, offsets=data["offsets"].data_ptr(),
                          full_lens=data["full_lens"].data_ptr(), calib_offset=data["calib_offset"].data_ptr(),
                          calib_scale=data["calib_scale"].data_ptr())
    # a non-default torch stream: its handle is passed to the C ABI, so torch's events bracket the kernels
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    stream = tstream.cuda_stream
    assert stream != 0

    def step():
        _lib.check(L.adb_detect_dev(ctx.handle, C.byref(batch), C.byref(cfg), w_dev.data_ptr() if w_dev is not None else None,
                                    records.data_ptr(), status.data_ptr(), C.c_void_p(stream)))

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    # clocks are sampled by rank 0 only (its own GPU): one nvidia-smi process every 100 ms per rank would load the host
    # and the driver lock of an 8-GPU box enough to show in the end-to-end number
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    L.adb_ctx_set_timing(ctx.handle, 1)
    launches0 = ctx.launches
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launches - launches0
    tim = (C.c_double * 16)()
    L.adb_ctx_get_timing(ctx.handle, tim)
    L.adb_ctx_set_timing(ctx.handle, 0)
    n_pass = int(torch.frombuffer(records.cpu().numpy(), dtype=torch.int32).reshape(n, 128)[:, 0].sum().item())
    lost = int((status != 0).sum().item())
    gsel_fallbacks = ctx.query("global_select_fallbacks") if flat["primary_method"] == 0 else 0
    val_handovers = ctx.query("validate_handovers")

    # ---------------- e2e: pinned host -> H2D -> kernels -> D2H ----------------
    host = {k: data[k].cpu().pin_memory() for k in ("adc", "offsets", "full_lens", "calib_offset", "calib_scale")}
    rec_host = torch.zeros(n * 512, dtype=torch.uint8).pin_memory()
    st_host = torch.zeros(n_batches, dtype=torch.int32).pin_memory()
    hbatch = _lib.AdbBatch(signal=host["adc"].data_ptr(), sig_type=_lib.SIG_I16, n_reads=n, m=m,
                           batch_size=args.minibatch, offsets=host["offsets"].data_ptr(),
                           full_lens=host["full_lens"].data_ptr(), calib_offset=host["calib_offset"].data_ptr(),
                           calib_scale=host["calib_scale"].data_ptr())

    if args.chunk_batches <= 0:  # copy-bound LLR path: short tail; kernel-bound CNN path: fewer per-chunk overheads
        args.chunk_batches = 32 if flat["primary_method"] == 1 else 16
    config["e2e_chunk_minibatches"] = args.chunk_batches

    def e2e_step():
        _lib.check(L.adb_detect_pipelined_host(ctx.handle, C.byref(hbatch), C.byref(cfg),
                                               w_host.data_ptr() if w_host is not None else None, rec_host.data_ptr(),
                                               st_host.data_ptr(), args.chunk_batches))

    e2e_steps = max(1, min(args.steps, 3))
    if args.profile_steps_only:
        e2e_steps = 0
    e2e_step()
    barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    barrier()
    e2e_s = time.perf_counter() - t0
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=2)
    h2d = sum(host[k].numel() * host[k].element_size() for k in host)
    d2h = rec_host.numel() + st_host.numel() * 4
    same = bool(torch.equal(rec_host, records.cpu()))

    # ---------------- reduce over ranks ----------------
    ms_t = torch.tensor([ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    tot = torch.tensor([n, samples, launches, n_pass, lost], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms_max, e2e_ms_max = float(ms_t[0]), float(ms_t[1])
    tot_reads, tot_samples = float(tot[0]), float(tot[1])
    value = tot_reads * args.steps / (ms_max / 1e3)
    e2e_value = tot_reads * e2e_steps / (e2e_ms_max / 1e3) if e2e_steps else None

    # roofline of the dominant kernel (rank-local): algorithmic bytes = 2 B/sample + 8 B calib + 4 B length + 512 B record
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # timing classes of adb_ctx_get_timing: 0 the streaming pass of the minibatch-global median/MAD (gsb_pass_kernel),
    # 1 its small sample / plan / finish kernels, 2 validate_fast_kernel, 3 llr_primary_kernel, 4 the length sort +
    # mvs_series_kernel, 5 CNN, 6 start-peak, 7 hand-over kernels (exact multi-pass select, histogram validate kernel)
    cls_names = ["global_select_pass", "global_select_small", "validate_fast_kernel", "llr_primary_kernel",
                 "mvs_series_kernel", "cnn_conv_kernels", "cnn_pre_post" if flat["primary_method"] == 1 else "start_peak",
                 "handover_kernels"]
    per_cls = {cls_names[i]: {"ms": tim[2 * i], "launches": int(tim[2 * i + 1])} for i in range(8) if tim[2 * i + 1] > 0}
    dom = max(range(8), key=lambda i: tim[2 * i])
    alg_bytes_per_step = 2.0 * samples + (8 + 4 + 512) * n
    if dom in (0, 3):  # these kernels stream the first max_obs_trace samples of every read once per launch
        launches_dom = max(per_cls[cls_names[dom]]["launches"], 1)
        alg_launch = 2.0 * float(torch.clamp(data["offsets"][1:] - data["offsets"][:-1], max=flat["max_obs_trace"]).sum().item())
    else:
        launches_dom = max(per_cls[cls_names[dom]]["launches"], 1)
        if dom in (1, 6, 7):
            # classes of several different kernels per step (small select kernels, CNN pre / post-processing, hand-over
            # kernels): the unit is the class's time per step, not an average over unlike launches
            launches_dom = args.steps
        alg_launch = alg_bytes_per_step
    avg_ms = tim[2 * dom] / launches_dom
    achieved = alg_launch / (avg_ms / 1e3) / 1e9 if avg_ms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        # the committed captures are per workload: RNA002 keys are bare class names, others carry the chemistry
        if dom == 5:
            traffic = tj.get("cnn")
        elif args.chemistry.lower() == "rna002":
            traffic = tj.get(cls_names[dom])
        else:
            traffic = tj.get(f"{cls_names[dom]} ({args.chemistry.lower()})")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": cls_names[dom], "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                "algorithmic_bytes_per_launch": alg_launch, "avg_launch_ms": avg_ms, "kernel_classes": per_cls,
                "whole_step_frac": (alg_bytes_per_step / ((ms / args.steps) / 1e3) / 1e9) / peak}
    if dom == 5:
        # the tensor-core convolutions (cnn_conv64_tc_kernel<2>, <3>; two launches per chunk of reads) are a dense
        # contraction: 64 556 800 algorithmic flop per read; every product is executed as three fp16 split
        # products, so the tensor pipe does 3x this work.  Peak: measured dense bf16/fp16 tensor throughput.
        tpeak = float(peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops", 2250.0)))
        flops_step = 64556800.0 * n
        tf = flops_step * args.steps / (tim[2 * dom] / 1e3) / 1e12 if tim[2 * dom] > 0 else 0.0
        roofline.update({"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s", "frac": tf / tpeak,
                         "algorithmic_flops_per_launch": flops_step * args.steps / launches_dom,
                         "executed_tensor_flops_factor": 3.0,
                         "hbm_view": {"achieved_GBps": achieved, "peak_GBps": peak, "frac": achieved / peak}})
        roofline.pop("algorithmic_bytes_per_launch", None)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64 traces / f32 statistics / i16 input", "data": "synthetic",
                "config": config, "samples_per_sec": tot_samples * args.steps / (ms_max / 1e3),
                "e2e": {"value": e2e_value, "unit": "reads/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "records_identical_to_device_run": same,
                        "api": "adb_detect_pipelined_host (pinned host int16 -> H2D -> kernels -> D2H records)"},
                "gpu_launches": int(tot[2]), "roofline": roofline, "clocks": sampler.summary(),
                "pass_fraction": float(tot[3]) / tot_reads, "lost_minibatches": int(tot[4]),
                "global_select_handovers_rank0": int(gsel_fallbacks), "validate_handovers_rank0": int(val_handovers)}
        if world == 1 and not args.no_cpu_baseline:
            # the CPU arm runs in a fresh interpreter: forking pool workers out of a process that has initialised CUDA
            # and torch's thread pools deadlocks the torch-CPU convolutions of the CNN oracle
            rpw = args.cpu_reads_per_worker or (256 if flat["primary_method"] == 1 else 4096)
            try:
                out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--chemistry",
                                      args.chemistry, "--steps", "1", "--warmup", "1", "--cpu-reads-per-worker", str(rpw)],
                                     capture_output=True, text=True, timeout=420).stdout
                ref = json.loads(out.strip().splitlines()[-1])
                line["cpu_baseline"] = {"value": ref["value"], "unit": "reads/s", "cores": ref["cpu_baseline"]["cores"],
                                        "kind": "port", "samples_per_sec": ref["samples_per_sec"],
                                        "sample": ref["cpu_baseline"]["sample"] + f" (oracle/detect_ref.py, {ref['ms_per_step'] / 1e3:.1f} s)"}
            except Exception as e:  # noqa: BLE001 -- the GPU line must still be printed
                line["cpu_baseline"] = {"value": None, "unit": "reads/s", "cores": cores, "kind": "port",
                                        "sample": f"failed: {type(e).__name__}: {e}"}
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main() or 0)
