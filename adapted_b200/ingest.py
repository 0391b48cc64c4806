"""File-level driver of the hot path: signal container -> GPU detection -> boundary tables.

Stands in for the part of ``adapted detect`` / ``adapted continue`` between the pod5 reader and the CSV files
(adapted/file_proc.py:143-214 producer, 217-266 worker, 312-457 saver threads, 97-140 continue logic).  ``pod5`` is
not installable in the build image, so the input is the native container -- the same
information a pod5 file holds per read: int16 ADC samples truncated to the preload window, the calibration pair, the
untruncated length and the read id.  Reads are cut into minibatches in file order (parser.py:95-99), results are
written 4000 reads per table like the reference's savers; a minibatch the reference loses on an exception outside its
per-read ``try`` is logged and appears in neither table, as there.
"""
from __future__ import annotations

import logging
import os
from typing import Any, Dict, Optional, Sequence, Set

import numpy as np

from .config import flatten_config
from .output import BoundaryTableWriter

CONTAINER_KEYS = ("adc", "offsets", "full_lens", "calib_offset", "calib_scale", "read_ids")


def write_container(path: str, adc: np.ndarray, offsets: np.ndarray, full_lens: np.ndarray, calib_offset: np.ndarray,
                    calib_scale: np.ndarray, read_ids: Sequence[str]) -> str:
    """One file: int16 blob + int64 offsets + int32 lengths + float32 calibration pairs + read ids."""
    if not path.endswith(".npz"):
        path += ".npz"
    np.savez(path, adc=np.ascontiguousarray(adc, np.int16), offsets=np.ascontiguousarray(offsets, np.int64),
             full_lens=np.ascontiguousarray(full_lens, np.int32), calib_offset=np.ascontiguousarray(calib_offset, np.float32),
             calib_scale=np.ascontiguousarray(calib_scale, np.float32), read_ids=np.asarray(list(read_ids), dtype="U36"))
    return path


def read_container(path: str) -> Dict[str, np.ndarray]:
    with np.load(path) as z:
        return {k: z[k] for k in CONTAINER_KEYS}


def processed_read_ids(continue_from: str, failed_only: bool = False) -> Set[str]:
    """Read ids already present in the tables of a previous run (scan_processed_reads, file_proc.py:103-130)."""
    done: Set[str] = set()
    subs = [("failed_reads", "failed_reads_")] + ([] if failed_only else [("boundaries", "detected_boundaries_")])
    for sub, prefix in subs:
        d = os.path.join(continue_from, sub)
        if not os.path.isdir(d):
            continue
        for fn in os.listdir(d):
            if fn.startswith(prefix) and fn.endswith(".csv"):
                with open(os.path.join(d, fn)) as f:
                    done.update(line.split(",")[0] for line in f.readlines()[1:])
    return done


def detect_file(path: str, out_dir: str, spc: Any, model: Any = None, minibatch_size: int = 1000,
                batch_size_output: int = 4000, continue_run: bool = False, device: int = 0,
                reads_per_call: int = 64000) -> Dict[str, int]:
    """Run the detection over every read of a container and write ``<out_dir>/boundaries/detected_boundaries_<i>.csv``
    and ``<out_dir>/failed_reads/failed_reads_<i>.csv``.  ``continue_run`` skips reads already in the tables of
    ``out_dir`` and continues the file numbering (``adapted continue``)."""
    from .detect import detect_reads

    c = read_container(path)
    ids = c["read_ids"]
    keep = np.arange(ids.size)
    if continue_run:
        done = processed_read_ids(out_dir)
        keep = np.array([i for i in keep if str(ids[i]) not in done], dtype=np.int64)
    flat = flatten_config(spc)
    method = flat["primary_method"]
    log = "" if method == 0 else None
    if continue_run:
        writer = BoundaryTableWriter.continue_from(out_dir, method, batch_size_output=batch_size_output, llr_detect_log=log)
    else:
        writer = BoundaryTableWriter(os.path.join(out_dir, "boundaries"), os.path.join(out_dir, "failed_reads"), method,
                                     batch_size_output=batch_size_output, llr_detect_log=log)
    offsets = c["offsets"]
    stats = {"reads": int(keep.size), "pass": 0, "fail": 0, "lost": 0}
    per_call = max(minibatch_size, reads_per_call // minibatch_size * minibatch_size)
    with writer:
        for s in range(0, keep.size, per_call):
            sel = keep[s: s + per_call]
            # compact the selected reads into one ragged batch (file order)
            lens = (offsets[sel + 1] - offsets[sel]).astype(np.int64)
            off = np.zeros(sel.size + 1, dtype=np.int64)
            np.cumsum(lens, out=off[1:])
            if sel.size and sel[-1] - sel[0] + 1 == sel.size:
                adc = c["adc"][offsets[sel[0]]: offsets[sel[-1] + 1]]
            else:
                adc = np.concatenate([c["adc"][offsets[i]: offsets[i + 1]] for i in sel]) if sel.size else np.zeros(0, np.int16)
            recs, status = detect_reads(adc, off, c["full_lens"][sel], c["calib_offset"][sel], c["calib_scale"][sel], spc,
                                        model=model, minibatch_size=minibatch_size, device=device, return_records=True)
            for bi, st in enumerate(status):
                a, b = bi * minibatch_size, min((bi + 1) * minibatch_size, sel.size)
                if st != 0:
                    logging.error("minibatch of %d reads lost (status %d), like the reference's handle_completed_future",
                                  b - a, int(st))
                    stats["lost"] += b - a
                    continue
                r = recs[a:b]
                writer.add(r, [str(x) for x in ids[sel[a:b]]])
                n_ok = int((r["success"] != 0).sum())
                stats["pass"] += n_ok
                stats["fail"] += (b - a) - n_ok
    stats["files"] = len(writer.files)
    return stats
