"""Synthetic direct-RNA squiggles with known ground truth (BASELINE.json).

open pore N(220,3) -> adapter N(80,7) -> poly(A) N(108,2.5) -> RNA levels N(95,14) held for a few samples
+ N(0,3) noise, quantised to int16 ADC with a per-read calibration ``pA = (adc + offset) * scale``
(float32 operations, the product's definition of pod5's ``signal_pa``).

Two back ends produce the same *distribution*: numpy (seeded, used by tests and golden vectors) and
torch (any device, used by bench.py to build millions of reads directly in HBM).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import numpy as np

SCALE = 0.1755  # pA per ADC count, typical MinKNOW calibration


@dataclass
class SynthSpec:
    open_pore: Tuple[int, int]
    adapter: Tuple[int, int]
    polya: Tuple[int, int]
    rna: Tuple[int, int]
    hold: int


SPECS = {
    "rna004": SynthSpec((20, 200), (2500, 4500), (300, 3000), (3000, 40000), 12),
    "rna002": SynthSpec((20, 200), (4000, 8000), (600, 6000), (6000, 80000), 40),
}


@dataclass
class ReadBatch:
    """Ragged int16 reads truncated to the preload window, plus calibration and ground truth."""

    adc: np.ndarray        # int16 [sum(min(len, m))]
    offsets: np.ndarray    # int64 [N+1] element offsets into adc
    full_lens: np.ndarray  # int32 [N] untruncated lengths
    calib_offset: np.ndarray  # float32 [N]
    calib_scale: np.ndarray   # float32 [N]
    truth: np.ndarray      # int32 [N, 3]: open-pore end, adapter end, poly(A) end
    m: int

    @property
    def n(self) -> int:
        return int(self.full_lens.size)

    def to_dense_pa(self, rows: Optional[slice] = None) -> np.ndarray:
        """float32 [N, m] NaN-padded pA matrix, the layout the reference's seam takes
        (adapted/file_proc.py:160-175)."""
        idx = range(self.n)[rows] if rows is not None else range(self.n)
        out = np.full((len(idx), self.m), np.nan, dtype=np.float32)
        for j, i in enumerate(idx):
            a = self.adc[self.offsets[i]: self.offsets[i + 1]]
            out[j, : a.size] = calibrate(a, self.calib_offset[i], self.calib_scale[i])
        return out


def calibrate(adc: np.ndarray, offset, scale) -> np.ndarray:
    """pA = (float32(adc) + offset) * scale, every step rounded to float32."""
    return (adc.astype(np.float32) + np.float32(offset)) * np.float32(scale)


def make_reads(n: int, chemistry: str, m: int, seed: int, stress: bool = False,
               short_frac: float = 0.0) -> ReadBatch:
    """Seeded numpy generator.  `stress` draws poly(A) lengths up to the preload limit (BASELINE config 4);
    `short_frac` makes that fraction of reads end inside the adapter / poly(A) (short-read edge cases)."""
    spec = SPECS[chemistry.lower()]
    rng = np.random.default_rng(seed)
    chunks, offs, lens, truth = [], [0], [], []
    c_off = rng.uniform(-240.0, -200.0, size=n).astype(np.float32)
    c_scale = np.full(n, SCALE, dtype=np.float32)
    for i in range(n):
        n_op = int(rng.integers(*spec.open_pore))
        n_ad = int(rng.integers(*spec.adapter))
        n_pa = int(rng.integers(*spec.polya)) if not stress else int(rng.integers(spec.polya[0], m))
        n_rna = int(rng.integers(*spec.rna))
        full = n_op + n_ad + n_pa + n_rna
        if short_frac > 0 and rng.random() < short_frac:
            full = int(rng.integers(50, n_op + n_ad + n_pa + 500))
        k = min(full, m)
        t = np.arange(k)
        pa = np.empty(k, dtype=np.float64)
        e0, e1, e2 = n_op, n_op + n_ad, n_op + n_ad + n_pa
        seg0, seg1, seg2 = t < e0, (t >= e0) & (t < e1), (t >= e1) & (t < e2)
        seg3 = t >= e2
        pa[seg0] = rng.normal(220.0, 3.0, size=int(seg0.sum()))
        pa[seg1] = rng.normal(80.0, 7.0, size=int(seg1.sum()))
        pa[seg2] = rng.normal(108.0, 2.5, size=int(seg2.sum()))
        n3 = int(seg3.sum())
        if n3:
            levels = rng.normal(95.0, 14.0, size=n3 // spec.hold + 1)
            pa[seg3] = np.repeat(levels, spec.hold)[:n3] + rng.normal(0.0, 3.0, size=n3)
        adc = np.clip(np.rint(pa / SCALE - c_off[i]), -32768, 32767).astype(np.int16)
        chunks.append(adc)
        offs.append(offs[-1] + k)
        lens.append(full)
        truth.append((e0, e1, e2))
    return ReadBatch(
        adc=np.concatenate(chunks) if chunks else np.zeros(0, np.int16),
        offsets=np.asarray(offs, dtype=np.int64), full_lens=np.asarray(lens, dtype=np.int32),
        calib_offset=c_off, calib_scale=c_scale, truth=np.asarray(truth, dtype=np.int32).reshape(-1, 3), m=m)


def make_reads_torch(n: int, chemistry: str, m: int, seed: int, device="cuda", chunk: int = 4000, stress: bool = False,
                     short_frac: float = 0.0, short_min: int = 50):
    """Same squiggle distribution as :func:`make_reads`, generated with torch on `device` (bench.py builds
    10^5..10^7 reads directly in HBM).  `stress`: poly(A) lengths up to the preload limit (BASELINE config 4: long
    poly(A) / truncated preload); `short_frac`: that fraction of reads ends inside the adapter / poly(A).  Returns a dict of tensors: adc int16 [sum k_i], offsets int64 [n+1],
    full_lens int32 [n], calib_offset / calib_scale float32 [n], truth int32 [n, 3]."""
    import torch

    spec = SPECS[chemistry.lower()]
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    dev = torch.device(device)

    def randint(lo, hi, size):
        return torch.randint(lo, hi, (size,), generator=g, device=dev)

    parts, lens_all, k_all, off_all, truth_all = [], [], [], [], []
    t = torch.arange(m, device=dev)[None, :]
    for s in range(0, n, chunk):
        c = min(chunk, n - s)
        n_op, n_ad = randint(*spec.open_pore, c), randint(*spec.adapter, c)
        n_pa = randint(spec.polya[0], m, c) if stress else randint(*spec.polya, c)
        n_rna = randint(*spec.rna, c)
        e0, e1, e2 = n_op, n_op + n_ad, n_op + n_ad + n_pa
        full = e2 + n_rna
        if short_frac > 0:
            short = torch.rand((c,), generator=g, device=dev) < short_frac
            cut = short_min + (torch.rand((c,), generator=g, device=dev) * (e2 + 450).to(torch.float32)).to(torch.int64)
            full = torch.where(short, cut, full)
        k = torch.clamp(full, max=m)
        z = torch.randn((c, m), generator=g, device=dev)
        nlev = m // spec.hold + 2
        levels = 95.0 + 14.0 * torch.randn((c, nlev), generator=g, device=dev)
        li = torch.clamp((t - e2[:, None]) // spec.hold, 0, nlev - 1)
        pa = torch.gather(levels, 1, li) + 3.0 * z
        pa = torch.where(t < e2[:, None], 108.0 + 2.5 * z, pa)
        pa = torch.where(t < e1[:, None], 80.0 + 7.0 * z, pa)
        pa = torch.where(t < e0[:, None], 220.0 + 3.0 * z, pa)
        coff = -240.0 + 40.0 * torch.rand((c,), generator=g, device=dev)
        adc = torch.clamp(torch.round(pa / SCALE - coff[:, None]), -32768, 32767).to(torch.int16)
        parts.append(adc[t < k[:, None]])
        lens_all.append(full.to(torch.int32))
        k_all.append(k)
        off_all.append(coff.to(torch.float32))
        truth_all.append(torch.stack([e0, e1, e2], dim=1).to(torch.int32))
        del z, levels, li, pa, adc
    k = torch.cat(k_all)
    offsets = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    offsets[1:] = torch.cumsum(k, 0)
    return dict(adc=torch.cat(parts), offsets=offsets, full_lens=torch.cat(lens_all),
                calib_offset=torch.cat(off_all), calib_scale=torch.full((n,), SCALE, dtype=torch.float32, device=dev),
                truth=torch.cat(truth_all), m=m)
