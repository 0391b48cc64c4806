"""Signal-processing configuration: host-side mirror of the reference's ``SigProcConfig``.

Section and field names follow the reference's TOML schema (adapted/config/sig_proc.py:22-221,
adapted/config/config_files/rna00*.toml) so the shipped chemistry files -- and any custom ``--config``
TOML -- load unchanged.  The detectors accept *either* this class or the reference's own
``SigProcConfig`` object (attribute access only), and :func:`flatten_config` turns both into the
fixed-layout ``adb_config`` block the CUDA library reads (include/adapted_b200.h).

The values of the two shipped chemistries are restated in :data:`_PRESETS` (they are the kernel
parameters).
"""
from __future__ import annotations

import copy
import math
from dataclasses import dataclass, field, fields, is_dataclass
from typing import Any, Dict, Optional, Tuple

Range = Tuple[Optional[float], Optional[float]]
INF = math.inf


@dataclass
class CoreConfig:
    min_obs_adapter: int = 1000
    max_obs_adapter: int = 6500
    min_obs_polya: int = 100
    downscale_factor: int = 10
    max_obs_trace: int = 16000
    sig_norm_outlier_thresh: float = 5.0


@dataclass
class CNNBoundariesConfig:
    cnn_detect: bool = True
    model_name: str = "rna004_130bps@v0.2.4.pth"
    polya_cand_k: int = 15
    fallback_to_llr_short_reads: bool = True


@dataclass
class LLRBoundariesConfig:
    llr_detect: bool = False
    adapter_peak_prominence: float = 1.0
    adapter_peak_rel_height: float = 1.0
    adapter_peak_width: int = 1000
    polya_peak_prominence: float = 1.0
    polya_peak_rel_height: float = 0.5
    polya_peak_width: int = 50


@dataclass
class MVSPolyAConfig:
    mvs_detect_check: bool = True
    mvs_detect_overwrite: bool = False
    search_window: int = 500
    pA_mean_window: int = 20
    pA_mean_range: Range = (None, None)
    pA_var_window: int = 100
    pA_var_range: Range = (None, 20.0)
    median_shift_range: Range = (20.0, None)
    median_shift_window: int = 2000
    polyA_window: int = 300
    polyA_med_range: Range = (90.0, 130.0)
    polyA_local_range: Range = (0.0, 15.0)
    pA_mean_adapter_med_scale_range: Range = (1.3, None)


@dataclass
class StreamingConfig:
    """adapted/config/sig_proc.py:140-158 (defaults: RNA002 live data) -- parameters of mean_var_shift_polyA_detect."""
    min_obs_adapter: int = 2500
    min_obs_post_loc: int = 300
    search_increment_step: int = 100
    pA_mean_window: int = 20
    pA_mean_range: Range = (90.0, 130.0)
    pA_var_window: int = 100
    pA_var_range: Range = (None, 20.0)
    median_shift_window: int = 2000
    median_shift_range: Range = (20.0, None)
    polyA_window: int = 300
    polyA_med_range: Range = (90.0, 130.0)
    polyA_local_range: Range = (0.0, 10.0)


@dataclass
class RNAStartPeakConfig:
    detect_rna_start_peak: bool = False
    downscale_factor: int = 10
    start_peak_max_idx: int = 150
    offset1: int = 10
    offset2: int = 100
    open_pore_pa: float = 195.0


@dataclass
class MedShiftConfig:
    detect_med_shift: bool = False
    med_shift_window: int = 2000
    med_shift_range: Range = (20.0, None)


@dataclass
class RealRangeConfig:
    detect_open_pores: bool = True
    real_signal_check: bool = True
    mean_window: int = 300
    mean_start_range: Range = (50.0, 100.0)
    mean_end_range: Range = (75.0, 120.0)
    max_obs_local_range: int = 5000
    local_range: Range = (10.0, 30.0)
    adapter_mad_range: Range = (3.0, 12.0)


_SECTIONS = {
    "core": CoreConfig,
    "llr_boundaries": LLRBoundariesConfig,
    "mvs_polya": MVSPolyAConfig,
    "real_range": RealRangeConfig,
    "cnn_boundaries": CNNBoundariesConfig,
    "med_shift": MedShiftConfig,
    "rna_start_peak": RNAStartPeakConfig,
    "streaming": StreamingConfig,  # Optional in the reference (sig_proc.py: `streaming: Optional[StreamingConfig]`)
}


@dataclass
class SigProcConfig:
    core: CoreConfig = field(default_factory=CoreConfig)
    llr_boundaries: LLRBoundariesConfig = field(default_factory=LLRBoundariesConfig)
    mvs_polya: MVSPolyAConfig = field(default_factory=MVSPolyAConfig)
    real_range: RealRangeConfig = field(default_factory=RealRangeConfig)
    cnn_boundaries: CNNBoundariesConfig = field(default_factory=CNNBoundariesConfig)
    med_shift: MedShiftConfig = field(default_factory=MedShiftConfig)
    rna_start_peak: RNAStartPeakConfig = field(default_factory=RNAStartPeakConfig)
    streaming: Optional[StreamingConfig] = None
    primary_method: Optional[str] = None
    sig_preload_size: int = 0

    def __post_init__(self):
        self.update_primary_method()
        self.update_sig_preload_size()

    # adapted/config/sig_proc.py:182-190
    def update_sig_preload_size(self) -> None:
        extra = 0
        if self.mvs_polya.mvs_detect_check:
            extra = self.mvs_polya.search_window + max(
                self.mvs_polya.median_shift_window, self.mvs_polya.polyA_window
            )
        self.sig_preload_size = self.core.max_obs_trace + extra

    # adapted/config/sig_proc.py:192-208
    def update_primary_method(self) -> None:
        flags = (
            bool(self.llr_boundaries.llr_detect),
            bool(self.cnn_boundaries.cnn_detect),
            bool(self.rna_start_peak.detect_rna_start_peak),
        )
        if sum(flags) != 1:
            raise ValueError("Exactly one primary method must be enabled")
        self.primary_method = ("llr", "cnn", "start_peak")[flags.index(True)]

    def copy(self) -> "SigProcConfig":
        return copy.deepcopy(self)


# Values of the two shipped chemistry files (rna002_70bps@v0.2.4.toml, rna004_130bps@v0.2.4.toml).
_COMMON_MVS = dict(
    mvs_detect_check=True, mvs_detect_overwrite=False, search_window=500, pA_mean_window=20,
    pA_var_window=100, median_shift_range=(5.0, INF), median_shift_window=1000,
    polyA_med_range=(-INF, INF), polyA_local_range=(-INF, INF),
    pA_mean_adapter_med_scale_range=(1.3, INF),
)
_COMMON_RR = dict(
    detect_open_pores=True, real_signal_check=True, mean_window=300, mean_start_range=(-INF, INF),
    mean_end_range=(-INF, INF), max_obs_local_range=5000, local_range=(7.0, 35.0),
    adapter_mad_range=(3.0, 12.0),
)
_PRESETS: Dict[str, Dict[str, Dict[str, Any]]] = {
    "rna002": {
        "core": dict(max_obs_trace=25000, min_obs_adapter=2000, max_obs_adapter=12000, min_obs_polya=100,
                     downscale_factor=20, sig_norm_outlier_thresh=5.0),
        "cnn_boundaries": dict(cnn_detect=False, model_name="rna002_70bps@v0.2.4.pth", polya_cand_k=15,
                               fallback_to_llr_short_reads=True),
        "llr_boundaries": dict(llr_detect=True, adapter_peak_prominence=1.0, adapter_peak_rel_height=1.0,
                               adapter_peak_width=1500, polya_peak_prominence=1.0,
                               polya_peak_rel_height=0.5, polya_peak_width=50),
        "mvs_polya": dict(_COMMON_MVS, pA_var_range=(-INF, 20.0)),
        "real_range": dict(_COMMON_RR),
        "med_shift": dict(detect_med_shift=False, med_shift_window=1000, med_shift_range=(5.0, INF)),
        "rna_start_peak": dict(detect_rna_start_peak=False),
    },
    "rna004": {
        "core": dict(max_obs_trace=16000, min_obs_adapter=1000, max_obs_adapter=6500, min_obs_polya=100,
                     downscale_factor=10, sig_norm_outlier_thresh=5.0),
        "cnn_boundaries": dict(cnn_detect=True, model_name="rna004_130bps@v0.2.4.pth", polya_cand_k=10,
                               fallback_to_llr_short_reads=True),
        "llr_boundaries": dict(llr_detect=False, adapter_peak_prominence=1.0, adapter_peak_rel_height=1.0,
                               adapter_peak_width=1000, polya_peak_prominence=1.0,
                               polya_peak_rel_height=0.5, polya_peak_width=50),
        "mvs_polya": dict(_COMMON_MVS, pA_var_range=(-INF, 30.0)),
        "real_range": dict(_COMMON_RR),
        "med_shift": dict(detect_med_shift=False, med_shift_window=2000, med_shift_range=(5.0, INF)),
        "rna_start_peak": dict(detect_rna_start_peak=False, downscale_factor=10, start_peak_max_idx=150,
                               offset1=10, offset2=100, open_pore_pa=195.0),
    },
}


def config_from_dict(d: Dict[str, Any]) -> SigProcConfig:
    """Build a config from a parsed TOML dict; unknown sections/keys are rejected like the reference
    does (adapted/config/base.py:120-174)."""
    kwargs = {}
    for section, content in d.items():
        if section not in _SECTIONS:
            raise ValueError(f"Invalid config file. Unknown key(s): {section}")
        cls = _SECTIONS[section]
        valid = {f.name for f in fields(cls)}
        bad = [k for k in content if k not in valid]
        if bad:
            raise ValueError(f"Invalid config file. Could not parse section {section}: unknown {bad}")
        conv = {k: (tuple(v) if isinstance(v, list) else v) for k, v in content.items()}
        kwargs[section] = cls(**conv)
    return SigProcConfig(**kwargs)


def load_config(path: str) -> SigProcConfig:
    import toml

    return config_from_dict(toml.load(path))


def get_chemistry_specific_config(chemistry: str) -> SigProcConfig:
    """Same call as adapted/config/sig_proc.py:245-255; primary method / preload size are already
    up to date on return (the reference needs two extra update_* calls)."""
    key = chemistry.lower()
    if key not in _PRESETS:
        raise ValueError(f"Unknown chemistry: {chemistry}")
    return config_from_dict(copy.deepcopy(_PRESETS[key]))


def start_peak_config(chemistry: str = "RNA004") -> SigProcConfig:
    """The start-peak companion configuration of BASELINE config 2:
    start-peak primary, MVS check off, median-shift check on."""
    d = copy.deepcopy(_PRESETS[chemistry.lower()])
    d["rna_start_peak"]["detect_rna_start_peak"] = True
    d["llr_boundaries"]["llr_detect"] = False
    d["cnn_boundaries"]["cnn_detect"] = False
    d["mvs_polya"]["mvs_detect_check"] = False
    d["med_shift"]["detect_med_shift"] = True
    return config_from_dict(d)


# ---------------------------------------------------------------------------------------------
# flattening for the C-ABI
# ---------------------------------------------------------------------------------------------

def _lo_hi(rng) -> Tuple[float, float]:
    """None / missing ends become -inf / +inf (adapted/detect/utils.py:16-26)."""
    if rng is None:
        return (-INF, INF)
    lo, hi = rng[0], rng[1]
    return (-INF if lo is None else float(lo), INF if hi is None else float(hi))


def range_is_empty(rng) -> bool:
    """adapted/detect/utils.py:29-36."""
    if rng is None:
        return True
    return (rng[0] == -INF and rng[1] == INF) or (rng[0] is None and rng[1] is None)


def flatten_streaming_config(p: Any) -> Dict[str, Any]:
    """StreamingConfig (local or the reference's) -> the scalar dict that fills ``adb_stream_config``."""
    d = {k: int(getattr(p, k)) for k in ("min_obs_adapter", "min_obs_post_loc", "search_increment_step", "pA_mean_window",
                                         "pA_var_window", "median_shift_window", "polyA_window")}
    for k in ("pA_mean_range", "pA_var_range", "median_shift_range", "polyA_med_range", "polyA_local_range"):
        d[k] = _lo_hi(getattr(p, k))
    return d


_METHOD_CODE = {"llr": 0, "cnn": 1, "start_peak": 2}


def flatten_config(spc: Any) -> Dict[str, Any]:
    """Flatten a (reference or local) SigProcConfig into the scalar dict that fills ``adb_config``."""
    core, llr, mvs, rr = spc.core, spc.llr_boundaries, spc.mvs_polya, spc.real_range
    cnn, ms, sp = spc.cnn_boundaries, spc.med_shift, spc.rna_start_peak
    out: Dict[str, Any] = dict(
        max_obs_trace=int(core.max_obs_trace), min_obs_adapter=int(core.min_obs_adapter),
        max_obs_adapter=int(core.max_obs_adapter), min_obs_polya=int(core.min_obs_polya),
        downscale_factor=int(core.downscale_factor),
        sig_norm_outlier_thresh=float(core.sig_norm_outlier_thresh),
        adapter_peak_prominence=float(llr.adapter_peak_prominence),
        adapter_peak_rel_height=float(llr.adapter_peak_rel_height),
        adapter_peak_width=int(llr.adapter_peak_width),
        polya_cand_k=int(cnn.polya_cand_k),
        fallback_to_llr_short_reads=int(bool(cnn.fallback_to_llr_short_reads)),
        mvs_detect_check=int(bool(mvs.mvs_detect_check)),
        mvs_detect_overwrite=int(bool(mvs.mvs_detect_overwrite)),
        search_window=int(mvs.search_window), pA_mean_window=int(mvs.pA_mean_window),
        pA_var_window=int(mvs.pA_var_window), median_shift_window=int(mvs.median_shift_window),
        polyA_window=int(mvs.polyA_window),
        pA_mean_range=_lo_hi(mvs.pA_mean_range), pA_var_range=_lo_hi(mvs.pA_var_range),
        median_shift_range=_lo_hi(mvs.median_shift_range), polyA_med_range=_lo_hi(mvs.polyA_med_range),
        polyA_local_range=_lo_hi(mvs.polyA_local_range),
        pA_mean_scale_range=_lo_hi(mvs.pA_mean_adapter_med_scale_range),
        pA_mean_range_empty=int(range_is_empty(mvs.pA_mean_range)),
        pA_mean_scale_range_empty=int(range_is_empty(mvs.pA_mean_adapter_med_scale_range)),
        detect_open_pores=int(bool(rr.detect_open_pores)), real_signal_check=int(bool(rr.real_signal_check)),
        mean_window=int(rr.mean_window), max_obs_local_range=int(rr.max_obs_local_range),
        mean_start_range=_lo_hi(rr.mean_start_range), mean_end_range=_lo_hi(rr.mean_end_range),
        local_range=_lo_hi(rr.local_range), adapter_mad_range=_lo_hi(rr.adapter_mad_range),
        detect_med_shift=int(bool(ms.detect_med_shift)), med_shift_window=int(ms.med_shift_window),
        med_shift_range=_lo_hi(ms.med_shift_range),
        sp_downscale_factor=int(sp.downscale_factor), start_peak_max_idx=int(sp.start_peak_max_idx),
        sp_offset1=int(sp.offset1), sp_offset2=int(sp.offset2), open_pore_pa=float(sp.open_pore_pa),
        primary_method=_METHOD_CODE[spc.primary_method],
        sig_preload_size=int(spc.sig_preload_size),
    )
    return out


def config_as_dict(spc: Any) -> Dict[str, Any]:
    """Nested plain dict (for fixtures / command.json style dumps)."""
    out = {}
    for name in _SECTIONS:
        sec = getattr(spc, name, None)
        if sec is None:  # the optional [streaming] section
            continue
        if is_dataclass(sec):
            out[name] = {f.name: getattr(sec, f.name) for f in fields(sec)}
        else:
            out[name] = dict(vars(sec))
    return out
