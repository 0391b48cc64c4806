// adb_csv.cpp -- native writer of the boundary tables.
//
// Produces, byte for byte, the text that the reference's save_detected_boundaries (adapted/output.py:26-51) writes
// for a list of ReadResult objects: pandas.DataFrame([to_summary_dict() ...]) -> drop success / llr_trace
// (/ fail_reason) -> round(3) -> to_csv(index=False).  The source is the fixed-layout adb_record array of the CUDA
// library, so no DetectResults objects and no pandas are needed at GPU rates.  What has to be restated is pandas'
// per-column type inference (a column of ints with one None is a float64 column and prints "123.0"), numpy's
// round-half-even-on-the-scaled-value rounding, Python's shortest float repr, numpy's str() of 1-D int arrays (with
// its 75-column wrapping) and csv.QUOTE_MINIMAL.  Column order = dataclass field order of DetectResults
// (adapted/container_types.py:22-94) behind read_id (ReadResult.to_summary_dict, container_types.py:112-120).
#include <charconv>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

#include "../../include/adapted_b200.h"

namespace {

enum Kind { K_NONE = 0, K_INT, K_FLT, K_BOOL, K_STR, K_ARR };
struct Cell {
    Kind kind = K_NONE;
    int64_t i = 0;
    double f = 0.0;
    const char *s = nullptr;
    const int32_t *arr = nullptr;
    int n = 0;
    bool f32 = false;  // the reference holds this value as an np.float32 scalar
};

enum Col {
    C_READ_ID = 0, C_SIGNAL_LEN, C_PRELOADED,
    C_AD_START, C_AD_END, C_AD_LEN, C_AD_MEAN, C_AD_STD, C_AD_MED, C_AD_MAD,
    C_PA_START, C_PA_END, C_PA_LEN, C_PA_MEAN, C_PA_STD, C_PA_MED, C_PA_MAD, C_PA_TRUNC, C_PA_CAND,
    C_RNA_START, C_RNA_LEN, C_RNA_MEAN, C_RNA_STD, C_RNA_MED, C_RNA_MAD,
    C_SP_IDX, C_SP_PA, C_SP_NEXT_IDX, C_SP_NEXT_PA, C_SP_OP_IDX, C_SP_OP_TYPE,
    C_MED_SHIFT, C_LLR_AE, C_LLR_PE, C_CNN_AE, C_CNN_PE, C_SPK_AE, C_SPK_PE,
    C_LLR_AE_ADJ, C_LLR_PE_ADJ, C_LLR_EARLY_STOP, C_MVS_ADJ_IGNORED, C_MVS_TO_EARLY_STOP, C_MVS_AE,
    C_MVS_MEAN, C_MVS_VAR, C_MVS_MED, C_MVS_RANGE, C_MVS_SHIFT,
    C_REAL_START, C_REAL_END, C_REAL_RANGE, C_OPEN_PORES, C_LLR_LOG, C_FAIL_REASON, N_COLS
};

const char *const COL_NAMES[N_COLS] = {
    "read_id", "signal_len", "preloaded",
    "adapter_start", "adapter_end", "adapter_len", "adapter_mean", "adapter_std", "adapter_med", "adapter_mad",
    "polya_start", "polya_end", "polya_len", "polya_mean", "polya_std", "polya_med", "polya_mad", "polya_truncated",
    "polya_candidates",
    "rna_preloaded_start", "rna_preloaded_len", "rna_preloaded_mean", "rna_preloaded_std", "rna_preloaded_med",
    "rna_preloaded_mad",
    "start_peak_idx", "start_peak_pa", "start_peak_next_max_idx", "start_peak_next_max_pa",
    "start_peak_open_pore_idx", "start_peak_open_pore_type",
    "adapter_rna_median_shift", "llr_adapter_end", "llr_polya_end", "cnn_adapter_end", "cnn_polya_end",
    "start_peak_adapter_end", "start_peak_polya_end",
    "llr_adapter_end_adjust", "llr_polya_end_adjust", "llr_trace_early_stop_pos",
    "mvs_llr_polya_end_adjust_ignored", "mvs_llr_polya_end_to_early_stop", "mvs_adapter_end",
    "mvs_detect_mean_at_loc", "mvs_detect_var_at_loc", "mvs_detect_polya_med", "mvs_detect_polya_local_range",
    "mvs_detect_med_shift",
    "real_adapter_mean_start", "real_adapter_mean_end", "real_adapter_local_range", "open_pores", "llr_detect_log",
    "fail_reason"};

const char *fail_text(int code) {
    switch (code) {  // the strings of adapted/detect/combined.py:396-580 and of the exceptions caught at :225,304,350
    case ADB_FAIL_NO_ADAPTER: return "No adapter detected (primary)";
    case ADB_FAIL_ADAPTER_MAD: return "adapter MAD check failed";
    case ADB_FAIL_OPEN_PORE: return "Open pore too close to boundary";
    case ADB_FAIL_REAL_RANGE: return "Real signal check failed";
    case ADB_FAIL_NO_POLYA: return "No polya detected (primary)";
    case ADB_FAIL_MVS_NOT_ENOUGH: return "MVS polya check failed: not enough signal";
    case ADB_FAIL_MVS_NO_ADAPTER: return "No adapter detected in range (mvs_detect)";
    case ADB_FAIL_MED_SHIFT: return "Median shift check failed";
    case ADB_FAIL_EXC_PA_MEAN_RANGE: return "pA_mean_range is not specified";
    case ADB_FAIL_EXC_TOPK_NONE: return "'NoneType' object is not iterable";
    case ADB_FAIL_EXC_EMPTY_TRACE: return "attempt to get argmin of an empty sequence";
    case ADB_FAIL_EXC_SLICE_INDEX: return "slice indices must be integers or None or have an __index__ method";
    case ADB_FAIL_EXC_MAD_ZERO: return "MAD normalization failed: scale is 0";
    default: return nullptr;
    }
}

// fail_reason of one record (scratch owns composed strings)
const char *fail_reason(const adb_record &r, std::string &scratch) {
    const int code = r.fail_code;
    if (code == 0) return nullptr;
    if (code == ADB_FAIL_MVS_CHECKS) {  // combined.py:497-515
        static const char *const names[5] = {"mean ", "var ", "med ", "range ", "shift"};
        scratch = "MVS polya check failed: ";
        for (int i = 0; i < 5; i++)
            if (r.mvs_fail_mask >> i & 1) scratch += names[i];
        while (!scratch.empty() && scratch.back() == ' ') scratch.pop_back();
    } else {
        const char *t = fail_text(code);
        scratch = t ? t : "unknown failure";
    }
    if ((r.valid & ADB_V_FIELDS) && (r.valid & ADB_V_START_PEAK) && (r.valid & ADB_V_SP_OPEN_PORE) &&
        (r.sp_flag == 1 || r.sp_flag == 2)) {  // combined.py:340-347
        scratch += "+";
        scratch += r.sp_flag == 1 ? "open pore in adapter" : "potential concatemer adapter-only read";
    }
    return scratch.c_str();
}

Cell cell_int(int64_t v) { Cell c; c.kind = K_INT; c.i = v; return c; }
Cell cell_flt(double v) { Cell c; c.kind = K_FLT; c.f = v; return c; }
Cell cell_f32(double v) { Cell c; c.kind = K_FLT; c.f = v; c.f32 = true; return c; }
Cell cell_bool(bool v) { Cell c; c.kind = K_BOOL; c.i = v; return c; }
Cell cell_str(const char *s) { Cell c; if (s) { c.kind = K_STR; c.s = s; } return c; }
Cell cell_arr(const int32_t *a, int n) { Cell c; c.kind = K_ARR; c.arr = a; c.n = n; return c; }

// full open-pore lists of the records whose list does not fit the record (adb_open_pores_host): index[i] = -1 or the
// row k of record i, whose positions are pos[offs[k] .. offs[k + 1])
struct OpOver {
    const int32_t *index = nullptr;
    const int64_t *offs = nullptr;
    const int32_t *pos = nullptr;
};

// One cell of the table; mirrors adapted_b200/records.py:records_to_results field by field.
Cell get_cell(const adb_record &r, int col, int method, const char *read_id, const char *llr_log, const char *reason,
              const OpOver &ov = OpOver(), int rec_idx = -1) {
    if (col == C_READ_ID) return cell_str(read_id ? read_id : "");
    if (col == C_FAIL_REASON) return cell_str(reason);
    const uint32_t v = r.valid;
    if (!(v & ADB_V_FIELDS)) return Cell();  // DetectResults(success=False, fail_reason=str(e)): everything None
    const int seg_start[3] = {r.adapter_start, r.adapter_end, r.polya_end};
    const int seg_end[3] = {r.adapter_end, r.polya_end, r.preloaded};
    const uint32_t seg_bit[3] = {ADB_V_ADAPTER_STATS, ADB_V_POLYA_STATS, ADB_V_RNA_STATS};
    auto seg_stat = [&](int s, int q) { return (v & seg_bit[s]) ? cell_flt(r.stats[s][q]) : Cell(); };
    auto seg_len = [&](int s) { return (v & seg_bit[s]) ? cell_int((int64_t)seg_end[s] - seg_start[s]) : Cell(); };
    const bool polya_none = (v & ADB_V_POLYA_NONE) != 0;
    switch (col) {
    case C_SIGNAL_LEN: return cell_int(r.signal_len);
    case C_PRELOADED: return cell_int(r.preloaded);
    case C_AD_START: return cell_int(r.adapter_start);
    case C_AD_END: return cell_int(r.adapter_end);
    case C_AD_LEN: return seg_len(0);
    case C_AD_MEAN: case C_AD_STD: case C_AD_MED: case C_AD_MAD: return seg_stat(0, col - C_AD_MEAN);
    case C_PA_START: return cell_int(r.adapter_end);
    case C_PA_END: return polya_none ? Cell() : cell_int(r.polya_end);
    case C_PA_LEN: return seg_len(1);
    case C_PA_MEAN: case C_PA_STD: case C_PA_MED: case C_PA_MAD: return seg_stat(1, col - C_PA_MEAN);
    case C_PA_CAND: return (v & ADB_V_CAND) ? cell_arr(r.cand, r.n_cand) : Cell();
    case C_RNA_START: return polya_none ? Cell() : cell_int(r.polya_end);
    case C_RNA_LEN: return seg_len(2);
    case C_RNA_MEAN: case C_RNA_STD: case C_RNA_MED: case C_RNA_MAD: return seg_stat(2, col - C_RNA_MEAN);
    case C_SP_IDX: return (v & ADB_V_START_PEAK) ? cell_int(r.sp_idx) : Cell();
    case C_SP_PA: return (v & ADB_V_START_PEAK) ? cell_f32(r.sp_pa) : Cell();
    case C_SP_NEXT_IDX: return (v & ADB_V_START_PEAK) ? cell_int(r.sp_next_idx) : Cell();
    case C_SP_NEXT_PA: return (v & ADB_V_START_PEAK) ? cell_f32(r.sp_next_pa) : Cell();
    case C_SP_OP_IDX:
        return ((v & ADB_V_START_PEAK) && (v & ADB_V_SP_OPEN_PORE)) ? cell_int(r.sp_open_pore_idx) : Cell();
    case C_SP_OP_TYPE:
        if ((v & ADB_V_START_PEAK) && (v & ADB_V_SP_OPEN_PORE)) {
            if (r.sp_flag == 1) return cell_str("open pore in adapter");
            if (r.sp_flag == 2) return cell_str("potential concatemer adapter-only read");
        }
        return Cell();
    case C_MED_SHIFT: return (v & ADB_V_MED_SHIFT) ? cell_f32(r.med_shift) : Cell();
    case C_LLR_AE: return method == ADB_METHOD_LLR ? cell_int(r.primary_adapter_end) : Cell();
    case C_LLR_PE: return method == ADB_METHOD_LLR ? cell_int(r.primary_polya_end) : Cell();
    case C_CNN_AE: return method == ADB_METHOD_CNN ? cell_int(r.primary_adapter_end) : Cell();
    case C_CNN_PE: return method == ADB_METHOD_CNN ? cell_int(r.primary_polya_end) : Cell();
    case C_SPK_AE: return method == ADB_METHOD_START_PEAK ? cell_int(r.primary_adapter_end) : Cell();
    case C_SPK_PE: return method == ADB_METHOD_START_PEAK ? cell_int(r.primary_polya_end) : Cell();
    case C_MVS_ADJ_IGNORED: return cell_bool(false);
    case C_MVS_TO_EARLY_STOP: return cell_bool((v & ADB_V_TO_EARLY_STOP) != 0);
    case C_MVS_AE: return (v & ADB_V_MVS_ADAPTER_END) ? cell_int(r.mvs_adapter_end) : Cell();
    case C_MVS_MEAN: case C_MVS_VAR: case C_MVS_MED: case C_MVS_RANGE: case C_MVS_SHIFT:
        return (v & ADB_V_MVS) ? cell_flt(r.mvs[col - C_MVS_MEAN]) : Cell();
    case C_REAL_START: return (v & ADB_V_REAL_MEANS) ? cell_f32(r.real[0]) : Cell();
    case C_REAL_END: return (v & ADB_V_REAL_MEANS) ? cell_f32(r.real[1]) : Cell();
    case C_REAL_RANGE: return (v & ADB_V_REAL_RANGE) ? cell_flt(r.real[2]) : Cell();
    case C_OPEN_PORES:
        if (!(v & ADB_V_OPEN_PORES)) return Cell();
        if (r.n_open_pores > ADB_MAX_OPEN_PORES && ov.index && rec_idx >= 0 && ov.index[rec_idx] >= 0) {
            const int k = ov.index[rec_idx];
            return cell_arr(ov.pos + ov.offs[k], (int)(ov.offs[k + 1] - ov.offs[k]));
        }
        // (a list beyond the record's capacity without its overflow row is refused by adb_format_csv_ex)
        return cell_arr(r.open_pores, r.n_open_pores < ADB_MAX_OPEN_PORES ? r.n_open_pores : ADB_MAX_OPEN_PORES);
    case C_LLR_LOG: return cell_str(llr_log);
    default: return Cell();  // polya_truncated, llr_*_adjust, llr_trace_early_stop_pos: None in v0.2.4
    }
}

struct Out {
    char *p;
    int64_t cap, len = 0;
    void put(const char *s, size_t n) {
        if (len + (int64_t)n <= cap) memcpy(p + len, s, n);
        len += (int64_t)n;
    }
    void put(const std::string &s) { put(s.data(), s.size()); }
    void put(char c) { put(&c, 1); }
};

// A value that is a multiple of 1/1000 with at most 15 significant digits (what round(3) leaves): its shortest
// round-trip text is k / 1000 written out, so the digits come from integer arithmetic (libstdc++'s floating-point
// to_chars serialises threads).  Returns false when the value is not of that form.
static bool fmt_thousandths(double x, double max_k, std::string &s) {
    const double kd = std::nearbyint(x * 1000.0);
    if (!(std::fabs(kd) < max_k) || kd / 1000.0 != x) return false;
    long long k = (long long)kd;
    char buf[40];
    char *p = buf;
    if (std::signbit(x)) { *p++ = '-'; k = -k; }
    auto r = std::to_chars(p, buf + sizeof buf, k / 1000);
    p = r.ptr;
    *p++ = '.';
    const int frac = (int)(k % 1000);
    if (frac == 0) *p++ = '0';
    else {
        char d[3] = {(char)('0' + frac / 100), (char)('0' + (frac / 10) % 10), (char)('0' + frac % 10)};
        int n = 3;
        while (n > 1 && d[n - 1] == '0') n--;
        for (int i = 0; i < n; i++) *p++ = d[i];
    }
    s.assign(buf, p);
    return true;
}

// Python's repr(float) for the magnitudes that occur here (|x| < 1e16 after round(3)): shortest round-trip digits in
// fixed notation, always with a fractional part.
void fmt_float(double x, std::string &s) {
    if (std::isfinite(x) && fmt_thousandths(x, 1e15, s)) return;
    s.clear();
    if (std::isinf(x)) { s = x < 0 ? "-inf" : "inf"; return; }
    char buf[400];
    const double ax = std::fabs(x);
    if (ax != 0.0 && (ax >= 1e16 || ax < 1e-4)) {
        // repr switches to the exponent form outside [1e-4, 1e16): d.ddde+XX with at least two exponent digits
        auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
        s.assign(buf, r.ptr);
        const size_t e = s.find('e');
        if (e != std::string::npos) {
            std::string mant = s.substr(0, e), ex = s.substr(e + 1);
            char sign = '+';
            if (!ex.empty() && (ex[0] == '+' || ex[0] == '-')) { sign = ex[0]; ex.erase(0, 1); }
            while (ex.size() > 2 && ex[0] == '0') ex.erase(0, 1);
            if (ex.size() < 2) ex.insert(0, 2 - ex.size(), '0');
            s = mant + "e" + sign + ex;
        }
        return;
    }
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);
    s.assign(buf, r.ptr);
    if (s.find('.') == std::string::npos) s += ".0";
}

// the same for a float32 column (all cells np.float32, no None): pandas keeps float32, numpy rounds in float32 and
// the text is the shortest float32 round-trip
void fmt_float32(float x, std::string &s) {
    // (float32: six significant digits are unambiguous)
    if (std::isfinite(x) && std::fabs(x) < 999.0f) {
        const float kf = std::nearbyintf(x * 1000.0f);
        if (kf / 1000.0f == x && fmt_thousandths((double)kf / 1000.0, 1e6, s)) return;
    }
    s.clear();
    if (std::isinf(x)) { s = x < 0 ? "-inf" : "inf"; return; }
    char buf[128];
    const float ax = std::fabs(x);
    if (ax != 0.0f && (ax >= 1e16f || ax < 1e-4f)) {
        auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::scientific);
        s.assign(buf, r.ptr);
        const size_t e = s.find('e');
        if (e != std::string::npos) {
            std::string mant = s.substr(0, e), ex = s.substr(e + 1);
            char sign = '+';
            if (!ex.empty() && (ex[0] == '+' || ex[0] == '-')) { sign = ex[0]; ex.erase(0, 1); }
            while (ex.size() > 2 && ex[0] == '0') ex.erase(0, 1);
            if (ex.size() < 2) ex.insert(0, 2 - ex.size(), '0');
            s = mant + "e" + sign + ex;
        }
        return;
    }
    auto r = std::to_chars(buf, buf + sizeof buf, x, std::chars_format::fixed);
    s.assign(buf, r.ptr);
    if (s.find('.') == std::string::npos) s += ".0";
}

float round3f(float x) {
    if (!std::isfinite(x)) return x;
    volatile float t = x * 1000.0f;  // individually rounded float32 steps
    t = std::nearbyintf(t);
    return t / 1000.0f;
}

// numpy.round(x, 3) on float64: rint(x * 1000) / 1000 (numpy/_core/src/multiarray/calculation.c), which is what
// DataFrame.round applies to every float64 column
double round3(double x) {
    if (!std::isfinite(x)) return x;
    return std::nearbyint(x * 1000.0) / 1000.0;
}

// str(numpy int array), 1-D: elements right-aligned to the widest, one blank between, wrapped like
// numpy/_core/arrayprint.py:_formatArray/_extendLine with linewidth 75 (continuation lines indented by one blank)
void fmt_int_array(const int32_t *a, int n, std::string &s) {
    s.clear();
    if (n <= 0) { s = "[]"; return; }
    size_t width = 0;
    char buf[16];
    for (int i = 0; i < n; i++) {
        const size_t l = (size_t)(std::to_chars(buf, buf + sizeof buf, a[i]).ptr - buf);
        if (l > width) width = l;
    }
    const size_t elem_width = 75 - 1;
    std::string line = " ", word;
    for (int i = 0; i < n; i++) {
        const int l = (int)(std::to_chars(buf, buf + sizeof buf, a[i]).ptr - buf);
        word.assign(width - (size_t)l, ' ');
        word.append(buf, (size_t)l);
        if (line.size() + word.size() > elem_width && line.size() > 1) {
            while (!line.empty() && line.back() == ' ') line.pop_back();
            s += line;
            s += "\n";
            line = " ";
        }
        line += word;
        if (i + 1 < n) line += " ";
    }
    s += line;
    s = "[" + s.substr(1) + "]";
}

void put_field(Out &o, const char *s, size_t n) {  // csv.QUOTE_MINIMAL with the default dialect
    bool quote = false;
    for (size_t i = 0; i < n; i++)
        if (s[i] == ',' || s[i] == '"' || s[i] == '\n' || s[i] == '\r') { quote = true; break; }
    if (!quote) { o.put(s, n); return; }
    o.put('"');
    for (size_t i = 0; i < n; i++) {
        if (s[i] == '"') o.put('"');
        o.put(s[i]);
    }
    o.put('"');
}

}  // namespace

// rows [k0, k1) of the table body into `o` (column types already inferred)
static void format_rows(Out &o, int k0, int k1, int n_cols, const adb_record *recs, const int32_t *sel, const char *const *read_ids,
                        int primary_method, const char *llr_detect_log, bool save_fail_reasons, const uint8_t *has_none,
                        const uint8_t *not_f32, const OpOver &ov) {
    std::string text, scratch;
    char buf[32];
    for (int k = k0; k < k1; k++) {
        const int idx = sel ? sel[k] : k;
        const adb_record &r = recs[idx];
        const char *id = read_ids ? read_ids[idx] : nullptr;
        const char *reason = save_fail_reasons ? fail_reason(r, scratch) : nullptr;
        for (int c = 0; c < n_cols; c++) {
            if (c) o.put(',');
            const Cell cell = get_cell(r, c, primary_method, id, llr_detect_log, reason, ov, idx);
            switch (cell.kind) {
            case K_NONE: break;
            case K_INT:
                // ints + None in one column -> float64 column (NaN for None): "123.0"
                if (has_none[c]) { fmt_float((double)cell.i, text); o.put(text); }
                else {
                    auto res = std::to_chars(buf, buf + sizeof buf, (long long)cell.i);
                    o.put(buf, (size_t)(res.ptr - buf));
                }
                break;
            case K_FLT:
                if (std::isnan(cell.f)) break;  // na_rep = ""
                if (!not_f32[c]) fmt_float32(round3f((float)cell.f), text);
                else fmt_float(round3(cell.f), text);
                o.put(text);
                break;
            case K_BOOL: o.put(cell.i ? "True" : "False", cell.i ? 4 : 5); break;
            case K_STR: put_field(o, cell.s, strlen(cell.s)); break;
            case K_ARR: fmt_int_array(cell.arr, cell.n, text); put_field(o, text.data(), text.size()); break;
            }
        }
        o.put('\n');
    }
}

extern "C" int64_t adb_format_csv_ex(const adb_record *recs, const int32_t *sel, int32_t n_sel,
                                     const char *const *read_ids, int32_t primary_method, const char *llr_detect_log,
                                     int32_t save_fail_reasons, const int32_t *op_index, const int64_t *op_offsets,
                                     const int32_t *op_pos, char *out, int64_t cap) {
    if ((n_sel > 0 && !recs) || n_sel < 0 || (cap > 0 && !out)) return ADB_ERR_ARG;
    OpOver ov;
    if (op_index && op_offsets && op_pos) { ov.index = op_index; ov.offs = op_offsets; ov.pos = op_pos; }
    // never truncate silently: a list longer than the record keeps needs its overflow row
    for (int k = 0; k < n_sel; k++) {
        const int idx = sel ? sel[k] : k;
        const adb_record &r = recs[idx];
        if ((r.valid & ADB_V_FIELDS) && (r.valid & ADB_V_OPEN_PORES) && r.n_open_pores > ADB_MAX_OPEN_PORES) {
            if (!ov.index || ov.index[idx] < 0) return ADB_ERR_OVERFLOW;
            const int row = ov.index[idx];
            if (ov.offs[row + 1] - ov.offs[row] != r.n_open_pores) return ADB_ERR_OVERFLOW;
        }
    }
    Out o{out, cap};
    if (n_sel == 0) {  // pd.DataFrame([]).round(3).to_csv(index=False) writes one empty line
        o.put('\n');
        return o.len;
    }
    const int n_cols = save_fail_reasons ? N_COLS : N_COLS - 1;
    // pass 1: per column, is there a None (pandas' maybe_convert_objects turns an int column with a None into float64)
    // and are all cells np.float32 scalars (the column then stays float32).  Only the validity bits and the method
    // decide that, not the values: one cheap probe per row.
    std::vector<uint8_t> has_none(n_cols, 0), not_f32(n_cols, 0);
    {
        std::string scratch;
        uint32_t seen_valid_masks[64];
        int n_seen = 0;
        for (int k = 0; k < n_sel; k++) {
            const adb_record &r = recs[sel ? sel[k] : k];
            // the None / float32 pattern of a row is a function of (valid bits, sp_flag in {1, 2}, fail_code != 0)
            const uint32_t key = r.valid ^ ((uint32_t)(r.sp_flag == 1 || r.sp_flag == 2) << 30) ^ ((uint32_t)(r.fail_code != 0) << 31);
            bool known = false;
            for (int t = 0; t < n_seen; t++) known |= (seen_valid_masks[t] == key);
            if (known) continue;
            if (n_seen < 64) seen_valid_masks[n_seen++] = key;
            const char *reason = save_fail_reasons ? fail_reason(r, scratch) : nullptr;
            for (int c = 0; c < n_cols; c++) {
                const Cell cell = get_cell(r, c, primary_method, "", llr_detect_log, reason);
                if (cell.kind == K_NONE) has_none[c] = 1;
                if (!(cell.kind == K_FLT && cell.f32)) not_f32[c] = 1;
            }
        }
    }
    for (int c = 0; c < n_cols; c++) {
        if (c) o.put(',');
        o.put(COL_NAMES[c], strlen(COL_NAMES[c]));
    }
    o.put('\n');
    // pass 2: the rows; large tables are formatted by several threads into private buffers and stitched in order
    unsigned hw = std::thread::hardware_concurrency();
    int n_thr = (n_sel >= 8192) ? (int)std::min<unsigned>(hw ? hw : 1u, 16u) : 1;
    if (n_thr <= 1) {
        format_rows(o, 0, n_sel, n_cols, recs, sel, read_ids, primary_method, llr_detect_log, save_fail_reasons != 0,
                    has_none.data(), not_f32.data(), ov);
        return o.len;
    }
    std::vector<std::vector<char>> parts(n_thr);
    std::vector<int64_t> lens(n_thr, 0);
    std::vector<std::thread> pool;
    const int per = (n_sel + n_thr - 1) / n_thr;
    for (int t = 0; t < n_thr; t++) {
        pool.emplace_back([&, t]() {
            const int k0 = std::min(n_sel, t * per), k1 = std::min(n_sel, k0 + per);
            parts[t].resize((size_t)(k1 - k0) * 512 + 64);
            for (;;) {
                Out po{parts[t].data(), (int64_t)parts[t].size()};
                format_rows(po, k0, k1, n_cols, recs, sel, read_ids, primary_method, llr_detect_log, save_fail_reasons != 0,
                            has_none.data(), not_f32.data(), ov);
                lens[t] = po.len;
                if (po.len <= (int64_t)parts[t].size()) break;
                parts[t].resize((size_t)po.len);
            }
        });
    }
    for (auto &th : pool) th.join();
    for (int t = 0; t < n_thr; t++) o.put(parts[t].data(), (size_t)lens[t]);
    return o.len;
}

extern "C" int64_t adb_format_csv(const adb_record *recs, const int32_t *sel, int32_t n_sel,
                                  const char *const *read_ids, int32_t primary_method, const char *llr_detect_log,
                                  int32_t save_fail_reasons, char *out, int64_t cap) {
    return adb_format_csv_ex(recs, sel, n_sel, read_ids, primary_method, llr_detect_log, save_fail_reasons, nullptr,
                             nullptr, nullptr, out, cap);
}
