// start-peak primary (placeholder).
#pragma once
#include "adb_common.cuh"
#include "adb_ctx.cuh"
static int start_peak_primary(adb_ctx *, const BatchDev &, const adb_config &, int *, adb_record *, int *, cudaStream_t) {
    set_err("start-peak primary method not built yet");
    return ADB_ERR_UNSUPPORTED;
}
static int start_peak_finish(adb_ctx *, const BatchDev &, const adb_config &, adb_record *, cudaStream_t) { return ADB_OK; }
