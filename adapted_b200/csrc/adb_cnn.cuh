// CNN primary boundaries (placeholder until the conv kernels land).
#pragma once
#include "adb_common.cuh"
#include "adb_ctx.cuh"
#define ADB_CNN_NPARAMS 58882
static int cnn_primary_boundaries(adb_ctx *, const BatchDev &, const adb_config &, const float *, int *, cudaStream_t) {
    set_err("CNN primary method not built yet");
    return ADB_ERR_UNSUPPORTED;
}
extern "C" int adb_cnn_scores_host(adb_ctx *, const float *, int32_t, int32_t, const float *, float *) {
    set_err("CNN primary method not built yet");
    return ADB_ERR_UNSUPPORTED;
}
