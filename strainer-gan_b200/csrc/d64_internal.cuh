// Internal hooks of csrc/d64.cu used by the training path (csrc/d64_train.cu); not part of the public C ABI.
#pragma once
#include "common.cuh"

extern "C" {
// One stage of the TRAINING forward in SG_CONV_FP16 arithmetic on a sg_d64 workspace: layer 1 = conv1 + LeakyReLU,
// layers 2..4 = raw conv output + batch statistics (scale | shift | mean | rstd | gamma, 512 floats each, to `bn_save`;
// running statistics parked in the workspace until sg_d64_train_commit_), layer 5 = head (logit, prob).
int sg_d64_train_layer_(const float* x, int64_t batch, const void* packed, void* workspace, int layer, float* prob,
                        float* logit, float* const* running_stats, float momentum, float eps, float* bn_save,
                        int32_t* status, void* stream);
// byte offsets of {status words, act1, act2, act3, act4} inside a SG_CONV_FP16 workspace for `batch` images
void sg_d64_workspace_offsets_(int64_t batch, size_t* act_off);
// commits the parked running statistics unless the status words report an fp16 overflow
int sg_d64_train_commit_(void* workspace, int64_t batch, float* const* running_stats, const int32_t* status, void* stream);
}
