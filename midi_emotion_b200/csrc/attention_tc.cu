// Tensor-core (tcgen05) relative attention -- placeholder until the kernels land; the SIMT path is
// the only implementation for now, and asking for ME_ATTN_TENSOR is an error (never a fallback).
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {
int launch_attn_fwd_tc(const me_attn_args*) {
  set_error("me_attention_forward: ME_ATTN_TENSOR is not built in this version");
  return 1;
}
int launch_attn_bwd_tc(const me_attn_bwd_args*) {
  set_error("me_attention_backward: ME_ATTN_TENSOR is not built in this version");
  return 1;
}
}  // namespace me
