// Token pipeline of the training loader on the device: data/loader.py:132-195 (Loader.__getitem__ after the
// random draws) with data/data_processing.py:225-247 (transpose, tensor_to_ind_tensor) for a whole batch in one
// launch.  The reference does this per sample in Python: a loop over every (event, value) tuple for the
// transposition and a dict lookup per tuple for the token id.
//
//   tuples[b]  -> transpose pitches of transposable events by n_transpose[b] when the result stays in
//                 [min_pitch, max_pitch]                                         (data_processing.py:225-232)
//              -> token id through the tuple -> index table                      (data_processing.py:234-247)
//              -> crop [start, start + input_len + 1) when the sample does not start at a bar, else keep all
//                 (loader.py:141-154; <START>, <CLS>, emotion tokens are the caller-built `prefix`, :143-149,156-160,
//                 164-170)
//              -> trim to input_len + 1, pad with the pad id                     (loader.py:176-182)
//              -> input = seq[:-1], target = seq[1:] left-padded by `target_left_pad` pads (loader.py:185-193)
// HBM-bound integer work: 4 bytes read and 16 bytes written per token; one thread block per sample.
#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

constexpr int TP_THREADS = 256;

__global__ void __launch_bounds__(TP_THREADS) token_pipeline_kernel(me_token_pipeline_args a) {
  const int b = blockIdx.x;
  const int n = a.n_events[b];
  const int shift = a.n_transpose ? a.n_transpose[b] : 0;
  const int start = a.start ? a.start[b] : -1;
  const int npre = a.n_prefix ? a.n_prefix[b] : 0;
  const bool cropped = start >= 0;
  const int first = cropped ? start : 0;
  const int avail = cropped ? min(a.input_len + 1, n - first) : n;  // tokens taken from the tuples
  const int16_t* ev = a.events + static_cast<int64_t>(b) * a.max_events * 2;
  int64_t* in_row = a.input + static_cast<int64_t>(b) * a.input_len;
  int64_t* tg_row = a.target ? a.target + static_cast<int64_t>(b) * (a.input_len + a.target_left_pad) : nullptr;
  int bad = 0;
  if (tg_row)
    for (int p = threadIdx.x; p < a.target_left_pad; p += TP_THREADS) tg_row[p] = a.pad_token;
  for (int p = threadIdx.x; p <= a.input_len; p += TP_THREADS) {
    int64_t tok = a.pad_token;
    if (p < npre) {
      tok = a.prefix[b * ME_TP_MAX_PREFIX + p];
    } else if (p - npre < avail) {
      const int src = first + p - npre;
      const int e = ev[2 * src];
      int v = ev[2 * src + 1];
      if (e >= 0 && e < a.n_event_types) {
        if (shift != 0 && a.transposable[e] && v + shift <= a.max_pitch && v + shift >= a.min_pitch) v += shift;
        const int id = (v >= 0 && v < a.n_values) ? a.lut[e * a.n_values + v] : -1;
        if (id < 0) bad = 1;
        tok = id < 0 ? a.pad_token : id;
      } else {
        bad = 1;
      }
    }
    if (p < a.input_len) in_row[p] = tok;
    if (tg_row && p >= 1) tg_row[a.target_left_pad + p - 1] = tok;
  }
  if (a.status) {
    const int any = __syncthreads_or(bad);
    if (threadIdx.x == 0) a.status[b] = any ? 1 : 0;
  }
}

}  // namespace me

extern "C" int me_sizeof_token_pipeline_args(void) { return static_cast<int>(sizeof(me_token_pipeline_args)); }

extern "C" int me_token_pipeline(const me_token_pipeline_args* a) {
  ME_CHECK(a != nullptr, "me_token_pipeline: NULL args");
  ME_CHECK(a->B > 0 && a->max_events > 0 && a->input_len > 0, "me_token_pipeline: bad dims");
  ME_CHECK(a->events && a->n_events && a->lut && a->transposable && a->input, "me_token_pipeline: NULL pointer");
  ME_CHECK(a->n_event_types > 0 && a->n_values > 0, "me_token_pipeline: empty tuple table");
  ME_CHECK(a->target_left_pad >= 0 && (a->n_prefix == nullptr || a->prefix != nullptr), "me_token_pipeline: bad prefix/pad arguments");
  me::token_pipeline_kernel<<<a->B, me::TP_THREADS, 0, static_cast<cudaStream_t>(a->stream)>>>(*a);
  ME_LAUNCH_CHECK();
  return 0;
}
