// Library plumbing: error text, device query, TMA descriptor construction, attention dispatch.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"
#include "../../include/midi_emotion_b200.h"

namespace me {

static thread_local char g_err[512] = "";
unsigned long long g_launch_count = 0;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  // off unless ME_PDL=1: measured on the decode step (B = 256, t = 1024) 2.56 ms with it against 2.41 ms without --
  // the early-launched CTAs of the next kernel take SM slots from the tail of the running one
  static const bool on = [] { const char* e = getenv("ME_PDL"); return e != nullptr && e[0] == '1'; }();
  return on;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point: the library links only cudart.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_elems,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = encode_fn();
  ME_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  ME_CHECK(box_inner * 2 <= 128 && box_outer <= 256, "tensor map box %ux%u too large", box_inner, box_outer);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {pitch_elems * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ME_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(2d) failed: %d (inner=%llu outer=%llu pitch=%llu box=%ux%u base=%p)",
           static_cast<int>(r), (unsigned long long)inner, (unsigned long long)outer,
           (unsigned long long)pitch_elems, box_inner, box_outer, base);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2,
                      uint64_t pitch1_elems, uint64_t pitch2_elems, uint32_t b0, uint32_t b1, uint32_t b2) {
  EncodeTiledFn fn = encode_fn();
  ME_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[3] = {d0, d1, d2};
  cuuint64_t strides[2] = {pitch1_elems * 2, pitch2_elems * 2};
  cuuint32_t box[3] = {b0, b1, b2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ME_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(3d) failed: %d", static_cast<int>(r));
  return 0;
}

int make_tmap_nd_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                      const uint64_t* strides_elems, const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  ME_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point not available");
  ME_CHECK(rank >= 1 && rank <= 5, "tensor map rank %d", rank);
  ME_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map base must be 16-byte aligned");
  cuuint64_t d[5], st[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
  for (int i = 0; i + 1 < rank; ++i) {
    ME_CHECK(strides_elems[i] % 8 == 0, "tensor map stride %d (%llu elements) must be a multiple of 8", i,
             (unsigned long long)strides_elems[i]);
    st[i] = strides_elems[i] * 2;
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), d, st, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  ME_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(rank %d) failed: %d", rank, static_cast<int>(r));
  return 0;
}

int launch_attn_fwd_simt(const me_attn_args* a);
int launch_attn_bwd_simt(const me_attn_bwd_args* a);
int launch_attn_fwd_tc(const me_attn_args* a);
int launch_attn_bwd_tc(const me_attn_bwd_args* a);

}  // namespace me

extern "C" const char* me_last_error(void) { return me::g_err; }
extern "C" int me_version(void) { return 100; }
extern "C" unsigned long long me_launch_count(void) { return me::g_launch_count; }

extern "C" int me_device_is_sm100(void) {
  static int cached = -1;
  if (cached < 0) {
    int dev = 0, major = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return 0;
    cached = (major == 10) ? 1 : 0;
  }
  return cached;
}

extern "C" int me_attention_forward(const me_attn_args* a) {
  using namespace me;
  ME_CHECK(a != nullptr, "me_attention_forward: NULL args");
  if (a->impl == ME_ATTN_TENSOR) return launch_attn_fwd_tc(a);
  return launch_attn_fwd_simt(a);
}

extern "C" int me_attention_backward(const me_attn_bwd_args* a) {
  using namespace me;
  ME_CHECK(a != nullptr, "me_attention_backward: NULL args");
  if (a->f.impl == ME_ATTN_TENSOR) return launch_attn_bwd_tc(a);
  return launch_attn_bwd_simt(a);
}

// Binding self-check: hosts mirror the argument structs (ctypes / cgo / JNI) and compare sizes at load.
extern "C" int me_sizeof_attn_args(void) { return static_cast<int>(sizeof(me_attn_args)); }
extern "C" int me_sizeof_attn_bwd_args(void) { return static_cast<int>(sizeof(me_attn_bwd_args)); }
extern "C" int me_sizeof_layer_args(void) { return static_cast<int>(sizeof(me_layer_args)); }
extern "C" int me_sizeof_layer_bwd_args(void) { return static_cast<int>(sizeof(me_layer_bwd_args)); }
extern "C" int me_sizeof_decode_layer_args(void) { return static_cast<int>(sizeof(me_decode_layer_args)); }
extern "C" int me_sizeof_sample_args(void) { return static_cast<int>(sizeof(me_sample_args)); }

// ---------------------------------------------------------------------------------------------
// Live kernel timing for bench.py: while enabled, every tcgen05 GEMM launch (class 0) and every tensor-core
// attention kernel launch (class 1 forward, 2 backward key side, 3 backward query side) is bracketed by CUDA
// events on its own stream; me_profile_collect_class() sums durations and algorithmic FLOPs per class.
// ---------------------------------------------------------------------------------------------
namespace me {
struct ProfSlot { cudaEvent_t a, b; double flops; int cls; };
static ProfSlot* g_prof = nullptr;
static int g_prof_cap = 0, g_prof_n = 0, g_prof_on = 0;

cudaEvent_t prof_begin(double flops, cudaStream_t st, int cls) {
  if (!g_prof_on || g_prof_n >= g_prof_cap) return nullptr;
  ProfSlot& s = g_prof[g_prof_n];
  s.flops = flops;
  s.cls = cls;
  cudaEventRecord(s.a, st);
  return s.b;
}
void prof_end(cudaEvent_t e, cudaStream_t st) {
  if (!e) return;
  cudaEventRecord(e, st);
  ++g_prof_n;
}
}  // namespace me

extern "C" int me_profile_enable(int capacity) {
  using namespace me;
  if (capacity > g_prof_cap) {
    ProfSlot* n = static_cast<ProfSlot*>(realloc(g_prof, sizeof(ProfSlot) * capacity));
    ME_CHECK(n != nullptr, "me_profile_enable: out of memory");
    g_prof = n;
    for (int i = g_prof_cap; i < capacity; ++i) {
      ME_CUDA(cudaEventCreate(&g_prof[i].a));
      ME_CUDA(cudaEventCreate(&g_prof[i].b));
    }
    g_prof_cap = capacity;
  }
  g_prof_n = 0;
  g_prof_on = capacity > 0 ? 1 : 0;
  return 0;
}

extern "C" int me_profile_collect_class(int cls, double* total_ms, double* total_flops, int* launches) {
  using namespace me;
  double ms = 0, fl = 0;
  int n = 0;
  for (int i = 0; i < g_prof_n; ++i) {
    if (g_prof[i].cls != cls) continue;
    ME_CUDA(cudaEventSynchronize(g_prof[i].b));
    float t = 0.f;
    ME_CUDA(cudaEventElapsedTime(&t, g_prof[i].a, g_prof[i].b));
    ms += t;
    fl += g_prof[i].flops;
    ++n;
  }
  if (total_ms) *total_ms = ms;
  if (total_flops) *total_flops = fl;
  if (launches) *launches = n;
  return 0;
}

extern "C" int me_profile_collect(double* total_ms, double* total_flops, int* launches) {
  const int rc = me_profile_collect_class(0, total_ms, total_flops, launches);
  me::g_prof_n = 0;
  me::g_prof_on = 0;
  return rc;
}
