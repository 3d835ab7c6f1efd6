// Host runtime glue: thread-local error string and TMA tensor-map encoding through the driver entry
// point resolved at run time (no link-time dependency on libcuda: the library must load on GPU-less hosts).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <vector>

#include "common.cuh"
#include "ops.h"

namespace gyre {

static thread_local char g_err[1024] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

const char* last_error() { return g_err; }

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int encode_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const uint32_t* elem_strides, bool swizzle128) {
  return encode_tmap_f16_sw(out, base, rank, dims, strides_bytes, box, elem_strides, swizzle128 ? 128 : 0);
}

int encode_tmap_f16_sw(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, const uint32_t* elem_strides,
                       int swizzle_bytes) {
  EncodeTiledFn fn = get_encode();
  const CUtensorMapSwizzle sw = swizzle_bytes == 128  ? CU_TENSOR_MAP_SWIZZLE_128B
                                : swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B
                                : swizzle_bytes == 32 ? CU_TENSOR_MAP_SWIZZLE_32B
                                                      : CU_TENSOR_MAP_SWIZZLE_NONE;
  GYRE_REQUIRE(fn != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) {
    d[i] = dims[i];
    b[i] = box[i];
    e[i] = elem_strides[i];
  }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), d, s,
                  b, e, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed: %d (rank %d dims %llu %llu box %u %u stride0 %llu)", (int)r, rank,
                   (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0), box[0],
                   rank > 1 ? box[1] : 0, (unsigned long long)(rank > 1 ? strides_bytes[0] : 0));
    return -3;
  }
  return 0;
}


// ------------------------------------------------------------------ tunables
static const char* kTunableNames[TUNE_COUNT] = {"ATT_VARIANT", "PDL", "GELU_FAST", "GN_CHUNKS", "UPCONV_FOLD",
                                                "CTX_KV_CACHE", "XATTN", "GN_PHASE", "MCAST", "ATT_D128", "STREAMK", "FORCE_BN", "GEMM_STAGES", "DEBUG", "LN_SUB", "GN_THREADS", "LN_FUSE", "CFG_SHARE", "GN_FUSE", "SK_MIN"};
// defaults = the configuration measured fastest on B200 (profiles/); 0 restores the plain round-1 kernels
static const int kTunableDefaults[TUNE_COUNT] = {2002, 1, 1, 32, 1, 1, 1, 0, 2, 1, 1, 0, 0, 0, 1, 256, 1, 1, 1, 14000};
static std::atomic<int> g_tunables[TUNE_COUNT];
static std::once_flag g_tunables_once;

static void init_tunables() {
  std::call_once(g_tunables_once, [] {
    for (int i = 0; i < TUNE_COUNT; ++i) {
      int v = kTunableDefaults[i];
      char name[64];
      snprintf(name, sizeof(name), "GYRE_B200_%s", kTunableNames[i]);
      if (const char* e = getenv(name)) v = atoi(e);
      g_tunables[i].store(v);
    }
  });
}

int& coop_pdl_state() {
  static int state = 0;
  return state;
}

int tunable(int id) {
  init_tunables();
  return (id >= 0 && id < TUNE_COUNT) ? g_tunables[id].load(std::memory_order_relaxed) : 0;
}

int set_tunable_by_name(const char* name, int value) {
  init_tunables();
  GYRE_REQUIRE(name != nullptr, "set_tunable: null name");
  for (int i = 0; i < TUNE_COUNT; ++i)
    if (strcmp(name, kTunableNames[i]) == 0) {
      g_tunables[i].store(value);
      return 0;
    }
  set_last_error("set_tunable: unknown tunable '%s'", name);
  return -2;
}

int get_tunable_by_name(const char* name, int* value) {
  init_tunables();
  GYRE_REQUIRE(name != nullptr && value != nullptr, "get_tunable: null argument");
  for (int i = 0; i < TUNE_COUNT; ++i)
    if (strcmp(name, kTunableNames[i]) == 0) {
      *value = g_tunables[i].load();
      return 0;
    }
  set_last_error("get_tunable: unknown tunable '%s'", name);
  return -2;
}

// ------------------------------------------------------------------ launch counter + per-family event profiler
namespace prof {

static std::atomic<unsigned long long> g_launches{0};
static std::atomic<int> g_enabled{0};
struct Rec {
  int family;
  double flops, bytes;
  cudaEvent_t e0, e1;
};
static std::mutex g_mu;
static std::vector<Rec> g_recs;
static std::vector<cudaEvent_t> g_pool;

static cudaEvent_t get_event() {
  if (!g_pool.empty()) {
    cudaEvent_t e = g_pool.back();
    g_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}

unsigned long long launch_count() { return g_launches.load(); }
void enable(int on) { g_enabled.store(on); }
bool enabled() { return g_enabled.load() != 0; }

void reset() {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& r : g_recs) {
    g_pool.push_back(r.e0);
    g_pool.push_back(r.e1);
  }
  g_recs.clear();
}

Scope::Scope(int family, double flops, double bytes, cudaStream_t st, int kernels)
    : family_(family), flops_(flops), bytes_(bytes), st_(st), e0_(nullptr) {
  g_launches.fetch_add(static_cast<unsigned long long>(kernels));
  if (enabled()) {
    std::lock_guard<std::mutex> lk(g_mu);
    e0_ = get_event();
    cudaEventRecord(static_cast<cudaEvent_t>(e0_), st_);
  }
}

Scope::~Scope() {
  if (e0_) {
    std::lock_guard<std::mutex> lk(g_mu);
    cudaEvent_t e1 = get_event();
    cudaEventRecord(e1, st_);
    g_recs.push_back(Rec{family_, flops_, bytes_, static_cast<cudaEvent_t>(e0_), e1});
  }
}

// Sum over the family's launches of max(flops / peak_flops, bytes / peak_bw): the time the launches would take if each
// ran exactly on whichever of the two roofs binds it (a family mixes tensor-bound and HBM-bound shapes).
int read_roofline_ms(int family, double peak_tflops, double peak_gbs, double* ideal_ms) {
  GYRE_REQUIRE(peak_tflops > 0 && peak_gbs > 0 && ideal_ms, "prof_read_roofline: bad arguments");
  std::lock_guard<std::mutex> lk(g_mu);
  double t = 0;
  for (auto& r : g_recs) {
    if (r.family != family) continue;
    const double tf = r.flops / (peak_tflops * 1e12), tb = r.bytes / (peak_gbs * 1e9);
    t += (tf > tb ? tf : tb) * 1e3;
  }
  *ideal_ms = t;
  return 0;
}

int read(int family, unsigned long long* count, double* ms, double* flops, double* bytes) {
  GYRE_CHECK_CUDA(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(g_mu);
  unsigned long long c = 0;
  double t = 0, f = 0, b = 0;
  for (auto& r : g_recs) {
    if (r.family != family) continue;
    float e = 0.f;
    if (cudaEventElapsedTime(&e, r.e0, r.e1) == cudaSuccess) t += e;
    ++c;
    f += r.flops;
    b += r.bytes;
  }
  if (count) *count = c;
  if (ms) *ms = t;
  if (flops) *flops = f;
  if (bytes) *bytes = b;
  return 0;
}

}  // namespace prof
}  // namespace gyre
