#include "common.h"

#include <atomic>
#include <cstdarg>
#include <cstring>
#include <map>
#include <mutex>
#include <unordered_map>
#include <string>
#include <vector>

namespace w2v {

namespace {
thread_local char g_err[1024] = "";
std::atomic<long long> g_launches{0};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
std::once_flag g_encode_once;
int g_num_sms = 0;
}  // namespace

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
      g_num_sms = n;
    else
      g_num_sms = 148;
  }
  return g_num_sms;
}

static int make_tmap_2d_impl(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                             uint64_t row_stride_elems, uint32_t box0, uint32_t box1, CUtensorMapSwizzle sw,
                             int elem_bytes = 2);

int make_tmap_2d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                      uint64_t row_stride_elems, uint32_t box0, uint32_t box1) {
  return make_tmap_2d_impl(out, base, dim0, dim1, row_stride_elems, box0, box1, CU_TENSOR_MAP_SWIZZLE_128B);
}
int make_tmap_2d_f32(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                     uint64_t row_stride_elems, uint32_t box0, uint32_t box1) {
  return make_tmap_2d_impl(out, base, dim0, dim1, row_stride_elems, box0, box1, CU_TENSOR_MAP_SWIZZLE_128B, 4);
}
int make_tmap_2d_bf16_sw64(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                           uint64_t row_stride_elems, uint32_t box0, uint32_t box1) {
  return make_tmap_2d_impl(out, base, dim0, dim1, row_stride_elems, box0, box1, CU_TENSOR_MAP_SWIZZLE_64B);
}

// Tensor maps are pure functions of (address, shape, strides, box, swizzle, type), and the forward pass asks for
// the same ~400 of them every step (fixed workspace, fixed shapes): a small per-thread cache replaces the driver
// call (cuTensorMapEncodeTiled) by a hash lookup. Entries never go stale: a map encodes no memory contents.
namespace {
struct TmapKey {
  uint64_t base, dim0, dim1, dim2, s1, s2;
  uint32_t box0, box1, sw, elem;
  bool operator==(const TmapKey& o) const {
    return base == o.base && dim0 == o.dim0 && dim1 == o.dim1 && dim2 == o.dim2 && s1 == o.s1 && s2 == o.s2 &&
           box0 == o.box0 && box1 == o.box1 && sw == o.sw && elem == o.elem;
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    uint64_t h = 0x9E3779B97F4A7C15ull;
    auto mix = [&](uint64_t v) { h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); };
    mix(k.base); mix(k.dim0); mix(k.dim1); mix(k.dim2); mix(k.s1); mix(k.s2);
    mix(((uint64_t)k.box0 << 32) | k.box1); mix(((uint64_t)k.sw << 32) | k.elem);
    return (size_t)h;
  }
};
thread_local std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> t_tmap_cache;
constexpr size_t kTmapCacheMax = 4096;   // beyond this the cache is simply cleared (ragged shapes keep adding keys)
bool tmap_cache_get(const TmapKey& k, CUtensorMap* out) {
  auto it = t_tmap_cache.find(k);
  if (it == t_tmap_cache.end()) return false;
  *out = it->second;
  return true;
}
void tmap_cache_put(const TmapKey& k, const CUtensorMap& m) {
  if (t_tmap_cache.size() >= kTmapCacheMax) t_tmap_cache.clear();
  t_tmap_cache.emplace(k, m);
}
}  // namespace

static int make_tmap_2d_impl(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1,
                             uint64_t row_stride_elems, uint32_t box0, uint32_t box1, CUtensorMapSwizzle sw,
                             int elem_bytes) {
  const TmapKey key = {reinterpret_cast<uint64_t>(base), dim0, dim1, 0, row_stride_elems, 0, box0, box1,
                       (uint32_t)sw, (uint32_t)elem_bytes};
  if (tmap_cache_get(key, out)) return 0;
  std::call_once(g_encode_once, [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) ==
            cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  });
  if (g_encode == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return W2VSEG_ERR_CUDA;
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0 || (row_stride_elems * elem_bytes) % 16 != 0) {
    set_error("tensor map: base %p / row stride %llu not 16-byte aligned", base,
              (unsigned long long)row_stride_elems);
    return W2VSEG_ERR_ARG;
  }
  cuuint64_t gdim[2] = {dim0, dim1};
  cuuint64_t gstride[1] = {row_stride_elems * (uint64_t)elem_bytes};  // bytes, dim1 stride
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(out, elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16,
                        2, const_cast<void*>(base), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu, stride %llu, box "
              "%u x %u)",
              (int)r, (unsigned long long)dim0, (unsigned long long)dim1,
              (unsigned long long)row_stride_elems, box0, box1);
    return W2VSEG_ERR_CUDA;
  }
  tmap_cache_put(key, *out);
  return 0;
}

int make_tmap_3d_bf16(CUtensorMap* out, const void* base, uint64_t dim0, uint64_t dim1, uint64_t dim2,
                      uint64_t stride1_elems, uint64_t stride2_elems, uint32_t box0, uint32_t box1) {
  const TmapKey key = {reinterpret_cast<uint64_t>(base), dim0, dim1, dim2, stride1_elems, stride2_elems, box0, box1,
                       (uint32_t)CU_TENSOR_MAP_SWIZZLE_128B, 2u};
  if (tmap_cache_get(key, out)) return 0;
  CUtensorMap probe;
  W2V_TRY(make_tmap_2d_bf16(&probe, base, dim0, dim1, stride1_elems, box0, box1));  // loads the encoder, checks alignment
  if ((stride2_elems * 2) % 16 != 0) {
    set_error("tensor map: plane stride %llu not 16-byte aligned", (unsigned long long)stride2_elems);
    return W2VSEG_ERR_ARG;
  }
  cuuint64_t gdim[3] = {dim0, dim1, dim2};
  cuuint64_t gstride[2] = {stride1_elems * 2, stride2_elems * 2};
  cuuint32_t box[3] = {box0, box1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim,
                        gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (3-D) failed with CUresult %d", (int)r);
    return W2VSEG_ERR_CUDA;
  }
  tmap_cache_put(key, *out);
  return 0;
}

// ---- profiling -----------------------------------------------------------------------------
namespace {
struct ProfRec { std::string name; cudaEvent_t a, b; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
thread_local const char* g_prof_tag = nullptr;
}  // namespace

bool prof_enabled() { return g_prof_on; }
void prof_tag(const char* tag) { g_prof_tag = tag; }

ProfScope::ProfScope(cudaStream_t s, const char* default_name) : stream(s), slot(-1) {
  const char* name = g_prof_tag ? g_prof_tag : default_name;
  g_prof_tag = nullptr;
  if (!g_prof_on) return;
  ProfRec r;
  r.name = name;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, stream);
}

}  // namespace w2v

extern "C" {

int32_t w2vseg_profile_enable(int32_t on) {
  w2v::g_prof_on = on != 0;
  return 0;
}

// Synchronises, then writes one line per kernel name: "<name> <launches> <total_ms>\n".
// Returns the number of bytes written (excluding the terminating NUL) or a negative error.
int64_t w2vseg_profile_collect(char* buf, size_t cap) {
  if (cudaDeviceSynchronize() != cudaSuccess) {
    w2v::set_error("profile_collect: device synchronize failed");
    return W2VSEG_ERR_CUDA;
  }
  std::map<std::string, std::pair<long long, double>> acc;
  std::vector<std::string> order;
  for (auto& r : w2v::g_prof) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, r.a, r.b);
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
    auto it = acc.find(r.name);
    if (it == acc.end()) { order.push_back(r.name); acc[r.name] = {1, (double)ms}; }
    else { it->second.first += 1; it->second.second += ms; }
  }
  w2v::g_prof.clear();
  size_t off = 0;
  for (auto& n : order) {
    char line[256];
    int k = snprintf(line, sizeof(line), "%s %lld %.6f\n", n.c_str(), acc[n].first, acc[n].second);
    if (off + (size_t)k + 1 > cap) break;
    memcpy(buf + off, line, (size_t)k);
    off += (size_t)k;
  }
  if (cap > 0) buf[off < cap ? off : cap - 1] = 0;
  return (int64_t)off;
}

int32_t w2vseg_abi_version(void) { return W2VSEG_ABI_VERSION; }
const char* w2vseg_last_error(void) { return w2v::g_err; }
int64_t w2vseg_launch_count(void) { return w2v::g_launches.load(std::memory_order_relaxed); }

int32_t w2vseg_device_ok(void) {
  static int cached_dev = -1;  // cudaGetDeviceProperties costs milliseconds: ask once per device
  int cur = -2;
  if (cudaGetDevice(&cur) == cudaSuccess && cur == cached_dev) return 0;
  int dev = 0;
  cudaDeviceProp prop;
  if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
    w2v::set_error("no CUDA device available");
    return W2VSEG_ERR_CUDA;
  }
  if (prop.major != 10) {
    w2v::set_error("device '%s' is sm_%d%d; libw2vseg contains sm_100a code only", prop.name,
                   prop.major, prop.minor);
    return W2VSEG_ERR_CUDA;
  }
  cached_dev = dev;
  return 0;
}

// conv stack of wav2vec 2.0: kernels (10,3,3,3,3,2,2), strides (5,2,2,2,2,2,2), no padding
int32_t w2vseg_num_frames(int64_t n) {
  static const int k[7] = {10, 3, 3, 3, 3, 2, 2};
  static const int s[7] = {5, 2, 2, 2, 2, 2, 2};
  for (int l = 0; l < 7; ++l) {
    if (n < k[l]) return 0;
    n = (n - k[l]) / s[l] + 1;
  }
  return (int32_t)n;
}

int32_t w2vseg_frame_stride(int64_t l_max) {
  int64_t r = (l_max + 319) / 320;
  if (r < 2) r = 2;
  return (int32_t)r;
}

}  // extern "C"
