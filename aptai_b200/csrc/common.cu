#include "common.h"

#include <atomic>
#include <mutex>
#include <string.h>

namespace aptai {

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};
static thread_local int g_reverse = 0;

int traversal_reverse() { return g_reverse; }

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int after_launch(const char* what) {
  count_launch(1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return static_cast<int>(e);
  }
  return APTAI_OK;
}

struct DevInfo {
  int major = -1, minor = -1, sms = 0;
};
static DevInfo g_dev[64];
static std::mutex g_mu;

static const DevInfo* dev_info() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  if (g_dev[dev].major < 0) {
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return nullptr;
    g_dev[dev].major = p.major;
    g_dev[dev].minor = p.minor;
    g_dev[dev].sms = p.multiProcessorCount;
  }
  return &g_dev[dev];
}

int check_arch() {
  const DevInfo* d = dev_info();
  if (!d) {
    cudaGetLastError();
    set_error("no CUDA device available (aptai_b200 has no CPU fallback)");
    return APTAI_ERR_ARCH;
  }
  if (d->major != 10) {
    set_error("device is sm_%d%d; aptai_b200 kernels are built for sm_100a only", d->major, d->minor);
    return APTAI_ERR_ARCH;
  }
  return APTAI_OK;
}

int num_sms() {
  const DevInfo* d = dev_info();
  return d ? d->sms : 148;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int encode_tmap_typed(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank,
                             const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle128);

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                     const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  return encode_tmap_typed(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, base, rank, dims, strides_bytes, box, swizzle128);
}
int encode_tmap_f32(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                    const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  return encode_tmap_typed(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, base, rank, dims, strides_bytes, box, swizzle128);
}

static int encode_tmap_typed(CUtensorMap* map, CUtensorMapDataType dtype, const void* base, int rank,
                             const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box, int swizzle128) {
  if (!g_encode) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult qres;
      cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
      if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
        cudaGetLastError();
        set_error("cuTensorMapEncodeTiled not available from the driver");
        return APTAI_ERR_DRIVER;
      }
      g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides_bytes[i];
  CUresult r = g_encode(map, dtype, static_cast<cuuint32_t>(rank),
                        const_cast<void*>(base), gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]",
              static_cast<int>(r), rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return APTAI_ERR_DRIVER;
  }
  return APTAI_OK;
}

}  // namespace aptai

extern "C" {
int aptai_version(void) { return 100; }
const char* aptai_last_error_string(void) { return aptai::g_err; }
int64_t aptai_launch_count(void) { return aptai::g_launches.load(std::memory_order_relaxed); }
void aptai_set_traversal(int reverse) { aptai::g_reverse = reverse ? 1 : 0; }
}
