// Host-side plumbing of the C ABI: error string, launch counter, device info, cfg validation.
#include <stdlib.h>
#include "common.cuh"

#include <string.h>

namespace tpspp {

static thread_local char g_err[512] = "";
static thread_local int g_launches = 0;
static thread_local bool g_pdl = false;

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
// ---- optional per-launch profiling (tpspp_launch_profile): one CUDA event after every kernel launch ----
constexpr int PROF_MAX = 64;
static thread_local bool g_prof_on = false;
static thread_local cudaEvent_t g_prof_ev[PROF_MAX + 1];
static thread_local int g_prof_have = 0;      // events created so far
static thread_local int g_prof_n = 0;         // launches recorded in the current call
static thread_local cudaStream_t g_prof_stream = nullptr;

static void prof_record() {
  if (g_prof_n > PROF_MAX) return;
  while (g_prof_have <= g_prof_n) {
    if (cudaEventCreate(&g_prof_ev[g_prof_have]) != cudaSuccess) { g_prof_on = false; return; }
    ++g_prof_have;
  }
  cudaEventRecord(g_prof_ev[g_prof_n], g_prof_stream);
  ++g_prof_n;
}
void prof_begin(cudaStream_t st) {            // called at the top of an entry point that wants its launches timed
  if (!g_prof_on) return;
  g_prof_stream = st;
  g_prof_n = 0;
  prof_record();
}
void count_launch(int n) {
  g_launches += n;
  if (g_prof_on && g_prof_n > 0) prof_record();
}
void reset_launch_count() { g_launches = 0; }
static bool pdl_env() {
  static const bool on = [] { const char* e = getenv("TPSPP_PDL"); return !(e != nullptr && e[0] == '0'); }();
  return on;
}
bool pdl_on() { return g_pdl && !g_prof_on; }      // per-launch timing (events between launches) wants the grids serialised
void pdl_scope(bool on) { g_pdl = on && pdl_env(); }

int sm_count() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    cached = v;
    cached_dev = dev;
  }
  return cached;
}

int validate_cfg(const tpspp_warp_cfg* cfg) {
  TPSPP_REQUIRE(cfg != nullptr, "cfg is NULL");
  TPSPP_REQUIRE(cfg->batch >= 0, "batch must be >= 0 (got %d)", cfg->batch);
  TPSPP_REQUIRE(cfg->channels0 > 0 && cfg->src0_h > 0 && cfg->src0_w > 0,
                "src0 geometry must be positive (C=%d H=%d W=%d)", cfg->channels0, cfg->src0_h, cfg->src0_w);
  TPSPP_REQUIRE(cfg->channels1 >= 0, "channels1 must be >= 0");
  if (cfg->channels1 > 0)
    TPSPP_REQUIRE(cfg->src1_h > 0 && cfg->src1_w > 0, "src1 geometry must be positive");
  TPSPP_REQUIRE(cfg->out_h > 0 && cfg->out_w > 0, "rectified size must be positive");
  TPSPP_REQUIRE(cfg->num_fiducial > 0 && cfg->num_fiducial <= 125,
                "num_fiducial must be in [1,125] (got %d)", cfg->num_fiducial);
  TPSPP_REQUIRE(cfg->mode == TPSPP_MODE_ATTENTION || cfg->mode == TPSPP_MODE_CLASSICAL,
                "unknown mode %d", cfg->mode);
  TPSPP_REQUIRE(cfg->feat_dtype == TPSPP_F32 || cfg->feat_dtype == TPSPP_BF16 || cfg->feat_dtype == TPSPP_SRC0_BF16,
                "unknown feat_dtype %d", cfg->feat_dtype);
  TPSPP_REQUIRE((long long)cfg->src0_h * cfg->src0_w < (1LL << 30) &&
                    (long long)cfg->out_h * cfg->out_w < (1LL << 30),
                "plane too large");
  return TPSPP_OK;
}

}  // namespace tpspp

extern "C" int tpspp_version(void) { return TPSPP_ABI_VERSION; }
extern "C" const char* tpspp_last_error(void) { return tpspp::g_err; }
extern "C" int tpspp_last_launch_count(void) { return tpspp::g_launches; }

extern "C" int tpspp_launch_profile(int enable) {
  tpspp::g_prof_on = enable != 0;
  tpspp::g_prof_n = 0;
  return TPSPP_OK;
}
extern "C" int tpspp_launch_profile_read(float* ms, int capacity, int* count) {
  using namespace tpspp;
  TPSPP_REQUIRE(ms != nullptr && count != nullptr && capacity >= 0, "tpspp_launch_profile_read: bad arguments");
  *count = 0;
  if (g_prof_n < 2) return TPSPP_OK;
  TPSPP_CHECK_CUDA(cudaEventSynchronize(g_prof_ev[g_prof_n - 1]));
  const int n = g_prof_n - 1 < capacity ? g_prof_n - 1 : capacity;
  for (int i = 0; i < n; ++i) TPSPP_CHECK_CUDA(cudaEventElapsedTime(&ms[i], g_prof_ev[i], g_prof_ev[i + 1]));
  *count = n;
  return TPSPP_OK;
}

extern "C" int tpspp_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  TPSPP_CHECK_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  TPSPP_CHECK_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  TPSPP_CHECK_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  TPSPP_CHECK_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  if (maj != 10) {
    tpspp::set_error("libtpspp holds sm_100a code only; device is sm_%d%d", maj, min);
    return TPSPP_E_NO_DEVICE;
  }
  return TPSPP_OK;
}
