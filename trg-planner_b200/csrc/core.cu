// Error state, device selection, launch geometry and the CUDA-event profiler shared by all
// kernels of libtrgb_kernels.so.
#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "common.cuh"

namespace trgb {

// TRGB_BACKTRACE=1: native backtrace on SIGSEGV / SIGFPE / SIGABRT (debugging aid; Python's faulthandler
// only shows the Python frames)
static void crash_handler(int sig) {
  void* bt[64];
  const int n = backtrace(bt, 64);
  const char msg[] = "[trgb] fatal signal, native backtrace:\n";
  if (write(2, msg, sizeof(msg) - 1) < 0) {}
  backtrace_symbols_fd(bt, n, 2);
  _exit(128 + sig);
}
static const bool g_crash_handler_installed = [] {
  if (!std::getenv("TRGB_BACKTRACE")) return false;
  signal(SIGSEGV, crash_handler);
  signal(SIGFPE, crash_handler);
  signal(SIGABRT, crash_handler);
  return true;
}();

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  char buf[512];
  snprintf(buf, sizeof(buf), "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file,
           line, what);
  g_err = buf;
  cudaGetLastError();  // clear sticky-less errors
  return e == cudaErrorMemoryAllocation ? TRGB_E_NOMEM : TRGB_E_CUDA;
}

int sm_count() {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
      sms = 148;  // B200
  }
  return sms;
}

int grid_for_warps(int64_t n_warps, int ctas_per_sm) {
  const int64_t need = (n_warps + kWarpsPerCta - 1) / kWarpsPerCta;
  const int64_t cap = (int64_t)sm_count() * ctas_per_sm;
  int64_t g = need < cap ? need : cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ---- profiler ---------------------------------------------------------------------------
struct Pending {
  const char* name;
  cudaEvent_t a, b;
  double bytes;
};
struct Acc {
  int64_t launches = 0;
  double ms = 0, bytes = 0;
};
static std::mutex g_pmx;
static bool g_prof_on = false;
static std::vector<Pending> g_pending;
static std::vector<cudaEvent_t> g_pool;
static std::map<std::string, Acc> g_acc;
static std::atomic<int64_t> g_launches{0};

bool prof_enabled() { return g_prof_on; }
void count_launches(int64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

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

static void drain_locked() {
  for (auto& p : g_pending) {
    float ms = 0.f;
    if (cudaEventSynchronize(p.b) == cudaSuccess && cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) {
      Acc& a = g_acc[p.name];
      a.launches++;
      a.ms += ms;
      a.bytes += p.bytes;
    }
    g_pool.push_back(p.a);
    g_pool.push_back(p.b);
  }
  g_pending.clear();
}

ProfScope::ProfScope(const char* name, cudaStream_t s, double bytes) : name_(name), s_(s), bytes_(bytes) {
  g_launches.fetch_add(1, std::memory_order_relaxed);
  if (!g_prof_on) return;
  std::lock_guard<std::mutex> lk(g_pmx);
  start_ = get_event();
  stop_ = get_event();
  cudaEventRecord(start_, s_);
}
ProfScope::~ProfScope() {
  if (!start_) return;
  std::lock_guard<std::mutex> lk(g_pmx);
  cudaEventRecord(stop_, s_);
  // No draining here: synchronising on events in the middle of a run would stall whichever host
  // thread happens to launch next. Pending pairs are resolved by trgb_prof_collect / _reset; the
  // event pool therefore grows to two events per launch of the profiled region and is recycled.
  g_pending.push_back({name_, start_, stop_, bytes_});
}

}  // namespace trgb

using namespace trgb;

extern "C" const char* trgb_last_error(void) { return g_err.c_str(); }

extern "C" int trgb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

extern "C" int trgb_set_device(int device) {
  TRGB_CUDA(cudaSetDevice(device));
  return TRGB_OK;
}

extern "C" int trgb_prof_enable(int on) {
  std::lock_guard<std::mutex> lk(g_pmx);
  g_prof_on = on != 0;
  // Pre-create the events of a few profiled 10 M-point builds (2 per launch, ~14 k launches each):
  // creating them on demand put bursts of cudaEventCreate - and the driver's bookkeeping for tens
  // of thousands of new timing events - inside the region being timed.
  if (g_prof_on) {
    constexpr size_t kPrefill = 1u << 17;
    while (g_pool.size() + 2 * g_pending.size() < kPrefill) {
      cudaEvent_t e = nullptr;
      if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); break; }
      g_pool.push_back(e);
    }
  }
  return TRGB_OK;
}

extern "C" int trgb_prof_reset(void) {
  std::lock_guard<std::mutex> lk(g_pmx);
  drain_locked();
  g_acc.clear();
  g_launches.store(0);
  return TRGB_OK;
}

extern "C" int trgb_prof_collect(TrgbProfEntry* out, int cap) {
  std::lock_guard<std::mutex> lk(g_pmx);
  drain_locked();
  int i = 0;
  for (auto& kv : g_acc) {
    if (out && i < cap) {
      memset(&out[i], 0, sizeof(TrgbProfEntry));
      strncpy(out[i].name, kv.first.c_str(), sizeof(out[i].name) - 1);
      out[i].launches = kv.second.launches;
      out[i].total_ms = kv.second.ms;
      out[i].units = kv.second.bytes;
    }
    ++i;
  }
  return i;
}

extern "C" int64_t trgb_launch_count(void) { return g_launches.load(); }
