// Error string, version and launch accounting for libgcg.so.
#include <atomic>
#include <string.h>
#include <string>
#include <vector>

#include "gcg_common.cuh"

namespace gcg {
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
}  // namespace gcg

extern "C" int gcg_version(void) { return 100; }
extern "C" const char* gcg_last_error(void) { return gcg::g_err; }
extern "C" int64_t gcg_launch_count(void) { return gcg::g_launches.load(); }
extern "C" void gcg_launch_count_reset(void) { gcg::g_launches.store(0); }

// ------------------------------------------------------------------------------------------ epoch programs
// "host-side C++ gcg_epoch() that enqueues a1..a12 on one stream (CUDA-graph capturable)".
// The reference's epoch is ONE call of a compiled Theano function (mlpconv.py:265, :295); here the equivalent
// compiled object is a gcg_epoch: the ordered list of libgcg calls of one f_train, recorded once from the layer
// code (the layers decide shapes, kernels, buffers -- exactly what Theano's compilation step decides) and then
// replayed by gcg_epoch_run() without returning to Python between launches.
struct gcg_epoch {
  std::vector<std::function<int(void*)>> calls;
  std::vector<std::string> names;
  bool recording = false;
};

namespace gcg {
static thread_local gcg_epoch* tls_epoch = nullptr;
bool epoch_recording() { return tls_epoch != nullptr; }
void epoch_record(const char* name, std::function<int(void*)> call) {
  gcg_epoch* e = tls_epoch;
  if (!e) return;
  e->calls.push_back(std::move(call));
  e->names.emplace_back(name);
}
}  // namespace gcg

extern "C" int gcg_epoch_create(gcg_epoch** out) {
  GCG_CHECK_ARG(out != nullptr, "gcg_epoch_create: out is NULL");
  *out = new gcg_epoch();
  return GCG_OK;
}

extern "C" int gcg_epoch_destroy(gcg_epoch* e) {
  if (!e) return GCG_OK;
  if (gcg::tls_epoch == e) gcg::tls_epoch = nullptr;
  delete e;
  return GCG_OK;
}

extern "C" int gcg_epoch_record_begin(gcg_epoch* e) {
  GCG_CHECK_ARG(e != nullptr, "gcg_epoch_record_begin: epoch is NULL");
  GCG_CHECK_ARG(gcg::tls_epoch == nullptr, "gcg_epoch_record_begin: this thread is already recording");
  e->calls.clear();
  e->names.clear();
  e->recording = true;
  gcg::tls_epoch = e;
  return GCG_OK;
}

extern "C" int gcg_epoch_record_end(gcg_epoch* e) {
  GCG_CHECK_ARG(e != nullptr && gcg::tls_epoch == e, "gcg_epoch_record_end: this epoch is not being recorded");
  e->recording = false;
  gcg::tls_epoch = nullptr;
  return GCG_OK;
}

extern "C" int64_t gcg_epoch_size(const gcg_epoch* e) { return e ? (int64_t)e->calls.size() : 0; }

extern "C" const char* gcg_epoch_call_name(const gcg_epoch* e, int64_t i) {
  if (!e || i < 0 || i >= (int64_t)e->names.size()) return "";
  return e->names[(size_t)i].c_str();
}

extern "C" int gcg_epoch_run(const gcg_epoch* e, void* stream) {
  GCG_CHECK_ARG(e != nullptr, "gcg_epoch_run: epoch is NULL");
  GCG_CHECK_ARG(!e->recording, "gcg_epoch_run: the epoch is still being recorded");
  gcg_epoch* outer = gcg::tls_epoch;       // a replay inside another recording is not re-recorded call by call
  gcg::tls_epoch = nullptr;
  int rc = GCG_OK;
  for (size_t i = 0; i < e->calls.size(); ++i) {
    rc = e->calls[i](stream);
    if (rc != GCG_OK) break;               // the failing entry point has set gcg_last_error
  }
  gcg::tls_epoch = outer;
  return rc;
}
