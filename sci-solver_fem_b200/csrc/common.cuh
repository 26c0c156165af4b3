// common.cuh — device buffers, error handling and CUB wrappers shared by every stage.
// Product code (sm_100a only).  Nothing here touches oracle/.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace fsb {

struct CudaError : public std::runtime_error {
  explicit CudaError(const std::string& s) : std::runtime_error(s) {}
};

#define FSB_CUDA(call)                                                                           \
  do {                                                                                           \
    cudaError_t e__ = (call);                                                                    \
    if (e__ != cudaSuccess)                                                                      \
      throw ::fsb::CudaError(std::string(#call) + " failed: " + cudaGetErrorString(e__) + " at " + \
                             __FILE__ + ":" + std::to_string(__LINE__));                         \
  } while (0)

#define FSB_CHECK_LAUNCH() FSB_CUDA(cudaGetLastError())

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Function attributes (opt-in dynamic shared memory) are per DEVICE: a process may hold solvers on several
// devices, so "already set" is tracked per device ordinal.
struct PerDeviceOnce {
  bool done[64] = {};
  bool first(int device) {
    if (device < 0 || device >= 64) return true;
    if (done[device]) return false;
    done[device] = true;
    return true;
  }
};

// Device-visible description of the GPU group of a sharded solve (one process per GPU; peers'
// arenas are mapped through CUDA IPC and reached over NVLink).  nranks == 1: everything is local.
//
// Exchanges inside the iteration use a FLAG-IN-DATA protocol (the idea of NCCL's LL protocol): a value travels
// as one 16-byte store {lo32, epoch, hi32, epoch} into the receiver's buffer; 8-byte stores are atomic, so a
// receiver that polls its slot until both epoch words match has the value — no fence, no separate flag, no
// barrier, one NVLink one-way latency.  The epoch is a device counter that every exchange kernel of the solve
// bumps once (all GPUs run the same sequence of exchanges, skipped consistently once `done` is set).
// The all-reduce of the dot products uses the same slots ([2][kMaxRanks], parity double-buffered).
// Only the all-gather of the solution at the end of a solve still uses the fenced store + flag handshake.
constexpr int kMaxRanks = 8;
constexpr int kMaxChan = 47;   // exchange channels: 3 of the PCG driver + 5 per sharded level
struct DistDev {
  int rank = 0, nranks = 1;
  unsigned long long* my_flags = nullptr;             // [kMaxRanks] epoch flags of the fenced handshake, written by the peers
  unsigned long long* peer_flags[kMaxRanks] = {};     // the same array in every peer's arena
  uint4* my_red = nullptr;                            // [2][kMaxRanks] all-reduce slots {lo, epoch, hi, epoch}
  uint4* peer_red[kMaxRanks] = {};
  unsigned long long* epoch = nullptr;                // [2] local counters: [0] all-reduce, [1] fenced handshake
  unsigned* xchg = nullptr;                           // epoch of the flag-in-data exchanges (bumped by every exchange kernel)
  int* error = nullptr;                               // set when a peer wait times out; every later wait returns at once
};
// one flag-in-data exchange as seen by this GPU
struct LLXchg {
  const int* sidx = nullptr; const int* sptr = nullptr; int stotal = 0;  // what I send: indices into the source vector, segment per destination
  const int* ridx = nullptr; int rtotal = 0;                              // where slot k of my receive buffer goes in the destination vector
  uint4* mybuf = nullptr;                                                 // my receive buffer of the channel (rtotal slots)
  uint4* peerbuf[kMaxRanks] = {};                                         // start of MY segment inside peer q's receive buffer
};

// Optional per-launch timing (CUDA events around every solve-phase kernel; used by bench.py's
// roofline pass, never inside a timed region).
struct Profiler {
  struct Rec { const char* name; int level; cudaEvent_t a, b; };
  bool on = false;
  int cur_level = 0;
  std::vector<Rec> recs;
  std::vector<cudaEvent_t> pool;
  cudaEvent_t get() {
    if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
  }
  void clear() { for (auto& r : recs) { pool.push_back(r.a); pool.push_back(r.b); } recs.clear(); }
  ~Profiler() { clear(); for (auto e : pool) cudaEventDestroy(e); }
};

// The stream every kernel of a solver instance is launched on.
struct Ctx {
  cudaStream_t stream = nullptr;
  int device = 0;
  int num_sms = 148;
  Profiler* prof = nullptr;
  DistDev dist;                        // nranks == 1 unless a sharded solve is connected
  unsigned int* dist_ticket = nullptr; // last-CTA ticket of the push kernels
  cudaStream_t side[2] = {nullptr, nullptr};  // helper streams: kernels that may run next to each other (fork/join by events)
  cudaEvent_t ev_fork = nullptr, ev_join[2] = {nullptr, nullptr};
  cudaStream_t xstream = nullptr;      // sharded solve: a halo exchange runs here while the interior rows of its consumer run on `stream`
  cudaEvent_t ev_xfork = nullptr, ev_xjoin = nullptr;
};

struct RowRange { int begin = 0, end = -1; };  // end < 0: all rows
struct PeerPtrs { double* p[kMaxRanks]; };

struct ProfScope {
  Profiler* p; size_t idx; cudaStream_t s;
  ProfScope(const Ctx& c, const char* name) : p(c.prof && c.prof->on ? c.prof : nullptr), idx(0), s(c.stream) {
    if (!p) return;
    Profiler::Rec r{name, p->cur_level, p->get(), p->get()};
    cudaEventRecord(r.a, s);
    idx = p->recs.size();
    p->recs.push_back(r);
  }
  ~ProfScope() { if (p) cudaEventRecord(p->recs[idx].b, s); }
};

// Stream-ordered device buffer (cudaMallocAsync pool: setup makes hundreds of temporaries).
template <typename T>
class DevBuf {
 public:
  DevBuf() = default;
  DevBuf(size_t n, cudaStream_t s) { alloc(n, s); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept { swap(o); }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); swap(o); }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t n, cudaStream_t s) {
    release();
    n_ = n; s_ = s;
    if (n) FSB_CUDA(cudaMallocAsync((void**)&p_, n * sizeof(T), s));
  }
  void release() {
    if (p_ && !view_) cudaFreeAsync(p_, s_);
    p_ = nullptr; n_ = 0; view_ = false;
  }
  // non-owning view (multi-GPU: vectors that peers write live in one IPC-shared arena)
  void view(T* p, size_t n, cudaStream_t s) { release(); p_ = p; n_ = n; s_ = s; view_ = true; }
  void swap(DevBuf& o) { std::swap(p_, o.p_); std::swap(n_, o.n_); std::swap(s_, o.s_); std::swap(view_, o.view_); }
  T* get() const { return p_; }
  operator T*() const { return p_; }
  size_t size() const { return n_; }
  void zero() { if (n_) FSB_CUDA(cudaMemsetAsync(p_, 0, n_ * sizeof(T), s_)); }
  void fill_bytes(int v) { if (n_) FSB_CUDA(cudaMemsetAsync(p_, v, n_ * sizeof(T), s_)); }
  void from_host(const T* h, size_t n) { FSB_CUDA(cudaMemcpyAsync(p_, h, n * sizeof(T), cudaMemcpyHostToDevice, s_)); }
  void from_device(const T* d, size_t n) { FSB_CUDA(cudaMemcpyAsync(p_, d, n * sizeof(T), cudaMemcpyDeviceToDevice, s_)); }
  void to_host(T* h, size_t n) const {
    FSB_CUDA(cudaMemcpyAsync(h, p_, n * sizeof(T), cudaMemcpyDeviceToHost, s_));
    FSB_CUDA(cudaStreamSynchronize(s_));
  }
  std::vector<T> to_vector() const { std::vector<T> v(n_); if (n_) to_host(v.data(), n_); return v; }
  T read(size_t i) const { T v; FSB_CUDA(cudaMemcpyAsync(&v, p_ + i, sizeof(T), cudaMemcpyDeviceToHost, s_)); FSB_CUDA(cudaStreamSynchronize(s_)); return v; }

 private:
  T* p_ = nullptr;
  size_t n_ = 0;
  cudaStream_t s_ = nullptr;
  bool view_ = false;
};

typedef DevBuf<int> IBuf;
typedef DevBuf<double> DBuf;

struct DCsr {
  int nrows = 0, ncols = 0, nnz = 0;
  IBuf ptr, col;
  DBuf val;
};

// Sliced ELL (SELL-32): 32 consecutive rows form a slice stored column-major with the slice's own
// width, so a warp streams (col, val) with one fully coalesced 128/256-byte request per step and no
// staging; padding entries carry val = 0 and a valid column.
struct Sell {
  int nrows = 0, ncols = 0, nslices = 0;
  long long nstored = 0;
  DevBuf<long long> sptr;  // nslices + 1 entry offsets
  IBuf col;
  DBuf val;
  IBuf rowmap;             // optional: list position -> row (rows sorted by length inside windows of `window` rows)
  int window = 0;
  bool ready() const { return nslices > 0; }
};

// ---- primitives implemented in prims.cu (CUB under the hood; setup-time plumbing only) ----
void sort_pairs_u64_u32(const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, size_t n, int end_bit, cudaStream_t s);
void sort_keys_u64(const uint64_t* kin, uint64_t* kout, size_t n, int end_bit, cudaStream_t s);
void sort_pairs_i32_i32(const int* kin, int* kout, const int* vin, int* vout, size_t n, int end_bit, cudaStream_t s);  // stable, keys >= 0
void exclusive_scan_i32(const int* in, int* out, size_t n, cudaStream_t s);
void inclusive_scan_i32(const int* in, int* out, size_t n, cudaStream_t s);
int reduce_max_i32(const int* in, size_t n, cudaStream_t s);
long long reduce_sum_i32(const int* in, size_t n, cudaStream_t s);
int count_equal_i32(const int* in, size_t n, int value, cudaStream_t s);
void iota_i32(int* p, size_t n, cudaStream_t s);
void fill_i32(int* p, size_t n, int v, cudaStream_t s);
void fill_f64(double* p, size_t n, double v, cudaStream_t s);
int bits_for(long long maxval);  // number of low bits needed to represent values in [0, maxval]

}  // namespace fsb
