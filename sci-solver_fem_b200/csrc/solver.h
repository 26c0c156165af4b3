// solver.h — the host layer of the path: FEMSolver-equivalent state machine
// (mesh -> assemble -> AMG setup -> solve) over the device kernels.
#pragma once
#include <map>
#include <string>
#include <vector>

#include "fsb_internal.h"

namespace fsb {

// parameter surface of the reference's FEMSolver (src/FEMSolver.h:28-66, defaults FEMSolver.cu:11-34)
struct Params {
  int verbose = 0;
  int maxLevels = 100, maxIters = 100, preInnerIters = 5, postInnerIters = 5, postRelaxes = 1, cycleIters = 1;
  int dsType = 0, topSize = 256, randMisParameters = 90102, partitionMaxSize = 512, aggregatorType = 0;
  int convergeType = 0, cycleType = 0, solverType = 0, device = 0, blockSize = 256;
  double tolerance = 1e-6, smootherWeight = 1.0, proOmega = 0.67;
  // additive surface: deterministic aggregation seed; reference level-0 quirk switch
  unsigned seed = 0;
  int refLevel0NoPerm = 0;
  int useGraphs = 1;   // capture one PCG iteration into a CUDA graph
  int checkEvery = 2;  // PCG iterations enqueued between convergence polls
  int profile = 0;     // 1: time every solve-phase kernel with CUDA events (disables graphs for that solve)
};

// device-resident PCG state: no scalar ever crosses PCIe inside the iteration
struct PcgScalars {
  double rz_old, rz_new, py, alpha, beta, rr, bnorm, tol;
  int done, niter, maxit, hist_len;
  int hist_cap, err;
  unsigned int ticket[4];
};

// per-partition descriptor of the shared-memory variant (levels >= 1)
struct alignas(16) SellgDesc {
  int r0, np;        // first row / rows of the partition
  int w0, nw;        // first warp slab / number of warp slabs
  long long base;    // offset of the first slab (entries)
  int slots, pad;    // entries of all slabs
};
constexpr int kEllClasses = 3;
// everything a CTA of the fine-level smoother needs to find its slabs, in one 48-byte load
struct alignas(16) EllDesc {
  int r0, np;            // first row / rows of the partition
  long long base;        // offset of the partition's first warp slab (entries)
  unsigned char K[32];   // slab width of each warp
};
inline int ell_class(int rows) { return rows <= 256 ? 0 : rows <= 384 ? 1 : 2; }

struct LevelData {
  int n = 0, nnout = 0, nparts = 0, maxPartRows = 0, level_id = 0;
  DCsr A;             // permuted numbering (rows of a partition contiguous); coarsest: external numbering
  DBuf diag;
  DCsr P, R;          // P: rows internal(l), cols external(l+1); R = P^T
  DCsr RA;            // R A, levels >= 1 above the dense tail: restricted residual as R b - (R A) x in one kernel
  Aggregation agg;
  IBuf pstart;        // first row of each partition (nparts+1)
  DCsr Aout;          // inter-partition entries (rows internal numbering, global columns)
  Sell sA, sAout, sP; // SELL-32 streaming copies (large levels only): operator, A_out, prolongator
  bool use_ell = false;  // intra-partition off-diagonal entries as per-partition column-major ELL slabs
  int ellMaxK = 0, coopG = 1;
  int clusterC = 1;      // CTAs per partition of the cluster smoother
  int maxChunkNnz = 0;   // largest number of CSR entries of one CTA's row chunk
  int maxChunkRows = 0;
  int smemBytes = 0;     // > 0: the cluster smoother is usable (chunk fits in shared memory)
  // fine-level smoother: partitions by CTA size class (<= 256 / <= 384 / more rows) keep occupancy up
  DevBuf<EllDesc> plist[kEllClasses];
  int nlist[kEllClasses] = {0, 0, 0};
  DevBuf<EllDesc> plistOwn[kEllClasses];  // the same lists restricted to this GPU's partitions (sharded solve)
  int nlistOwn[kEllClasses] = {0, 0, 0};
  int ownP0 = 0, ownPn = 0;               // this GPU's partition range of a sharded level (dist.cu)
  std::vector<EllDesc> ellDescHost;       // one descriptor per partition (host copy, used to build the lists)
  // sorted, warp-sliced ELL slabs (hierarchy.cu: split_partitions)
  bool use_sellg = false;          // sorted slabs with ellG lanes per row, shared-memory resident (smooth_sellg_kernel)
  int ellG = 1;
  DevBuf<SellgDesc> sellgDesc;
  int sellgMaxSlots = 0;
  DevBuf<unsigned short> ellrow;  // n: local row handled by thread t of the partition's CTA (rows by decreasing length)
  IBuf pwarp;                     // nparts+1: first warp slab of each partition
  DevBuf<long long> ellwptr;      // warps+1: slab offsets (entries); width of a slab = (ellwptr[w+1]-ellwptr[w]) / 32
  int ellWarps = 0;
  long long ellSlots = 0;         // stored entries (incl. padding)
  DBuf ellval;
  DevBuf<unsigned short> ellcol;  // position of the column in the partition's sorted order; bit 15: copy B of the x tile
  // dense partition blocks (dense_tail.cu: build_block_smoother): on a level with few, small partitions the nu sweeps of a
  // smoothing stage are latency (one barrier-separated pass per sweep), but the stage is a fixed linear map per partition:
  //   pre  x = S(nu1) b          post  x = G^nu2 x_in + S(nu2-1) b'
  // so the three m x m blocks per partition are formed once at setup and a stage is ONE pass of small dense GEMVs
  bool use_blockdense = false;
  DBuf bdS1, bdGp, bdS2;          // packed row-major blocks, partition p at bdOff[p]
  DevBuf<long long> bdOff;        // nparts + 1
  IBuf bdWork;                    // per CTA: partition, first local row (2 ints)
  int bdCtas = 0;
  IBuf xadj, adj;     // graph handed to the aggregator (external numbering)
  DBuf b, x, x2, r;   // work vectors, internal numbering
  DBuf bc, xc;        // restricted residual / coarse correction, external numbering of level l+1
};

class Solver {
 public:
  explicit Solver(int device);
  ~Solver();
  Params prm;
  std::string last_error;

  // stage 1
  void set_mesh(int nv, const double* xyz, int ne, int npe, const int* elems, const int* labels, bool on_device);
  void assemble();                      // pattern + element loop  (FEMSolver::getMatrixFromMesh)
  int rows() const { return A0.nrows; }
  int nnz() const { return A0.nnz; }
  void get_matrix(int* ptr, int* col, double* val);
  void set_matrix_values(const double* val, bool on_device);
  void set_matrix_csr(int n, int nnz, const int* ptr, const int* col, const double* val);  // e.g. readMatlabSparseMatrix result
  // stage 2
  void setup();                         // AMG::setup
  int num_levels() const { return (int)levels.size(); }
  const LevelData& level(int l) const { return levels.at(l); }
  // stage 3
  void solve(const double* b, double* x, bool on_device);  // AMG::solve; x = initial guess in, solution out
  int iterations = 0;
  double final_relres = -1;
  std::vector<double> resid_history;
  std::map<std::string, double> times_ms;
  // kernel-level entry points for tests / microbenchmarks (device pointers, level-0 internal numbering)
  void spmv_fine(const double* x, double* y);
  void apply_matrix(const double* x, double* y);   // y = A x with the user-order matrix (device pointers)
  void precondition(const double* r, double* z);   // one V-cycle, z = M^-1 r
  long long launches = 0;              // kernels launched by the last solve()
  // ---- stage 4: sharded solve over the GPUs of one box (one process per GPU, replicated setup) ----
  // dist_prepare(): after setup(); deals the partitions of every level that is large enough out to the ranks
  // (contiguous, nnz-balanced ranges), builds the push lists of every exchange and allocates the IPC arena.
  // dist_connect(): maps the peers' arenas.  Then solve() (PCG) runs sharded.
  void dist_prepare(int rank, int nranks);
  void dist_get_blob(void* blob /* kDistBlobBytes */, long long* arena_bytes);
  void dist_connect(const void* blobs /* nranks x kDistBlobBytes, rank order */);
  void dist_disconnect();
  double dist_bench_exchange(int chan, int reps);  // tools: microseconds per back-to-back exchange (chan < 0: all-reduce)
  // one exchange pattern: the values of mine that peers need, grouped by destination peer, and the places where
  // the values the peers send me go (same index in every GPU's copy of the exchanged vector)
  struct PushList {
    IBuf idx, ptr;               // send entries / nranks+1 segment offsets
    int total = 0;
    IBuf ridx;                   // receive entries, grouped by source peer (ascending index inside a segment — the order
    int rptr[kMaxRanks + 1] = {};  //   in which that peer's list sends them)
    int rtotal = 0;
  };
  // channel ids: 0 p halo, 1 x0 halo (initial residual), then 5 per sharded level
  enum { kChanP = 0, kChanX0 = 1, kChanLevel0 = 2, kChanPerLevel = 5 };
  enum { kXPre = 0, kRes = 1, kDown = 2, kUp = 3, kXPost = 4 };
  struct DistLevel {
    int pbeg[kMaxRanks + 1] = {};  // partition ranges
    int rbeg[kMaxRanks + 1] = {};  // row ranges (internal numbering of the level)
    int abeg[kMaxRanks + 1] = {};  // row ranges of the next level in ITS external numbering = aggregates of the owned partitions
    PushList sendA;                // operator columns across the cut (x after pre-smoothing / after the coarse correction, p, x0)
    PushList sendR;                // residual rows the peers' restriction rows reference
    PushList sendDown;             // restricted residual: entries a peer owns on the next (sharded) level, or — next level
                                   //   replicated — my whole slice to everybody (all-gather)
    PushList sendUp;               // coarse corrections of my next-level rows that a peer's prolongator rows reference
    size_t off_x = 0, off_r = 0, off_bc = 0, off_xc = 0;
    // interior ranges (1024-aligned, inside this GPU's ranges): rows whose operator / prolongator rows and coarse rows whose
    // restriction rows reference nothing that arrives through an exchange — the consumer kernel runs on them WHILE the
    // exchange is in flight (begin >= end: no overlap on this level)
    RowRange intA, intR, intP;
  };
  struct Channel {                 // receive buffer of one exchange site + where my segments start in the peers' buffers
    const PushList* list = nullptr;
    size_t buf_off = 0;            // my receive buffer (arena offset, rtotal slots of 16 B)
    LLXchg dev;                    // filled by dist_connect
  };
  struct DistHost {
    int rank = 0, nranks = 1;
    bool connected = false;
    int nshard = 0;                // levels 0 .. nshard-1 are sharded, the rest is computed redundantly on every GPU
    std::vector<DistLevel> lev;
    Channel chan[kMaxChan];
    int nchan = 0;
    int user_lo = 0, user_hi = 0;  // user-numbering range that covers this GPU's fine rows (host-buffer solves copy only this slice)
    char* arena = nullptr;
    size_t arena_bytes = 0;
    size_t off_flags = 0, off_red = 0, off_p = 0, off_cgx = 0;
    char* peer[kMaxRanks] = {};
    DevBuf<unsigned long long> epoch;
    DevBuf<unsigned> xchg;
    IBuf error;
    DevBuf<unsigned int> ticket;
  } dist;
  // what every rank publishes before dist_connect (all-gathered by the caller): the IPC handle of its arena and, per
  // channel, the offset of its receive buffer and the segment offsets of the source peers inside it
  struct DistBlob {
    unsigned char handle[64];
    int nchan, nranks;
    unsigned long long buf_off[kMaxChan];
    int rptr[kMaxChan][kMaxRanks + 1];
  };
  static constexpr int kDistBlobBytes = 2560;
  static_assert(sizeof(DistBlob) <= kDistBlobBytes, "blob size");
  bool sharded(int lev) const { return dist.connected && dist.nranks > 1 && cg_active_ && lev < dist.nshard; }
  void exchange_chan(int chan, const double* src, double* dst, const int* done, bool side_stream = false);
  template <typename F>
  void with_exchange(bool on, int lev, int which, double* v, RowRange full, RowRange interior, const int* done, F&& consumer);
  void exchange(int lev, int which, const double* src, double* dst, const int* done) { exchange_chan(kChanLevel0 + kChanPerLevel * lev + which, src, dst, done); }
  PeerPtrs peers_at(size_t off) const { PeerPtrs p = {}; for (int q = 0; q < dist.nranks; q++) p.p[q] = reinterpret_cast<double*>(dist.peer[q] + off); return p; }
  std::string profile_report();        // "name level launches total_ms" lines of the last profiled solve
  Profiler profiler;

  Ctx ctx;
  Mesh mesh;
  Pattern pat;
  DCsr A0;                              // user-order fine matrix
  bool custom_matrix = false;
  std::vector<LevelData> levels;
  DBuf Ainv;                            // dense inverse of the coarsest operator
  // dense tail (dense_tail.cu): one V-cycle on level tail_level_ (and everything below it) as a dense matrix
  static constexpr int kDenseTailMaxRows = 3072;
  DBuf Mtail;
  int tail_level_ = -1;
  double tail_key_[4] = {0, 0, 0, 0};    // smoother parameters Mtail was built for
  void* cublas_ = nullptr;
  void build_dense_tail();
  void build_block_smoothers();          // dense partition blocks of the small levels above the tail
  void ensure_dense_tail();
  void destroy_cublas();
  bool has_setup = false;

 private:
  void vcycle(int lev, const double* b_ext, const int* gather, double* x_ext, const int* scatter, double* x_int_out);
  void pcg(const double* b_user, double* x_user);
  void enqueue_pcg_iteration();
  void tic(const char* name);
  void toc(const char* name);
  cudaEvent_t ev0_ = nullptr, ev1_ = nullptr;
  PcgScalars* poll_host_ = nullptr;            // pinned double buffer of the convergence poll
  cudaEvent_t poll_ev_[2] = {nullptr, nullptr};
  // PCG state
  DBuf cg_b, cg_x, cg_r, cg_z, cg_p, cg_y, partials, hist;
  DevBuf<PcgScalars> scal;
  cudaGraphExec_t iter_graph_ = nullptr;
  struct GraphKey { void* p[9]; int pre, post, relaxes; double w; };
  GraphKey graph_key_ = {};
  bool cg_active_ = false;
  long long iter_launches_ = 0;
  void destroy_graph();
};

}  // namespace fsb
