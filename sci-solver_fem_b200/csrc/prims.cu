// prims.cu — sort / scan / reduce plumbing for the one-off setup stages.
// These wrap CUB device primitives (part of the CUDA toolkit).  They are used only
// where the reference itself calls Thrust (sort_by_key, scan, reduce, count); every
// numerically substantive kernel of the path is hand-written elsewhere.
#include <cub/cub.cuh>

#include "common.cuh"

namespace fsb {

int bits_for(long long maxval) {
  int b = 1;
  while (b < 63 && (1LL << b) <= maxval) b++;
  return b;
}

namespace {
struct Temp {
  void* p = nullptr; size_t bytes = 0; cudaStream_t s;
  explicit Temp(cudaStream_t st) : s(st) {}
  void alloc() { if (bytes) FSB_CUDA(cudaMallocAsync(&p, bytes, s)); }
  ~Temp() { if (p) cudaFreeAsync(p, s); }
};
}  // namespace

void sort_pairs_u64_u32(const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, size_t n, int end_bit, cudaStream_t s) {
  if (n == 0) return;
  Temp t(s);
  FSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t.bytes, kin, kout, vin, vout, (int64_t)n, 0, end_bit, s));
  t.alloc();
  FSB_CUDA(cub::DeviceRadixSort::SortPairs(t.p, t.bytes, kin, kout, vin, vout, (int64_t)n, 0, end_bit, s));
}

void sort_keys_u64(const uint64_t* kin, uint64_t* kout, size_t n, int end_bit, cudaStream_t s) {
  if (n == 0) return;
  Temp t(s);
  FSB_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, t.bytes, kin, kout, (int64_t)n, 0, end_bit, s));
  t.alloc();
  FSB_CUDA(cub::DeviceRadixSort::SortKeys(t.p, t.bytes, kin, kout, (int64_t)n, 0, end_bit, s));
}

void sort_pairs_i32_i32(const int* kin, int* kout, const int* vin, int* vout, size_t n, int end_bit, cudaStream_t s) {
  if (n == 0) return;
  Temp t(s);
  const unsigned* k0 = reinterpret_cast<const unsigned*>(kin);
  unsigned* k1 = reinterpret_cast<unsigned*>(kout);
  FSB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, t.bytes, k0, k1, vin, vout, (int64_t)n, 0, end_bit, s));
  t.alloc();
  FSB_CUDA(cub::DeviceRadixSort::SortPairs(t.p, t.bytes, k0, k1, vin, vout, (int64_t)n, 0, end_bit, s));
}

void exclusive_scan_i32(const int* in, int* out, size_t n, cudaStream_t s) {
  if (n == 0) return;
  Temp t(s);
  FSB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, t.bytes, in, out, (int64_t)n, s));
  t.alloc();
  FSB_CUDA(cub::DeviceScan::ExclusiveSum(t.p, t.bytes, in, out, (int64_t)n, s));
}

void inclusive_scan_i32(const int* in, int* out, size_t n, cudaStream_t s) {
  if (n == 0) return;
  Temp t(s);
  FSB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, t.bytes, in, out, (int64_t)n, s));
  t.alloc();
  FSB_CUDA(cub::DeviceScan::InclusiveSum(t.p, t.bytes, in, out, (int64_t)n, s));
}

int reduce_max_i32(const int* in, size_t n, cudaStream_t s) {
  if (n == 0) return INT_MIN;
  Temp t(s);
  DevBuf<int> out(1, s);
  FSB_CUDA(cub::DeviceReduce::Max(nullptr, t.bytes, in, out.get(), (int64_t)n, s));
  t.alloc();
  FSB_CUDA(cub::DeviceReduce::Max(t.p, t.bytes, in, out.get(), (int64_t)n, s));
  return out.read(0);
}

long long reduce_sum_i32(const int* in, size_t n, cudaStream_t s) {
  if (n == 0) return 0;
  Temp t(s);
  DevBuf<long long> out(1, s);
  cub::TransformInputIterator<long long, cub::CastOp<long long>, const int*> it(in, cub::CastOp<long long>());
  FSB_CUDA(cub::DeviceReduce::Sum(nullptr, t.bytes, it, out.get(), (int64_t)n, s));
  t.alloc();
  FSB_CUDA(cub::DeviceReduce::Sum(t.p, t.bytes, it, out.get(), (int64_t)n, s));
  return out.read(0);
}

namespace {
struct EqOp {
  int v;
  __host__ __device__ int operator()(const int& x) const { return x == v ? 1 : 0; }
};
__global__ void iota_kernel(int* p, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = (int)i;
}
__global__ void fill_i_kernel(int* p, size_t n, int v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
__global__ void fill_d_kernel(double* p, size_t n, double v) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
}  // namespace

int count_equal_i32(const int* in, size_t n, int value, cudaStream_t s) {
  if (n == 0) return 0;
  Temp t(s);
  DevBuf<int> out(1, s);
  cub::TransformInputIterator<int, EqOp, const int*> it(in, EqOp{value});
  FSB_CUDA(cub::DeviceReduce::Sum(nullptr, t.bytes, it, out.get(), (int64_t)n, s));
  t.alloc();
  FSB_CUDA(cub::DeviceReduce::Sum(t.p, t.bytes, it, out.get(), (int64_t)n, s));
  return out.read(0);
}

void iota_i32(int* p, size_t n, cudaStream_t s) { if (n) iota_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, n); }
void fill_i32(int* p, size_t n, int v, cudaStream_t s) { if (n) fill_i_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, n, v); }
void fill_f64(double* p, size_t n, double v, cudaStream_t s) { if (n) fill_d_kernel<<<cdiv(n, 256), 256, 0, s>>>(p, n, v); }

}  // namespace fsb
