// cal_dev.h — thin device layer of the engine.
//
// Product build (nvcc, sm_100a): CUDA runtime, CUB for the sparse-hit-list sorts/scans.
// Test-only build (-DCAL_HOSTSIM, g++): the same kernel bodies and the same engine code run serially on the host so that
// tests can check the engine's logic and plumbing where no GPU exists.  The hostsim library is built by tests/hostsim/ into
// tests/hostsim/_build and is never part of calitas_b200/*.so: the product has no CPU path.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>

#ifndef CAL_HOSTSIM
// ======================================================== CUDA ===========================================================
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#define CAL_KERNEL __global__ void
#define CAL_MAXNREG(n) __maxnreg__(n)
#define CAL_PHASE(k)
#define CAL_SHARED_DYN(type, name) extern __shared__ __align__(16) unsigned char cal_smem_raw_[]; type* name = reinterpret_cast<type*>(cal_smem_raw_)
#define CAL_LAUNCH(kernel, grid, block, smem, stream, nphases, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)

namespace cal { namespace dev {

struct Error { std::string msg; };
inline void check(cudaError_t e, const char* what) { if (e != cudaSuccess) throw Error{ std::string(what) + ": " + cudaGetErrorString(e) }; }

typedef cudaStream_t Stream;
typedef cudaEvent_t Event;

inline void init(int device) {
  int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) throw Error{ std::string("no CUDA device available (this engine has no CPU fallback): ") + cudaGetErrorString(e) };
  if (device < 0 || device >= n) throw Error{ "device id out of range" };
  check(cudaSetDevice(device), "cudaSetDevice");
}
inline void set_device(int device) { check(cudaSetDevice(device), "cudaSetDevice"); }
inline int sm_count(int device) { int v = 0; check(cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device), "attr"); return v; }
inline Stream stream_create() { Stream s; check(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate"); return s; }
// high != 0: the greatest priority the device offers, else the least (pending thread blocks of a high-priority stream are dispatched first)
inline Stream stream_create_prio(int high) {
  int least = 0, greatest = 0; check(cudaDeviceGetStreamPriorityRange(&least, &greatest), "cudaDeviceGetStreamPriorityRange");
  Stream s; check(cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, high ? greatest : least), "cudaStreamCreateWithPriority"); return s;
}
inline void stream_wait(Stream s, cudaEvent_t e) { check(cudaStreamWaitEvent(s, e, 0), "cudaStreamWaitEvent"); }
inline void event_sync(cudaEvent_t e) { check(cudaEventSynchronize(e), "cudaEventSynchronize"); }
inline void stream_destroy(Stream s) { cudaStreamDestroy(s); }
inline void stream_sync(Stream s) { check(cudaStreamSynchronize(s), "cudaStreamSynchronize"); }
inline Event event_create() { Event e; check(cudaEventCreate(&e), "cudaEventCreate"); return e; }
inline void event_destroy(Event e) { cudaEventDestroy(e); }
inline void event_record(Event e, Stream s) { check(cudaEventRecord(e, s), "cudaEventRecord"); }
inline double event_ms(Event a, Event b) { float ms = 0; check(cudaEventElapsedTime(&ms, a, b), "cudaEventElapsedTime"); return ms; }
inline void* alloc(size_t bytes) { void* p = nullptr; check(cudaMalloc(&p, bytes ? bytes : 1), "cudaMalloc"); return p; }
inline void free_(void* p) { if (p) cudaFree(p); }
inline void* alloc_host(size_t bytes) { void* p = nullptr; check(cudaMallocHost(&p, bytes ? bytes : 1), "cudaMallocHost"); return p; }
inline void free_host(void* p) { if (p) cudaFreeHost(p); }
// Pinned host memory the device can write directly (zero-copy).  Counters the host waits for go through it instead of through cudaMemcpyAsync:
// a 4-byte D2H copy queues behind every bulk D2H copy already submitted to the copy engine (measured: each chunk's tail waited ~60 ms for the
// previous chunk's 3-GB hit copy at d = 6), a store from a kernel does not.
inline void* alloc_host_mapped(size_t bytes, void** dev_alias) {
  void* p = nullptr; check(cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocMapped), "cudaHostAlloc"); check(cudaHostGetDevicePointer(dev_alias, p, 0), "cudaHostGetDevicePointer"); return p;
}
inline void h2d(void* d, const void* h, size_t n, Stream s) { if (n) check(cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s), "cudaMemcpyAsync H2D"); }
inline void d2h(void* h, const void* d, size_t n, Stream s) { if (n) check(cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s), "cudaMemcpyAsync D2H"); }
inline void d2d(void* dst, const void* src, size_t n, Stream s) { if (n) check(cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, s), "cudaMemcpyAsync D2D"); }
inline void zero(void* d, size_t n, Stream s) { if (n) check(cudaMemsetAsync(d, 0, n, s), "cudaMemsetAsync"); }
inline void launch_check(const char* what) { check(cudaGetLastError(), what); }

// stable LSD radix sorts (CUB) — library plumbing for the sparse hit lists, not a hot op
inline size_t sort_pairs_u64_tmp(size_t n, int begin_bit, int end_bit) {
  size_t bytes = 0; cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, n, begin_bit, end_bit);
  return bytes;
}
inline void sort_pairs_u64(void* tmp, size_t tmp_bytes, const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, size_t n, int begin_bit, int end_bit, Stream s) {
  check(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, kin, kout, vin, vout, n, begin_bit, end_bit, s), "cub SortPairs");
}
inline size_t sort_keys_u64_tmp(size_t n, int begin_bit, int end_bit) {
  size_t bytes = 0; cub::DeviceRadixSort::SortKeys(nullptr, bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, n, begin_bit, end_bit);
  return bytes;
}
inline void sort_keys_u64(void* tmp, size_t tmp_bytes, const uint64_t* kin, uint64_t* kout, size_t n, int begin_bit, int end_bit, Stream s) {
  check(cub::DeviceRadixSort::SortKeys(tmp, tmp_bytes, kin, kout, n, begin_bit, end_bit, s), "cub SortKeys");
}
inline size_t exclusive_sum_u32_tmp(size_t n) { size_t bytes = 0; cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, n); return bytes; }
inline void exclusive_sum_u32(void* tmp, size_t tmp_bytes, const uint32_t* in, uint32_t* out, size_t n, Stream s) {
  check(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, in, out, n, s), "cub ExclusiveSum");
}

}}  // namespace cal::dev

#else
// ====================================================== HOSTSIM (tests only) ==================================================
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

namespace cal { namespace sim {
struct Dim3 { unsigned x = 1, y = 1, z = 1; Dim3() {} Dim3(unsigned a) : x(a) {} };
extern thread_local Dim3 threadIdx_, blockIdx_, blockDim_, gridDim_;      // thread_local: engines may be driven from several host threads
extern thread_local int phase_;
extern thread_local unsigned char smem_[256 * 1024];
template <class F> void launch(unsigned grid, unsigned block, int nphases, F f) {
  gridDim_ = Dim3(grid); blockDim_ = Dim3(block);
  for (unsigned b = 0; b < grid; ++b) { blockIdx_ = Dim3(b);
    for (int ph = 0; ph < nphases; ++ph) { phase_ = ph;
      for (unsigned t = 0; t < block; ++t) { threadIdx_ = Dim3(t); f(); } } }
}
}}
#define threadIdx cal::sim::threadIdx_
#define blockIdx cal::sim::blockIdx_
#define blockDim cal::sim::blockDim_
#define gridDim cal::sim::gridDim_
#define CAL_KERNEL static void
#define CAL_MAXNREG(n)
#define CAL_PHASE(k) if (cal::sim::phase_ == (k))
#define CAL_SHARED_DYN(type, name) type* name = reinterpret_cast<type*>(cal::sim::smem_)
#define CAL_LAUNCH(kernel, grid, block, smem, stream, nphases, ...) cal::sim::launch((unsigned)(grid), (unsigned)(block), (nphases), [&] { kernel(__VA_ARGS__); })
#define __launch_bounds__(...)
#define __restrict__
inline void __syncthreads() {}
template <class T> inline T __ldg(const T* p) { return *p; }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { unsigned o = *p; *p += v; return o; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }

namespace cal { namespace dev {
struct Error { std::string msg; };
typedef int Stream;
typedef std::chrono::steady_clock::time_point* Event;
inline void init(int) {}
inline void set_device(int) {}
inline int sm_count(int) { return 4; }
inline Stream stream_create() { return 0; }
inline Stream stream_create_prio(int) { return 0; }
inline void stream_destroy(Stream) {}
inline void stream_sync(Stream) {}
inline Event event_create() { return new std::chrono::steady_clock::time_point(); }
inline void event_destroy(Event e) { delete e; }
inline void event_record(Event e, Stream) { *e = std::chrono::steady_clock::now(); }
inline double event_ms(Event a, Event b) { return std::chrono::duration<double, std::milli>(*b - *a).count(); }
inline void stream_wait(Stream, Event) {}
inline void event_sync(Event) {}
inline void* alloc(size_t bytes) { return std::calloc(bytes ? bytes : 1, 1); }
inline void free_(void* p) { std::free(p); }
inline void* alloc_host(size_t bytes) { return std::calloc(bytes ? bytes : 1, 1); }
inline void free_host(void* p) { std::free(p); }
inline void* alloc_host_mapped(size_t bytes, void** dev_alias) { void* p = std::calloc(bytes ? bytes : 1, 1); *dev_alias = p; return p; }
inline void __threadfence_system_() {}
inline void h2d(void* d, const void* h, size_t n, Stream) { if (n) std::memcpy(d, h, n); }
inline void d2h(void* h, const void* d, size_t n, Stream) { if (n) std::memcpy(h, d, n); }
inline void d2d(void* dst, const void* src, size_t n, Stream) { if (n) std::memmove(dst, src, n); }
inline void zero(void* d, size_t n, Stream) { if (n) std::memset(d, 0, n); }
inline void launch_check(const char*) {}
inline size_t sort_pairs_u64_tmp(size_t, int, int) { return 1; }
inline void sort_pairs_u64(void*, size_t, const uint64_t* kin, uint64_t* kout, const uint32_t* vin, uint32_t* vout, size_t n, int begin_bit, int end_bit, Stream) {
  std::vector<uint32_t> ord(n); std::iota(ord.begin(), ord.end(), 0u);
  uint64_t mask = (end_bit - begin_bit >= 64) ? ~0ull : (((1ull << (end_bit - begin_bit)) - 1) << begin_bit);
  std::stable_sort(ord.begin(), ord.end(), [&](uint32_t a, uint32_t b) { return (kin[a] & mask) < (kin[b] & mask); });
  for (size_t i = 0; i < n; ++i) { kout[i] = kin[ord[i]]; vout[i] = vin[ord[i]]; }
}
inline size_t sort_keys_u64_tmp(size_t, int, int) { return 1; }
inline void sort_keys_u64(void*, size_t, const uint64_t* kin, uint64_t* kout, size_t n, int begin_bit, int end_bit, Stream) {
  std::vector<uint64_t> k(kin, kin + n);
  uint64_t mask = (end_bit - begin_bit >= 64) ? ~0ull : (((1ull << (end_bit - begin_bit)) - 1) << begin_bit);
  std::stable_sort(k.begin(), k.end(), [&](uint64_t a, uint64_t b) { return (a & mask) < (b & mask); });
  std::copy(k.begin(), k.end(), kout);
}
inline size_t exclusive_sum_u32_tmp(size_t) { return 1; }
inline void exclusive_sum_u32(void*, size_t, const uint32_t* in, uint32_t* out, size_t n, Stream) { uint32_t acc = 0; for (size_t i = 0; i < n; ++i) { uint32_t v = in[i]; out[i] = acc; acc += v; } }
}}  // namespace cal::dev
#endif
