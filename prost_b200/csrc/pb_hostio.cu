// pb_hostio.cu -- copies between device memory and HOST buffers handed in through the C ABI.
//
// The reference moves whole iterates with thrust::copy into std::vector (solver.cu:152-167,
// backend_pdhg.cu:513-563): pageable memory, so the driver stages through its own bounce buffers with a
// single-threaded copy and -- for a freshly allocated destination -- one page fault per 4 KB.  At
// 4096^2 the final x, z, y, w read is 400 MB and took longer than 2000 PDHG iterations.  Here:
//   * pinned / registered host memory (cudaPointerGetAttributes) is read / written by one DMA;
//   * pageable memory goes through two pinned staging buffers per context: the DMA of chunk i+1
//     overlaps the copy-out of chunk i, and the copy-out is split over several host threads (each
//     thread also takes the first-touch page faults of its own part);
//   * prefault_host_range() lets the solver loop fault the destination in on a helper thread while
//     the GPU is still iterating.
#include "pb_common.cuh"

#include <sys/mman.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <map>
#include <mutex>
#include <thread>

namespace pb {

// ---- device memory cache (pb_common.cuh: device_alloc / device_free) ---------------------------------------
namespace {

struct DeviceCache {
  std::mutex mu;
  std::multimap<std::pair<int, size_t>, void*> blocks;    // (device, bytes) -> released block
  size_t cached = 0;
  size_t cap = [] {
    const char* e = getenv("PB_POOL_MB");
    return (e ? static_cast<size_t>(atoll(e)) : size_t(32768)) << 20;
  }();
};
DeviceCache& device_cache() {
  static DeviceCache* c = new DeviceCache();     // never destroyed: no CUDA calls during static teardown
  return *c;
}
constexpr size_t kMinCachedBytes = 1u << 20;     // small blocks are cheap to allocate; keep the map short

void release_blocks_of(int device) {              // device < 0: all devices
  DeviceCache& c = device_cache();
  std::vector<std::pair<int, void*>> victims;
  {
    std::lock_guard<std::mutex> lock(c.mu);
    for (auto it = c.blocks.begin(); it != c.blocks.end();) {
      if (device < 0 || it->first.first == device) {
        victims.emplace_back(it->first.first, it->second);
        c.cached -= it->first.second;
        it = c.blocks.erase(it);
      } else {
        ++it;
      }
    }
  }
  int cur = 0;
  cudaGetDevice(&cur);
  for (auto& v : victims) {
    cudaSetDevice(v.first);
    cudaFree(v.second);
  }
  cudaSetDevice(cur);
  cudaGetLastError();
}

}  // namespace

void* device_alloc(size_t bytes) {
  int dev = 0;
  cudaGetDevice(&dev);
  DeviceCache& c = device_cache();
  if (bytes >= kMinCachedBytes) {
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.blocks.find(std::make_pair(dev, bytes));
    if (it != c.blocks.end()) {
      void* p = it->second;
      c.cached -= bytes;
      c.blocks.erase(it);
      return p;
    }
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {             // give the cached blocks back and try once more
    cudaGetLastError();
    release_blocks_of(dev);
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    std::ostringstream ss;
    ss << "Out of memory: cudaMalloc of " << bytes << " bytes failed (" << cudaGetErrorString(e) << ")";
    fail(e == cudaErrorMemoryAllocation ? PB_ERR_OOM : PB_ERR_CUDA, ss.str());
  }
  return p;
}

void device_free(void* p, size_t bytes, int device) {
  if (!p) return;
  DeviceCache& c = device_cache();
  if (bytes >= kMinCachedBytes && c.cap > 0) {
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur != device) cudaSetDevice(device);
    // the block may still be in use by enqueued work: wait, exactly like cudaFree would
    const cudaError_t e = cudaDeviceSynchronize();
    if (cur != device) cudaSetDevice(cur);
    if (e == cudaSuccess) {
      std::lock_guard<std::mutex> lock(c.mu);
      if (c.cached + bytes <= c.cap) {
        c.blocks.emplace(std::make_pair(device, bytes), p);
        c.cached += bytes;
        return;
      }
    } else {
      cudaGetLastError();
    }
  }
  cudaFree(p);
  cudaGetLastError();
}

// ---- pinned host memory pool (pb_host_alloc / pb_host_free) -------------------------------------------------
// Result vectors of a solve (x, z, y, w: 400 MB at 4096^2) that live in pinned memory are written by one DMA at
// PCIe speed; cudaHostAlloc itself costs ~0.2 ms per MB, so released blocks are kept for exact-size reuse.
namespace {
struct HostPool {
  std::mutex mu;
  std::multimap<size_t, void*> blocks;
  std::map<void*, size_t> live;
  size_t cached = 0;
  size_t cap = [] {
    const char* e = getenv("PB_HOST_POOL_MB");
    return (e ? static_cast<size_t>(atoll(e)) : size_t(4096)) << 20;
  }();
};
HostPool& host_pool() {
  static HostPool* p = new HostPool();
  return *p;
}
}  // namespace

void* host_pool_alloc(size_t bytes) {
  HostPool& hp = host_pool();
  {
    std::lock_guard<std::mutex> lock(hp.mu);
    auto it = hp.blocks.find(bytes);
    if (it != hp.blocks.end()) {
      void* p = it->second;
      hp.blocks.erase(it);
      hp.cached -= bytes;
      hp.live[p] = bytes;
      return p;
    }
  }
  void* p = nullptr;
  const cudaError_t e = cudaHostAlloc(&p, std::max<size_t>(bytes, 1), cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    fail(PB_ERR_OOM, std::string("Out of memory: cudaHostAlloc failed (") + cudaGetErrorString(e) + ")");
  }
  std::lock_guard<std::mutex> lock(hp.mu);
  hp.live[p] = bytes;
  return p;
}

void host_pool_free(void* p) {
  if (!p) return;
  HostPool& hp = host_pool();
  size_t bytes = 0;
  {
    std::lock_guard<std::mutex> lock(hp.mu);
    auto it = hp.live.find(p);
    if (it == hp.live.end()) return;            // not ours
    bytes = it->second;
    hp.live.erase(it);
    if (hp.cached + bytes <= hp.cap) {
      hp.blocks.emplace(bytes, p);
      hp.cached += bytes;
      return;
    }
  }
  cudaFreeHost(p);
  cudaGetLastError();
}

void host_pool_release() {
  HostPool& hp = host_pool();
  std::vector<void*> victims;
  {
    std::lock_guard<std::mutex> lock(hp.mu);
    for (auto& b : hp.blocks) victims.push_back(b.second);
    hp.blocks.clear();
    hp.cached = 0;
  }
  for (void* p : victims) cudaFreeHost(p);
  cudaGetLastError();
}

void device_cache_release() { release_blocks_of(-1); host_pool_release(); }

namespace {

constexpr size_t kStageBytes = 32u << 20;
constexpr size_t kDirectBytes = 1u << 20;       // small copies: plain cudaMemcpyAsync
constexpr int kCopyThreads = 8;

bool host_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void ensure_staging(Context* ctx) {
  for (int i = 0; i < 2; ++i) {
    if (!ctx->stage[i]) PB_CUDA(cudaHostAlloc(&ctx->stage[i], kStageBytes, cudaHostAllocDefault));
    if (!ctx->stage_ev[i]) PB_CUDA(cudaEventCreateWithFlags(&ctx->stage_ev[i], cudaEventDisableTiming));
  }
}

// dst <- src, split over kCopyThreads threads on 4 KB boundaries
void parallel_copy(char* dst, const char* src, size_t bytes) {
  if (bytes < (4u << 20)) { std::memcpy(dst, src, bytes); return; }
  const size_t per = ((bytes / kCopyThreads) + 4095) & ~size_t(4095);
  std::thread th[kCopyThreads - 1];
  int started = 0;
  for (int t = 1; t < kCopyThreads; ++t) {
    const size_t off = per * t;
    if (off >= bytes) break;
    const size_t len = std::min(per, bytes - off);
    th[started++] = std::thread([=] { std::memcpy(dst + off, src + off, len); });
  }
  std::memcpy(dst, src, std::min(per, bytes));
  for (int t = 0; t < started; ++t) th[t].join();
}

}  // namespace

void release_host_staging(Context* ctx) {
  for (int i = 0; i < 2; ++i) {
    if (ctx->stage[i]) cudaFreeHost(ctx->stage[i]);
    if (ctx->stage_ev[i]) cudaEventDestroy(ctx->stage_ev[i]);
    ctx->stage[i] = nullptr;
    ctx->stage_ev[i] = nullptr;
  }
}

void download_to_host(Context* ctx, float* h, const float* d, size_t n) {
  const size_t bytes = n * sizeof(float);
  if (bytes == 0) return;
  cudaStream_t s = ctx->stream;
  if (bytes <= kDirectBytes || host_pinned(h)) {
    PB_CUDA(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaStreamSynchronize(s));
    return;
  }
  ensure_staging(ctx);
  const char* src = reinterpret_cast<const char*>(d);
  char* dst = reinterpret_cast<char*>(h);
  const size_t chunks = (bytes + kStageBytes - 1) / kStageBytes;
  auto issue = [&](size_t c) {
    const size_t off = c * kStageBytes, len = std::min(kStageBytes, bytes - off);
    PB_CUDA(cudaMemcpyAsync(ctx->stage[c & 1], src + off, len, cudaMemcpyDeviceToHost, s));
    PB_CUDA(cudaEventRecord(ctx->stage_ev[c & 1], s));
  };
  issue(0);
  if (chunks > 1) issue(1);
  for (size_t c = 0; c < chunks; ++c) {
    const size_t off = c * kStageBytes, len = std::min(kStageBytes, bytes - off);
    PB_CUDA(cudaEventSynchronize(ctx->stage_ev[c & 1]));
    parallel_copy(dst + off, static_cast<const char*>(ctx->stage[c & 1]), len);
    if (c + 2 < chunks) issue(c + 2);
  }
}

void upload_from_host(Context* ctx, float* d, const float* h, size_t n) {
  const size_t bytes = n * sizeof(float);
  if (bytes == 0) return;
  cudaStream_t s = ctx->stream;
  if (bytes <= kDirectBytes || host_pinned(h)) {
    PB_CUDA(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
    PB_CUDA(cudaStreamSynchronize(s));
    return;
  }
  ensure_staging(ctx);
  const char* src = reinterpret_cast<const char*>(h);
  char* dst = reinterpret_cast<char*>(d);
  const size_t chunks = (bytes + kStageBytes - 1) / kStageBytes;
  for (size_t c = 0; c < chunks; ++c) {
    const size_t off = c * kStageBytes, len = std::min(kStageBytes, bytes - off);
    if (c >= 2) PB_CUDA(cudaEventSynchronize(ctx->stage_ev[c & 1]));     // the DMA out of this slot is done
    parallel_copy(static_cast<char*>(ctx->stage[c & 1]), src + off, len);
    PB_CUDA(cudaMemcpyAsync(dst + off, ctx->stage[c & 1], len, cudaMemcpyHostToDevice, s));
    PB_CUDA(cudaEventRecord(ctx->stage_ev[c & 1], s));
  }
  PB_CUDA(cudaStreamSynchronize(s));
}

void prefault_host_range(void* p, size_t bytes) {
  if (!p || bytes < kDirectBytes || host_pinned(p)) return;
  const long page = sysconf(_SC_PAGESIZE) > 0 ? sysconf(_SC_PAGESIZE) : 4096;
  char* b = static_cast<char*>(p);
  char* e = b + bytes;
  char* ab = reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(b) + page - 1) & ~uintptr_t(page - 1));
  char* ae = reinterpret_cast<char*>(reinterpret_cast<uintptr_t>(e) & ~uintptr_t(page - 1));
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
  if (ae > ab && madvise(ab, static_cast<size_t>(ae - ab), MADV_POPULATE_WRITE) == 0) return;
  // older kernels: rewrite one byte per page with its own value (allocates the page, keeps the content)
  for (volatile char* q = b; q < e; q += page) *q = *q;
}

}  // namespace pb
