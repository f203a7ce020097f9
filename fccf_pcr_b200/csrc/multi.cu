// multi.cu — the parts of libfccf that span host threads and GPUs:
//   CopyPool                   pageable host memory -> device through pinned chunks filled by worker threads
//   fccf_register_batch_multi  BASELINE config 4 inside the library: one host thread per context / GPU, the
//                              pairs cut into contiguous blocks, no data-path collective
//   fccf_score_sharded         BASELINE config 3: an ordered hypothesis list cut into contiguous ranges over
//                              the contexts (static voxel table and moving cloud replicated), the global best
//                              by ONE 8-byte ncclAllReduce(max) of the packed (score, index) word over NVLink
//                              (libnccl.so.2 is loaded at run time; without it the 8-byte words are compared on
//                              the host and *used_nccl says so)
#include <dlfcn.h>
#include <chrono>
#include <cstdio>
#include <nccl.h>
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include "hostcopy.h"
#include "fccf_internal.h"

namespace fccf {

bool host_pointer_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

CopyPool::CopyPool(int device, int nworkers, size_t chunk_bytes) : device_(device), chunk_(chunk_bytes) {
  cudaSetDevice(device_);
  if (const char* e = getenv("FCCF_COPY_CHUNK_KB")) { long kb = atol(e); if (kb >= 64) chunk_ = (size_t)kb << 10; }
  if (const char* e = getenv("FCCF_COPY_SLOTS")) { int v = atoi(e); if (v >= 2 && v <= 4) nslots_ = v; }
  w_.resize(nworkers);
  ok_ = true;
  for (Worker& w : w_) {
    ok_ = ok_ && cudaStreamCreateWithFlags(&w.stream, cudaStreamNonBlocking) == cudaSuccess;
    for (int s = 0; s < nslots_; s++) {
      ok_ = ok_ && cudaMallocHost(&w.pinned[s], chunk_) == cudaSuccess;
      ok_ = ok_ && cudaEventCreateWithFlags(&w.slot_free[s], cudaEventDisableTiming) == cudaSuccess;
    }
    ok_ = ok_ && cudaEventCreateWithFlags(&w.issued, cudaEventDisableTiming) == cudaSuccess;
  }
  if (!ok_) { cudaGetLastError(); return; }
  for (int i = 0; i < nworkers; i++) w_[i].th = std::thread(&CopyPool::run, this, i);
}

CopyPool::~CopyPool() {
  { std::lock_guard<std::mutex> l(m_); stop_ = true; }
  cv_job_.notify_all();
  for (Worker& w : w_) if (w.th.joinable()) w.th.join();
  if (getenv("FCCF_DEBUG_TIMING")) for (size_t i = 0; i < w_.size(); i++) fprintf(stderr, "[fccf] copy worker %zu: %.1f MB, wait %.1f ms, memcpy %.1f ms (%.1f GB/s), issue %.1f ms\n", i, w_[i].bytes / 1e6, 1e3 * w_[i].t_wait, 1e3 * w_[i].t_copy, w_[i].bytes / 1e9 / (w_[i].t_copy > 0 ? w_[i].t_copy : 1), 1e3 * w_[i].t_issue);
  cudaSetDevice(device_);
  for (Worker& w : w_) {
    if (w.stream) { cudaStreamSynchronize(w.stream); cudaStreamDestroy(w.stream); }
    for (int s = 0; s < 4; s++) { if (w.pinned[s]) cudaFreeHost(w.pinned[s]); if (w.slot_free[s]) cudaEventDestroy(w.slot_free[s]); }
    if (w.issued) cudaEventDestroy(w.issued);
  }
}

void CopyPool::run(int wi) {
  cudaSetDevice(device_);
  Worker& w = w_[wi];
  bool primed[4] = {false, false, false, false};
  while (true) {
    Job j;
    {
      std::unique_lock<std::mutex> l(m_);
      cv_job_.wait(l, [&] { return stop_ || !q_.empty(); });
      if (stop_ && q_.empty()) return;
      j = q_.front(); q_.pop_front();
    }
    const int s = w.next; w.next = (w.next + 1) % nslots_;
    cudaError_t e = cudaSuccess;
    auto t0 = std::chrono::steady_clock::now();
    if (primed[s]) e = cudaEventSynchronize(w.slot_free[s]);     // the DMA that last used this pinned chunk is done
    auto t1 = std::chrono::steady_clock::now();
    if (e == cudaSuccess) {
      memcpy(w.pinned[s], j.src, j.bytes);
      auto t2 = std::chrono::steady_clock::now();
      e = cudaMemcpyAsync(j.dst, w.pinned[s], j.bytes, cudaMemcpyHostToDevice, w.stream);
      if (e == cudaSuccess) e = cudaEventRecord(w.slot_free[s], w.stream);
      primed[s] = true;
      auto t3 = std::chrono::steady_clock::now();
      w.t_wait += std::chrono::duration<double>(t1 - t0).count(); w.t_copy += std::chrono::duration<double>(t2 - t1).count(); w.t_issue += std::chrono::duration<double>(t3 - t2).count(); w.bytes += j.bytes;
    }
    {
      std::lock_guard<std::mutex> l(m_);
      if (e != cudaSuccess && err_ == cudaSuccess) err_ = e;
      w.used = true;
      inflight_--;
    }
    cv_done_.notify_all();
  }
}

void CopyPool::add(void* dst, const void* src, size_t bytes) {
  std::lock_guard<std::mutex> l(m_);
  for (size_t off = 0; off < bytes; off += chunk_) {
    Job j; j.dst = (char*)dst + off; j.src = (const char*)src + off; j.bytes = std::min(chunk_, bytes - off);
    q_.push_back(j); inflight_++;
  }
  cv_job_.notify_all();
}

cudaError_t CopyPool::flush_into(cudaStream_t stream) {
  std::unique_lock<std::mutex> l(m_);
  cv_done_.wait(l, [&] { return inflight_ == 0; });
  cudaError_t e = err_; err_ = cudaSuccess;
  for (Worker& w : w_) {
    if (!w.used) continue;
    w.used = false;
    if (e == cudaSuccess) e = cudaEventRecord(w.issued, w.stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(stream, w.issued, 0);
  }
  return e;
}

// ---- NCCL, loaded at run time ---------------------------------------------------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  NcclApi() {
    if (const char* e = getenv("FCCF_NO_NCCL")) if (e[0] == '1') return;
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    CommInitAll = (decltype(CommInitAll))dlsym(h, "ncclCommInitAll");
    CommDestroy = (decltype(CommDestroy))dlsym(h, "ncclCommDestroy");
    AllReduce = (decltype(AllReduce))dlsym(h, "ncclAllReduce");
    GroupStart = (decltype(GroupStart))dlsym(h, "ncclGroupStart");
    GroupEnd = (decltype(GroupEnd))dlsym(h, "ncclGroupEnd");
    GetErrorString = (decltype(GetErrorString))dlsym(h, "ncclGetErrorString");
    ok = CommInitAll && CommDestroy && AllReduce && GroupStart && GroupEnd;
  }
};
static NcclApi& nccl_api() { static NcclApi a; return a; }

struct NcclComms { std::vector<ncclComm_t> comm; std::vector<long long*> d_send, d_recv; };
static std::mutex g_nccl_mutex;
static std::map<std::string, NcclComms> g_nccl_comms;      // by device list

// communicators (and one send / receive word per device) for the given devices, created once
static NcclComms* nccl_comms_for(const std::vector<int>& devs) {
  NcclApi& api = nccl_api();
  if (!api.ok) return nullptr;
  std::string key;
  for (int d : devs) key += std::to_string(d) + ",";
  std::lock_guard<std::mutex> l(g_nccl_mutex);
  auto it = g_nccl_comms.find(key);
  if (it != g_nccl_comms.end()) return it->second.comm.empty() ? nullptr : &it->second;
  NcclComms c;
  c.comm.resize(devs.size());
  if (api.CommInitAll(c.comm.data(), (int)devs.size(), devs.data()) != ncclSuccess) { g_nccl_comms[key] = NcclComms(); return nullptr; }
  for (int d : devs) {
    long long *s = nullptr, *r = nullptr;
    cudaSetDevice(d);
    if (cudaMalloc(&s, 8) != cudaSuccess || cudaMalloc(&r, 8) != cudaSuccess) { cudaGetLastError(); g_nccl_comms[key] = NcclComms(); return nullptr; }
    c.d_send.push_back(s); c.d_recv.push_back(r);
  }
  g_nccl_comms[key] = c;
  return &g_nccl_comms[key];
}

// Signed maximum of one 8-byte word per device: ncclAllReduce(max, int64) on the given streams.  words_dev[i] is a
// device pointer on devs[i]; the result lands in out[i] (host).  false: NCCL unavailable.
bool nccl_allreduce_max_i64(const std::vector<int>& devs, const std::vector<cudaStream_t>& streams, const std::vector<const long long*>& words_dev, long long* out) {
  NcclComms* c = nccl_comms_for(devs);
  if (!c) return false;
  NcclApi& api = nccl_api();
  const size_t n = devs.size();
  for (size_t i = 0; i < n; i++) { cudaSetDevice(devs[i]); if (cudaMemcpyAsync(c->d_send[i], words_dev[i], 8, cudaMemcpyDeviceToDevice, streams[i]) != cudaSuccess) return false; }
  if (api.GroupStart() != ncclSuccess) return false;
  bool ok = true;
  for (size_t i = 0; i < n; i++) ok = ok && api.AllReduce(c->d_send[i], c->d_recv[i], 1, ncclInt64, ncclMax, c->comm[i], streams[i]) == ncclSuccess;
  if (api.GroupEnd() != ncclSuccess) ok = false;
  for (size_t i = 0; i < n; i++) {
    cudaSetDevice(devs[i]);
    if (cudaMemcpyAsync(&out[i], c->d_recv[i], 8, cudaMemcpyDeviceToHost, streams[i]) != cudaSuccess) ok = false;
    if (cudaStreamSynchronize(streams[i]) != cudaSuccess) ok = false;
  }
  return ok;
}

}  // namespace fccf
