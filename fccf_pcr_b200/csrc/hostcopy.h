// hostcopy.h — host->device copies from PAGEABLE caller memory at pinned-memory speed.
//
// The reference hands its clouds over as pcl::PointCloud / std::vector storage (FCCF.cpp:1655-1661, 1683):
// pageable memory, which cudaMemcpyAsync moves through the driver's single staging buffer at ~11 GB/s
// (r01: 84 % of the 10M-point command-line run).  CopyPool keeps a few worker threads, each with a few pinned
// chunks, its own stream and events: a copy is cut into chunks, the workers memcpy chunk after chunk into
// their pinned buffers and issue one cudaMemcpyAsync per chunk, so the CPU copy of one chunk overlaps the DMA
// of the previous ones and several cores feed the link.  The caller's stream then waits for the workers' events.
#pragma once
#include <cuda_runtime.h>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

namespace fccf {

class CopyPool {
 public:
  CopyPool(int device, int nworkers, size_t chunk_bytes);
  ~CopyPool();
  bool ok() const { return ok_; }
  // queue a pageable host -> device copy (cut into chunks)
  void add(void* dst_device, const void* src_host, size_t bytes);
  // wait until every queued chunk has been issued, then make `stream` wait for all of them
  cudaError_t flush_into(cudaStream_t stream);
  int workers() const { return (int)w_.size(); }

 private:
  struct Job { char* dst; const char* src; size_t bytes; };
  struct Worker {
    std::thread th;
    cudaStream_t stream = nullptr;
    char* pinned[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t slot_free[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t issued = nullptr;      // recorded after the worker's last issued chunk
    bool used = false;                 // issued something since the last flush
    int next = 0;
    double t_wait = 0, t_copy = 0, t_issue = 0, bytes = 0;   // FCCF_DEBUG_TIMING
  };
  void run(int wi);
  int device_; size_t chunk_; int nslots_ = 3; bool ok_ = false;
  std::vector<Worker> w_;
  std::mutex m_; std::condition_variable cv_job_, cv_done_;
  std::deque<Job> q_;
  int inflight_ = 0; bool stop_ = false; cudaError_t err_ = cudaSuccess;
};

// true when the runtime knows `p` as pinned / device / managed memory (a plain cudaMemcpyAsync is the fast path)
bool host_pointer_is_pinned(const void* p);

}  // namespace fccf
