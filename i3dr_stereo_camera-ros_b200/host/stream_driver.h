// B200FrameStream -- the frame-stream driver of SURVEY.md section 8e in C++: stereo pairs are independent, so a stream is
// dealt frame by frame over the GPUs of one box (frame i -> GPU i mod N).  One engine handle, one worker thread and `lanes`
// frames in flight per GPU; no NCCL, no torch, nothing crosses between GPUs.  Results land in the caller's buffers, which a
// caller that wants copies overlapped with kernels takes from b200sgm_host_alloc (page-locked).
// The reference has no counterpart: its node processes one frame at a time on one spinner thread
// (/root/reference/src/generate_disparity.cpp:526 "//TODO multithread this", :999).
#ifndef B200_STREAM_DRIVER_H
#define B200_STREAM_DRIVER_H

#include <condition_variable>
#include <cstdint>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "b200sgm.h"

class B200FrameStream
{
public:
  // devices: CUDA ordinals, one worker each.  Every engine is created for width x height x params.numDisparities with `lanes` lanes.
  B200FrameStream(const std::vector<int> &devices, int width, int height, const b200sgm_params &params, int lanes);
  ~B200FrameStream();
  B200FrameStream(const B200FrameStream &) = delete;
  B200FrameStream &operator=(const B200FrameStream &) = delete;

  bool ok() const { return error_.empty(); }
  const std::string &error() const { return error_; }
  int gpus() const { return int(workers_.size()); }

  // Queues frame number `id` of the stream (ids are dealt to GPU id mod N; submit them in stream order).  The three buffers are
  // tight CV_8UC1 / CV_8UC1 / CV_16S images of the stream's size.  Blocks while 2 * lanes frames already wait for that GPU, so
  // at most 3 * lanes frames per GPU are pending: a buffer may be reused once 3 * lanes + 1 later frames of the same GPU were
  // submitted (or after drain()).
  void submit(uint64_t id, const uint8_t *left, const uint8_t *right, int16_t *disp);
  // Blocks until every submitted frame is finished.  Returns 0, a positive warning code if some frame raised one, or the first
  // negative error code (error() then carries the message).
  int drain();
  // GPU that processed / will process frame `id`
  int owner(uint64_t id) const { return int(id % workers_.size()); }

private:
  struct Job { uint64_t id; const uint8_t *left, *right; int16_t *disp; };
  struct Worker {
    int device = 0;
    b200sgm_handle engine = nullptr;
    std::thread thread;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> queue;
    bool stop = false, idle = true;
    int status = 0;           // first error (< 0) or last warning (> 0)
    std::string message;
  };
  void run(Worker &w);

  int width_, height_, lanes_;
  std::vector<Worker *> workers_;
  std::string error_;
};

#endif // B200_STREAM_DRIVER_H
