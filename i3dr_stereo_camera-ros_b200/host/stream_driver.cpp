#include "stream_driver.h"

B200FrameStream::B200FrameStream(const std::vector<int> &devices, int width, int height, const b200sgm_params &params, int lanes)
    : width_(width), height_(height), lanes_(lanes < 1 ? 1 : lanes)
{
  if (devices.empty()) { error_ = "no devices"; return; }
  for (int dev : devices) {
    Worker *w = new Worker();
    w->device = dev;
    int rc = b200sgm_create(dev, width, height, params.numDisparities > 0 ? params.numDisparities : 1, lanes_, &w->engine);
    if (rc == B200SGM_OK) rc = b200sgm_set_params(w->engine, &params);
    if (rc != B200SGM_OK) {
      error_ = "device " + std::to_string(dev) + ": " + (w->engine ? b200sgm_last_error(w->engine) : "b200sgm_create failed");
      if (w->engine) b200sgm_destroy(w->engine);
      delete w;
      return;
    }
    workers_.push_back(w);
  }
  for (Worker *w : workers_) w->thread = std::thread([this, w] { run(*w); });
}

B200FrameStream::~B200FrameStream()
{
  for (Worker *w : workers_) {
    { std::lock_guard<std::mutex> lk(w->mu); w->stop = true; }
    w->cv.notify_all();
  }
  for (Worker *w : workers_) {
    if (w->thread.joinable()) w->thread.join();
    if (w->engine) b200sgm_destroy(w->engine);
    delete w;
  }
}

void B200FrameStream::submit(uint64_t id, const uint8_t *left, const uint8_t *right, int16_t *disp)
{
  if (workers_.empty()) return;
  Worker &w = *workers_[id % workers_.size()];
  {
    // back-pressure: at most 2 * lanes frames wait per GPU (with the `lanes` in flight that bounds what a caller must keep alive)
    std::unique_lock<std::mutex> lk(w.mu);
    w.cv.wait(lk, [&] { return w.queue.size() < size_t(2 * lanes_); });
    w.queue.push_back(Job{id, left, right, disp});
    w.idle = false;
  }
  w.cv.notify_all();
}

int B200FrameStream::drain()
{
  int status = 0;
  for (Worker *w : workers_) {
    std::unique_lock<std::mutex> lk(w->mu);
    w->cv.wait(lk, [w] { return w->idle; });
    if (w->status < 0 && status >= 0) { status = w->status; error_ = w->message; }
    else if (w->status > 0 && status == 0) status = w->status;
    w->status = 0;
  }
  return status;
}

// Worker of one GPU: frames of its queue go to lanes 0, 1, ..., lanes-1, 0, ...; a lane is waited on right before it is reused, and
// every lane once the queue runs dry (so drain() only returns when the results are in the caller's buffers).
void B200FrameStream::run(Worker &w)
{
  std::vector<char> busy(size_t(lanes_), 0);
  uint64_t issued = 0;
  auto note = [&](int rc) {
    if (rc == 0) return;
    std::lock_guard<std::mutex> lk(w.mu);
    if (rc < 0 && w.status >= 0) { w.status = rc; w.message = b200sgm_last_error(w.engine); }
    else if (rc > 0 && w.status == 0) w.status = rc;
  };
  for (;;) {
    Job job;
    {
      std::unique_lock<std::mutex> lk(w.mu);
      if (w.queue.empty()) {
        lk.unlock();
        for (int ln = 0; ln < lanes_; ln++)
          if (busy[ln]) { note(b200sgm_wait(w.engine, ln)); busy[ln] = 0; }
        lk.lock();
        if (w.queue.empty()) { w.idle = true; w.cv.notify_all(); }
        w.cv.wait(lk, [&w] { return w.stop || !w.queue.empty(); });
        if (w.queue.empty()) return;   // stop
      }
      job = w.queue.front();
      w.queue.pop_front();
    }
    w.cv.notify_all();     // a submit() may be waiting for room
    const int ln = int(issued++ % uint64_t(lanes_));
    if (busy[ln]) { note(b200sgm_wait(w.engine, ln)); busy[ln] = 0; }
    const int rc = b200sgm_enqueue(w.engine, ln, job.left, size_t(width_), job.right, size_t(width_), width_, height_, job.disp,
                                   size_t(width_) * 2);
    if (rc == 0) busy[ln] = 1;
    else note(rc);
  }
}
