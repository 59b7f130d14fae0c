// TEST INFRASTRUCTURE - CPU oracle (see oracle/model.h header).
// threadpool.h: the CPU-baseline dispatcher, same contract as mjpc/threadpool.{h,cc}:30-85
// (FIFO queue under one mutex, thread-local WorkerId(), counter-based WaitCount/ResetCount).
#pragma once
#include <condition_variable>
#include <functional>
#include <mutex>
#include <queue>
#include <thread>
#include <vector>

namespace oracle {

class ThreadPool {
 public:
  explicit ThreadPool(int num_threads) : ctr_(0), stop_(false) {
    for (int i = 0; i < num_threads; i++) threads_.emplace_back(&ThreadPool::WorkerThread, this, i);
  }
  ~ThreadPool() {
    { std::unique_lock<std::mutex> lock(m_); stop_ = true; }
    cv_in_.notify_all();
    for (auto& t : threads_) t.join();
  }
  int NumThreads() const { return (int)threads_.size(); }
  static int WorkerId() { return worker_id_; }
  void Schedule(std::function<void()> task) {
    { std::unique_lock<std::mutex> lock(m_); queue_.push(std::move(task)); }
    cv_in_.notify_one();
  }
  void WaitCount(int value) {
    std::unique_lock<std::mutex> lock(m_);
    cv_ext_.wait(lock, [&]() { return ctr_ >= value; });
  }
  int GetCount() { std::unique_lock<std::mutex> lock(m_); return ctr_; }
  void ResetCount() { std::unique_lock<std::mutex> lock(m_); ctr_ = 0; }

 private:
  void WorkerThread(int i) {
    worker_id_ = i;
    while (true) {
      std::function<void()> task;
      {
        std::unique_lock<std::mutex> lock(m_);
        cv_in_.wait(lock, [&]() { return stop_ || !queue_.empty(); });
        if (stop_ && queue_.empty()) return;
        task = std::move(queue_.front());
        queue_.pop();
      }
      task();
      { std::unique_lock<std::mutex> lock(m_); ++ctr_; }
      cv_ext_.notify_all();
    }
  }
  static thread_local int worker_id_;
  std::vector<std::thread> threads_;
  std::mutex m_;
  std::condition_variable cv_in_, cv_ext_;
  std::queue<std::function<void()>> queue_;
  int ctr_;
  bool stop_;
};
inline thread_local int ThreadPool::worker_id_ = -1;

}  // namespace oracle
