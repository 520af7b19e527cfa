// host_pool.cpp -- persistent host worker threads behind r3d::parallel_for.
//
// The host tail of a matching batch (bucket copy, (i,j) sort, coordinate de-duplication with the
// reference's own std::set -- match_post.cpp) and the per-round AC-RANSAC state machines are short
// parallel loops issued hundreds of times per call, from several tail threads at once.  Spawning
// std::threads per loop cost more than the loop bodies; this pool keeps the threads alive and lets
// concurrent loops share them.  A loop may be entered from any thread, including from inside another
// pooled loop (the caller always works on its own loop, so nesting cannot deadlock).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <fstream>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace r3d {

namespace {

struct Job {
  size_t n = 0;
  const std::function<void(size_t)>* f = nullptr;
  std::atomic<size_t> next{0};
  std::atomic<size_t> done{0};
  std::atomic<int> helpers_wanted{0};
  std::mutex mu;
  std::condition_variable cv;
};

class HostPool {
 public:
  HostPool() {
    unsigned hw = std::thread::hardware_concurrency();
    if (hw < 2) hw = 2;
    // one process per GPU (torchrun / mpirun): share the host cores between the local ranks
    unsigned local_world = 1;
    for (const char* name : {"LOCAL_WORLD_SIZE", "OMPI_COMM_WORLD_LOCAL_SIZE", "SLURM_NTASKS_PER_NODE"}) {
      const char* v = std::getenv(name);
      if (v && std::atoi(v) > 0) { local_world = (unsigned)std::atoi(v); break; }
    }
    // a container may see every hardware thread but own only a CPU-time quota (cgroup v2 cpu.max = "quota period"):
    // more runnable threads than that burn the period's quota in a burst and are then throttled for the rest of it
    {
      std::ifstream f("/sys/fs/cgroup/cpu.max");
      std::string q;
      long long period = 0;
      if (f >> q >> period && q != "max" && period > 0) {
        const long long cpus = (std::atoll(q.c_str()) + period - 1) / period;
        if (cpus >= 1 && (unsigned long long)cpus < hw) hw = (unsigned)cpus;
      }
    }
    const char* e = std::getenv("R3D_HOST_THREADS");
    unsigned n = e && std::atoi(e) > 0 ? (unsigned)std::atoi(e) : std::max(4u, std::min(hw / local_world, 96u));
    for (unsigned t = 0; t + 1 < n; ++t) threads_.emplace_back([this]() { worker(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : threads_) t.join();
  }
  void run(int n_threads, size_t n, const std::function<void(size_t)>& f) {
    auto job = std::make_shared<Job>();
    job->n = n;
    job->f = &f;
    const int helpers = (int)std::min<size_t>((size_t)std::max(0, n_threads - 1), n - 1);
    job->helpers_wanted.store(helpers);
    if (helpers > 0) {
      {
        std::lock_guard<std::mutex> lk(mu_);
        jobs_.push_back(job);
      }
      if (helpers >= (int)threads_.size()) cv_.notify_all();
      else for (int k = 0; k < helpers; ++k) cv_.notify_one();
    }
    work(*job);
    if (helpers > 0) {
      {
        std::lock_guard<std::mutex> lk(mu_);  // nobody may join the finished loop any more
        for (auto it = jobs_.begin(); it != jobs_.end(); ++it)
          if (it->get() == job.get()) { jobs_.erase(it); break; }
      }
      std::unique_lock<std::mutex> lk(job->mu);
      job->cv.wait(lk, [&]() { return job->done.load() == job->n; });
    }
  }

 private:
  static void work(Job& j) {
    size_t mine = 0;
    for (;;) {
      const size_t i = j.next.fetch_add(1);
      if (i >= j.n) break;
      (*j.f)(i);
      ++mine;
    }
    if (mine && j.done.fetch_add(mine) + mine == j.n) {
      std::lock_guard<std::mutex> lk(j.mu);
      j.cv.notify_all();
    }
  }
  void worker() {
    for (;;) {
      std::shared_ptr<Job> job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&]() { return stop_ || pick(job); });
        if (!job) return;  // stop_
      }
      work(*job);
    }
  }
  // called with mu_ held: a queued loop that still wants helpers and still has items
  bool pick(std::shared_ptr<Job>& out) {
    for (auto& j : jobs_) {
      if (j->next.load() >= j->n) continue;
      if (j->helpers_wanted.fetch_sub(1) > 0) { out = j; return true; }
      j->helpers_wanted.fetch_add(1);
    }
    return false;
  }
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_;
  std::deque<std::shared_ptr<Job>> jobs_;
  bool stop_ = false;
};

HostPool& pool() {
  static HostPool* p = new HostPool();  // intentionally leaked: worker threads must outlive static destructors
  return *p;
}

}  // namespace

void pool_parallel_for(int n_threads, size_t n, const std::function<void(size_t)>& f) {
  if (n == 0) return;
  if (n_threads <= 1 || n == 1) {
    for (size_t i = 0; i < n; ++i) f(i);
    return;
  }
  pool().run(n_threads, n, f);
}

}  // namespace r3d
