// Process-level watchdog that does not need the Python GIL.
//
// A rank whose main thread is wedged inside a native call (a stream synchronise behind a kernel that spins on a
// peer's flag, a collective that never completes) cannot run Python timers reliably: whoever holds the GIL decides.
// This watchdog is one detached C++ thread: armed with a deadline and a complete (path, argv, envp) image, it either
// gets disarmed in time or replaces the process (execve keeps the PID, so a torchrun agent keeps supervising the
// same worker; the driver tears the CUDA context down with the old image, which also removes its spinning kernels).
// With an empty path the thread only reports and leaves with `exit_code`.
//
// Used by bench.py (fall back from the fused peer-memory path to the collective path instead of hanging a
// multi-GPU measurement) and by utils/watchdog.py (HCTR_STEP_TIMEOUT_ABORT).
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

struct Image {
  std::string path;
  std::vector<std::string> argv, envp;
  std::string message;
  int fd = 2;
  int exit_code = 124;
};

// never destroyed: a waiter may still sit on them while the process runs its static destructors at exit
// (pthread_cond_destroy would block on that waiter)
std::mutex& g_mu = *new std::mutex;
std::condition_variable& g_cv = *new std::condition_variable;
unsigned long long g_generation = 0;   // bumped by every arm / disarm: a sleeping thread of an older one retires
bool g_armed = false;

void fire(const Image& im) {
  if (!im.message.empty()) {
    // write(2), not stdio: the main thread may hold the stdio locks
    ssize_t r = ::write(im.fd, im.message.data(), im.message.size());
    (void)r;
  }
  if (im.path.empty()) ::_exit(im.exit_code);
  std::vector<char*> av, ev;
  for (auto& s : im.argv) av.push_back(const_cast<char*>(s.c_str()));
  for (auto& s : im.envp) ev.push_back(const_cast<char*>(s.c_str()));
  av.push_back(nullptr);
  ev.push_back(nullptr);
  ::execve(im.path.c_str(), av.data(), ev.data());
  const char* m = "[hctr watchdog] execve failed\n";
  ssize_t r = ::write(2, m, std::strlen(m));
  (void)r;
  ::_exit(im.exit_code);
}

}  // namespace

extern "C" {

// argv / envp: NULL-terminated arrays (copied).  path NULL or "": exit with `exit_code` instead of exec.
int hctr_exec_watchdog_arm(double seconds, const char* path, const char* const* argv, const char* const* envp,
                           const char* message, int message_fd, int exit_code) {
  Image im;
  if (path) im.path = path;
  for (const char* const* p = argv; p && *p; ++p) im.argv.emplace_back(*p);
  for (const char* const* p = envp; p && *p; ++p) im.envp.emplace_back(*p);
  if (message) im.message = message;
  im.fd = message_fd;
  im.exit_code = exit_code;
  unsigned long long gen;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    gen = ++g_generation;
    g_armed = true;
  }
  g_cv.notify_all();
  std::thread([im, gen, seconds]() {
    std::unique_lock<std::mutex> lk(g_mu);
    auto deadline = std::chrono::steady_clock::now() + std::chrono::duration<double>(seconds);
    while (g_generation == gen && g_armed) {
      if (g_cv.wait_until(lk, deadline) == std::cv_status::timeout) break;
    }
    if (g_generation != gen || !g_armed) return;
    lk.unlock();
    fire(im);
  }).detach();
  return 0;
}

void hctr_exec_watchdog_disarm() {
  {
    std::lock_guard<std::mutex> lk(g_mu);
    ++g_generation;
    g_armed = false;
  }
  g_cv.notify_all();
}

int hctr_exec_watchdog_armed() {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_armed ? 1 : 0;
}

}  // extern "C"
