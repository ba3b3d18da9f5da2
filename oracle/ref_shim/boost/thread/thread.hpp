// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <thread>
#include <vector>
#include <functional>
#include "mutex.hpp"
namespace boost { using thread = std::thread; using std::ref;
class thread_group { std::vector<std::thread> t; public:
  thread_group() {} thread_group(const thread_group&) = delete;
  template <class F> void create_thread(F f) { t.emplace_back(f); }
  void join_all() { for (auto& x : t) if (x.joinable()) x.join(); }
  ~thread_group() { join_all(); } }; }
