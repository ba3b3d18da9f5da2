// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <mutex>
namespace boost { using mutex = std::mutex; using recursive_mutex = std::recursive_mutex;
template <class M> using lock_guard = std::lock_guard<M>; template <class M> using unique_lock = std::unique_lock<M>; }
