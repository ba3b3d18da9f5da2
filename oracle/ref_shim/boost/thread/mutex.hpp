// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
#include <mutex>
#include <condition_variable>
namespace boost {
class mutex : public std::mutex { public: typedef std::unique_lock<mutex> scoped_lock; };
class recursive_mutex : public std::recursive_mutex { public: typedef std::unique_lock<recursive_mutex> scoped_lock; };
template <class M> using lock_guard = std::lock_guard<M>; template <class M> using unique_lock = std::unique_lock<M>; }
