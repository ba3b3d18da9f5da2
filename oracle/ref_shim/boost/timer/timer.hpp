// Stand-in for a header of an absent third-party library (Boost / OpenBabel), written for oracle/_ref only:
// it lets the reference's own Vina headers compile where they lie under /root/reference. No arithmetic lives here.
#pragma once
namespace boost { namespace timer { struct cpu_times { long long wall = 0, user = 0, system = 0; };
class cpu_timer { public: void start() {} void stop() {} void resume() {} cpu_times elapsed() const { return cpu_times(); } }; } }
