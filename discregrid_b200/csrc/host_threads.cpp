// host_threads(): how many host threads are worth starting.  = CPUs in the affinity mask, capped by the cgroup v2 (cpu.max) or
// v1 (cpu.cfs_quota_us / cpu.cfs_period_us) quota, overridable with DG_HOST_THREADS.  Used by every multithreaded host pass of the
// library (BVH build, reduceField, OBJ reader, the connectivity table of dg_add_function_sdf).
#include "bvh_build.h"

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <sched.h>

namespace dgb {

static unsigned probe_host_threads()
{
    if (const char* e = std::getenv("DG_HOST_THREADS")) {
        const long v = std::strtol(e, nullptr, 10);
        if (v > 0) return (unsigned)v;
    }
    unsigned n = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof set, &set) == 0) n = (unsigned)CPU_COUNT(&set);
    if (n == 0) n = std::thread::hardware_concurrency();
    if (n == 0) n = 1;
    double quota = 0.0;
    if (FILE* f = std::fopen("/sys/fs/cgroup/cpu.max", "r")) {                 // cgroup v2: "<quota|max> <period>"
        char q[64] = {0}; long long period = 0;
        if (std::fscanf(f, "%63s %lld", q, &period) == 2 && q[0] != 'm' && period > 0) quota = std::strtod(q, nullptr) / (double)period;
        std::fclose(f);
    } else {
        long long q = -1, p = 0;
        if (FILE* fq = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_quota_us", "r")) { if (std::fscanf(fq, "%lld", &q) != 1) q = -1; std::fclose(fq); }
        if (FILE* fp = std::fopen("/sys/fs/cgroup/cpu/cpu.cfs_period_us", "r")) { if (std::fscanf(fp, "%lld", &p) != 1) p = 0; std::fclose(fp); }
        if (q > 0 && p > 0) quota = (double)q / (double)p;
    }
    if (quota >= 1.0 && quota < (double)n) n = (unsigned)quota;
    return n ? n : 1;
}

unsigned host_threads()
{
    static std::once_flag once;
    static unsigned n = 1;
    std::call_once(once, [] { n = probe_host_threads(); });
    return n;
}

}  // namespace dgb
