// Status plumbing of the C ABI (include/afs_b200.h).
#include "common.cuh"

#include <atomic>

namespace afs {
thread_local int g_last_cuda_error = 0;
static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}

extern "C" int afs_abi_version(void) { return AFS_ABI_VERSION; }

extern "C" const char* afs_status_string(int status) {
  switch (status) {
    case AFS_OK: return "ok";
    case AFS_ERR_INVALID_ARG: return "invalid argument";
    case AFS_ERR_UNSUPPORTED: return "unsupported configuration";
    case AFS_ERR_CUDA: return "CUDA runtime error";
    case AFS_ERR_WORKSPACE: return "workspace too small";
    default: return "unknown status";
  }
}

extern "C" int afs_last_cuda_error(void) { return afs::g_last_cuda_error; }

extern "C" uint64_t afs_launch_count(void) { return afs::g_launches.load(std::memory_order_relaxed); }
