// Misc C-ABI entry points (include/sppark_b200.h): device probing, error-message ownership,
// introspection.  Reference: util/all_gpus.cpp:65-86.
#include "util/gpu.cuh"

std::atomic<uint64_t> g_launch_count{0};

extern "C" int cuda_available(void)
{
    try { return ngpus() != 0; } catch (...) { return 0; }
}

extern "C" void drop_error_message(char* msg) { free(msg); }

extern "C" int sppark_b200_sm_count(int device_id)
{
    try { return select_gpu(device_id).sm_count(); } catch (...) { return -1; }
}

extern "C" const char* sppark_b200_version(void) { return "sppark_b200 0.1 (sm_100a)"; }

extern "C" uint64_t sppark_b200_launch_count(void) { return g_launch_count.load(); }
