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

phase_profile_t g_profile;

extern "C" void sppark_b200_profile_enable(int on) { g_profile.enabled = on != 0; g_profile.reset(); }

// writes up to `cap` (name, ms) pairs for the phases of the last profiled call; returns count.
// ms[i] is the time from mark i to mark i+1.  Call after synchronising the stream.
extern "C" int sppark_b200_profile_read(const char** names, float* ms, int cap)
{
    int k = 0;
    for (int i = 0; i + 1 < g_profile.n && k < cap; i++, k++) {
        float t = 0;
        if (cudaEventElapsedTime(&t, g_profile.ev[i], g_profile.ev[i + 1]) != cudaSuccess) t = -1;
        names[k] = g_profile.name[i];
        ms[k] = t;
    }
    return k;
}
