// Library-level C-ABI entry points (version, error strings, launch counter).
#include "common.cuh"

using namespace simulst;

extern "C" {

int simulst_version(void) { return SIMULST_VERSION; }

const char* simulst_error_string(int code) {
    switch (code) {
        case SIMULST_OK: return "ok";
        case SIMULST_E_ARG: return "invalid argument (null pointer, bad dtype enum or flag combination)";
        case SIMULST_E_SHAPE: return "unsupported shape (negative, or source length beyond SIMULST_MMA_MAX_SRC)";
        case SIMULST_E_ARCH: return "current CUDA device is not sm_100 (B200)";
        case SIMULST_E_LAUNCH: return "CUDA launch failed";
        case SIMULST_E_ALIGN: return "pointer not aligned for its element type";
        default: return "unknown error";
    }
}

long long simulst_launch_count(void) { return LaunchCounter::value().load(std::memory_order_relaxed); }
void simulst_reset_launch_count(void) { LaunchCounter::value().store(0, std::memory_order_relaxed); }

}  // extern "C"
