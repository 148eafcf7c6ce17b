// Misc entry points of libnmb200: version, error string, device query.
#include "common.cuh"

namespace nmb {
char *last_error_buffer() {
    static thread_local char buf[512] = {0};
    return buf;
}
}  // namespace nmb

extern "C" {

int nmb_abi_version(void) { return NMB_ABI_VERSION; }

const char *nmb_last_error(void) { return nmb::last_error_buffer(); }

int nmb_device_sm_count(void) {
    int dev = 0, n = 0;
    NMB_CUDA(cudaGetDevice(&dev));
    NMB_CUDA(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
    return n;
}

int nmb_program_bytes(void) { return nmb::kProgramBytesPerMotif; }

}  // extern "C"
