// Error plumbing and version entry points of the C ABI (include/dasv_b200.h).
#include "common.cuh"
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

namespace dasv {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// Launch-configuration errors only (no host sync): asynchronous faults surface at the caller's
// next synchronisation, exactly like a stock PyTorch op on the same stream.
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}

bool pdl_enabled() {
    const char* e = getenv("DASV_PDL");
    return !(e && e[0] == '0');
}

int sm_count() {
    static int cached[64] = {0};                 // per device ordinal; the attribute never changes
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return 148; }
    int n = cached[dev];
    if (n <= 0) {
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) { cudaGetLastError(); n = 148; }
        cached[dev] = n;                         // benign race: every thread writes the same value
    }
    return n;
}

}  // namespace dasv

extern "C" int dasv_abi_version(void) { return 1; }
extern "C" const char* dasv_last_error(void) { return dasv::g_err; }

// Host-side helper of extract.py: the utterances of a batch lie scattered in the loader's pinned buffer; one plain
// cudaMemcpyAsync per utterance from C costs ~1.5 us of host time each (through torch's copy_ it is ~10 us, which at 256
// utterances was 8 % of BASELINE configs[3]).
extern "C" int dasv_h2d_segments(void* dst, const void* src_host, const long long* src_off, const long long* dst_off,
                                 const long long* nbytes, int n, void* stream) {
    if (n <= 0) return 0;
    if (!dst || !src_host || !src_off || !dst_off || !nbytes) { dasv::set_error("h2d_segments: null argument"); return 1; }
    for (int i = 0; i < n; ++i) {
        if (nbytes[i] <= 0) continue;
        if (src_off[i] < 0 || dst_off[i] < 0) { dasv::set_error("h2d_segments: negative offset in segment %d", i); return 1; }
        cudaError_t e = cudaMemcpyAsync(static_cast<char*>(dst) + dst_off[i], static_cast<const char*>(src_host) + src_off[i],
                                        static_cast<size_t>(nbytes[i]), cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream));
        if (e != cudaSuccess) { dasv::set_error("h2d_segments: segment %d: %s", i, cudaGetErrorString(e)); return 1; }
    }
    return 0;
}
