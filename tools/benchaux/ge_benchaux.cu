// ge_benchaux.cu — measurement helpers for bench.py.  NOT part of the product library (libgame_engine_b200.so):
// nothing here touches game state.  Built in-tree as tools/benchaux/libge_benchaux.so by game_engine_b200/build.py.
#include <cuda_runtime.h>

// Occupies a stream for about `ns` nanoseconds (one thread spinning on %globaltimer).  bench.py enqueues it in front
// of its first timing event so that the host can queue the timed launches while the device is still busy — the
// timed region then measures the device, not the host's launch rate from a cold queue.  The spin itself is
// outside the timed region.
__global__ void k_delay(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do { asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}

extern "C" int bx_stream_delay(int device, void* cuda_stream, unsigned microseconds) {
    if (microseconds > 1000000u) return -1;
    if (cudaSetDevice(device) != cudaSuccess) return -2;
    k_delay<<<1, 1, 0, (cudaStream_t)cuda_stream>>>((unsigned long long)microseconds * 1000ull);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
