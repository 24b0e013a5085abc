// tools/visibility_probe.cu -- when does a store into mapped pinned host memory become visible to the polling host if the kernel
// that issued it KEEPS RUNNING (here: spins 30 us)?  Variants: store width (1 B plain, 1 B st.wt, 4 B, 16 B, a full 32-byte sector
// written by 8 lanes) and an optional fence.sys after the store.  Decides how the status bytes of the host transport must be written.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/visibility_probe tools/visibility_probe.cu && tools/visibility_probe
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 1; } } while (0)
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

__global__ void probe(uint8_t* host, int mode, int fence, uint32_t tag, unsigned long long spin_ns, uint8_t* dev_sink, int traffic) {
    const unsigned long long t0 = gtimer();
    uint8_t* slot = host + (size_t)blockIdx.x * 64;               // one 64-byte line per CTA
    const int lane = threadIdx.x;
    if (mode == 0) { if (lane == 0) slot[3] = (uint8_t)tag; }
    else if (mode == 1) { if (lane == 0) __stwt(slot + 3, (uint8_t)tag); }
    else if (mode == 2) { if (lane == 0) *reinterpret_cast<volatile uint32_t*>(slot) = tag * 0x01010101u; }
    else if (mode == 3) { if (lane == 0) *reinterpret_cast<uint4*>(slot) = make_uint4(tag * 0x01010101u, 1, 2, 3); }
    else if (mode == 4) { if (lane < 8) reinterpret_cast<uint32_t*>(slot)[lane] = tag * 0x01010101u; }      // a full sector, one warp store
    else if (mode == 5) { if (lane < 16) reinterpret_cast<uint32_t*>(slot)[lane] = tag * 0x01010101u; }     // a full line
    if (fence && lane == 0) __threadfence_system();
    // keep running (and optionally stream stores to device memory like the frame writer does)
    uint4* sink = reinterpret_cast<uint4*>(dev_sink) + (size_t)blockIdx.x * 65536;
    uint32_t i = 0;
    while (gtimer() - t0 < spin_ns) {
        if (traffic) { sink[(i * 32 + lane) & 65535] = make_uint4(i, tag, lane, 0); i++; }
    }
}

int main() {
    CK(cudaSetDevice(0));
    const int CTAS = 128;
    uint8_t* h; CK(cudaHostAlloc(&h, CTAS * 64, cudaHostAllocMapped));
    uint8_t* d; CK(cudaMalloc(&d, (size_t)CTAS * 65536 * 16));
    cudaStream_t s; CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    const char* names[] = {"1 B plain store", "1 B st.wt", "4 B volatile", "16 B vector", "32 B sector (8 lanes)", "64 B line (16 lanes)"};
    printf("kernel: %d CTAs x 32 threads, each stores into its own host line, then spins 30 us; host polls all lines\n", CTAS);
    for (int traffic = 0; traffic < 2; traffic++)
        for (int fence = 0; fence < 2; fence++)
            for (int mode = 0; mode < 6; mode++) {
                double first = 0, last = 0;
                const int reps = 200;
                for (int r = 1; r <= reps + 20; r++) {
                    memset(h, 0, CTAS * 64);
                    const uint32_t tag = (r & 0x7F) | 0x80;
                    const double t0 = now_us();
                    probe<<<CTAS, 32, 0, s>>>(h, mode, fence, tag, 30000ull, d, traffic);
                    const double t1 = now_us();
                    double tf = 0, tl = 0;
                    int seen = 0;
                    bool got[CTAS] = {};
                    while (seen < CTAS && now_us() - t1 < 200.0) {
                        for (int c = 0; c < CTAS; c++)
                            if (!got[c] && ((volatile uint8_t*)h)[c * 64 + 3] == (uint8_t)tag) { got[c] = true; seen++; tl = now_us(); if (seen == 1) tf = tl; }
                    }
                    CK(cudaStreamSynchronize(s));
                    if (r > 20) { first += tf - t1; last += tl - t1; }
                    (void)t0;
                }
                printf("  %s device traffic, %s fence.sys | %-22s: first line visible %6.2f us, last %6.2f us after the launch call returned\n",
                       traffic ? "with" : "no  ", fence ? "with" : "no  ", names[mode], first / reps, last / reps);
            }
    return 0;
}
