// tools/pcie_probe.cu -- what does the GPU <-> pinned-host-memory path cost on this box?  Decides the host transport design:
//   A  scattered small stores from the GPU into mapped pinned host memory (the device patching the caller's frame mirror itself):
//      kernel time for n_worlds x stores_per_world stores of 1 / 4 / 12 (3x4) bytes at "random frame" addresses
//   B  launch -> first host-visible flag latency (one launch per step)
//   C  doorbell round trip with a resident kernel: host writes seq, the kernel (polling the mapped word) echoes it back
//      (optionally reading a 128-byte action slice per CTA and fencing system-wide before the echo)
//   D  cost of __threadfence_system() after sysmem stores, in-kernel (%globaltimer)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pcie_probe tools/pcie_probe.cu && tools/pcie_probe
#include <cuda_runtime.h>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)
static double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

// A: world w patches `cells` cells of its frame (frame_bytes apart): per cell 2 rows x (1B + 4B + 1B) like span6, or 4 rows x 3 words
__global__ void scatter(uint8_t* host, int n, uint32_t frame_bytes, int cells, int mode, uint32_t salt) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    uint8_t* f = host + (size_t)w * frame_bytes;
    uint32_t h = (uint32_t)w * 2654435761u + salt;
    for (int c = 0; c < cells; c++) {
        h = h * 1664525u + 1013904223u;
        const uint32_t cell = (h >> 8) % 441u, r = cell / 21u, col = cell % 21u;
        uint8_t* p = f + (size_t)(4 * r) * 252 + 12 * col;
        if (mode == 0) {                                          // two 6-byte spans at offset 3 (rows 1, 2)
            for (int y = 1; y < 3; y++) { uint8_t* q = p + y * 252 + 3; q[0] = 1; *reinterpret_cast<uint32_t*>(q + 1) = h; q[5] = 2; }
        } else if (mode == 1) {                                   // four full 12-byte rows as 3 words
            for (int y = 0; y < 4; y++) { uint32_t* q = reinterpret_cast<uint32_t*>(p + y * 252); q[0] = h; q[1] = h; q[2] = h; }
        } else {                                                  // two full 12-byte rows (rows 1, 2) as 3 words
            for (int y = 1; y < 3; y++) { uint32_t* q = reinterpret_cast<uint32_t*>(p + y * 252); q[0] = h; q[1] = h; q[2] = h; }
        }
    }
}
// full frames written by a CTA with 16-byte stores (a re-seeded world rendered straight into host memory)
__global__ void frames16(uint4* host, int nframes, uint32_t frame_bytes) {
    for (int f = blockIdx.x; f < nframes; f += gridDim.x) {
        uint4* dst = host + (size_t)f * (frame_bytes / 16);
        for (uint32_t i = threadIdx.x; i < frame_bytes / 16; i += blockDim.x) dst[i] = make_uint4(i, f, 3, 4);
    }
}
// B
__global__ void flag_kernel(volatile uint32_t* flag, uint32_t v) { if (threadIdx.x == 0 && blockIdx.x == 0) { *flag = v; } }
// B2: n worlds, each thread writes a 16-byte record then (per CTA) fence + flag
__global__ void records_kernel(uint4* rec, volatile uint32_t* flags, int n, uint32_t seq) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w < n) rec[w] = make_uint4(w, seq, seq, seq);
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) flags[blockIdx.x] = seq;
}
// C: resident kernel.  CTA 0 thread 0 polls the doorbell; everybody else polls a device word.
__global__ void doorbell_kernel(volatile uint32_t* bell, volatile uint32_t* echo, const uint8_t* actions, uint8_t* sink,
                                uint32_t* go, int rounds, int read_actions, int fence, unsigned long long* dev_ns) {
    __shared__ uint32_t s_sum;
    for (int k = 1; k <= rounds; k++) {
        if (blockIdx.x == 0 && threadIdx.x == 0) {
            const unsigned long long t0 = gtimer();
            while (*bell != (uint32_t)k) { if (gtimer() - t0 > 2000000000ull) { *echo = 0xDEADu; return; } }
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(go), "r"(k) : "memory");
        }
        if (threadIdx.x == 0) {
            uint32_t v;
            const unsigned long long t0 = gtimer();
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(go) : "memory"); } while (v < (uint32_t)k && gtimer() - t0 < 2000000000ull);
            s_sum = 0;
        }
        __syncthreads();
        const unsigned long long t1 = gtimer();
        uint32_t a = 0;
        if (read_actions) a = *(reinterpret_cast<const volatile uint8_t*>(actions) + blockIdx.x * blockDim.x + threadIdx.x);
        sink[blockIdx.x * blockDim.x + threadIdx.x] = (uint8_t)(a + k);    // a host-mapped byte per thread (reward/done stand-in)
        if (fence) __threadfence_system();
        __syncthreads();
        if (threadIdx.x == 0) {
            echo[blockIdx.x] = (uint32_t)k;
            if (blockIdx.x == 0 && dev_ns) dev_ns[k] = gtimer() - t1;
        }
    }
}
// D
__global__ void fence_cost(uint32_t* host, unsigned long long* out, int reps) {
    unsigned long long acc = 0;
    for (int i = 0; i < reps; i++) {
        host[threadIdx.x + 32 * i] = i;
        const unsigned long long t0 = gtimer();
        __threadfence_system();
        acc += gtimer() - t0;
    }
    if (threadIdx.x == 0) out[0] = acc / reps;
}

int main() {
    CK(cudaSetDevice(0));
    const uint32_t FB = 21168;
    const int NMAX = 131072;
    uint8_t* h_frames;
    CK(cudaHostAlloc(&h_frames, (size_t)NMAX * FB, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(h_frames, 0, (size_t)NMAX * FB);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    cudaStream_t s;
    CK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    printf("== A: scattered stores into mapped pinned host memory (kernel time incl. flush, CUDA events)\n");
    for (int n : {4096, 16384, 131072}) {
        for (int mode = 0; mode < 3; mode++) {
            for (int cells : {2}) {
                for (int it = 0; it < 3; it++) scatter<<<(n + 127) / 128, 128, 0, s>>>(h_frames, n, FB, cells, mode, it);
                CK(cudaStreamSynchronize(s));
                const int reps = 20;
                CK(cudaEventRecord(e0, s));
                for (int it = 0; it < reps; it++) scatter<<<(n + 127) / 128, 128, 0, s>>>(h_frames, n, FB, cells, mode, 100 + it);
                CK(cudaEventRecord(e1, s));
                CK(cudaStreamSynchronize(s));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                const int stores = mode == 0 ? 6 * cells : (mode == 1 ? 12 * cells : 6 * cells);
                printf("  n=%6d mode=%d (%s) cells=%d: %8.2f us per launch, %6.1f M stores/s\n", n, mode,
                       mode == 0 ? "2 rows x 1+4+1 B" : (mode == 1 ? "4 rows x 3 words " : "2 rows x 3 words "), cells, ms * 1e3 / reps,
                       (double)n * stores * reps / (ms * 1e-3) / 1e6);
            }
        }
    }
    for (int nf : {14, 28, 437}) {
        frames16<<<nf < 148 ? nf : 148, 256, 0, s>>>((uint4*)h_frames, nf, FB);
        CK(cudaStreamSynchronize(s));
        CK(cudaEventRecord(e0, s));
        for (int it = 0; it < 10; it++) frames16<<<nf < 148 ? nf : 148, 256, 0, s>>>((uint4*)h_frames, nf, FB);
        CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("  %d full frames (16-byte stores): %8.2f us per launch, %6.2f GB/s\n", nf, ms * 1e2, (double)nf * FB * 10 / (ms * 1e-3) / 1e9);
    }

    printf("== B: launch -> host-visible\n");
    uint32_t* h_flags;
    CK(cudaHostAlloc(&h_flags, 4096 * 4, cudaHostAllocMapped));
    memset(h_flags, 0, 4096 * 4);
    uint4* h_rec;
    CK(cudaHostAlloc(&h_rec, (size_t)NMAX * 16, cudaHostAllocMapped));
    {
        double acc = 0, accl = 0;
        const int reps = 2000;
        for (int k = 1; k <= reps + 100; k++) {
            const double t0 = now_us();
            flag_kernel<<<1, 32, 0, s>>>(h_flags, k);
            const double t1 = now_us();
            while (*(volatile uint32_t*)h_flags != (uint32_t)k) { }
            const double t2 = now_us();
            if (k > 100) { acc += t2 - t0; accl += t1 - t0; }
        }
        printf("  1-thread flag kernel: launch call %.2f us, launch -> flag visible %.2f us (total)\n", accl / reps, acc / reps);
        for (int n : {4096, 131072}) {
            const int ctas = (n + 127) / 128;
            acc = 0;
            for (int k = 1; k <= reps / 4 + 20; k++) {
                const double t0 = now_us();
                records_kernel<<<ctas, 128, 0, s>>>(h_rec, h_flags, n, k);
                for (int c = 0; c < ctas; c++) while (((volatile uint32_t*)h_flags)[c] != (uint32_t)k) { }
                if (k > 20) acc += now_us() - t0;
            }
            printf("  %d records + per-CTA flags (%d CTAs): launch -> all flags seen %.2f us\n", n, ctas, acc / (reps / 4));
        }
    }

    printf("== C: doorbell round trip with a resident kernel\n");
    {
        uint32_t *h_bell, *h_echo, *d_go;
        uint8_t *h_act, *h_sink;
        unsigned long long* d_ns;
        CK(cudaHostAlloc(&h_bell, 64, cudaHostAllocMapped));
        CK(cudaHostAlloc(&h_echo, 4096 * 4, cudaHostAllocMapped));
        CK(cudaHostAlloc(&h_act, NMAX, cudaHostAllocMapped));
        CK(cudaHostAlloc(&h_sink, NMAX, cudaHostAllocMapped));
        CK(cudaMalloc(&d_go, 4));
        const int rounds = 2000;
        CK(cudaMalloc(&d_ns, (rounds + 1) * 8));
        for (int ctas : {1, 32, 128}) {
            for (int variant = 0; variant < 3; variant++) {
                const int read_actions = variant >= 1, fence = variant >= 2;
                *h_bell = 0; memset(h_echo, 0, 4096 * 4);
                CK(cudaMemset(d_go, 0, 4));
                CK(cudaDeviceSynchronize());
                doorbell_kernel<<<ctas, 128, 0, s>>>(h_bell, h_echo, h_act, h_sink, d_go, rounds, read_actions, fence, d_ns);
                double acc = 0;
                bool dead = false;
                for (int k = 1; k <= rounds && !dead; k++) {
                    memset(h_act, k & 7, ctas * 128);
                    const double t0 = now_us();
                    __atomic_store_n(h_bell, (uint32_t)k, __ATOMIC_RELEASE);
                    for (int c = 0; c < ctas; c++) {
                        while (((volatile uint32_t*)h_echo)[c] != (uint32_t)k) { if (now_us() - t0 > 3e6) { dead = true; break; } }
                        if (dead) break;
                    }
                    if (k > 100) acc += now_us() - t0;
                }
                CK(cudaStreamSynchronize(s));
                unsigned long long ns[8];
                CK(cudaMemcpy(ns, d_ns + 1000, sizeof(ns), cudaMemcpyDeviceToHost));
                printf("  %3d CTAs, %s: round trip %.2f us%s  (device: GO seen -> echo issued %llu ns)\n", ctas,
                       variant == 0 ? "echo only          " : (variant == 1 ? "+ action read      " : "+ action read+fence"), acc / (rounds - 100),
                       dead ? " [TIMED OUT]" : "", ns[0]);
            }
        }
    }
    printf("== D: __threadfence_system after a sysmem store (one warp)\n");
    {
        unsigned long long* d_out; CK(cudaMalloc(&d_out, 8));
        fence_cost<<<1, 32, 0, s>>>((uint32_t*)h_frames, d_out, 100);
        CK(cudaStreamSynchronize(s));
        unsigned long long v; CK(cudaMemcpy(&v, d_out, 8, cudaMemcpyDeviceToHost));
        printf("  fence.sc.sys after a host store: %llu ns\n", v);
    }
    return 0;
}
