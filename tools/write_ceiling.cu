// tools/write_ceiling.cu -- how fast can this B200 WRITE?  The roofline denominator in bench.py is the measured COPY
// bandwidth (MEASURED_PEAKS.json: read + write mix); the env kernel is a ~98 % write stream.  This probe issues nothing but
// TMA bulk stores of a constant shared-memory frame (no compute, no loads) over the same footprint as config 2 / config 4
// and reports GB/s, i.e. the ceiling any frame writer can reach.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/write_ceiling tools/write_ceiling.cu && tools/write_ceiling
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>

__global__ void __launch_bounds__(128, 4) blast(uint8_t* dst, int64_t nframes, uint32_t frame_bytes, int depth) {
    extern __shared__ __align__(128) uint8_t sm[];
    for (uint32_t i = threadIdx.x; i < frame_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(sm)[i] = 0x01020304u * (i + 1);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        uint64_t pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        int pending = 0;
        for (int64_t f = blockIdx.x; f < nframes; f += gridDim.x) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(dst + f * frame_bytes),
                         "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(frame_bytes), "l"(pol) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++pending >= depth) { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); pending = 1; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

int main() {
    const uint32_t frame_bytes = 21168;
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaFuncSetAttribute(blast, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024);
    for (int64_t worlds : {4096ll, 131072ll}) {
        const int ring = worlds == 4096 ? 4 : 2;                  // same footprint as bench.py: larger than the 126 MB L2
        uint8_t* buf;
        if (cudaMalloc(&buf, (size_t)ring * worlds * frame_bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
        for (int per_sm : {2, 4, 8}) for (int depth : {2, 4}) {
            const int grid = sms * per_sm, iters = worlds == 4096 ? 2000 : 60;
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int i = 0; i < 20; i++) blast<<<grid, 128, frame_bytes>>>(buf + (size_t)(i % ring) * worlds * frame_bytes, worlds, frame_bytes, depth);
            cudaEventRecord(e0);
            for (int i = 0; i < iters; i++) blast<<<grid, 128, frame_bytes>>>(buf + (size_t)(i % ring) * worlds * frame_bytes, worlds, frame_bytes, depth);
            cudaEventRecord(e1);
            if (cudaEventSynchronize(e1) != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            printf("pure TMA store stream: %7lld frames x %u B per launch, %d CTAs/SM, depth %d: %7.2f us per launch, %7.1f GB/s\n",
                   (long long)worlds, frame_bytes, per_sm, depth, ms * 1e3 / iters, (double)worlds * frame_bytes * iters / (ms * 1e-3) / 1e9);
        }
        cudaFree(buf);
    }
    return 0;
}
