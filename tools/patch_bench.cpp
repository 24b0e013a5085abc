// tools/patch_bench.cpp -- CPU-only microbenchmark of the delta-transport patch loop (cw_host.cu): random-walk agents, synthetic pre-digested
// records, N worlds x 21x21.  g++ -O3 -o /tmp/pb tools/patch_bench.cpp -lpthread && /tmp/pb N threads mode(0 patch, 1 +prefetch, 2 prefetch only) hugepages
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sys/mman.h>
#include <thread>
#include <vector>
static inline double now_us() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
struct Rec { uint32_t x, y, z, w; };
static uint8_t r6[16][8], r12[16][16];
static inline void store6(uint8_t* p, int c) { memcpy(p, r6[c], 4); memcpy(p + 4, r6[c] + 4, 2); }
static inline void store12(uint8_t* p, int c) { memcpy(p, r12[c], 8); memcpy(p + 8, r12[c] + 8, 4); }
int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 4096, T = argc > 2 ? atoi(argv[2]) : 1, mode = argc > 3 ? atoi(argv[3]) : 0, huge = argc > 4 ? atoi(argv[4]) : 0;
    const int W = 21, steps = 400;
    const size_t fb = 48 * W * W, rowb = 12 * W;
    size_t bytes = (size_t)N * fb;
    uint8_t* frames = (uint8_t*)mmap(nullptr, bytes + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    frames = (uint8_t*)(((uintptr_t)frames + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
    if (huge) madvise(frames, bytes, MADV_HUGEPAGE);
    memset(frames, 0, bytes);
    for (int c = 0; c < 16; c++) { for (int k = 0; k < 12; k++) r12[c][k] = c * 16 + k; for (int k = 0; k < 6; k++) r6[c][k] = c * 16 + k; }
    std::vector<uint32_t> pos(N);
    std::vector<Rec> recs((size_t)steps * N);
    std::vector<int32_t> rew(N); std::vector<uint8_t> dn(N);
    uint32_t h = 12345;
    for (int w = 0; w < N; w++) { h = h * 1664525u + 1013904223u; pos[w] = ((h >> 8) % 21) | (((h >> 16) % 21) << 8); }
    for (int s = 0; s < steps; s++) for (int w = 0; w < N; w++) {
        h = h * 1664525u + 1013904223u;
        int a = (h >> 10) % 6, r = pos[w] & 0xFF, c = pos[w] >> 8, nr = r, nc = c;
        if (a == 0) nr = r > 0 ? r - 1 : 0; else if (a == 2) nr = r < 20 ? r + 1 : 20; else if (a == 1) nc = c < 20 ? c + 1 : 20; else if (a == 3) nc = c > 0 ? c - 1 : 0;
        Rec& q = recs[(size_t)s * N + w];
        q.x = nr | (nc << 8); q.y = 0; q.z = r | (c << 6) | (((h >> 20) & 7) << 12) | (((h >> 23) & 7) << 16) | ((a >= 4 && ((h >> 26) & 3) == 0) ? 1u << 20 : 0); q.w = -1;
        pos[w] = nr | (nc << 8);
    }
    auto job = [&](int tid, int s) {
        const int64_t lo = (int64_t)N * tid / T, hi = (int64_t)N * (tid + 1) / T;
        const Rec* rr = recs.data() + (size_t)s * N;
        for (int64_t w = lo; w < hi; w++) {
            if (mode >= 1 && w + 8 < hi) {
                const Rec& p = rr[w + 8];
                uint8_t* f = frames + (w + 8) * fb;
                const uint8_t* po = f + (size_t)(4 * (p.z & 63u) + 1) * rowb + 12 * ((p.z >> 6) & 63u) + 3;
                const uint8_t* pn = f + (size_t)(4 * (p.x & 0xFFu) + 1) * rowb + 12 * ((p.x >> 8) & 0xFFu) + 3;
                __builtin_prefetch(po, 1); __builtin_prefetch(po + rowb, 1); __builtin_prefetch(pn, 1); __builtin_prefetch(pn + rowb, 1);
            }
            const Rec r = rr[w];
            rew[w] = r.w; dn[w] = (r.z >> 24) & 1;
            if (mode == 2) continue;
            uint8_t* frame = frames + w * fb;
            const uint32_t z = r.z;
            const int orow = z & 63, ocol = (z >> 6) & 63, ocode = (z >> 12) & 15, ncode = (z >> 16) & 15;
            const int nrow = r.x & 0xFF, ncol = (r.x >> 8) & 0xFF, hold = (r.x >> 16) & 0xFF;
            const bool objchg = (z >> 20) & 1, moved = (orow != nrow) | (ocol != ncol);
            if (!moved && !objchg) continue;
            uint8_t* pn = frame + (size_t)(4 * nrow) * rowb + 12 * ncol;
            if (moved) { uint8_t* po = frame + (size_t)(4 * orow + 1) * rowb + 12 * ocol + 3; store6(po, ocode); store6(po + rowb, ocode); }
            if (objchg) for (int y = 0; y < 4; y++) store12(pn + y * rowb, ncode);
            store6(pn + rowb + 3, 9); store6(pn + 2 * rowb + 3, hold ? hold : 9);
        }
    };
    for (int rep = 0; rep < 2; rep++) {
        double t0 = now_us();
        for (int s = 0; s < steps; s++) {
            std::vector<std::thread> th;
            if (T == 1) job(0, s);
            else { for (int t = 0; t < T; t++) th.emplace_back(job, t, s); for (auto& x : th) x.join(); }
        }
        double dt = now_us() - t0;
        printf("N=%d T=%d mode=%d huge=%d: %.2f us/step, %.1f ns/world/thread\n", N, T, mode, huge, dt / steps, dt * 1e3 / steps / N * T);
    }
}
