// cw_host.cu -- host-buffer API: the batched env behind an opaque handle (cw_host_*, see include/cw_b200.h).
//
// Every argument is a HOST pointer.  One call = one reference-style `env.step(actions)` for N worlds:
// actions go host->device, the fused step+reset+render launch runs in slices, reward/done (and the frames when
// an obs buffer is passed) come back device->host.  Slices alternate between two streams so the D2H copy of
// slice j overlaps the kernel of slice j+1; PCIe, not the kernel, bounds this path when frames are returned.
#include <cuda_runtime.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "cw_b200.h"

// ---- a small persistent worker pool for the host-side frame patching of the delta transport -----------------
namespace {
inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#else
    std::this_thread::yield();
#endif
}
class WorkerPool {
public:
    explicit WorkerPool(int nthreads) : n_(nthreads < 1 ? 1 : nthreads) {
        for (int i = 1; i < n_; i++) th_.emplace_back([this, i] { loop(i); });
    }
    ~WorkerPool() {
        { std::lock_guard<std::mutex> lk(m_); stop_ = true; gen_.fetch_add(1); }
        cv_.notify_all();
        for (auto& t : th_) t.join();
    }
    // start job(tid, nthreads) on the worker threads (tid 1..n-1); the caller then does whatever it wants (e.g. launch
    // the kernel whose output the workers are already polling for), runs tid 0 itself and joins
    void start(const std::function<void(int, int)>& job) {
        job_ = &job;
        remaining_.store(n_ - 1, std::memory_order_release);
        { std::lock_guard<std::mutex> lk(m_); gen_.fetch_add(1, std::memory_order_release); }
        cv_.notify_all();
    }
    bool done() const { return remaining_.load(std::memory_order_acquire) == 0; }
    int size() const { return n_; }
private:
    void loop(int tid) {
        uint64_t seen = 0;
        for (;;) {
            // stay hot while the caller is stepping back to back (a step is tens of microseconds); block after ~0.5 ms idle
            const auto t0 = std::chrono::steady_clock::now();
            for (int spin = 0; gen_.load(std::memory_order_acquire) == seen; spin++) {
                if ((spin & 255) == 255 && std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(500)) break;
            }
            if (gen_.load(std::memory_order_acquire) == seen) {
                std::unique_lock<std::mutex> lk(m_);
                cv_.wait(lk, [&] { return gen_.load(std::memory_order_acquire) != seen; });
            }
            seen = gen_.load(std::memory_order_acquire);
            if (stop_) return;
            (*job_)(tid, n_);
            remaining_.fetch_sub(1, std::memory_order_acq_rel);
        }
    }
    int n_;
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable cv_;
    std::atomic<uint64_t> gen_{0};
    std::atomic<int> remaining_{0};
    const std::function<void(int, int)>* job_ = nullptr;
    bool stop_ = false;
};

// COLORS_N (ray.py:28-30)
constexpr uint8_t kLut[9][3] = {{0, 0, 0},   {110, 69, 39},  {255, 105, 180}, {100, 100, 200}, {100, 100, 100},
                            {0, 128, 0}, {205, 133, 63}, {197, 91, 97},   {240, 230, 140}};

// the same colours repeated over 4 pixels (a full cell row) and over 2 pixels (an overlay row)
struct LutRows {
    uint8_t r12[9][12], r6[9][6];
    constexpr LutRows() : r12(), r6() {
        for (int c = 0; c < 9; c++) {
            for (int k = 0; k < 12; k++) r12[c][k] = kLut[c][k % 3];
            for (int k = 0; k < 6; k++) r6[c][k] = kLut[c][k % 3];
        }
    }
};
constexpr LutRows kRows;
constexpr auto& kLut12 = kRows.r12;
constexpr auto& kLut6 = kRows.r6;
const uint8_t kWhite6[6] = {255, 255, 255, 255, 255, 255};

// one cell of a host frame: 4x4 pixels of the object's colour, agent overlay on top (ray.py:550-557)
inline void patch_cell(uint8_t* frame, int W, int cell, int code, bool agent_here, int hold) {
    const int r = cell / W, c = cell - r * W;
    const size_t rowb = (size_t)12 * W;
    uint8_t px[12];
    for (int k = 0; k < 4; k++) memcpy(px + 3 * k, kLut[code], 3);
    uint8_t* p = frame + (size_t)(4 * r) * rowb + 12 * c;
    for (int y = 0; y < 4; y++) memcpy(p + y * rowb, px, 12);
    if (agent_here) {
        memset(p + rowb + 3, 255, 6);                                     // ray.py:555
        if (hold) { memcpy(p + 2 * rowb + 3, kLut[hold], 3); memcpy(p + 2 * rowb + 6, kLut[hold], 3); }   // ray.py:556-557
        else memset(p + 2 * rowb + 3, 255, 6);
    }
}
// a whole host frame from a grid tile (ray.py:442-486): used for re-seeded worlds only
inline void render_frame(uint8_t* frame, int H, int W, const uint8_t* g, uint32_t agent) {
    const size_t rowb = (size_t)12 * W;
    for (int br = 0; br < H; br++) {
        uint8_t* row = frame + (size_t)(4 * br) * rowb;
        for (int bc = 0; bc < W; bc++)
            for (int k = 0; k < 4; k++) memcpy(row + 12 * bc + 3 * k, kLut[g[br * W + bc]], 3);
        for (int k = 1; k < 4; k++) memcpy(row + k * rowb, row, rowb);
    }
    const int ar = agent & 0xFF, ac = (agent >> 8) & 0xFF, h = (agent >> 16) & 0xFF;
    patch_cell(frame, W, ar * W + ac, g[ar * W + ac], true, h);
}
}  // namespace

struct CwHostEnv {
    uint32_t magic;
    CwConfig cfg;
    CwState st;
    int device, flags;
    size_t frame_bytes;
    uint8_t *d_actions, *d_done, *d_obs, *d_goal_obs;
    int32_t* d_reward;
    int64_t* d_stats;
    uint8_t *h_actions, *h_done;      // pinned staging for the small vectors
    int32_t* h_reward;
    int64_t* h_stats;
    uint8_t* h_frames;                // pinned staging for frames when the caller's buffer is pageable
    size_t h_frames_bytes;
    cudaStream_t streams[2];
    int64_t slice;                    // worlds per slice
    // delta transport (CW_F_DELTA_TRANSPORT)
    uint4* h_delta;                   // pinned + mapped: per-world delta records written by the kernel
    uint32_t* h_fresh;                // pinned + mapped: sparse records of re-seeded worlds
    uint8_t* m_grid;                  // host mirror of the grids (the state the caller's frames show)
    uint32_t* m_agent;                // host mirror of the agent words
    uint8_t* mirror_obs;              // caller buffers the mirror currently describes
    uint8_t* mirror_goal;
    WorkerPool* pool;
    uint32_t seq;                     // sequence tag of the last delta step (1..63)
    const uint8_t* pinned_actions;    // last caller action buffer found to be page-locked
    bool nopatch;                     // CW_HOST_NOPATCH=1 (diagnostics only): consume the records, skip the frame patching
    bool trace;                       // CW_HOST_TRACE=1 (diagnostics only): phase times of the delta step, printed at destroy
    double tr_launch, tr_first, tr_total;   // accumulated microseconds: launch call, launch -> first record seen, whole call
    uint64_t tr_steps;
};

#define CW_HOST_MAGIC 0x43574845u
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

// worker threads of the delta transport: the cores this process may run on, shared fairly between the ranks of the node
// (LOCAL_WORLD_SIZE is set by torchrun), at most 16, at least 1; CW_HOST_THREADS overrides.
static int host_threads(int64_t n) {
    if (const char* s = getenv("CW_HOST_THREADS")) { const int v = atoi(s); if (v > 0) return v > 64 ? 64 : v; }
    int cores = 0;
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cores = CPU_COUNT(&set);
    if (cores <= 0) cores = (int)std::thread::hardware_concurrency();
    if (cores <= 0) cores = 4;
    int ranks = 1;
    if (const char* s = getenv("LOCAL_WORLD_SIZE")) { const int v = atoi(s); if (v > 0) ranks = v; }
    int nt = cores / ranks;
    if (nt > 16) nt = 16;
    if (nt > (int)(n / 128 + 1)) nt = (int)(n / 128 + 1);
    return nt < 1 ? 1 : nt;
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

static CwState slice_state(const CwHostEnv* e, int64_t off, int64_t cnt) {
    CwState s = e->st;
    s.grid += off * e->cfg.cell_stride; s.init_grid += off * e->cfg.cell_stride;
    s.agent += off; s.goal += off; s.t += off; s.episode += off;
    s.n = cnt; s.env_id_base += (uint64_t)off;
    return s;
}

static int ensure_frame_staging(CwHostEnv* e) {
    const size_t need = (size_t)e->st.n * e->frame_bytes;
    if (e->h_frames_bytes >= need) return 0;
    if (e->h_frames) cudaFreeHost(e->h_frames);
    e->h_frames = nullptr; e->h_frames_bytes = 0;
    CK(cudaMallocHost(&e->h_frames, need));
    e->h_frames_bytes = need;
    return 0;
}

extern "C" {

int cw_host_create(const CwConfig* cfg, int64_t n, int device, uint64_t seed, uint64_t env_id_base, int flags, CwHostEnv** out) {
    if (!cfg || !out) return CW_E_NULLPTR;
    if (n < 1) return CW_E_BADCONFIG;
    if (flags & ~(CW_F_AUTO_RESET | CW_F_DELTA_TRANSPORT)) return CW_E_BADFLAGS;
    CK(cudaSetDevice(device));
    CwHostEnv* e = new (std::nothrow) CwHostEnv();
    if (!e) return (int)cudaErrorMemoryAllocation;
    memset(e, 0, sizeof(*e));
    e->magic = CW_HOST_MAGIC; e->cfg = *cfg; e->device = device; e->flags = flags;
    e->frame_bytes = (size_t)48 * cfg->H * cfg->W;
    e->st.n = n; e->st.seed = seed; e->st.env_id_base = env_id_base;
    const size_t gb = (size_t)n * cfg->cell_stride;
    int rc = 0;
#define TRY(x) do { if (!rc) { cudaError_t e_ = (x); if (e_ != cudaSuccess) rc = (int)e_; } } while (0)
    TRY(cudaMalloc(&e->st.grid, gb)); TRY(cudaMalloc(&e->st.init_grid, gb));
    TRY(cudaMalloc(&e->st.agent, n * 4)); TRY(cudaMalloc(&e->st.goal, n * 4));
    TRY(cudaMalloc(&e->st.t, n * 4)); TRY(cudaMalloc(&e->st.episode, n * 4));
    TRY(cudaMalloc(&e->d_actions, n));
    TRY(cudaMalloc(&e->d_reward, n * 5));                       // [reward int32 x n][done uint8 x n], one block
    if (!rc) e->d_done = reinterpret_cast<uint8_t*>(e->d_reward) + n * 4;
    TRY(cudaMalloc(&e->d_obs, (size_t)n * e->frame_bytes)); TRY(cudaMalloc(&e->d_goal_obs, (size_t)n * e->frame_bytes));
    TRY(cudaMalloc(&e->d_stats, CW_STATS_REPLICAS * CW_STATS_LEN * 8));
    TRY(cudaMallocHost(&e->h_actions, n)); TRY(cudaMallocHost(&e->h_reward, n * 5));
    if (!rc) e->h_done = reinterpret_cast<uint8_t*>(e->h_reward) + n * 4;
    TRY(cudaMallocHost(&e->h_stats, CW_STATS_REPLICAS * CW_STATS_LEN * 8));
    if (flags & CW_F_DELTA_TRANSPORT) {
        TRY(cudaMallocHost(&e->h_delta, n * sizeof(uint4)));
        TRY(cudaMallocHost(&e->h_fresh, n * CW_FRESH_WORDS * sizeof(uint32_t)));
        if (!rc) {
            e->m_grid = (uint8_t*)malloc(gb);
            e->m_agent = (uint32_t*)malloc(n * 4);
            if (!e->m_grid || !e->m_agent) rc = (int)cudaErrorMemoryAllocation;
            e->pool = new (std::nothrow) WorkerPool(host_threads(n));
            if (!rc && e->h_delta) memset(e->h_delta, 0, n * sizeof(uint4));   // tag 0 = never written
            if (const char* np = getenv("CW_HOST_NOPATCH")) e->nopatch = *np == '1';
            if (const char* tr = getenv("CW_HOST_TRACE")) e->trace = *tr == '1';
        }
    }
    TRY(cudaStreamCreateWithFlags(&e->streams[0], cudaStreamNonBlocking));
    TRY(cudaStreamCreateWithFlags(&e->streams[1], cudaStreamNonBlocking));
    if (!rc) {
        TRY(cudaMemset(e->st.grid, 0, gb)); TRY(cudaMemset(e->st.init_grid, 0, gb));
        TRY(cudaMemset(e->st.agent, 0, n * 4)); TRY(cudaMemset(e->st.goal, 0, n * 4));
        TRY(cudaMemset(e->st.t, 0, n * 4)); TRY(cudaMemset(e->st.episode, 0, n * 4));
        TRY(cudaMemset(e->d_stats, 0, CW_STATS_REPLICAS * CW_STATS_LEN * 8));
        TRY(cudaDeviceSynchronize());
    }
#undef TRY
    // slices: enough to pipeline copies against kernels, large enough to fill the GPU
    int64_t slice = (n + 7) / 8;
    if (slice < 512) slice = n < 512 ? n : 512;
    e->slice = slice;
    if (rc) { cw_host_destroy(e); return rc; }
    *out = e;
    return 0;
}

int cw_host_reset(CwHostEnv* e, uint8_t* obs_host, uint8_t* goal_obs_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    CK(cudaSetDevice(e->device));
    cudaStream_t s = e->streams[0];
    int rc = cw_reset(&e->cfg, &e->st, nullptr, e->d_obs, e->d_goal_obs, nullptr, s);
    if (rc) return rc;
    const size_t total = (size_t)e->st.n * e->frame_bytes;
    for (int which = 0; which < 2; which++) {
        uint8_t* dst = which ? goal_obs_host : obs_host;
        const uint8_t* src = which ? e->d_goal_obs : e->d_obs;
        if (!dst) continue;
        if (is_pinned(dst)) { CK(cudaMemcpyAsync(dst, src, total, cudaMemcpyDeviceToHost, s)); CK(cudaStreamSynchronize(s)); }
        else {
            rc = ensure_frame_staging(e); if (rc) return rc;
            CK(cudaMemcpyAsync(e->h_frames, src, total, cudaMemcpyDeviceToHost, s));
            CK(cudaStreamSynchronize(s));
            memcpy(dst, e->h_frames, total);
        }
    }
    CK(cudaStreamSynchronize(s));
    if (e->flags & CW_F_DELTA_TRANSPORT) {                       // (re)build the host mirror
        CK(cudaMemcpy(e->m_grid, e->st.grid, (size_t)e->st.n * e->cfg.cell_stride, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(e->m_agent, e->st.agent, (size_t)e->st.n * 4, cudaMemcpyDeviceToHost));
        e->mirror_obs = obs_host;
        e->mirror_goal = goal_obs_host;
    }
    return 0;
}

// delta transport: one launch writing 16-byte records into mapped pinned memory; the pool patches the frames WHILE the
// kernel runs.  There is no stream synchronisation on this path: every record carries the step's 6-bit sequence tag
// (one 16-byte store, preceded system-wide by the sparse record of a re-seeded world), the workers are started before
// the launch and poll the records of their slice; when every record of the step has been consumed the step is complete.
static int host_step_delta(CwHostEnv* e, const uint8_t* act_src, int32_t* reward_host, uint8_t* done_host, uint8_t* obs_host) {
    cudaStream_t s = e->streams[0];
    const int64_t n = e->st.n;
    const int H = e->cfg.H, W = e->cfg.W, cs = e->cfg.cell_stride;
    if (obs_host != e->mirror_obs) {                              // unknown buffer: one full refresh, then deltas
        int rc = cw_render(&e->cfg, e->st.grid, e->st.agent, e->d_obs, n, s);
        if (rc) return rc;
        CK(cudaMemcpyAsync(obs_host, e->d_obs, (size_t)n * e->frame_bytes, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(e->m_grid, e->st.grid, (size_t)n * cs, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(e->m_agent, e->st.agent, (size_t)n * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        e->mirror_obs = obs_host;
    }
    e->seq = e->seq >= 63 ? 1 : e->seq + 1;                       // 1..63; 0 is the never-written state of the buffer
    const uint32_t seq = e->seq;
    const size_t fb = e->frame_bytes;
    uint8_t* goal = e->mirror_goal;
    std::atomic<int> failed{0};
    const bool nopatch = e->nopatch;
    const std::function<void(int, int)> job = [&, n, H, W, cs, seq, fb, goal, nopatch](int tid, int nt) {
        const int64_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
        uint8_t tmp[CW_MAX_SIDE * CW_MAX_SIDE];
        const size_t rowb = (size_t)12 * W;
        // Records do not arrive in order: a re-seeded world's record follows ~6 us after its neighbours' (its warp runs the
        // Philox reset + imagine_obs first).  Worlds whose record is not there yet are deferred and revisited after the
        // rest of the slice, so one late record does not stall the patching behind it.
        constexpr int kMaxDeferred = 128;
        int64_t deferred[kMaxDeferred];
        int ndef = 0;
        const int64_t total = hi - lo;
        auto ready = [&](int64_t w) { return (reinterpret_cast<const volatile uint32_t*>(e->h_delta + w)[2] >> 26) == seq; };
        auto wait_for = [&](int64_t w) {                          // poll (the GPU's write invalidates the line); false: launch failed
            uint64_t spins = 0;
            while (!ready(w)) {
                if (failed.load(std::memory_order_relaxed)) return false;
                if ((++spins & 0xFFFFF) == 0 && tid == 0 && cudaStreamQuery(s) != cudaErrorNotReady) {
                    // the launch has finished (or failed): every record is in host memory now, or never will be
                    if (!ready(w)) { failed.store(1); return false; }
                }
                cpu_relax();
            }
            return true;
        };
        if (total > 0 && !wait_for(lo)) return;                   // the step's records start to land
        for (int64_t it = 0; it < total + ndef; it++) {
            const bool second = it >= total;
            const int64_t w = second ? deferred[it - total] : lo + it;
            if (!ready(w)) {
                if (!second && ndef < kMaxDeferred) { deferred[ndef++] = w; continue; }
                if (!wait_for(w)) return;
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            const uint4 r = e->h_delta[w];
            const uint32_t flags = r.z >> 24;
            reward_host[w] = (int32_t)r.w;
            done_host[w] = (uint8_t)(flags & 1u);
            if (nopatch) continue;
            uint8_t* g = e->m_grid + w * cs;
            uint8_t* frame = obs_host + w * fb;
            if (flags & 2u) {                                     // re-seeded: rebuild tile + frame (+ goal frame)
                const uint32_t* fr = e->h_fresh + w * CW_FRESH_WORDS;
                memset(g, 0, cs);
                for (int k = 0; k < 8; k++) if (fr[k] >> 16) g[fr[k] & 0xFFFFu] = (uint8_t)(fr[k] >> 16);
                e->m_agent[w] = r.x;
                render_frame(frame, H, W, g, r.x);
                if (goal) {
                    memset(tmp, 0, (size_t)H * W);
                    for (int k = 8; k < 16; k++) if (fr[k] >> 16) tmp[fr[k] & 0xFFFFu] = (uint8_t)(fr[k] >> 16);
                    render_frame(goal + w * fb, H, W, tmp, fr[16]);
                }
                continue;
            }
            // render_edit (ray.py:522-557) on the <= 2 cells a step can change.  A cell whose OBJECT is unchanged differs
            // only in the centred 2x2 overlay block, so it costs two 6-byte writes instead of four 12-byte rows.
            const uint32_t old = e->m_agent[w];
            const int wcell = (int)(r.z & 0xFFFFu);
            if (wcell != 0xFFFF) g[wcell] = (uint8_t)((r.z >> 16) & 0xFFu);
            else if (old == r.x) continue;                        // nothing visible changed
            e->m_agent[w] = r.x;
            const int orow = (int)(old & 0xFF), ocol = (int)((old >> 8) & 0xFF);
            const int nrow = (int)(r.x & 0xFF), ncol = (int)((r.x >> 8) & 0xFF), hold = (int)((r.x >> 16) & 0xFF);
            const int oc = orow * W + ocol, nc = nrow * W + ncol;
            uint8_t* po = frame + (size_t)(4 * orow) * rowb + 12 * ocol;
            uint8_t* pn = frame + (size_t)(4 * nrow) * rowb + 12 * ncol;
            if (oc != nc) {                                       // the agent left `oc`; its object did not change
                const uint8_t* col = kLut6[g[oc]];
                memcpy(po + rowb + 3, col, 6); memcpy(po + 2 * rowb + 3, col, 6);
            }
            if (wcell != 0xFFFF && wcell != nc) patch_cell(frame, W, wcell, g[wcell], false, 0);   // (not produced by step())
            if (wcell == nc) {                                    // the object under the agent changed: all four rows
                const uint8_t* col = kLut12[g[nc]];
                for (int y = 0; y < 4; y++) memcpy(pn + y * rowb, col, 12);
            }
            memset(pn + rowb + 3, 255, 6);                        // ray.py:555
            memcpy(pn + 2 * rowb + 3, hold ? kLut6[hold] : kWhite6, 6);   // ray.py:556-557
        }
    };
    const auto t_begin = std::chrono::steady_clock::now();
    e->pool->start(job);                                          // workers poll while the launch is on its way
    int rc = cw_step_delta(&e->cfg, &e->st, act_src, e->h_delta, e->h_fresh, e->d_stats, (e->flags & CW_F_AUTO_RESET) | CW_F_HOST_ACTIONS,
                           (int)seq, s);
    if (rc) failed.store(1);
    const auto t_launched = std::chrono::steady_clock::now();
    if (e->trace && !rc) {                                        // when does the first record of the caller's slice land?
        const volatile uint32_t* tag = &reinterpret_cast<const volatile uint32_t*>(e->h_delta)[2];
        while ((*tag >> 26) != seq && !failed.load(std::memory_order_relaxed)) cpu_relax();
        e->tr_first += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_launched).count();
    }
    job(0, e->pool->size());
    // watchdog: once the launch has left the stream every record is in host memory; workers still polling 100 ms later
    // will never be served (a failed launch) -- release them instead of hanging the caller
    for (uint64_t k = 1; !e->pool->done(); k++) {
        if ((k & 0x3FFF) == 0 && !rc && cudaStreamQuery(s) != cudaErrorNotReady) {
            const auto t0 = std::chrono::steady_clock::now();
            while (!e->pool->done() && std::chrono::steady_clock::now() - t0 < std::chrono::milliseconds(100)) cpu_relax();
            if (!e->pool->done()) failed.store(1);
        }
        cpu_relax();
    }
    if (e->trace) {
        e->tr_launch += std::chrono::duration<double, std::micro>(t_launched - t_begin).count();
        e->tr_total += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t_begin).count();
        e->tr_steps++;
    }
    if (rc) return rc;
    if (failed.load()) {
        cudaError_t ce = cudaStreamSynchronize(s);
        return (int)(ce != cudaSuccess ? ce : cudaErrorUnknown);
    }
    return 0;
}

int cw_host_step(CwHostEnv* e, const uint8_t* actions_host, int32_t* reward_host, uint8_t* done_host, uint8_t* obs_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (!actions_host || !reward_host || !done_host) return CW_E_NULLPTR;
    CK(cudaSetDevice(e->device));
    const int64_t n = e->st.n;
    const uint8_t* act_src = actions_host;
    if (actions_host != e->pinned_actions) {                     // the attribute query costs ~1 us: remember a pinned buffer
        if (is_pinned(actions_host)) e->pinned_actions = actions_host;
        else { memcpy(e->h_actions, actions_host, n); act_src = e->h_actions; }
    }
    if (obs_host && (e->flags & CW_F_DELTA_TRANSPORT)) return host_step_delta(e, act_src, reward_host, done_host, obs_host);
    const bool direct = obs_host && is_pinned(obs_host);
    uint8_t* frames_dst = obs_host;
    if (obs_host && !direct) { int rc = ensure_frame_staging(e); if (rc) return rc; frames_dst = e->h_frames; }
    if (!obs_host) {
        // Frames stay in HBM for a device-side consumer.  Zero-copy: the kernel reads the actions from, and writes
        // reward/done to, mapped pinned host memory (UVA), so a step is ONE launch + ONE stream sync -- no memcpy
        // launches on the critical path.
        cudaStream_t s = e->streams[0];
        int rc = cw_step_render(&e->cfg, &e->st, act_src /* pinned: the caller's own buffer or our staging copy */, e->h_reward, e->h_done, e->d_obs, e->d_goal_obs, nullptr,
                                e->d_stats, e->flags & CW_F_AUTO_RESET, s);
        if (rc) return rc;
        CK(cudaStreamSynchronize(s));
    } else {
        int k = 0;
        for (int64_t off = 0; off < n; off += e->slice, k ^= 1) {
            const int64_t cnt = (n - off) < e->slice ? (n - off) : e->slice;
            cudaStream_t s = e->streams[k];
            CK(cudaMemcpyAsync(e->d_actions + off, act_src + off, cnt, cudaMemcpyHostToDevice, s));
            CwState sl = slice_state(e, off, cnt);
            int rc = cw_step_render(&e->cfg, &sl, e->d_actions + off, e->d_reward + off, e->d_done + off,
                                    e->d_obs + (size_t)off * e->frame_bytes, e->d_goal_obs + (size_t)off * e->frame_bytes, nullptr,
                                    e->d_stats, e->flags & CW_F_AUTO_RESET, s);
            if (rc) return rc;
            CK(cudaMemcpyAsync(e->h_reward + off, e->d_reward + off, cnt * 4, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(e->h_done + off, e->d_done + off, cnt, cudaMemcpyDeviceToHost, s));
            CK(cudaMemcpyAsync(frames_dst + (size_t)off * e->frame_bytes, e->d_obs + (size_t)off * e->frame_bytes,
                               (size_t)cnt * e->frame_bytes, cudaMemcpyDeviceToHost, s));
        }
        CK(cudaStreamSynchronize(e->streams[0]));
        CK(cudaStreamSynchronize(e->streams[1]));
    }
    memcpy(reward_host, e->h_reward, n * 4);
    memcpy(done_host, e->h_done, n);
    if (obs_host && !direct) memcpy(obs_host, e->h_frames, (size_t)n * e->frame_bytes);
    return 0;
}

int cw_host_stats(CwHostEnv* e, int64_t* stats_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (!stats_host) return CW_E_NULLPTR;
    CK(cudaSetDevice(e->device));
    CK(cudaMemcpyAsync(e->h_stats, e->d_stats, CW_STATS_REPLICAS * CW_STATS_LEN * 8, cudaMemcpyDeviceToHost, e->streams[0]));
    CK(cudaStreamSynchronize(e->streams[0]));
    for (int k = 0; k < CW_STATS_LEN; k++) {                     // sum the replicas
        int64_t v = 0;
        for (int r = 0; r < CW_STATS_REPLICAS; r++) v += e->h_stats[r * CW_STATS_LEN + k];
        stats_host[k] = v;
    }
    return 0;
}

int cw_host_device_state(CwHostEnv* e, CwState* out_state, uint8_t** out_obs) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (out_state) *out_state = e->st;
    if (out_obs) *out_obs = e->d_obs;
    return 0;
}

int cw_host_destroy(CwHostEnv* e) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    cudaSetDevice(e->device);
    if (e->trace && e->tr_steps)
        fprintf(stderr, "cw_host trace: %llu delta steps; launch call %.2f us, launch -> first record %.2f us, whole call %.2f us\n",
                (unsigned long long)e->tr_steps, e->tr_launch / e->tr_steps, e->tr_first / e->tr_steps, e->tr_total / e->tr_steps);
    if (e->streams[0]) cudaStreamSynchronize(e->streams[0]);     // the delta path returns without a stream sync
    if (e->streams[1]) cudaStreamSynchronize(e->streams[1]);
    cudaFree(e->st.grid); cudaFree(e->st.init_grid); cudaFree(e->st.agent); cudaFree(e->st.goal); cudaFree(e->st.t);
    cudaFree(e->st.episode); cudaFree(e->d_actions); cudaFree(e->d_reward); cudaFree(e->d_obs);
    cudaFree(e->d_goal_obs); cudaFree(e->d_stats);
    cudaFreeHost(e->h_actions); cudaFreeHost(e->h_reward); cudaFreeHost(e->h_stats);
    if (e->h_frames) cudaFreeHost(e->h_frames);
    if (e->h_delta) cudaFreeHost(e->h_delta);
    if (e->h_fresh) cudaFreeHost(e->h_fresh);
    free(e->m_grid); free(e->m_agent);
    delete e->pool;
    if (e->streams[0]) cudaStreamDestroy(e->streams[0]);
    if (e->streams[1]) cudaStreamDestroy(e->streams[1]);
    e->magic = 0;
    delete e;
    cudaGetLastError();
    return 0;
}

}  // extern "C"
