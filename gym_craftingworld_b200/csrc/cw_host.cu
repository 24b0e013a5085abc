// cw_host.cu -- host-buffer API: the batched env behind an opaque handle (cw_host_*, see include/cw_b200.h).
//
// Every argument is a HOST pointer.  One call = one reference-style `env.step(actions)` for N worlds.  Three transports:
//   device consumer (obs_host == NULL)   the frames stay in HBM (a ring of four frame buffers); reward / done come back as one
//       self-validating status byte per world in mapped pinned host memory and the call returns when all of them have landed --
//       no stream synchronisation, the frames of step k are written behind step k+1.  Single steps of batches <= 16384 worlds:
//       a two-launch pipeline on two streams (host_step_pipe: thread-per-world step launch + render launch of the state
//       snapshot it publishes); larger batches and cw_host_step_many: one fused launch per step, chained by per-group dataflow.
//   delta (CW_F_DELTA_TRANSPORT)         a thread-per-world kernel writes one pre-digested 16-byte record per world into
//       mapped pinned memory; a small worker pool patches the <= 2 changed cells of each world in the caller's frame buffer
//       while the kernel runs (the reference's render_edit, ray.py:522-557).  No stream synchronisation either.
//   frames                               every rendered frame copied device -> host over PCIe (sliced, two streams).
#include <cuda_runtime.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <new>
#include <thread>
#include <vector>

#include "cw_b200.h"
#include "cw_internal.h"

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

// the same colours repeated over 4 pixels (a full cell row, padded to 16 bytes) and over 2 pixels (an overlay row, padded to
// 8); index 9 = the agent's white.  Codes are 4-bit fields of a record: the tables cover all 16 values.
struct LutRows {
    uint8_t r12[16][16], r6[16][8];
    constexpr LutRows() : r12(), r6() {
        for (int c = 0; c < 16; c++) {
            for (int k = 0; k < 12; k++) r12[c][k] = c < 9 ? kLut[c][k % 3] : 255;
            for (int k = 0; k < 6; k++) r6[c][k] = c < 9 ? kLut[c][k % 3] : 255;
        }
    }
};
constexpr LutRows kRows;
constexpr int kWhite = 9;

inline void store6(uint8_t* p, int code) { memcpy(p, kRows.r6[code], 4); memcpy(p + 4, kRows.r6[code] + 4, 2); }
inline void store12(uint8_t* p, int code) { memcpy(p, kRows.r12[code], 8); memcpy(p + 8, kRows.r12[code] + 8, 4); }

// one cell of a host frame: 4x4 pixels of the object's colour, agent overlay on top (ray.py:550-557)
inline void patch_cell(uint8_t* frame, int W, int cell, int code, bool agent_here, int hold) {
    const int r = cell / W, c = cell - r * W;
    const size_t rowb = (size_t)12 * W;
    uint8_t* p = frame + (size_t)(4 * r) * rowb + 12 * c;
    for (int y = 0; y < 4; y++) store12(p + y * rowb, code);
    if (agent_here) {
        store6(p + rowb + 3, kWhite);                                    // ray.py:555
        store6(p + 2 * rowb + 3, hold ? hold : kWhite);                  // ray.py:556-557
    }
}
// a whole host frame of a world that holds at most 8 objects (a re-seeded world or its imagined goal state, ray.py:442-486):
// everything else is black, so the frame is one memset plus <= 9 cells
// zero a frame that is almost certainly not in this core's cache with non-temporal stores: no read-for-ownership of the 330 lines
inline void zero_frame(uint8_t* frame, size_t bytes) {
#if defined(__SSE2__)
    if ((reinterpret_cast<uintptr_t>(frame) & 15u) == 0 && (bytes & 15u) == 0) {
        const __m128i z = _mm_setzero_si128();
        __m128i* p = reinterpret_cast<__m128i*>(frame);
        for (size_t i = 0; i < bytes / 16; i++) _mm_stream_si128(p + i, z);
        _mm_sfence();                                             // before the cached stores of the cells that follow
        return;
    }
#endif
    memset(frame, 0, bytes);
}
inline void render_sparse(uint8_t* frame, int H, int W, const uint32_t* objs8, uint32_t agent) {
    zero_frame(frame, (size_t)48 * H * W);
    const int acell = (int)(agent & 0xFF) * W + (int)((agent >> 8) & 0xFF), hold = (int)((agent >> 16) & 0xFF);
    int under = 0;
    for (int k = 0; k < 8; k++) {
        const int code = (int)(objs8[k] >> 16), cell = (int)(objs8[k] & 0xFFFFu);
        if (!code) continue;
        if (cell == acell) under = code;
        else patch_cell(frame, W, cell, code, false, 0);
    }
    patch_cell(frame, W, acell, under, true, hold);
}
}  // namespace

struct Bound {                          // a caller buffer declared with cw_host_bind: host address + its device alias when page-locked
    void* host = nullptr;
    void* dev = nullptr;                // non-null: mapped pinned memory, the device reads / writes it directly
};

struct CwHostEnv {
    uint32_t magic;
    CwConfig cfg;
    CwState st;
    int device, flags;
    size_t frame_bytes;
    uint8_t *d_actions, *d_done, *d_obs[4], *d_goal_obs;
    int nring;                        // frame buffers of the device-consumer transport (2..4; CW_HOST_RING)
    int32_t* d_reward;
    int64_t* d_stats;
    uint8_t *h_actions, *h_done;      // pinned (mapped) staging for the small vectors
    int32_t* h_reward;
    int64_t* h_stats;
    uint8_t* h_frames;                // pinned staging for frames when the caller's buffer is pageable
    size_t h_frames_bytes;
    cudaStream_t streams[2];
    int64_t slice;                    // worlds per slice (frames transport)
    int cur;                          // frame buffer holding the current observation (device consumer: alternates)
    Bound b_actions;                  // cw_host_bind_actions
    // device-consumer transport: chained launches + one self-validating status byte per world and step
    uint32_t* d_chain;                // [CW_CHAIN_MAX_POS + N] chain words of cw_step_render_chained
    uint8_t* h_status;                // pinned + mapped: [K][status_stride] status bytes (grown on demand), then [K][N] staged actions
    size_t h_status_bytes;
    int chain_pos;                    // position of the next launch in the open chain (0: the next launch opens one)
    std::vector<uint8_t> pending_lines;   // scratch of collect_status: status lines not yet complete
    // pipelined single steps of the device-consumer transport (small batches): step launch on `s_step`, render launch on streams[0]
    bool pipe_ok;                     // the handle has the snapshot buffers (CW_HOST_PIPE=0 turns the path off)
    bool pipe_active;                 // the last step went through the pipe (a switch to another path drains both streams first)
    cudaStream_t s_step;
    cw::PipeSnap snap[cw::kPipeSlots];
    uint8_t* d_snap;                  // one allocation behind the slots
    uint32_t* d_pipe_words;           // [kPipeSlots] slot-consumed words, then one epoch word per 32 worlds
    uint32_t pipe_seq;                // number of the last pipelined step (never reset: the words on the device are compared with it)
    // delta transport (CW_F_DELTA_TRANSPORT)
    uint4* h_delta;                   // pinned + mapped: per-world delta records written by the kernel
    uint32_t* h_fresh;                // pinned + mapped: sparse records of re-seeded worlds
    uint8_t* mirror_obs;              // caller buffers the delta records currently describe (nullptr: unknown -> refresh)
    uint8_t* mirror_goal;
    WorkerPool* pool;
    uint32_t seq;                     // sequence tag of the last delta step (1..63)
    bool nopatch;                     // CW_HOST_NOPATCH=1 (diagnostics only): consume the records, skip the frame patching
    bool trace;                       // CW_HOST_TRACE=1 (diagnostics only): phase times, printed at destroy
    double tr_launch, tr_first, tr_total, tr_first_byte, tr_mid;   // accumulated microseconds: launch call, launch -> first record / flag seen, whole call
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

// device alias of a page-locked host buffer (cudaHostAlloc / cudaHostRegister / torch pin_memory), nullptr for pageable memory
static void* device_alias(const void* p) {
    cudaPointerAttributes a;
    if (!p || cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

static CwState slice_state(const CwHostEnv* e, int64_t off, int64_t cnt) {
    CwState s = e->st;
    s.grid += off * e->cfg.cell_stride; s.init_grid += off * e->cfg.cell_stride;
    s.agent += off; s.goal += off; s.t += off; s.episode += off;
    s.n = cnt; s.env_id_base += (uint64_t)off;
    return s;
}

static int ensure_pinned(uint8_t** buf, size_t* have, size_t need) {
    if (*have >= need) return 0;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr; *have = 0;
    CK(cudaMallocHost(buf, need));
    *have = need;
    return 0;
}
static int ensure_frame_staging(CwHostEnv* e) { return ensure_pinned(&e->h_frames, &e->h_frames_bytes, (size_t)e->st.n * e->frame_bytes); }

// Wait for the status bytes of one step (mapped pinned memory, zero before the launch; the kernel stores 0x80 | success << 1 | done
// per world, the host pre-sets the padding of the last 64-byte line) and unpack them into the caller's reward / done arrays.
// The bytes do not arrive in order, and every line the device writes is invalidated in this core's cache: a scan that waits on
// line after line pays one serialized miss per line (64 lines = ~5 us at 4096 worlds, measured).  So the lines are swept in blocks
// -- prefetch the block's pending lines, then examine them; complete lines are unpacked and dropped, the rest stay for the next
// sweep -- and the misses overlap.  A watchdog on the stream turns a failed launch into an error instead of a hang.
static int collect_status(const uint8_t* status, int64_t n, int32_t max_steps, int32_t* reward, uint8_t* done, cudaStream_t s,
                          uint8_t* pending, double* t_first_us = nullptr) {
    if (t_first_us) {                                             // (CW_HOST_TRACE) when does the first byte land?
        uint64_t spins = 0;
        const auto t0 = std::chrono::steady_clock::now();
        while (!(*reinterpret_cast<const volatile uint8_t*>(status) & 0x80u) && spins++ < (1ull << 26)) cpu_relax();
        *t_first_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    }
    const int64_t nlines = (n + 63) / 64;
    memset(pending, 1, (size_t)nlines);
    int64_t remaining = nlines;
    // Each status line receives several partial writes from the device; a core that keeps re-reading the line forces every one of
    // them to take it back first.  So do not touch the lines while they fill: watch a single sentinel byte (the last world's),
    // and only then sweep -- by then the other lines are complete or about to be.  (CW_HOST_NO_SENTINEL=1: sweep from the start.)
    static const bool no_sentinel = getenv("CW_HOST_NO_SENTINEL") && *getenv("CW_HOST_NO_SENTINEL") == '1';
    if (!no_sentinel) {
        const volatile uint8_t* last = status + n - 1;
        for (uint64_t spins = 1; !(*last & 0x80u); spins++) {
            cpu_relax();
            if ((spins & 0xFFFF) == 0 && cudaStreamQuery(s) != cudaErrorNotReady) break;   // (the sweeps below sort it out)
        }
    }
    auto unpack_scalar = [&](int64_t w0, int64_t w1) {
        for (int64_t w = w0; w < w1; w++) {
            const uint8_t b = status[w];
            done[w] = b & 1u;
            reward[w] = (b & 2u) ? max_steps : -1;
        }
    };
    int drained = 0;
    for (uint64_t sweeps = 1; remaining; sweeps++) {
        for (int64_t b0 = 0; b0 < nlines; b0 += 32) {
            const int64_t b1 = b0 + 32 < nlines ? b0 + 32 : nlines;
            bool any = false;
            for (int64_t l = b0; l < b1; l++)
                if (pending[l]) { __builtin_prefetch(status + 64 * l); any = true; }
            if (!any) continue;
            for (int64_t l = b0; l < b1; l++) {
                if (!pending[l]) continue;
                const int64_t w0 = 64 * l, cnt = n - w0 < 64 ? n - w0 : 64;
#if defined(__SSE2__)
                const __m128i* p = reinterpret_cast<const __m128i*>(status + w0);
                const __m128i v[4] = {_mm_load_si128(p), _mm_load_si128(p + 1), _mm_load_si128(p + 2), _mm_load_si128(p + 3)};
                if ((_mm_movemask_epi8(v[0]) & _mm_movemask_epi8(v[1]) & _mm_movemask_epi8(v[2]) & _mm_movemask_epi8(v[3])) != 0xFFFF) continue;
                if (cnt == 64) {
                    const __m128i one = _mm_set1_epi8(1), two = _mm_set1_epi8(2);
                    const __m128i hit = _mm_set1_epi32(max_steps + 1), minus1 = _mm_set1_epi32(-1);
                    for (int c = 0; c < 4; c++) {
                        _mm_storeu_si128(reinterpret_cast<__m128i*>(done + w0 + 16 * c), _mm_and_si128(v[c], one));
                        const __m128i succ = _mm_cmpeq_epi8(_mm_and_si128(v[c], two), two);   // 0xFF where reward == max_steps
                        const __m128i lo = _mm_unpacklo_epi8(succ, succ), hi = _mm_unpackhi_epi8(succ, succ);
                        const __m128i m[4] = {_mm_unpacklo_epi16(lo, lo), _mm_unpackhi_epi16(lo, lo), _mm_unpacklo_epi16(hi, hi), _mm_unpackhi_epi16(hi, hi)};
                        for (int q = 0; q < 4; q++)               // -1 + (max_steps + 1) where successful
                            _mm_storeu_si128(reinterpret_cast<__m128i*>(reward + w0 + 16 * c + 4 * q), _mm_add_epi32(minus1, _mm_and_si128(m[q], hit)));
                    }
                } else {
                    unpack_scalar(w0, w0 + cnt);
                }
#else
                bool all = true;
                for (int64_t w = w0; w < w0 + 64; w++) all &= (reinterpret_cast<const volatile uint8_t*>(status)[w] & 0x80u) != 0;
                if (!all) continue;
                unpack_scalar(w0, w0 + cnt);
#endif
                pending[l] = 0;
                remaining--;
            }
        }
        if (!remaining) break;
        if (drained) return drained;                              // (that was the last look after the stream drained)
        asm volatile("" ::: "memory");
        cpu_relax();
        if ((sweeps & 0x3FFF) == 0) {
            const cudaError_t q = cudaStreamQuery(s);
            if (q != cudaErrorNotReady) drained = (int)(q != cudaSuccess ? q : cudaErrorLaunchFailure);   // the bytes are final now
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
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
    e->magic = CW_HOST_MAGIC; e->cfg = *cfg; e->device = device; e->flags = flags;
    e->frame_bytes = (size_t)48 * cfg->H * cfg->W;
    e->st.n = n; e->st.seed = seed; e->st.env_id_base = env_id_base;
    const size_t gb = (size_t)n * cfg->cell_stride;
    const size_t chain_words = (size_t)CW_CHAIN_MAX_POS + (size_t)n;
    int rc = 0;
#define TRY(x) do { if (!rc) { cudaError_t e_ = (x); if (e_ != cudaSuccess) rc = (int)e_; } } while (0)
    TRY(cudaMalloc(&e->st.grid, gb)); TRY(cudaMalloc(&e->st.init_grid, gb));
    TRY(cudaMalloc(&e->st.agent, n * 4)); TRY(cudaMalloc(&e->st.goal, n * 4));
    TRY(cudaMalloc(&e->st.t, n * 4)); TRY(cudaMalloc(&e->st.episode, n * 4));
    TRY(cudaMalloc(&e->d_actions, n));
    TRY(cudaMalloc(&e->d_reward, n * 5));                       // [reward int32 x n][done uint8 x n], one block
    if (!rc) e->d_done = reinterpret_cast<uint8_t*>(e->d_reward) + n * 4;
    TRY(cudaMalloc(&e->d_obs[0], (size_t)n * e->frame_bytes)); TRY(cudaMalloc(&e->d_goal_obs, (size_t)n * e->frame_bytes));
    e->nring = 1;
    if (!(flags & CW_F_DELTA_TRANSPORT)) {                        // (a delta handle renders on the device only to refresh)
        // chained launches overlap step k+1 with the draining stores of step k: that needs >= 2 frame buffers, and a launch may
        // not store before the launch `nring` positions back has completed -- small batches get more slack
        e->nring = (size_t)n * e->frame_bytes * 4 <= ((size_t)1 << 30) ? 4 : 2;
        if (const char* r = getenv("CW_HOST_RING")) { const int v = atoi(r); if (v >= 2 && v <= 4) e->nring = v; }
        for (int i = 1; i < e->nring; i++) TRY(cudaMalloc(&e->d_obs[i], (size_t)n * e->frame_bytes));
    }
    TRY(cudaMalloc(&e->d_stats, CW_STATS_REPLICAS * CW_STATS_LEN * 8));
    TRY(cudaMalloc(&e->d_chain, chain_words * 4));
    // pipelined single steps: worth it where a step is a latency problem (one CTA wave of frames), i.e. for small batches
    const char* pipe_env = getenv("CW_HOST_PIPE");
    const char* nochain_env = getenv("CW_NO_CHAIN");              // (debugger / sanitizer sessions: no launch may spin on another)
    e->pipe_ok = !(flags & CW_F_DELTA_TRANSPORT) && n <= 16384 && !(pipe_env && *pipe_env == '0') && !(nochain_env && *nochain_env == '1');
    const size_t pipe_words = (size_t)cw::kPipeSlots + (size_t)((n + 31) / 32);
    if (e->pipe_ok) {
        const size_t per_slot = 2 * gb + (size_t)n * sizeof(uint4);
        TRY(cudaMalloc(&e->d_snap, per_slot * cw::kPipeSlots));
        TRY(cudaMalloc(&e->d_pipe_words, pipe_words * 4));
        if (!rc) {
            TRY(cudaMemset(e->d_pipe_words, 0, pipe_words * 4));
            for (int i = 0; i < cw::kPipeSlots; i++) {
                uint8_t* base = e->d_snap + per_slot * i;         // (gb is a multiple of 16: every part stays 16-byte aligned)
                e->snap[i].grid = base; e->snap[i].goal = base + gb; e->snap[i].meta = reinterpret_cast<uint4*>(base + 2 * gb);
            }
        }
        int lo = 0, hi = 0;                                       // the step launches must never queue behind frame traffic
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        TRY(cudaStreamCreateWithPriority(&e->s_step, cudaStreamNonBlocking, hi));
    }
    TRY(cudaMallocHost(&e->h_actions, n)); TRY(cudaMallocHost(&e->h_reward, n * 5));
    if (!rc) e->h_done = reinterpret_cast<uint8_t*>(e->h_reward) + n * 4;
    TRY(cudaMallocHost(&e->h_stats, CW_STATS_REPLICAS * CW_STATS_LEN * 8));
    if (flags & CW_F_DELTA_TRANSPORT) {
        TRY(cudaMallocHost(&e->h_delta, n * sizeof(uint4)));
        TRY(cudaMallocHost(&e->h_fresh, n * CW_FRESH_WORDS * sizeof(uint32_t)));
        if (!rc) {
            e->pool = new (std::nothrow) WorkerPool(host_threads(n));
            if (!e->pool) rc = (int)cudaErrorMemoryAllocation;
            if (e->h_delta) memset(e->h_delta, 0, n * sizeof(uint4));   // tag 0 = never written
        }
    }
    if (const char* np = getenv("CW_HOST_NOPATCH")) e->nopatch = *np == '1';
    if (const char* tr = getenv("CW_HOST_TRACE")) e->trace = *tr == '1';
    TRY(cudaStreamCreateWithFlags(&e->streams[0], cudaStreamNonBlocking));
    TRY(cudaStreamCreateWithFlags(&e->streams[1], cudaStreamNonBlocking));
    if (!rc) {
        TRY(cudaMemset(e->st.grid, 0, gb)); TRY(cudaMemset(e->st.init_grid, 0, gb));
        TRY(cudaMemset(e->st.agent, 0, n * 4)); TRY(cudaMemset(e->st.goal, 0, n * 4));
        TRY(cudaMemset(e->st.t, 0, n * 4)); TRY(cudaMemset(e->st.episode, 0, n * 4));
        TRY(cudaMemset(e->d_stats, 0, CW_STATS_REPLICAS * CW_STATS_LEN * 8));
        TRY(cudaMemset(e->d_chain, 0, chain_words * 4));
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

int cw_host_bind_actions(CwHostEnv* e, const uint8_t* actions_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    CK(cudaSetDevice(e->device));
    e->b_actions.host = (void*)actions_host; e->b_actions.dev = device_alias(actions_host);
    return 0;
}

// a path other than the pipelined single step is about to touch the state: drain the pipe's two streams first
static int leave_pipe(CwHostEnv* e) {
    if (!e->pipe_active) return 0;
    CK(cudaStreamSynchronize(e->s_step));
    CK(cudaStreamSynchronize(e->streams[0]));
    e->pipe_active = false;
    e->chain_pos = 0;
    return 0;
}

int cw_host_sync(CwHostEnv* e) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    CK(cudaSetDevice(e->device));
    if (e->s_step) CK(cudaStreamSynchronize(e->s_step));
    CK(cudaStreamSynchronize(e->streams[0]));
    CK(cudaStreamSynchronize(e->streams[1]));
    return 0;
}

int cw_host_reset(CwHostEnv* e, uint8_t* obs_host, uint8_t* goal_obs_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    CK(cudaSetDevice(e->device));
    cudaStream_t s = e->streams[0];
    int rc = leave_pipe(e);
    if (rc) return rc;
    e->chain_pos = 0; e->cur = 0;
    rc = cw_reset(&e->cfg, &e->st, nullptr, e->d_obs[0], e->d_goal_obs, nullptr, s);
    if (rc) return rc;
    const size_t total = (size_t)e->st.n * e->frame_bytes;
    if (obs_host) CK(cudaMemcpyAsync(obs_host, e->d_obs[0], total, cudaMemcpyDeviceToHost, s));       // (pageable targets are staged by the runtime)
    if (goal_obs_host) CK(cudaMemcpyAsync(goal_obs_host, e->d_goal_obs, total, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    e->mirror_obs = (e->flags & CW_F_DELTA_TRANSPORT) ? obs_host : nullptr;
    e->mirror_goal = (e->flags & CW_F_DELTA_TRANSPORT) ? goal_obs_host : nullptr;
    return 0;
}

int cw_host_load_state(CwHostEnv* e, const uint8_t* grid_host, const uint32_t* agent_host, const uint32_t* goal_host,
                       const int32_t* t_host, uint8_t* obs_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    CK(cudaSetDevice(e->device));
    { int rc = leave_pipe(e); if (rc) return rc; }
    CK(cudaStreamSynchronize(e->streams[0])); CK(cudaStreamSynchronize(e->streams[1]));
    const int64_t n = e->st.n;
    const size_t gb = (size_t)n * e->cfg.cell_stride;
    if (grid_host) {                                              // INIT_OBS_VECTOR := the injected state (ray.py:183)
        CK(cudaMemcpy(e->st.grid, grid_host, gb, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(e->st.init_grid, grid_host, gb, cudaMemcpyHostToDevice));
    }
    if (agent_host) CK(cudaMemcpy(e->st.agent, agent_host, n * 4, cudaMemcpyHostToDevice));
    if (goal_host) CK(cudaMemcpy(e->st.goal, goal_host, n * 4, cudaMemcpyHostToDevice));
    if (t_host) CK(cudaMemcpy(e->st.t, t_host, n * 4, cudaMemcpyHostToDevice));
    e->chain_pos = 0; e->cur = 0;
    e->mirror_obs = nullptr;                                      // the caller's frames no longer describe the device state
    int rc = cw_render(&e->cfg, e->st.grid, e->st.agent, e->d_obs[0], n, e->streams[0]);
    if (rc) return rc;
    if (obs_host) CK(cudaMemcpyAsync(obs_host, e->d_obs[0], (size_t)n * e->frame_bytes, cudaMemcpyDeviceToHost, e->streams[0]));
    CK(cudaStreamSynchronize(e->streams[0]));
    if (obs_host && (e->flags & CW_F_DELTA_TRANSPORT)) e->mirror_obs = obs_host;
    return 0;
}

// delta transport: one launch writing 16-byte records into mapped pinned memory; the pool patches the frames WHILE the
// kernel runs.  There is no stream synchronisation on this path: every record carries the step's 6-bit sequence tag
// (one 16-byte store, preceded system-wide by the sparse record of a re-seeded world), the workers are started before
// the launch and poll the records of their slice; when every record of the step has been consumed the step is complete.
// A record is pre-digested by the device -- where the agent stood, where it stands, the object codes of both cells after the
// step, whether the object under the agent changed -- so the host keeps no copy of the grid and the patch loop has no
// dependent loads: it prefetches the <= 4 destination lines of a world a few worlds ahead and then only stores.
static int host_step_delta(CwHostEnv* e, const uint8_t* act_src, int32_t* reward_host, uint8_t* done_host, uint8_t* obs_host) {
    cudaStream_t s = e->streams[0];
    const int64_t n = e->st.n;
    const int H = e->cfg.H, W = e->cfg.W;
    e->chain_pos = 0;
    if (obs_host != e->mirror_obs) {                              // unknown buffer: one full refresh, then deltas
        int rc = cw_render(&e->cfg, e->st.grid, e->st.agent, e->d_obs[0], n, s);
        if (rc) return rc;
        CK(cudaMemcpyAsync(obs_host, e->d_obs[0], (size_t)n * e->frame_bytes, cudaMemcpyDeviceToHost, s));
        CK(cudaStreamSynchronize(s));
        e->mirror_obs = obs_host; e->cur = 0;
    }
    e->seq = e->seq >= 63 ? 1 : e->seq + 1;                       // 1..63; 0 is the never-written state of the buffer
    const uint32_t seq = e->seq;
    const size_t fb = e->frame_bytes;
    uint8_t* goal = e->mirror_goal;
    std::atomic<int> failed{0};
    const bool nopatch = e->nopatch;
    const uint4* recs = e->h_delta;
    const uint32_t* fresh = e->h_fresh;
    const std::function<void(int, int)> job = [&, n, H, W, seq, fb, goal, nopatch, recs, fresh](int tid, int nt) {
        const int64_t lo = n * tid / nt, hi = n * (tid + 1) / nt;
        const size_t rowb = (size_t)12 * W;
        // Records do not arrive in order: a re-seeded world's record follows ~6 us after its neighbours' (its warp runs the
        // Philox reset + imagine_obs first).  Worlds whose record is not there yet are deferred and revisited after the
        // rest of the slice, so one late record does not stall the patching behind it.
        constexpr int kMaxDeferred = 128, kAhead = 8;
        int64_t deferred[kMaxDeferred];
        int ndef = 0;
        const int64_t total = hi - lo;
        auto word = [&](int64_t w, int i) { return reinterpret_cast<const volatile uint32_t*>(recs + w)[i]; };
        auto ready = [&](int64_t w) { return (word(w, 2) >> 26) == seq; };
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
        auto prefetch = [&](int64_t w) {                          // the overlay rows of the cell left and of the cell entered
            const uint32_t x = word(w, 0), z = word(w, 2);
            if ((z >> 26) != seq || ((z >> 24) & 2u)) return;
            uint8_t* frame = obs_host + w * fb;
            const uint8_t* po = frame + (size_t)(4 * (z & 63u) + 1) * rowb + 12 * ((z >> 6) & 63u) + 3;
            const uint8_t* pn = frame + (size_t)(4 * (x & 0xFFu) + 1) * rowb + 12 * ((x >> 8) & 0xFFu) + 3;
            __builtin_prefetch(po, 1); __builtin_prefetch(po + rowb, 1);
            __builtin_prefetch(pn, 1); __builtin_prefetch(pn + rowb, 1);
        };
        if (total > 0 && !wait_for(lo)) return;                   // the step's records start to land
        if (!nopatch) for (int64_t w = lo; w < lo + kAhead && w < hi; w++) prefetch(w);
        for (int64_t it = 0; it < total + ndef; it++) {
            const bool second = it >= total;
            const int64_t w = second ? deferred[it - total] : lo + it;
            if (!second && (w & 3) == 0) __builtin_prefetch(recs + w + 32);   // record lines (4 records each), 8 lines ahead of the scan
            if (!second && !nopatch && w + kAhead < hi) prefetch(w + kAhead);
            if (!ready(w)) {
                if (!second && ndef < kMaxDeferred) { deferred[ndef++] = w; continue; }
                if (!wait_for(w)) return;
            }
            std::atomic_thread_fence(std::memory_order_acquire);
            const uint4 r = recs[w];
            const uint32_t flags = r.z >> 24;
            reward_host[w] = (int32_t)r.w;
            done_host[w] = (uint8_t)(flags & 1u);
            if (nopatch) continue;
            uint8_t* frame = obs_host + w * fb;
            if (flags & 2u) {                                     // re-seeded: the new world (+ its goal frame) from the sparse record
                const uint32_t* fr = fresh + w * CW_FRESH_WORDS;
                render_sparse(frame, H, W, fr, r.x);
                if (goal) render_sparse(goal + w * fb, H, W, fr + 8, fr[16]);
                continue;
            }
            // render_edit (ray.py:522-557) on the <= 2 cells a step can change.  A cell whose OBJECT is unchanged differs
            // only in the centred 2x2 overlay block, so it costs two 6-byte writes instead of four 12-byte rows.
            const uint32_t z = r.z;
            const int orow = (int)(z & 63u), ocol = (int)((z >> 6) & 63u), ocode = (int)((z >> 12) & 15u), ncode = (int)((z >> 16) & 15u);
            const int nrow = (int)(r.x & 0xFF), ncol = (int)((r.x >> 8) & 0xFF), hold = (int)((r.x >> 16) & 0xFF);
            const bool objchg = (z >> 20) & 1u, moved = (orow != nrow) | (ocol != ncol);
            if (!moved && !objchg) continue;                      // nothing visible changed
            uint8_t* pn = frame + (size_t)(4 * nrow) * rowb + 12 * ncol;
            if (moved) {                                          // the agent left a cell whose object did not change
                uint8_t* po = frame + (size_t)(4 * orow + 1) * rowb + 12 * ocol + 3;
                store6(po, ocode); store6(po + rowb, ocode);
            }
            if (objchg) {                                         // the object under the agent changed: all four rows
                for (int y = 0; y < 4; y++) store12(pn + y * rowb, ncode);
            }
            store6(pn + rowb + 3, kWhite);                        // ray.py:555
            store6(pn + 2 * rowb + 3, hold ? hold : kWhite);      // ray.py:556-557
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
    if (rc) { e->mirror_obs = nullptr; return rc; }
    if (failed.load()) {
        e->mirror_obs = nullptr;
        cudaError_t ce = cudaStreamSynchronize(s);
        return (int)(ce != cudaSuccess ? ce : cudaErrorUnknown);
    }
    return 0;
}

// device-consumer transport, K >= 1 consecutive steps: K chained launches of the fused kernel (the frame buffers rotate), each
// storing one status byte per world into mapped host memory the moment that world has stepped.  The host zeroes the bytes,
// enqueues the launches and unpacks reward / done row by row as the bytes arrive; it returns after the LAST step's bytes --
// the frames keep draining on the stream (cw_host_sync, or stream order for a device consumer).
static int host_steps_device(CwHostEnv* e, const uint8_t* act_host, bool act_mapped, int32_t* reward_host, uint8_t* done_host, int K) {
    cudaStream_t s = e->streams[0];
    const int64_t n = e->st.n;
    const size_t stride = ((size_t)n + 63) & ~(size_t)63;
    int rc = ensure_pinned(&e->h_status, &e->h_status_bytes, (size_t)K * stride + (act_mapped ? 0 : (size_t)K * n) + 1024);
    if (rc) return rc;
    memset(e->h_status, 0, (size_t)K * stride);
    if ((size_t)n != stride)                                      // the padding of each row's last line counts as "written"
        for (int k = 0; k < K; k++) memset(e->h_status + (size_t)k * stride + n, 0x80, stride - (size_t)n);
    if (e->pending_lines.size() < stride / 64) e->pending_lines.resize(stride / 64);
    const uint8_t* act_dev = act_host;                            // (UVA: a page-locked host address is valid on the device)
    if (!act_mapped) { uint8_t* a = e->h_status + (size_t)K * stride; memcpy(a, act_host, (size_t)K * n); act_dev = a; }
    const auto t_begin = std::chrono::steady_clock::now();
    static const bool nostatus = getenv("CW_HOST_NOSTATUS") && *getenv("CW_HOST_NOSTATUS") == '1';   // (experiment: cost of the status bytes)
    if (nostatus) {
        for (int k = 0; k < K; k++) {
            e->cur = (e->cur + 1) % e->nring;
            rc = cw::step_render_chained_notify(&e->cfg, &e->st, act_dev + (size_t)k * n, e->d_reward, e->d_done, e->d_obs[e->cur], e->d_goal_obs,
                                                nullptr, e->d_stats, e->flags & CW_F_AUTO_RESET, e->d_chain, e->chain_pos, e->nring, nullptr, s);
            if (rc) return rc;
            e->chain_pos = (e->chain_pos + 1) % CW_CHAIN_MAX_POS;
        }
        CK(cudaStreamSynchronize(s));
        return 0;
    }
    for (int k = 0; k < K; k++) {
        e->cur = (e->cur + 1) % e->nring;
        rc = cw::step_render_chained_notify(&e->cfg, &e->st, act_dev + (size_t)k * n, nullptr, nullptr, e->d_obs[e->cur], e->d_goal_obs,
                                            nullptr, e->d_stats, e->flags & CW_F_AUTO_RESET, e->d_chain, e->chain_pos, e->nring,
                                            e->h_status + (size_t)k * stride, s);
        if (rc) { e->chain_pos = 0; return rc; }
        e->chain_pos = (e->chain_pos + 1) % CW_CHAIN_MAX_POS;
    }
    const auto t_launched = std::chrono::steady_clock::now();
    for (int k = 0; k < K; k++) {
        rc = collect_status(e->h_status + (size_t)k * stride, n, e->cfg.max_steps, reward_host + (size_t)k * n, done_host + (size_t)k * n, s,
                            e->pending_lines.data(), (e->trace && K == 1) ? &e->tr_first_byte : nullptr);
        if (rc) { e->chain_pos = 0; return rc; }
    }
    if (e->trace && K == 1) {
        const auto t_end = std::chrono::steady_clock::now();
        e->tr_launch += std::chrono::duration<double, std::micro>(t_launched - t_begin).count();
        e->tr_first += std::chrono::duration<double, std::micro>(t_end - t_launched).count();
        e->tr_total += std::chrono::duration<double, std::micro>(t_end - t_begin).count();
        e->tr_steps++;
    }
    return 0;
}

// device-consumer transport, ONE step of a small batch: two launches on two streams (cw_internal.h).  The step launch answers in
// the time a thread-per-world kernel needs; the render launch of the same step follows on the frame stream and is waited for by
// nobody but the snapshot slot's next user, kPipeSlots steps later.  The call returns when every status byte has landed.
static int host_step_pipe(CwHostEnv* e, const uint8_t* act_host, const uint8_t* act_dev, int32_t* reward_host, uint8_t* done_host) {
    const int64_t n = e->st.n;
    const size_t stride = ((size_t)n + 63) & ~(size_t)63;
    int rc = ensure_pinned(&e->h_status, &e->h_status_bytes, stride + 1024);
    if (rc) return rc;
    memset(e->h_status, 0, (size_t)n);
    if ((size_t)n != stride) memset(e->h_status + n, 0x80, stride - (size_t)n);
    if (e->pending_lines.size() < stride / 64) e->pending_lines.resize(stride / 64);
    if (!e->pipe_active) {                                        // entering from another path: its launches own the live state
        CK(cudaStreamSynchronize(e->streams[0]));
        e->chain_pos = 0;
        e->pipe_active = true;
    }
    const auto t_begin = std::chrono::steady_clock::now();
    const uint32_t seq = ++e->pipe_seq;
    if (seq == 0) return CW_E_BADCONFIG;                          // (2^32 steps on one handle)
    const int slot = (int)(seq % cw::kPipeSlots);
    uint32_t* words = e->d_pipe_words;
    static const bool skip_render = getenv("CW_PIPE_SKIP_RENDER") && *getenv("CW_PIPE_SKIP_RENDER") == '1';   // (experiment: the step chain alone)
    const uint32_t want = (seq > (uint32_t)cw::kPipeSlots && !skip_render) ? seq - (uint32_t)cw::kPipeSlots : 0u;
    rc = cw::step_snap_launch(&e->cfg, &e->st, act_dev, act_host, e->h_status, &e->snap[slot], words + cw::kPipeSlots, seq, words + slot, want,
                              e->d_stats, e->flags & CW_F_AUTO_RESET, e->s_step);
    if (rc) return rc;
    const auto t_mid = std::chrono::steady_clock::now();
    if (!skip_render) {
        e->cur = (e->cur + 1) % e->nring;
        rc = cw::render_pipe_launch(&e->cfg, n, &e->snap[slot], e->d_obs[e->cur], e->d_goal_obs, e->d_chain, e->chain_pos, e->nring,
                                    words + cw::kPipeSlots, seq, words + slot, e->streams[0]);
        if (rc) return rc;
        e->chain_pos = (e->chain_pos + 1) % CW_CHAIN_MAX_POS;
    }
    const auto t_launched = std::chrono::steady_clock::now();
    if (e->trace) e->tr_mid += std::chrono::duration<double, std::micro>(t_mid - t_begin).count();
    rc = collect_status(e->h_status, n, e->cfg.max_steps, reward_host, done_host, e->s_step, e->pending_lines.data(),
                        e->trace ? &e->tr_first_byte : nullptr);
    if (rc) return rc;
    if (e->trace) {
        const auto t_end = std::chrono::steady_clock::now();
        e->tr_launch += std::chrono::duration<double, std::micro>(t_launched - t_begin).count();
        e->tr_first += std::chrono::duration<double, std::micro>(t_end - t_launched).count();
        e->tr_total += std::chrono::duration<double, std::micro>(t_end - t_begin).count();
        e->tr_steps++;
    }
    return 0;
}

int cw_host_step(CwHostEnv* e, const uint8_t* actions_host, int32_t* reward_host, uint8_t* done_host, uint8_t* obs_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (!actions_host || !reward_host || !done_host) return CW_E_NULLPTR;
    CK(cudaSetDevice(e->device));
    const int64_t n = e->st.n;
    // An action array declared with cw_host_bind_actions and found page-locked is read in place (zero-copy); anything else goes
    // through the handle's own pinned staging -- no pointer is ever assumed to be pinned because it once was.
    const bool act_direct = actions_host == e->b_actions.host && e->b_actions.dev;
    const uint8_t* act_src = actions_host;                        // readable by the host AND (mapped) by the device
    if (!act_direct) { memcpy(e->h_actions, actions_host, n); act_src = e->h_actions; }
    if (e->flags & CW_F_DELTA_TRANSPORT) {
        // a delta handle keeps the caller's frame AND goal mirrors current by records alone (no device frames exist that could
        // refresh the goal mirror), so every step of it must be a delta step
        if (!obs_host) return CW_E_BADCONFIG;
        return host_step_delta(e, act_src, reward_host, done_host, obs_host);
    }
    if (!obs_host) {                                              // (the device reads through the array's device alias: for cudaHostRegister-ed
        const uint8_t* act_dev = act_direct ? (const uint8_t*)e->b_actions.dev : act_src;                                        //  memory it may differ)
        if (e->pipe_ok) return host_step_pipe(e, act_src, act_dev, reward_host, done_host);
        return host_steps_device(e, act_dev, true, reward_host, done_host, 1);
    }
    { int rc = leave_pipe(e); if (rc) return rc; }
    // frames transport: every frame crosses PCIe; slices alternate between two streams so copies overlap kernels
    uint8_t* frames_dst = obs_host;
    const bool direct = device_alias(obs_host) != nullptr;
    if (!direct) { int rc = ensure_frame_staging(e); if (rc) return rc; frames_dst = e->h_frames; }
    e->chain_pos = 0; e->cur = 0;
    int k = 0;
    for (int64_t off = 0; off < n; off += e->slice, k ^= 1) {
        const int64_t cnt = (n - off) < e->slice ? (n - off) : e->slice;
        cudaStream_t s = e->streams[k];
        CK(cudaMemcpyAsync(e->d_actions + off, act_src + off, cnt, cudaMemcpyHostToDevice, s));
        CwState sl = slice_state(e, off, cnt);
        int rc = cw_step_render(&e->cfg, &sl, e->d_actions + off, e->d_reward + off, e->d_done + off,
                                e->d_obs[0] + (size_t)off * e->frame_bytes, e->d_goal_obs + (size_t)off * e->frame_bytes, nullptr,
                                e->d_stats, e->flags & CW_F_AUTO_RESET, s);
        if (rc) return rc;
        CK(cudaMemcpyAsync(e->h_reward + off, e->d_reward + off, cnt * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(e->h_done + off, e->d_done + off, cnt, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(frames_dst + (size_t)off * e->frame_bytes, e->d_obs[0] + (size_t)off * e->frame_bytes,
                           (size_t)cnt * e->frame_bytes, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(e->streams[0]));
    CK(cudaStreamSynchronize(e->streams[1]));
    memcpy(reward_host, e->h_reward, n * 4);
    memcpy(done_host, e->h_done, n);
    if (!direct) memcpy(obs_host, e->h_frames, (size_t)n * e->frame_bytes);
    return 0;
}

int cw_host_step_many(CwHostEnv* e, const uint8_t* actions_host, int K, int32_t* reward_host, uint8_t* done_host, uint8_t* obs_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (K < 0) return CW_E_BADCONFIG;
    if (K == 0) return 0;
    if (!actions_host || !reward_host || !done_host) return CW_E_NULLPTR;
    const int64_t n = e->st.n;
    if (obs_host) {                                               // host frames: K single steps (the frames after the last one remain)
        for (int k = 0; k < K; k++) {
            int rc = cw_host_step(e, actions_host + (size_t)k * n, reward_host + (size_t)k * n, done_host + (size_t)k * n, obs_host);
            if (rc) return rc;
        }
        return 0;
    }
    if (e->flags & CW_F_DELTA_TRANSPORT) return CW_E_BADCONFIG;   // (see cw_host_step)
    CK(cudaSetDevice(e->device));
    { int rc = leave_pipe(e); if (rc) return rc; }
    // open-loop run for a device consumer: K chained launches enqueued back to back, reward / done rows unpacked as they land
    const void* alias = device_alias(actions_host);
    return host_steps_device(e, alias ? (const uint8_t*)alias : actions_host, alias != nullptr, reward_host, done_host, K);
}

int cw_host_stats(CwHostEnv* e, int64_t* stats_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (!stats_host) return CW_E_NULLPTR;
    CK(cudaSetDevice(e->device));
    if (e->pipe_active) CK(cudaStreamSynchronize(e->s_step));     // (the episode statistics are the step launches')
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
    if (out_obs) *out_obs = e->d_obs[e->cur];
    return 0;
}

int cw_host_fetch_frames(CwHostEnv* e, uint8_t* obs_host, uint8_t* goal_obs_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (e->flags & CW_F_DELTA_TRANSPORT) return CW_E_BADCONFIG;   // a delta handle has no current device frames
    CK(cudaSetDevice(e->device));
    const size_t total = (size_t)e->st.n * e->frame_bytes;
    if (obs_host) CK(cudaMemcpyAsync(obs_host, e->d_obs[e->cur], total, cudaMemcpyDeviceToHost, e->streams[0]));
    if (goal_obs_host) CK(cudaMemcpyAsync(goal_obs_host, e->d_goal_obs, total, cudaMemcpyDeviceToHost, e->streams[0]));
    CK(cudaStreamSynchronize(e->streams[0]));
    return 0;
}

int cw_host_stream(CwHostEnv* e, void** out_stream) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (!out_stream) return CW_E_NULLPTR;
    *out_stream = (void*)e->streams[0];
    return 0;
}

int cw_host_destroy(CwHostEnv* e) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    cudaSetDevice(e->device);
    if (e->trace && e->tr_steps)
        fprintf(stderr, "cw_host trace: %llu calls; launch call(s) %.2f us (first of two: %.2f us), launch -> first record / all status bytes %.2f us (first byte %.2f us), whole call %.2f us\n",
                (unsigned long long)e->tr_steps, e->tr_launch / e->tr_steps, e->tr_mid / e->tr_steps, e->tr_first / e->tr_steps, e->tr_first_byte / e->tr_steps, e->tr_total / e->tr_steps);
    if (e->s_step) cudaStreamSynchronize(e->s_step);
    if (e->streams[0]) cudaStreamSynchronize(e->streams[0]);     // neither fast path ends with a stream sync
    if (e->streams[1]) cudaStreamSynchronize(e->streams[1]);
    cudaFree(e->st.grid); cudaFree(e->st.init_grid); cudaFree(e->st.agent); cudaFree(e->st.goal); cudaFree(e->st.t);
    cudaFree(e->st.episode); cudaFree(e->d_actions); cudaFree(e->d_reward); for (int i = 0; i < 4; i++) cudaFree(e->d_obs[i]);
    cudaFree(e->d_goal_obs); cudaFree(e->d_stats); cudaFree(e->d_chain); cudaFree(e->d_snap); cudaFree(e->d_pipe_words);
    if (e->s_step) cudaStreamDestroy(e->s_step);
    cudaFreeHost(e->h_actions); cudaFreeHost(e->h_reward); cudaFreeHost(e->h_stats);
    if (e->h_status) cudaFreeHost(e->h_status);
    if (e->h_frames) cudaFreeHost(e->h_frames);
    if (e->h_delta) cudaFreeHost(e->h_delta);
    if (e->h_fresh) cudaFreeHost(e->h_fresh);
    delete e->pool;
    if (e->streams[0]) cudaStreamDestroy(e->streams[0]);
    if (e->streams[1]) cudaStreamDestroy(e->streams[1]);
    e->magic = 0;
    delete e;
    cudaGetLastError();
    return 0;
}

}  // extern "C"
