// cw_host.cu -- host-buffer API: the batched env behind an opaque handle (cw_host_*, see include/cw_b200.h).
//
// Every argument is a HOST pointer.  One call = one reference-style `env.step(actions)` for N worlds:
// actions go host->device, the fused step+reset+render launch runs in slices, reward/done (and the frames when
// an obs buffer is passed) come back device->host.  Slices alternate between two streams so the D2H copy of
// slice j overlaps the kernel of slice j+1; PCIe, not the kernel, bounds this path when frames are returned.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <new>

#include "cw_b200.h"

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
};

#define CW_HOST_MAGIC 0x43574845u
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return (int)e_; } while (0)

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
    if (flags & ~CW_F_AUTO_RESET) return CW_E_BADFLAGS;
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
    TRY(cudaMalloc(&e->d_stats, CW_STATS_LEN * 8));
    TRY(cudaMallocHost(&e->h_actions, n)); TRY(cudaMallocHost(&e->h_reward, n * 5));
    if (!rc) e->h_done = reinterpret_cast<uint8_t*>(e->h_reward) + n * 4;
    TRY(cudaMallocHost(&e->h_stats, CW_STATS_LEN * 8));
    TRY(cudaStreamCreateWithFlags(&e->streams[0], cudaStreamNonBlocking));
    TRY(cudaStreamCreateWithFlags(&e->streams[1], cudaStreamNonBlocking));
    if (!rc) {
        TRY(cudaMemset(e->st.grid, 0, gb)); TRY(cudaMemset(e->st.init_grid, 0, gb));
        TRY(cudaMemset(e->st.agent, 0, n * 4)); TRY(cudaMemset(e->st.goal, 0, n * 4));
        TRY(cudaMemset(e->st.t, 0, n * 4)); TRY(cudaMemset(e->st.episode, 0, n * 4));
        TRY(cudaMemset(e->d_stats, 0, CW_STATS_LEN * 8));
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
    return 0;
}

int cw_host_step(CwHostEnv* e, const uint8_t* actions_host, int32_t* reward_host, uint8_t* done_host, uint8_t* obs_host) {
    if (!e || e->magic != CW_HOST_MAGIC) return CW_E_BADHANDLE;
    if (!actions_host || !reward_host || !done_host) return CW_E_NULLPTR;
    CK(cudaSetDevice(e->device));
    const int64_t n = e->st.n;
    const bool direct = obs_host && is_pinned(obs_host);
    uint8_t* frames_dst = obs_host;
    if (obs_host && !direct) { int rc = ensure_frame_staging(e); if (rc) return rc; frames_dst = e->h_frames; }
    const uint8_t* act_src = actions_host;
    if (!is_pinned(actions_host)) { memcpy(e->h_actions, actions_host, n); act_src = e->h_actions; }
    if (!obs_host) {
        // Frames stay in HBM for a device-side consumer.  Zero-copy: the kernel reads the actions from, and writes
        // reward/done to, mapped pinned host memory (UVA), so a step is ONE launch + ONE stream sync -- no memcpy
        // launches on the critical path.
        cudaStream_t s = e->streams[0];
        int rc = cw_step_render(&e->cfg, &e->st, act_src /* pinned: the caller's own buffer or our staging copy */, e->h_reward, e->h_done, e->d_obs, e->d_goal_obs, nullptr,
                                e->d_stats, e->flags, s);
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
                                    e->d_stats, e->flags, s);
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
    CK(cudaMemcpyAsync(e->h_stats, e->d_stats, CW_STATS_LEN * 8, cudaMemcpyDeviceToHost, e->streams[0]));
    CK(cudaStreamSynchronize(e->streams[0]));
    memcpy(stats_host, e->h_stats, CW_STATS_LEN * 8);
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
    cudaFree(e->st.grid); cudaFree(e->st.init_grid); cudaFree(e->st.agent); cudaFree(e->st.goal); cudaFree(e->st.t);
    cudaFree(e->st.episode); cudaFree(e->d_actions); cudaFree(e->d_reward); cudaFree(e->d_obs);
    cudaFree(e->d_goal_obs); cudaFree(e->d_stats);
    cudaFreeHost(e->h_actions); cudaFreeHost(e->h_reward); cudaFreeHost(e->h_stats);
    if (e->h_frames) cudaFreeHost(e->h_frames);
    if (e->streams[0]) cudaStreamDestroy(e->streams[0]);
    if (e->streams[1]) cudaStreamDestroy(e->streams[1]);
    e->magic = 0;
    delete e;
    cudaGetLastError();
    return 0;
}

}  // extern "C"
