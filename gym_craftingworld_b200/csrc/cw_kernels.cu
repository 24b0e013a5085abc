// cw_kernels.cu -- sm_100a kernels + C-ABI launchers of the batched CraftingWorld hot path.
//
//   cw_env_kernel   one CTA per world (grid-stride): [step] -> [auto/forced Philox reset (+imagine_obs goal frame)]
//                   -> render.  The world's grid tile is staged in shared memory; the RGB frame is composed in
//                   shared memory and streamed out with TMA bulk stores (cp.async.bulk shared::cta -> global), so
//                   the global write stream is full-line and issues no LSU store instructions.
//                   Bound: HBM write bandwidth (48*H*W bytes per world-step).
//   cw_step_kernel  one thread per world, K steps per launch, warp-cooperative auto-reset; compact observations.
//                   Bound: latency / issue (tens of bytes per world-step).
//   cw_onehot_kernel one-hot observation_vector expansion (12 bytes per cell).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "cw_b200.h"
#include "cw_device.cuh"

namespace cw {

enum : int {  // internal mode bits of cw_env_kernel
    M_STEP = 1, M_AUTO_RESET = 2, M_FORCE_RESET = 4, M_RENDER = 8, M_IMAGINE_ONLY = 16
};

struct EnvArgs {
    const uint8_t* actions;
    const uint8_t* mask;
    int32_t* reward;
    uint8_t* done;
    uint8_t* obs;
    uint8_t* goal_obs;
    uint8_t* init_obs;       // INIT_OBS copy of the first frame of a new episode (ray.py:193), nullable
    unsigned long long* stats;
    const uint8_t* rgrid;    // render-only entry: grid / agent given directly (state may be partial)
    const uint32_t* ragent;
    int mode;
    int bands_per_chunk;
    uint32_t w_magic;        // floor(2^32 / W) + 1
};

// dynamic shared memory: [sgrid cell_stride][simag cell_stride][frame: kBuf * chunk_bytes]
template <int kBuf>
__global__ void __launch_bounds__(128) cw_env_kernel(const CwConfig cfg, const CwState st, const EnvArgs args) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_lut[9];
    __shared__ uint32_t s_agent, s_goal_agent;
    __shared__ int s_reset;

    uint8_t* sgrid = smem;
    uint8_t* simag = smem + cfg.cell_stride;
    const int H = cfg.H, W = cfg.W;
    const uint32_t band_bytes = 48u * (uint32_t)W;                // 4 pixel rows x 4W pixels x 3 bytes
    const uint32_t chunk_bytes = band_bytes * (uint32_t)args.bands_per_chunk;
    const size_t frame_bytes = (size_t)band_bytes * H;
    uint8_t* fbuf = smem + 2 * cfg.cell_stride;
    const int tid = threadIdx.x;
    const int nchunk16 = cfg.cell_stride >> 4;
    const int mode = args.mode;
    const uint8_t* grid_in = args.rgrid ? args.rgrid : st.grid;
    const uint32_t* agent_in = args.ragent ? args.ragent : st.agent;

    if (tid < 9) s_lut[tid] = kColorLUT[tid];
    int buf = 0;

    for (int64_t n = blockIdx.x; n < st.n; n += gridDim.x) {
        const bool forced = (mode & M_FORCE_RESET) && (!args.mask || args.mask[n]);
        if ((mode & M_FORCE_RESET) && !forced) continue;          // masked reset: untouched worlds are skipped
        // ---- A: stage the grid tile; thread 0 fetches the scalars --------------------------------------
        if (!forced && tid < nchunk16)
            reinterpret_cast<uint4*>(sgrid)[tid] = reinterpret_cast<const uint4*>(grid_in + n * cfg.cell_stride)[tid];
        uint32_t agent = 0, goal = 0;
        int t = 0, a = 6;
        if (tid == 0) {
            agent = agent_in[n];
            if (mode & (M_STEP | M_IMAGINE_ONLY)) goal = st.goal[n];
            if (mode & M_STEP) { t = st.t[n]; a = args.actions[n]; }
            if (kBuf == 1) bulk_wait_read<0>(); else bulk_wait_read<kBuf - 1>();   // frame buffer `buf` is free again
        }
        __syncthreads();
        // ---- B: step (one thread; every cell it needs is in shared memory, init cells in L2) -----------
        if (tid == 0) {
            int do_reset = forced ? 1 : 0;
            if (mode & M_STEP) {
                bool dn; int wcell, wval;
                const int rew = step_core(cfg, sgrid, st.init_grid + n * cfg.cell_stride, agent, goal, t, a, dn, wcell, wval);
                if (wcell >= 0) st.grid[n * cfg.cell_stride + wcell] = (uint8_t)wval;
                args.reward[n] = rew;
                args.done[n] = dn ? 1 : 0;
                if (dn && (mode & M_AUTO_RESET)) {
                    do_reset = 1;
                    if (args.stats) stats_add(cfg, args.stats, goal, t, rew);
                } else {
                    st.agent[n] = agent; st.goal[n] = goal; st.t[n] = t;
                }
            }
            s_agent = agent; s_goal_agent = goal;                 // s_goal_agent carries `goal` into phase C
            s_reset = do_reset;
        }
        __syncthreads();
        // ---- C: reset (+ goal frame) by warp 0 ----------------------------------------------------------
        const bool resetting = s_reset != 0;
        const bool imagining = (resetting && args.goal_obs) || (mode & M_IMAGINE_ONLY);
        if (resetting || imagining) {
            if (tid < 32) {
                WarpPhilox rng;
                uint32_t ag = s_agent, gl = s_goal_agent;
                if (resetting) {
                    reset_warp(cfg, st, n, sgrid, rng, ag, gl);
                    if (tid == 0) { st.agent[n] = ag; st.goal[n] = gl; st.t[n] = 0; s_agent = ag; }
                } else {
                    rng.init(st.seed, st.env_id_base + (uint64_t)n, st.episode[n]);
                }
                if (imagining) {
                    for (int ch = tid; ch < nchunk16; ch += 32)
                        reinterpret_cast<uint4*>(simag)[ch] = reinterpret_cast<const uint4*>(sgrid)[ch];
                    __syncwarp();
                    uint32_t gag = ag;
                    imagine_warp(cfg, simag, gag, gl >> 16, rng);
                    if (tid == 0) s_goal_agent = gag;
                }
            }
            __syncthreads();
            if (imagining) {                                      // goal frame: rare path, chunks serialised
                uint8_t* gdst = args.goal_obs + (size_t)n * frame_bytes;
                for (int band0 = 0; band0 < H; band0 += args.bands_per_chunk) {
                    const int nb = min(args.bands_per_chunk, H - band0);
                    if (tid == 0) bulk_wait_read<0>();
                    __syncthreads();
                    compose_bands(cfg, simag, s_goal_agent, band0, nb, reinterpret_cast<uint32_t*>(fbuf), s_lut, args.w_magic);
                    fence_proxy_async_smem();
                    __syncthreads();
                    if (tid == 0) { bulk_store(gdst + (size_t)band0 * band_bytes, fbuf, band_bytes * nb); bulk_commit(); }
                }
                if (tid == 0) bulk_wait_read<0>();
                __syncthreads();
                buf = 0;
            }
        }
        // ---- D: render the (possibly fresh) state: compose in shared memory, TMA bulk store -------------
        if (mode & M_RENDER) {
            const uint32_t ag = s_agent;
            uint8_t* gdst = args.obs + (size_t)n * frame_bytes;
            for (int band0 = 0; band0 < H; band0 += args.bands_per_chunk) {
                const int nb = min(args.bands_per_chunk, H - band0);
                uint8_t* fb = fbuf + (size_t)buf * chunk_bytes;
                if (band0 > 0) {                                  // further chunks of the same world
                    if (tid == 0) { if (kBuf == 1) bulk_wait_read<0>(); else bulk_wait_read<kBuf - 1>(); }
                    __syncthreads();
                }
                compose_bands(cfg, sgrid, ag, band0, nb, reinterpret_cast<uint32_t*>(fb), s_lut, args.w_magic);
                fence_proxy_async_smem();
                __syncthreads();
                if (tid == 0) {
                    bulk_store(gdst + (size_t)band0 * band_bytes, fb, band_bytes * nb);
                    if (resetting && args.init_obs)               // same chunk, second destination
                        bulk_store(args.init_obs + (size_t)n * frame_bytes + (size_t)band0 * band_bytes, fb, band_bytes * nb);
                    bulk_commit();
                }
                buf = (buf + 1 == kBuf) ? 0 : buf + 1;
            }
        } else {
            __syncthreads();                                      // sgrid / s_* are rewritten next iteration
        }
    }
    if (tid == 0) bulk_wait_all();
}

// one thread per world, K steps per launch (K = 1: cw_step; K > 1: cw_rollout)
__global__ void __launch_bounds__(128) cw_step_kernel(const CwConfig cfg, const CwState st, const uint8_t* __restrict__ actions,
                                                      int32_t* __restrict__ reward, uint8_t* __restrict__ done,
                                                      unsigned long long* stats, int K, int flags) {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = n < st.n;
    const int64_t nn = valid ? n : 0;
    uint8_t* g = st.grid + nn * cfg.cell_stride;
    const uint8_t* ig = st.init_grid + nn * cfg.cell_stride;
    uint32_t agent = 0, goal = 0;
    int t = 0;
    if (valid) { agent = st.agent[n]; goal = st.goal[n]; t = st.t[n]; }
    for (int k = 0; k < K; k++) {
        bool dn = false;
        if (valid) {
            const int a = actions[(size_t)k * st.n + n];
            int wcell, wval;
            const int rew = step_core(cfg, g, ig, agent, goal, t, a, dn, wcell, wval);
            if (reward) reward[(size_t)k * st.n + n] = rew;
            if (done) done[(size_t)k * st.n + n] = dn ? 1 : 0;
            if (dn && (flags & CW_F_AUTO_RESET) && stats) stats_add(cfg, stats, goal, t, rew);
        }
        if (flags & CW_F_AUTO_RESET) {                            // finished worlds are re-seeded by the whole warp
            uint32_t m = __ballot_sync(0xffffffffu, valid && dn);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int64_t env = __shfl_sync(0xffffffffu, n, src);
                WarpPhilox rng;
                uint32_t ag, gl;
                reset_warp(cfg, st, env, nullptr, rng, ag, gl);
                if (lane_id() == src) { agent = ag; goal = gl; t = 0; }
            }
        }
    }
    if (valid) { st.agent[n] = agent; st.goal[n] = goal; st.t[n] = t; }
}

// observation_vector (ray.py:94-98, 605-613): one 32-bit word (4 of the 12 channel bytes of a cell) per thread
__global__ void __launch_bounds__(256) cw_onehot_kernel(const CwConfig cfg, const uint8_t* __restrict__ grid,
                                                        const uint32_t* __restrict__ agent, uint32_t* __restrict__ out,
                                                        int64_t n_words, uint32_t hw3_magic_unused) {
    const int HW = cfg.H * cfg.W;
    for (int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_words; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t cellg = w / 3;
        const int j = (int)(w - cellg * 3);
        const int64_t env = cellg / HW;
        const int cell = (int)(cellg - env * HW);
        const int code = grid[env * cfg.cell_stride + cell];
        const uint32_t ag = agent[env];
        const int acell = (int)(ag & 0xFF) * cfg.W + (int)((ag >> 8) & 0xFF);
        uint32_t v = 0;
        const int ch = code - 1;                                  // object channel 0..7 (or -1)
        if (ch >= 0 && (ch >> 2) == j) v |= 1u << (8 * (ch & 3));
        if (cell == acell && j == 2) {
            v |= 1u;                                              // channel 8: agent
            const int h = (ag >> 16) & 0xFF;
            if (h) v |= 1u << (8 * h);                            // channels 9..11: holding
        }
        out[w] = v;
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
struct OccEntry { int nbuf; size_t smem; int per_sm; };
struct DeviceInfo { int sms = 0; int smem_optin = 0; bool ok = false; bool attr_set[2] = {false, false}; int max_dyn = 0; int n_occ = 0; OccEntry occ[8]; };
static DeviceInfo g_dev[64];

static int device_info(DeviceInfo** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return CW_E_BADCONFIG;
    DeviceInfo& d = g_dev[dev];
    if (!d.ok) {
        e = cudaDeviceGetAttribute(&d.sms, cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return (int)e;
        e = cudaDeviceGetAttribute(&d.smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (e != cudaSuccess) return (int)e;
        d.ok = true;
    }
    *out = &d;
    return 0;
}

static int check_config(const CwConfig* cfg) {
    if (!cfg) return CW_E_NULLPTR;
    if (cfg->H < 1 || cfg->W < 1 || cfg->H > CW_MAX_SIDE || cfg->W > CW_MAX_SIDE) return CW_E_BADCONFIG;
    if (cfg->cell_stride != (cfg->H * cfg->W + 15) / 16 * 16) return CW_E_BADCONFIG;
    if (cfg->max_steps < 1) return CW_E_BADCONFIG;
    if (cfg->n_selected < 1 || cfg->n_selected > 9) return CW_E_BADCONFIG;
    if (cfg->number_of_tasks < 1 || cfg->number_of_tasks > cfg->n_selected) return CW_E_BADCONFIG;
    for (int i = 0; i < cfg->n_selected; i++)
        if (cfg->selected[i] > 15) return CW_E_BADCONFIG;
    return 0;
}
static int check_reset_config(const CwConfig* cfg) {
    return (cfg->H * cfg->W >= 9) ? 0 : CW_E_BADCONFIG;   // 9 distinct cells are needed (the reference needs 12, ray.py:608)
}

static int env_tunable(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}

// frame chunking: the largest number of bands whose chunk fits the per-buffer budget
static int pick_bands(const CwConfig* cfg, int budget_bytes) {
    int forced = env_tunable("CW_BANDS_PER_CHUNK", 0);
    if (forced > 0) return forced < cfg->H ? forced : cfg->H;
    int bands = budget_bytes / (48 * cfg->W);
    if (bands < 1) bands = 1;
    if (bands >= cfg->H) return cfg->H;
    int nchunks = (cfg->H + bands - 1) / bands;                   // equalise the chunks
    return (cfg->H + nchunks - 1) / nchunks;
}

static int launch_env_kernel(const CwConfig* cfg, const CwState* st, EnvArgs args, cudaStream_t stream) {
    DeviceInfo* dev;
    int rc = device_info(&dev);
    if (rc) return rc;
    if (st->n <= 0) return 0;
    const int nbuf = env_tunable("CW_FRAME_BUFFERS", 1) >= 2 ? 2 : 1;
    const bool needs_frame = (args.mode & M_RENDER) || args.goal_obs;
    args.bands_per_chunk = needs_frame ? pick_bands(cfg, env_tunable("CW_CHUNK_BYTES", 25 * 1024)) : 1;
    args.w_magic = (uint32_t)(0x100000000ull / (uint64_t)cfg->W) + 1u;
    const size_t smem = 2 * (size_t)cfg->cell_stride + (needs_frame ? (size_t)nbuf * 48 * cfg->W * args.bands_per_chunk : 0);
    auto kern = nbuf == 2 ? cw_env_kernel<2> : cw_env_kernel<1>;
    // occupancy is queried once per (device, buffers, smem size)
    int per_sm = 0;
    for (int i = 0; i < dev->n_occ; i++)
        if (dev->occ[i].nbuf == nbuf && dev->occ[i].smem == smem) per_sm = dev->occ[i].per_sm;
    if (!dev->attr_set[nbuf - 1]) {   // once per device and kernel: allow any dynamic size up to the opt-in maximum
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return (int)e;
        dev->max_dyn = dev->smem_optin - (int)fa.sharedSizeBytes;   // the opt-in limit covers static + dynamic
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, dev->max_dyn);
        if (e != cudaSuccess) return (int)e;
        dev->attr_set[nbuf - 1] = true;
    }
    if (smem > (size_t)dev->max_dyn) return CW_E_BADCONFIG;
    if (per_sm == 0) {
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
        if (e != cudaSuccess) return (int)e;
        if (per_sm < 1) per_sm = 1;
        if (dev->n_occ < 8) { dev->occ[dev->n_occ].nbuf = nbuf; dev->occ[dev->n_occ].smem = smem; dev->occ[dev->n_occ].per_sm = per_sm; dev->n_occ++; }
    }
    const int cap = env_tunable("CW_CTAS_PER_SM", 0);
    if (cap > 0 && cap < per_sm) per_sm = cap;
    int64_t blocks = (int64_t)dev->sms * per_sm;
    if (blocks > st->n) blocks = st->n;
    kern<<<(unsigned)blocks, 128, smem, stream>>>(*cfg, *st, args);
    return (int)cudaGetLastError();
}

}  // namespace cw

using namespace cw;

extern "C" {

int cw_abi_version(void) { return CW_ABI_VERSION; }

const char* cw_error_string(int code) {
    switch (code) {
        case 0: return "ok";
        case CW_E_BADCONFIG: return "cw: invalid CwConfig / size";
        case CW_E_NULLPTR: return "cw: required pointer is NULL";
        case CW_E_BADFLAGS: return "cw: invalid flags";
        case CW_E_BADHANDLE: return "cw: invalid handle";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "cw: unknown error";
    }
}

static int check_state(const CwState* st) {
    if (!st) return CW_E_NULLPTR;
    if (st->n < 0) return CW_E_BADCONFIG;
    if (st->n > 0 && (!st->grid || !st->init_grid || !st->agent || !st->goal || !st->t || !st->episode)) return CW_E_NULLPTR;
    if (st->n_fixed < 0 || (st->n_fixed > 0 && (!st->fixed_grid || !st->fixed_agent))) return CW_E_NULLPTR;
    return 0;
}

int cw_reset(const CwConfig* cfg, const CwState* st, const uint8_t* mask, uint8_t* obs, uint8_t* goal_obs, uint8_t* init_obs,
             void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_reset_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    EnvArgs a = {};
    if (init_obs && !obs) return CW_E_NULLPTR;
    if (st->n_fixed > 0 && (!st->fixed_grid || !st->fixed_agent)) return CW_E_NULLPTR;
    a.mask = mask; a.obs = obs; a.goal_obs = goal_obs; a.init_obs = init_obs;
    a.mode = M_FORCE_RESET | (obs ? M_RENDER : 0);
    return launch_env_kernel(cfg, st, a, (cudaStream_t)stream);
}

int cw_step(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done, int64_t* stats,
            int flags, void* stream) {
    return cw_rollout(cfg, st, actions, reward, done, stats, 1, flags, stream);
}

int cw_rollout(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done, int64_t* stats,
               int K, int flags, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (flags & ~CW_F_AUTO_RESET) return CW_E_BADFLAGS;
    if ((flags & CW_F_AUTO_RESET) && (rc = check_reset_config(cfg))) return rc;
    if (K < 0) return CW_E_BADCONFIG;
    if (st->n == 0 || K == 0) return 0;
    if (!actions) return CW_E_NULLPTR;
    const int64_t blocks = (st->n + 127) / 128;
    cw_step_kernel<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(*cfg, *st, actions, reward, done,
                                                                        (unsigned long long*)stats, K, flags);
    return (int)cudaGetLastError();
}

int cw_render(const CwConfig* cfg, const uint8_t* grid, const uint32_t* agent, uint8_t* obs, int64_t n, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    if (n < 0) return CW_E_BADCONFIG;
    if (n == 0) return 0;
    if (!grid || !agent || !obs) return CW_E_NULLPTR;
    CwState st = {};
    st.n = n;
    EnvArgs a = {};
    a.obs = obs; a.rgrid = grid; a.ragent = agent; a.mode = M_RENDER;
    return launch_env_kernel(cfg, &st, a, (cudaStream_t)stream);
}

int cw_step_render(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done, uint8_t* obs,
                   uint8_t* goal_obs, uint8_t* init_obs, int64_t* stats, int flags, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (flags & ~CW_F_AUTO_RESET) return CW_E_BADFLAGS;
    if ((flags & CW_F_AUTO_RESET) && (rc = check_reset_config(cfg))) return rc;
    if (st->n == 0) return 0;
    if (!actions || !reward || !done || !obs) return CW_E_NULLPTR;
    EnvArgs a = {};
    a.actions = actions; a.reward = reward; a.done = done; a.obs = obs; a.goal_obs = goal_obs; a.init_obs = init_obs;
    a.stats = (unsigned long long*)stats;
    a.mode = M_STEP | M_RENDER | ((flags & CW_F_AUTO_RESET) ? M_AUTO_RESET : 0);
    return launch_env_kernel(cfg, st, a, (cudaStream_t)stream);
}

int cw_imagine(const CwConfig* cfg, const CwState* st, uint8_t* goal_obs, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (st->n == 0) return 0;
    if (!goal_obs) return CW_E_NULLPTR;
    EnvArgs a = {};
    a.goal_obs = goal_obs; a.mode = M_IMAGINE_ONLY;
    return launch_env_kernel(cfg, st, a, (cudaStream_t)stream);
}

int cw_onehot(const CwConfig* cfg, const uint8_t* grid, const uint32_t* agent, uint8_t* onehot, int64_t n, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    if (n < 0) return CW_E_BADCONFIG;
    if (n == 0) return 0;
    if (!grid || !agent || !onehot) return CW_E_NULLPTR;
    DeviceInfo* dev;
    rc = device_info(&dev); if (rc) return rc;
    const int64_t n_words = n * cfg->H * cfg->W * 3;
    int64_t blocks = (n_words + 255) / 256;
    const int64_t cap = (int64_t)dev->sms * 16;
    if (blocks > cap) blocks = cap;
    cw_onehot_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*cfg, grid, agent, (uint32_t*)onehot, n_words, 0);
    return (int)cudaGetLastError();
}

}  // extern "C"
