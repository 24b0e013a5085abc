// cw_kernels.cu -- sm_100a kernels + C-ABI launchers of the batched CraftingWorld hot path.
//
//   cw_env_kernel<V>   groups of worlds per CTA iteration: [step] -> [auto/forced Philox reset (+imagine_obs goal frame)]
//                      -> render.  Grid tiles are staged in shared memory (cp.async), frames composed in a shared-memory
//                      ring and streamed out with TMA bulk stores (cp.async.bulk shared::cta -> global): the global
//                      write stream is full-line and issues no LSU store instructions.
//                      Bound: HBM write bandwidth (48*H*W bytes per world-step).
//                      V_PLAIN   ordinary launch (whole-grid dependency on the previous launch, PDL)
//                      V_CHAINED consecutive step launches linked by per-group dataflow (cw_step_render_chained, 3.1b)
//                      V_LIST    work-list launch: re-seed + render the worlds a preceding step launch queued (3.5)
//                      V_PIPE    render-only member of a chain, fed by the snapshots of cw_step_snap_kernel (host path, 3.6)
//   cw_step_kernel<E>  one thread per world, K steps per launch, warp-cooperative auto-reset; compact observations.
//                      Bound: latency / issue (tens of bytes per world-step).  E = true: the thread also patches the
//                      world's device frame (render_edit) and queues finished worlds instead of re-seeding them.
//   cw_step_chained_kernel   the compact step as a member of a chain: one-warp CTAs linked per warp by release / acquire
//                      marks (3.2).  Bound: the per-warp dependency chain (latency).
//   cw_step_snap_kernel      the step half of the pipelined host path: status bytes, live state, a state snapshot per step (3.6)
//   cw_delta_kernel    thread-per-world step that emits pre-digested 16-byte records for a host-side frame mirror (3.6)
//   cw_onehot_kernel, cw_render_alt_kernel   observation-format expanders: work items staged in shared memory at the
//                      destination's 16-byte phase, TMA bulk stores, two stages (3.4).  Bound: HBM write.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "cw_b200.h"
#include "cw_device.cuh"
#include "cw_internal.h"

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
    uint32_t* list;          // work-list launch (cw_step_render_edit): [0] count, [1] exit ticket, [2..] world ids to re-seed
    // host-buffer API (chained launches): one self-validating status byte per world, 0x80 | success << 1 | done, stored into
    // mapped pinned host memory the moment the world has stepped -- long before its frame is composed.  The host zeroes the
    // bytes before the launch and polls them: no fence, no counter, no stream synchronisation (a system-scope fence issued
    // while the SM streams frames out was measured to return only when the launch was all but over).
    uint8_t* status;
    int ctas_cap;            // > 0: at most this many CTAs per SM (host-driven chained launches: consecutive launches must co-reside)
    const uint8_t* rgrid;    // render-only entry: grid / agent given directly (state may be partial)
    const uint32_t* ragent;
    int mode;
    int group;               // G: worlds per CTA iteration (<= 32); their steps run lane-parallel in warp 0
    int nbuf;                // F: frame-chunk ring slots (2..4)
    int bands_per_chunk;
    int first_split;         // the first frame of a launch is stored in this many pieces (earlier first store)
    uint32_t w_magic;        // floor(2^32 / W) + 1
    // chained launches (cw_step_render_chained): per-group dataflow instead of a whole-grid dependency
    uint32_t* chain;         // [CW_CHAIN_MAX_POS] finished-CTA counters per chain position, then one epoch word per group
    int chain_pos;           // position of this launch in its chain (0 = ordinary launch that opens a chain)
    int chain_ring;          // the chain's frame buffers rotate with this period (>= 1)
    // pipelined host transport (V_PIPE, cw_host.cu): render-only launch fed by the snapshots of cw_step_snap_kernel
    const uint4* pmeta;      // [N] {agent, agent of the imagined goal state, flags, -} of the snapshot slot (rgrid = its grid rows)
    const uint8_t* pgoal;    // [N][cell_stride] imagined goal state of the worlds re-seeded in this step
    const uint32_t* pepoch;  // [ceil(N / 32)] one word per warp of the step launch: the step number it has published
    uint32_t* pslot;         // the slot's "consumed" word: the last CTA out stores pseq (the step kernel waits for it before reuse)
    uint32_t pseq;           // number of this step (1, 2, ...; compared modulo 2^32)
};

// ---- chained launches: acquire / release on the chain words (gpu scope), bounded spins --------------------------
__device__ __forceinline__ uint32_t ld_acquire_gpu(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(uint32_t* p, uint32_t v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// spin until *p >= want.  A chain that was set up wrongly must not hang the GPU for ever: after g_chain_timeout_ns (10 s unless
// CW_CHAIN_TIMEOUT_MS says otherwise; a legitimately slow predecessor under a debugger / sanitizer / time-slicing needs far
// less, and CW_NO_CHAIN=1 turns chained launches into ordinary ones for such sessions) the kernel traps -- loudly, instead of
// returning frames that were never produced.
__device__ unsigned long long g_chain_timeout_ns = 10000000000ull;
__device__ __forceinline__ void chain_wait_ge(const uint32_t* p, uint32_t want) {
    if (ld_acquire_gpu(p) >= want) return;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        __nanosleep(64);
        if (ld_acquire_gpu(p) >= want) return;
        if (global_timer_ns() - t0 > g_chain_timeout_ns) __trap();
    }
}

// the same for a step NUMBER that may wrap: spin until *p has reached `want` (signed distance)
__device__ __forceinline__ void chain_wait_seq(const uint32_t* p, uint32_t want) {
    if ((int32_t)(ld_acquire_gpu(p) - want) >= 0) return;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        __nanosleep(64);
        if ((int32_t)(ld_acquire_gpu(p) - want) >= 0) return;
        if (global_timer_ns() - t0 > g_chain_timeout_ns) __trap();
    }
}

#ifdef CW_TIMING
__device__ unsigned long long* g_dbg = nullptr;   // [CTA][16] globaltimer stamps (experiments only)
__device__ __forceinline__ unsigned long long gtimer() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
__device__ int g_dbg_per_pos = 0;                  // > 0: the stamps of chain position p go to rows [p * g_dbg_per_pos, ...) (timeline of a whole chain)
#define CW_STAMP(slot) do { if (g_dbg && (threadIdx.x == 0 || (slot) >= 8) ) g_dbg[((size_t)dbg_row0 + blockIdx.x) * 16 + (slot)] = gtimer(); } while (0)
#define CW_WSTAMP(slot) do { if (g_dbg && (threadIdx.x & 31) == 0) g_dbg[((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 16 + (slot)] = gtimer(); } while (0)
#define CW_CSTAMP(slot) do { if (g_dbg && g_dbg_per_pos > 0 && lane_id() == 0 && (int)blockIdx.x < g_dbg_per_pos) g_dbg[((size_t)cpos * g_dbg_per_pos + blockIdx.x) * 16 + (slot)] = gtimer(); } while (0)
#define CW_SSTAMP(slot) do { if (g_dbg && (threadIdx.x & 31) == 0 && blockIdx.x < 400) g_dbg[((size_t)600 + blockIdx.x) * 16 + (slot)] = gtimer(); } while (0)
#else
#define CW_STAMP(slot) do { } while (0)
#define CW_WSTAMP(slot) do { } while (0)
#define CW_SSTAMP(slot) do { } while (0)
#define CW_CSTAMP(slot) do { } while (0)
#endif

enum : int { FL_RENDER = 1, FL_FRESH = 2, FL_GOAL = 4, FL_PENDING = 8 /* reset warp has work on this world */ };
constexpr int kComposeThreads = 128;   // warps 0..3 expand frames (warp 0 also steps the group)
constexpr int kEnvThreads = 160;       // + warp 4: the reset warp
enum : int { BAR_COMPOSE = 1, BAR_RESET_DONE = 2 };

// One CTA iteration handles a GROUP of G consecutive worlds:
//   A  grid tiles of the group arrive in shared memory (cp.async, prefetched one group ahead; scalars are
//      prefetched into registers of warp 0, lane i = world i of the group)
//   B  warp 0 steps the G worlds lane-parallel on the shared tiles and publishes per-world flags
//   R  warp 4 (the reset warp) re-seeds finished / forced worlds -- Philox reset + imagine_obs into a scratch
//      tile: ~2.5k dependent instructions on one warp -- WHILE
//   C  warps 0..3 expand the other worlds one at a time into a ring of F frame slots; thread 0 streams each slot
//      out with a TMA bulk store and only waits for the store issued F-1 slots earlier, so composing overlaps
//      the stores.  Worlds the reset warp worked on are emitted last, after its named-barrier arrival.
// dynamic shared memory: [tiles 2 x G x cell_stride][imagine scratch G x cell_stride][ring F x chunk_bytes]
// Three instantiations (V_PLAIN ordinary launch, V_CHAINED chain protocol, V_LIST work-list re-seed): separate code, so the
// ordinary launch is exactly what it was -- the extra control flow measurably slowed it (6 % at 131072 worlds) when all
// shared one kernel body.
// V_PIPE (pipelined host transport, section 3.6): a render-only member of a chain.  The state it renders is the SNAPSHOT a
// cw_step_snap_kernel launch on another stream published (per 32 worlds, release / acquire), so it never waits for a grid and
// nothing waits for it except the snapshot slot's next user and the frame buffer's next writer.
enum : int { V_PLAIN = 0, V_CHAINED = 1, V_LIST = 2, V_PIPE = 3 };
template <int kVariant>
__global__ void __launch_bounds__(kEnvThreads, 4) cw_env_kernel(const CwConfig cfg, const CwState st, const EnvArgs args) {
    constexpr bool kChained = kVariant == V_CHAINED, kList = kVariant == V_LIST, kPipe = kVariant == V_PIPE;
    constexpr bool kLinked = kChained || kPipe;                   // launches linked by dataflow: no whole-grid wait at the start
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ uint32_t s_lut[9];
    __shared__ uint32_t s_agent[32], s_gagent[32], s_goal[32], s_ep[32];
    __shared__ uint32_t s_flag[32];
    __shared__ uint32_t s_obj[8];
    __shared__ uint32_t s_anypend;
    __shared__ uint32_t s_pre[kLinked ? 5 : 1][32];                            // chained: next group's scalars, parked by the reset warp

    const int H = cfg.H, W = cfg.W, cs = cfg.cell_stride;
    const int G = args.group, F = args.nbuf, mode = args.mode;
    const uint32_t band_bytes = 48u * (uint32_t)W;                // 4 pixel rows x 4W pixels x 3 bytes
    const uint32_t chunk_bytes = band_bytes * (uint32_t)args.bands_per_chunk;
    const size_t frame_bytes = (size_t)band_bytes * H;
    uint8_t* tiles = smem;
    uint8_t* simag = smem + 2 * G * cs;
    uint8_t* ring = smem + 3 * G * cs;
    const int tid = threadIdx.x;
    const bool composer = tid < kComposeThreads;
    const int nchunk16 = cs >> 4;
    const uint8_t* grid_in = args.rgrid ? args.rgrid : st.grid;
    const uint32_t* agent_in = args.ragent ? args.ragent : st.agent;
    int64_t ngroups = (st.n + G - 1) / G;
    // a pure (masked) reset overwrites the tiles of the worlds it touches and skips the others: nothing to load
    const bool need_tiles = (mode & (M_STEP | M_IMAGINE_ONLY)) || !(mode & M_FORCE_RESET);
    // Chained launch (cw_step_render_chained, position > 0): the previous launch in the stream is the same kernel on the
    // same worlds, one chain position earlier.  Instead of waiting for that whole grid (and for its frame stores to
    // drain), a CTA waits per GROUP for the predecessor's state of exactly the worlds it is about to step (epoch word,
    // release/acquire), and -- before its first frame store -- for the launch that last wrote the same frame buffer
    // (finished-CTA counter, `chain_ring` positions back).  griddepcontrol.wait moves to the END of the kernel, so
    // grids still COMPLETE in stream order.
    const uint32_t cpos = (uint32_t)args.chain_pos;
    constexpr bool chained = kChained;
    const bool chain_follow = kLinked && cpos > 0;
    uint32_t* const c_fin = args.chain;
    uint32_t* const c_epoch = args.chain + CW_CHAIN_MAX_POS;

    // Programmatic dependent launch: let the next launch in the stream start its prologue now, and do not touch
    // anything the previous launch wrote (state, frames) until it has fully completed.
#ifdef CW_TIMING
    const size_t dbg_row0 = (size_t)(args.chain ? args.chain_pos : 0) * (size_t)g_dbg_per_pos;
#endif
    CW_STAMP(0);
    pdl_launch_dependents();
    if (tid < 9) s_lut[tid] = kColorLUT[tid];
    if (!chain_follow) pdl_wait();
    CW_STAMP(1);
    // work-list launch: one world per CTA iteration (G == 1), the worlds are the ids the preceding step launch appended
    if (kList) ngroups = (int64_t)__ldcg(args.list);
    auto first_world = [&](int64_t g) -> int64_t { return kList ? (int64_t)__ldcg(args.list + 2 + g) : g * G; };

    // tile + scalar prefetch of group `g` (tiles -> stage `sgi`; scalars -> registers of warp 0).  Ordinary launches:
    // issued by all threads at the top of the iteration before.
    uint32_t p_agent = 0, p_goal = 0, p_ep = 0;
    int p_t = 0, p_a = 6, p_forced = 0;
    int64_t p_e0 = 0;
    auto prefetch = [&](int64_t g, int sgi) {
        if (g < ngroups) {
            const int64_t e0 = first_world(g);
            p_e0 = e0;
            const int cnt = (int)min((int64_t)G, st.n - e0);
            const uint8_t* src = grid_in + e0 * cs;
            uint8_t* dst = tiles + (size_t)sgi * G * cs;
            if (need_tiles)
                for (int i = tid; i < cnt * nchunk16; i += kEnvThreads) cp_async16(dst + 16 * i, src + 16 * i);
            if (tid < cnt) {                                      // L2 loads (.cg): streaming, no reuse in L1
                const int64_t e = e0 + tid;
                p_agent = __ldcg(agent_in + e);
                if (mode & (M_STEP | M_IMAGINE_ONLY)) p_goal = __ldcg(st.goal + e);
                if (mode & M_STEP) { p_t = __ldcg(st.t + e); p_a = args.actions[e]; }
                if (mode & M_FORCE_RESET) p_forced = (!args.mask || args.mask[e]) ? 1 : 0;
                if (mode & (M_AUTO_RESET | M_FORCE_RESET | M_IMAGINE_ONLY)) p_ep = __ldcg(st.episode + e);
            }
        }
        cp_async_commit();
    };
    // Chained launches: the RESET WARP alone fetches the next group, after its own work of the current iteration and
    // while the compose warps are busy -- it acquires the group's epoch word (the predecessor launch has published the
    // state of exactly these worlds), issues the tile copies and parks the scalars in shared memory for warp 0.  The
    // acquire (an L2 round trip + L1 invalidate) is therefore never on the compose warps' path.
    auto prefetch_chained = [&](int64_t g, int sgi) {             // reset warp only (step + render [+ auto-reset] mode)
        const int lane = tid - kComposeThreads;
        if (g < ngroups) {
            if (cpos > 0) chain_wait_ge(c_epoch + g, cpos);
            const int64_t e0 = g * G;
            const int cnt = (int)min((int64_t)G, st.n - e0);
            const uint8_t* src = grid_in + e0 * cs;
            uint8_t* dst = tiles + (size_t)sgi * G * cs;
            for (int i = lane; i < cnt * nchunk16; i += 32) cp_async16(dst + 16 * i, src + 16 * i);
            if (lane < cnt) {
                const int64_t e = e0 + lane;
                s_pre[0][lane] = __ldcg(agent_in + e);
                s_pre[1][lane] = __ldcg(st.goal + e);
                s_pre[2][lane] = (uint32_t)__ldcg(st.t + e);
                s_pre[3][lane] = args.actions[e];
                s_pre[4][lane] = (mode & M_AUTO_RESET) ? __ldcg(st.episode + e) : 0u;
            }
        }
        cp_async_commit();
    };

    // Pipelined launches: the same division of labour; the group's worlds were published by the step launch's warps
    // [e0 / 32, (e0 + cnt - 1) / 32] (one epoch word each; G <= 32: at most two), its tiles are rows of the snapshot.
    auto prefetch_pipe = [&](int64_t g, int sgi) {                // reset warp only
        const int lane = tid - kComposeThreads;
        if (g < ngroups) {
            const int64_t e0 = g * G;
            const int cnt = (int)min((int64_t)G, st.n - e0);
            const int64_t b0 = e0 >> 5, b1 = (e0 + cnt - 1) >> 5;
            chain_wait_seq(args.pepoch + b0, args.pseq);
            if (b1 != b0) chain_wait_seq(args.pepoch + b1, args.pseq);
            const uint8_t* src = grid_in + e0 * cs;
            uint8_t* dst = tiles + (size_t)sgi * G * cs;
            for (int i = lane; i < cnt * nchunk16; i += 32) cp_async16(dst + 16 * i, src + 16 * i);
            if (lane < cnt) {
                const uint4 m = __ldcg(args.pmeta + e0 + lane);
                s_pre[0][lane] = m.x; s_pre[1][lane] = m.y; s_pre[2][lane] = m.z;
            }
        }
        cp_async_commit();
    };

    int slot = 0;                                                 // ring slot of the next frame chunk
    const int first_bands = max(1, (args.bands_per_chunk + args.first_split - 1) / args.first_split);
    // expand one world (tile `src`) into the ring and stream it to dst (and dst2 when non-null); compose warps only
    auto emit_frame = [&](const uint8_t* src, uint32_t ag, uint8_t* dst, uint8_t* dst2, int bands_per_store) {
        for (int band0 = 0; band0 < H; band0 += bands_per_store) {
            const int nb = min(bands_per_store, H - band0);
            uint8_t* fb = ring + (size_t)slot * chunk_bytes;
            compose_bands(cfg, src, ag, band0, nb, reinterpret_cast<uint32_t*>(fb), s_lut, args.w_magic, tid, kComposeThreads);
            fence_proxy_async_smem();
            if (tid == 0) bulk_wait_read_dyn(F - 2);              // frees the slot the NEXT chunk composes into
            bar_sync(BAR_COMPOSE, kComposeThreads);
            if (tid == 0) {
                bulk_store(dst + (size_t)band0 * band_bytes, fb, band_bytes * nb);
                if (dst2) bulk_store(dst2 + (size_t)band0 * band_bytes, fb, band_bytes * nb);
                bulk_commit();
            }
            slot = (slot + 1 == F) ? 0 : slot + 1;
        }
    };

    int stage = 0;
    if constexpr (kLinked) {
        if (!composer) {
            if constexpr (kPipe) prefetch_pipe(blockIdx.x, 0); else prefetch_chained(blockIdx.x, 0);
            // frame buffers rotate with period chain_ring: the launch that last wrote the buffer this one is about to
            // write must have completed (all its CTAs counted).  With a ring >= 2 that launch is long gone; with a
            // single buffer this makes the launch wait for its predecessor like an ordinary one.
            if (args.chain_pos >= args.chain_ring) chain_wait_ge(c_fin + (cpos - (uint32_t)args.chain_ring), gridDim.x);
        }
    } else {
        prefetch(blockIdx.x, 0);
    }
    for (int64_t gi = blockIdx.x; gi < ngroups; gi += gridDim.x) {
        // ---- A: this group's scalars move to `c_*`; the next group's tiles + scalars start loading ------------
        uint32_t c_agent = p_agent, c_goal = p_goal, c_ep = p_ep;
        int c_t = p_t, c_a = p_a;
        const int c_forced = p_forced;
        const int64_t c_e0 = p_e0;
        if constexpr (kLinked) {
            cp_async_wait<0>();                                   // (reset warp) this group's tiles have landed
        } else {
            prefetch(gi + gridDim.x, stage ^ 1);
            cp_async_wait<1>();                                   // everything but the newest group has landed
        }
        __syncthreads();
        if constexpr (kLinked) {
            if (tid == 0 && gi == (int64_t)blockIdx.x) fence_proxy_async_all();   // orders the bulk stores after the acquire above
            if (tid < 32) { c_agent = s_pre[0][tid]; c_goal = s_pre[1][tid]; c_t = (int)s_pre[2][tid]; c_a = (int)s_pre[3][tid]; c_ep = s_pre[4][tid]; }
        }
        CW_STAMP(2);
        uint8_t* gt = tiles + (size_t)stage * G * cs;             // this group's tiles
        const int64_t e0 = kList ? c_e0 : gi * G;
        // ---- B: warp 0 steps the group lane-parallel -------------------------------------------------------------
        if (tid < 32) {
            const int lane = tid;
            const int64_t e = e0 + lane;
            const bool valid = lane < G && e < st.n;
            uint32_t agent = c_agent, goal = c_goal, flag = 0;
            if (valid) {
                const bool skip = (mode & M_FORCE_RESET) && !c_forced;   // masked reset: untouched worlds are skipped
                if (!skip && (mode & M_RENDER)) flag |= FL_RENDER;
                if ((mode & M_FORCE_RESET) && c_forced) flag |= FL_PENDING;
                if (mode & M_IMAGINE_ONLY) flag |= FL_PENDING;
                if constexpr (kPipe) {                            // c_goal / c_t carry the snapshot's goal-state agent word / flags
                    if ((c_t & 1) && args.goal_obs) flag |= FL_PENDING;      // re-seeded in this step: its goal frame is due
                }
                if (mode & M_STEP) {
                    int t = c_t, wcell, wval;
                    bool dn;
                    const int rew = step_core(cfg, gt + lane * cs, st.init_grid + e * cs, agent, goal, t, c_a, dn, wcell, wval);
                    if (wcell >= 0) st.grid[e * cs + wcell] = (uint8_t)wval;
                    if (args.reward) args.reward[e] = rew;
                    if (args.done) args.done[e] = dn ? 1 : 0;
                    if constexpr (!kList) {
                        // st.wt: written THROUGH the L2 to system memory now -- a plain store of a partial sector can sit in L2 until
                        // the frame stream evicts it or the grid ends, i.e. the host would see it when the launch is all but over
                        if (args.status) __stwt(args.status + e, (uint8_t)(0x80u | (rew == cfg.max_steps ? 2u : 0u) | (dn ? 1u : 0u)));
                    }
                    if (dn && (mode & M_AUTO_RESET)) {
                        flag |= FL_PENDING;
                        // (replica chosen by the WORLD, not by the CTA: the same counts whatever the launch geometry)
                        if (args.stats) stats_add(cfg, args.stats + ((e >> 5) % CW_STATS_REPLICAS) * CW_STATS_LEN, goal, t, rew);
                    } else {
                        st.agent[e] = agent; st.goal[e] = goal; st.t[e] = t;
                    }
                }
            }
            if (lane < G) { s_agent[lane] = agent; s_goal[lane] = goal; s_flag[lane] = flag; s_ep[lane] = kPipe ? (uint32_t)c_t : c_ep; }
            const uint32_t pend = __ballot_sync(0xffffffffu, (flag & FL_PENDING) != 0);
            if (lane == 0) s_anypend = pend;
        }
        __syncthreads();
        CW_STAMP(3);
        if (!composer) {
            // ---- R: the reset warp works through the pending worlds, then arrives on BAR_RESET_DONE -----------------
            const int lane = tid - kComposeThreads;
            // chained: publish this group's state for the next chain position.  Without a re-seeded world the state is
            // final right here (warp 0 wrote it before the barrier); otherwise thread 0 publishes after the re-seeded
            // worlds' frames (goal / init frames have no ring) have been written completely.
            if (chained && s_anypend == 0 && lane == 0) st_release_gpu(c_epoch + gi, cpos + 1u);   // release is cumulative over the barrier
            for (int i = 0; i < G; i++) {
                uint32_t flag = s_flag[i];
                if (!(flag & FL_PENDING)) continue;
                const int64_t er = e0 + i;
                uint8_t* tile = gt + i * cs;
                if constexpr (kPipe) {
                    // the imagined goal state was drawn by the step launch: fetch its tile.  The goal frames have no ring: if the
                    // episode that just ended was shorter than the frame ring, the launch that wrote this world's previous goal
                    // frame may still be running -- wait for those launches (1 .. length positions back) to complete first.
                    uint8_t* im = simag + i * cs;
                    for (int ch = lane; ch < nchunk16; ch += 32)
                        reinterpret_cast<uint4*>(im)[ch] = __ldcg(reinterpret_cast<const uint4*>(args.pgoal + er * cs) + ch);
                    const uint32_t len = s_ep[i] >> 8;
                    for (uint32_t j = 1; j <= len && j < (uint32_t)args.chain_ring && j <= cpos; j++) chain_wait_ge(c_fin + (cpos - j), gridDim.x);
                    if (lane == 0) { s_gagent[i] = s_goal[i]; s_flag[i] = flag | FL_GOAL; }
                    __syncwarp();
                    continue;
                }
                WarpPhilox rng;
                uint32_t ag = s_agent[i], gl = s_goal[i];
                if (!(mode & M_IMAGINE_ONLY)) {                   // reset(): ray.py:156-218
                    Sparse8 objs;
                    reset_warp(cfg, st, er, tile, rng, ag, gl, s_ep[i], &objs, s_obj);
                    if (lane == 0) { st.agent[er] = ag; st.goal[er] = gl; st.t[er] = 0; s_agent[i] = ag; }
                    flag |= FL_FRESH;
                    if (lane == 0 && st.init_agent) st.init_agent[er] = ag;
                    if (args.goal_obs || st.goal_grid) {          // desired_goal = imagine_obs(): ray.py:191, 220-299
                        uint32_t gag = ag;
                        imagine_fresh(cfg, objs, gag, gl >> 16, rng);      // closed form on the 8-object list
                        tile_from_objects(objs, nchunk16, simag + i * cs);
                        if (st.goal_grid) {                       // compact goal state (one-hot observation family)
                            for (int ch = lane; ch < nchunk16; ch += 32)
                                reinterpret_cast<uint4*>(st.goal_grid + er * cs)[ch] = reinterpret_cast<const uint4*>(simag + i * cs)[ch];
                            if (lane == 0) st.goal_agent[er] = gag;
                        }
                        if (lane == 0) s_gagent[i] = gag;
                        if (args.goal_obs) flag |= FL_GOAL;
                    }
                } else {                                          // cw_imagine: arbitrary (dense) injected state
                    rng.init(st.seed, st.env_id_base + (uint64_t)er, s_ep[i]);
                    uint8_t* im = simag + i * cs;
                    for (int ch = lane; ch < nchunk16; ch += 32)
                        reinterpret_cast<uint4*>(im)[ch] = reinterpret_cast<const uint4*>(tile)[ch];
                    __syncwarp();
                    uint32_t gag = ag;
                    imagine_warp(cfg, im, gag, gl >> 16, rng);
                    if (st.goal_grid) {
                        for (int ch = lane; ch < nchunk16; ch += 32)
                            reinterpret_cast<uint4*>(st.goal_grid + er * cs)[ch] = reinterpret_cast<const uint4*>(im)[ch];
                        if (lane == 0) st.goal_agent[er] = gag;
                    }
                    if (lane == 0) s_gagent[i] = gag;
                    if (args.goal_obs) flag |= FL_GOAL;
                }
                if (lane == 0) s_flag[i] = flag;
                __syncwarp();
            }
            __threadfence_block();                                // tiles / s_* written above are visible to the composers
#ifdef CW_TIMING
            if (tid == kComposeThreads && g_dbg) { int np = 0; for (int i = 0; i < G; i++) np += (s_flag[i] & FL_PENDING) ? 1 : 0; g_dbg[((size_t)dbg_row0 + blockIdx.x) * 16 + 9] = np; }
#endif
            if (tid == kComposeThreads) CW_STAMP(8);
            bar_arrive(BAR_RESET_DONE, kEnvThreads);
            if constexpr (kChained) prefetch_chained(gi + gridDim.x, stage ^ 1);   // s_pre was consumed before this iteration's 2nd barrier
            if constexpr (kPipe) prefetch_pipe(gi + gridDim.x, stage ^ 1);
        } else {
            // ---- C: expand + stream out: untouched worlds first, worlds from the reset warp after its arrival ------
            // the very first frame of the launch goes out in quarter-frame stores: the launch is bound by the DRAM
            // write-back window that opens with the first store, so open it as early as possible
            bool first = gi == (int64_t)blockIdx.x;
            for (int i = 0; i < G; i++) {
                const uint32_t flag = s_flag[i];
                if (!(flag & FL_RENDER) || (flag & FL_PENDING)) continue;
                const size_t off = (size_t)(e0 + i) * frame_bytes;
                emit_frame(gt + i * cs, s_agent[i], args.obs + off, nullptr, first ? first_bands : args.bands_per_chunk);
                first = false;
            }
            CW_STAMP(4);
            bar_sync(BAR_RESET_DONE, kEnvThreads);
            CW_STAMP(5);
            for (int i = 0; i < G; i++) {
                const uint32_t flag = s_flag[i];
                if (!(flag & FL_PENDING)) continue;
                const size_t off = (size_t)(e0 + i) * frame_bytes;
                if (flag & FL_GOAL) emit_frame(simag + i * cs, s_gagent[i], args.goal_obs + off, nullptr, args.bands_per_chunk);
                if (flag & FL_RENDER)
                    emit_frame(gt + i * cs, s_agent[i], args.obs + off, ((flag & FL_FRESH) && args.init_obs) ? args.init_obs + off : nullptr,
                               args.bands_per_chunk);
            }
            if (chained && s_anypend != 0 && tid == 0) { bulk_wait_all(); st_release_gpu(c_epoch + gi, cpos + 1u); }
        }
        CW_STAMP(6);
        __syncthreads();                                          // tiles / s_* of this stage are rewritten next
        stage ^= 1;
    }
    cp_async_wait<0>();
    if (tid == 0) {
        bulk_wait_all();
        if (kLinked) {                                            // this CTA's frames of chain position cpos are complete
            __threadfence();
            const uint32_t before = atomicAdd(c_fin + cpos, 1u);
            if (kPipe && before == gridDim.x - 1) st_release_gpu(args.pslot, args.pseq);   // every CTA is done with the snapshot slot
            // The grid must not COMPLETE before its predecessor grid has (whatever follows the chain in the stream sees all of
            // it).  ONE resident thread is enough for that -- the last CTA out; if every CTA waited here, each would hold its
            // SM slot and shared memory until the predecessor's completion flush, delaying the launch after this one.
            if (chain_follow && before == gridDim.x - 1) pdl_wait();
        }
        if (kList) {                                              // the last CTA out (all have read the count) empties the list
            __threadfence();
            if (atomicAdd(args.list + 1, 1u) == gridDim.x - 1) { args.list[0] = 0; args.list[1] = 0; }
        }
    }
    CW_STAMP(7);
}

// one thread per world, K steps per launch (K = 1: cw_step; K > 1: cw_rollout)
// render_edit (ray.py:522-557) on the device frame of one world: the <= 2 cells a step can change.  A cell whose OBJECT
// is unchanged differs only in the centred 2x2 overlay block (two 6-byte spans at byte offset 3 of the cell's rows 1, 2);
// the cell the step wrote (always the agent's cell) is rewritten in full.  `frame` is 16-byte aligned, a row 12*W bytes.
__device__ __forceinline__ void span6(uint8_t* p, uint32_t rgb) {   // R G B R G B at p, p % 4 == 3
    p[0] = (uint8_t)rgb;
    *reinterpret_cast<uint32_t*>(p + 1) = __byte_perm(rgb, 0, 0x1021);   // G B R G
    p[5] = (uint8_t)(rgb >> 16);
}
__device__ __forceinline__ void frame_edit(const CwConfig& cfg, uint8_t* frame, const uint8_t* g, uint32_t old_agent, uint32_t agent,
                                           int wcell, int wval) {
    const int W = cfg.W;
    const size_t rowb = (size_t)12 * W;
    const int orow = old_agent & 0xFF, ocol = (old_agent >> 8) & 0xFF;
    const int nrow = agent & 0xFF, ncol = (agent >> 8) & 0xFF, hold = (agent >> 16) & 0xFF;
    const int oc = orow * W + ocol, nc = nrow * W + ncol;
    uint8_t* pn = frame + (size_t)(4 * nrow) * rowb + 12 * ncol;
    if (oc != nc) {                                               // the agent left `oc`; the object there did not change
        uint8_t* po = frame + (size_t)(4 * orow) * rowb + 12 * ocol;
        const uint32_t rgb = kColorLUT[g[oc]];
        span6(po + rowb + 3, rgb); span6(po + 2 * rowb + 3, rgb);
    }
    if (wcell == nc) {                                            // the object under the agent changed: all four rows
        const uint32_t rgb = kColorLUT[wval];
        const uint32_t w0 = __byte_perm(rgb, 0, 0x0210), w1 = __byte_perm(rgb, 0, 0x1021), w2 = __byte_perm(rgb, 0, 0x2102);
#pragma unroll
        for (int y = 0; y < 4; y += 3) {                          // rows 0 and 3 in full; rows 1, 2 get their outer pixels
            uint32_t* q = reinterpret_cast<uint32_t*>(pn + y * rowb);
            q[0] = w0; q[1] = w1; q[2] = w2;
        }
#pragma unroll
        for (int y = 1; y < 3; y++) {
            uint8_t* q = pn + y * rowb;
            q[0] = (uint8_t)rgb; q[1] = (uint8_t)(rgb >> 8); q[2] = (uint8_t)(rgb >> 16);
            q[9] = (uint8_t)rgb; q[10] = (uint8_t)(rgb >> 8); q[11] = (uint8_t)(rgb >> 16);
        }
    }
    span6(pn + rowb + 3, 0x00FFFFFFu);                            // ray.py:555
    span6(pn + 2 * rowb + 3, hold ? kColorLUT[hold] : 0x00FFFFFFu);   // ray.py:556-557
}

// ---- pre-drawn reset records of the compact step kernel ---------------------------------------------------------------------
// What a reset draws depends only on (seed, global id, episode), so it can be drawn BEFORE the episode ends.  st.reset_rec keeps
// one 32-byte record per world as two self-validating 16-byte halves, each written and read with ONE 16-byte access:
//     { episode tag, desired mask | agent cell << 16, cells 0|1, cells 2|3 }   { episode tag, cells 4|5, cells 6|7, 0 }
// A finished world's warp then only copies the placement into the grid (reset_apply); the ~2500 dependent instructions of the
// Philox sampling run in extra "refill" CTAs of the NEXT launch, next to -- not in front of -- the stepping warps.  A world
// whose record is missing or stale (a tag of either half differs: first use, a 1-step episode whose successor is still being
// drawn, a reset done elsewhere) falls back to the inline draw; both routes produce the same bits.  st.reset_list is the queue
// between the two: [0] tail (appends so far), [1] limit (tail at the end of the previous launch), [2] head, [3] exit ticket,
// then a ring of 2N world ids.
constexpr int kListHeader = 4;
__device__ __forceinline__ void reset_record_write(uint32_t* rec, uint32_t tag, uint32_t des, const uint32_t (&cells)[9]) {
    const int lane = lane_id();
    uint4 v = lane == 0 ? make_uint4(tag, des | (cells[8] << 16), cells[0] | (cells[1] << 16), cells[2] | (cells[3] << 16))
                        : make_uint4(tag, cells[4] | (cells[5] << 16), cells[6] | (cells[7] << 16), 0u);
    if (lane < 2) __stcg(reinterpret_cast<uint4*>(rec) + lane, v);
}
// refill warps: draw the next reset of every queued world (or of all worlds: prefill)
__device__ __forceinline__ void reset_refill(const CwConfig& cfg, const CwState& st, int64_t first_warp, int64_t n_warps, bool all) {
    uint32_t* list = st.reset_list;
    const uint32_t cap = 2u * (uint32_t)st.n;
    const uint32_t head = all ? 0u : __ldcg(list + 2), limit = all ? (uint32_t)st.n : __ldcg(list + 1);
    for (uint32_t i = head + (uint32_t)first_warp; (int32_t)(limit - i) > 0; i += (uint32_t)n_warps) {
        const int64_t env = all ? (int64_t)i : (int64_t)__ldcg(list + kListHeader + i % cap);
        const uint32_t ep = __ldcg(st.episode + env);
        WarpPhilox rng;
        uint32_t des, cells[9];
        reset_sample(cfg, st, env, rng, ep, des, cells);
        reset_record_write(st.reset_rec + env * 8, ep, des, cells);
    }
}

// one thread per world, K steps per launch (K = 1: cw_step; K > 1: cw_rollout).  `obs` (nullable): the world's device frame is
// kept current by render_edit.  CW_F_DEFER_RESET: a finished world is counted and reported but NOT re-seeded here (the caller
// follows up with a masked cw_reset, which also renders the new episode's frames).  CTAs >= step_blocks are refill CTAs (above).
#ifndef CW_STEP_MINBLOCKS
#define CW_STEP_MINBLOCKS 6   /* <= 80 registers: 512 stepping + up to 376 refill CTAs of a 65536-world launch fit one wave */
#endif
template <bool kEdit>   // (a separate instantiation: the frame patching must not cost the compact path registers)
__global__ void __launch_bounds__(128, CW_STEP_MINBLOCKS) cw_step_kernel(const CwConfig cfg, const CwState st, const uint8_t* __restrict__ actions,
                                                      int32_t* __restrict__ reward, uint8_t* __restrict__ done,
                                                      unsigned long long* stats, uint8_t* obs, uint32_t* list, int K, int flags,
                                                      int step_blocks) {
    CW_WSTAMP(0);
    pdl_launch_dependents();
    pdl_wait();
    CW_WSTAMP(1);
    const bool records = !kEdit && st.reset_rec != nullptr;
    if (!kEdit && (int)blockIdx.x >= step_blocks) {               // refill CTAs: the draws of the resets queued by the previous launch
        reset_refill(cfg, st, (int64_t)(blockIdx.x - step_blocks) * 4 + (threadIdx.x >> 5), (int64_t)(gridDim.x - step_blocks) * 4,
                     step_blocks == 0);
    } else {
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = n < st.n;
    const int64_t nn = valid ? n : 0;
    uint8_t* g = st.grid + nn * cfg.cell_stride;
    const uint8_t* ig = st.init_grid + nn * cfg.cell_stride;
    uint32_t agent = 0, goal = 0, ep = 0;
    int t = 0;
    bool queued = false;                                          // this world is already in the refill queue of this launch
    if (valid) { agent = st.agent[n]; goal = st.goal[n]; t = st.t[n]; ep = ((flags & CW_F_AUTO_RESET) && !(flags & CW_F_DEFER_RESET)) ? st.episode[n] : 0u; }
    for (int k = 0; k < K; k++) {
        bool dn = false;
        if (valid) {
            const int a = actions[(size_t)k * st.n + n];
            int wcell, wval;
            const uint32_t old_agent = agent;
            const int rew = step_core(cfg, g, ig, agent, goal, t, a, dn, wcell, wval);
            if (kEdit && (wcell >= 0 || agent != old_agent))
                frame_edit(cfg, obs + (size_t)n * 48 * cfg.H * cfg.W, g, old_agent, agent, wcell, wval);
            if (reward) reward[(size_t)k * st.n + n] = rew;
            if (done) done[(size_t)k * st.n + n] = dn ? 1 : 0;
            if (dn && (flags & CW_F_AUTO_RESET) && stats) stats_add(cfg, stats + (blockIdx.x % CW_STATS_REPLICAS) * CW_STATS_LEN, goal, t, rew);
        }
        if (kEdit && list && (flags & CW_F_AUTO_RESET)) {          // deferred reset: queue the finished worlds (one atomic per warp)
            const uint32_t m = __ballot_sync(0xffffffffu, valid && dn);
            if (m) {
                uint32_t base = 0;
                if (lane_id() == __ffs(m) - 1) base = atomicAdd(list, (uint32_t)__popc(m));
                base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
                if (valid && dn) list[2 + base + __popc(m & ((1u << lane_id()) - 1u))] = (uint32_t)n;
            }
        }
        if (!kEdit && (flags & CW_F_AUTO_RESET) && !(flags & CW_F_DEFER_RESET)) {   // finished worlds are re-seeded by the whole warp
            uint32_t m = __ballot_sync(0xffffffffu, valid && dn);
            CW_WSTAMP(2);
#ifdef CW_TIMING
            if (g_dbg && (threadIdx.x & 31) == 0) g_dbg[((size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 16 + 9] = __popc(m);
#endif
            uint32_t q = 0, qbase = 0;
            if (records && m) {                                   // queue them for the refill CTAs of the next launch (once per launch);
                q = __ballot_sync(0xffffffffu, valid && dn && !queued);   // the atomic is issued now, its result used after the re-seeding
                if (q && lane_id() == __ffs(q) - 1) qbase = atomicAdd(st.reset_list, (uint32_t)__popc(q));
            }
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int64_t env = __shfl_sync(0xffffffffu, n, src);
                const uint32_t env_ep = __shfl_sync(0xffffffffu, ep, src);
                uint32_t ag, gl;
                bool fast = false;
                if (records) {
                    // lanes 0 / 1 fetch one 16-byte half each (L2); the record is used only if BOTH halves carry this episode's tag
                    const int lane = lane_id();
                    uint4 h = make_uint4(env_ep, 0u, 0u, 0u);
                    if (lane < 2) h = __ldcg(reinterpret_cast<const uint4*>(st.reset_rec + env * 8) + lane);
                    fast = __all_sync(0xffffffffu, h.x == env_ep);
                    if (fast) {
                        const uint32_t a1 = __shfl_sync(0xffffffffu, h.y, 0), a2 = __shfl_sync(0xffffffffu, h.z, 0), a3 = __shfl_sync(0xffffffffu, h.w, 0);
                        const uint32_t b1 = __shfl_sync(0xffffffffu, h.y, 1), b2 = __shfl_sync(0xffffffffu, h.z, 1);
                        const uint32_t cells[9] = {a2 & 0xFFFFu, a2 >> 16, a3 & 0xFFFFu, a3 >> 16, b1 & 0xFFFFu, b1 >> 16, b2 & 0xFFFFu, b2 >> 16, a1 >> 16};
                        reset_apply(cfg, st, env, nullptr, a1 & 0xFFFFu, cells, env_ep, ag, gl);
                    }
                }
                if (!fast) {
                    WarpPhilox rng;
                    reset_warp(cfg, st, env, nullptr, rng, ag, gl, env_ep);
                }
                if (lane_id() == src) { agent = ag; goal = gl; t = 0; ep = env_ep + 1; if (st.init_agent) st.init_agent[env] = ag; }
            }
            if (q) {
                qbase = __shfl_sync(0xffffffffu, qbase, __ffs(q) - 1);
                if (valid && dn && !queued) {
                    st.reset_list[kListHeader + (qbase + __popc(q & ((1u << lane_id()) - 1u))) % (2u * (uint32_t)st.n)] = (uint32_t)n;
                    queued = true;
                }
            }
        }
    }
    CW_WSTAMP(3);
    if (valid) { st.agent[n] = agent; st.goal[n] = goal; st.t[n] = t; }
    CW_WSTAMP(4);
    }
    if (records) {                                                // the last CTA out advances the queue window for the next launch
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            uint32_t* l = st.reset_list;
            if (atomicAdd(l + 3, 1u) == gridDim.x - 1) {
                if (step_blocks == 0) { l[0] = 0; l[1] = 0; l[2] = 0; }          // prefill: every record is fresh, the queue empty
                else { l[2] = l[1]; l[1] = *reinterpret_cast<volatile uint32_t*>(l); }
                l[3] = 0;
            }
        }
    }
}

// Compact step as a member of a CHAIN (cw_step_chained, open-loop tape, one launch per step): an ordinary launch waits for its
// whole predecessor grid -- the slowest warp's re-seed, the refill CTAs -- plus a 2-3 us hand-over (section 3.2), together more
// than twice the 2.6 us of stepping.  A warp steps the same 32 worlds in every launch, so launch i+1 does not wait for grid i: each warp (= CTA: a held-up
// warp then holds nothing but its own slot) waits for ONE word, its own mark of position i (release / acquire).  A finished world
// is re-seeded from its pre-drawn record when there is one (a copy); the ~2500 dependent instructions of the NEXT record's draw
// run AFTER the warp has published its mark, i.e. beside the successor warp's step, not in front of it.  The last CTA out waits
// for the predecessor grid, so grids complete in stream order.
constexpr int kChainStepThreads = 32;
__global__ void __launch_bounds__(kChainStepThreads, 16) cw_step_chained_kernel(const CwConfig cfg, const CwState st, const uint8_t* __restrict__ actions,
                                                                               int32_t* __restrict__ reward, uint8_t* __restrict__ done,
                                                                               unsigned long long* stats, uint32_t* chain, uint32_t cpos,
                                                                               int flags) {
    CW_CSTAMP(0);
    pdl_launch_dependents();
    if (cpos == 0) pdl_wait();                                    // position 0 is an ordinary launch
    uint32_t* const c_fin = chain;
    const int lane = lane_id();
    const int64_t n = (int64_t)blockIdx.x * kChainStepThreads + lane;
    uint32_t* const my_epoch = chain + CW_CHAIN_MAX_POS + blockIdx.x;
    if (cpos > 0) {
        // Launches run ahead of the dependency chain until the SMs are full of waiting warps; a warp whose predecessor is itself
        // still waiting (mark two or more positions behind) polls slowly, only the next in line polls fast.
        if (lane == 0) {
            uint32_t v = ld_acquire_gpu(my_epoch);
            if (v < cpos) {
                const unsigned long long t0 = global_timer_ns();
                for (;;) {
                    __nanosleep(v + 1u >= cpos ? 32 : 800);
                    v = ld_acquire_gpu(my_epoch);
                    if (v >= cpos) break;
                    if (global_timer_ns() - t0 > g_chain_timeout_ns) __trap();
                }
            }
        }
        __syncwarp();
    }
    CW_CSTAMP(1);
    const bool valid = n < st.n;
    const int64_t nn = valid ? n : 0;
    const bool records = st.reset_rec != nullptr && st.n_fixed == 0;
    uint32_t agent = 0, goal = 0, ep = 0;
    int t = 0;
    bool dn = false;
    if (valid) {
        const int a = actions[n];
        agent = __ldcg(st.agent + n); goal = __ldcg(st.goal + n); t = __ldcg(st.t + n);
        ep = (flags & CW_F_AUTO_RESET) ? __ldcg(st.episode + n) : 0u;
        int wcell, wval;
        const int rew = step_core<true>(cfg, st.grid + nn * cfg.cell_stride, st.init_grid + nn * cfg.cell_stride, agent, goal, t, a, dn, wcell, wval);
        if (reward) reward[n] = rew;
        if (done) done[n] = dn ? 1 : 0;
        if (dn && (flags & CW_F_AUTO_RESET)) {
            if (stats) stats_add(cfg, stats + ((blockIdx.x >> 2) % CW_STATS_REPLICAS) * CW_STATS_LEN, goal, t, rew);   // (the replica cw_step uses)
        } else {
            st.agent[n] = agent; st.goal[n] = goal; st.t[n] = t;
        }
    }
    CW_CSTAMP(2);
    const uint32_t fin = (flags & CW_F_AUTO_RESET) ? __ballot_sync(0xffffffffu, valid && dn) : 0u;   // finished worlds, re-seeded by the whole warp
    for (uint32_t m = fin; m;) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int64_t env = __shfl_sync(0xffffffffu, n, src);
        const uint32_t env_ep = __shfl_sync(0xffffffffu, ep, src);
        uint32_t ag, gl;
        bool fast = false;
        if (records) {                                            // (see cw_step_kernel: used only if BOTH halves carry this episode's tag)
            uint4 h = make_uint4(env_ep, 0u, 0u, 0u);
            if (lane < 2) h = __ldcg(reinterpret_cast<const uint4*>(st.reset_rec + env * 8) + lane);
            fast = __all_sync(0xffffffffu, h.x == env_ep);
            if (fast) {
                const uint32_t a1 = __shfl_sync(0xffffffffu, h.y, 0), a2 = __shfl_sync(0xffffffffu, h.z, 0), a3 = __shfl_sync(0xffffffffu, h.w, 0);
                const uint32_t b1 = __shfl_sync(0xffffffffu, h.y, 1), b2 = __shfl_sync(0xffffffffu, h.z, 1);
                const uint32_t cells[9] = {a2 & 0xFFFFu, a2 >> 16, a3 & 0xFFFFu, a3 >> 16, b1 & 0xFFFFu, b1 >> 16, b2 & 0xFFFFu, b2 >> 16, a1 >> 16};
                reset_apply(cfg, st, env, nullptr, a1 & 0xFFFFu, cells, env_ep, ag, gl);
            }
        }
        if (!fast) {
            WarpPhilox rng;
            reset_warp(cfg, st, env, nullptr, rng, ag, gl, env_ep);   // ray.py:156-218
        }
        if (lane == 0) { st.agent[env] = ag; st.goal[env] = gl; st.t[env] = 0; if (st.init_agent) st.init_agent[env] = ag; }
    }
    __syncwarp();                                                 // the lanes' writes are ordered before the mark
    if (lane == 0) st_release_gpu(my_epoch, cpos + 1u);          // (the release is cumulative over the warp barrier: no separate fence)
    CW_CSTAMP(3);
    if (records) {                                                // off the chain: the draws of the re-seeded worlds' NEXT resets
        for (uint32_t m = fin; m;) {
            const int src = __ffs(m) - 1;
            m &= m - 1;
            const int64_t env = __shfl_sync(0xffffffffu, n, src);
            const uint32_t next_ep = __shfl_sync(0xffffffffu, ep, src) + 1u;
            WarpPhilox rng;
            uint32_t des, cells[9];
            reset_sample(cfg, st, env, rng, next_ep, des, cells);
            reset_record_write(st.reset_rec + env * 8, next_ep, des, cells);
        }
    }
    CW_CSTAMP(4);
    if (lane == 0) {                                              // (only finds the last CTA out: nothing is published through this counter)
        const uint32_t before = atomicAdd(c_fin + cpos, 1u);
        if (cpos > 0 && before == gridDim.x - 1) pdl_wait();
    }
    CW_CSTAMP(5);
}

// Delta transport (cw_step_delta): one thread per world, records straight into (possibly host-mapped) memory.  The step of a
// host-buffer call is a synchronous round trip, so what matters here is LATENCY to the records: no tile staging, no shared
// memory, 32 CTAs at 4096 worlds; only the warps that hold a finished world pay for its re-seed (warp-cooperative Philox
// reset + closed-form imagine_obs), every other record is on its way ~1 us after the actions arrive.
constexpr int kParamActions = 4096;               // a batch up to this size can carry its actions in the kernel parameters
struct ActionBlock { uint8_t a[kParamActions]; };
template <bool kInParams>                         // kInParams: the actions arrive WITH the launch (no PCIe read of mapped host memory)
__global__ void __launch_bounds__(128) cw_delta_kernel(const CwConfig cfg, const CwState st, const uint8_t* __restrict__ actions,
                                                       const __grid_constant__ ActionBlock pa, uint4* __restrict__ delta,
                                                       uint32_t* __restrict__ fresh, unsigned long long* stats, uint32_t seq,
                                                       int flags) {
    __shared__ uint32_t s_obj[4][8];
    pdl_launch_dependents();
    pdl_wait();
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = n < st.n;
    const int64_t nn = valid ? n : 0;
    uint8_t* g = st.grid + nn * cfg.cell_stride;
    uint32_t agent = 0, goal = 0, ep = 0, rew_u = 0;
    int t = 0;
    bool dn = false;
    if (valid) {
        const int a = kInParams ? pa.a[n] : actions[n];           // (host-mapped memory: the longest latency, issued first)
        agent = __ldcg(st.agent + n); goal = __ldcg(st.goal + n); t = __ldcg(st.t + n);
        ep = (flags & CW_F_AUTO_RESET) ? __ldcg(st.episode + n) : 0u;
        int wcell, wval, ocode, ncode;
        const uint32_t opos = (agent & 0x3Fu) | (((agent >> 8) & 0x3Fu) << 6);   // where the agent stood (row | col << 6)
        const int rew = step_core_ex<true>(cfg, g, st.init_grid + nn * cfg.cell_stride, agent, goal, t, a, dn, wcell, wval, ocode, ncode);
        rew_u = (uint32_t)rew;
        if (dn && (flags & CW_F_AUTO_RESET)) {
            if (stats) stats_add(cfg, stats + (blockIdx.x % CW_STATS_REPLICAS) * CW_STATS_LEN, goal, t, rew);
        } else {
            st.agent[n] = agent; st.goal[n] = goal; st.t[n] = t;
            // the record is pre-digested: the consumer repaints from it alone (no copy of the grid on its side)
            delta[n] = make_uint4(agent, goal, opos | ((uint32_t)ocode << 12) | ((uint32_t)ncode << 16) | (wcell >= 0 ? 1u << 20 : 0u) |
                                  (((dn ? 1u : 0u) | (seq << 2)) << 24), rew_u);
        }
    }
    if (!(flags & CW_F_AUTO_RESET)) return;
    uint32_t m = __ballot_sync(0xffffffffu, valid && dn);        // finished worlds of this warp, re-seeded by the whole warp
    const int lane = lane_id();
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int64_t env = __shfl_sync(0xffffffffu, n, src);
        const uint32_t env_ep = __shfl_sync(0xffffffffu, ep, src), env_rew = __shfl_sync(0xffffffffu, rew_u, src);
        WarpPhilox rng;
        Sparse8 objs;
        uint32_t ag, gl;
        reset_warp(cfg, st, env, nullptr, rng, ag, gl, env_ep, &objs, s_obj[threadIdx.x >> 5]);   // ray.py:156-218
        if (lane == 0) { st.agent[env] = ag; st.goal[env] = gl; st.t[env] = 0; if (st.init_agent) st.init_agent[env] = ag; }
        uint32_t* fr = fresh + (size_t)env * CW_FRESH_WORDS;
        uint32_t word = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) word = lane == k ? (objs.cell[k] | (objs.code[k] << 16)) : word;
        if (lane < 8) fr[lane] = word;
        uint32_t gag = ag;
        imagine_fresh(cfg, objs, gag, gl >> 16, rng);             // desired_goal = imagine_obs(): ray.py:191, 220-299
        word = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) word = lane == k ? (objs.cell[k] | (objs.code[k] << 16)) : word;
        if (lane < 8) fr[8 + lane] = word;
        if (lane == 8) fr[16] = gag;
        __threadfence_system();                                   // the 16-byte record goes last (see cw_env_kernel)
        __syncwarp();
        if (lane == 0) delta[env] = make_uint4(ag, gl, (3u /* done | fresh */ | (seq << 2)) << 24, env_rew);
    }
}

// Pipelined host transport (cw_host.cu, device consumer, small batches): the STEP half.  A host-driven closed loop needs
// reward / done of step k before it can issue step k+1 -- a latency problem -- while the frames are a bandwidth problem; a launch
// that does both only reports the last world when its last CTA has found an SM slot, i.e. when the previous launch's frames are
// out.  So the step runs here, one thread per world on its own stream: status byte to the host first, then the state that the
// render launch of this step (cw_env_kernel<V_PIPE>, another stream) needs is copied into a SNAPSHOT slot -- the grid rows of the
// warp's 32 worlds are one contiguous block -- and published per CTA with a release.  Live state belongs to this kernel alone,
// so step k+1 never waits for frames; a slot is reused only after the render launch that read it has finished (`slot_free`).
// One WARP per CTA: a step CTA must find room on an SM that is full of render CTAs which may be spinning on its result (4 x 160
// threads x <= 96 registers leave 4096 registers: 32 threads x <= 128).  cw_host.cu keeps the two launch configurations in step.
constexpr int kSnapThreads = 32;
template <bool kInParams>
__global__ void __launch_bounds__(kSnapThreads, 16) cw_step_snap_kernel(const CwConfig cfg, const CwState st, const uint8_t* __restrict__ actions,
                                                                        const __grid_constant__ ActionBlock pa, uint8_t* __restrict__ status,
                                                                        const PipeSnap snap, uint32_t* __restrict__ epoch, uint32_t seq,
                                                                        const uint32_t* slot_free, uint32_t slot_want,
                                                                        unsigned long long* stats, int flags) {
    __shared__ uint32_t s_obj[8];
    // Consecutive step launches are linked by dataflow too: a dependent launch would wait for its predecessor's COMPLETION, and
    // this kernel's tail -- re-seeds, the snapshot copy, the grid hand-over (section 3.2) -- is longer than its step.  A warp steps
    // the same 32 worlds in every launch, so it waits for ONE word: its own mark of the step before.
    CW_SSTAMP(0);
    pdl_launch_dependents();
    const int lane = lane_id();
    const int64_t n = (int64_t)blockIdx.x * kSnapThreads + lane;
    const bool valid = n < st.n;
    const int64_t nn = valid ? n : 0;
    const int cs = cfg.cell_stride;
    uint32_t* const my_epoch = epoch + blockIdx.x;
    if (lane == 0) chain_wait_seq(my_epoch, seq - 1u);           // (acquire: also drops this SM's stale L1 lines)
    __syncwarp();
    CW_SSTAMP(1);
    uint32_t agent = 0, goal = 0, ep = 0, gag = 0, mflags = 0;
    int t = 0;
    bool dn = false;
    if (valid) {
        const int a = kInParams ? pa.a[n] : actions[n];
        agent = __ldcg(st.agent + n); goal = __ldcg(st.goal + n); t = __ldcg(st.t + n);
        ep = (flags & CW_F_AUTO_RESET) ? __ldcg(st.episode + n) : 0u;
        int wcell, wval;
        const int rew = step_core<true>(cfg, st.grid + nn * cs, st.init_grid + nn * cs, agent, goal, t, a, dn, wcell, wval);
        CW_SSTAMP(2);
        // (st.wt: through the L2 to system memory now, see cw_env_kernel)
        __stwt(status + n, (uint8_t)(0x80u | (rew == cfg.max_steps ? 2u : 0u) | (dn ? 1u : 0u)));
        if (dn && (flags & CW_F_AUTO_RESET)) {
            if (stats) stats_add(cfg, stats + (blockIdx.x % CW_STATS_REPLICAS) * CW_STATS_LEN, goal, t, rew);
        } else {
            st.agent[n] = agent; st.goal[n] = goal; st.t[n] = t;
        }
    }
    uint32_t m = (flags & CW_F_AUTO_RESET) ? __ballot_sync(0xffffffffu, valid && dn) : 0u;   // finished worlds, re-seeded by the whole warp
    // the slot's previous reader (the render launch kPipeSlots steps back) must be done with it before anything is written there
    CW_SSTAMP(3);
    if (slot_want) {
        if (lane == 0) chain_wait_seq(slot_free, slot_want);
        __syncwarp();
    }
    CW_SSTAMP(4);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const int64_t env = __shfl_sync(0xffffffffu, n, src);
        const uint32_t env_ep = __shfl_sync(0xffffffffu, ep, src);
        WarpPhilox rng;
        Sparse8 objs;
        uint32_t ag, gl;
        reset_warp(cfg, st, env, nullptr, rng, ag, gl, env_ep, &objs, s_obj);   // ray.py:156-218
        if (lane == 0) { st.agent[env] = ag; st.goal[env] = gl; st.t[env] = 0; if (st.init_agent) st.init_agent[env] = ag; }
        uint32_t g2 = ag;
        imagine_fresh(cfg, objs, g2, gl >> 16, rng);              // desired_goal = imagine_obs(): ray.py:191, 220-299
        tile_from_objects(objs, cs >> 4, snap.goal + env * cs);
        if (lane == src) { agent = ag; gag = g2; mflags = 1u | ((uint32_t)min(t, 255) << 8); }
    }
    if (valid) snap.meta[n] = make_uint4(agent, gag, mflags, 0u);
    __syncwarp();                                                 // the lanes' writes (step cell, re-seeded rows) are ordered before the copy
    {                                                             // the grid rows of the warp's 32 worlds are one contiguous block
        const int64_t w0 = n - lane;
        const int cnt = (int)max((int64_t)0, min((int64_t)32, st.n - w0));
        const uint4* src = reinterpret_cast<const uint4*>(st.grid + w0 * cs);
        uint4* dst = reinterpret_cast<uint4*>(snap.grid + w0 * cs);
        const int nq = cnt * (cs >> 4);
        for (int i0 = lane; i0 < nq; i0 += 32 * 8) {              // eight 16-byte loads in flight per lane (a load-store loop pays
            uint4 v[8];                                           //  one L2 round trip per iteration: 5.6 us for 28 of them, measured)
#pragma unroll
            for (int u = 0; u < 8; u++) if (i0 + 32 * u < nq) v[u] = __ldcg(src + i0 + 32 * u);
#pragma unroll
            for (int u = 0; u < 8; u++) if (i0 + 32 * u < nq) dst[i0 + 32 * u] = v[u];
        }
    }
    // live state final, copied (the next step will write into these rows) and the snapshot of these 32 worlds complete: ONE mark
    // for both waiters, the same warp of the next step launch and the render launch of this step
    __syncwarp();
    CW_SSTAMP(5);
    if (lane == 0) st_release_gpu(my_epoch, seq);                 // (the release is cumulative over the warp barrier: no separate fence)
    CW_SSTAMP(6);
    pdl_wait();                                                   // grids still complete in stream order (a stream sync means what it says)
    CW_SSTAMP(7);
}

// ---- observation-format expanders (one-hot state, AltObs frames): staged in shared memory, streamed out by TMA ------
// Both outputs are contiguous over (world, ...), so a work item is a contiguous BYTE RANGE of the output: it is composed
// in shared memory at the same address phase (mod 16) as its destination; the 16-byte aligned body leaves with ONE TMA
// bulk store (no LSU store instructions), the few head / tail bytes of a range that does not start / end on a 16-byte
// boundary with element-sized stores.  Two stages: composing item i+1 overlaps the store of item i.  The global write
// stream is therefore full-line whatever H, W and the world count are.
constexpr int kExpThreads = 256;
constexpr int kOneHotIters = 7;
constexpr uint32_t kOneHotCells = kOneHotIters * kExpThreads;    // cells per work item (21 KB of output)
constexpr uint32_t kOneHotStage = kOneHotCells * 12 + 64;   // + phase + the unused cells of a last partial quad
constexpr uint32_t kAltBudget = 32 * 1024;       // bytes of AltObs output per work item
constexpr uint32_t kAltStage = kAltBudget + 48;
constexpr int kAltIters = 3;                     // cells per item <= kAltBudget / 54 = 606 <= 3 x 256

template <int kGran>   // element size of head / tail copies: 2 or 4 bytes (s and g have the same address modulo 16)
__device__ __forceinline__ void stream_out_same_phase(const uint8_t* s, uint8_t* g, uint32_t nbytes, int tid) {
    const uint32_t head = min((16u - (uint32_t)((uintptr_t)g & 15u)) & 15u, nbytes);
    const uint32_t body = (nbytes - head) & ~15u;
    for (uint32_t i = tid * kGran; i < head; i += kExpThreads * kGran)
        for (int q = 0; q < kGran; q++) g[i + q] = s[i + q];
    for (uint32_t i = head + body + tid * kGran; i < nbytes; i += kExpThreads * kGran)
        for (int q = 0; q < kGran; q++) g[i + q] = s[i + q];
    if (tid == 0) {
        if (body) bulk_store(g + head, s + head, body);
        bulk_commit();                                            // (an empty group keeps the stage accounting uniform)
    }
}

// observation_vector (ray.py:94-98, 605-613): uint8[N][H][W][12].  A work item is kOneHotCells consecutive cells of the
// flattened (world, cell) index; a thread owns QUADS of 4 consecutive cells = 48 output bytes = three 16-byte words built in
// registers (one index division per quad; a quad crosses at most one world boundary because H*W >= 4).
__global__ void __launch_bounds__(kExpThreads) cw_onehot_kernel(const CwConfig cfg, const uint8_t* __restrict__ grid,
                                                                const uint32_t* __restrict__ agent, uint8_t* __restrict__ out,
                                                                int64_t n_cells, int64_t n, uint32_t hw_magic) {
    extern __shared__ __align__(16) uint8_t xsm[];
    const uint32_t HW = (uint32_t)(cfg.H * cfg.W), W = (uint32_t)cfg.W;
    const int tid = threadIdx.x;
    const int64_t items = (n_cells + kOneHotCells - 1) / kOneHotCells;
    constexpr int kQuadIters = (kOneHotCells / 4 + kExpThreads - 1) / kExpThreads;   // 2
    struct Loaded { uint32_t code[kQuadIters][4], agA[kQuadIters], agB[kQuadIters], cell0[kQuadIters]; };
    // (world, cell) of the first cell of this CTA's current item, advanced by a constant stride per iteration: the only
    // 64-bit divisions of the kernel happen here, once
    const int64_t c_first = (int64_t)blockIdx.x * kOneHotCells, c_stride = (int64_t)gridDim.x * kOneHotCells;
    int64_t ld_env = c_first / HW;
    uint32_t ld_rem = (uint32_t)(c_first - ld_env * HW);
    const int64_t step_env = c_stride / HW;
    const uint32_t step_rem = (uint32_t)(c_stride - step_env * HW);
    const uint32_t wrap_off = (uint32_t)cfg.cell_stride - HW;    // byte offset correction for a cell that belongs to the next world
    auto load_item = [&](int64_t it, Loaded& L) {                // all loads of an item in flight together; advances (ld_env, ld_rem)
        const int64_t c0 = it * kOneHotCells;
        const uint32_t cnt = (uint32_t)min((int64_t)kOneHotCells, n_cells - c0);
        const uint32_t nquad = (cnt + 3u) >> 2;
#pragma unroll
        for (int u = 0; u < kQuadIters; u++) {
            const uint32_t q = tid + u * kExpThreads;
            const uint32_t j0 = q < nquad ? 4u * q : 0u;
            const uint32_t idx = ld_rem + j0;
            const uint32_t de = __umulhi(idx, hw_magic);          // idx / HW  (idx < kOneHotCells + HW)
            const int64_t env = ld_env + de;
            const uint32_t c = idx - de * HW;
            L.cell0[u] = c;
            const uint8_t* gp = grid + env * cfg.cell_stride + c;
#pragma unroll
            for (int i = 0; i < 4; i++)
                L.code[u][i] = (j0 + i < cnt) ? gp[i + ((c + i >= HW) ? wrap_off : 0u)] : 0u;
            L.agA[u] = agent[env];
            L.agB[u] = agent[env + 1 < n ? env + 1 : env];
        }
        ld_env += step_env; ld_rem += step_rem;
        if (ld_rem >= HW) { ld_rem -= HW; ld_env += 1; }
    };
    int stage = 0;
    Loaded cur, nxt;
    if ((int64_t)blockIdx.x < items) load_item(blockIdx.x, cur);
    for (int64_t it = blockIdx.x; it < items; it += gridDim.x, stage ^= 1) {
        const int64_t c0 = it * kOneHotCells;
        const uint32_t cnt = (uint32_t)min((int64_t)kOneHotCells, n_cells - c0);
        const uint32_t nquad = (cnt + 3u) >> 2;
        uint8_t* dst = out + c0 * 12;
        const uint32_t phase = (uint32_t)((uintptr_t)dst & 15u);  // 0, 4, 8 or 12 (c0 * 12 is a multiple of 16)
        uint8_t* buf = xsm + stage * kOneHotStage + phase;
        if (it + gridDim.x < items) load_item(it + gridDim.x, nxt);   // next item's loads fly while this one is composed
        if (tid == 0) bulk_wait_read<1>();                        // the store that last read this stage has drained
        __syncthreads();
#pragma unroll
        for (int u = 0; u < kQuadIters; u++) {
            const uint32_t q = tid + u * kExpThreads;
            if (q >= nquad) continue;
            // agent words of the quad's world (A) and of the next world (B, for cells past the world boundary): position of
            // the agent cell relative to the quad, and channels 8..11 (agent; holding sticks / axe / hammer) as one word
            const uint32_t a = cur.agA[u], bb = cur.agB[u], c = cur.cell0[u];
            const uint32_t ia = (a & 0xFF) * W + ((a >> 8) & 0xFF) - c;               // 0..3 if the agent of A is in the quad
            const uint32_t ib = (bb & 0xFF) * W + ((bb >> 8) & 0xFF) + HW - c;        // same for B (cells c+i >= HW)
            const uint32_t ha = (a >> 16) & 0xFF, hb = (bb >> 16) & 0xFF;
            const uint32_t wa = 1u | ((ha >= 1 && ha <= 3) ? 1u << (8 * ha) : 0u), wb = 1u | ((hb >= 1 && hb <= 3) ? 1u << (8 * hb) : 0u);
            const uint32_t nA = HW - c;                                               // cells i < nA belong to world A
            uint32_t w[12];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const uint32_t cd = cur.code[u][i];
                const unsigned long long m = (cd >= 1 && cd <= 8) ? 1ull << (8 * (cd - 1u)) : 0ull;   // object channels 0..7
                w[3 * i + 0] = (uint32_t)m;
                w[3 * i + 1] = (uint32_t)(m >> 32);
                w[3 * i + 2] = ((uint32_t)i < nA) ? ((uint32_t)i == ia ? wa : 0u) : ((uint32_t)i == ib ? wb : 0u);
            }
            uint8_t* p = buf + 48u * q;
            if (phase == 0) {
                reinterpret_cast<uint4*>(p)[0] = make_uint4(w[0], w[1], w[2], w[3]);
                reinterpret_cast<uint4*>(p)[1] = make_uint4(w[4], w[5], w[6], w[7]);
                reinterpret_cast<uint4*>(p)[2] = make_uint4(w[8], w[9], w[10], w[11]);
            } else {
#pragma unroll
                for (int k = 0; k < 12; k++) reinterpret_cast<uint32_t*>(p)[k] = w[k];
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        stream_out_same_phase<4>(buf, dst, cnt * 12, tid);
        cur = nxt;
    }
    if (tid == 0) bulk_wait_all();
}
// fallback for an output pointer that is not 4-byte aligned: one byte-wise cell per thread
__global__ void __launch_bounds__(256) cw_onehot_unaligned_kernel(const CwConfig cfg, const uint8_t* __restrict__ grid,
                                                                  const uint32_t* __restrict__ agent, uint8_t* __restrict__ out,
                                                                  int64_t n_cells) {
    const int HW = cfg.H * cfg.W;
    for (int64_t fc = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; fc < n_cells; fc += (int64_t)gridDim.x * blockDim.x) {
        const int64_t env = fc / HW;
        const int cell = (int)(fc - env * HW);
        const int code = grid[env * cfg.cell_stride + cell];
        const uint32_t ag = agent[env];
        const int acell = (int)(ag & 0xFF) * cfg.W + (int)((ag >> 8) & 0xFF), h = (ag >> 16) & 0xFF;
        uint8_t* o = out + fc * 12;
        for (int ch = 0; ch < 12; ch++)
            o[ch] = (uint8_t)((ch < 8 && code == ch + 1) || (cell == acell && (ch == 8 || (h >= 1 && h <= 3 && ch == 8 + h))));
    }
}

// AltObs renderer (craftingworld_altobs.py:489-548): int16[N][3H+3][3W][3].  Sub-pixel k of a cell is lit with
// CPV_COLORS[k] x the multiplicity of channel k, so a frame is almost all zeros: a work item (several whole worlds, or a
// band of cell rows of one large world) is zero-filled in shared memory with 16-byte stores, the <= 3 lit sub-pixels of
// each cell are scattered into it, and the range is streamed out.
__device__ __constant__ int16_t kCPV[9][3] = {{45, 82, 160},  {255, 102, 102}, {204, 204, 0},   {211, 211, 211}, {34, 133, 34},
                                              {0, 215, 255},  {153, 52, 255},  {10, 215, 100},  {0, 0, 255}};   // altobs.py:26-27
struct AltPlan { int worlds_per_item, bands, rows_per_band; int64_t items; uint32_t w_magic, cells_magic; };
__global__ void __launch_bounds__(kExpThreads) cw_render_alt_kernel(const CwConfig cfg, const uint8_t* __restrict__ grid,
                                                                    const uint32_t* __restrict__ agent, int16_t* __restrict__ out,
                                                                    int64_t n, const AltPlan plan) {
    extern __shared__ __align__(16) uint8_t xsm[];
    const int H = cfg.H, W = cfg.W, PW = 3 * W;
    const uint32_t P = (uint32_t)(3 * H + 3) * PW;                // pixels per frame
    const int tid = threadIdx.x;
    int stage = 0;
    for (int64_t it = blockIdx.x; it < plan.items; it += gridDim.x, stage ^= 1) {
        int64_t e0;
        int nw, r0, r1;
        bool last;
        if (plan.bands == 1) { e0 = it * plan.worlds_per_item; nw = (int)min((int64_t)plan.worlds_per_item, n - e0); r0 = 0; r1 = H; last = true; }
        else {
            e0 = it / plan.bands; nw = 1;
            const int b = (int)(it - e0 * plan.bands);
            r0 = b * plan.rows_per_band; r1 = min(H, r0 + plan.rows_per_band); last = b == plan.bands - 1;
        }
        const uint32_t rows = (uint32_t)(r1 - r0);
        const uint32_t item_pixels = nw > 1 ? (uint32_t)nw * P : (3u * rows + (last ? 3u : 0u)) * PW;
        const uint32_t nbytes = item_pixels * 6;
        uint8_t* dst = reinterpret_cast<uint8_t*>(out) + ((size_t)e0 * P + (size_t)(3 * r0) * PW) * 6;
        uint8_t* sbase = xsm + stage * kAltStage;
        uint8_t* buf = sbase + ((uintptr_t)dst & 15u);            // even: int16 output
        const uint32_t band_cells = rows * (uint32_t)W, ncell = (uint32_t)nw * band_cells;
        uint32_t code[kAltIters], ag[kAltIters], wi[kAltIters], lr[kAltIters], cc[kAltIters];
#pragma unroll
        for (int u = 0; u < kAltIters; u++) {                     // all loads of the item in flight together
            const uint32_t j = tid + u * kExpThreads, jj = j < ncell ? j : 0u;
            wi[u] = __umulhi(jj, plan.cells_magic);               // j / band_cells
            const uint32_t lc = jj - wi[u] * band_cells;          // cell within the band
            lr[u] = __umulhi(lc, plan.w_magic);                   // row within the band
            cc[u] = lc - lr[u] * (uint32_t)W;
            const int64_t env = e0 + wi[u];
            code[u] = grid[env * cfg.cell_stride + (uint32_t)(r0 + lr[u]) * (uint32_t)W + cc[u]];
            ag[u] = agent[env];
        }
        if (tid == 0) bulk_wait_read<1>();                        // the store that last read this stage has drained
        __syncthreads();
        for (uint32_t i = tid; i < ((nbytes + 30u) >> 4); i += kExpThreads) reinterpret_cast<uint4*>(sbase)[i] = make_uint4(0, 0, 0, 0);
        __syncthreads();
        int16_t* px = reinterpret_cast<int16_t*>(buf);
#pragma unroll
        for (int u = 0; u < kAltIters; u++) {
            if (tid + u * kExpThreads >= ncell) continue;
            const uint32_t cell = (uint32_t)(r0 + lr[u]) * (uint32_t)W + cc[u];
            const bool here = cell == (ag[u] & 0xFF) * (uint32_t)W + ((ag[u] >> 8) & 0xFF);
            const uint32_t h = here ? ((ag[u] >> 16) & 0xFF) : 0u, cd = code[u];
            int16_t* cellp = px + ((size_t)wi[u] * P + (size_t)(3 * lr[u]) * PW + 3 * cc[u]) * 3;
            auto light = [&](uint32_t k, int mult) {              // sub-pixel k = (k / 3, k % 3) of the cell, altobs.py:45-51
                int16_t* q = cellp + ((k / 3) * PW + (k % 3)) * 3;
                q[0] = (int16_t)(mult * kCPV[k][0]); q[1] = (int16_t)(mult * kCPV[k][1]); q[2] = (int16_t)(mult * kCPV[k][2]);
            };
            const bool held = h >= 1 && h <= 3;                   // a held item adds onto channels 0..2, altobs.py:531-533
            if (cd >= 1 && cd <= 8) light(cd - 1, 1 + ((held && h == cd) ? 1 : 0));
            if (here) {
                light(8, 1);                                      // agent channel
                if (held && h != cd) light(h - 1, 1);
            }
        }
        if (last) {                                               // status strip, altobs.py:542-545: columns 3..5 white while holding
            for (uint32_t j = tid; j < (uint32_t)nw * 27u; j += kExpThreads) {
                const uint32_t w = j / 27u, q = j - w * 27u, y = q / 9u, xq = q - y * 9u;
                if (((agent[e0 + w] >> 16) & 0xFF) != 0)
                    px[((size_t)w * P + (size_t)(3 * rows + y) * PW + 3) * 3 + xq] = 255;
            }
        }
        fence_proxy_async_smem();
        __syncthreads();
        stream_out_same_phase<2>(buf, dst, nbytes, tid);
    }
    if (tid == 0) bulk_wait_all();
}

// A stand-in DEVICE CONSUMER of the frames (closed-loop measurements and tests): reads every byte of every world's frame and
// turns it into that world's next action, so step k+1 depends on frame k the way it does under a policy network.
//   h = sum_i word_i * (2 i + 1)  (uint32 words of the frame, wrap-around),  action = ((h ^ h >> 16) & 0xFFFF) % 6
// One CTA per world (grid-stride), 16-byte streaming loads, warp-shuffle + shared-memory reduction.  Bound: HBM read.
constexpr int kPolicyThreads = 256, kPolicyLoads = 6;    // 256 x 6 x 16 B = one 21x21 frame (21 168 B) in ONE round of loads
__global__ void __launch_bounds__(kPolicyThreads) cw_frame_policy_kernel(const uint4* __restrict__ obs, int64_t n, uint32_t words16,
                                                                         uint8_t* __restrict__ actions) {
    __shared__ uint32_t s_part[kPolicyThreads / 32];
    pdl_launch_dependents();
    pdl_wait();
    for (int64_t w = blockIdx.x; w < n; w += gridDim.x) {
        const uint4* f = obs + (size_t)w * words16;
        uint32_t h = 0;
        for (uint32_t i0 = threadIdx.x; i0 < words16; i0 += kPolicyLoads * kPolicyThreads) {   // independent 16-byte loads in flight per thread
            uint4 v[kPolicyLoads];
#pragma unroll
            for (int u = 0; u < kPolicyLoads; u++)
                v[u] = (i0 + (uint32_t)(kPolicyThreads * u) < words16) ? __ldcs(f + i0 + kPolicyThreads * u) : make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
            for (int u = 0; u < kPolicyLoads; u++) {
                const uint32_t k = 8u * (i0 + (uint32_t)(kPolicyThreads * u)) + 1u;     // 2 * (4 i) + 1
                h += v[u].x * k + v[u].y * (k + 2u) + v[u].z * (k + 4u) + v[u].w * (k + 6u);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
        if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = h;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int q = 0; q < kPolicyThreads / 32; q++) t += s_part[q];
            actions[w] = (uint8_t)(((t ^ (t >> 16)) & 0xFFFFu) % 6u);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------------
struct OccEntry { size_t smem; int per_sm; };
struct KernelInfo { bool attr_set = false; int max_dyn = 0; int n_occ = 0; OccEntry occ[64]; };   // per kernel instantiation
struct DeviceInfo { int sms = 0; int smem_optin = 0; bool ok = false; bool alt_attr_set = false; KernelInfo k[4]; };
static DeviceInfo g_dev[64];
static std::mutex g_dev_mu;   // guards the per-device attribute / occupancy cache (entry points may be called from several host threads)

static int device_info(DeviceInfo** out) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    if (dev < 0 || dev >= 64) return CW_E_BADCONFIG;
    DeviceInfo& d = g_dev[dev];
    std::lock_guard<std::mutex> lk(g_dev_mu);
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
    if (cfg->H < 2 || cfg->W < 2 || cfg->H > CW_MAX_SIDE || cfg->W > CW_MAX_SIDE) return CW_E_BADCONFIG;   // (the reciprocal-multiply divisions need a divisor >= 2)
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

// Launch with the programmatic-stream-serialization attribute (PDL): back-to-back launches of the step graph
// overlap the next kernel's launch + prologue with this kernel's tail.  CW_PDL=0 disables it.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args... args) {
    static int use_pdl = -1;
    if (use_pdl < 0) { const char* e = getenv("CW_PDL"); use_pdl = (e && *e == '0') ? 0 : 1; }
    cudaLaunchConfig_t lc = {};
    lc.gridDim = grid; lc.blockDim = block; lc.dynamicSmemBytes = smem; lc.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    lc.attrs = attr; lc.numAttrs = use_pdl ? 1 : 0;
    return cudaLaunchKernelEx(&lc, kern, static_cast<KArgs>(args)...);
}

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return (s && *s) ? atoi(s) : dflt;
}
// experiment knobs (tools/sweep*.sh), read from the environment ONCE: a host-buffer step is tens of microseconds and six
// getenv scans per launch were a measurable part of it
struct Tunables {
    int bands_per_chunk = env_int("CW_BANDS_PER_CHUNK", 0), chunk_bytes = env_int("CW_CHUNK_BYTES", 25 * 1024);
    int frame_buffers = env_int("CW_FRAME_BUFFERS", 0), first_split = env_int("CW_FIRST_SPLIT", 4);
    int ctas_per_sm = env_int("CW_CTAS_PER_SM", 0), group = env_int("CW_GROUP", 0);
    int refill_ctas = env_int("CW_REFILL_CTAS", 0);             // experiment: cap on the refill CTAs of the compact step launch
    int no_chain = env_int("CW_NO_CHAIN", 0);                   // 1: cw_step_render_chained degrades to ordinary launches
    int chain_timeout_ms = env_int("CW_CHAIN_TIMEOUT_MS", 0);   // > 0: spin limit of the chain waits
};
static const Tunables& tunables() {
    static const Tunables t;
    return t;
}

// frame chunking: the largest number of bands whose chunk fits the per-buffer budget
static int pick_bands(const CwConfig* cfg, int budget_bytes) {
    int forced = tunables().bands_per_chunk;
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
    const int variant = args.pepoch ? V_PIPE : (args.chain ? V_CHAINED : (args.list ? V_LIST : V_PLAIN));
    auto kern = variant == V_PIPE ? cw_env_kernel<V_PIPE>
                                  : (variant == V_CHAINED ? cw_env_kernel<V_CHAINED> : (variant == V_LIST ? cw_env_kernel<V_LIST> : cw_env_kernel<V_PLAIN>));
    KernelInfo* ki = &dev->k[variant];
    std::unique_lock<std::mutex> lk(g_dev_mu);
    if (!ki->attr_set) {   // once per device: allow any dynamic size up to the opt-in maximum, prefer shared memory
        cudaFuncAttributes fa;
        cudaError_t e = cudaFuncGetAttributes(&fa, kern);
        if (e != cudaSuccess) return (int)e;
        ki->max_dyn = dev->smem_optin - (int)fa.sharedSizeBytes;   // the opt-in limit covers static + dynamic
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ki->max_dyn);
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) return (int)e;
        ki->attr_set = true;
    }
    const bool needs_frame = (args.mode & M_RENDER) || args.goal_obs;
    args.bands_per_chunk = needs_frame ? pick_bands(cfg, tunables().chunk_bytes) : 1;
    // Ring depth F.  A launch of one or a few CTA waves (small batches) is fastest with F = 2 and 4 CTAs/SM; a persistent
    // launch over many groups wants a deeper ring and larger groups (fewer CTAs/SM, the per-group step phase amortised
    // over more worlds): measured on B200 at 21x21 (tools/sweep3.sh) F = 3 wins around 0.7 GB of frames per launch and
    // F = 4 from ~1.3 GB up (0.90 -> 0.95 of the HBM roofline at 131072 worlds).  Multi-chunk frames keep F = 2.
    int F = tunables().frame_buffers;
    if (F <= 0) {
        const double frame_total = (double)st->n * 48.0 * cfg->H * cfg->W;
        const bool single_chunk = args.bands_per_chunk >= cfg->H;
        F = (!needs_frame || !single_chunk || args.list) ? 2 : (frame_total >= 1.2e9 ? 4 : (frame_total >= 0.6e9 ? 3 : 2));
    }
    F = F < 2 ? 2 : (F > 6 ? 6 : F);
    args.nbuf = F;
    args.w_magic = (uint32_t)(0x100000000ull / (uint64_t)cfg->W) + 1u;
    args.first_split = tunables().first_split;
    if (args.first_split < 1) args.first_split = 1;
    const size_t ring = needs_frame ? (size_t)F * 48 * cfg->W * args.bands_per_chunk : 0;
    const int cap = tunables().ctas_per_sm > 0 ? tunables().ctas_per_sm : args.ctas_cap;
    // Pipelined launches spin on words that a step launch on ANOTHER stream publishes: that launch must always find room, whatever
    // mix of render launches is resident.  The register file guarantees it: at most 4 render CTAs per SM (launch bounds; 5 warps x
    // <= 96 registers each -- with more per thread only three fit) leave >= 4096 registers, one 32-thread step CTA of <= 128.
    auto smem_for = [&](int G) { return 3 * (size_t)G * cfg->cell_stride + ring; };
    // group size G: the per-SM critical path is (CTA waves) x G worlds; pick the G that minimises it
    int bestG = 0, best_per_sm = 1;
    int64_t best_cost = 0;
    const int forcedG = args.list ? 1 : tunables().group;
    const int gmax = 16384 / cfg->cell_stride < 1 ? 1 : (16384 / cfg->cell_stride > 16 ? 16 : 16384 / cfg->cell_stride);
    for (int G = (forcedG > 0 ? forcedG : 1); G <= (forcedG > 0 ? forcedG : gmax); G++) {
        if (G > 32) break;
        const size_t smem = smem_for(G);
        if (smem > (size_t)ki->max_dyn) break;
        int per_sm = 0;
        for (int i = 0; i < ki->n_occ; i++)
            if (ki->occ[i].smem == smem) per_sm = ki->occ[i].per_sm;
        if (per_sm == 0) {
            cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kEnvThreads, smem);
            if (e != cudaSuccess) return (int)e;
            if (per_sm < 1) per_sm = 1;
            if (ki->n_occ < 64) { ki->occ[ki->n_occ].smem = smem; ki->occ[ki->n_occ].per_sm = per_sm; ki->n_occ++; }
        }
        if (cap > 0 && cap < per_sm) per_sm = cap;
        const int64_t slots = (int64_t)dev->sms * per_sm;
        const int64_t groups = (st->n + G - 1) / G;
        const int64_t waves = (groups + slots - 1) / slots;
        const int64_t cost = waves * G + waves;                  // + a fixed per-wave cost (load / step latency)
        if (bestG == 0 || cost < best_cost || (cost == best_cost && G > bestG)) { bestG = G; best_cost = cost; best_per_sm = per_sm; }
    }
    lk.unlock();
    if (bestG == 0) return CW_E_BADCONFIG;
    args.group = bestG;
    const size_t smem = smem_for(bestG);
    int64_t blocks = (int64_t)dev->sms * best_per_sm;
    const int64_t groups = (st->n + bestG - 1) / bestG;
    if (blocks > groups) blocks = groups;                         // (work-list launch: the count lives on the device; groups == n)
    if (args.chain && tunables().chain_timeout_ms > 0) {          // once: the spin limit of the chain waits
        static std::once_flag once;
        std::call_once(once, [] {
            const unsigned long long ns = (unsigned long long)tunables().chain_timeout_ms * 1000000ull;
            cudaMemcpyToSymbol(g_chain_timeout_ns, &ns, sizeof(ns));
        });
    }
    if (args.chain && args.chain_pos == 0) {                      // a chain opens: clear its counters and epoch words
        cudaError_t me = cudaMemsetAsync(args.chain, 0, sizeof(uint32_t) * (size_t)(CW_CHAIN_MAX_POS + st->n), stream);
        if (me != cudaSuccess) return (int)me;
    }
    cudaError_t le = launch_pdl(kern, dim3((unsigned)blocks), dim3(kEnvThreads), smem, stream, *cfg, *st, args);
    return (int)(le != cudaSuccess ? le : cudaGetLastError());
}


static int check_state(const CwState* st) {
    if (!st) return CW_E_NULLPTR;
    if (st->n < 0) return CW_E_BADCONFIG;
    if (st->n > 0 && (!st->grid || !st->init_grid || !st->agent || !st->goal || !st->t || !st->episode)) return CW_E_NULLPTR;
    if (st->n_fixed < 0 || (st->n_fixed > 0 && (!st->fixed_grid || !st->fixed_agent))) return CW_E_NULLPTR;
    if (st->goal_grid && !st->goal_agent) return CW_E_NULLPTR;
    return 0;
}

// pre-drawn reset records apply to auto-resetting launches of worlds without a fixed pool, when the caller provides both buffers
static bool reset_records_usable(const CwState* st, int flags) {
    return st->reset_rec && st->reset_list && st->n_fixed == 0 && (flags & CW_F_AUTO_RESET) && st->n <= 0x3FFFFFFF;
}

int step_render_chained_notify(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
                               uint8_t* obs, uint8_t* goal_obs, uint8_t* init_obs, int64_t* stats, int flags, uint32_t* chain,
                               int chain_pos, int obs_ring, uint8_t* status, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (flags & ~CW_F_AUTO_RESET) return CW_E_BADFLAGS;
    if ((flags & CW_F_AUTO_RESET) && (rc = check_reset_config(cfg))) return rc;
    if (chain_pos < 0 || chain_pos >= CW_CHAIN_MAX_POS || obs_ring < 1) return CW_E_BADCONFIG;
    if (st->n == 0) return 0;
    if (!actions || !obs || !chain) return CW_E_NULLPTR;
    if (!status && (!reward || !done)) return CW_E_NULLPTR;
    EnvArgs a = {};
    a.actions = actions; a.reward = reward; a.done = done; a.obs = obs; a.goal_obs = goal_obs; a.init_obs = init_obs;
    a.stats = (unsigned long long*)stats;
    a.mode = M_STEP | M_RENDER | ((flags & CW_F_AUTO_RESET) ? M_AUTO_RESET : 0);
    a.status = status;
    if (tunables().no_chain) return launch_env_kernel(cfg, st, a, (cudaStream_t)stream);   // (debugger / sanitizer sessions, experiments)
    a.chain = chain; a.chain_pos = chain_pos; a.chain_ring = obs_ring;
    // Chains of small batches (about one CTA wave of frames per launch): with 4 CTAs per SM one launch fills the GPU, and the next
    // launch's CTAs only get SM slots -- and only then run their 4 us prologues -- as this one's leave, in bursts.  3 per SM leaves
    // room for the head of the next launch.  Host-driven chains, whose step phases the host is waiting for: 15.5 -> 13.6 us per
    // step at 4096 worlds (K = 128 stream launches).  Graph chains: the ramp of a chain that starts on an idle GPU shrinks (20-step
    // windows at config 2: 14.2-14.9 -> 13.9-14.1 us per step, most on the slower boxes of the pool) for 1 % of the steady state
    // (12.9 -> 13.0 us), profiles/r2_sweep_short_window.txt.
    const double frame_total = (double)st->n * 48.0 * cfg->H * cfg->W;
    if (st->n <= 16384 && (status || frame_total <= 256e6)) a.ctas_cap = 3;
    return launch_env_kernel(cfg, st, a, (cudaStream_t)stream);
}

// ---- pipelined host transport: the two launches of one step (cw_host.cu) -------------------------------------------------------
int step_snap_launch(const CwConfig* cfg, const CwState* st, const uint8_t* actions_dev, const uint8_t* actions_host, uint8_t* status,
                     const PipeSnap* snap, uint32_t* epoch, uint32_t seq, const uint32_t* slot_free, uint32_t slot_want,
                     int64_t* stats, int flags, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (flags & ~CW_F_AUTO_RESET) return CW_E_BADFLAGS;
    if ((flags & CW_F_AUTO_RESET) && (rc = check_reset_config(cfg))) return rc;
    if (st->n == 0) return 0;
    if ((!actions_dev && !actions_host) || !status || !snap || !snap->grid || !snap->meta || !snap->goal || !epoch || !slot_free) return CW_E_NULLPTR;
    if (st->goal_grid) return CW_E_BADCONFIG;
    const int64_t blocks = (st->n + kSnapThreads - 1) / kSnapThreads;
    cudaError_t le;
    if (actions_host && st->n <= kParamActions) {                 // small batch: the actions travel in the kernel parameters
        static thread_local ActionBlock blk;
        memcpy(blk.a, actions_host, (size_t)st->n);
        le = launch_pdl(cw_step_snap_kernel<true>, dim3((unsigned)blocks), dim3(kSnapThreads), 0, (cudaStream_t)stream, *cfg, *st, actions_dev, blk, status,
                        *snap, epoch, seq, slot_free, slot_want, (unsigned long long*)stats, flags);
    } else {
        if (!actions_dev) return CW_E_NULLPTR;
        static const ActionBlock none = {};
        le = launch_pdl(cw_step_snap_kernel<false>, dim3((unsigned)blocks), dim3(kSnapThreads), 0, (cudaStream_t)stream, *cfg, *st, actions_dev, none, status,
                        *snap, epoch, seq, slot_free, slot_want, (unsigned long long*)stats, flags);
    }
    return (int)(le != cudaSuccess ? le : cudaGetLastError());
}

int render_pipe_launch(const CwConfig* cfg, int64_t n, const PipeSnap* snap, uint8_t* obs, uint8_t* goal_obs, uint32_t* chain, int chain_pos,
                       int obs_ring, const uint32_t* epoch, uint32_t seq, uint32_t* slot_free, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    if (n < 0 || chain_pos < 0 || chain_pos >= CW_CHAIN_MAX_POS || obs_ring < 1) return CW_E_BADCONFIG;
    if (n == 0) return 0;
    if (!snap || !snap->grid || !snap->meta || !snap->goal || !obs || !chain || !epoch || !slot_free) return CW_E_NULLPTR;
    CwState st = {};
    st.n = n;
    EnvArgs a = {};
    a.obs = obs; a.goal_obs = goal_obs; a.rgrid = snap->grid; a.mode = M_RENDER;
    a.pmeta = snap->meta; a.pgoal = snap->goal; a.pepoch = epoch; a.pslot = slot_free; a.pseq = seq;
    a.chain = chain; a.chain_pos = chain_pos; a.chain_ring = obs_ring;
    a.ctas_cap = 3;
    return launch_env_kernel(cfg, &st, a, (cudaStream_t)stream);
}

}  // namespace cw

using namespace cw;

extern "C" {

int cw_abi_version(void) { return CW_ABI_VERSION; }

#ifdef CW_TIMING
int cw_debug_set_timing(void* dev_ptr) { return (int)cudaMemcpyToSymbol(cw::g_dbg, &dev_ptr, sizeof(void*)); }
int cw_debug_set_timing_rows_per_position(int rows) { return (int)cudaMemcpyToSymbol(cw::g_dbg_per_pos, &rows, sizeof(int)); }
#endif

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
    int64_t refill = 0;
    CwState s2 = *st;
    if (reset_records_usable(st, flags)) {                        // refill CTAs beside the stepping ones
        DeviceInfo* dev;
        rc = device_info(&dev); if (rc) return rc;
        refill = blocks < (int64_t)dev->sms * 2 ? blocks : (int64_t)dev->sms * 2;   // (stepping + refill CTAs of a 65536-world launch: one wave)
        if (tunables().refill_ctas > 0 && tunables().refill_ctas < refill) refill = tunables().refill_ctas;
    } else {
        s2.reset_rec = nullptr; s2.reset_list = nullptr;
    }
    cudaError_t le = launch_pdl(cw_step_kernel<false>, dim3((unsigned)(blocks + refill)), dim3(128), 0, (cudaStream_t)stream, *cfg, s2, actions,
                                reward, done, (unsigned long long*)stats, (uint8_t*)nullptr, (uint32_t*)nullptr, K, flags, (int)blocks);
    return (int)(le != cudaSuccess ? le : cudaGetLastError());
}

int cw_step_chained(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done, int64_t* stats,
                    int flags, uint32_t* chain, int chain_pos, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (flags & ~CW_F_AUTO_RESET) return CW_E_BADFLAGS;
    if ((flags & CW_F_AUTO_RESET) && (rc = check_reset_config(cfg))) return rc;
    if (chain_pos < 0 || chain_pos >= CW_CHAIN_MAX_POS) return CW_E_BADCONFIG;
    if (st->n == 0) return 0;
    if (!actions || !chain) return CW_E_NULLPTR;
    // Above ~128k worlds a launch is several waves of work and throughput-bound: the ordinary kernel's 128-thread CTAs are the better
    // geometry there (262 144 worlds: 13.8 us per launch against 15.7), and a whole-grid dependent launch is a valid chain member.
    if (tunables().no_chain || st->n > 131072) return cw_rollout(cfg, st, actions, reward, done, stats, 1, flags, stream);
    if (chain_pos == 0) {                                         // a chain opens: clear its counters and marks
        cudaError_t me = cudaMemsetAsync(chain, 0, sizeof(uint32_t) * (size_t)(CW_CHAIN_MAX_POS + (st->n + 31) / 32), (cudaStream_t)stream);
        if (me != cudaSuccess) return (int)me;
    }
    const int64_t blocks = (st->n + kChainStepThreads - 1) / kChainStepThreads;
    cudaError_t le = launch_pdl(cw_step_chained_kernel, dim3((unsigned)blocks), dim3(kChainStepThreads), 0, (cudaStream_t)stream, *cfg, *st, actions,
                                reward, done, (unsigned long long*)stats, chain, (uint32_t)chain_pos, flags);
    return (int)(le != cudaSuccess ? le : cudaGetLastError());
}

int cw_prefill_resets(const CwConfig* cfg, const CwState* st, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_reset_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (st->n == 0) return 0;
    if (!reset_records_usable(st, CW_F_AUTO_RESET)) return CW_E_NULLPTR;
    DeviceInfo* dev;
    rc = device_info(&dev); if (rc) return rc;
    const int64_t want = (st->n + 3) / 4;                         // one warp per world, at most a few waves
    const int64_t blocks = want < (int64_t)dev->sms * 16 ? want : (int64_t)dev->sms * 16;
    cudaError_t le = launch_pdl(cw_step_kernel<false>, dim3((unsigned)blocks), dim3(128), 0, (cudaStream_t)stream, *cfg, *st, (const uint8_t*)nullptr,
                                (int32_t*)nullptr, (uint8_t*)nullptr, (unsigned long long*)nullptr, (uint8_t*)nullptr, (uint32_t*)nullptr, 0,
                                CW_F_AUTO_RESET, 0);
    return (int)(le != cudaSuccess ? le : cudaGetLastError());
}

int cw_step_render_edit(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done, uint8_t* obs,
                        uint8_t* goal_obs, uint8_t* init_obs, int64_t* stats, int flags, uint32_t* scratch, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (flags & ~CW_F_AUTO_RESET) return CW_E_BADFLAGS;
    if ((flags & CW_F_AUTO_RESET) && (rc = check_reset_config(cfg))) return rc;
    if (st->n == 0) return 0;
    if (!actions || !reward || !done || !obs) return CW_E_NULLPTR;
    if ((flags & CW_F_AUTO_RESET) && !scratch) return CW_E_NULLPTR;
    if (st->n > 0xFFFFFFFFll) return CW_E_BADCONFIG;
    const int64_t blocks = (st->n + 127) / 128;
    cudaError_t le = launch_pdl(cw_step_kernel<true>, dim3((unsigned)blocks), dim3(128), 0, (cudaStream_t)stream, *cfg, *st, actions,
                                reward, done, (unsigned long long*)stats, obs, scratch, 1,
                                flags | ((flags & CW_F_AUTO_RESET) ? CW_F_DEFER_RESET : 0), (int)blocks);
    if (le != cudaSuccess) return (int)le;
    if (!(flags & CW_F_AUTO_RESET)) return (int)cudaGetLastError();
    EnvArgs a = {};                                               // reset of the queued worlds + their three frames, one world per CTA
    a.list = scratch; a.obs = obs; a.goal_obs = goal_obs; a.init_obs = init_obs;
    a.mode = M_FORCE_RESET | M_RENDER;
    return launch_env_kernel(cfg, st, a, (cudaStream_t)stream);
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
    if (!actions || !reward || !done) return CW_E_NULLPTR;
    if (!obs && (goal_obs || init_obs)) return CW_E_NULLPTR;
    EnvArgs a = {};
    a.actions = actions; a.reward = reward; a.done = done; a.obs = obs; a.goal_obs = goal_obs; a.init_obs = init_obs;
    a.stats = (unsigned long long*)stats;
    a.mode = M_STEP | (obs ? M_RENDER : 0) | ((flags & CW_F_AUTO_RESET) ? M_AUTO_RESET : 0);
    return launch_env_kernel(cfg, st, a, (cudaStream_t)stream);
}

int cw_step_render_chained(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
                           uint8_t* obs, uint8_t* goal_obs, uint8_t* init_obs, int64_t* stats, int flags, uint32_t* chain,
                           int chain_pos, int obs_ring, void* stream) {
    return cw::step_render_chained_notify(cfg, st, actions, reward, done, obs, goal_obs, init_obs, stats, flags, chain, chain_pos,
                                          obs_ring, nullptr, stream);
}

int cw_step_delta(const CwConfig* cfg, const CwState* st, const uint8_t* actions, void* delta, uint32_t* fresh, int64_t* stats,
                  int flags, int seq, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (flags & ~(CW_F_AUTO_RESET | CW_F_HOST_ACTIONS)) return CW_E_BADFLAGS;
    if ((flags & CW_F_AUTO_RESET) && (rc = check_reset_config(cfg))) return rc;
    if (st->n == 0) return 0;
    if (!actions || !delta || !fresh) return CW_E_NULLPTR;
    if (seq < 0 || seq > 63) return CW_E_BADCONFIG;
    if (st->goal_grid) return CW_E_BADCONFIG;                     // the compact goal state is maintained by cw_reset / cw_step_render only
    const int64_t blocks = (st->n + 127) / 128;
    const int kflags = flags & CW_F_AUTO_RESET;
    cudaError_t le;
    if ((flags & CW_F_HOST_ACTIONS) && st->n <= kParamActions) {   // small batch, actions readable here: ship them in the launch
        static thread_local ActionBlock blk;
        memcpy(blk.a, actions, (size_t)st->n);
        le = launch_pdl(cw_delta_kernel<true>, dim3((unsigned)blocks), dim3(128), 0, (cudaStream_t)stream, *cfg, *st, actions, blk,
                        (uint4*)delta, fresh, (unsigned long long*)stats, (uint32_t)seq, kflags);
    } else {
        static const ActionBlock none = {};
        le = launch_pdl(cw_delta_kernel<false>, dim3((unsigned)blocks), dim3(128), 0, (cudaStream_t)stream, *cfg, *st, actions, none,
                        (uint4*)delta, fresh, (unsigned long long*)stats, (uint32_t)seq, kflags);
    }
    return (int)(le != cudaSuccess ? le : cudaGetLastError());
}

int cw_imagine(const CwConfig* cfg, const CwState* st, uint8_t* goal_obs, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    rc = check_state(st); if (rc) return rc;
    if (st->n == 0) return 0;
    if (!goal_obs && !st->goal_grid) return CW_E_NULLPTR;
    EnvArgs a = {};
    a.goal_obs = goal_obs; a.mode = M_IMAGINE_ONLY;
    return launch_env_kernel(cfg, st, a, (cudaStream_t)stream);
}

int cw_frame_policy(const CwConfig* cfg, const uint8_t* obs, int64_t n, uint8_t* actions, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    if (n < 0) return CW_E_BADCONFIG;
    if (n == 0) return 0;
    if (!obs || !actions) return CW_E_NULLPTR;
    if ((uintptr_t)obs & 15u) return CW_E_BADCONFIG;
    DeviceInfo* dev;
    rc = device_info(&dev); if (rc) return rc;
    // every CTA gets the same number of worlds (4096 worlds on 2368 resident CTAs would give half of them two worlds and the
    // other half one: the launch then lasts two worlds with half the GPU idle in the second)
    const int64_t cap = (int64_t)dev->sms * (2048 / kPolicyThreads);
    const int64_t per_cta = (n + cap - 1) / cap;
    const int64_t blocks = (n + per_cta - 1) / per_cta;
    cudaError_t le = launch_pdl(cw_frame_policy_kernel, dim3((unsigned)blocks), dim3(kPolicyThreads), 0, (cudaStream_t)stream,
                                reinterpret_cast<const uint4*>(obs), n, (uint32_t)(3 * cfg->H * cfg->W), actions);
    return (int)(le != cudaSuccess ? le : cudaGetLastError());
}

int cw_onehot(const CwConfig* cfg, const uint8_t* grid, const uint32_t* agent, uint8_t* onehot, int64_t n, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    if (n < 0) return CW_E_BADCONFIG;
    if (n == 0) return 0;
    if (!grid || !agent || !onehot) return CW_E_NULLPTR;
    DeviceInfo* dev;
    rc = device_info(&dev); if (rc) return rc;
    const int64_t n_cells = n * cfg->H * cfg->W;
    if ((uintptr_t)onehot & 3u) {                                 // foreign, unaligned buffer: byte-wise path
        int64_t blocks = (n_cells + 255) / 256;
        if (blocks > (int64_t)dev->sms * 16) blocks = (int64_t)dev->sms * 16;
        cw_onehot_unaligned_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*cfg, grid, agent, onehot, n_cells);
        return (int)cudaGetLastError();
    }
    const int64_t items = (n_cells + kOneHotCells - 1) / kOneHotCells;
    int64_t blocks = items < (int64_t)dev->sms * 5 ? items : (int64_t)dev->sms * 5;   // 2 x 21.5 KB stages: 5 CTAs per SM
    const uint32_t hw_magic = (uint32_t)(0x100000000ull / (uint64_t)(cfg->H * cfg->W)) + 1u;
    cw_onehot_kernel<<<(unsigned)blocks, kExpThreads, 2 * kOneHotStage, (cudaStream_t)stream>>>(*cfg, grid, agent, onehot, n_cells, n, hw_magic);
    return (int)cudaGetLastError();
}

int cw_render_alt(const CwConfig* cfg, const uint8_t* grid, const uint32_t* agent, int16_t* obs, int64_t n, void* stream) {
    int rc = check_config(cfg); if (rc) return rc;
    if (n < 0) return CW_E_BADCONFIG;
    if (n == 0) return 0;
    if (!grid || !agent || !obs) return CW_E_NULLPTR;
    if ((uintptr_t)obs & 1u) return CW_E_BADCONFIG;
    DeviceInfo* dev;
    rc = device_info(&dev); if (rc) return rc;
    const uint32_t row_bytes = 3u * 3u * (uint32_t)cfg->W * 6u;   // one row of cells = 3 pixel rows
    const uint32_t strip_bytes = row_bytes;                       // the status strip is 3 pixel rows too
    const uint32_t world_bytes = row_bytes * (uint32_t)cfg->H + strip_bytes;
    AltPlan plan;
    if (world_bytes <= kAltBudget) {
        plan.worlds_per_item = (int)(kAltBudget / world_bytes); if (plan.worlds_per_item > 16) plan.worlds_per_item = 16;
        plan.bands = 1; plan.rows_per_band = cfg->H;
        plan.items = (n + plan.worlds_per_item - 1) / plan.worlds_per_item;
    } else {
        plan.worlds_per_item = 1;
        plan.rows_per_band = (int)((kAltBudget - strip_bytes) / row_bytes); if (plan.rows_per_band < 1) plan.rows_per_band = 1;
        plan.bands = (cfg->H + plan.rows_per_band - 1) / plan.rows_per_band;
        plan.items = n * plan.bands;
    }
    plan.w_magic = (uint32_t)(0x100000000ull / (uint64_t)cfg->W) + 1u;
    const uint32_t band_cells = (uint32_t)(plan.bands == 1 ? cfg->H : plan.rows_per_band) * (uint32_t)cfg->W;
    plan.cells_magic = (uint32_t)(0x100000000ull / (uint64_t)band_cells) + 1u;
    {
        std::lock_guard<std::mutex> lk(g_dev_mu);
        if (!dev->alt_attr_set) {                                 // 2 stages x 32 KB: above the 48 KB default
            cudaError_t e = cudaFuncSetAttribute(cw_render_alt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * kAltStage));
            if (e != cudaSuccess) return (int)e;
            dev->alt_attr_set = true;
        }
    }
    int64_t blocks = plan.items < (int64_t)dev->sms * 3 ? plan.items : (int64_t)dev->sms * 3;
    cw_render_alt_kernel<<<(unsigned)blocks, kExpThreads, 2 * kAltStage, (cudaStream_t)stream>>>(*cfg, grid, agent, obs, n, plan);
    return (int)cudaGetLastError();
}

}  // extern "C"
