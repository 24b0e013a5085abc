// cw_device.cuh -- device-side building blocks of the batched CraftingWorld hot path (sm_100a).
//
// Semantics follow the reference env, gym_craftingworld/envs/craftingworld_ray.py ("ray.py") and
// envs/coordinates.py; the citations beside each block are the lines it re-implements for N worlds.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cw_b200.h"

namespace cw {

enum : int { EMPTY = 0, STICKS, AXE, HAMMER, ROCK, TREE, BREAD, HOUSE, WHEAT };  // ray.py:21 (+1)
enum : int {                                                                      // ray.py:40-41
    T_MAKE_BREAD = 0, T_EAT_BREAD, T_BUILD_HOUSE, T_CHOP_TREE, T_CHOP_ROCK, T_GO_TO_HOUSE, T_MOVE_AXE,
    T_MOVE_HAMMER, T_MOVE_STICKS
};

// COLORS_N (ray.py:28-30) packed R | G<<8 | B<<16
#define CW_RGB(r, g, b) ((uint32_t)(r) | ((uint32_t)(g) << 8) | ((uint32_t)(b) << 16))
__device__ __constant__ uint32_t kColorLUT[9] = {
    CW_RGB(0, 0, 0),       CW_RGB(110, 69, 39),  CW_RGB(255, 105, 180), CW_RGB(100, 100, 200), CW_RGB(100, 100, 100),
    CW_RGB(0, 128, 0),     CW_RGB(205, 133, 63), CW_RGB(197, 91, 97),   CW_RGB(240, 230, 140)};

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ------------------------------------------------------------------------------------------------------
// Philox4x32-10, warp-replicated stream.  Stream layout (the spec is oracle/compact.py PhiloxStream):
// key = (seed_lo, seed_hi), counter = (env_lo, env_hi, episode, block), words of a block consumed in order.
// Lane l holds block (base + l), so one refill yields 128 words; every lane tracks the same position and
// sees the same draws, so all control flow on the draws is warp-uniform.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t k0,
                                              uint32_t k1) {
#pragma unroll
    for (int i = 0; i < 10; i++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

struct WarpPhilox {
    uint32_t k0, k1, e0, e1, ep, base;
    uint32_t w0, w1, w2, w3;
    int pos;  // 0..128

    __device__ __forceinline__ void init(uint64_t seed, uint64_t env_id, uint32_t episode) {
        k0 = (uint32_t)seed; k1 = (uint32_t)(seed >> 32);
        e0 = (uint32_t)env_id; e1 = (uint32_t)(env_id >> 32); ep = episode;
        base = 0; pos = 128;
    }
    __device__ __forceinline__ void refill() {
        w0 = e0; w1 = e1; w2 = ep; w3 = base + (uint32_t)lane_id();
        philox4x32_10(w0, w1, w2, w3, k0, k1);
        base += 32; pos = 0;
    }
    __device__ __forceinline__ uint32_t next32() {
        if (pos == 128) refill();
        const int j = pos & 3;
        const uint32_t mine = j == 0 ? w0 : (j == 1 ? w1 : (j == 2 ? w2 : w3));
        const uint32_t v = __shfl_sync(0xffffffffu, mine, pos >> 2);
        pos++;
        return v;
    }
    // unbiased integer in [0,n): Lemire multiply-shift with rejection
    __device__ __forceinline__ uint32_t uniform(uint32_t n) {
        uint64_t m = (uint64_t)next32() * n;
        uint32_t lo = (uint32_t)m;
        if (lo < n) {
            const uint32_t thresh = (0u - n) % n;
            while (lo < thresh) { m = (uint64_t)next32() * n; lo = (uint32_t)m; }
        }
        return (uint32_t)(m >> 32);
    }
};

// ------------------------------------------------------------------------------------------------------
// step(action): ray.py:301-378.  `g` is this world's grid (global or shared), `ig` its initial grid (global).
// All cell reads are issued up front (no dependent load chain); at most one cell is written.
// Returns the reward; done/changed by reference.  `wcell`/`wval` report the single grid write (wcell < 0: none)
// so a caller that steps on a shared-memory copy can mirror it to global memory.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t setbit(uint32_t m, int bit, bool on) {
    return on ? (m | (1u << bit)) : (m & ~(1u << bit));
}

// kEagerInit: read the two INIT_OBS cells unconditionally, together with the grid cells (one dependent load level less on
// a latency-critical path), instead of only when something is held.
// `ocode` / `ncode`: the object codes, AFTER the step, of the cell the agent stood on before the step and of the cell it stands
// on now (the same cell unless it moved) -- everything a frame consumer needs to repaint the <= 2 cells a step touches
// (render_edit, ray.py:522-557) without a copy of the grid.  Dead code for callers that ignore them.
template <bool kEagerInit = false>
__device__ __forceinline__ int step_core_ex(const CwConfig& cfg, uint8_t* __restrict__ g, const uint8_t* __restrict__ ig,
                                            uint32_t& agent, uint32_t& goal, int& t, int a, bool& done, int& wcell,
                                            int& wval, int& ocode, int& ncode) {
    const int W = cfg.W, H = cfg.H, M = cfg.max_steps;
    int r = agent & 0xFF, c = (agent >> 8) & 0xFF, h = (agent >> 16) & 0xFF;
    uint32_t ach = goal & 0xFFFFu;
    const uint32_t des = goal >> 16;
    t += 1;                                                                      // ray.py:309
    wcell = -1; wval = 0;
    bool changed;
    const int cell = r * W + c;
    if (a >= 4) {
        const int here = g[cell];
        if (a == 4) {                                                            // pickup, ray.py:314-327
            changed = (here >= STICKS) & (here <= HAMMER) & (h == 0);            // ray.py:317-322
            if (changed) { h = here; wcell = cell; wval = EMPTY; }               // ray.py:326-327
        } else if (a == 5) {                                                     // drop, ray.py:329-341
            changed = (h != 0) & (here == EMPTY);                                // ray.py:332-335
            if (changed) { wcell = cell; wval = h; h = 0; }                      // ray.py:339-341
        } else {
            changed = false;  // out-of-range action: defined no-op (reference: IndexError, ray.py:308)
        }
        ocode = ncode = (wcell >= 0) ? wval : here;
    } else {                                                                     // move, ray.py:343-346, 380-440
        const int dr = (a == 2) - (a == 0), dc = (a == 1) - (a == 3);            // ray.py:130-131
        const int nr = min(max(r + dr, 0), H - 1), nc = min(max(c + dc, 0), W - 1);   // coordinates.py:22-25
        const int ncell = nr * W + nc;
        int here = g[cell];
        ocode = here;                                                            // a move never changes the cell it leaves
        const int T = g[ncell];
        int ih = 0, ihn = 0;
        if (kEagerInit || h != 0) { ih = __ldcg(ig + cell); ihn = __ldcg(ig + ncell); }   // L2: a co-resident chained launch may have re-seeded it
        int old = EMPTY;  // EMPTY == "None": matches no predicate below (ray.py:655 uses 100)
        bool moved = ncell != cell;                                              // ray.py:395-396
        const bool blocked = ((T == ROCK) & (h != HAMMER)) | ((T == TREE) & (h != AXE));   // ray.py:401-405
        moved = moved & !blocked;
        if (moved) {
            r = nr; c = nc; old = T; ih = ihn;                                   // ray.py:407-411
            int nv = T;
            if (T == ROCK || T == BREAD) nv = EMPTY;                             // ray.py:423-425
            else if (T == TREE) nv = STICKS;                                     // ray.py:426-428
            else if (T == STICKS && h == HAMMER) nv = HOUSE;                     // ray.py:429-432
            else if (T == WHEAT && h == AXE) nv = BREAD;                         // ray.py:433-438
            if (nv != T) { wcell = ncell; wval = nv; }
            here = nv;
        }
        changed = moved;
        ncode = here;
        // eval_task_edit: for EVERY move action, successful or not (ray.py:345-346, 646-703)
        if (old == BREAD) ach |= 1u << T_EAT_BREAD;                              // ray.py:657-659
        else if (old == ROCK) ach |= 1u << T_CHOP_ROCK;                          // ray.py:660-662
        else if (old == TREE) ach |= 1u << T_CHOP_TREE;                          // ray.py:663-665
        ach = setbit(ach, T_GO_TO_HOUSE, here == HOUSE);                         // ray.py:668 (level triggered)
        if (h == STICKS) {                                                       // ray.py:672-684
            const bool home = (ih == STICKS) | ((ih == TREE) & ((ach >> T_CHOP_TREE) & 1u));
            ach = setbit(ach, T_MOVE_STICKS, !home);
        } else if (h == AXE) {                                                   // ray.py:685-693
            if (old == WHEAT) ach |= 1u << T_MAKE_BREAD;
            ach = setbit(ach, T_MOVE_AXE, ih != AXE);
        } else if (h == HAMMER) {                                                // ray.py:694-702
            if (old == STICKS) ach |= 1u << T_BUILD_HOUSE;
            ach = setbit(ach, T_MOVE_HAMMER, ih != HAMMER);
        }
    }
    if (wcell >= 0) g[wcell] = (uint8_t)wval;
    int reward = -1;                                                             // ray.py:362-363
    if (changed) {                                                               // ray.py:348, 361
        const bool success = cfg.subset_reward ? ((des & ~ach) == 0)             // ray.py:763-767
                                               : (ach == des);                   // ray.py:747-761
        if (success) reward = M;
    }
    done = (t >= M) | (reward == M);                                             // ray.py:367
    agent = (uint32_t)r | ((uint32_t)c << 8) | ((uint32_t)h << 16);
    goal = ach | (des << 16);
    return reward;
}

template <bool kEagerInit = false>
__device__ __forceinline__ int step_core(const CwConfig& cfg, uint8_t* __restrict__ g, const uint8_t* __restrict__ ig,
                                         uint32_t& agent, uint32_t& goal, int& t, int a, bool& done, int& wcell,
                                         int& wval) {
    int ocode, ncode;
    return step_core_ex<kEagerInit>(cfg, g, ig, agent, goal, t, a, done, wcell, wval, ocode, ncode);
}

// episode statistics of a finished episode (host reduces them across ranks with one small all-reduce)
__device__ __forceinline__ void stats_add(const CwConfig& cfg, unsigned long long* stats, uint32_t goal, int t, int reward) {
    const bool success = reward == cfg.max_steps;
    const long long ret = success ? (long long)cfg.max_steps - (t - 1) : -(long long)t;
    atomicAdd(stats + 0, 1ull);
    if (success) atomicAdd(stats + 1, 1ull);
    atomicAdd(stats + 2, (unsigned long long)ret);
    atomicAdd(stats + 3, (unsigned long long)t);
#pragma unroll
    for (int i = 0; i < 9; i++) {
        if ((goal >> i) & 1u) atomicAdd(stats + 4 + i, 1ull);
        if ((goal >> (16 + i)) & 1u) atomicAdd(stats + 13 + i, 1ull);
    }
}

struct Sparse8 {   // a world with (at most) one object per entry: cell index + object code, static indexing only
    uint32_t cell[8];
    uint32_t code[8];
};

// desired_goal_vector: n_tasks = U{1..number_of_tasks} if stacking else 1; partial Fisher-Yates over selected_tasks, kept as
// 4-bit entries of one 64-bit word (ray.py:169-174).  Serial walk of the stream (the spec, oracle/compact.py).
__device__ __forceinline__ uint32_t reset_tasks_serial(const CwConfig& cfg, WarpPhilox& rng) {
    uint64_t perm = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) perm |= (uint64_t)(cfg.selected[i] & 15) << (4 * i);
    uint32_t des = 0;
    const int ntask = cfg.stacking ? (int)rng.uniform((uint32_t)cfg.number_of_tasks) + 1 : 1;
    for (int i = 0; i < ntask; i++) {
        const int j = i + (int)rng.uniform((uint32_t)(cfg.n_selected - i));
        const uint64_t vi = (perm >> (4 * i)) & 15, vj = (perm >> (4 * j)) & 15;
        perm = (perm & ~((uint64_t)15 << (4 * i)) & ~((uint64_t)15 << (4 * j))) | (vj << (4 * i)) | (vi << (4 * j));
        des |= 1u << (uint32_t)vj;
    }
    return des;
}

// The RANDOM part of reset() for a world without a fixed pool: task sampling (ray.py:169-174) + sample_state's 9 distinct
// cells (605-613; cells[0..7] the objects sticks..wheat, cells[8] the agent), by one full warp, from the stream
// (seed, global id of world n, episode ep).  Writes nothing: the result depends only on (seed, id, ep), so it can be drawn
// long before the reset happens (cw_step_kernel's pre-drawn reset records).  `rng` is left positioned after the placement
// draws so imagine_* can continue the same stream.
__device__ __forceinline__ void reset_sample(const CwConfig& cfg, const CwState& st, int64_t n, WarpPhilox& rng, uint32_t ep,
                                             uint32_t& des_out, uint32_t (&cells)[9]) {
    const int lane = lane_id();
    rng.init(st.seed, st.env_id_base + (uint64_t)n, ep);
    rng.refill();                                                // 128 draws buffered: draw j = lane j>>2, slot j&3
    const uint32_t ncell = (uint32_t)(cfg.H * cfg.W);
    uint64_t perm = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) perm |= (uint64_t)(cfg.selected[i] & 15) << (4 * i);
    uint32_t des = 0;
    bool sampled = false;
    // ---- lane-parallel sampling with exact redraw semantics --------------------------------------------------------
    // The serial algorithm (the fallback below; spec: oracle/compact.py) walks the stream: draw i+1 follows the last draw
    // used by draw i, and a draw is repeated when it lands in Lemire's rejection zone or (placement) on an occupied cell.
    // Here lane i evaluates position i from stream offset (base + i + shift_i); the FIRST violating position k is exactly
    // where the serial loop would redraw, so all positions >= k move one draw forward and the round repeats.  Rounds are
    // rare (8 % of resets at 21x21 need one), every round is a handful of shuffles instead of a dependent loop.
    {
        auto fetch = [&](int d) {                                // draw d (per lane) of the buffered 128-draw block
            const int src = d >> 2, sl = d & 3;
            const uint32_t a0 = __shfl_sync(0xffffffffu, rng.w0, src), a1 = __shfl_sync(0xffffffffu, rng.w1, src);
            const uint32_t a2 = __shfl_sync(0xffffffffu, rng.w2, src), a3 = __shfl_sync(0xffffffffu, rng.w3, src);
            return sl == 0 ? a0 : (sl == 1 ? a1 : (sl == 2 ? a2 : a3));
        };
        bool ok = true;
        int consumed = 0, ntask = 1;
        if (cfg.stacking) {                                      // ray.py:169 (warp-uniform)
            const uint32_t nn = (uint32_t)cfg.number_of_tasks, thresh = (0u - nn) % nn;
            for (;;) {
                const uint64_t m = (uint64_t)fetch(consumed++) * nn;
                if ((uint32_t)m >= thresh) { ntask = (int)(m >> 32) + 1; break; }
                if (consumed > 32) { ok = false; break; }
            }
        }
        const int li = lane < 8 ? lane : 8;
        int j_mine = 0;
        if (ok) {                                                // Fisher-Yates: j_i = i + uniform(n_selected - i)
            const uint32_t nsel_i = lane < ntask ? (uint32_t)(cfg.n_selected - li) : 1u;
            const uint32_t thresh_i = (0u - nsel_i) % nsel_i;
            int shift = 0;
            for (;;) {
                const uint64_t mf = (uint64_t)fetch(consumed + li + shift) * nsel_i;
                j_mine = li + (int)(mf >> 32);
                const uint32_t bad = __ballot_sync(0xffffffffu, lane < ntask && (uint32_t)mf < thresh_i);
                if (!bad) break;
                if (lane >= __ffs(bad) - 1) shift++;
                if (consumed + 9 + __shfl_sync(0xffffffffu, shift, 8) >= 100) { ok = false; break; }   // stay inside the block
            }
            consumed += ntask + __shfl_sync(0xffffffffu, shift, ntask - 1);
        }
        uint32_t cell_mine = 0;
        if (ok) {                                                // placement: 9 distinct uniform cells, ray.py:605-613
            const uint32_t thresh_c = (0u - ncell) % ncell;
            int shift = 0;
            for (;;) {
                const uint64_t mc = (uint64_t)fetch(consumed + li + shift) * ncell;
                cell_mine = (uint32_t)(mc >> 32);
                const uint32_t twins = __match_any_sync(0xffffffffu, lane < 9 ? cell_mine : 0x80000000u + (uint32_t)lane);
                const bool viol = lane < 9 && ((uint32_t)mc < thresh_c || (twins & ((1u << lane) - 1u)) != 0);
                const uint32_t bad = __ballot_sync(0xffffffffu, viol);
                if (!bad) break;
                if (lane >= __ffs(bad) - 1) shift++;
                if (consumed + 9 + __shfl_sync(0xffffffffu, shift, 8) >= 127) { ok = false; break; }
            }
            consumed += 9 + __shfl_sync(0xffffffffu, shift, 8);
        }
        if (ok) {
            for (int i = 0; i < ntask; i++) {
                const int j = __shfl_sync(0xffffffffu, j_mine, i);
                const uint64_t vi = (perm >> (4 * i)) & 15, vj = (perm >> (4 * j)) & 15;
                perm = (perm & ~((uint64_t)15 << (4 * i)) & ~((uint64_t)15 << (4 * j))) | (vj << (4 * i)) | (vi << (4 * j));
                des |= 1u << (uint32_t)vj;
            }
#pragma unroll
            for (int k = 0; k < 9; k++) cells[k] = __shfl_sync(0xffffffffu, cell_mine, k);
            rng.pos = consumed;
            sampled = true;
        } else {
            rng.pos = 0;                                         // ultra-rare: rewind and walk the stream serially
        }
    }
    if (!sampled) {
        des = reset_tasks_serial(cfg, rng);
#pragma unroll
        for (int k = 0; k < 9; k++) {                            // sample_state: 9 distinct uniform cells (ray.py:605-613)
            uint32_t cell;
            bool dup;
            do {
                cell = rng.uniform(ncell);
                dup = false;
#pragma unroll
                for (int q = 0; q < k; q++) dup |= cells[q] == cell;
            } while (dup);
            cells[k] = cell;
        }
    }
    des_out = des;
}

// The DETERMINISTIC part of reset() (ray.py:176-203) for a drawn placement: grid + init_grid rows as 16-byte chunks composed in
// registers (one chunk per lane per pass; `sg`, if non-null, receives a shared-memory copy), episode counter, agent / goal words.
__device__ __forceinline__ void reset_apply(const CwConfig& cfg, const CwState& st, int64_t n, uint8_t* sg, uint32_t des,
                                            const uint32_t (&cells)[9], uint32_t ep, uint32_t& agent_out, uint32_t& goal_out,
                                            Sparse8* objs = nullptr) {
    const int lane = lane_id();
    const int nchunk = cfg.cell_stride >> 4;
    uint4* gg = reinterpret_cast<uint4*>(st.grid + n * cfg.cell_stride);
    uint4* gi = reinterpret_cast<uint4*>(st.init_grid + n * cfg.cell_stride);
    for (int ch = lane; ch < nchunk; ch += 32) {
        uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if ((int)(cells[k] >> 4) == ch) {
                const uint32_t v = (uint32_t)(k + 1) << (8 * (cells[k] & 3));
                const int wi = (cells[k] >> 2) & 3;
                w0 |= wi == 0 ? v : 0; w1 |= wi == 1 ? v : 0; w2 |= wi == 2 ? v : 0; w3 |= wi == 3 ? v : 0;
            }
        }
        const uint4 v4 = make_uint4(w0, w1, w2, w3);
        gg[ch] = v4;
        gi[ch] = v4;                                                         // INIT_OBS_VECTOR, ray.py:183
        if (sg) reinterpret_cast<uint4*>(sg)[ch] = v4;
    }
    const uint32_t ar = cells[8] / (uint32_t)cfg.W;
    if (objs) {
#pragma unroll
        for (int k = 0; k < 8; k++) { objs->cell[k] = cells[k]; objs->code[k] = (uint32_t)(k + 1); }
    }
    if (lane == 0) st.episode[n] = ep + 1;
    agent_out = ar | ((cells[8] - ar * (uint32_t)cfg.W) << 8);                   // holding nothing
    goal_out = des << 16;                                                        // achieved = 0, ray.py:176
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------
// reset(): ray.py:156-218, executed by one full warp for world `n`.
// Draw order (spec: oracle/compact.py reset_env): task count, task subset (169-174), placement (605-613).
// Writes grid + init_grid (global; and `sg` if non-null, a shared copy) and episode[n]; returns agent/goal.
// `rng` is left positioned after the placement draws so imagine_warp can continue the same stream.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void reset_warp(const CwConfig& cfg, const CwState& st, int64_t n, uint8_t* sg, WarpPhilox& rng,
                                           uint32_t& agent_out, uint32_t& goal_out, uint32_t ep, Sparse8* objs = nullptr,
                                           uint32_t* s_scratch = nullptr) {
    // `ep` = st.episode[n], loaded by the caller together with the world's other scalars (no dependent load here)
    if (st.n_fixed == 0) {
        uint32_t des, cells[9];
        reset_sample(cfg, st, n, rng, ep, des, cells);
        reset_apply(cfg, st, n, sg, des, cells, ep, agent_out, goal_out, objs);
        return;
    }
    // generate_fixed_initial_state: uniform pick from the pre-sampled pool (ray.py:636-644)
    const int lane = lane_id();
    rng.init(st.seed, st.env_id_base + (uint64_t)n, ep);
    rng.refill();
    const uint32_t des = reset_tasks_serial(cfg, rng);
    const int nchunk = cfg.cell_stride >> 4;
    uint4* gg = reinterpret_cast<uint4*>(st.grid + n * cfg.cell_stride);
    uint4* gi = reinterpret_cast<uint4*>(st.init_grid + n * cfg.cell_stride);
    const uint32_t idx = rng.uniform((uint32_t)st.n_fixed);
    const uint4* src = reinterpret_cast<const uint4*>(st.fixed_grid + (size_t)idx * cfg.cell_stride);
    for (int ch = lane; ch < nchunk; ch += 32) {
        const uint4 v4 = src[ch];
        gg[ch] = v4;
        gi[ch] = v4;
        if (sg) reinterpret_cast<uint4*>(sg)[ch] = v4;
    }
    if (objs) {   // object list of the pooled world (one of each code): found by one pass over the tile
        __syncwarp();
        for (int ch = lane; ch < nchunk; ch += 32) {
            const uint4 v4 = src[ch];
            const uint32_t w[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int b = 0; b < 16; b++) {
                const uint32_t c = (w[b >> 2] >> (8 * (b & 3))) & 0xFFu;
                if (c >= 1 && c <= 8) s_scratch[c - 1] = (uint32_t)(16 * ch + b);
            }
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 8; k++) { objs->cell[k] = s_scratch[k]; objs->code[k] = (uint32_t)(k + 1); }
        __syncwarp();
    }
    if (lane == 0) st.episode[n] = ep + 1;
    agent_out = st.fixed_agent[idx] & 0xFFFFu;                                   // holding nothing
    goal_out = des << 16;                                                        // achieved = 0, ray.py:176
    __syncwarp();
}

// warp-cooperative scans over a shared-memory grid tile, in np.where (row-major) order, 32 cells per pass.
// (A 16-cells-per-lane byte-SIMD variant was measured slower on this latency-bound path)
// position of the j-th (0-based) set bit of m; j < popc(m).  (__fns is software-emulated and slow.)
__device__ __forceinline__ int nth_set_bit(uint32_t m, int j) {
    for (int i = 0; i < j; i++) m &= m - 1;
    return __ffs(m) - 1;
}
__device__ __forceinline__ int warp_count(const uint8_t* g, int n, int code, int skip) {
    int cnt = 0;
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        const bool p = (i < n) && (g[i] == code) && (i != skip);
        cnt += __popc(__ballot_sync(0xffffffffu, p));
    }
    return cnt;
}
__device__ __forceinline__ int warp_nth(const uint8_t* g, int n, int code, int k, int skip) {
    for (int base = 0; base < n; base += 32) {
        const int i = base + lane_id();
        const bool p = (i < n) && (g[i] == code) && (i != skip);
        const uint32_t m = __ballot_sync(0xffffffffu, p);
        const int c = __popc(m);
        if (k < c) return base + nth_set_bit(m, k);
        k -= c;
    }
    return -1;
}
__device__ __forceinline__ void warp_set(uint8_t* g, int cell, int code) {
    __syncwarp();
    if (lane_id() == 0) g[cell] = (uint8_t)code;
    __syncwarp();
}

// imagine_obs: ray.py:220-299, by one full warp on a shared-memory scratch copy `g` of the initial grid.
// Skills are applied in the reference's fixed order; a skill with no candidate object is skipped (the reference
// would raise on randint(0); unreachable from sample_state worlds).
__device__ __forceinline__ void imagine_warp(const CwConfig& cfg, uint8_t* g, uint32_t& agent, uint32_t des, WarpPhilox& rng) {
    const int n = cfg.H * cfg.W, W = cfg.W;
    int r = agent & 0xFF, c = (agent >> 8) & 0xFF;
    const int acell = r * W + c;
    int cnt, k, src, dst, fr;
    if ((des >> T_MAKE_BREAD) & 1u) {                                            // ray.py:226-231
        src = warp_nth(g, n, WHEAT, 0, -1);
        if (src >= 0) warp_set(g, src, BREAD);
    }
    if ((des >> T_EAT_BREAD) & 1u) {                                             // ray.py:232-237
        cnt = warp_count(g, n, BREAD, -1);
        if (cnt) { k = (int)rng.uniform((uint32_t)cnt); warp_set(g, warp_nth(g, n, BREAD, k, -1), EMPTY); }
    }
    if ((des >> T_CHOP_TREE) & 1u) {                                             // ray.py:238-243
        src = warp_nth(g, n, TREE, 0, -1);
        if (src >= 0) warp_set(g, src, STICKS);
    }
    if ((des >> T_MOVE_STICKS) & 1u) {                                           // ray.py:244-257
        cnt = warp_count(g, n, STICKS, -1);
        if (cnt) {
            k = (int)rng.uniform((uint32_t)cnt);
            fr = warp_count(g, n, EMPTY, acell);                                 // [:9] -> agent cell is occupied
            if (fr) {
                const int spot = (int)rng.uniform((uint32_t)fr);
                src = warp_nth(g, n, STICKS, k, -1);
                dst = warp_nth(g, n, EMPTY, spot, acell);
                warp_set(g, src, EMPTY);
                warp_set(g, dst, STICKS);
            }
        }
    }
    if ((des >> T_BUILD_HOUSE) & 1u) {                                           // ray.py:258-264
        cnt = warp_count(g, n, STICKS, -1);
        if (cnt) { k = (int)rng.uniform((uint32_t)cnt); warp_set(g, warp_nth(g, n, STICKS, k, -1), HOUSE); }
    }
    if ((des >> T_CHOP_ROCK) & 1u) {                                             // ray.py:265-268
        src = warp_nth(g, n, ROCK, 0, -1);
        if (src >= 0) warp_set(g, src, EMPTY);
    }
    if ((des >> T_GO_TO_HOUSE) & 1u) {                                           // ray.py:269-276
        cnt = warp_count(g, n, HOUSE, -1);
        if (cnt) {
            k = (int)rng.uniform((uint32_t)cnt);
            dst = warp_nth(g, n, HOUSE, k, -1);
            r = dst / W; c = dst - r * W;
        }
    }
    if ((des >> T_MOVE_AXE) & 1u) {                                              // ray.py:277-286
        src = warp_nth(g, n, AXE, 0, -1);
        if (src >= 0) {
            fr = warp_count(g, n, EMPTY, -1);                                    // [:8] -> agent cell allowed
            if (fr) {
                const int spot = (int)rng.uniform((uint32_t)fr);
                dst = warp_nth(g, n, EMPTY, spot, -1);
                warp_set(g, src, EMPTY);
                warp_set(g, dst, AXE);
            }
        }
    }
    if ((des >> T_MOVE_HAMMER) & 1u) {                                           // ray.py:287-297
        src = warp_nth(g, n, HAMMER, 0, -1);
        if (src >= 0) {
            fr = warp_count(g, n, EMPTY, -1);
            if (fr) {
                const int spot = (int)rng.uniform((uint32_t)fr);
                dst = warp_nth(g, n, EMPTY, spot, -1);
                warp_set(g, src, EMPTY);
                warp_set(g, dst, HAMMER);
            }
        }
    }
    agent = (agent & 0xFFFF0000u) | (uint32_t)r | ((uint32_t)c << 8);
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------
// imagine_obs (ray.py:220-299) in closed form for a world FRESH from reset(): sample_state / the fixed pool put
// exactly one of each object on the grid and the agent on an empty cell, so the world is an 8-entry list
// (index k = object k+1 initially) and every "np.where" candidate set of the reference has at most two
// members with known indices.  Same draws, same order, same results as imagine_warp's grid scans (the oracle
// checks both); ~10x fewer dependent instructions, which matters because it runs on one warp per reset.
// Row-major (np.where) order of two objects = order of their cell indices.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t nth_free_cell(const Sparse8& o, uint32_t extra, uint32_t spot) {
    // spot-th (0-based) cell, in row-major order, that holds no object and is not `extra`:
    // least fixed point of c = spot + #{occupied <= c}
    uint32_t c = spot;
    for (;;) {
        uint32_t cnt = extra <= c ? 1u : 0u;
#pragma unroll
        for (int k = 0; k < 8; k++) cnt += (o.code[k] != 0 && o.cell[k] <= c) ? 1u : 0u;
        const uint32_t nx = spot + cnt;
        if (nx == c) return c;
        c = nx;
    }
}
__device__ __forceinline__ uint32_t count_objects(const Sparse8& o) {
    uint32_t n = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) n += o.code[k] != 0 ? 1u : 0u;
    return n;
}

__device__ __forceinline__ void imagine_fresh(const CwConfig& cfg, Sparse8& o, uint32_t& agent, uint32_t des, WarpPhilox& rng) {
    enum { S = 0, A = 1, HM = 2, R = 3, T = 4, B = 5, HO = 6, WH = 7 };   // list index of each initial object
    const uint32_t HW = (uint32_t)(cfg.H * cfg.W), W = (uint32_t)cfg.W;
    const uint32_t acell = (agent & 0xFF) * W + ((agent >> 8) & 0xFF);
    uint32_t anew = acell;
    if ((des >> T_MAKE_BREAD) & 1u) o.code[WH] = BREAD;                             // ray.py:226-231
    if ((des >> T_EAT_BREAD) & 1u) {                                                // ray.py:232-237
        const bool two = o.code[WH] == BREAD;                                       // breads: B, and WH once baked
        const uint32_t k = rng.uniform(two ? 2u : 1u);
        const bool b_first = !two || o.cell[B] < o.cell[WH];
        if ((k == 0) == b_first) o.code[B] = EMPTY; else o.code[WH] = EMPTY;
    }
    if ((des >> T_CHOP_TREE) & 1u) o.code[T] = STICKS;                              // ray.py:238-243
    if ((des >> T_MOVE_STICKS) & 1u) {                                              // ray.py:244-257
        const bool two = o.code[T] == STICKS;                                       // sticks: S, and T once chopped
        const uint32_t k = rng.uniform(two ? 2u : 1u);
        const bool s_first = !two || o.cell[S] < o.cell[T];
        const bool move_s = (k == 0) == s_first;
        const uint32_t fr = HW - count_objects(o) - 1u;                             // [:9]: the agent cell is occupied
        if ((int)fr > 0) {
            const uint32_t dst = nth_free_cell(o, acell, rng.uniform(fr));
            if (move_s) o.cell[S] = dst; else o.cell[T] = dst;
        }
    }
    if ((des >> T_BUILD_HOUSE) & 1u) {                                              // ray.py:258-264
        const bool two = o.code[T] == STICKS;
        const uint32_t k = rng.uniform(two ? 2u : 1u);
        const bool s_first = !two || o.cell[S] < o.cell[T];
        if ((k == 0) == s_first) o.code[S] = HOUSE; else o.code[T] = HOUSE;
    }
    if ((des >> T_CHOP_ROCK) & 1u) o.code[R] = EMPTY;                               // ray.py:265-268
    if ((des >> T_GO_TO_HOUSE) & 1u) {                                              // ray.py:269-276
        const bool s_house = o.code[S] == HOUSE, t_house = o.code[T] == HOUSE;      // at most one of them was built
        const bool two = s_house || t_house;
        const uint32_t other = s_house ? o.cell[S] : o.cell[T];
        const uint32_t k = rng.uniform(two ? 2u : 1u);
        const bool ho_first = !two || o.cell[HO] < other;
        anew = ((k == 0) == ho_first) ? o.cell[HO] : other;
    }
    if ((des >> T_MOVE_AXE) & 1u) {                                                 // ray.py:277-286
        const uint32_t fr = HW - count_objects(o);                                  // [:8]: the agent cell is allowed
        if ((int)fr > 0) o.cell[A] = nth_free_cell(o, 0xFFFFFFFFu, rng.uniform(fr));
    }
    if ((des >> T_MOVE_HAMMER) & 1u) {                                              // ray.py:287-297
        const uint32_t fr = HW - count_objects(o);
        if ((int)fr > 0) o.cell[HM] = nth_free_cell(o, 0xFFFFFFFFu, rng.uniform(fr));
    }
    const uint32_t ar = anew / W;
    agent = (agent & 0xFFFF0000u) | ar | ((anew - ar * W) << 8);
}

// grid tile (16-byte chunks, one per lane per pass) from an object list
__device__ __forceinline__ void tile_from_objects(const Sparse8& o, int nchunk, uint8_t* tile) {
    for (int ch = lane_id(); ch < nchunk; ch += 32) {
        uint32_t w0 = 0, w1 = 0, w2 = 0, w3 = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if ((int)(o.cell[k] >> 4) == ch && o.code[k] != 0) {
                const uint32_t v = o.code[k] << (8 * (o.cell[k] & 3));
                const int wi = (o.cell[k] >> 2) & 3;
                w0 |= wi == 0 ? v : 0; w1 |= wi == 1 ? v : 0; w2 |= wi == 2 ? v : 0; w3 |= wi == 3 ? v : 0;
            }
        }
        reinterpret_cast<uint4*>(tile)[ch] = make_uint4(w0, w1, w2, w3);
    }
    __syncwarp();
}

// ------------------------------------------------------------------------------------------------------
// render(state): ray.py:442-486.  Expands `nbands` cell rows starting at `band0` of the shared-memory grid `sg`
// into the shared-memory frame chunk `frame` (uint32 words; a pixel row is 3*W words = 12 bytes per cell).
// A thread owns one cell: colour LUT -> the three 32-bit words of its 4-pixel RGB span -> 4 pixel rows; the
// owner of the agent cell patches rows 1,2 itself (2x2 white block :483, bottom row = held colour :484-486),
// so no second pass / barrier is needed.  Lanes hit consecutive cells => word stride 3, which is conflict-free only while a
// warp stays inside one cell row: with W = 21 a warp straddles cell rows (a jump of 4 pixel rows = 9*W words) and ncu counts
// ~49 % of the shared-store wavefronts as conflict replays (profiles/r1f_env_kernel_chained_cfg4_ncu_full.csv).  A
// conflict-free one-band-per-warp mapping was measured slower: the kernel is bound by the HBM write stream.
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void compose_cell(const uint8_t* __restrict__ src, int i, int b, int col, int roww, int acell,
                                             uint32_t hc, uint32_t* __restrict__ frame, const uint32_t* __restrict__ slut) {
    const uint32_t rgb = slut[src[i]];
    const uint32_t w0 = __byte_perm(rgb, 0, 0x0210), w1 = __byte_perm(rgb, 0, 0x1021), w2 = __byte_perm(rgb, 0, 0x2102);
    uint32_t* p = frame + b * (4 * roww) + 3 * col;
    p[0] = w0; p[1] = w1; p[2] = w2;
    p[roww + 0] = w0; p[roww + 1] = w1; p[roww + 2] = w2;
    p[2 * roww + 0] = w0; p[2 * roww + 1] = w1; p[2 * roww + 2] = w2;
    p[3 * roww + 0] = w0; p[3 * roww + 1] = w1; p[3 * roww + 2] = w2;
    if (i == acell) {                                          // one thread per frame: agent overlay (ray.py:483-486)
        p[roww + 0] = w0 | 0xFF000000u; p[roww + 1] = 0xFFFFFFFFu; p[roww + 2] = w2 | 0x000000FFu;
        p[2 * roww + 0] = __byte_perm(w0, hc, 0x4210);
        p[2 * roww + 1] = __byte_perm(hc, 0, 0x1021);
        p[2 * roww + 2] = __byte_perm(w2, hc, 0x3216);
    }
}
__device__ __forceinline__ void compose_bands(const CwConfig& cfg, const uint8_t* __restrict__ sg, uint32_t agent,
                                              int band0, int nbands, uint32_t* __restrict__ frame,
                                              const uint32_t* __restrict__ slut, uint32_t w_magic, int ctid, int cthreads) {
    const int W = cfg.W, roww = 3 * W;
    const int ar = agent & 0xFF, ac = (agent >> 8) & 0xFF, ah = (agent >> 16) & 0xFF;
    const int acell = ar * W + ac - band0 * W;
    const uint32_t hc = ah ? slut[ah] : 0x00FFFFFFu;
    const uint8_t* src = sg + band0 * W;
    const int ncells = nbands * W;
    for (int i = ctid; i < ncells; i += cthreads) {
        const int b = (int)__umulhi((uint32_t)i, w_magic);   // i / W
        compose_cell(src, i, b, i - b * W, roww, acell, hc, frame, slut);
    }
}

// ---- TMA bulk store (shared::cta -> global), sm_90+ : SASS UBLKCP ---------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    uint64_t pol;   // streaming frames: ask L2 to evict them first (measured A/B: +2.9% at 4096 worlds, +0.5% at 131072)
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(smem_u32(ssrc)),
                 "r"(bytes), "l"(pol) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(kPending) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_dyn(int pending) {   // pending in 0..4
    if (pending <= 0) bulk_wait_read<0>();
    else if (pending == 1) bulk_wait_read<1>();
    else if (pending == 2) bulk_wait_read<2>();
    else if (pending == 3) bulk_wait_read<3>();
    else bulk_wait_read<4>();
}
// 16-byte async copy global -> shared through L2 (LDGSTS), for the grid tiles
__device__ __forceinline__ void cp_async16(void* sdst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(sdst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int kPending>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(kPending) : "memory"); }

// ---- named barriers (producer arrives, consumers sync; `count` = all participating threads) ----------------
__device__ __forceinline__ void bar_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

// ---- programmatic dependent launch (sm_90+) -------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

}  // namespace cw
