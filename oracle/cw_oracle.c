/* TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference CraftingWorld hot path.
 *
 * Not product code: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, as the checker or as the timed CPU baseline.  It follows
 * gym_craftingworld/envs/craftingworld_ray.py ("ray.py") and envs/coordinates.py of the reference; each
 * function cites the lines it restates.  It is pinned (tests/test_oracle_golden.py) against traces produced
 * by the unmodified reference (the .npz files under tests/golden) and against oracle/compact.py.
 *
 * State layout (shared with the CUDA library so the same arrays can be fed to both):
 *   grid, init_grid  uint8 [N][cell_stride]   cell = r*W + c; code 0 empty, k+1 = OBJECTS[k] (ray.py:21)
 *   agent            uint32[N]                r | c<<8 | hold<<16      (hold 0 none 1 sticks 2 axe 3 hammer)
 *   goal             uint32[N]                achieved | desired<<16   (bit i = TASK_LIST[i], ray.py:40-41)
 *   t                int32 [N]                step_num (ray.py:203, 309)
 *   episode          uint32[N]                resets performed so far (Philox counter word)
 *   obs              uint8 [N][4H][4W][3]
 *   stats            int64 [24]               0 episodes 1 successes 2 return_sum 3 length_sum
 *                                             4..12 achieved-skill counts 13..21 desired-skill counts
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { EMPTY, STICKS, AXE, HAMMER, ROCK, TREE, BREAD, HOUSE, WHEAT };
enum { T_MAKE_BREAD, T_EAT_BREAD, T_BUILD_HOUSE, T_CHOP_TREE, T_CHOP_ROCK, T_GO_TO_HOUSE, T_MOVE_AXE,
       T_MOVE_HAMMER, T_MOVE_STICKS };

typedef struct {
    int32_t H, W, cell_stride, max_steps;
    int32_t subset_reward;   /* reward_style is not None (ray.py:71-74) */
    int32_t stacking;        /* ray.py:83, 169 */
    int32_t n_selected;      /* len(selected_tasks) */
    int32_t number_of_tasks; /* ray.py:79-81 */
    uint8_t selected[16];    /* task bit of each selected task (ray.py:174) */
} CwoConfig;

/* COLORS_N, ray.py:28-30 */
static const uint8_t LUT[9][3] = {{0, 0, 0},       {110, 69, 39},  {255, 105, 180}, {100, 100, 200}, {100, 100, 100},
                                  {0, 128, 0},     {205, 133, 63}, {197, 91, 97},   {240, 230, 140}};

/* ------------------------------------------------------------------ Philox4x32-10 (Random123) ---- */
void cwo_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int i = 0; i < 10; i++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

typedef struct { uint32_t key[2], ctr[4], buf[4]; int pos; } Stream;

static void stream_init(Stream *s, uint64_t seed, uint64_t env_id, uint32_t episode) {
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->ctr[0] = (uint32_t)env_id; s->ctr[1] = (uint32_t)(env_id >> 32); s->ctr[2] = episode; s->ctr[3] = 0;
    s->pos = 4;
}
static uint32_t next32(Stream *s) {
    if (s->pos == 4) { cwo_philox4x32_10(s->ctr, s->key, s->buf); s->ctr[3]++; s->pos = 0; }
    return s->buf[s->pos++];
}
/* unbiased integer in [0,n): Lemire multiply-shift with rejection */
static uint32_t uniform(Stream *s, uint32_t n) {
    uint64_t m = (uint64_t)next32(s) * n;
    uint32_t lo = (uint32_t)m;
    if (lo < n) {
        uint32_t thresh = (uint32_t)(0u - n) % n;
        while (lo < thresh) { m = (uint64_t)next32(s) * n; lo = (uint32_t)m; }
    }
    return (uint32_t)(m >> 32);
}
/* test hook: first `count` uniform(n) draws of a stream */
void cwo_stream_uniform(uint64_t seed, uint64_t env_id, uint32_t episode, uint32_t n, int count, uint32_t *out) {
    Stream s; stream_init(&s, seed, env_id, episode);
    for (int i = 0; i < count; i++) out[i] = uniform(&s, n);
}

/* ------------------------------------------------------------------ step ------------------------- */
static inline uint32_t setbit(uint32_t m, int bit, int on) { return on ? (m | (1u << bit)) : (m & ~(1u << bit)); }

/* ray.py:301-378 (+380-440 move, 646-703 task eval, 747-767 reward, coordinates.py:22-35 clamp).
 * Returns reward; *done_out, *changed_out, cells touched in chg[2] (nchg) for the incremental renderer. */
static int32_t step_one(const CwoConfig *cfg, uint8_t *g, const uint8_t *ig, uint32_t *agent, uint32_t *goal,
                        int32_t *t, int a, uint8_t *done_out, int *nchg, int chg[2]) {
    const int W = cfg->W, H = cfg->H, M = cfg->max_steps;
    int r = *agent & 0xFF, c = (*agent >> 8) & 0xFF, h = (*agent >> 16) & 0xFF;
    uint32_t ach = *goal & 0xFFFF, des = *goal >> 16;
    int changed = 1;
    *t += 1;                                                                     /* ray.py:309 */
    *nchg = 0;
    if (a == 4) {                                                                /* pickup ray.py:314-327 */
        int here = g[r * W + c];
        if (here < STICKS || here > HAMMER || h != 0) changed = 0;               /* ray.py:317-322 */
        else { h = here; g[r * W + c] = EMPTY; chg[(*nchg)++] = r * W + c; }     /* ray.py:326-327 */
    } else if (a == 5) {                                                         /* drop ray.py:329-341 */
        if (h == 0 || g[r * W + c] != EMPTY) changed = 0;                        /* ray.py:332-335 */
        else { g[r * W + c] = (uint8_t)h; h = 0; chg[(*nchg)++] = r * W + c; }   /* ray.py:339-341 */
    } else if (a >= 0 && a <= 3) {                                               /* move ray.py:343-346 */
        static const int DR[4] = {-1, 0, 1, 0}, DC[4] = {0, 1, 0, -1};           /* ray.py:130-131 */
        int old = -1;                                                            /* None -> 100, ray.py:655 */
        int nr = r + DR[a], nc = c + DC[a];
        nr = nr < 0 ? 0 : (nr > H - 1 ? H - 1 : nr);                             /* coordinates.py:22-25 */
        nc = nc < 0 ? 0 : (nc > W - 1 ? W - 1 : nc);
        if (nr == r && nc == c) changed = 0;                                     /* ray.py:395-396 */
        else {
            int T = g[nr * W + nc];
            if ((T == ROCK && h != HAMMER) || (T == TREE && h != AXE)) changed = 0;   /* ray.py:401-405 */
            else {
                chg[(*nchg)++] = r * W + c; chg[(*nchg)++] = nr * W + nc;
                r = nr; c = nc;                                                  /* ray.py:407-410 */
                if (T != EMPTY) old = T;                                         /* ray.py:411, 417-419 */
                if (T == ROCK || T == BREAD) g[r * W + c] = EMPTY;               /* ray.py:423-425 */
                else if (T == TREE) g[r * W + c] = STICKS;                       /* ray.py:426-428 */
                else if (T == STICKS && h == HAMMER) g[r * W + c] = HOUSE;       /* ray.py:429-432 */
                else if (T == WHEAT && h == AXE) g[r * W + c] = BREAD;           /* ray.py:433-438 */
            }
        }
        /* eval_task_edit -- for EVERY move action, successful or not (ray.py:345-346, 646-703) */
        if (old == BREAD) ach |= 1u << T_EAT_BREAD;                              /* ray.py:657-659 */
        else if (old == ROCK) ach |= 1u << T_CHOP_ROCK;                          /* ray.py:660-662 */
        else if (old == TREE) ach |= 1u << T_CHOP_TREE;                          /* ray.py:663-665 */
        ach = setbit(ach, T_GO_TO_HOUSE, g[r * W + c] == HOUSE);                 /* ray.py:668 */
        int ih = ig[r * W + c];
        if (h == STICKS) {                                                       /* ray.py:672-684 */
            int home = ih == STICKS || (ih == TREE && ((ach >> T_CHOP_TREE) & 1));
            ach = setbit(ach, T_MOVE_STICKS, !home);
        } else if (h == AXE) {                                                   /* ray.py:685-693 */
            if (old == WHEAT) ach |= 1u << T_MAKE_BREAD;
            ach = setbit(ach, T_MOVE_AXE, ih != AXE);
        } else if (h == HAMMER) {                                                /* ray.py:694-702 */
            if (old == STICKS) ach |= 1u << T_BUILD_HOUSE;
            ach = setbit(ach, T_MOVE_HAMMER, ih != HAMMER);
        }
    } else changed = 0; /* out-of-range action: defined no-op (the reference raises IndexError, ray.py:308) */
    int32_t reward = -1;                                                         /* ray.py:362-363 */
    if (changed) {                                                               /* ray.py:348, 361 */
        int success = cfg->subset_reward ? ((des & ~ach) == 0)                   /* ray.py:763-767 */
                                         : (ach == des);                         /* ray.py:747-761 */
        if (success) reward = M;
    }
    *done_out = (uint8_t)((*t >= M) || (reward == M));                           /* ray.py:367 */
    *agent = (uint32_t)r | ((uint32_t)c << 8) | ((uint32_t)h << 16);
    *goal = ach | (des << 16);
    return reward;
}

/* ------------------------------------------------------------------ render ----------------------- */
/* render(state): ray.py:442-486 */
static void render_one(const CwoConfig *cfg, const uint8_t *g, uint32_t agent, uint8_t *obs) {
    const int W = cfg->W, H = cfg->H, PW = 4 * W;
    const size_t rowb = (size_t)PW * 3;
    for (int br = 0; br < H; br++) {                       /* colour LUT + x4 upsample, ray.py:477-479 */
        uint8_t *row = obs + (size_t)(4 * br) * rowb;
        for (int bc = 0; bc < W; bc++) {
            const uint8_t *col = LUT[g[br * W + bc]];
            for (int k = 0; k < 4; k++) memcpy(row + (size_t)(4 * bc + k) * 3, col, 3);
        }
        for (int k = 1; k < 4; k++) memcpy(row + k * rowb, row, rowb);   /* the 4 pixel rows of a cell row are equal */
    }
    int r = agent & 0xFF, c = (agent >> 8) & 0xFF, h = (agent >> 16) & 0xFF;
    for (int y = 4 * r + 1; y < 4 * r + 3; y++)
        for (int x = 4 * c + 1; x < 4 * c + 3; x++) memset(obs + ((size_t)y * PW + x) * 3, 255, 3);              /* :483 */
    if (h) for (int x = 4 * c + 1; x < 4 * c + 3; x++) memcpy(obs + ((size_t)(4 * r + 2) * PW + x) * 3, LUT[h], 3); /* :484-486 */
}
/* render_edit(change_idxs): ray.py:522-557 (255 - COLORS_H[h] == COLORS_N[h], ray.py:31) */
static void render_edit_one(const CwoConfig *cfg, const uint8_t *g, uint32_t agent, uint8_t *obs, int nchg, const int chg[2]) {
    const int W = cfg->W, PW = 4 * W;
    int r = agent & 0xFF, c = (agent >> 8) & 0xFF, h = (agent >> 16) & 0xFF;
    for (int i = 0; i < nchg; i++) {
        int x0 = chg[i] / W, y0 = chg[i] % W; /* (row, col) */
        for (int y = 4 * x0; y < 4 * x0 + 4; y++)
            for (int x = 4 * y0; x < 4 * y0 + 4; x++) memcpy(obs + ((size_t)y * PW + x) * 3, LUT[g[chg[i]]], 3);   /* :550-551 */
        if (x0 == r && y0 == c) {                                                                                  /* :553 */
            for (int y = 4 * r + 1; y < 4 * r + 3; y++)
                for (int x = 4 * c + 1; x < 4 * c + 3; x++) memset(obs + ((size_t)y * PW + x) * 3, 255, 3);        /* :555 */
            if (h) for (int x = 4 * c + 1; x < 4 * c + 3; x++) memcpy(obs + ((size_t)(4 * r + 2) * PW + x) * 3, LUT[h], 3); /* :556-557 */
        }
    }
}

/* ------------------------------------------------------------------ reset ------------------------ */
static int nth_cell(const uint8_t *g, int n, int code, int k, int skip) { /* k-th cell == code, row-major (np.where order) */
    for (int i = 0; i < n; i++) if (g[i] == code && i != skip) { if (k == 0) return i; k--; }
    return -1;
}
static int count_cells(const uint8_t *g, int n, int code, int skip) {
    int k = 0; for (int i = 0; i < n; i++) k += (g[i] == code && i != skip); return k;
}
/* imagine_obs: ray.py:220-299.  g is modified in place; *agent may move (GoToHouse). */
static void imagine_one(const CwoConfig *cfg, uint8_t *g, uint32_t *agent, uint32_t des, Stream *s) {
    const int n = cfg->H * cfg->W, W = cfg->W;
    int r = *agent & 0xFF, c = (*agent >> 8) & 0xFF, k, cnt, src, dst;
    if ((des >> T_MAKE_BREAD) & 1) { src = nth_cell(g, n, WHEAT, 0, -1); if (src >= 0) g[src] = BREAD; }          /* :226-231 */
    if ((des >> T_EAT_BREAD) & 1) { cnt = count_cells(g, n, BREAD, -1);                                            /* :232-237 */
        if (cnt) { k = uniform(s, cnt); g[nth_cell(g, n, BREAD, k, -1)] = EMPTY; } }
    if ((des >> T_CHOP_TREE) & 1) { src = nth_cell(g, n, TREE, 0, -1); if (src >= 0) g[src] = STICKS; }           /* :238-243 */
    if ((des >> T_MOVE_STICKS) & 1) { cnt = count_cells(g, n, STICKS, -1);                                         /* :244-257 */
        if (cnt) { k = uniform(s, cnt); int fr = count_cells(g, n, EMPTY, r * W + c);   /* [:9]: agent cell excluded */
            if (fr) { int spot = uniform(s, fr); src = nth_cell(g, n, STICKS, k, -1); dst = nth_cell(g, n, EMPTY, spot, r * W + c);
                      g[src] = EMPTY; g[dst] = STICKS; } } }
    if ((des >> T_BUILD_HOUSE) & 1) { cnt = count_cells(g, n, STICKS, -1);                                         /* :258-264 */
        if (cnt) { k = uniform(s, cnt); g[nth_cell(g, n, STICKS, k, -1)] = HOUSE; } }
    if ((des >> T_CHOP_ROCK) & 1) { src = nth_cell(g, n, ROCK, 0, -1); if (src >= 0) g[src] = EMPTY; }            /* :265-268 */
    if ((des >> T_GO_TO_HOUSE) & 1) { cnt = count_cells(g, n, HOUSE, -1);                                          /* :269-276 */
        if (cnt) { k = uniform(s, cnt); dst = nth_cell(g, n, HOUSE, k, -1); r = dst / W; c = dst % W; } }
    if ((des >> T_MOVE_AXE) & 1) { src = nth_cell(g, n, AXE, 0, -1);                                               /* :277-286 */
        if (src >= 0) { int fr = count_cells(g, n, EMPTY, -1);                           /* [:8]: agent cell allowed */
            if (fr) { int spot = uniform(s, fr); dst = nth_cell(g, n, EMPTY, spot, -1); g[src] = EMPTY; g[dst] = AXE; } } }
    if ((des >> T_MOVE_HAMMER) & 1) { src = nth_cell(g, n, HAMMER, 0, -1);                                         /* :287-297 */
        if (src >= 0) { int fr = count_cells(g, n, EMPTY, -1);
            if (fr) { int spot = uniform(s, fr); dst = nth_cell(g, n, EMPTY, spot, -1); g[src] = EMPTY; g[dst] = HAMMER; } } }
    *agent = (*agent & 0xFFFF0000u) | (uint32_t)r | ((uint32_t)c << 8);
}

/* fixed_init_state pool (ray.py:116-118, 149-154): set once by the test, read by every reset */
static const uint8_t *g_fixed_grid = NULL; static const uint32_t *g_fixed_agent = NULL; static int64_t g_fixed_n = 0;
void cwo_set_fixed_pool(const uint8_t *grid, const uint32_t *agent, int64_t n) { g_fixed_grid = grid; g_fixed_agent = agent; g_fixed_n = n; }

/* reset(): ray.py:156-218 on the Philox stream (draw order: task count, task subset, placement, imagine) */
static void reset_one(const CwoConfig *cfg, uint8_t *g, uint8_t *ig, uint32_t *agent, uint32_t *goal, int32_t *t,
                      uint32_t *episode, uint64_t seed, uint64_t env_id, uint8_t *goal_obs) {
    Stream s; stream_init(&s, seed, env_id, *episode);
    *episode += 1;
    /* task sampling, ray.py:169-174 */
    int n = cfg->stacking ? (int)uniform(&s, cfg->number_of_tasks) + 1 : 1;
    uint8_t sel[16]; memcpy(sel, cfg->selected, 16);
    uint32_t des = 0;
    for (int i = 0; i < n; i++) {
        int j = i + (int)uniform(&s, cfg->n_selected - i);
        uint8_t tmp = sel[i]; sel[i] = sel[j]; sel[j] = tmp;
        des |= 1u << sel[i];
    }
    if (g_fixed_n > 0) {
        /* generate_fixed_initial_state, ray.py:630-644: uniform pick from the pre-sampled pool */
        uint32_t idx = uniform(&s, (uint32_t)g_fixed_n);
        memcpy(g, g_fixed_grid + (size_t)idx * cfg->cell_stride, cfg->cell_stride);
        *agent = g_fixed_agent[idx] & 0xFFFFu;
    } else {
        /* sample_state, ray.py:605-613 */
        int cells[9];
        for (int k = 0; k < 9; k++) {
            int cell, dup;
            do { cell = (int)uniform(&s, cfg->H * cfg->W); dup = 0; for (int q = 0; q < k; q++) dup |= cells[q] == cell; } while (dup);
            cells[k] = cell;
        }
        memset(g, 0, cfg->cell_stride);
        for (int k = 0; k < 8; k++) g[cells[k]] = (uint8_t)(k + 1);
        *agent = (uint32_t)(cells[8] / cfg->W) | ((uint32_t)(cells[8] % cfg->W) << 8);
    }
    memcpy(ig, g, cfg->cell_stride);                                             /* ray.py:183 */
    *goal = des << 16;                                                           /* ray.py:176 */
    *t = 0;                                                                      /* ray.py:203 */
    if (goal_obs) {                                                              /* ray.py:191 */
        uint8_t *tmp = (uint8_t *)malloc(cfg->cell_stride);
        memcpy(tmp, g, cfg->cell_stride);
        uint32_t ag = *agent;
        imagine_one(cfg, tmp, &ag, des, &s);
        render_one(cfg, tmp, ag, goal_obs);
        free(tmp);
    }
}

static void stats_add(const CwoConfig *cfg, int64_t *stats, uint32_t goal, int32_t t, int32_t reward) {
    int success = reward == cfg->max_steps;
    stats[0] += 1; stats[1] += success;
    stats[2] += success ? (int64_t)cfg->max_steps - (t - 1) : -(int64_t)t;
    stats[3] += t;
    for (int i = 0; i < 9; i++) { stats[4 + i] += (goal >> i) & 1; stats[13 + i] += (goal >> (16 + i)) & 1; }
}

/* ------------------------------------------------------------------ exported batch entry points -- */
#define FRAME(cfg) ((size_t)48 * (cfg)->H * (cfg)->W)

void cwo_step(const CwoConfig *cfg, uint8_t *grid, const uint8_t *init_grid, uint32_t *agent, uint32_t *goal,
              int32_t *t, const uint8_t *actions, int32_t *reward, uint8_t *done, int64_t N) {
    int nchg, chg[2];
    for (int64_t n = 0; n < N; n++)
        reward[n] = step_one(cfg, grid + n * cfg->cell_stride, init_grid + n * cfg->cell_stride, agent + n, goal + n,
                             t + n, actions[n], done + n, &nchg, chg);
}

void cwo_render(const CwoConfig *cfg, const uint8_t *grid, const uint32_t *agent, uint8_t *obs, int64_t N) {
    for (int64_t n = 0; n < N; n++) render_one(cfg, grid + n * cfg->cell_stride, agent[n], obs + n * FRAME(cfg));
}

/* mask: NULL = all envs, else reset where mask[n] != 0.  goal_obs may be NULL. */
void cwo_reset(const CwoConfig *cfg, uint8_t *grid, uint8_t *init_grid, uint32_t *agent, uint32_t *goal, int32_t *t,
               uint32_t *episode, const uint8_t *mask, uint64_t seed, uint64_t env_id_base, uint8_t *goal_obs, int64_t N) {
    for (int64_t n = 0; n < N; n++)
        if (!mask || mask[n])
            reset_one(cfg, grid + n * cfg->cell_stride, init_grid + n * cfg->cell_stride, agent + n, goal + n, t + n,
                      episode + n, seed, env_id_base + (uint64_t)n, goal_obs ? goal_obs + n * FRAME(cfg) : NULL);
}

/* imagine_obs on given states (test hook for injected worlds): goal image of env n with stream (seed, id, episode). */
void cwo_imagine(const CwoConfig *cfg, const uint8_t *grid, const uint32_t *agent, const uint32_t *goal,
                 const uint32_t *episode, uint64_t seed, uint64_t env_id_base, uint8_t *out_grid, uint32_t *out_agent, int64_t N) {
    for (int64_t n = 0; n < N; n++) {
        Stream s; stream_init(&s, seed, env_id_base + (uint64_t)n, episode[n]);
        uint8_t *g = out_grid + n * cfg->cell_stride;
        memcpy(g, grid + n * cfg->cell_stride, cfg->cell_stride);
        out_agent[n] = agent[n];
        imagine_one(cfg, g, out_agent + n, goal[n] >> 16, &s);
    }
}

/* One batched env step as the product's fused entry point defines it: step, then (auto_reset && done) ->
 * stats + Philox reset of that env, then render of the (possibly fresh) state.  obs / goal_obs / stats may be NULL. */
void cwo_step_full(const CwoConfig *cfg, uint8_t *grid, uint8_t *init_grid, uint32_t *agent, uint32_t *goal, int32_t *t,
                   uint32_t *episode, const uint8_t *actions, int32_t *reward, uint8_t *done, uint8_t *obs,
                   uint8_t *goal_obs, int64_t *stats, int auto_reset, uint64_t seed, uint64_t env_id_base, int64_t N) {
    int nchg, chg[2];
    for (int64_t n = 0; n < N; n++) {
        uint8_t *g = grid + n * cfg->cell_stride, *ig = init_grid + n * cfg->cell_stride;
        reward[n] = step_one(cfg, g, ig, agent + n, goal + n, t + n, actions[n], done + n, &nchg, chg);
        if (auto_reset && done[n]) {
            if (stats) stats_add(cfg, stats, goal[n], t[n], reward[n]);
            reset_one(cfg, g, ig, agent + n, goal + n, t + n, episode + n, seed, env_id_base + (uint64_t)n,
                      goal_obs ? goal_obs + n * FRAME(cfg) : NULL);
        }
        if (obs) render_one(cfg, g, agent[n], obs + n * FRAME(cfg));
    }
}

/* ------------------------------------------------------------------ CPU baseline loop ------------- */
/* K steps x N envs on `nthreads` pthreads (env slices), auto-reset on, actions u8[K][N].
 * render_mode 0: none; 1: full render every step (what the GPU kernel does); 2: incremental render_edit on the
 * <=2 changed cells + full render after a reset (what the reference does, ray.py:192, 358). */
typedef struct {
    const CwoConfig *cfg; uint8_t *grid, *init_grid; uint32_t *agent, *goal; int32_t *t; uint32_t *episode;
    const uint8_t *actions; int32_t *reward; uint8_t *done; uint8_t *obs; int64_t stats[24];
    int render_mode; uint64_t seed, env_id_base; int64_t n0, n1, N; int K;
} Job;

static void *job_main(void *p) {
    Job *j = (Job *)p; const CwoConfig *cfg = j->cfg; int nchg, chg[2];
    for (int k = 0; k < j->K; k++)
        for (int64_t n = j->n0; n < j->n1; n++) {
            uint8_t *g = j->grid + n * cfg->cell_stride, *ig = j->init_grid + n * cfg->cell_stride;
            uint8_t *o = j->obs ? j->obs + n * FRAME(cfg) : NULL;
            j->reward[n] = step_one(cfg, g, ig, j->agent + n, j->goal + n, j->t + n, j->actions[(size_t)k * j->N + n],
                                    j->done + n, &nchg, chg);
            int fresh = 0;
            if (j->done[n]) {
                stats_add(cfg, j->stats, j->goal[n], j->t[n], j->reward[n]);
                reset_one(cfg, g, ig, j->agent + n, j->goal + n, j->t + n, j->episode + n, j->seed, j->env_id_base + (uint64_t)n, NULL);
                fresh = 1;
            }
            if (o && (j->render_mode == 1 || fresh)) render_one(cfg, g, j->agent[n], o);
            else if (o && j->render_mode == 2) render_edit_one(cfg, g, j->agent[n], o, nchg, chg);
        }
    return NULL;
}

void cwo_run_threads(const CwoConfig *cfg, uint8_t *grid, uint8_t *init_grid, uint32_t *agent, uint32_t *goal, int32_t *t,
                     uint32_t *episode, const uint8_t *actions, int32_t *reward, uint8_t *done, uint8_t *obs, int64_t *stats,
                     int render_mode, uint64_t seed, uint64_t env_id_base, int64_t N, int K, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    Job *jobs = (Job *)calloc(nthreads, sizeof(Job));
    pthread_t *th = (pthread_t *)calloc(nthreads, sizeof(pthread_t));
    for (int i = 0; i < nthreads; i++) {
        Job j = {cfg, grid, init_grid, agent, goal, t, episode, actions, reward, done, render_mode ? obs : NULL, {0},
                 render_mode, seed, env_id_base, N * i / nthreads, N * (i + 1) / nthreads, N, K};
        jobs[i] = j;
        pthread_create(&th[i], NULL, job_main, &jobs[i]);
    }
    for (int i = 0; i < nthreads; i++) {
        pthread_join(th[i], NULL);
        if (stats) for (int q = 0; q < 24; q++) stats[q] += jobs[i].stats[q];
    }
    free(jobs); free(th);
}

int cwo_version(void) { return 1; }
