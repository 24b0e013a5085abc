/* cw_b200.h -- C ABI of the B200-native batched CraftingWorld hot path (libcw_b200.so).
 *
 * The reference (lauradarcy/gym-craftingworld) is pure Python and has no FFI of its own; its boundary is the Gym
 * Env protocol of CraftingWorldEnvRay (gym_craftingworld/envs/craftingworld_ray.py, "ray.py" below).  Each entry
 * point here replaces the per-env Python method(s) cited beside it, for N independent worlds per call.  Host side:
 * gym_craftingworld_b200/env.py (same method names / argument meaning as the reference class) binds these with
 * ctypes; INTEGRATION.md shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every `uint8_t* / uint32_t* / int32_t* / int64_t*` below that is not marked
 *     "host" is a DEVICE pointer owned by the caller (PyTorch in our host layer); the library never allocates,
 *     frees or retains them (the cw_host_* family is the exception: it owns its own device + pinned buffers).
 *   - alignment: grid / init_grid / goal_grid rows and every frame buffer (obs, goal_obs, init_obs) must start on a
 *     16-byte boundary (tiles move as 16-byte chunks, frames leave through TMA bulk stores); word arrays on their natural
 *     alignment.  cudaMalloc / PyTorch allocations satisfy this.  cw_onehot / cw_render_alt accept any 4- / 2-byte aligned
 *     output (cw_onehot even an unaligned one, through a slower byte-wise path).
 *   - `stream` is a cudaStream_t passed as void*; all device entry points are asynchronous on it and never
 *     synchronise the host.
 *   - return value: 0 on success, a positive cudaError_t, or a negative CW_E_* argument error. Never throws.
 *   - out-of-range actions (>5) are a defined no-op that still advances step_num (the reference raises
 *     IndexError, ray.py:308); the Python layer can validate in debug mode.
 *
 * Device state (SoA, all integer), N = state->n worlds:
 *   grid, init_grid  uint8 [N][cell_stride]  cell = row*W + col, cell_stride = roundup(H*W,16);
 *                                            code 0 empty, k+1 = OBJECTS[k] (ray.py:21): 1 sticks 2 axe 3 hammer
 *                                            4 rock 5 tree 6 bread 7 house 8 wheat.  init_grid = INIT_OBS_VECTOR
 *                                            object codes (ray.py:183), only read by the Move* predicates.
 *   agent            uint32[N]               row | col<<8 | hold<<16   hold: 0 none 1 sticks 2 axe 3 hammer
 *   goal             uint32[N]               achieved | desired<<16    bit i = TASK_LIST[i] (ray.py:40-41)
 *   t                int32 [N]               step_num (ray.py:203, 309)
 *   episode          uint32[N]               resets performed so far = Philox counter word of the next reset
 *   obs, goal_obs    uint8 [N][4H][4W][3]    RGB frames (ray.py:442-486); goal_obs = imagine_obs (ray.py:220-299)
 *   stats            int64 [CW_STATS_REPLICAS][CW_STATS_LEN]  (sum over replicas:) 0 episodes 1 successes 2 return_sum 3 length_sum
 *                                            4..12 achieved-skill counts 13..21 desired-skill counts (finished eps)
 */
#ifndef CW_B200_H
#define CW_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CW_ABI_VERSION 4
#define CW_STATS_LEN 24
#define CW_STATS_REPLICAS 16 /* the stats buffer is int64[CW_STATS_REPLICAS][CW_STATS_LEN]: finished episodes are added to
                                replica (block index % 16) so same-address atomics do not serialise; consumers sum the replicas */
#define CW_MAX_SIDE 64 /* 2 <= H, W <= 64 (cell_stride <= 4096) */
#define CW_CHAIN_MAX_POS 1024 /* chained launches: positions 0..1023; the chain buffer is uint32[CW_CHAIN_MAX_POS + N] */
#define CW_FRESH_WORDS 18 /* delta transport: uint32 words of a re-seeded world's sparse record (see cw_step_delta) */

/* argument errors */
#define CW_E_BADCONFIG (-1)
#define CW_E_NULLPTR (-2)
#define CW_E_BADFLAGS (-3)
#define CW_E_BADHANDLE (-4)

/* flags */
#define CW_F_AUTO_RESET 1 /* on done: add the episode to stats, Philox-reset the world in the same launch; the
                             returned reward/done are the finished episode's, state/obs the new episode's */
#define CW_F_DEFER_RESET 8 /* internal (cw_step_render_edit): count and report a finished world but do not re-seed it in the step launch */
#define CW_F_HOST_ACTIONS 16 /* cw_step_delta only: `actions` is ALSO readable by the calling CPU thread (mapped pinned host memory);
                                for batches of <= 4096 worlds the library then ships the actions inside the kernel parameters, which
                                saves the device a PCIe read on the latency-critical path */
#define CW_F_DELTA_TRANSPORT 2 /* cw_host_create only: keep the caller's frame buffer current by delta records + host-side
                                  patching of the changed cells instead of copying every frame over PCIe */

/* Constructor arguments of the reference that reach the hot path (ray.py:59-83). */
typedef struct CwConfig {
    int32_t H, W;            /* size; square upstream (non-square is broken there) */
    int32_t cell_stride;     /* roundup(H*W, 16) */
    int32_t max_steps;       /* MAX_STEPS, also the success reward (ray.py:46, 759) */
    int32_t subset_reward;   /* reward_style is not None -> compute_reward_subset (ray.py:71-74, 763-767) */
    int32_t stacking;        /* ray.py:83, 169 */
    int32_t n_selected;      /* len(selected_tasks) in 1..9 */
    int32_t number_of_tasks; /* ray.py:79-81, in 1..n_selected */
    uint8_t selected[16];    /* task bit of each selected task: task_list.index(selected_tasks[i]) (ray.py:174) */
} CwConfig;

typedef struct CwState {
    uint8_t* grid;
    uint8_t* init_grid;
    uint32_t* agent;
    uint32_t* goal;
    int32_t* t;
    uint32_t* episode;
    int64_t n;            /* worlds in this slice */
    uint64_t seed;        /* Philox key */
    uint64_t env_id_base; /* global id of world 0 of this slice: streams are keyed by GLOBAL id, so results do
                             not depend on how worlds are sharded over ranks */
    /* fixed_init_state pool (ray.py:116-118, 149-154, 630-644): when n_fixed > 0 a reset draws
     * uniform(n_fixed) and copies that pre-sampled world instead of sampling a new placement */
    const uint8_t* fixed_grid;   /* uint8 [n_fixed][cell_stride], nullable */
    const uint32_t* fixed_agent; /* uint32[n_fixed] */
    int64_t n_fixed;
    /* optional compact outputs of reset() for the one-hot observation family (CraftingWorldEnvOneHot,
     * carftingworld_onehot.py:203, 310): the imagined goal STATE and the agent word of INIT_OBS. Nullable; written by
     * cw_reset / cw_step_render for every world they (re)seed. */
    uint8_t* goal_grid;   /* uint8 [N][cell_stride]: imagine_obs final_state, object codes */
    uint32_t* goal_agent; /* uint32[N]: agent word of the imagined state */
    uint32_t* init_agent; /* uint32[N]: agent word at reset (INIT_OBS_VECTOR's agent channel) */
    /* optional, compact step path (cw_step / cw_rollout with CW_F_AUTO_RESET, no fixed pool): pre-drawn reset records.  What a
     * reset draws depends only on (seed, global id, episode), so it is drawn ahead of time by extra CTAs of the step launches and
     * a finished world is re-seeded by a copy instead of the Philox sampling (results are bit-identical; a missing / stale record
     * falls back to the inline draw).  Both nullable (then every reset draws inline); zero both once, then call
     * cw_prefill_resets after every cw_reset / change of seed. */
    uint32_t* reset_rec;  /* uint32[N][8], 32-byte aligned: two 16-byte halves, each {episode tag, ...placement} (see cw_kernels.cu) */
    uint32_t* reset_list; /* uint32[4 + 2N]: queue of worlds whose next record is due (tail, limit, head, ticket, ring) */
} CwState;

int cw_abi_version(void);
const char* cw_error_string(int code);

/* reset(): ray.py:156-218 -- task sampling (169-174), sample_state (599-628), INIT copy (183), counters (203);
 * goal_obs (nullable) receives imagine_obs (220-299), obs (nullable) the first frame (192) and init_obs
 * (nullable, needs obs) a second copy of it (INIT_OBS, 193).
 * mask (nullable, device uint8[N]): reset only worlds with mask[n] != 0; others are untouched (and not rendered).
 * RNG: Philox4x32-10, key = seed, counter = (global env id, episode[n], block); episode[n] += 1. */
int cw_reset(const CwConfig* cfg, const CwState* st, const uint8_t* mask, uint8_t* obs, uint8_t* goal_obs,
             uint8_t* init_obs, void* stream);

/* step(action) without pixels: ray.py:301-378 (pickup 314-327, drop 329-341, __move_agent 380-440 with the Coord
 * clamp coordinates.py:22-35, eval_task_edit 646-703, compute_reward_equal/_subset 747-767, done 367).
 * One thread per world.  reward int32[N] (-1 or max_steps), done uint8[N]; stats nullable. */
int cw_step(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
            int64_t* stats, int flags, void* stream);

/* render(state): ray.py:442-486 (== the incremental render_edit 522-557 on every reachable state). */
int cw_render(const CwConfig* cfg, const uint8_t* grid, const uint32_t* agent, uint8_t* obs, int64_t n,
              void* stream);

/* step + (auto-reset) + render fused in one launch: what one reference `obs, r, d, info = env.step(a)` does
 * (ray.py:301-378 incl. render_edit 358), for N worlds.  goal_obs / init_obs (nullable) are rewritten only for
 * worlds that auto-reset in this call (desired_goal and init_observation of the new episode, ray.py:191-196).
 * obs may be NULL (no pixels: step + auto-reset + the compact goal outputs of CwState only). */
int cw_step_render(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
                   uint8_t* obs, uint8_t* goal_obs, uint8_t* init_obs, int64_t* stats, int flags, void* stream);

/* step + auto-reset with INCREMENTAL rendering -- the reference's own algorithm (`render_edit`, ray.py:522-557, called from
 * step at ray.py:358): `obs` is a persistent frame buffer (the same pointer every call, last written by cw_reset / cw_render /
 * this function) and only the <= 2 cells a step changes are rewritten in it, by the thread that steps the world; worlds that
 * finish are queued and re-seeded by a second launch (one world per CTA) that renders their new obs / goal_obs / init_obs frames.
 * After the call `obs` holds exactly what cw_step_render would have written.  HBM traffic per env-step drops from 48*H*W bytes
 * to a few sectors, so this path is latency- not bandwidth-bound.
 * `scratch` (device uint32[N + 2], zeroed once by the caller, needed with CW_F_AUTO_RESET): the work list of finished worlds
 * that links the two launches; the library leaves it empty again after every call. */
int cw_step_render_edit(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
                        uint8_t* obs, uint8_t* goal_obs, uint8_t* init_obs, int64_t* stats, int flags, uint32_t* scratch,
                        void* stream);

/* cw_step_render for an OPEN-LOOP run of steps (action tape known in advance, e.g. a CUDA graph of K steps): the same
 * launch per step, but consecutive launches are chained by per-group dataflow instead of whole-grid dependencies.
 *   chain      device uint32[CW_CHAIN_MAX_POS + N], owned by the caller, used only by the chain
 *   chain_pos  0 opens a chain: an ordinary launch (ordered after everything earlier in the stream) that also clears
 *              `chain`.  chain_pos = i > 0: the caller guarantees that the operation immediately before it in `stream` is
 *              the launch with chain_pos = i-1 of the same chain (same cfg / state / N / flags / stats / reward / done /
 *              goal_obs / init_obs) and that `actions` (and every other input) was final before the chain opened.  Such a
 *              launch does not wait for its predecessor GRID: the CTA that steps a group of worlds waits only for the
 *              predecessor's state of that group, so composing step i overlaps the draining frame stores of step i-1.
 *   obs_ring   the `obs` pointers of the chain rotate over this many distinct buffers (1 = the same buffer every step):
 *              position i stores its first frame only after position i-obs_ring has completed.
 * Grids still complete in stream order, so anything after the chain in the stream sees all of it.  Results are
 * identical to the same sequence of cw_step_render calls. */
int cw_step_render_chained(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
                           uint8_t* obs, uint8_t* goal_obs, uint8_t* init_obs, int64_t* stats, int flags, uint32_t* chain,
                           int chain_pos, int obs_ring, void* stream);

/* step + auto-reset WITHOUT device frames, for a host-side frame mirror ("delta transport"): instead of 48*H*W bytes of
 * pixels per world, each world gets one PRE-DIGESTED 16-byte record, written with ONE 16-byte store
 *     delta[n] = { agent, goal, digest, reward }
 *     digest   = orow | ocol<<6 | ocode<<12 | ncode<<16 | objchg<<20 | flags<<24
 *                (orow, ocol) where the agent stood before the step; ocode / ncode the object codes -- after the step -- of that
 *                cell and of the cell it stands on now (`agent`); objchg: the object under the agent changed (pickup, drop, a
 *                transforming move); flags: 1 done, 2 fresh (re-seeded), bits 2..7 = `seq` (0..63), the caller's tag.
 * Everything render_edit (ray.py:522-557) needs is in the record, so the consumer keeps no copy of the grid.  A world
 * re-seeded in this call additionally gets (written and fenced system-wide BEFORE its delta record, so a consumer that polls
 * the tag finds them complete)
 *     fresh[n][0..7]  = cell | code<<16 of its 8 objects        fresh[n][8..15] = same for the imagined goal state
 *     fresh[n][16]    = agent word of the imagined goal state
 * `delta` / `fresh` may point into mapped pinned HOST memory (zero-copy).  st->goal_grid must be NULL (the compact goal state
 * is maintained by cw_reset / cw_step_render only). */
int cw_step_delta(const CwConfig* cfg, const CwState* st, const uint8_t* actions, void* delta /* uint4[N] */,
                  uint32_t* fresh /* [N][CW_FRESH_WORDS] */, int64_t* stats, int flags, int seq, void* stream);

/* K consecutive steps in one launch (open-loop action tape actions[K][N]); reward/done [K][N] nullable.
 * Same per-step semantics as cw_step. */
int cw_rollout(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
               int64_t* stats, int K, int flags, void* stream);

/* cw_step for an OPEN-LOOP run of single-step launches (action tape known in advance, e.g. a CUDA graph of K steps whose rewards
 * feed something between the launches' outputs, or a tape too long for one cw_rollout buffer): the same step per launch, but
 * consecutive launches are linked per WARP of 32 worlds by dataflow instead of whole-grid dependencies -- a dependent launch
 * otherwise waits for its predecessor's slowest warp and refill CTAs plus a 2-3 us grid hand-over, more than the step itself takes.
 *   chain      device uint32[CW_CHAIN_MAX_POS + ceil(N / 32)] (a cw_step_render_chained buffer is large enough), used only by the chain
 *   chain_pos  0 opens a chain (an ordinary launch that also clears `chain`); i > 0: the operation immediately before it in
 *              `stream` is position i-1 of the same chain (same cfg / state / N / flags / stats), all inputs final before it opened.
 * reward / done (nullable) int32[N] / uint8[N] of this step.  With pre-drawn reset records (st->reset_rec; st->reset_list is not
 * used here) a finished world is re-seeded by a copy and its NEXT record is drawn by the same warp after it has released its
 * successor; without them it is re-seeded inline, which delays that warp only.  Results equal the same sequence of cw_step calls. */
int cw_step_chained(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
                    int64_t* stats, int flags, uint32_t* chain, int chain_pos, void* stream);

/* Draw the NEXT reset of every world into st->reset_rec (one warp per world) and empty st->reset_list: call after cw_reset,
 * cw_host-style state injection that changes episode counters, or a change of seed.  Needs both buffers and n_fixed == 0. */
int cw_prefill_resets(const CwConfig* cfg, const CwState* st, void* stream);

/* imagine_obs on the CURRENT state (ray.py:220-299) without resetting: goal image of each world using the stream
 * (seed, global id, episode[n]).  For states injected with load_state. */
int cw_imagine(const CwConfig* cfg, const CwState* st, uint8_t* goal_obs, void* stream);

/* A stand-in device-side CONSUMER of the frames, for closed-loop measurements and tests (the reference has no counterpart: its
 * consumer is the user's policy): reads every byte of each world's frame obs[n] (uint8[4H][4W][3], 16-byte aligned) and derives
 * that world's next action from it,  h = sum_i word_i * (2 i + 1) over the frame's uint32 words (wrap-around),
 * actions[n] = ((h ^ h >> 16) & 0xFFFF) % 6.  Step k+1 then depends on frame k the way it does under a policy network. */
int cw_frame_policy(const CwConfig* cfg, const uint8_t* obs, int64_t n, uint8_t* actions, void* stream);

/* One-hot observation_vector (ray.py:94-98, 605-613; the obs of CraftingWorldEnvOneHot): uint8[N][H][W][12],
 * channels 0..7 objects, 8 agent, 9..11 holding sticks/axe/hammer (at the agent cell). */
int cw_onehot(const CwConfig* cfg, const uint8_t* grid, const uint32_t* agent, uint8_t* onehot, int64_t n,
              void* stream);

/* AltObs renderer (craftingworld_altobs.py:489-548, unregistered upstream): int16[N][3H+3][3W][3]; 3x3 sub-pixels per
 * cell, sub-pixel k lit with CPV_COLORS[k] x multiplicity of channel k (a held item adds to channels 0..2 at the agent
 * cell, so values reach 510 -- hence int16), plus a 3-row status strip (255 at columns 3..5 while holding). */
int cw_render_alt(const CwConfig* cfg, const uint8_t* grid, const uint32_t* agent, int16_t* obs, int64_t n,
                  void* stream);

/* ---- host-buffer API: the env behind an opaque handle, all arguments HOST pointers ----------------------
 * The drop-in for a host-language caller without device memory of its own: the library owns the device state, pinned
 * staging buffers and streams.  One cw_host_step = one reference-style `obs, reward, done, info = env.step(action)`
 * (ray.py:301-378) for N worlds.  Calls on one handle are not re-entrant (like the reference env object).
 *
 * Where the frames go is chosen per handle / per call:
 *   obs_host == NULL            DEVICE CONSUMER: the frames are produced in HBM (rotating buffers, cw_host_device_state); only
 *                               actions (in) and reward / done (out) cross PCIe, through mapped pinned memory (one status byte
 *                               per world comes back).  The call returns as soon as every world's reward / done is in host
 *                               memory -- the frames of this step may still be on their way on the handle's stream
 *                               (cw_host_stream; cw_host_sync waits for them).  Up to 16 384 worlds a step is two launches on two
 *                               streams (a thread-per-world step launch, and the render launch of the state snapshot it
 *                               publishes), so step k+1 never waits for the frames of step k; larger batches and cw_host_step_many
 *                               use one fused launch per step, chained (cw_step_render_chained).
 *   CW_F_DELTA_TRANSPORT handle HOST FRAMES by delta records: obs_host (and the goal buffer given to cw_host_reset) are
 *                               persistent mirrors owned by the caller, pass the same pointers every call; the library patches
 *                               them in place from 16-byte records (a different obs pointer triggers one full refresh).  After
 *                               every call they hold exactly the frames a full device render + copy would have produced.
 *                               Such a handle takes delta steps only (obs_host == NULL is CW_E_BADCONFIG).
 *   otherwise                   HOST FRAMES by copy: every rendered frame crosses PCIe (sliced over two streams). */
typedef struct CwHostEnv CwHostEnv;
int cw_host_create(const CwConfig* cfg, int64_t n, int device, uint64_t seed, uint64_t env_id_base, int flags,
                   CwHostEnv** out);
int cw_host_reset(CwHostEnv* env, uint8_t* obs_host /*nullable*/, uint8_t* goal_obs_host /*nullable*/);
/* Declare the caller's persistent action array (nullable to undeclare).  If it is page-locked (cudaHostAlloc, cudaHostRegister,
 * a pinned torch tensor) the device reads it IN PLACE whenever cw_host_step is called with exactly this pointer; the caller keeps
 * it allocated and page-locked until the next cw_host_bind_actions or cw_host_destroy.  Any other action pointer (or a pageable
 * one) is staged through the handle's own pinned buffer on every call -- the library never assumes a buffer is page-locked
 * because it once was.  reward / done / frame arrays may be any host memory. */
int cw_host_bind_actions(CwHostEnv* env, const uint8_t* actions_host);
int cw_host_step(CwHostEnv* env, const uint8_t* actions_host, int32_t* reward_host, uint8_t* done_host,
                 uint8_t* obs_host /*nullable: device consumer*/);
/* K consecutive steps on an open-loop action tape actions_host[K][N]; reward_host / done_host are [K][N].  Device consumer
 * (obs_host == NULL): K chained launches enqueued back to back and ONE wait -- the round trip is paid once per K steps; with
 * obs_host the call is K cw_host_step calls (the frames after the last step remain). */
int cw_host_step_many(CwHostEnv* env, const uint8_t* actions_host, int K, int32_t* reward_host, uint8_t* done_host,
                      uint8_t* obs_host /*nullable*/);
/* Inject state (each array nullable = leave as is): grid uint8[N][cell_stride] (also becomes INIT_OBS_VECTOR, ray.py:183),
 * agent / goal words uint32[N], step counters int32[N]; obs_host (nullable) receives the frames of the new state (goal frames
 * stay those of the last reset).  For parity tests (reference-generated states) and staggered starts. */
int cw_host_load_state(CwHostEnv* env, const uint8_t* grid_host, const uint32_t* agent_host, const uint32_t* goal_host,
                       const int32_t* t_host, uint8_t* obs_host /*nullable*/);
int cw_host_stats(CwHostEnv* env, int64_t* stats_host /*[CW_STATS_LEN]*/);
/* device pointers of the handle's state and of the frame buffer holding the CURRENT observation, for callers that DO have a
 * device-side consumer (e.g. a policy): order the consumer after the env kernels through cw_host_stream (a cudaStream_t, the
 * stream the frames are written on).  The frame buffer is one of four that rotate: work enqueued on that stream before the next
 * cw_host_step reads complete frames of this step, and the buffer is not written again for three more steps.  The STATE arrays are
 * advanced by the next cw_host_step as soon as it is called (small batches step on a second stream of the handle): read them
 * between cw_host_sync and the next step. */
int cw_host_device_state(CwHostEnv* env, CwState* out_state, uint8_t** out_obs);
int cw_host_stream(CwHostEnv* env, void** out_stream);
/* copy the CURRENT device frames (and the goal frames) to host memory, each nullable; synchronises the handle's stream.
 * For inspection / recording by a caller whose steps leave the frames on the device (not for delta handles). */
int cw_host_fetch_frames(CwHostEnv* env, uint8_t* obs_host, uint8_t* goal_obs_host);
int cw_host_sync(CwHostEnv* env); /* wait for everything the handle has enqueued (frames of the last step included) */
int cw_host_destroy(CwHostEnv* env);

#ifdef __cplusplus
}
#endif
#endif /* CW_B200_H */
