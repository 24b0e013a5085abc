// cw_internal.h -- entry points shared between the translation units of libcw_b200.so that are NOT part of the C ABI.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cw_b200.h"

namespace cw {

// cw_step_render_chained with a per-world STATUS byte for the host-buffer API (cw_host.cu).  `status` (nullable) points into
// mapped pinned host memory, uint8[N], zeroed by the host before the launch: the moment a world has stepped -- before any frame
// is composed -- the kernel stores 0x80 | success << 1 | done into its byte (success: reward == max_steps), so the polling host
// has reward / done of the whole batch a few microseconds into the launch.  With `status`, `reward` / `done` may be null.
int step_render_chained_notify(const CwConfig* cfg, const CwState* st, const uint8_t* actions, int32_t* reward, uint8_t* done,
                               uint8_t* obs, uint8_t* goal_obs, uint8_t* init_obs, int64_t* stats, int flags, uint32_t* chain,
                               int chain_pos, int obs_ring, uint8_t* status, void* stream);

// ---- pipelined host transport (device consumer, small batches) -------------------------------------------
// One host-driven step = two launches on two streams.  step_snap_launch (thread per world): status bytes to the host, live state
// advanced in place, then the state the frames need copied into a snapshot slot and published per 32 worlds (`epoch`, one word per
// warp, value `seq`) -- to the same warp of the NEXT step launch (no step launch waits for its predecessor's completion) and to
// the render launch of this step.  render_pipe_launch (the fused kernel's render half, a member of a chain like cw_step_render_chained): waits per group
// for the snapshot of its worlds, writes the frames, and marks the slot consumed (`*slot_free = seq`).  The step launch of step
// seq = k + kPipeSlots reuses the slot and therefore waits for `*slot_free` to reach `slot_want` = k (0: no wait).
constexpr int kPipeSlots = 4;
struct PipeSnap {
    uint8_t* grid;      // uint8[N][cell_stride]: the grid after the step (of the NEW episode for a world that finished)
    uint4* meta;        // [N] {agent, agent of the imagined goal state, flags, 0}; flags: 1 = re-seeded in this step,
                        //     bits 8.. = length of the episode that ended (saturated at 255)
    uint8_t* goal;      // uint8[N][cell_stride]: imagined goal state, written for re-seeded worlds only
};
// actions: `actions_host` (nullable) is readable by the calling thread -- batches of <= 4096 worlds then carry their actions in the
// kernel parameters; otherwise `actions_dev` (device-readable, e.g. the device alias of mapped pinned memory) is read by the kernel.
int step_snap_launch(const CwConfig* cfg, const CwState* st, const uint8_t* actions_dev, const uint8_t* actions_host, uint8_t* status,
                     const PipeSnap* snap, uint32_t* epoch, uint32_t seq, const uint32_t* slot_free, uint32_t slot_want,
                     int64_t* stats, int flags, void* stream);
int render_pipe_launch(const CwConfig* cfg, int64_t n, const PipeSnap* snap, uint8_t* obs, uint8_t* goal_obs, uint32_t* chain, int chain_pos,
                       int obs_ring, const uint32_t* epoch, uint32_t seq, uint32_t* slot_free, void* stream);

}  // namespace cw
