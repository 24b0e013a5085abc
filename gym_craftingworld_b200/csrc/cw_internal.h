// cw_internal.h -- entry points shared between the translation units of libcw_b200.so that are NOT part of the C ABI.
#pragma once
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

}  // namespace cw
