"""TEST INFRASTRUCTURE ONLY -- NumPy / pure-Python restatement of the reference CraftingWorld hot path.

This is the readable spec the CUDA kernels (and the C restatement ``cw_oracle.c``) are checked against.
It restates ``gym_craftingworld/envs/craftingworld_ray.py`` ("ray.py") and ``envs/coordinates.py`` of the
reference on the compact state encoding; each function cites the reference lines it follows.  It is pinned
against outputs of the unmodified reference (``tests/golden/*.npz``, ``tests/test_oracle_*``).

Encodings (``ray.py:15-21, 40-41, 130-131``)
  cell code g: 0 empty, k+1 = OBJECTS[k]: 1 sticks 2 axe 3 hammer 4 rock 5 tree 6 bread 7 house 8 wheat
  hold h     : 0 none, 1 sticks, 2 axe, 3 hammer            (same code as the cell code of the item)
  action a   : 0 up(-1,0) 1 right(0,+1) 2 down(+1,0) 3 left(0,-1) 4 pickup 5 drop
  task bit i : 0 MakeBread 1 EatBread 2 BuildHouse 3 ChopTree 4 ChopRock 5 GoToHouse 6 MoveAxe
               7 MoveHammer 8 MoveSticks

The reset RNG is NOT the reference's (``gym.utils.seeding`` -> MT19937 ``RandomState``, an un-vendored,
unpinned third-party dependency, ``requirements.txt:1``): BASELINE.json asks for a Philox counter-based
reset, whose *distribution* must match ``sample_state`` / task sampling / ``imagine_obs``.  The Philox
stream layout defined here (``PhiloxStream``) is the spec for ``cw_oracle.c`` and the CUDA reset kernel.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

EMPTY, STICKS, AXE, HAMMER, ROCK, TREE, BREAD, HOUSE, WHEAT = range(9)
UP, RIGHT, DOWN, LEFT, PICKUP, DROP = range(6)
T_MAKE_BREAD, T_EAT_BREAD, T_BUILD_HOUSE, T_CHOP_TREE, T_CHOP_ROCK, T_GO_TO_HOUSE, T_MOVE_AXE, \
    T_MOVE_HAMMER, T_MOVE_STICKS = range(9)
TASK_LIST = ['MakeBread', 'EatBread', 'BuildHouse', 'ChopTree', 'ChopRock', 'GoToHouse', 'MoveAxe',
             'MoveHammer', 'MoveSticks']                                               # ray.py:40-41
# colour LUT by cell code, COLORS_N (ray.py:28-30)
LUT = np.array([(0, 0, 0), (110, 69, 39), (255, 105, 180), (100, 100, 200), (100, 100, 100), (0, 128, 0),
                (205, 133, 63), (197, 91, 97), (240, 230, 140)], dtype=np.uint8)
DR = (-1, 0, 1, 0)   # ray.py:130-131
DC = (0, 1, 0, -1)


@dataclass
class Config:
    """Constructor arguments of the reference that reach the hot path (``ray.py:59-83``)."""
    H: int = 21
    W: int = 21
    max_steps: int = 300
    subset_reward: bool = False                 # reward_style is not None -> compute_reward_subset (ray.py:71-74)
    stacking: bool = True
    selected: tuple = tuple(range(9))           # task-bit index of each selected task (ray.py:174)
    number_of_tasks: int = 9                    # clipped to len(selected) (ray.py:79-81)

    def __post_init__(self):
        self.number_of_tasks = min(int(self.number_of_tasks), len(self.selected))


@dataclass
class EnvState:
    grid: np.ndarray                            # uint8[H,W] object codes
    init_grid: np.ndarray                       # uint8[H,W] object codes of INIT_OBS_VECTOR (ray.py:183)
    r: int
    c: int
    hold: int = 0
    achieved: int = 0                           # 9-bit achieved_goal_vector (ray.py:176)
    desired: int = 0                            # 9-bit desired_goal_vector (ray.py:170-174)
    t: int = 0                                  # step_num (ray.py:203)
    episode: int = 0

    def copy(self):
        return EnvState(self.grid.copy(), self.init_grid.copy(), self.r, self.c, self.hold, self.achieved,
                        self.desired, self.t, self.episode)


# ----------------------------------------------------------------------------------------------------
# step  (ray.py:301-378 + 380-440 + 646-703 + 747-767; coordinates.py:22-35)
# ----------------------------------------------------------------------------------------------------

def step_env(s: EnvState, a: int, cfg: Config):
    """One ``step(action)``.  Mutates ``s``; returns ``(reward, done, changed)``."""
    g = s.grid
    s.t += 1                                                                    # ray.py:309
    changed = True                                                              # ray.py:313
    if a == PICKUP:                                                             # ray.py:314-327
        here = int(g[s.r, s.c])
        if here not in (STICKS, AXE, HAMMER) or s.hold != 0:                    # ray.py:317-322
            changed = False
        else:
            s.hold = here                                                       # ray.py:326
            g[s.r, s.c] = EMPTY                                                 # ray.py:327
    elif a == DROP:                                                             # ray.py:329-341
        if s.hold == 0 or g[s.r, s.c] != EMPTY:                                 # ray.py:332-335
            changed = False
        else:
            g[s.r, s.c] = s.hold                                                # ray.py:339-340
            s.hold = 0                                                          # ray.py:341
    elif 0 <= a <= 3:                                                           # ray.py:343-346
        old = -1                                                                # "None" -> old_object = 100 (ray.py:655)
        nr = max(0, min(s.r + DR[a], cfg.H - 1))                                # coordinates.py:22-25
        nc = max(0, min(s.c + DC[a], cfg.W - 1))
        if nr == s.r and nc == s.c:                                             # ray.py:395-396
            changed = False
        else:
            T = int(g[nr, nc])
            if (T == ROCK and s.hold != HAMMER) or (T == TREE and s.hold != AXE):   # ray.py:401-405
                changed = False
            else:
                s.r, s.c = nr, nc                                               # ray.py:407-410
                if T != EMPTY:                                                  # ray.py:417-419: empty -> None
                    old = T                                                     # ray.py:411
                if T in (ROCK, BREAD):                                          # ray.py:423-425
                    g[nr, nc] = EMPTY
                elif T == TREE:                                                 # ray.py:426-428
                    g[nr, nc] = STICKS
                elif T == STICKS and s.hold == HAMMER:                          # ray.py:429-432
                    g[nr, nc] = HOUSE
                elif T == WHEAT and s.hold == AXE:                              # ray.py:433-438
                    g[nr, nc] = BREAD
        # eval_task_edit: runs for EVERY move action, successful or not (ray.py:345-346, 646-703)
        ach = s.achieved
        if old == BREAD:                                                        # ray.py:657-659
            ach |= 1 << T_EAT_BREAD
        elif old == ROCK:                                                       # ray.py:660-662
            ach |= 1 << T_CHOP_ROCK
        elif old == TREE:                                                       # ray.py:663-665
            ach |= 1 << T_CHOP_TREE
        here = int(g[s.r, s.c])
        ach = _setbit(ach, T_GO_TO_HOUSE, here == HOUSE)                        # ray.py:668 (level-triggered)
        init_here = int(s.init_grid[s.r, s.c])
        if s.hold == STICKS:                                                    # ray.py:672-684
            home = init_here == STICKS or (init_here == TREE and (ach >> T_CHOP_TREE) & 1)
            ach = _setbit(ach, T_MOVE_STICKS, not home)
        elif s.hold == AXE:                                                     # ray.py:685-693
            if old == WHEAT:
                ach |= 1 << T_MAKE_BREAD
            ach = _setbit(ach, T_MOVE_AXE, init_here != AXE)
        elif s.hold == HAMMER:                                                  # ray.py:694-702
            if old == STICKS:
                ach |= 1 << T_BUILD_HOUSE
            ach = _setbit(ach, T_MOVE_HAMMER, init_here != HAMMER)
        s.achieved = ach
    else:
        # the reference raises IndexError for a not in [0,6) (ray.py:308); the batched API defines
        # an out-of-range action as a no-op that still advances step_num (DESIGN.md "Errors").
        changed = False
    if changed:                                                                 # ray.py:348-363
        if cfg.subset_reward:
            success = (s.desired & ~s.achieved) == 0                            # ray.py:763-767
        else:
            success = s.achieved == s.desired                                   # ray.py:747-761
        reward = cfg.max_steps if success else -1
    else:
        reward = -1
    done = s.t >= cfg.max_steps or reward == cfg.max_steps                      # ray.py:367
    return reward, bool(done), changed


def _setbit(mask: int, bit: int, on) -> int:
    return (mask | (1 << bit)) if on else (mask & ~(1 << bit))


# ----------------------------------------------------------------------------------------------------
# render  (ray.py:442-486; incremental form ray.py:522-557 yields the same image on reachable states)
# ----------------------------------------------------------------------------------------------------

def render(grid, r, c, hold):
    """``uint8[4H,4W,3]`` frame: LUT colour x4 upsample, 2x2 white agent block, bottom row = held colour."""
    img = LUT[np.asarray(grid)]                                                 # ray.py:462-477
    img = np.repeat(np.repeat(img, 4, axis=0), 4, axis=1)                       # ray.py:478-479
    img[4 * r + 1:4 * r + 3, 4 * c + 1:4 * c + 3, :] = 255                      # ray.py:483
    if hold != 0:
        img[4 * r + 2, 4 * c + 1:4 * c + 3, :] = LUT[hold]                      # ray.py:484-486
    return img


# AltObs renderer (craftingworld_altobs.py:26-27, 45-51, 489-548): 3x3 sub-pixels per cell, sub-pixel k (row-major)
# lit with CPV_COLORS[k] times the multiplicity of channel k (objects 0..7, agent 8; a held item adds 1 to
# channel 0..2 at the agent cell, so values reach 2 x 255 -> int16), plus a 3-row status strip at the bottom.
CPV = np.array([(45, 82, 160), (255, 102, 102), (204, 204, 0), (211, 211, 211), (34, 133, 34), (0, 215, 255),
                (153, 52, 255), (10, 215, 100), (0, 0, 255)], dtype=np.int16)                # altobs.py:26-27


def render_alt(grid, r, c, hold):
    """``int16[3H+3, 3W, 3]`` AltObs frame of one world (``craftingworld_altobs.py:489-548``)."""
    grid = np.asarray(grid)
    H, W = grid.shape
    m = np.zeros((H, W, 9), np.int16)
    for k in range(8):
        m[:, :, k] = grid == k + 1                                              # objects_new[..., :8]
    m[r, c, 8] = 1                                                              # agent channel
    if hold:
        m[r, c, hold - 1] += 1                                                  # holding_new added onto channels 0..2 (:531-533)
    img = np.zeros((3 * H + 3, 3 * W, 3), np.int16)
    for k in range(9):                                                          # OBJECT_ENCODING_M / COLORS_A_M (:45-51, :540-541)
        img[k // 3:3 * H:3, k % 3::3] = m[:, :, k, None] * CPV[k]
    if hold:
        img[3 * H:, 3:6] = 255                                                  # :543-545
    return img


# ----------------------------------------------------------------------------------------------------
# Philox4x32-10 counter-based stream (D. E. Shaw Research "Random123"; Salmon et al., SC'11)
# ----------------------------------------------------------------------------------------------------
_M0, _M1 = 0xD2511F53, 0xCD9E8D57
_W0, _W1 = 0x9E3779B9, 0xBB67AE85
_MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = (int(x) & _MASK for x in ctr)
    k0, k1 = (int(x) & _MASK for x in key)
    for _ in range(10):
        p0 = _M0 * c0
        p1 = _M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & _MASK, p1 & _MASK, ((p0 >> 32) ^ c3 ^ k1) & _MASK, p0 & _MASK
        k0 = (k0 + _W0) & _MASK
        k1 = (k1 + _W1) & _MASK
    return c0, c1, c2, c3


class PhiloxStream:
    """32-bit draws for (seed, global env id, episode):  key=(seed_lo, seed_hi),
    counter=(env_lo, env_hi, episode, block) with block = 0,1,2,...; words of a block are consumed in order."""

    def __init__(self, seed: int, env_id: int, episode: int):
        self.key = (seed & _MASK, (seed >> 32) & _MASK)
        self.base = (env_id & _MASK, (env_id >> 32) & _MASK, episode & _MASK)
        self.block = 0
        self.buf = ()
        self.pos = 4

    def next32(self) -> int:
        if self.pos == 4:
            self.buf = philox4x32_10(self.base + (self.block,), self.key)
            self.block += 1
            self.pos = 0
        v = self.buf[self.pos]
        self.pos += 1
        return v

    def uniform(self, n: int) -> int:
        """Unbiased integer in [0, n) (Lemire 2019, multiply-shift with rejection)."""
        m = self.next32() * n
        lo = m & _MASK
        if lo < n:
            thresh = ((1 << 32) - n) % n
            while lo < thresh:
                m = self.next32() * n
                lo = m & _MASK
        return m >> 32


# ----------------------------------------------------------------------------------------------------
# reset  (ray.py:156-218, 169-176 task sampling, 599-628 sample_state, 220-299 imagine_obs)
# ----------------------------------------------------------------------------------------------------

def sample_tasks(rng: PhiloxStream, cfg: Config) -> int:
    """desired_goal_vector: n = U{1..number_of_tasks} (1 if not stacking) distinct entries of
    selected_tasks, uniformly (``ray.py:169-174``: randint + first n of a shuffle == partial Fisher-Yates)."""
    n = rng.uniform(cfg.number_of_tasks) + 1 if cfg.stacking else 1
    sel = list(cfg.selected)
    m = len(sel)
    desired = 0
    for i in range(n):
        j = i + rng.uniform(m - i)
        sel[i], sel[j] = sel[j], sel[i]
        desired |= 1 << sel[i]
    return desired


def sample_state(rng: PhiloxStream, cfg: Config):
    """One of each of the 8 objects + the agent on 9 distinct uniformly random cells (``ray.py:605-613``:
    a uniform permutation of the H*W cells puts rows 0..8 of the diag on a uniform ordered 9-tuple)."""
    cells = []
    for _ in range(9):
        while True:
            cell = rng.uniform(cfg.H * cfg.W)
            if cell not in cells:
                break
        cells.append(cell)
    grid = np.zeros((cfg.H, cfg.W), np.uint8)
    for k in range(8):
        grid[cells[k] // cfg.W, cells[k] % cfg.W] = k + 1
    return grid, cells[8] // cfg.W, cells[8] % cfg.W


def _nth(mask2d, k):
    """k-th True cell in row-major order (the order ``np.where`` enumerates, ``ray.py:229-295``)."""
    rr, cc = np.nonzero(mask2d)
    return int(rr[k]), int(cc[k])


def imagine(grid, r, c, hold, desired, rng: PhiloxStream):
    """Goal-state imagination, ``imagine_obs`` (``ray.py:220-299``): apply the desired skills to a copy of
    the initial state in the reference's fixed order with uniform choices among candidates.  Returns
    ``(grid, r, c, hold)`` of the imagined final state (the caller renders it, ``ray.py:299``).

    Candidate counts are >= 1 on every state ``sample_state`` can produce; on injected states lacking a
    required object the reference crashes (``randint(0)``) -- here a skill with no candidate is skipped.
    """
    g = np.array(grid, dtype=np.uint8, copy=True)

    def has(bit):
        return (desired >> bit) & 1

    if has(T_MAKE_BREAD) and (g == WHEAT).any():                                # ray.py:226-231 (first wheat)
        g[_nth(g == WHEAT, 0)] = BREAD
    if has(T_EAT_BREAD) and (g == BREAD).any():                                 # ray.py:232-237
        k = rng.uniform(int((g == BREAD).sum()))
        g[_nth(g == BREAD, k)] = EMPTY
    if has(T_CHOP_TREE) and (g == TREE).any():                                  # ray.py:238-243 (first tree)
        g[_nth(g == TREE, 0)] = STICKS
    if has(T_MOVE_STICKS) and (g == STICKS).any():                              # ray.py:244-257
        k = rng.uniform(int((g == STICKS).sum()))
        free = g == EMPTY                                                       # [:,:,:9]: objects AND agent
        free[r, c] = False
        if free.any():
            spot = rng.uniform(int(free.sum()))
            src, dst = _nth(g == STICKS, k), _nth(free, spot)
            g[src] = EMPTY
            g[dst] = STICKS
    if has(T_BUILD_HOUSE) and (g == STICKS).any():                              # ray.py:258-264
        k = rng.uniform(int((g == STICKS).sum()))
        g[_nth(g == STICKS, k)] = HOUSE
    if has(T_CHOP_ROCK) and (g == ROCK).any():                                  # ray.py:265-268 (first rock)
        g[_nth(g == ROCK, 0)] = EMPTY
    if has(T_GO_TO_HOUSE) and (g == HOUSE).any():                               # ray.py:269-276
        k = rng.uniform(int((g == HOUSE).sum()))
        r, c = _nth(g == HOUSE, k)                                              # agent + held bits move
    if has(T_MOVE_AXE) and (g == AXE).any():                                    # ray.py:277-286 (first axe)
        free = g == EMPTY                                                       # [:,:,:8]: agent cell allowed
        if free.any():
            spot = rng.uniform(int(free.sum()))
            src, dst = _nth(g == AXE, 0), _nth(free, spot)
            g[src] = EMPTY
            g[dst] = AXE
    if has(T_MOVE_HAMMER) and (g == HAMMER).any():                              # ray.py:287-297 (first hammer)
        free = g == EMPTY
        if free.any():
            spot = rng.uniform(int(free.sum()))
            src, dst = _nth(g == HAMMER, 0), _nth(free, spot)
            g[src] = EMPTY
            g[dst] = HAMMER
    return g, r, c, hold


def reset_env(seed: int, env_id: int, episode: int, cfg: Config, with_goal: bool = False):
    """``reset()`` (``ray.py:156-218``) on the Philox stream.  Draw order: task count, task subset,
    placement, then (only if ``with_goal``) the imagine_obs draws.  Returns ``EnvState`` (and the imagined
    goal state tuple if ``with_goal``)."""
    rng = PhiloxStream(seed, env_id, episode)
    desired = sample_tasks(rng, cfg)                                            # ray.py:169-174
    grid, r, c = sample_state(rng, cfg)                                         # ray.py:178-179
    s = EnvState(grid, grid.copy(), r, c, 0, 0, desired, 0, episode)            # ray.py:176, 183, 203
    if with_goal:
        return s, imagine(grid, r, c, 0, desired, rng)                          # ray.py:191
    return s
