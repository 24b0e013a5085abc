#!/usr/bin/env python
"""bench.py -- env-steps/s of the batched CraftingWorld hot path on N B200s (one process per GPU).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3|cfg4|cfg5] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of worlds: ONE fused step + auto-reset + render launch
(cw_env_kernel) per step for the pixel workloads, one cw_step_chained_kernel launch for the compact workload.  Worlds are
sharded over ranks by global id with no data-path collective (weak scaling: per-GPU batch fixed); with N > 1 the
24 x int64 episode-statistics vector is snapshotted on the step stream and all-reduced over NCCL on a side stream at
a graph-replay boundary every 128 env steps (BASELINE config 4's cadence, counted across the timed windows): the
reduction of the statistics up to step k overlaps steps k+1.., and the window it falls into only closes once it has finished.

Prints ONE JSON line (rank 0):
  value         device-resident throughput of `--workload` (default cfg2 = BASELINE.json configs[1]): exactly K steps
                replayed from CUDA graphs, actions already in HBM, frames left in HBM, CUDA events on the launching
                stream behind a device-side gate (host launch latency is outside the window), median of several
                windows, max over ranks.
  workloads     the same measurement, shorter, for the other BASELINE configs (cfg3, cfg4, cfg5), each with its own
                roofline record.
  closed_loop   actions of step k+1 computed ON THE DEVICE from the pixels of frame k (cw_frame_policy reads every byte).
  e2e           the same metric through the host-buffer C entry points (HostCraftingWorldEnv -> cw_host_step) with HOST
                arrays: `e2e.value` = frames produced in HBM for a device-side consumer, actions from / reward + done
                back to host memory every step; `e2e.host_frames_delta` additionally keeps the frames current in HOST
                memory; `e2e.full_frame_copy` copies every frame over PCIe.  Wall clock, MAX over ranks, median of 3 windows
                of >= 1000 calls after 200 untimed calls.
  roofline      the fused kernel's algorithmic bytes per launch / its mean launch duration vs the measured HBM bandwidth.
  cpu_baseline  the reference's own CraftingWorldEnvRay (oracle/_ref, installed unmodified by oracle/build_ref.py) on the
                host cores of this box, the Python port and the C port beside it.
`--impl reference` times only the reference's CPU env loop (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "env-steps/sec w/ pixel obs"
UNIT = "env-steps/s"
TAPE = 128                      # length of the pre-generated action tape == steps per CUDA graph == stats period
WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on at N=1
    "cfg2": dict(envs=4096, size=21, obs="pixels", dense=False,
                 desc="4096 envs default 21x21 grid per GPU, pixel obs, nine-skill random tasks, auto-reset"),
    "cfg3": dict(envs=65536, size=21, obs="compact", dense=False,
                 desc="65536 envs default grid per GPU, compact-state obs (no render), step kernel only"),
    "cfg4": dict(envs=131072, size=21, obs="pixels", dense=False,
                 desc="131072 envs default grid per GPU (1M over 8), pixel obs, stats all-reduce every 128 steps"),
    "cfg5": dict(envs=16384, size=32, obs="pixels", dense=True,
                 desc="16384 envs 32x32 grid per GPU, dense object placement (p=0.5 per cell, kept dense: no auto-reset), pixel obs"),
}


def algorithmic_bytes(size, obs):
    """bytes one env-step must move.  pixels: frame write + grid read + 32 B scalars; compact: 32 B."""
    return 48 * size * size + size * size + 32 if obs == "pixels" else 32


# ---------------------------------------------------------------------------------------------------------
# CPU baselines: the reference's own env class (oracle/_ref), its Python port, the C port -- one process / thread per core
# ---------------------------------------------------------------------------------------------------------
def reference_installed():
    """The unmodified reference as installed by oracle/build_ref.py (bench.py never reads /root/reference)."""
    ref = os.path.join(ROOT, "oracle", "_ref")
    if os.path.isfile(os.path.join(ref, "gym_craftingworld", "envs", "craftingworld_ray.py")):
        os.environ["CW_REFERENCE"] = ref                          # first candidate of oracle/ref_shim.py
        return True
    return False


def cpu_loop_throughput(size, envs_per_proc, steps, warmup, procs=None, kind="port"):
    import multiprocessing as mp
    from oracle import pyenv
    cores = sorted(os.sched_getaffinity(0))
    procs = procs or len(cores)
    ctx = mp.get_context("fork")
    args = [(envs_per_proc, steps, warmup, (size, size), 300, 1000 + i) for i in range(procs)]
    with ctx.Pool(procs) as pool:
        res = pool.map(pyenv.run_worker_reference if kind == "reference" else pyenv.run_worker, args)
    total = sum(n for n, _ in res)
    slowest = max(dt for _, dt in res)
    return total / slowest, procs, total


def cpu_c_port_throughput(size, envs, steps, render_mode, threads):
    """The C restatement (oracle/cw_oracle.c) on `threads` pthreads: a much stronger CPU baseline than the reference's
    own Python loop; reported beside it for context."""
    import numpy as np
    from oracle import native
    ob = native.OracleBatch(native.make_config(H=size, W=size), envs, seed=1)
    ob.reset()
    obs = ob.render() if render_mode else None
    acts = np.random.RandomState(0).randint(0, 6, (steps, envs)).astype(np.uint8)
    ob.run_threads(acts[:2], render_mode=render_mode, obs=obs, nthreads=threads)
    t0 = time.perf_counter()
    ob.run_threads(acts, render_mode=render_mode, obs=obs, nthreads=threads)
    return envs * steps / (time.perf_counter() - t0)


def cpu_baseline_block(size, quick=False):
    cores = len(os.sched_getaffinity(0))
    out = {}
    if reference_installed():
        steps = 400 if quick else 1500
        v, procs, total = cpu_loop_throughput(size, 16, steps, 50, kind="reference")
        out = {"value": v, "unit": UNIT, "cores": procs, "kind": "reference",
               "sample": f"{procs} processes x 16 worlds x {steps} steps ({total} env-steps) of the UNMODIFIED reference "
                         f"CraftingWorldEnvRay(size=({size},{size})) env loop (step, reset on done; nine-skill random tasks), "
                         "imported from oracle/_ref under the gym/matplotlib import shim"}
    steps = 1500 if quick else 6000
    v, procs, total = cpu_loop_throughput(size, 16, steps, 100, kind="port")
    port = {"value": v, "unit": UNIT, "cores": procs, "kind": "port",
            "sample": f"{procs} processes x 16 worlds x {steps} steps ({total} env-steps) of the {size}x{size} nine-skill "
                      f"env loop, Python/NumPy port of the reference (oracle/pyenv.py: incremental render_edit, reset on done)"}
    if out:
        out["port"] = port
    else:
        out = port
    try:
        k = 40 if quick else 150
        out["c_port"] = {
            "incremental_render": cpu_c_port_throughput(size, 4096, k * 4, 2, cores),
            "full_render_every_step": cpu_c_port_throughput(size, 4096, k, 1, cores),
            "unit": UNIT, "cores": cores,
            "note": "oracle/cw_oracle.c on pthreads (not the reference's implementation language); context only"}
    except Exception as e:  # noqa: BLE001
        out["c_port"] = {"error": repr(e)}
    return out


# ---------------------------------------------------------------------------------------------------------
# clocks: NVML polled from a thread (~1 ms period) so that even a sub-millisecond timed region has samples inside
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, local_rank):
        self.samples, self.stop_flag, self.thread, self.sm_max, self.src = [], False, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = self._handle(local_rank)
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.src = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:  # noqa: BLE001
            self.nv = None
        deadline = time.perf_counter() + 2.0                      # do not start measuring before the first sample exists
        while self.nv is not None and not self.samples and time.perf_counter() < deadline:
            time.sleep(0.001)

    def _handle(self, local_rank):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = local_rank
        if vis:
            tok = vis.split(",")[local_rank].strip()
            if tok.startswith("GPU-"):
                return self.nv.nvmlDeviceGetHandleByUUID(tok)
            idx = int(tok)
        return self.nv.nvmlDeviceGetHandleByIndex(idx)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:  # noqa: BLE001
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                self.samples.append((time.perf_counter(), float(sm), int(rs)))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.0005)

    def stop(self):
        self.stop_flag = True

    def summary(self, t0, t1, t_load0):
        rows = [s for s in self.samples if t0 <= s[0] <= t1]
        window = "timed region"
        if len(rows) < 3:
            rows = [s for s in self.samples if t_load0 <= s[0] <= t1 + 0.05]
            window = "warm-up + timed region (timed region shorter than 3 samples)"
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": [], "samples": 0, "window": window, "source": self.src}
        sm = sorted(r[1] for r in rows)
        bits = 0
        for r in rows:
            bits |= r[2]
        reasons = sorted(nm for nm, b in self.REASONS.items() if bits & b)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.sm_max, "reasons": reasons, "samples": len(sm),
                "samples_in_timed_region": len([s for s in self.samples if t0 <= s[0] <= t1]), "window": window, "source": self.src}


# ---------------------------------------------------------------------------------------------------------
def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth, burst)"
    except Exception:  # noqa: BLE001
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


def ncu_traffic(workload):
    """dram read+write bytes per launch of the dominant kernel from the committed ncu --set full capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f).get(workload)
    except Exception:  # noqa: BLE001
        return None


def dense_worlds(env, torch, seed):
    """BASELINE config 5 placement: each non-agent cell occupied w.p. 0.5, type uniform over the 8 objects, >= 1 of
    each type, agent on an empty cell, nothing held."""
    N, H, W = env.num_envs, env.cfg.H, env.cfg.W
    g = torch.Generator(device=env.device).manual_seed(seed)
    occ = torch.rand((N, H * W), generator=g, device=env.device) < 0.5
    typ = torch.randint(1, 9, (N, H * W), generator=g, device=env.device, dtype=torch.uint8)
    grid = torch.where(occ, typ, torch.zeros_like(typ))
    cells = torch.rand((N, H * W), generator=g, device=env.device).argsort(dim=1)[:, :9]
    for k in range(8):
        grid.scatter_(1, cells[:, k:k + 1], torch.full((N, 1), k + 1, dtype=torch.uint8, device=env.device))
    ac = cells[:, 8]
    grid.scatter_(1, ac.unsqueeze(1), torch.zeros((N, 1), dtype=torch.uint8, device=env.device))
    desired = torch.randint(1, 512, (N,), generator=g, device=env.device)
    env.load_state(grid.cpu().numpy().reshape(N, H, W), (ac // W).cpu().numpy(), (ac % W).cpu().numpy(),
                   torch.zeros(N).numpy(), desired.cpu().numpy())


class Ctx:
    """Per-process measurement context: device, ranks, stream, barrier, timing of exactly K steps."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.stream = torch.cuda.Stream(device=self.dev)
        import gym_craftingworld_b200 as cw
        fair = max(1, len(os.sched_getaffinity(0)) // int(os.environ.get("LOCAL_WORLD_SIZE", "1")))
        bound = cw.bind_to_gpu_numa_node(self.local_rank)         # host-side polling / patching next to the GPU's PCIe root
        if bound:                                                 # the worker pool keeps its fair share of ALL cores (not of one node's)
            os.environ.setdefault("CW_HOST_THREADS", str(min(16, fair)))
        self.numa = f"bound to {len(bound)} CPUs of the GPU's NUMA node" if bound else "not changed (single node or unknown topology)"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, values):
        """element-wise MAX over ranks of a list of floats; also returns every rank's list (rank-major)"""
        torch = self.torch
        t = torch.tensor(values, device=self.dev, dtype=torch.float64)
        if self.world == 1:
            return list(values), [list(values)]
        allr = [torch.zeros_like(t) for _ in range(self.world)]
        self.dist.all_gather(allr, t)
        stack = torch.stack(allr)
        return stack.max(dim=0).values.tolist(), stack.tolist()

    def time_steps(self, enqueue, K, windows=None):
        """Device time of exactly K steps, `windows` times.  `enqueue()` puts the K steps on self.stream.  Each window is
        bracketed by barrier + synchronize; a ~0.3 ms device-side gate in front of the first event keeps the host's launch
        latency out of the CUDA-event window (everything is queued before the GPU reaches the start event)."""
        torch = self.torch
        with torch.cuda.stream(self.stream):
            enqueue()                                             # untimed: also tells how long a window is
            torch.cuda.synchronize()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream); enqueue(); e1.record(self.stream)
            torch.cuda.synchronize()
            est = e0.elapsed_time(e1)
            if windows is None:
                windows = 5 if est < 400 else (3 if est < 3000 else 1)
            ms = []
            w0 = time.perf_counter()
            for _ in range(windows):
                self.barrier()
                ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda._sleep(int(0.3e-3 * 1.9e9))           # the gate
                ev0.record(self.stream)
                enqueue()
                ev1.record(self.stream)
                torch.cuda.synchronize()
                ms.append(ev0.elapsed_time(ev1))
            w1 = time.perf_counter()
            self.barrier()
        mx, per_rank = self.max_over_ranks(ms)                    # per window: MAX over ranks
        order = sorted(range(len(mx)), key=lambda i: mx[i])
        mid = order[len(order) // 2]
        ranks_mid = sorted(r[mid] for r in per_rank)
        return {"ms": mx[mid], "windows_ms": mx, "rank_ms": {"min": ranks_mid[0], "median": ranks_mid[len(ranks_mid) // 2], "max": ranks_mid[-1]},
                "t0": w0, "t1": w1}


def frame_ring(args, name, main=True):
    """number of rotating frame buffers of a pixel workload: the ring must exceed the 126 MB L2 and chained launches need >= 2"""
    wl = WORKLOADS[name]
    N = (args.envs or wl["envs"]) if main else wl["envs"]
    ring = 1
    while N * 48 * wl["size"] * wl["size"] * ring < 300e6 and ring < 64:
        ring *= 2
    ring = max(ring, 2)
    if args.ring and main:
        ring = args.ring
    return ring


def workload_config(args, world):
    """`config` of the JSON line: what is measured, as a pure function of the command line -- BOTH arms print exactly this dict
    (how each arm measures it is under `method` / `cpu_baseline.sample`)."""
    wl = WORKLOADS[args.workload]
    N, size = args.envs or wl["envs"], wl["size"]
    cfg = {"workload": f"{args.workload}: {wl['desc']}", "envs_per_gpu": N, "grid": f"{size}x{size}", "obs": wl["obs"],
           "actions": "uniform iid over the 6 actions, a pre-generated 128-step tape, cycled",
           "parallelism": f"dp{world} (worlds sharded by global id, no data-path collective)"}
    if wl["obs"] == "pixels":
        ring = frame_ring(args, args.workload)
        cfg["l2"] = (f"GPU arm: frames written round-robin into {ring} buffers = {ring * N * 48 * size * size / 1e6:.0f} MB > 126 MB L2 (inputs larger "
                     "than L2; no flush needed); CPU reference arm: not applicable")
    else:
        cfg["l2"] = "GPU arm: state of 65536 worlds x ~0.9 KB, the step kernel is latency bound; CPU reference arm: not applicable"
    return cfg


def stagger(env, torch, seed):
    """Spread the episode clocks uniformly over [0, max_steps): every step then sees the steady-state share of time-outs and
    re-seeds (N / max_steps worlds) instead of a synchronised storm every max_steps steps."""
    g = torch.Generator(device=env.device).manual_seed(seed)
    env.t.copy_(torch.randint(0, env.MAX_STEPS, (env.num_envs,), generator=g, device=env.device, dtype=torch.int32))


def pixel_or_compact_leg(cx, name, K, W_, main):
    """One BASELINE workload on this rank's GPU: returns the record (value, roofline, ...) -- the full set of legs when it is
    the main workload, the chained + independent-launch legs otherwise."""
    import gym_craftingworld_b200 as cw
    torch, args = cx.torch, cx.args
    wl = dict(WORKLOADS[name])
    if args.envs and main:
        wl["envs"] = args.envs
    N, size, pixels = wl["envs"], wl["size"], wl["obs"] == "pixels"
    frame_bytes = 48 * size * size
    ring = frame_ring(args, name, main) if pixels else 1          # rotate frame buffers so the ring exceeds the 126 MB L2
    auto_reset = not wl["dense"]                                  # cfg5 stays dense: a re-seed would replace a dense world by a 9-object one
    env = cw.BatchedCraftingWorldEnv(N, size=(size, size), seed=args.seed, device=cx.dev, auto_reset=auto_reset, obs_mode=wl["obs"],
                                     env_id_base=cx.rank * N, obs_buffers=ring, goal_images=not args.no_goal_images,
                                     max_steps=args.max_steps, collect_stats=not args.no_stats)
    env.reset()
    if wl["dense"]:
        dense_worlds(env, torch, 99 + cx.rank)
    else:
        stagger(env, torch, 7 + cx.rank)
    gen = torch.Generator(device=cx.dev).manual_seed(1234 + cx.rank)
    tape = torch.randint(0, 6, (TAPE, N), generator=gen, device=cx.dev, dtype=torch.uint8)
    reducer = cw.StatsReducer(env, every=TAPE, inline=os.environ.get("CW_STATS_INLINE", "0") == "1") if cx.world > 1 else None
    stream = cx.stream
    peak, peak_src = measured_peak()
    B = algorithmic_bytes(size, wl["obs"])
    rec = {"workload": f"{name}: {wl['desc']}", "envs_per_gpu": N, "grid": f"{size}x{size}", "obs": wl["obs"], "steps": K}

    with torch.cuda.stream(stream):
        for k in range(W_):                                       # eager warm-up (also warms the launch path)
            env.step(tape[k % TAPE])
        torch.cuda.synchronize()
        chain = not args.no_chain                                 # (compact observations: cw_step_chained, launches linked per warp)
        if chain:
            env.step(tape[0], chain_pos=0)                        # warm the chained launch path outside the capture

        def capture(nsteps, body):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                for k in range(nsteps):
                    body(k)
            return g

        def graph_runner(body):
            """exactly K steps as replays of graphs of <= TAPE steps; the statistics reduction fires at EVERY replay boundary
            (snapshot on this stream, NCCL on a side stream) and the window only closes once it has finished"""
            n_full, rem = divmod(K, TAPE)
            g_full = capture(TAPE, body) if n_full else None
            g_rem = capture(rem, body) if rem else None

            def enqueue():
                # BASELINE config 4: all-reduce the statistics every 128 env steps.  At a replay boundary at which 128 more steps
                # have accumulated (every replay of a full graph; every ~6th window when K = 20) the statistics are snapshotted
                # on this stream, the copy goes to NCCL on the side stream and stepping carries on -- the reduction overlaps the
                # next replay, and the timed window only closes once it has finished.
                for _ in range(n_full):
                    g_full.replay()
                    if reducer is not None:
                        reducer.step(TAPE)
                if g_rem is not None:
                    g_rem.replay()
                    if reducer is not None:
                        reducer.step(rem)
                if reducer is not None:
                    reducer.wait()                                # `stream` waits for the side stream: the all-reduce is inside the window
            return enqueue

    res = cx.time_steps(graph_runner(lambda k: env.step(tape[k], chain_pos=k if chain else None)), K)
    ms = res["ms"]
    rec.update(value=N * cx.world * K / (ms / 1e3), unit=UNIT, ms_per_step=ms / K, windows_ms=res["windows_ms"], rank_ms=res["rank_ms"],
               gpu_launches=K, t0=res["t0"], t1=res["t1"],
               launch=(f"CUDA graphs of <= {TAPE} steps, one launch per step"
                       + (("; launches chained by per-group dataflow (cw_step_render_chained: open-loop action tape, step i+1 overlaps "
                           "the draining frame stores of step i; results identical)" if pixels else
                           "; launches linked per warp of 32 worlds by dataflow (cw_step_chained: open-loop action tape, no launch waits "
                           "for its predecessor grid; results identical)") if chain else "")),
               l2=(f"frames written round-robin into {ring} buffers = {ring * N * frame_bytes / 1e6:.0f} MB > 126 MB L2 (inputs larger than "
                   "L2; no flush needed)") if pixels else "state 65536 x ~0.9 KB; step kernel is latency bound",
               episodes=("kept dense: auto-reset off, worlds step past done as upstream allows" if wl["dense"] else
                         "episode clocks staggered uniformly over [0, max_steps): steady-state share of time-outs / re-seeds in every step"),
               stats_allreduce=({"every_env_steps": TAPE, "issued_so_far": reducer.reductions,
                                 "note": "NCCL SUM all-reduce of the 16x24 int64 statistics snapshot, on a side stream, every 128 env steps "
                                         "(BASELINE config 4's cadence) counted across the windows; each is issued at a graph-replay boundary "
                                         "and joined before the window's stop event"} if reducer is not None else None))
    launch_s = ms / 1e3 / K
    achieved = B * N / launch_s / 1e9
    if pixels:
        rec["roofline"] = {"bound": "hbm", "kernel": "cw_env_kernel<V_CHAINED>" if chain else "cw_env_kernel<V_PLAIN>", "achieved": achieved,
                           "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(name),
                           "algorithmic_bytes_per_env_step": B, "units_per_launch": N, "launch_us": launch_s * 1e6, "peak_source": peak_src}
    else:
        rec["roofline"] = {"bound": "latency/issue (reported, not an HBM roofline: 32 algorithmic bytes per env-step)",
                           "kernel": "cw_step_chained_kernel" if chain else "cw_step_kernel<false>",
                           "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": ncu_traffic(name),
                           "algorithmic_bytes_per_env_step": B, "units_per_launch": N, "launch_us": launch_s * 1e6, "peak_source": peak_src}

    if main and pixels and K < 1280 and not args.no_steady:
        # K steps from an idle GPU to a drained one pay one pipeline fill + drain (~1 step of 20); the same launches over a long
        # window show the steady state the kernel sustains.  Reported beside `value`, never instead of it.
        Ks = 2560
        with torch.cuda.stream(stream):
            g_ss = capture(TAPE, lambda k: env.step(tape[k], chain_pos=k if chain else None))

        def run_ss():
            for _ in range(Ks // TAPE):
                g_ss.replay()
                if reducer is not None:
                    reducer.step(TAPE)
            if reducer is not None:
                reducer.wait()
        rs = cx.time_steps(run_ss, Ks, windows=3)
        rec["steady_state"] = {"steps": Ks, "value": N * cx.world * Ks / (rs["ms"] / 1e3), "unit": UNIT, "ms_per_step": rs["ms"] / Ks,
                               "roofline_frac": B * N / (rs["ms"] / 1e3 / Ks) / 1e9 / peak, "windows_ms": rs["windows_ms"],
                               "note": f"the same chained launches, {Ks} steps per window (20 graph replays): the K-step window of `value` "
                                       "runs from an idle GPU to a drained one, this one amortises that fill + drain"}

    if chain and not args.no_unchained:                           # the same K steps as independent launches, for comparison
        with torch.cuda.stream(stream):
            run = graph_runner(lambda k: env.step(tape[k]))
        r2 = cx.time_steps(run, K)
        rec["unchained"] = {"value": N * cx.world * K / (r2["ms"] / 1e3), "unit": UNIT, "ms_per_step": r2["ms"] / K,
                            "roofline_frac": B * N / (r2["ms"] / 1e3 / K) / 1e9 / peak,
                            "note": "the same K steps as independent (whole-grid dependent, PDL) launches"}

    if pixels and not args.no_closed_loop:
        # closed loop with a REAL device consumer: between two steps a kernel reads every byte of every frame and derives the
        # next actions from the pixels (cw_frame_policy), so step k+1 depends on frame k.  Bytes per step: the frames are
        # written once and read once.
        abuf = torch.zeros(N, dtype=torch.uint8, device=cx.dev)

        def body(k):
            env.frame_policy(out=abuf)
            env.step(abuf)
        with torch.cuda.stream(stream):
            body(0)
            run = graph_runner(body)
        r3 = cx.time_steps(run, K)
        Bc = B + frame_bytes + 1
        ach = Bc * N / (r3["ms"] / 1e3 / K) / 1e9
        rec["closed_loop"] = {"value": N * cx.world * K / (r3["ms"] / 1e3), "unit": UNIT, "ms_per_step": r3["ms"] / K, "gpu_launches": 2 * K,
                              "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                           "algorithmic_bytes_per_env_step": Bc},
                              "note": "per step: cw_frame_policy (a device consumer that reads every frame byte and computes the next "
                                      "action from the pixels) + one independent fused launch; the frames are written AND read"}
    rec["episode_stats_rank0"] = {k: v for k, v in env.episode_stats().items() if k in ("episodes", "successes", "mean_return", "mean_length")}

    # ---- compact workload: the same open-loop tape as ONE launch per 128 steps (cw_rollout) -------------------------
    if not pixels:
        reps = max(1, K // TAPE)

        def run():
            for _ in range(reps):
                env.rollout(tape, return_trace=False)
        r4 = cx.time_steps(run, reps * TAPE)
        rec["rollout"] = {"value": N * cx.world * reps * TAPE / (r4["ms"] / 1e3), "unit": UNIT, "steps": reps * TAPE, "gpu_launches": reps,
                          "ms_per_step": r4["ms"] / (reps * TAPE),
                          "note": f"cw_rollout: {TAPE} steps of the same tape per launch, state in registers across the steps (open loop only)"}

    # ---- the same workload with INCREMENTAL rendering (the reference's render_edit, cw_step_render_edit) --------------
    if pixels and main and not args.no_incremental and auto_reset:
        del env
        torch.cuda.empty_cache()
        ienv = cw.BatchedCraftingWorldEnv(N, size=(size, size), seed=args.seed, device=cx.dev, auto_reset=True, env_id_base=cx.rank * N,
                                          goal_images=not args.no_goal_images, max_steps=args.max_steps,
                                          collect_stats=not args.no_stats, render="incremental")
        ienv.reset()
        stagger(ienv, torch, 7 + cx.rank)
        reducer = None
        with torch.cuda.stream(stream):
            for k in range(3):
                ienv.step(tape[k])
            n_full, rem = divmod(K, TAPE)
            gi = capture(TAPE if n_full else rem, lambda k: ienv.step(tape[k]))
        per = TAPE if n_full else rem
        reps = max(1, K // per)
        r5 = cx.time_steps(lambda: [gi.replay() for _ in range(reps)], reps * per)
        rec["incremental_render"] = {"value": N * cx.world * reps * per / (r5["ms"] / 1e3), "unit": UNIT, "steps": reps * per,
                                     "ms_per_step": r5["ms"] / (reps * per), "gpu_launches": 2 * reps * per,
                                     "note": "render='incremental' (cw_step_render_edit): the persistent frame buffer in HBM is kept current by "
                                             "rewriting only the <= 2 cells a step changes (what the reference's step does, ray.py:358, 522-557) "
                                             "plus full frames for re-seeded worlds; identical pixels, ~100x fewer bytes, so NOT comparable with "
                                             "the roofline of the full-expansion path that `value` measures"}
        del ienv
    torch.cuda.empty_cache()
    return rec, tape


def e2e_legs(cx, name, K, tape):
    """The metric through the host-buffer C entry points (HostCraftingWorldEnv -> cw_host_step), host arrays in and out, wall clock,
    MAX over ranks."""
    import numpy as np
    import gym_craftingworld_b200 as cw
    torch, args = cx.torch, cx.args
    wl = WORKLOADS[name]
    N, size = (args.envs or wl["envs"]), wl["size"]
    frame_bytes = 48 * size * size
    e_steps = max(K, 1000 if not args.quick else 30)
    e_steps = min(e_steps, 4000)
    if N * frame_bytes > 1.5e9:
        e_steps = min(e_steps, 200)
    acts = tape.cpu().numpy()
    res = {}

    def timed(fn, n_calls, steps_per_call, windows=3):
        """median of `windows` wall-clock windows (each the MAX over ranks) of n_calls calls, like `value`'s device windows"""
        rates = []
        for _ in range(windows):
            cx.barrier()
            h0 = time.perf_counter()
            for k in range(n_calls):
                fn(k)
            henv.sync()                                           # device-consumer legs: the frames of the last step are complete
            dt = time.perf_counter() - h0
            mx, _ = cx.max_over_ranks([dt])
            rates.append(N * cx.world * n_calls * steps_per_call / mx[0])
        return sorted(rates)[len(rates) // 2]

    for variant in ("device", "delta", "frames"):
        henv = cw.HostCraftingWorldEnv(N, size=(size, size), seed=args.seed, device=cx.local_rank, env_id_base=cx.rank * N,
                                       return_frames=variant != "device", transport="delta" if variant == "delta" else "frames",
                                       max_steps=args.max_steps)
        henv.reset()
        henv.load_state(t=np.random.RandomState(11 + cx.rank).randint(0, args.max_steps, N))   # staggered episode clocks (as above)
        n_steps = e_steps if variant != "frames" else max(10, e_steps // 50)
        for k in range(200 if variant != "frames" else 3):        # untimed: page-faults of the mirrors, worker threads hot, host clocks up
            henv.step(acts[k % TAPE])
        v = timed(lambda k: henv.step(acts[k % TAPE]), n_steps, 1, windows=3 if variant != "frames" else 1)
        res[variant] = {"value": v, "unit": UNIT, "h2d_bytes_per_step": henv.h2d_bytes_per_step, "d2h_bytes_per_step": henv.d2h_bytes_per_step,
                        "steps": n_steps, "windows": 3 if variant != "frames" else 1}
        if variant == "device":                                   # the same transport, 128 steps of an open-loop tape per library call
            calls = max(1, min(e_steps // TAPE, 16))
            henv.step_many(acts)
            res["device_many"] = {"value": timed(lambda k: henv.step_many(acts), calls, TAPE, windows=1), "unit": UNIT, "steps": calls * TAPE,
                                  "h2d_bytes_per_step": N, "d2h_bytes_per_step": N}
        henv.close()
        cx.barrier()
    e2e = dict(res["device"])
    e2e["api"] = ("HostCraftingWorldEnv(return_frames=False).step -> cw_host_step(obs_host=NULL): every step the actions come from pinned HOST "
                  "memory and reward + done are back in HOST memory when the call returns (one status byte per world through mapped pinned "
                  "memory, no stream synchronisation); the pixel frames are produced in HBM for a device-side consumer (four rotating frame "
                  "buffers) -- the call a GPU policy loop makes.  Batches of <= 16384 worlds run as a two-launch pipeline per step: a "
                  "thread-per-world step launch (cw_step_snap_kernel: status bytes, live state, a state snapshot) on one stream and the "
                  "render launch of that snapshot (cw_env_kernel<V_PIPE>) on another, linked by release/acquire words -- the step of "
                  "call k+1 never waits for the frames of call k; larger batches use one fused chained launch per step. "
                  "tests/test_gpu_parity.py::test_host_env_device_consumer_matches_oracle")
    e2e["step_many_128"] = dict(res["device_many"], note="cw_host_step_many: 128 steps of an open-loop tape per library call (K chained launches, "
                                "reward/done rows unpacked as they land)")
    e2e["host_frames_delta"] = dict(res["delta"], note="transport='delta': additionally the CURRENT pixel frames are in HOST memory after every call "
                                    "(16-byte pre-digested records + host-side render_edit by a worker pool; bit-identical to a device render + copy). "
                                    "Bound by the host cores' memory system (~50 ns per world-step per core, tools/patch_bench.cpp), not by the GPU")
    e2e["full_frame_copy"] = dict(res["frames"], note="transport='frames': every rendered frame copied over PCIe (sliced, two streams); PCIe-bound at ~52 GB/s")
    return e2e


def run_ours(args):
    K, W_ = args.steps, max(args.warmup, 3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cpu_base = None
    if world == 1 and not args.no_cpu_baseline:                  # before CUDA is initialised (fork-safe)
        cpu_base = cpu_baseline_block(WORKLOADS[args.workload]["size"], quick=args.quick)
    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0"))) if rank == 0 else None   # before any GPU work; waits for its first sample
    t_load0 = time.perf_counter()
    cx = Ctx(args)
    main_rec, tape = pixel_or_compact_leg(cx, args.workload, K, W_, True)
    pixels = WORKLOADS[args.workload]["obs"] == "pixels"
    others = {}
    if not args.only:
        k_sub = max(3, min(K, 256))
        for name in ("cfg3", "cfg4", "cfg5"):
            if name == args.workload:
                continue
            # (a compact step launch is ~5 us: 20 of them are a 0.1 ms window, a third of it ramp -- the compact sub-record
            #  times at least one full 128-step graph, still under a millisecond)
            rec, _ = pixel_or_compact_leg(cx, name, max(k_sub, TAPE) if WORKLOADS[name]["obs"] == "compact" else k_sub, 3, False)
            for key in ("t0", "t1"):
                rec.pop(key, None)
            others[name] = rec
    if sampler:
        sampler.stop()                                           # the nvidia poller must not compete with the host threads of the e2e legs
    e2e = e2e_legs(cx, args.workload, K, tape) if (pixels and not args.no_e2e) else None

    if cx.rank == 0:
        t0, t1 = main_rec.pop("t0"), main_rec.pop("t1")
        roof = main_rec.pop("roofline")
        roof["note"] = ("launch duration = CUDA-event time of the timed region / launches (includes inter-kernel gaps). peak is the measured COPY "
                        "bandwidth (read+write mix); this kernel is a ~98% write stream, which does not pay a copy's read/write turnarounds, so "
                        "frac can slightly exceed 1. traffic = dram__bytes_read+write per launch from ncu (profiles/traffic.json)")
        line = {
            "metric": METRIC if pixels else "env-steps/sec compact obs (step kernel only)", "value": main_rec["value"], "unit": UNIT,
            "n_gpus": cx.world, "steps": K, "warmup": W_, "ms_per_step": main_rec["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, cx.world),
            "method": {"launch": main_rec["launch"], "l2": main_rec["l2"], "episodes": main_rec["episodes"],
                       "actions": "the tape is a uint8[128,N] tensor on the device",
                       "timing": "exactly K steps between two CUDA events on the launching stream, behind a ~0.3 ms device-side gate so that every "
                                 "launch is queued before the start event; median of len(windows_ms) such windows, each MAX over ranks",
                       "host_affinity_rank0": cx.numa},
            "windows_ms": main_rec["windows_ms"], "rank_ms": main_rec["rank_ms"], "stats_allreduce": main_rec["stats_allreduce"],
            "clocks": sampler.summary(t0, t1, t_load0) if sampler else None,
            "gpu_launches": main_rec["gpu_launches"],
            "steady_state": main_rec.get("steady_state"),
            "unchained": main_rec.get("unchained"), "closed_loop": main_rec.get("closed_loop"),
            "incremental_render": main_rec.get("incremental_render"), "rollout": main_rec.get("rollout"),
            "workloads": others or None,
            "e2e": e2e,
            "roofline": roof,
            "cpu_baseline": cpu_base,
            "episode_stats_rank0": main_rec["episode_stats_rank0"],
        }
        emit(line)
    if cx.world > 1:
        cx.dist.destroy_process_group()


def run_reference(args):
    """The reference arm: the reference's OWN CraftingWorldEnvRay (oracle/_ref, unmodified) on all host cores of the box, rank 0
    only; falls back to the Python port when the install is absent."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    wl = WORKLOADS[args.workload]
    K, W_ = args.steps, max(args.warmup, 3)
    cores = len(os.sched_getaffinity(0))
    kind = "reference" if reference_installed() else "port"
    # a "step" steps every world of the bounded sample once: `per_proc` worlds per process, sized from a short
    # calibration so that K steps take roughly 15 s of wall clock (never more than a few minutes)
    rate, _, _ = cpu_loop_throughput(wl["size"], 4, 100, 10, procs=cores, kind=kind)
    # (at most 128 worlds per process: a reference env object carries ~1 MB of int64 arrays, and with more of them per core the
    # loop falls out of cache -- the arm should show the reference at its best)
    per_proc = int(max(1, min(128, rate / cores * 15.0 / max(K + W_, 1))))
    value, procs, total = cpu_loop_throughput(wl["size"], per_proc, K, W_, procs=cores, kind=kind)
    what = ("the UNMODIFIED reference class CraftingWorldEnvRay (gym_craftingworld 0.1.9.8, installed into oracle/_ref by oracle/build_ref.py, "
            "imported under the gym/matplotlib shim)" if kind == "reference" else
            "Python/NumPy port of the reference (oracle/pyenv.py) -- oracle/_ref is absent on this box")
    sample = (f"{procs} processes x {per_proc} worlds, one step = every world stepped once ({procs * per_proc} env-steps), "
              f"{wl['size']}x{wl['size']} nine-skill env loop with reset on done; {what}")
    world = max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus)
    pixels = wl["obs"] == "pixels"
    line = {"impl": "reference", "metric": METRIC if pixels else "env-steps/sec compact obs (step kernel only)", "value": value, "unit": UNIT,
            "n_gpus": world, "steps": K, "warmup": W_, "ms_per_step": 1e3 * procs * per_proc / value, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": workload_config(args, world),
            "method": {"launch": "no GPU: one Python process per host core, each stepping its worlds one env.step() at a time (rank 0 only)",
                       "sample": "bounded CPU sample of the same env config (see cpu_baseline.sample)",
                       "timing": "wall clock around K steps of every world of the sample, after W warm-up steps"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line goes to the process's real stdout; everything else (NCCL banners, library chatter) was
    redirected to stderr by quiet_stdout()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def quiet_stdout() -> None:
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2560)
    ap.add_argument("--warmup", type=int, default=64)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--only", action="store_true", help="measure only --workload (skip the sub-records of the other BASELINE configs)")
    ap.add_argument("--envs", type=int, default=0, help="override worlds per GPU")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--quick", action="store_true", help="shorter CPU baseline / e2e legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--ring", type=int, default=0, help="override the number of rotating frame buffers")
    ap.add_argument("--no-chain", action="store_true", help="independent launches instead of chained ones")
    ap.add_argument("--no-unchained", action="store_true", help="skip the comparison leg with independent launches")
    ap.add_argument("--no-steady", action="store_true", help="skip the long-window (steady state) leg of the main workload")
    ap.add_argument("--no-closed-loop", action="store_true", help="skip the closed-loop leg (device consumer between the steps)")
    ap.add_argument("--no-incremental", action="store_true", help="skip the incremental-rendering leg")
    ap.add_argument("--no-stats", action="store_true", help="experiment: do not accumulate episode statistics")
    ap.add_argument("--no-goal-images", action="store_true", help="experiment: skip imagine_obs / goal + init frames")
    ap.add_argument("--max-steps", type=int, default=300, help="experiment: episode length (reference default 300)")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
