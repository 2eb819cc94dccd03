#!/usr/bin/env python3
"""bench.py -- env-steps/s of the batched UR5e reach simulator on N B200s, with roofline, end-to-end and CPU-baseline
figures on one JSON line (the contract is in the task statement; DESIGN.md section "Measurement" explains each key).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--task ID] [--envs-per-gpu M]
                    [--geometry capsule|hull]

Workload (BASELINE.json metric): UR5DynReach-v1, 1 Mi envs per GPU (config[4]: 8 Mi envs on 8 GPUs), i.i.d. U(-1,1)
actions resident in HBM, auto-reset on.  One step = one launch of the fused step kernel over all envs of the rank."""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

OBS_DIM = {"UR5OriReach-v1": 18, "UR5ObsReach-v1": 26, "UR5StaReach-v1": 29, "UR5DynReach-v1": 35}
# algorithmic bytes per env-step and per reset (SURVEY.md section 8d / BASELINE.md section 5)
BYTES_STEP = {"UR5OriReach-v1": 215, "UR5ObsReach-v1": 287, "UR5StaReach-v1": 323, "UR5DynReach-v1": 371}
BYTES_RESET = {"UR5OriReach-v1": 100, "UR5ObsReach-v1": 144, "UR5StaReach-v1": 168, "UR5DynReach-v1": 240}
METRIC = "env-steps/sec"
UNIT = "env-steps/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--task", default="UR5DynReach-v1", choices=sorted(OBS_DIM))
    ap.add_argument("--envs-per-gpu", type=int, default=1 << 20)
    ap.add_argument("--geometry", default="capsule", choices=["capsule", "hull"])
    ap.add_argument("--chains", type=int, default=2, help="independent env sub-ranges per GPU in the captured graph (1..8)")
    ap.add_argument("--graph-steps", type=int, default=32, help="env steps captured per CUDA graph replay (multiple of 8)")
    ap.add_argument("--e2e-steps", type=int, default=60)
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the GPU's NUMA node")
    ap.add_argument("--no-extra-configs", action="store_true", help="headline only: skip the other BASELINE configs / rollout loop")
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="wall-clock budget of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_config(args, world):
    return {"workload": f"{args.task}, {args.envs_per_gpu} envs per GPU, random actions U(-1,1) resident in HBM, "
                        f"auto-reset on, {args.geometry} geometry",
            "task": args.task, "envs_per_gpu": args.envs_per_gpu, "total_envs": args.envs_per_gpu * world,
            "geometry": args.geometry, "chains_per_gpu": args.chains if args.envs_per_gpu >= (1 << 19) else 1, "sharding": f"env index ranges over {world} rank(s), no data-path collective",
            "graph_steps": max(8, args.graph_steps // 8 * 8),
            "l2": "per-step traffic (state + actions + outputs) exceeds the 126 MB L2; an 8-deep ring of action buffers"}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def _cpu_worker(task, geom, seed, n_steps, conn):
    import numpy as np
    from oracle import oracle_env as oe
    env = oe.make(task, geom=geom, stream=oe.PhiloxStream(seed), env_index=seed, first_event=1)
    rng = np.random.default_rng(seed)
    acts = rng.uniform(-1, 1, (max(n_steps, 1), 6)).astype(np.float32)
    conn.send("ready")
    conn.recv()
    t0 = time.perf_counter()
    ev = 1
    for k in range(n_steps):                     # demo.py:10-15 without the GUI
        o, r, term, trunc, info = env.step(acts[k])
        ev += 1
        if term or trunc:
            env.reset(event=ev)
    conn.send(time.perf_counter() - t0)


def cpu_env_steps_per_s(task, geometry, steps_per_proc, procs):
    """P processes, one oracle env each (the reference allows one env per process: SURVEY.md Q10), random actions,
    reset on done.  Returns (env-steps/s over all processes, seconds)."""
    import multiprocessing as mp
    from oracle import oracle_env as oe
    oe.build_oracle()
    geom = oe.GEOM_CAPSULE if geometry == "capsule" else oe.GEOM_HULL
    ctx = mp.get_context("fork")
    pipes, ps = [], []
    for i in range(procs):
        a, b = ctx.Pipe()
        p = ctx.Process(target=_cpu_worker, args=(task, geom, i, steps_per_proc, b), daemon=True)
        p.start(); pipes.append(a); ps.append(p)
    for a in pipes:
        a.recv()
    t0 = time.perf_counter()
    for a in pipes:
        a.send("go")
    for a in pipes:
        a.recv()
    dt = time.perf_counter() - t0
    for p in ps:
        p.join()
    return procs * steps_per_proc / dt, dt


def calibrate_cpu(task, geometry, seconds, procs):
    rate1, _ = cpu_env_steps_per_s(task, geometry, 150, procs)
    per_proc = max(200, int(rate1 / procs * seconds))
    rate, dt = cpu_env_steps_per_s(task, geometry, per_proc, procs)
    return rate, per_proc, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    procs = os.cpu_count() or 1
    sample = 250                                  # env steps per process per bench "step"
    for _ in range(max(args.warmup, 0) and 1):
        cpu_env_steps_per_s(args.task, args.geometry, 50, procs)
    rate, dt = cpu_env_steps_per_s(args.task, args.geometry, sample * max(args.steps, 1), procs)
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(workload_config(args, world), env_steps_per_bench_step=procs * sample,
                           note="one bench step of this arm = `env_steps_per_bench_step` env steps (a bounded sample of the "
                                "workload on the host cores): value = env_steps_per_bench_step / (ms_per_step / 1000)"),
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
                             "sample": f"{procs} processes x {sample * max(args.steps, 1)} env steps of {args.task} (one oracle env per "
                                       "process, reset on done); the reference itself needs PyBullet, which is not installable "
                                       "here, so this is the CPU restatement (oracle/), not PyBullet"},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz, self.pause = index, [], set(), False, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            if self.pause:
                time.sleep(0.001)
                continue
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        import statistics
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ GPU arm
def bind_to_gpu_numa_node(local_rank):
    """Pin this process (and therefore the pages of the pinned host buffers it allocates afterwards) to the CPUs of the
    NUMA node its GPU hangs off.  Eight ranks that all sit on node 0 push every host copy of the far GPUs across the
    socket interconnect (round 1: e2e scaling efficiency 0.23 at N = 8)."""
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(local_rank)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = int(open(base + "/numa_node").read().strip())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        info.update({"pci": bdf, "numa_node": node, "local_cpulist": cpulist})
        if node >= 0 and use:
            os.sched_setaffinity(0, use)
            info.update({"bound": True, "cpus": len(use)})
    except Exception as e:                           # not fatal: the numbers are then measured unbound, and say so
        info["error"] = repr(e)
    return info


def pinned_copy_bandwidth(dev, mbytes=256, reps=5):
    """GB/s of one pinned-host <-> device cudaMemcpyAsync of `mbytes`, best of `reps`, each direction: the ceiling of the
    end-to-end path on this box, measured live (e2e.roofline.peak)."""
    import torch
    n = mbytes * (1 << 20)
    host, devb = torch.empty(n, dtype=torch.uint8, pin_memory=True), torch.empty(n, dtype=torch.uint8, device=dev)
    best = {}
    for name, (dst, src) in {"d2h": (host, devb), "h2d": (devb, host)}.items():
        t = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); dst.copy_(src, non_blocking=True); e1.record(); torch.cuda.synchronize(dev)
            t.append(e0.elapsed_time(e1))
        best[name] = n / (min(t) * 1e-3) / 1e9
    return best


def measure_device(ug, torch, dist, dev, local_rank, rank, world, task, n, geometry, steps, warm, chains, gs, offset,
                   link_dist="obstacle"):
    """One device-resident measurement: env-steps/s of `task` with n envs on this rank (inputs resident in HBM), max over
    ranks, plus the live per-kernel event timings.  Returns the dict that becomes a bench line or a `configs` entry."""
    import ctypes
    env = ug.UR5VecEnv(task, n, device=local_rank, seed=0, env_index_offset=offset, geometry=geometry, goal_buffers=True,
                       link_dist=link_dist)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    ring = [torch.rand((n, 6), device=dev, generator=g) * 2 - 1 for _ in range(8)]
    env.reset()
    for k in range(warm):
        env.step(ring[k % 8])
    chains = chains if n >= (1 << 19) else 1         # small batches: sub-ranges would not fill the 148 SMs
    gs = max(8, min(gs, max(steps, 8)) // 8 * 8)
    graph = env.capture_steps(ring * (gs // 8), chains=chains)
    graph.replay()
    env.stats(reset=True)
    n_replays, rem = divmod(steps, gs)               # exactly `steps` env steps: whole replays + a shorter graph
    graph_rem = env.capture_steps([ring[k % 8] for k in range(rem)], chains=chains) if rem else None
    if graph_rem is not None:
        graph_rem.replay()
        env.stats(reset=True)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    for k in range(n_replays):
        graph.replay()
    if graph_rem is not None:
        graph_rem.replay()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    st = env.stats(reset=True)                       # episodes finished inside the timed region only
    # per-kernel durations, right after the timed region (same clocks: the 0.25 s sampling repeat below would drop a
    # power-capped GPU to its sustained clock first): CUDA events recorded by the library on the launching stream directly around each kernel
    # (urgym_profile_enable), over back-to-back eager steps (one chain, serial kernels)
    ks = max(8, min(64, steps))
    sampler.pause = True                             # clocks are sampled over the timed region (and its repeat), not here
    env.L.urgym_profile_enable(env.h, 1)
    for k in range(ks):
        env.step(ring[k % 8])
    a_ms, b_ms, cnt = ctypes.c_double(), ctypes.c_double(), ctypes.c_int()
    env.L.urgym_profile_read(env.h, ctypes.byref(a_ms), ctypes.byref(b_ms), ctypes.byref(cnt))
    env.L.urgym_profile_enable(env.h, 0)
    env.stats(reset=True)
    torch.cuda.synchronize(dev)
    sampler.pause = False
    clock_note = "sampled during the timed region (NVML, 5 ms period)"
    if len(sampler.samples) < 8:
        t_end = time.perf_counter() + 0.25           # too short for the 5 ms sampler: an untimed repeat of the same loop
        while time.perf_counter() < t_end:
            graph.replay()
            torch.cuda.synchronize(dev)
        clock_note = "timed region shorter than the sampler period: sampled during an untimed 0.25 s repeat of the same loop"
    sampler.stop_flag = True
    barrier()
    env.stats(reset=True)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        st = ug.allreduce_stats(st, device=dev)
    ms_max = float(t.item())
    return dict(env=env, ring=ring, ms=ms, ms_max=ms_max, steps=steps, stats=st, chains=chains, graph_steps=gs,
                kernel_ms=a_ms.value, reset_kernel_ms=b_ms.value, clocks=dict(sampler.result(), how=clock_note),
                value=n * world * steps / (ms_max * 1e-3), launches=2 * steps * chains + (steps // gs + (1 if rem else 0)) * (chains > 1))


def load_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def roofline_of(m, task, n, world, geometry):
    """algorithmic bytes per launch / that kernel's average launch duration (capsule geometry, HBM-bound by
    construction); hull geometry is bound by the FP32 pipe and is reported against it."""
    peaks = load_peaks()
    steps, st = m["steps"], m["stats"]
    resets_per_launch = st["episodes"] / world / steps
    if geometry == "hull":
        # bound by the FP32 / issue pipes (GJK on convex hulls), not by HBM.  The bench measures the kernel time live;
        # the pipe utilisation comes from the committed ncu summary of the same kernel and is labelled as such.
        sm_max = peaks.get("sm_max_mhz", 1965.0)
        peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12              # FP32 FMA pipe, TFLOP/s at the maximum SM clock
        out = {"bound": "fp32", "achieved": None, "frac": None, "peak": peak, "unit": "TFLOP/s",
               "peak_source": f"148 SMs x 128 FP32 lanes x 2 x {sm_max:.0f} MHz (nominal)", "traffic": None,
               "kernel": f"urgym_step_kernel<{task}, hull>: {n} env-steps per launch", "kernel_ms": m["kernel_ms"],
               "reset_kernel_ms": m["reset_kernel_ms"]}
        try:
            out["ncu"] = json.load(open(os.path.join(ROOT, "profiles", "ncu_hull_kernel_summary.json")))
            # the FP32 FMA pipe's utilisation under ncu, scaled by (ncu duration / live duration): a committed figure, not
            # a live counter -- labelled as such
            out["frac"] = out["ncu"]["pipe_fma_pct"] / 100.0 * out["ncu"]["duration_ms"] / m["kernel_ms"]
            out["achieved"] = out["frac"] * peak
            out["frac_source"] = "pipe_fma_pct of " + out["ncu"]["source"].split(" ")[0] + " x (ncu duration / live kernel_ms)"
        except Exception:
            out["ncu"] = None
        return out
    peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json hbm_gbs)") if peaks.get("hbm_gbs") else (6650.0, "fallback (B200_PROFILING.md)")
    achieved = n * BYTES_STEP[task] / (m["kernel_ms"] * 1e-3) / 1e9
    whole = (n * BYTES_STEP[task] + resets_per_launch * BYTES_RESET[task]) / (m["ms"] * 1e-3 / steps) / 1e9
    traffic, traffic_src = None, None
    try:
        summ = json.load(open(os.path.join(ROOT, "profiles", "ncu_step_kernel_summary.json")))
        if n == summ.get("envs_per_launch", 1 << 20):
            traffic, traffic_src = summ["dram_bytes_per_launch"].get(task), summ.get("source")
    except Exception:
        pass
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "traffic_source": traffic_src, "peak_source": peak_src,
            "kernel": f"urgym_step_kernel<{task}, {geometry}>: {n} env-steps per launch x {BYTES_STEP[task]} B",
            "kernel_ms": m["kernel_ms"], "reset_kernel_ms": m["reset_kernel_ms"],
            "bytes_per_env_step": BYTES_STEP[task], "bytes_per_reset": BYTES_RESET[task], "resets_per_step": resets_per_launch,
            "whole_step": {"achieved": whole, "frac": whole / peak,
                           "how": "step + auto-reset kernels: (N*B_step + N_done*B_reset) / (timed region / steps)"}}


def measure_e2e(torch, dist, dev, world, env, ring, n, task, steps):
    """The same metric through the reference-facing host-buffer entry points: pinned host actions in, observation +
    reward + flags out, every step; two output slots in flight (urgym_step_host_async / urgym_host_wait)."""
    bufs = [env.alloc_host_buffers(terminal_obs=False) for _ in range(2)]
    host_ring = [r.cpu().pin_memory() for r in ring[:2]]
    for k in range(2):
        env.step_host(host_ring[k % 2], bufs[k % 2])
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    # three passes of `steps` steps, the median pass is reported (all three are in the line): one pass is ~0.2 s of host-side
    # work, and a single scheduling hiccup of the host process showed up as a 6x outlier in one of twenty runs
    passes = []
    for _ in range(3):
        t0 = time.perf_counter()
        env.step_host_async(0, host_ring[0], bufs[0])
        for k in range(1, steps):
            env.step_host_async(k % 2, host_ring[k % 2], bufs[k % 2])
            env.host_wait((k - 1) % 2)                   # (a consumer would read slot (k - 1) % 2 here)
        env.host_wait((steps - 1) % 2)
        torch.cuda.synchronize(dev)
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.barrier()
        passes.append(float(tt.item()))
    dt = sorted(passes)[1]
    D = OBS_DIM[task]
    h2d, d2h = n * 24, n * (4 * D + 4 + 3)
    bw = pinned_copy_bandwidth(dev)
    together = None
    if world > 1:
        # the ceiling that applies at N GPUs: every rank copying device->host at the same time (on the round-2 box eight
        # GPUs share ~93 GB/s of device->host bandwidth, 57 GB/s each when alone: profiles/pcie_bw_8gpu_r02.json)
        n_b = 256 << 20
        hb, db = torch.empty(n_b, dtype=torch.uint8, pin_memory=True), torch.empty(n_b, dtype=torch.uint8, device=dev)
        torch.cuda.synchronize(dev); dist.barrier()
        t1 = time.perf_counter()
        for _ in range(6):
            hb.copy_(db, non_blocking=True)
        torch.cuda.synchronize(dev)
        tw = torch.tensor([time.perf_counter() - t1], dtype=torch.float64, device=dev)
        dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        together = 6 * n_b / float(tw.item()) / 1e9
    ach = (h2d + d2h) * steps / dt / 1e9
    return {"value": n * world * steps / dt, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": steps,
            "passes_s": passes, "reported": "median of three passes of `steps` steps",
            "path": "urgym_step_host_async / urgym_host_wait, two slots in flight: pinned host actions -> device, step + "
                    "auto-reset kernels, observation + reward + 3 flag arrays -> pinned host",
            "roofline": {"bound": "pcie", "achieved": ach, "unit": "GB/s per GPU (both directions summed)",
                         "peak": together if together else bw["d2h"], "frac": ach / (together if together else bw["d2h"]),
                         "peak_alone_d2h": bw["d2h"], "peak_alone_h2d": bw["h2d"], "peak_all_ranks_at_once_d2h": together,
                         "peak_source": "256 MB pinned cudaMemcpyAsync device->host per GPU, measured in this run: `peak` is the rate "
                                        "with ALL ranks copying at once when N > 1 (the host side is shared), else this GPU alone"}}


def measure_rollout(ug, torch, dist, dev, world, m, task, n, steps):
    """BASELINE config C5's wording: the envs "driving a SAC rollout loop" (train.py:39-60).  The shipped policy of the
    task computes the actions from the observations on the device, the step follows, urgym_replay_write appends the
    transitions to a device-resident ring; captured as CUDA graphs of 8 steps.  SAC's gradient step is the caller's."""
    import numpy as np
    short = {"UR5OriReach-v1": "Ori", "UR5ObsReach-v1": "Obs", "UR5StaReach-v1": "Sta", "UR5DynReach-v1": "Dyn"}[task]
    w = np.load(os.path.join(ROOT, "tests", "golden", f"policy_{short}.npz"))
    policy = ug.mlp_policy({k: torch.as_tensor(w[k], device=dev) for k in w.files if not k.startswith("published")}, torch.bfloat16)
    env = m["env"]
    ring = ug.DeviceReplayRing(env, capacity=4 * n)
    noise = m["ring"]
    k = [0]

    def explore(obs):                                # SAC acts stochastically while it collects: policy mean + noise
        k[0] += 1
        return torch.clamp(policy(obs) + 0.3 * noise[k[0] % 8], -1.0, 1.0)

    ro = ug.Rollout(env, explore, ring)
    ro.capture(8)
    ro.replay(2)
    env.stats(reset=True)
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    reps = max(1, steps // 8)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    ro.replay(reps)
    e1.record()
    torch.cuda.synchronize(dev)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    st = env.stats(reset=True)
    if world > 1:
        st = ug.allreduce_stats(st, device=dev)
    ms = float(t.item())
    D = OBS_DIM[task]
    return {"value": n * world * reps * 8 / (ms * 1e-3), "unit": UNIT, "steps": reps * 8, "ms_per_step": ms / (reps * 8),
            "what": "actor MLP (bf16 torch matmuls, the shipped policy + exploration noise) -> urgym_step -> urgym_replay_write "
                    "into a device ring of 4 N transitions; CUDA graphs of 8 steps; no host traffic",
            "replay_bytes_per_env_step": 2 * (2 * 4 * D + 24 + 4 + 2), "episode_stats": ug.summarize(st)}


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    import urgym_b200 as ug

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the simulator has no CPU path (use --impl reference for the CPU arm)")
    numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else {"bound": False, "skipped": True}
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on stdout; the contract is ONE JSON line there, so everything but the final
        # print goes to stderr
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    n = args.envs_per_gpu
    steps, warm = max(args.steps, 1), max(args.warmup, 3)
    m = measure_device(ug, torch, dist, dev, local_rank, rank, world, args.task, n, args.geometry, steps, warm, args.chains,
                       args.graph_steps, rank * n)
    roofline = roofline_of(m, args.task, n, world, args.geometry)
    e2e = measure_e2e(torch, dist, dev, world, m["env"], m["ring"], n, args.task, max(args.e2e_steps, 2))
    cfg = workload_config(args, world)
    cfg.update({"chains_per_gpu": m["chains"], "graph_steps": m["graph_steps"], "numa": numa})
    line = {"metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warm,
            "ms_per_step": m["ms_max"] / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": cfg, "roofline": roofline,
            "e2e": e2e, "gpu_launches": m["launches"], "clocks": m["clocks"],
            "episode_stats": ug.summarize(m["stats"])}
    if not args.no_extra_configs:
        # the other BASELINE.json configs, measured the same way (shorter), and C5's rollout loop
        extra = {}
        if args.task == "UR5DynReach-v1" and args.geometry == "capsule":
            try:
                extra["C5_rollout_loop"] = dict(measure_rollout(ug, torch, dist, dev, world, m, args.task, n, 64),
                                                config=f"{args.task}, {n} envs per GPU x {world} GPU(s), policy-driven, replay ring on the device")
            except Exception as e:
                extra["C5_rollout_loop"] = {"error": repr(e)}
        m["env"].close()
        todo = [("C2_Ori_65536", "UR5OriReach-v1", 65536, 1), ("C3_Obs_1Mi", "UR5ObsReach-v1", 1 << 20, 1)]
        if world >= 2:
            todo.append(("C4_Sta_4Mi_sharded", "UR5StaReach-v1", (4 << 20) // world, world))
        for name, task, ne, need in todo:
            try:
                if need == 1 and world > 1:
                    # single-GPU configs under torchrun: every rank runs its own copy, the line reports ONE GPU
                    mm = measure_device(ug, torch, None, dev, local_rank, rank, 1, task, ne, "capsule", 104, 3, args.chains, 32, 0)
                    w_eff = 1
                else:
                    mm = measure_device(ug, torch, dist, dev, local_rank, rank, world, task, ne, "capsule", 104, 3, args.chains, 32, rank * ne)
                    w_eff = world
                extra[name] = {"config": f"{task}, {ne} envs per GPU x {w_eff} GPU(s), capsule geometry, random actions",
                               "value": mm["value"], "unit": UNIT, "n_gpus": w_eff, "steps": mm["steps"],
                               "ms_per_step": mm["ms_max"] / mm["steps"], "roofline": roofline_of(mm, task, ne, w_eff, "capsule"),
                               "episode_stats": ug.summarize(mm["stats"])}
                if ne < (1 << 19):
                    extra[name]["note"] = ("launch-bound: the per-step traffic of this batch fits the 126 MB L2, so the HBM fraction "
                                           "is not a roofline statement here (SURVEY.md 7.3-5)")
                mm["env"].close()
            except Exception as e:
                extra[name] = {"error": repr(e)}
        if world < 2:
            extra["C4_Sta_4Mi_sharded"] = {"skipped": "BASELINE.json shards this config over 2/4/8 GPUs; run with --gpus >= 2"}
        if rank == 0 and args.task == "UR5DynReach-v1":
            # the reference's own link geometry (convex hulls, GJK): the parity path, bound by the FP32 pipe, with the
            # capsule path's disagreement against it measured live on the hull path's own state distribution
            try:
                nh = 1 << 18
                mm = measure_device(ug, torch, None, dev, local_rank, rank, 1, args.task, nh, "hull", 24, 3, 1, 8, 0)
                extra["hull_geometry_Dyn"] = {"config": f"{args.task}, {nh} envs on one GPU, hull geometry (the reference's meshes), random actions",
                                              "value": mm["value"], "unit": UNIT, "n_gpus": 1, "steps": mm["steps"],
                                              "ms_per_step": mm["ms_max"] / mm["steps"], "roofline": roofline_of(mm, args.task, nh, 1, "hull")}
                mm["env"].close()
                sys.path.insert(0, os.path.join(ROOT, "tools"))
                import disagreement
                extra["hull_geometry_Dyn"]["capsule_vs_hull"] = disagreement.measure(args.task, 1 << 16, 20, warmup=60, device=local_rank)
            except Exception as e:
                extra["hull_geometry_Dyn"] = {"error": repr(e)}
        if rank == 0:
            # the motor-driven env (POSITION_CONTROL motors over 20 dynamic substeps), compute-bound
            try:
                nm = 1 << 18
                mv = ug.UR5MotorVecEnv(nm, device=local_rank, seed=0)
                mv.reset()
                gm = torch.Generator(device=dev).manual_seed(77)
                acts = [torch.rand((nm, 6), device=dev, generator=gm) * 2 - 1 for _ in range(4)]
                for k in range(4):
                    mv.step(acts[k])
                torch.cuda.synchronize(dev)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(24):
                    mv.step(acts[k % 4])
                e1.record()
                torch.cuda.synchronize(dev)
                mms = e0.elapsed_time(e1) / 24
                extra["motor_UR5IAIReach"] = {"config": f"UR5IAIReach-v1 (robot UR5, position motors over 20 substeps), {nm} envs on one GPU, "
                                                        "random actions, auto-reset on", "value": nm / (mms * 1e-3), "unit": UNIT, "n_gpus": 1,
                                              "steps": 24, "ms_per_step": mms, "bound": "fp32 (20 x [7 recursive Newton-Euler passes + 50 "
                                              "Gauss-Seidel sweeps] per env step; 60 B of state per env)",
                                              "episode_stats": ug.summarize(mv.stats())}
                mv.close()
            except Exception as e:
                extra["motor_UR5IAIReach"] = {"error": repr(e)}
        line["configs"] = extra
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            procs = os.cpu_count() or 1
            rate, per_proc, cdt = calibrate_cpu(args.task, args.geometry, args.cpu_seconds, procs)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
                                    "sample": f"{procs} processes x {per_proc} env steps of {args.task} on the oracle port "
                                              f"({cdt:.1f} s wall); PyBullet itself is not installable here"}
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
