#!/usr/bin/env python
"""bench.py — env-steps/sec of the lockstep simulation step with 64x64 birdview observations.

Workload (BASELINE.json configs[2], "C3"): 16,384 envs x 32 agents per GPU, Traffic Lights scenario
(synthetic lane mesh + stop lines + light schedule + log-replay NPCs), kinematics + all-pairs
collision + offroad/wrong-way/red-light + reward/termination + birdview rendering, auto-reset on.
A "step" is one tde_step over the whole batch.  Envs shard across GPUs with no collective on the
step path (weak scaling: 16,384 envs per GPU); NCCL only reduces the episode statistics at the end.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...  # CPU oracle port on the host cores
  (N > 1: launched by torch.distributed.run, one rank per GPU)

Prints ONE JSON line (rank 0).  At N = 1 the default line also carries
  other_configs          the other BASELINE configs (C2, C4, C5; C1 as the latency of one env behind the gym.Env API) measured the same way on this GPU,
  pursuit                the C3 step under a pure-pursuit policy (long episodes: junctions, red lights, late waypoints),
  cpu_baseline           the batched C oracle port on all host cores,
  cpu_baseline_per_env   BASELINE config C1 the way the reference runs: one process per core, one env each (B = 1),
                         the reference's call pattern per step (gym_env.py:369-437).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env_steps_per_sec_64x64_birdview"
UNIT = "env-steps/s"
WORKLOADS = {
    # name: (envs per GPU, agents, render)
    "c3": (16384, 32, True),
    "c2": (1024, 16, False),
    "c4": (65536, 64, False),   # stateless collision + offroad micro-benchmark (tde_collision_boxes / tde_offroad_boxes)
    "c5": (8192, 8, True),      # rollout collection: 65,536 envs over 8 GPUs, training-scenario mix, 3-frame stack into a GPU rollout buffer
}
C5_N_STACK, C5_N_STEPS = 3, 32


def build_scenarios(workload: str):
    from torchdriveenv_b200 import scenarios as S
    if workload == "c3":
        return S.traffic_lights(32), "C3: 16384 envs x 32 agents per GPU, Traffic Lights, birdview 3x64x64, auto-reset"
    if workload == "c2":
        return S.roundabout(16), "C2: 1024 envs x 16 agents, Roundabout, kinematics+collision+offroad+reward, no render"
    if workload == "c5":
        return S.training_mix(100, 8), ("C5: rollout collection, 8192 envs x 8 agents per GPU (65,536 over 8 GPUs), mix of the reference's 100 "
                                        "training polylines, 3-frame stack stored (scatter mode, no frame moved) into a GPU-resident rollout buffer, uniform random policy")
    raise ValueError(workload)


def config_dict(workload: str, desc: str, E: int, A: int, render: bool, world: int) -> dict:
    """The `config` object of the JSON line: the same for the CUDA arm and the reference arm."""
    return dict(workload=desc, envs_per_gpu=E, agents=A, render=render, global_envs=E * world,
                parallelism=f"env-sharded x{world}, no collectives on the step path",
                l2="per-step working set (obs write %.0f MB + state) exceeds the 126 MB L2" % (E * 12288 / 1e6)
                if render else "inputs smaller than L2 (C2 is latency-bound by design)",
                actions="U(-1,1) x U(-0.3,0.3), fresh per step, resident in HBM")


def make_actions(E: int, n: int, seed: int) -> np.ndarray:
    """a ~ U(-1,1), beta ~ U(-0.3,0.3) (the reference action space, gym_env.py:83-84)."""
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(-1, 1, (n, E)), rng.uniform(-0.3, 0.3, (n, E))], -1).astype(np.float32)


class ClockSampler:
    """SM clock, power and throttle reasons sampled DURING the timed region: NVML in a thread every 2 ms (a C3 run of
    100 steps lasts 18 ms, too short for `nvidia-smi -lms`), `nvidia-smi` as the fallback when NVML is not importable."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.rows, self.proc, self.thread, self.nvml, self.alive = [], None, None, None, True
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it lists plain indices
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [int(x) for x in vis.split(",")] if vis and all(x.strip().isdigit() for x in vis.split(",")) else None
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(ids[index] if ids and index < len(ids) else index)
            self.nvml = pynvml
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            sys.setswitchinterval(5e-4)   # the launching loop keeps the GIL busy: let the 2 ms poller in more often than every 5 ms
            self.thread = threading.Thread(target=self._poll_nvml, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll_nvml(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        while self.alive:
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                pw = float(n.nvmlDeviceGetPowerUsage(self.handle)) / 1e3
                try:
                    mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                flags = ["Active" if mask & b else "Not Active" for b, _ in bits]
                self.rows.append((time.time(), ", ".join([f"{sm}", f"{self.max_sm}", f"{pw}"] + flags)))
            except Exception:
                pass
            time.sleep(0.002)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None and self.nvml is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=[], samples=0)
        if self.nvml is not None:
            self.alive = False
            self.thread.join(timeout=1.0)
        else:
            time.sleep(0.12)
            self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1]
        window = "timed region"
        if len(rows) < 3:   # short timed region: fall back to every sample taken while the GPU was busy (warm-up .. end)
            rows, window = [r for (t, r) in self.rows if t <= t1 + 0.1], "warm-up + timed region"
        self.window = window
        sm, mx, reasons, pw = [], [], set(), []
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except Exception:
                continue
            for n, v in zip(self.NAMES, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=float(max(mx)) if mx else None,
                    power_w_max=float(max(pw)) if pw else None, reasons=sorted(reasons), samples=len(sm), window=self.window,
                    source="nvml" if self.nvml is not None else "nvidia-smi")


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(workload: str):
    """dram bytes per launch of the step kernel from the committed ncu capture, or None."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(p)).get(workload)
    except Exception:
        return None


def cpu_oracle_throughput(workload: str, budget_s: float, seed: int = 0):
    """The CPU oracle port (reference algorithm: brute-force mesh tests, per-primitive polygon fill) on a
    bounded sample of the same workload, all host threads (OpenMP)."""
    from oracle import oracle as O
    from torchdriveenv_b200._capi import default_config
    O.set_num_threads(os.cpu_count() or 1)     # torch.distributed.run exports OMP_NUM_THREADS=1
    _, A, render = WORKLOADS[workload]
    ss, _ = build_scenarios(workload)
    E = 1024
    env = O.OracleEnvSet(default_config(num_envs=E, max_agents=A, auto_reset=1), ss.pack(A))
    env.reset(seed=seed)
    acts = make_actions(E, 64, seed)
    env.step(acts[0], render=render)          # warm-up
    t0 = time.perf_counter()
    n = 0
    while True:
        env.step(acts[(n + 1) % 64], render=render)
        n += 1
        if time.perf_counter() - t0 >= budget_s or n >= 10000:
            break
    dt = time.perf_counter() - t0
    threads = O.num_threads()
    return dict(value=E * n / dt, unit=UNIT, cores=threads, kind="port",
                sample=f"{E} envs x {n} steps of the same workload ({dt:.1f} s), C oracle with OpenMP over envs, "
                       f"{threads} threads on {os.cpu_count()} visible cores")


def per_env_worker(seconds: float, seed: int) -> None:
    """One worker of the per-env CPU baseline: ONE env (B = 1, BASELINE config C1: Three Way, ego + 8 NPCs, 64x64
    birdview) stepped through the oracle with the reference's call pattern (gym_env.py:369-437): per step 12 get_state
    copies, sim.step, one render, the three infraction metrics evaluated for is_terminated and again for get_info,
    reward / waypoint arithmetic in Python floats, a reset when the episode ends.  Prints the steps it made."""
    import math
    os.environ["OMP_NUM_THREADS"] = "1"
    from oracle import oracle as O
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200._capi import default_config
    O.set_num_threads(1)
    ss = S.three_way(6)
    A = 9
    env = O.OracleEnvSet(default_config(num_envs=1, max_agents=A, auto_reset=0), ss.pack(A))
    wps = np.asarray(ss.scenarios[0].waypoints, np.float64)
    rng = np.random.default_rng(seed)
    get_state = lambda: env.state[:, :1].copy()          # simulator.get_state(): B x 1 x 4 (NPCs hidden)
    episode = 0

    def reset():
        nonlocal episode
        env.reset(seed=seed + episode); episode += 1
        return env.render(), 1, 0

    obs, target, steps_in_ep = reset()
    n = 0
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        a = np.array([[rng.uniform(-1, 1), rng.uniform(-0.3, 0.3)]], np.float32)
        st = get_state(); lx, ly, lpsi, lv = (float(st[0, 0, k]) for k in range(4))      # :371-375 (1)
        env.kinematics(a); steps_in_ep += 1                                              # :117 sim.step
        obs = env.render()                                                               # :122-124 get_obs
        x, y, psi = float(get_state()[0, 0, 0]), float(get_state()[0, 0, 1]), float(get_state()[0, 0, 2])   # :397-399 (3)
        d = math.dist((x, y), (lx, ly))
        reward = (1.0 if d > 0.5 else 0.0) - 25.0 * (1.0 - math.cos(psi - lpsi))
        if target < len(wps):                                                            # :391-394 check_reach_target (2)
            if math.dist((float(get_state()[0, 0, 0]), float(get_state()[0, 0, 1])), tuple(wps[target])) < 3.0:
                reward += 100.0
        inf = env.compute_infractions()[0, 0]                                            # :413-417 is_terminated: three metrics
        terminated = bool(inf[1] > 0 or inf[0] > 0 or inf[2] > 0)
        truncated = steps_in_ep >= 200                                                   # :134-135
        s4 = [get_state() for _ in range(4)]                                             # :420-423 get_info (4)
        inf2 = env.compute_infractions()[0, 0]                                           # :427-429 the metrics again
        info = dict(offroad=inf2[1], collision=inf2[0], traffic_light_violation=inf2[2],
                    psi_smoothness=abs(lpsi - float(s4[2][0, 0, 2])) / 0.1, speed_smoothness=abs(lv - float(s4[3][0, 0, 3])) / 0.1)
        if target < len(wps):                                                            # :378-383 check_reach_target again (2)
            if math.dist((float(get_state()[0, 0, 0]), float(get_state()[0, 0, 1])), tuple(wps[target])) < 3.0:
                target += 1
        n += 1
        if terminated or truncated:
            obs, target, steps_in_ep = reset()
    print(json.dumps(dict(steps=n, seconds=time.perf_counter() - t0, reward=reward, info_keys=len(info))), flush=True)


def cpu_per_env_baseline(seconds: float):
    """BASELINE.md section 3 'CPU-ref (per-env)': one process per host core, one env each; aggregate env-steps/s."""
    cores = os.cpu_count() or 1
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--per-env-worker", str(seconds), "--seed", str(100 + k)],
                              stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True) for k in range(cores)]
    steps, secs = 0, []
    for pr in procs:
        out, _ = pr.communicate()
        try:
            d = json.loads(out.strip().splitlines()[-1])
            steps += int(d["steps"]); secs.append(float(d["seconds"]))
        except Exception:
            pass
    if not secs:
        return None
    return dict(value=steps / max(secs), unit=UNIT, cores=cores, processes=len(secs), kind="port",
                config="C1: single env per process, Three Way, ego + 8 NPCs (2 predetermined, 6 log-replay), 64x64 birdview",
                sample=f"{len(secs)} processes x 1 env, {max(secs):.1f} s each, {steps} env-steps in total; per step the reference's call "
                       "pattern (gym_env.py:369-437): 12 get_state, sim.step, 1 render, the three infraction metrics twice, "
                       "Python-float reward; C oracle, OMP_NUM_THREADS=1 per process")


def c4_inputs(E: int, A: int, seed: int):
    from torchdriveenv_b200 import scenarios as S
    st, at = S.scatter_boxes(E, A, size=200.0, seed=seed)
    patch = S.scatter_patch(200.0, 10.0)
    return st, at, patch


def cpu_oracle_c4(budget_s: float, seed: int = 0):
    """CPU oracle on a bounded sample of config C4 (all-pairs SAT + brute-force corner-to-mesh distance)."""
    from oracle import oracle as O
    O.set_num_threads(os.cpu_count() or 1)
    E, A = 512, 64
    st, at, patch = c4_inputs(E, A, seed)
    O.collision_boxes(st[:8], at[:8])
    t0 = time.perf_counter(); n = 0
    while True:
        O.collision_boxes(st, at)
        O.offroad_boxes(patch.road_tris, 0.5, st, at)
        n += 1
        if time.perf_counter() - t0 >= budget_s or n >= 10000:
            break
    dt = time.perf_counter() - t0
    threads = O.num_threads()
    return dict(value=E * n / dt, unit="envs/s", cores=threads, kind="port",
                sample=f"{E} envs x 64 agents x {n} passes of collision + offroad ({dt:.1f} s), C oracle with OpenMP, "
                       f"{threads} threads on {os.cpu_count()} visible cores")


def run_c4(args, rank: int, local_rank: int, world: int):
    """Config C4: 65,536 envs x 64 agents per GPU, all-pairs SAT + lane-mesh offroad, vs the HBM roofline."""
    import torch
    import torch.distributed as dist
    from torchdriveenv_b200 import scenarios as S
    from torchdriveenv_b200.engine import Engine
    from torchdriveenv_b200.roofline import c4_bytes_per_env
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    E, A, _ = WORKLOADS["c4"]
    st, at, patch = c4_inputs(E, A, seed=12 + rank)
    eng = Engine(S.ScenarioSet([patch], [S.make_scenario(0, [[5, 5], [50, 5]], 0, 0, "p")]), 1, 1, device=str(dev))
    st_d, at_d = torch.from_numpy(st).to(dev), torch.from_numpy(at).to(dev)
    pin_s, pin_a = torch.from_numpy(st).pin_memory(), torch.from_numpy(at).pin_memory()
    out_h = torch.zeros((2, E, A), dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream(dev)
    K, W = args.steps, args.warmup

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(W):
        eng.collision_boxes(st_d, at_d); eng.offroad_boxes(0, st_d, at_d)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(K):
        eng.collision_boxes(st_d, at_d); eng.offroad_boxes(0, st_d, at_d)
    e1.record(stream)
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item()) / K
    # end to end: boxes from pinned host memory, both results back to the host, every step
    K2 = max(3, min(K, args.e2e_steps))
    barrier()
    e0.record(stream)
    for _ in range(K2):
        s2, a2 = pin_s.to(dev, non_blocking=True), pin_a.to(dev, non_blocking=True)
        out_h[0].copy_(eng.collision_boxes(s2, a2), non_blocking=True)
        out_h[1].copy_(eng.offroad_boxes(0, s2, a2), non_blocking=True)
        torch.cuda.synchronize(dev)
    e1.record(stream)
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        bpe = c4_bytes_per_env(A)
        achieved = bpe * E / (ms * 1e-3) / 1e9
        cpu = cpu_oracle_c4(args.cpu_seconds) if world == 1 and not args.no_cpu_baseline else None
        line = dict(metric="c4_collision_offroad_envs_per_sec", value=E * world / (ms * 1e-3), unit="envs/s", n_gpus=world, steps=K, warmup=W,
                    ms_per_step=ms, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload="C4: 65536 envs x 64 agents per GPU, all-pairs SAT + corner-to-lane-mesh offroad on scattered boxes "
                                         "(200 m checkerboard patch)", envs_per_gpu=E, agents=A,
                                l2="inputs 134 MB + outputs 34 MB per step exceed the 126 MB L2"),
                    clocks=clocks, gpu_launches=2 * K,
                    e2e=dict(value=E * world * K2 / (float(e2e_ms.item()) * 1e-3), unit="envs/s", h2d_bytes_per_step=2 * E * A * 16,
                             d2h_bytes_per_step=2 * E * A * 4, steps=K2, api="tde_collision_boxes + tde_offroad_boxes on tensors copied from pinned host memory"),
                    roofline=dict(bound="hbm", kernel="tde_collision_kernel + tde_offroad_kernel", achieved=achieved, peak=peak, unit="GB/s",
                                  frac=achieved / peak, traffic=None, algorithmic_bytes_per_launch=bpe * E, bytes_per_env=bpe,
                                  avg_launch_ms=ms, peak_source=peak_src,
                                  note="issue-bound by design: ~2,016 pair tests and 256 point-to-mesh queries per env (SURVEY 8d)"),
                    cpu_baseline=cpu)
        return line
    return None


def run_c5(args, rank: int, local_rank: int, world: int):
    """Config C5: on-policy rollout collection (examples/rl_training.py:159-160,178-181) on the training-scenario
    mix, frame stack fused into the rollout-buffer store (tde_step_rollout), uniform random policy on the GPU."""
    import torch
    import torch.distributed as dist
    from torchdriveenv_b200.distributed import reduce_episode_stats, summarize
    from torchdriveenv_b200.engine import Engine
    from torchdriveenv_b200.rollout import RolloutCollector, uniform_policy
    from torchdriveenv_b200.roofline import rollout_bytes_per_env_step, rollout_traffic_per_env_step
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    E, A, _ = WORKLOADS["c5"]
    ss, desc = build_scenarios("c5")
    eng = Engine(ss, E, A, device=str(dev), auto_reset=1, env_index_offset=rank * E)
    T = C5_N_STEPS
    # a rollout = one CUDA graph replay; frames kept once each in a ring (the consumer gathers the stack), or scattered into
    # the n_stack stacked observations they belong to (--c5-frame-copy scatter)
    col = RolloutCollector(eng, T, n_stack=C5_N_STACK, seed=0, cuda_graph=True, frame_copy=args.c5_frame_copy)
    policy = uniform_policy(seed=1000 + rank)
    K = max(T, (args.steps // T) * T)          # whole rollouts
    W = max(3, -(-args.warmup // T))           # >= 3: the eager rollout, the captured one, one replay
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    for _ in range(W):
        col.collect(policy)
    barrier()
    launches0 = eng.num_kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.time()
    e0.record(stream)
    for _ in range(K // T):
        col.collect(policy)
    e1.record(stream)
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    gpu_launches = eng.num_kernel_launches() - launches0
    tmax = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item()) / K
    # end to end with a host-side policy: actions from pinned host memory every step, reward + flags back to the
    # host every step; the observations stay in the GPU rollout buffer, where the learner reads them
    K2 = max(3, min(K, args.e2e_steps))
    pin_act = torch.from_numpy(make_actions(E, 8, seed=2000 + rank)).pin_memory()
    host_rew, host_flags = torch.zeros(E, dtype=torch.float32).pin_memory(), torch.zeros((2, E), dtype=torch.uint8).pin_memory()
    b = col.buffer

    def host_policy_step(t, k):
        a = pin_act[k % 8].to(dev, non_blocking=True)
        if col.frame_copy == "ring":
            eng.step_into(a, b.frames[t + C5_N_STACK], reward=b.rewards[t], terminated=b.terminated[t], truncated=b.truncated[t])
        else:
            eng.step_rollout_scatter(a, b.observations, t, C5_N_STACK, reward=b.rewards[t], terminated=b.terminated[t], truncated=b.truncated[t])
        host_rew.copy_(b.rewards[t], non_blocking=True)
        host_flags[0].copy_(b.terminated[t], non_blocking=True)
        host_flags[1].copy_(b.truncated[t], non_blocking=True)
        torch.cuda.synchronize(dev)

    host_policy_step(0, 0)
    barrier()
    e0.record(stream)
    for k in range(K2):
        host_policy_step(k % T, k)
    e1.record(stream)
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    stats = reduce_episode_stats(eng.episode_stats(), device=dev)
    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        bpe = rollout_bytes_per_env_step(A, C5_N_STACK)
        achieved = bpe * E / (ms * 1e-3) / 1e9
        cpu = cpu_oracle_throughput("c5", args.cpu_seconds) if world == 1 and not args.no_cpu_baseline else None
        line = dict(metric=METRIC, value=E * world / (ms * 1e-3), unit=UNIT, n_gpus=world, steps=K, warmup=W * T, ms_per_step=ms,
                    higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                    config=dict(workload=desc, envs_per_gpu=E, agents=A, render=True, global_envs=E * world, n_stack=C5_N_STACK,
                                rollout_steps=T, rollout_buffer_bytes=b.nbytes(), frame_copy=col.frame_copy,
                                parallelism=f"env-sharded x{world}, no collectives on the step path",
                                l2="each step writes a fresh %.0f MB buffer slot: larger than the 126 MB L2" % (E * 9 * 4096 / 1e6)),
                    clocks=clocks, gpu_launches=2 * K,      # one physics + one render launch per step (replayed from the captured graph)
                    e2e=dict(value=E * world * K2 / (float(e2e_ms.item()) * 1e-3), unit=UNIT, h2d_bytes_per_step=E * 8, d2h_bytes_per_step=E * 6,
                             steps=K2, api="Engine.step_into / step_rollout_scatter with actions from pinned host memory and reward/terminated/truncated read back "
                                           "every step (observations stay in the GPU rollout buffer)"),
                    roofline=dict(bound="hbm", kernel="tde_render_kernel (+ tde_physics_kernel, one launch each per step)", achieved=achieved,
                                  peak=peak, unit="GB/s", frac=achieved / peak, traffic=None, algorithmic_bytes_per_launch=bpe * E,
                                  bytes_per_env_step=bpe, avg_launch_ms=ms, peak_source=peak_src,
                                  moved_bytes_per_env_step=rollout_traffic_per_env_step(A, C5_N_STACK, col.frame_copy),
                                  note="algorithmic = one new frame per env-step; the ring stores it once, the scatter store n_stack times"),
                    cpu_baseline=cpu, episode_stats=summarize(stats))
        return line
    return None


def run_reference(args, rank: int, world: int):
    """The reference arm: the CPU implementation of the path on the box's host cores - the C oracle port (the reference's
    own implementation needs torchdrivesim, absent offline), every host thread, rank 0 only.  Each bench "step" is a
    bounded sample of the workload: a 1,024-env slice stepped for the step's share of ~150 s (at least one oracle step)."""
    if rank != 0:
        return
    workload = args.workload
    if workload == "c4":
        r = cpu_oracle_c4(max(2.0, min(20.0, 3.0 * (args.steps + args.warmup))))
        print(json.dumps(dict(impl="reference", metric="c4_collision_offroad_envs_per_sec", value=r["value"], unit="envs/s", n_gpus=args.gpus,
                              steps=args.steps, warmup=args.warmup, ms_per_step=1e3 * 65536 / r["value"], higher_is_better=True, scaling="weak",
                              vs_baseline=None, dtype="f32", data="synthetic", config=dict(workload="C4 (bounded 512-env sample per pass)"),
                              cpu_baseline=r, e2e=dict(value=r["value"], unit="envs/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))), flush=True)
        return
    E, A, render = WORKLOADS[workload]
    ss, desc = build_scenarios(workload)
    from oracle import oracle as O
    from torchdriveenv_b200._capi import default_config
    threads = O.set_num_threads(os.cpu_count() or 1)      # torch.distributed.run exports OMP_NUM_THREADS=1 to its workers
    ES = 1024
    env = O.OracleEnvSet(default_config(num_envs=ES, max_agents=A, auto_reset=1), ss.pack(A))
    env.reset(seed=0)
    acts = make_actions(ES, 64, 0)
    env.step(acts[0], render=render)
    per_step_budget = min(5.0, 150.0 / max(1, args.steps + args.warmup))
    n_total, t_total, k_act = 0, 0.0, 0
    t_all = time.perf_counter()
    for k in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        n = 0
        while True:
            env.step(acts[k_act % 64], render=render)
            k_act += 1; n += 1
            if time.perf_counter() - t0 >= per_step_budget:
                break
        if k >= args.warmup:
            n_total += n; t_total += time.perf_counter() - t0
    value = ES * n_total / t_total
    cb = dict(value=value, unit=UNIT, cores=threads, kind="port",
              sample=f"each of the {args.steps} bench steps = {n_total / max(1, args.steps):.1f} oracle steps of a {ES}-env slice of the workload "
                     f"({ES * n_total} env-steps in {t_total:.1f} s); C oracle with OpenMP over envs, {threads} threads on {os.cpu_count()} visible cores")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=1e3 * t_total / max(1, args.steps),           # measured: wall time of one bench step (a bounded sample)
                oracle_ms_per_step=1e3 * t_total / max(1, n_total), envs_per_oracle_step=ES,
                ms_per_full_batch=1e3 * E * max(1, args.gpus) / value,    # what one pass over the whole workload would take at this rate
                higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32", data="synthetic",
                config=config_dict(workload, desc, E, A, render, max(1, args.gpus)),
                cpu_baseline=cb, e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0),
                wall_s=time.perf_counter() - t_all)
    print(json.dumps(line), flush=True)


def pursuit_policy(eng, waypoints, gen, v_want: float = 8.0):
    """Pure pursuit of each env's current target waypoint, on the GPU (the driver of tests/golden/make_reference_golden.py)."""
    import torch
    st = eng.get_state()[:, 0]
    tgt = eng.get_env_vars()[:, 2].long().clamp_(0, waypoints.shape[0] - 1)
    w = waypoints[tgt]
    err = torch.atan2(w[:, 1] - st[:, 1], w[:, 0] - st[:, 0]) - st[:, 2]
    err = torch.remainder(err + np.pi, 2 * np.pi) - np.pi
    acc = (0.8 * (v_want - st[:, 3])).clamp_(-1, 1)
    steer = (0.35 * err + 0.02 * torch.randn(st.shape[0], generator=gen, device=st.device)).clamp_(-0.3, 0.3)
    return torch.stack([acc, steer], 1).contiguous()


def measure_pursuit(eng, ss, E: int, K: int, W: int, dev, stream, seed: int = 77, preroll: int = 150):
    """The C3 step under a policy that survives: the envs are first rolled out with pure pursuit (untimed) and the actions
    recorded; then the same seed is replayed - the episodes are deterministic - and the last K steps are timed.  By then
    the batch is spread along the whole route (junctions, stop lines, late waypoints), not bunched at the start."""
    import torch
    from torchdriveenv_b200.distributed import summarize
    wps = torch.from_numpy(np.asarray(ss.scenarios[0].waypoints, np.float32)).to(dev)
    gen = torch.Generator(device=dev); gen.manual_seed(seed)

    def restart():
        v = eng.get_env_vars(); v[:, 5] = 0; eng.set_env_vars(v)      # episode counters: the replay draws the same episodes
        eng.reset(seed=seed)
        eng.episode_stats(reset=True)

    restart()
    tape = torch.empty((preroll + W + K, E, 2), dtype=torch.float32, device=dev)
    for k in range(tape.shape[0]):
        tape[k] = pursuit_policy(eng, wps, gen)
        eng.step(tape[k], render=False)
    restart()
    for k in range(preroll):
        eng.step(tape[k], render=False)
    for k in range(preroll, preroll + W):
        eng.step(tape[k])
    eng.episode_stats(reset=True)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for k in range(preroll + W, preroll + W + K):
        eng.step(tape[k])
    e1.record(stream)
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / K
    v = eng.get_env_vars()
    return dict(value=E / (ms * 1e-3), unit=UNIT, ms_per_step=ms, steps=K, policy="pure pursuit of the target waypoint at 8 m/s, "
                f"recorded then replayed (deterministic episodes), {preroll} untimed steps first",
                mean_env_step_at_timing=float(v[:, 1].float().mean().item()), mean_target_index=float(v[:, 2].float().mean().item()),
                episode_stats=summarize(eng.episode_stats()))


def run_cuda(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    from torchdriveenv_b200.distributed import reduce_episode_stats, summarize
    from torchdriveenv_b200.engine import Engine
    from torchdriveenv_b200.roofline import bytes_per_env_step

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    workload = args.workload
    E_total_strong, A, render = WORKLOADS[workload]
    E = E_total_strong if args.scaling == "weak" else max(1, E_total_strong // world)
    offset = rank * E
    ss, desc = build_scenarios(workload)
    eng = Engine(ss, E, A, device=str(dev), auto_reset=1, env_index_offset=offset)
    eng.reset(seed=0)
    K, W = args.steps, args.warmup
    n_act = min(K + W, 64)
    acts = torch.from_numpy(make_actions(E, n_act, seed=1000 + rank)).to(dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None   # started early: nvidia-smi needs ~0.2 s to emit its first row
    for k in range(W):
        eng.step(acts[k % n_act], render=render)
    barrier()
    launches0 = eng.num_kernel_launches()
    # CUDA events on the launching stream bracket the timed region; inside it one more event every EV steps gives
    # the per-launch durations (an event record between every pair of launches costs about 1 % of a C3 step)
    EV = 10
    # the timed steps are replayed from CUDA graphs of up to 50 steps each (Engine.capture_steps): a small batch (C2: 1,024 envs)
    # steps in less time than the host needs to issue a launch, and on C3 the replay closes most of the gap between the two
    # kernels of a step.  `eager_ms_per_step` (one tde_step call per step from Python) is reported beside it.
    use_graph = args.cuda_graph != "off"
    if use_graph:   # exactly K steps are timed: the block length is the largest divisor of K up to 50
        EV = max(d for d in range(1, 51) if K % d == 0)
        use_graph = EV >= 5
        if not use_graph:
            EV = 10
    if use_graph:
        graph = eng.capture_steps(acts[(torch.arange(EV) + W) % n_act], render=render)
        per_step_launches = (eng.num_kernel_launches() - launches0) // EV
        graph.replay()          # the capture itself ran nothing
        barrier()
    marks = sorted(set(range(0, K, EV)) | {K})
    evs = {k: torch.cuda.Event(enable_timing=True) for k in marks}
    t0 = time.time()
    evs[0].record(stream)
    if use_graph:
        for k in range(0, K, EV):
            graph.replay()
            evs[k + EV].record(stream)
    else:
        for k in range(K):
            eng.step(acts[(W + k) % n_act], render=render)
            if (k + 1) in evs:
                evs[k + 1].record(stream)
    barrier()
    t1 = time.time()
    clocks = sampler.stop(t0, t1) if sampler else None
    total_ms = evs[0].elapsed_time(evs[K])
    per_launch_ms = float(np.sum([evs[a].elapsed_time(evs[b]) for a, b in zip(marks[:-1], marks[1:])]) / K)
    gpu_launches = per_step_launches * K if use_graph else eng.num_kernel_launches() - launches0
    eager_ms = None
    if use_graph:
        n_eager = min(K, 200)
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record(stream)
        for k in range(n_eager):
            eng.step(acts[k % n_act], render=render)
        ee1.record(stream)
        barrier()
        eager_ms = ee0.elapsed_time(ee1) / n_eager
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms_max = float(tmax.item())
    value = E * world * K / (total_ms_max * 1e-3)

    # end to end through the host-buffer C-ABI call: H2D actions + D2H obs/reward/flags/info every step
    K2 = max(3, min(K, args.e2e_steps))
    host_acts = make_actions(E, 8, seed=2000 + rank)

    def host_steps(mode):
        """env-steps/s of the whole job through tde_step_host; mode = what crosses PCIe for the observation
        ("rgb": the 12 KB planes; "classes": the 2 KB class image, expanded to the same planes by host threads)"""
        os.environ["TDE_HOST_OBS"] = mode
        for k in range(2):
            eng.step_host(host_acts[k], render=render)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(K2):
            eng.step_host(host_acts[k % 8], render=render)
        e1.record(stream)     # tde_step_host returns with the host buffers filled: the host-side expansion is inside
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return E * world * K2 / (float(ms.item()) * 1e-3)

    h2d = E * 2 * 4
    small = 4 + 1 + 1 + 16 * 4
    host_modes = {m: host_steps(m) for m in (("rgb", "classes") if render else ("rgb",))}
    e2e_mode = args.host_obs if render else "rgb"
    os.environ["TDE_HOST_OBS"] = e2e_mode
    e2e_value = host_modes[e2e_mode]
    d2h = E * ((3 * 64 * 64 if render and e2e_mode == "rgb" else 64 * 32 if render else 0) + small)

    # the only collective of this path: episode statistics, after the timed regions
    stats = reduce_episode_stats(eng.episode_stats(), device=dev)

    if rank != 0:
        return None
    peak, peak_src = measured_peak_hbm()
    bpe = bytes_per_env_step(A, render)
    achieved = bpe * E / (per_launch_ms * 1e-3) / 1e9
    roof = dict(bound="hbm", kernel="tde_render_kernel (+ tde_physics_kernel, one launch each per step)" if render else "tde_physics_kernel (one launch per step)", achieved=achieved, peak=peak, unit="GB/s", frac=achieved / peak,
                traffic=ncu_traffic(workload), algorithmic_bytes_per_launch=bpe * E, bytes_per_env_step=bpe,
                avg_launch_ms=per_launch_ms, peak_source=peak_src)
    cpu = cpu_oracle_throughput(workload, args.cpu_seconds) if world == 1 and not args.no_cpu_baseline else None
    line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=total_ms_max / K,
                launch_mode="CUDA graphs of %d steps (Engine.capture_steps)" % EV if use_graph else "one tde_step call per step",
                eager_ms_per_step=eager_ms,
                higher_is_better=True, scaling=args.scaling, vs_baseline=None, dtype="f32", data="synthetic",
                config=config_dict(workload, desc, E, A, render, world),
                clocks=clocks, gpu_launches=int(gpu_launches),
                e2e=dict(value=e2e_value, unit=UNIT, h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h, steps=K2,
                         api="tde_step_host (pinned host buffers, copies inside the timed region)",
                         observation_over_pcie=e2e_mode, by_mode=host_modes,
                         # the host-buffer number is bound by the bytes that cross PCIe, not by the kernels: its own roofline
                         pcie_gbs_achieved=(h2d + d2h) * e2e_value / E / world / 1e9, pcie_note="per GPU; PCIe Gen5 x16 moves ~55 GB/s device to host",
                         # what bounds the class-image mode: the 12 KB per env the host threads write into the caller's planes
                         host_fill_gbs=(3 * 64 * 64 if render else 0) * e2e_value / 1e9, host_fill_note="whole box, bytes written into the callers' observation buffers"),
                roofline=roof, cpu_baseline=cpu, episode_stats=summarize(stats))
    if world == 1 and workload == "c3" and args.policy in ("both", "pursuit"):
        # a second engine that does not end an episode at an infraction (cfg.terminated_at_infraction = False, one of the
        # reference's own configurations): the pursuit driver runs the red lights, so with termination on it would be back
        # at the start every ~33 steps; this way the episodes run their 200 steps through every junction of the route
        eng.close()
        eng2 = Engine(ss, E, A, device=str(dev), auto_reset=1, env_index_offset=offset, terminated_at_infraction=0)
        line["pursuit"] = measure_pursuit(eng2, ss, E, max(50, min(K, 200)), 10, dev, stream)
        line["pursuit"]["config"] = "C3 with terminated_at_infraction = False"
        eng2.close()
    return line


def run_c1_gym_api(local_rank: int, steps: int = 400):
    """Config C1 the way a user of the reference runs it, on the GPU: ONE env behind the reference's gym.Env API
    (SingleAgentWrapper(WaypointSuiteEnv), gym_env.py:440-487), numpy observation / float reward / bool flags / info dict back
    on the host every step.  One env cannot fill a GPU: this is the latency of the drop-in call, reported beside the
    per-env CPU baseline (`cpu_baseline_per_env`) that runs the same call pattern on the host cores."""
    from torchdriveenv_b200 import gym_env as G, scenarios as S
    cfg = G.EnvConfig(device=f"cuda:{local_rank}", seed=0)
    env = G.SingleAgentWrapper(G.WaypointSuiteEnv(cfg, S.three_way(6)))
    rng = np.random.default_rng(0)
    env.reset(seed=0)
    acts = np.stack([rng.uniform(-1, 1, steps + 20), rng.uniform(-0.3, 0.3, steps + 20)], 1).astype(np.float32)
    n_done = 0
    for k in range(20):
        _, _, te, tr, _ = env.step(acts[k])
        if te or tr:
            env.reset()
    t0 = time.perf_counter()
    for k in range(steps):
        obs, r, te, tr, info = env.step(acts[20 + k])
        if te or tr:
            n_done += 1
            env.reset()
    dt = time.perf_counter() - t0
    env.close()
    return dict(metric="env_steps_per_sec_single_env_gym_api", value=steps / dt, unit="env-steps/s", steps=steps, us_per_step=dt / steps * 1e6,
                episodes_finished=n_done, config="C1: one env (Three Way, 6 agents) behind SingleAgentWrapper(WaypointSuiteEnv).step, host numpy outputs, "
                                                 "resets included; wall clock (the call synchronises)")


def slim(line):
    """A sub-line of `other_configs`: the measurement without the nested extras."""
    if line is None:
        return None
    keep = ("metric", "value", "unit", "steps", "warmup", "ms_per_step", "launch_mode", "eager_ms_per_step", "config", "clocks", "gpu_launches", "e2e", "roofline", "episode_stats")
    return {k: line[k] for k in keep if k in line}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cuda-graph", choices=("auto", "on", "off"), default="auto",
                    help="replay the timed steps from CUDA graphs of 50 steps (auto = on)")
    ap.add_argument("--c5-frame-copy", choices=("ring", "scatter", "shift"), default="ring", help="how the C5 rollout buffer keeps the frame stack")
    ap.add_argument("--host-obs", choices=("rgb", "classes"), default="classes",
                    help="what tde_step_host sends over PCIe for the observation in the e2e leg (both are timed, this one is e2e.value)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--policy", default="both", choices=["random", "pursuit", "both"],
                    help="C3: the headline value is always the uniform random policy of SURVEY 8d; 'both' / 'pursuit' add the pursuit line")
    ap.add_argument("--no-other-configs", action="store_true", help="N = 1, C3: skip the C2 / C4 / C5 lines and the per-env CPU baseline")
    ap.add_argument("--per-env-worker", type=float, default=None, help=argparse.SUPPRESS)
    ap.add_argument("--seed", type=int, default=0, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.per_env_worker is not None:
        per_env_worker(args.per_env_worker, args.seed)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world != args.gpus and world == 1 and args.gpus > 1:
        # plain `python bench.py --gpus N`: re-launch under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 2000), os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    import torch
    import torch.distributed as dist
    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    runner = {"c4": run_c4, "c5": run_c5}.get(args.workload, run_cuda)
    line = runner(args, rank, local_rank, world)
    if rank == 0 and line is not None and world == 1 and args.workload == "c3" and not args.no_other_configs:
        # the other BASELINE configs on the same GPU, each long enough for >= 3 clock samples in its timed region
        others = {}
        for name, steps, warm in (("c2", 3000, 100), ("c4", 60, 5), ("c5", 128, 32)):
            sub = argparse.Namespace(**vars(args))
            sub.workload, sub.steps, sub.warmup, sub.no_cpu_baseline, sub.e2e_steps, sub.policy = name, steps, warm, True, 5, "random"
            try:
                others[name] = slim({"c4": run_c4, "c5": run_c5}.get(name, run_cuda)(sub, 0, local_rank, 1))
            except Exception as ex:      # a failing extra must not take the headline line with it
                others[name] = dict(error=f"{type(ex).__name__}: {ex}")
        try:
            others["c1_gym_api"] = run_c1_gym_api(local_rank)
        except Exception as ex:
            others["c1_gym_api"] = dict(error=f"{type(ex).__name__}: {ex}")
        line["other_configs"] = others
        line["cpu_baseline_per_env"] = None if args.no_cpu_baseline else cpu_per_env_baseline(min(8.0, args.cpu_seconds))
    if rank == 0 and line is not None:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
