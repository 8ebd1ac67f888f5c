"""A/B of tde_step_host's two observation paths on config C3 (16,384 envs x 32 agents): RGB planes over PCIe (12 KB per
env) against the 4-bit class image (2 KB per env) expanded by host threads.  Prints one JSON line per arm:
env-steps/s through host buffers, host threads, chunks.  Both arms fill the caller's pinned buffer with the same bytes
(checked once per arm against each other before timing).

usage: python tools/e2e_ab.py [--envs 16384] [--steps 20] [--threads 4,8,16] [--out gpurun_out/e2e_ab.json]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torchdriveenv_b200 import scenarios as S          # noqa: E402
from torchdriveenv_b200.engine import Engine           # noqa: E402


def run_arm(ss, E, A, steps, mode, threads, chunks, acts, simd=""):
    os.environ["TDE_HOST_OBS"] = mode
    if simd:
        os.environ["TDE_HOST_SIMD"] = simd
    else:
        os.environ.pop("TDE_HOST_SIMD", None)
    if threads:
        os.environ["TDE_HOST_THREADS"] = str(threads)
    if chunks:
        os.environ["TDE_HOST_CHUNKS"] = str(chunks)
    else:
        os.environ.pop("TDE_HOST_CHUNKS", None)
    eng = Engine(ss, E, A, auto_reset=1)          # a new handle: the thread count is read when the staging is allocated
    eng.reset(seed=1)
    for k in range(3):
        out = eng.step_host(acts[k])
    digest = int(np.frombuffer(out[0].tobytes(), np.uint64).sum())
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for k in range(steps):
        eng.step_host(acts[k % len(acts)])
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    eng.close()
    return dict(mode=mode, simd=simd or "widest", host_threads=threads or "default", chunks=chunks or "default", envs=E, steps=steps,
                env_steps_per_s=E * steps / dt, ms_per_step=dt / steps * 1e3,
                pcie_gbs=E * ((12288 if mode == "rgb" else 2048) + 82) * steps / dt / 1e9, obs_digest_after_3_steps=digest)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--threads", default="2,4,8,16")
    ap.add_argument("--chunks", default="0")
    ap.add_argument("--simd", default="widest", help="comma list of scalar|avx2|avx512|widest for the class-image arms")
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    E, A = args.envs, 32
    ss = S.traffic_lights(A)
    rng = np.random.default_rng(0)
    acts = [np.stack([rng.uniform(-1, 1, E), rng.uniform(-0.3, 0.3, E)], 1).astype(np.float32) for _ in range(8)]
    rows = [run_arm(ss, E, A, args.steps, "rgb", 0, 0, acts)]
    for ch in [int(c) for c in args.chunks.split(",")]:
        for th in [int(t) for t in args.threads.split(",")]:
            for simd in args.simd.split(","):
                rows.append(run_arm(ss, E, A, args.steps, "classes", th, ch, acts, simd if simd != "widest" else ""))
    assert len({r["obs_digest_after_3_steps"] for r in rows}) == 1, "the arms disagree on the observation bytes"
    for r in rows:
        print(json.dumps(r), flush=True)
    if args.out:
        with open(args.out, "w") as f:
            json.dump(dict(host_cpus=os.cpu_count(), rows=rows), f, indent=1)


if __name__ == "__main__":
    main()
