"""Algorithmic HBM bytes per env-step — the single formula used by bench.py and DESIGN.md
(SURVEY.md §8d): per agent 45 B (state in/out 32, length/width/lr 12, present 1), per NPC 16 B of
replay/target state, per env 50 B (action 8, reward 4, flags 2, infractions 12, bookkeeping 24),
observation 3*64*64 B.  Meshes, waypoints, stop lines and light schedules are per-scenario tables
shared by every env (L2 resident) and count as 0."""


def bytes_per_env_step(num_agents: int, render: bool) -> int:
    return 45 * num_agents + 16 * (num_agents - 1) + 50 + (3 * 64 * 64 if render else 0)


def c4_bytes_per_env(num_agents: int) -> int:
    """collision/offroad micro-bench: read x,y,psi,length,width,present (21 B), write 2 floats (8 B)."""
    return 29 * num_agents


def rollout_bytes_per_env_step(num_agents: int, n_stack: int) -> int:
    """Rollout collection (config C5): ALGORITHMIC bytes only - the step with ONE new frame per env (12,288 B) plus the
    action row kept in the buffer (8 B).  How often the implementation stores that frame (n_stack times in scatter mode;
    shift mode additionally reads n_stack - 1 frames) is traffic, not algorithm: see rollout_traffic_per_env_step."""
    return bytes_per_env_step(num_agents, True) + 8


def rollout_traffic_per_env_step(num_agents: int, n_stack: int, frame_copy: str = "scatter") -> int:
    """Bytes the rollout step actually moves per env: the ring of frames writes the frame once; scatter mode writes it
    n_stack times; shift mode reads n_stack - 1 frames and writes n_stack."""
    frame = 3 * 64 * 64
    frames = 1 if frame_copy == "ring" else n_stack if frame_copy == "scatter" else 2 * n_stack - 1
    return bytes_per_env_step(num_agents, False) + frames * frame + 8
