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
    """Rollout collection (config C5, tde_step_rollout): the step without its plain observation, plus the
    frame stack written into the next buffer slot (n_stack frames) from the previous slot (n_stack - 1
    frames read), plus the action row kept in the buffer (8 B)."""
    frame = 3 * 64 * 64
    return bytes_per_env_step(num_agents, False) + (2 * n_stack - 1) * frame + 8
