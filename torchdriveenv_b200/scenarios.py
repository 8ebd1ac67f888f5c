"""Scenario tables: the constructor-side inputs of the simulation hot path.

The reference builds these per episode inside ``build_simulator`` (gym_env.py:179-300) from
torchdrivesim's CARLA map assets (``find_map_config`` :312: road mesh :184, stop lines :183,
traffic-light controller :181), the waypoint suite (:314,326) and the replay car sequences
(:275-283).  The CARLA assets ship inside torchdrivesim and are not available offline
(SURVEY.md §8c), so the maps here are *synthetic*: lane ribbons extruded from the reference's own
waypoint polylines (torchdriveenv/data/validation_cases.yml:8-83), with lane markings, stop lines,
a light schedule and constant-speed log-replay NPCs.  Everything is seeded and deterministic.

Layout produced by :meth:`ScenarioSet.pack` is exactly ``tde_scenario_set`` in include/tde_b200.h.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

LIGHT_GREEN, LIGHT_YELLOW, LIGHT_RED = 0, 1, 2

# Waypoint polylines of the reference's validation suite (validation_cases.yml:8-83), rounded to
# millimetres.  Names follow README.md:21-23 / SURVEY.md §8c.
VALIDATION_POLYLINES: Dict[str, List[List[float]]] = {
    "three_way": [
        [-88.296, -118.394], [-96.581, -104.980], [-109.995, -101.429],
        [-128.538, -101.824], [-139.585, -101.824], [-147.475, -108.531],
    ],
    "parked_car": [
        [-27.822, -65.959], [-50.328, -65.959], [-70.877, -65.959], [-88.816, -65.632],
        [-102.515, -58.457], [-102.842, -43.779], [-102.842, -25.513],
    ],
    "chicken": [
        [152.294, -48.321], [152.798, -63.186], [153.050, -77.043], [153.050, -90.145],
        [153.302, -102.239],
    ],
    "roundabout": [
        [50.084, -5.234], [28.809, -4.891], [11.651, -15.528], [-5.506, -19.303],
        [-18.203, -10.381], [-20.262, 4.374], [-13.055, 16.385], [-7.565, 32.856],
        [-7.565, 50.013],
    ],
    "traffic_lights": [
        [196.553, 55.153], [169.712, 54.690], [157.680, 40.344], [157.680, 13.966],
        [142.872, -3.156], [123.898, -3.619], [104.925, -2.693], [87.339, 12.115],
        [87.802, 37.105], [87.339, 65.334], [87.339, 84.770],
    ],
}
# Predetermined NPCs of the Three Way case (validation_cases.yml:1291-1306): x y psi v / l w lr
THREE_WAY_NPCS = dict(
    states=[[-98.751, -79.730, 4.7094336, 10.0], [-116.110, -101.429, 3.1310002, 7.0]],
    attributes=[[5.0, 2.0, 2.0], [5.0, 2.0, 2.0]],
)


@dataclass
class MapData:
    """Static geometry shared by scenarios (what torchdrivesim's map config holds)."""
    road_tris: np.ndarray                      # (M, 8) ax ay bx by cx cy dir_cos dir_sin
    mark_tris: np.ndarray = field(default_factory=lambda: np.zeros((0, 6), np.float32))
    stoplines: np.ndarray = field(default_factory=lambda: np.zeros((0, 5), np.float32))  # x y len wid psi
    light_states: np.ndarray = field(default_factory=lambda: np.zeros((1, 0), np.uint8))  # (P, L)
    name: str = "map"


@dataclass
class ScenarioData:
    """One waypoint-suite entry (WaypointSuite :63-68 + Scenario :56-60)."""
    map_index: int
    waypoints: np.ndarray                      # (W, 2)
    start_heading: float
    agent_init: np.ndarray                     # (n, 4) x y psi v ; row 0 = ego placeholder
    agent_attr: np.ndarray                     # (n, 3) length width rear_axis_offset
    replay_states: Optional[np.ndarray] = None  # (T, n, 4)
    replay_mask: Optional[np.ndarray] = None    # (T, n)
    name: str = "scenario"


@dataclass
class ScenarioSet:
    maps: List[MapData]
    scenarios: List[ScenarioData]

    def max_agents(self) -> int:
        return max(int(s.agent_init.shape[0]) for s in self.scenarios)

    def pack(self, max_agents: int) -> Dict[str, np.ndarray]:
        """Flatten into the contiguous arrays of ``tde_scenario_set``."""
        A = int(max_agents)
        f32, i32 = np.float32, np.int32

        def offsets(counts):
            return np.concatenate([[0], np.cumsum(counts)]).astype(i32)

        maps, scen = self.maps, self.scenarios
        out: Dict[str, np.ndarray] = {}
        out["map_tri_offset"] = offsets([m.road_tris.shape[0] for m in maps])
        out["road_tris"] = _cat([m.road_tris.reshape(-1, 8) for m in maps], 8)
        out["map_mark_offset"] = offsets([m.mark_tris.shape[0] for m in maps])
        out["mark_tris"] = _cat([m.mark_tris.reshape(-1, 6) for m in maps], 6)
        out["map_stop_offset"] = offsets([m.stoplines.shape[0] for m in maps])
        out["stoplines"] = _cat([m.stoplines.reshape(-1, 5) for m in maps], 5)
        periods, lights = [], []
        for m in maps:
            ls = np.asarray(m.light_states, np.uint8)
            L = m.stoplines.shape[0]
            if ls.ndim != 2 or ls.shape[1] != L or ls.shape[0] < 1:
                ls = np.zeros((1, L), np.uint8)
            periods.append(ls.shape[0])
            lights.append(ls.reshape(-1))
        out["map_light_period"] = np.asarray(periods, i32)
        out["map_light_offset"] = offsets([x.size for x in lights])
        out["light_states"] = np.concatenate(lights).astype(np.uint8) if lights else np.zeros(0, np.uint8)

        Ns = len(scen)
        out["scen_map"] = np.asarray([s.map_index for s in scen], i32)
        out["scen_wp_offset"] = offsets([s.waypoints.shape[0] for s in scen])
        out["waypoints"] = _cat([np.asarray(s.waypoints, f32).reshape(-1, 2) for s in scen], 2)
        out["scen_start_heading"] = np.asarray([s.start_heading for s in scen], f32)
        out["scen_num_agents"] = np.asarray([s.agent_init.shape[0] for s in scen], i32)
        init = np.zeros((Ns, A, 4), f32)
        attr = np.zeros((Ns, A, 3), f32)
        attr[:, :, 0], attr[:, :, 1], attr[:, :, 2] = 5.0, 2.0, 1.0
        rep_T, rep_s, rep_m = [], [], []
        for k, s in enumerate(scen):
            n = s.agent_init.shape[0]
            if n > A:
                raise ValueError(f"scenario {s.name}: {n} agents > max_agents {A}")
            init[k, :n] = s.agent_init
            attr[k, :n] = s.agent_attr
            if s.replay_states is not None and s.replay_states.shape[0] > 0:
                T = s.replay_states.shape[0]
                rs = np.zeros((T, A, 4), f32)
                rm = np.zeros((T, A), np.uint8)
                rs[:, :n] = s.replay_states
                rm[:, :n] = s.replay_mask
                rm[:, 0] = 0  # the ego is never replayed (npc_mask[0] = False, gym_env.py:271)
                rep_T.append(T); rep_s.append(rs.reshape(-1, A * 4)); rep_m.append(rm)
            else:
                rep_T.append(0)
        out["agent_init"], out["agent_attr"] = init, attr
        out["scen_replay_T"] = np.asarray(rep_T, i32)
        out["scen_replay_offset"] = offsets(rep_T)
        out["replay_states"] = (np.concatenate(rep_s).astype(f32) if rep_s else np.zeros((0, A * 4), f32))
        out["replay_mask"] = (np.concatenate(rep_m).astype(np.uint8) if rep_m else np.zeros((0, A), np.uint8))
        return {k: np.ascontiguousarray(v) for k, v in out.items()}


def _cat(parts: Sequence[np.ndarray], width: int) -> np.ndarray:
    parts = [np.asarray(p, np.float32).reshape(-1, width) for p in parts]
    return np.concatenate(parts) if parts else np.zeros((0, width), np.float32)


# ----------------------------------------------------------------------------- geometry helpers

def resample_polyline(pts: np.ndarray, max_seg: float) -> np.ndarray:
    pts = np.asarray(pts, np.float64)
    out = [pts[0]]
    for a, b in zip(pts[:-1], pts[1:]):
        n = max(1, int(math.ceil(np.linalg.norm(b - a) / max_seg)))
        for k in range(1, n + 1):
            out.append(a + (b - a) * (k / n))
    return np.asarray(out)


def extend_polyline(pts: np.ndarray, before: float, after: float) -> np.ndarray:
    pts = np.asarray(pts, np.float64)
    d0 = pts[0] - pts[1]; d0 /= np.linalg.norm(d0)
    d1 = pts[-1] - pts[-2]; d1 /= np.linalg.norm(d1)
    return np.concatenate([[pts[0] + d0 * before], pts, [pts[-1] + d1 * after]])


def _miter_normals(pts: np.ndarray) -> np.ndarray:
    """Per-vertex offset direction so that ribbons of consecutive segments join without gaps."""
    seg = pts[1:] - pts[:-1]
    seg /= np.linalg.norm(seg, axis=1, keepdims=True)
    nrm = np.stack([-seg[:, 1], seg[:, 0]], axis=1)
    out = np.zeros_like(pts)
    out[0], out[-1] = nrm[0], nrm[-1]
    for i in range(1, len(pts) - 1):
        m = nrm[i - 1] + nrm[i]
        ln = np.linalg.norm(m)
        if ln < 1e-9:
            out[i] = nrm[i]
            continue
        m /= ln
        scale = 1.0 / max(0.5, float(np.dot(m, nrm[i])))
        out[i] = m * scale
    return out


def ribbon(pts: np.ndarray, off_a: float, off_b: float, reverse_dir: bool = False) -> np.ndarray:
    """Triangulated strip between lateral offsets off_a < off_b of a polyline; (M, 8) road triangles
    with the lane direction of each segment (reversed for an oncoming lane)."""
    pts = np.asarray(pts, np.float64)
    nrm = _miter_normals(pts)
    tris = []
    for i in range(len(pts) - 1):
        d = pts[i + 1] - pts[i]
        d /= np.linalg.norm(d)
        if reverse_dir:
            d = -d
        p0a, p0b = pts[i] + nrm[i] * off_a, pts[i] + nrm[i] * off_b
        p1a, p1b = pts[i + 1] + nrm[i + 1] * off_a, pts[i + 1] + nrm[i + 1] * off_b
        tris.append([*p0a, *p0b, *p1b, d[0], d[1]])
        tris.append([*p0a, *p1b, *p1a, d[0], d[1]])
    return np.asarray(tris, np.float32).reshape(-1, 8)


def strip_marking(pts: np.ndarray, offset: float, width: float, dash: Optional[float] = None,
                  gap: float = 0.0) -> np.ndarray:
    """Lane-marking triangles along a polyline at a lateral offset; dashed when dash is given."""
    pts = np.asarray(pts, np.float64)
    nrm = _miter_normals(pts)
    tris = []
    s_acc = 0.0
    for i in range(len(pts) - 1):
        a, b = pts[i], pts[i + 1]
        L = float(np.linalg.norm(b - a))
        if dash is None:
            pieces = [(0.0, 1.0)]
        else:
            pieces = []
            period = dash + gap
            k0 = int(math.floor(s_acc / period))
            s = k0 * period
            while s < s_acc + L:
                lo, hi = max(s, s_acc), min(s + dash, s_acc + L)
                if hi > lo:
                    pieces.append(((lo - s_acc) / L, (hi - s_acc) / L))
                s += period
        for (u0, u1) in pieces:
            q0, q1 = a + (b - a) * u0, a + (b - a) * u1
            n0 = nrm[i] * (1 - u0) + nrm[i + 1] * u0
            n1 = nrm[i] * (1 - u1) + nrm[i + 1] * u1
            p0a, p0b = q0 + n0 * (offset - width / 2), q0 + n0 * (offset + width / 2)
            p1a, p1b = q1 + n1 * (offset - width / 2), q1 + n1 * (offset + width / 2)
            tris.append([*p0a, *p0b, *p1b])
            tris.append([*p0a, *p1b, *p1a])
        s_acc += L
    return np.asarray(tris, np.float32).reshape(-1, 6)


def ring_road(center, r_in: float, r_out: float, n: int = 48, clockwise: bool = False) -> np.ndarray:
    cx, cy = center
    tris = []
    for k in range(n):
        a0, a1 = 2 * math.pi * k / n, 2 * math.pi * (k + 1) / n
        am = 0.5 * (a0 + a1)
        d = (-math.sin(am), math.cos(am))
        if clockwise:
            d = (-d[0], -d[1])
        p = lambda r, a: (cx + r * math.cos(a), cy + r * math.sin(a))
        i0, o0, i1, o1 = p(r_in, a0), p(r_out, a0), p(r_in, a1), p(r_out, a1)
        tris.append([*i0, *o0, *o1, *d])
        tris.append([*i0, *o1, *i1, *d])
    return np.asarray(tris, np.float32).reshape(-1, 8)


def subdivide_long_triangles(tris: np.ndarray, max_edge: float) -> np.ndarray:
    """Split triangles until no edge exceeds max_edge (the rasteriser's fixed-point range needs
    edges <= 200 m at the default 35 m field of view)."""
    tris = np.asarray(tris, np.float32)
    width = tris.shape[1]
    work = [t for t in tris]
    out = []
    while work:
        t = work.pop()
        v = t[:6].reshape(3, 2).astype(np.float64)
        e = [np.linalg.norm(v[(k + 1) % 3] - v[k]) for k in range(3)]
        k = int(np.argmax(e))
        if e[k] <= max_edge:
            out.append(t)
            continue
        a, b, c = v[k], v[(k + 1) % 3], v[(k + 2) % 3]
        m = 0.5 * (a + b)
        for tri in ((a, m, c), (m, b, c)):
            nt = t.copy()
            nt[:6] = np.asarray(tri, np.float32).reshape(-1)
            work.append(nt)
    return np.asarray(out, np.float32).reshape(-1, width)


# ----------------------------------------------------------------------------- traffic-light programs

def compile_light_program(phases, num_stoplines: int, dt: float = 0.1) -> np.ndarray:
    """Traffic-light controller program -> the periodic schedule table of ``tde_scenario_set``.

    The reference ticks a finite-state controller once per step inside IAIWrapper (gym_env.py:181-189,
    290-291): a cycle of phases, each holding every light in one state for a duration.  ``phases`` is a
    sequence of ``(duration_seconds, states)`` with ``states`` either a list of ``num_stoplines`` states
    or a dict ``{stopline_index: state}`` (missing lights stay as in the previous phase, initially red);
    states are ``LIGHT_GREEN / LIGHT_YELLOW / LIGHT_RED`` or the strings ``"green" / "yellow" / "red"``.
    Returns uint8 ``[period_steps][num_stoplines]``: row t is what the lights show ``t`` steps into the
    cycle, which is all the step kernel needs (env time = (steps + phase offset) mod period)."""
    names = {"green": LIGHT_GREEN, "yellow": LIGHT_YELLOW, "red": LIGHT_RED}
    cur = np.full(num_stoplines, LIGHT_RED, np.uint8)
    rows = []
    for duration, states in phases:
        if isinstance(states, dict):
            for k, v in states.items():
                cur[int(k)] = names[v] if isinstance(v, str) else int(v)
        else:
            vals = [names[v] if isinstance(v, str) else int(v) for v in states]
            if len(vals) != num_stoplines:
                raise ValueError(f"phase lists {len(vals)} lights, the map has {num_stoplines} stop lines")
            cur = np.asarray(vals, np.uint8)
        n = max(1, int(round(float(duration) / dt)))
        rows += [cur.copy()] * n
    if not rows:
        return np.zeros((1, num_stoplines), np.uint8)
    out = np.stack(rows).astype(np.uint8)
    if out.max(initial=0) > LIGHT_RED:
        raise ValueError("light states must be 0 (green), 1 (yellow) or 2 (red)")
    return out


# ----------------------------------------------------------------------------- path following (NPC replay)

class _Path:
    def __init__(self, pts: np.ndarray):
        self.pts = np.asarray(pts, np.float64)
        seg = self.pts[1:] - self.pts[:-1]
        self.len = np.linalg.norm(seg, axis=1)
        self.cum = np.concatenate([[0.0], np.cumsum(self.len)])
        self.total = float(self.cum[-1])

    def at(self, s: float):
        s = min(max(s, 0.0), self.total - 1e-6)
        i = int(np.searchsorted(self.cum, s, side="right") - 1)
        i = min(i, len(self.len) - 1)
        u = (s - self.cum[i]) / self.len[i]
        p = self.pts[i] + (self.pts[i + 1] - self.pts[i]) * u
        d = self.pts[i + 1] - self.pts[i]
        return p, math.atan2(d[1], d[0])


def offset_polyline(pts: np.ndarray, off: float) -> np.ndarray:
    pts = np.asarray(pts, np.float64)
    return pts + _miter_normals(pts) * off


def rollout_replay(path: _Path, s0: float, speed: float, T: int, dt: float = 0.1, reverse: bool = False):
    """Constant-speed path following = a synthetic log: (T, 4) states x y psi v."""
    out = np.zeros((T, 4), np.float32)
    for t in range(T):
        s = s0 + (-speed if reverse else speed) * dt * t
        p, psi = path.at(s)
        if reverse:
            psi = psi + math.pi
        psi = (psi + math.pi) % (2 * math.pi) - math.pi
        out[t] = [p[0], p[1], psi, speed]
    return out


# ----------------------------------------------------------------------------- scenario builders

LANE_W = 3.5


def build_polyline_map(polyline, name: str, handed: float = 1.0, max_seg: float = 6.0,
                       with_lights: bool = False, ring: Optional[dict] = None,
                       light_period: int = 300) -> MapData:
    """Two-lane road (ego lane centred on the waypoint polyline, oncoming lane beside it), junction
    arms at sharp turns, lane markings, optional stop lines + light schedule, optional ring."""
    poly = np.asarray(polyline, np.float64)
    ext = extend_polyline(poly, 25.0, 25.0)
    pts = resample_polyline(ext, max_seg)
    h = LANE_W / 2
    road = [ribbon(pts, -h, h), ribbon(pts, handed * h if handed > 0 else -3 * h,
                                       handed * 3 * h if handed > 0 else -h, reverse_dir=True)]
    marks = [strip_marking(pts, handed * h, 0.3, dash=3.0, gap=6.0),
             strip_marking(pts, -handed * h, 0.3), strip_marking(pts, handed * 3 * h, 0.3)]
    stop, sched_cols = [], []
    # junction arms where the route turns by more than 30 degrees
    for i in range(1, len(poly) - 1):
        d_in = poly[i] - poly[i - 1]; d_in /= np.linalg.norm(d_in)
        d_out = poly[i + 1] - poly[i]; d_out /= np.linalg.norm(d_out)
        turn = math.acos(float(np.clip(np.dot(d_in, d_out), -1, 1)))
        if turn > math.radians(30):
            arm = resample_polyline(np.stack([poly[i] - d_in * 2.0, poly[i] + d_in * 30.0]), max_seg)
            road += [ribbon(arm, -h, h), ribbon(arm, handed * h if handed > 0 else -3 * h,
                                               handed * 3 * h if handed > 0 else -h, reverse_dir=True)]
            marks.append(strip_marking(arm, handed * h, 0.3, dash=3.0, gap=6.0))
            if with_lights:
                c = poly[i] - d_in * 9.0
                stop.append([c[0], c[1], 0.8, LANE_W, math.atan2(d_in[1], d_in[0])])
                k = len(stop) - 1
                col = np.zeros(light_period, np.uint8)
                g, y = int(0.4 * light_period), int(0.1 * light_period)
                phase = (k * 97) % light_period
                for t in range(light_period):
                    u = (t + phase) % light_period
                    col[t] = LIGHT_GREEN if u < g else (LIGHT_YELLOW if u < g + y else LIGHT_RED)
                sched_cols.append(col)
    if ring is not None:
        road.append(ring_road(ring["center"], ring["r_in"], ring["r_out"], ring.get("n", 48)))
    road_t = np.concatenate(road).astype(np.float32)
    mark_t = np.concatenate(marks).astype(np.float32)
    stop_a = np.asarray(stop, np.float32).reshape(-1, 5)
    lights = (np.stack(sched_cols, axis=1) if sched_cols else np.zeros((1, 0), np.uint8))
    return MapData(road_tris=road_t, mark_tris=mark_t, stoplines=stop_a, light_states=lights, name=name)


def place_npcs(polyline, n_npcs: int, rng: np.random.Generator, handed: float = 1.0, T: int = 200,
               replay_fraction: float = 1.0, gap: float = 11.0):
    """NPCs on the ego lane (ahead of the start) and on the oncoming lane, with constant-speed
    path-following replays (the offline stand-in for IAI DRIVE, north_star)."""
    poly = extend_polyline(np.asarray(polyline, np.float64), 25.0, 25.0)
    ego_path = _Path(resample_polyline(poly, 2.0))
    opp_path = _Path(resample_polyline(offset_polyline(resample_polyline(poly, 2.0), handed * LANE_W), 2.0))
    states, attrs, replays, masks = [], [], [], []
    for k in range(n_npcs):
        oncoming = (k % 2 == 1)
        path = opp_path if oncoming else ego_path
        slot = k // 2
        s0 = (25.0 + 30.0 + slot * gap) if not oncoming else (path.total - 10.0 - slot * gap)
        s0 = float(np.clip(s0 + rng.uniform(-1.5, 1.5), 1.0, path.total - 1.0))
        speed = float(rng.uniform(5.0, 10.0))
        rep = rollout_replay(path, s0, speed, T, reverse=oncoming)
        length, width = float(rng.uniform(4.4, 5.2)), float(rng.uniform(1.8, 2.1))
        lr = float(rng.uniform(0.82, 0.97)) * length / 5.0 + 0.9
        states.append(rep[0]); attrs.append([length, width, lr]); replays.append(rep)
        masks.append(np.full(T, 1 if rng.uniform() < replay_fraction else 0, np.uint8))
    return (np.asarray(states, np.float32).reshape(-1, 4), np.asarray(attrs, np.float32).reshape(-1, 3),
            np.stack(replays, axis=1) if replays else np.zeros((T, 0, 4), np.float32),
            np.stack(masks, axis=1) if masks else np.zeros((T, 0), np.uint8))


def make_scenario(map_index: int, polyline, n_npcs: int, seed: int, name: str, handed: float = 1.0,
                  extra_npcs: Optional[dict] = None, T: int = 200, replay_fraction: float = 1.0) -> ScenarioData:
    rng = np.random.default_rng(seed)
    poly = np.asarray(polyline, np.float64)
    d = poly[1] - poly[0]
    heading = math.atan2(d[1], d[0])
    ego_attr = np.asarray([[5.0, 2.0, 0.9]], np.float32)
    ego_init = np.asarray([[poly[0][0], poly[0][1], heading, 0.0]], np.float32)
    st, at, rep, msk = place_npcs(poly, n_npcs, rng, handed=handed, T=T, replay_fraction=replay_fraction)
    if extra_npcs is not None:  # predetermined agents of the reference scenario: constant velocity, no replay
        es = np.asarray(extra_npcs["states"], np.float32).reshape(-1, 4)
        ea = np.asarray(extra_npcs["attributes"], np.float32).reshape(-1, 3)
        st = np.concatenate([es, st]); at = np.concatenate([ea, at])
        rep = np.concatenate([np.zeros((T, es.shape[0], 4), np.float32), rep], axis=1)
        msk = np.concatenate([np.zeros((T, es.shape[0]), np.uint8), msk], axis=1)
    init = np.concatenate([ego_init, st]).astype(np.float32)
    attr = np.concatenate([ego_attr, at]).astype(np.float32)
    n = init.shape[0]
    rs = np.zeros((T, n, 4), np.float32); rm = np.zeros((T, n), np.uint8)
    rs[:, 1:], rm[:, 1:] = rep, msk
    return ScenarioData(map_index=map_index, waypoints=poly.astype(np.float32), start_heading=float(heading),
                        agent_init=init, agent_attr=attr, replay_states=rs, replay_mask=rm, name=name)


def three_way(n_extra_npcs: int = 6, seed: int = 0) -> ScenarioSet:
    """BASELINE config C1: Three Way, ego + 2 predetermined + <=6 replay NPCs."""
    poly = VALIDATION_POLYLINES["three_way"]
    m = build_polyline_map(poly, "three_way")
    s = make_scenario(0, poly, n_extra_npcs, seed, "three_way", extra_npcs=THREE_WAY_NPCS)
    return ScenarioSet([m], [s])


def roundabout(n_agents: int = 16, n_variants: int = 1, seed: int = 0) -> ScenarioSet:
    """BASELINE config C2: Roundabout (waypoints circle the origin at r ~ 20 m)."""
    poly = VALIDATION_POLYLINES["roundabout"]
    m = build_polyline_map(poly, "roundabout", ring=dict(center=(0.0, 0.0), r_in=13.0, r_out=26.0, n=48))
    sc = [make_scenario(0, poly, n_agents - 1, seed + k, f"roundabout_{k}") for k in range(n_variants)]
    return ScenarioSet([m], sc)


def traffic_lights(n_agents: int = 32, n_variants: int = 1, seed: int = 0) -> ScenarioSet:
    """BASELINE config C3: Traffic Lights (11 waypoints, 228 m) with stop lines + light schedule."""
    poly = VALIDATION_POLYLINES["traffic_lights"]
    m = build_polyline_map(poly, "traffic_lights", with_lights=True)
    sc = [make_scenario(0, poly, n_agents - 1, seed + k, f"traffic_lights_{k}") for k in range(n_variants)]
    return ScenarioSet([m], sc)


def validation_mix(n_agents: int = 8, seed: int = 0) -> ScenarioSet:
    """All five validation polylines, one map each (a small stand-in for config C5's scenario mix)."""
    maps, scen = [], []
    for k, (name, poly) in enumerate(VALIDATION_POLYLINES.items()):
        maps.append(build_polyline_map(poly, name, with_lights=(name == "traffic_lights"),
                                       ring=dict(center=(0.0, 0.0), r_in=13.0, r_out=26.0) if name == "roundabout" else None))
        scen.append(make_scenario(k, poly, n_agents - 1, seed + k, name,
                                  extra_npcs=THREE_WAY_NPCS if name == "three_way" else None))
    return ScenarioSet(maps, scen)


def synthetic_training_polyline(rng: np.random.Generator) -> np.ndarray:
    """One waypoint polyline with the statistics of the reference's training suite
    (torchdriveenv/data/training_cases.yml:102-2980; SURVEY.md §8a row a10: 5-20 waypoints, mean 14.4,
    spacing 12.9-15.0 m): a heading random walk with an occasional junction-sized turn."""
    n = int(rng.integers(5, 21))
    pts = [rng.uniform(-200.0, 200.0, 2)]
    psi = float(rng.uniform(-math.pi, math.pi))
    for _ in range(n - 1):
        psi += float(rng.normal(0.0, 0.12)) + (float(rng.choice([-1.0, 1.0])) * float(rng.uniform(0.7, 1.4)) if rng.uniform() < 0.15 else 0.0)
        step = float(rng.uniform(12.9, 15.0))
        pts.append(pts[-1] + step * np.array([math.cos(psi), math.sin(psi)]))
    return np.round(np.asarray(pts, np.float64), 3)


def reference_training_polylines() -> List[np.ndarray]:
    """The 100 waypoint polylines of the reference's training suite (torchdriveenv/data/training_cases.yml:102-2980),
    from the packaged copy torchdriveenv_b200/data/training_cases.json (tools/make_packaged_suites.py)."""
    import json
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "training_cases.json")
    with open(path) as f:
        suite = json.load(f)
    return [np.asarray(p, np.float64) for p in suite["waypoint_suite"]]


def training_mix(n_scenarios: int = 100, n_agents: int = 8, seed: int = 0) -> ScenarioSet:
    """BASELINE config C5's scenario mix: the reference's own training polylines (the first ``n_scenarios`` of the 100 in
    training_cases.yml; beyond 100 synthetic ones with the same statistics follow), one map each (two-lane road, junction
    arms with stop lines + lights at sharp turns), ego + constant-speed log-replay NPCs (the offline stand-in for the
    Inverted AI agents)."""
    rng = np.random.default_rng(seed)
    polys = reference_training_polylines()
    maps, scen = [], []
    for k in range(n_scenarios):
        poly = polys[k] if k < len(polys) else synthetic_training_polyline(rng)
        maps.append(build_polyline_map(poly, f"train_{k}", with_lights=True))
        scen.append(make_scenario(k, poly, n_agents - 1, seed + 1000 + k, f"train_{k}"))
    return ScenarioSet(maps, scen)


def scatter_patch(size: float = 200.0, cell: float = 10.0) -> MapData:
    """Config C4's lane mesh: a checkerboard of road squares over a size x size patch, so that a
    uniformly scattered box is on/near/off the road with comparable probability."""
    n = int(size / cell)
    tris = []
    for i in range(n):
        for j in range(n):
            if (i + j) % 2:
                continue
            x0, y0 = i * cell, j * cell
            x1, y1 = x0 + cell, y0 + cell
            tris.append([x0, y0, x1, y0, x1, y1, 1.0, 0.0])
            tris.append([x0, y0, x1, y1, x0, y1, 1.0, 0.0])
    return MapData(road_tris=np.asarray(tris, np.float32), name="scatter_patch")


def scatter_boxes(num_envs: int, num_agents: int, size: float = 200.0, seed: int = 0, present_p: float = 1.0):
    """Config C4's synthetic boxes: state (E, A, 4), attr (E, A, 4)."""
    rng = np.random.default_rng(seed)
    st = np.zeros((num_envs, num_agents, 4), np.float32)
    at = np.zeros((num_envs, num_agents, 4), np.float32)
    st[..., 0] = rng.uniform(0, size, (num_envs, num_agents))
    st[..., 1] = rng.uniform(0, size, (num_envs, num_agents))
    st[..., 2] = rng.uniform(-math.pi, math.pi, (num_envs, num_agents))
    st[..., 3] = rng.uniform(0, 10, (num_envs, num_agents))
    at[..., 0] = rng.uniform(4.8, 5.5, (num_envs, num_agents))
    at[..., 1] = rng.uniform(1.8, 2.2, (num_envs, num_agents))
    at[..., 2] = rng.uniform(0.82, 0.97, (num_envs, num_agents))
    at[..., 3] = (rng.uniform(0, 1, (num_envs, num_agents)) < present_p).astype(np.float32)
    return st, at
