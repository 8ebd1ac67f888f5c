"""Independent numpy restatements of the birdview rule, used to pin the oracle's rasteriser.
Primitive list and painter's levels as in DESIGN.md §SPEC-render."""
import math

import cv2
import numpy as np

H = W = 64
CLS = dict(road=1, mark=2, green=3, yellow=4, red=5, waypoint=6, vehicle=7, ego=8, direction=9, ego_direction=10)


def world_primitives(packed, cfg, state, attr, env_vars, sincos):
    """[(class, (n,2) float32 world vertices)] for one env, in painter's order."""
    f = np.float32
    s, step, target, _, phase, _, m = [int(v) for v in env_vars[:7]]
    prims = []
    t0, t1 = packed["map_tri_offset"][m], packed["map_tri_offset"][m + 1]
    for t in packed["road_tris"][t0:t1]:
        prims.append((CLS["road"], t[:6].reshape(3, 2)))
    k0, k1 = packed["map_mark_offset"][m], packed["map_mark_offset"][m + 1]
    for t in packed["mark_tris"][k0:k1]:
        prims.append((CLS["mark"], t.reshape(3, 2)))

    def box(x, y, psi, length, width):
        sn, cs = sincos(np.array([psi], f)); sn, cs = f(sn[0]), f(cs[0])
        hl, hw = f(0.5) * f(length), f(0.5) * f(width)
        def pt(ox, oy):
            return [f(x) + (f(ox) * cs - f(oy) * sn), f(y) + (f(ox) * sn + f(oy) * cs)]
        quad = np.array([pt(hl, hw), pt(hl, -hw), pt(-hl, -hw), pt(-hl, hw)], f)
        tri = np.array([pt(hl, f(0)), pt(f(0.5) * hl, hw), pt(f(0.5) * hl, -hw)], f)
        return quad, tri

    l0, l1 = packed["map_stop_offset"][m], packed["map_stop_offset"][m + 1]
    L = l1 - l0
    P = int(packed["map_light_period"][m])
    for l in range(L):
        sl = packed["stoplines"][l0 + l]
        ls = int(packed["light_states"][packed["map_light_offset"][m] + ((step + phase) % P) * L + l]) if P > 0 else 0
        quad, _ = box(sl[0], sl[1], sl[4], sl[2], sl[3])
        prims.append(({0: CLS["green"], 1: CLS["yellow"], 2: CLS["red"]}[ls], quad))
    w0, w1 = packed["scen_wp_offset"][s], packed["scen_wp_offset"][s + 1]
    if target < w1 - w0:
        wp = packed["waypoints"][w0 + target]
        r = f(2.0)
        prims.append((CLS["waypoint"], np.array([[wp[0] + r, wp[1]], [wp[0], wp[1] + r], [wp[0] - r, wp[1]], [wp[0], wp[1] - r]], f)))
    quads, tris = [], []
    for a in range(state.shape[0]):
        if attr[a, 3] == 0:
            continue
        quad, tri = box(state[a, 0], state[a, 1], state[a, 2], attr[a, 0], attr[a, 1])
        prims.append((CLS["ego"] if a == 0 else CLS["vehicle"], quad))
        prims.append((CLS["ego_direction"] if a == 0 else CLS["direction"], tri))
    return prims


def camera(cfg, ego_state, sincos):
    f = np.float32
    sn, cs = sincos(np.array([ego_state[2]], f))
    ppm = f(W) / f(cfg.fov)
    return dict(ex=f(ego_state[0]), ey=f(ego_state[1]), ce=f(cs[0]), se=f(sn[0]), ppm=ppm,
                ppmy=ppm if cfg.left_handed_coordinates else -ppm)


def to_pixels32(cam, verts):
    f = np.float32
    v = np.asarray(verts, f)
    dx, dy = v[:, 0] - cam["ex"], v[:, 1] - cam["ey"]
    cx = dx * cam["ce"] + dy * cam["se"]
    cy = dy * cam["ce"] - dx * cam["se"]
    return cx * cam["ppm"] + f(0.5 * W), cy * cam["ppmy"] + f(0.5 * H)


def raster_fixed(prims, cam):
    """Same rule as the oracle, written independently: 1/16-px snap, integer edge functions at pixel
    centres, top-left ties, final class = highest class covering the pixel."""
    img = np.zeros((H, W), np.uint8)
    jj, ii = np.meshgrid(np.arange(H, dtype=np.int64), np.arange(W, dtype=np.int64), indexing="ij")
    px, py = 16 * ii + 8, 16 * jj + 8
    for cls, verts in prims:
        fx, fy = to_pixels32(cam, verts)
        if not (fx.max() >= -1 and fx.min() <= W + 1 and fy.max() >= -1 and fy.min() <= H + 1):
            continue
        X = np.clip(np.rint(fx * np.float32(16)), -8191, 8191).astype(np.int64)
        Y = np.clip(np.rint(fy * np.float32(16)), -8191, 8191).astype(np.int64)
        n = len(X)
        area2 = sum(X[k] * Y[(k + 1) % n] - X[(k + 1) % n] * Y[k] for k in range(n))
        if area2 == 0:
            continue
        if area2 < 0:
            X, Y = X[::-1], Y[::-1]
        inside = np.ones((H, W), bool)
        for k in range(n):
            dx, dy = X[(k + 1) % n] - X[k], Y[(k + 1) % n] - Y[k]
            if dx == 0 and dy == 0:
                continue
            E = dx * (py - Y[k]) - dy * (px - X[k])
            incl = (dy < 0) or (dy == 0 and dx > 0)
            inside &= (E > 0) | ((E == 0) & incl)
        img[inside & (img < cls)] = cls
    return img


def raster_float64(prims, cam):
    """No snapping, float64, closed edges: differs from the fixed-point rule only on primitive edges."""
    img = np.zeros((H, W), np.uint8)
    jj, ii = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    px, py = ii + 0.5, jj + 0.5
    c64 = {k: float(v) for k, v in cam.items()}
    for cls, verts in prims:
        v = np.asarray(verts, np.float64)
        dx, dy = v[:, 0] - c64["ex"], v[:, 1] - c64["ey"]
        fx = (dx * c64["ce"] + dy * c64["se"]) * c64["ppm"] + W / 2
        fy = (dy * c64["ce"] - dx * c64["se"]) * c64["ppmy"] + H / 2
        n = len(fx)
        area2 = sum(fx[k] * fy[(k + 1) % n] - fx[(k + 1) % n] * fy[k] for k in range(n))
        if area2 == 0:
            continue
        sgn = 1.0 if area2 > 0 else -1.0
        inside = np.ones((H, W), bool)
        for k in range(n):
            ex, ey = fx[(k + 1) % n] - fx[k], fy[(k + 1) % n] - fy[k]
            inside &= sgn * (ex * (py - fy[k]) - ey * (px - fx[k])) >= 0
        img[inside & (img < cls)] = cls
    return img


def raster_cv2(prims, cam):
    """cv2.fillPoly with 4 fractional bits, painter's order by class (paints boundary pixels too)."""
    img = np.zeros((H, W), np.uint8)
    for cls in sorted({c for c, _ in prims}):
        for c, verts in prims:
            if c != cls:
                continue
            fx, fy = to_pixels32(cam, verts)
            if not (fx.max() >= -1 and fx.min() <= W + 1 and fy.max() >= -1 and fy.min() <= H + 1):
                continue
            # cv2 pixel centres are at integer coordinates: shift by half a pixel
            pts = np.stack([np.rint((fx - 0.5) * 16), np.rint((fy - 0.5) * 16)], 1).astype(np.int32)
            cv2.fillPoly(img, [pts.reshape(-1, 1, 2)], int(cls), lineType=cv2.LINE_8, shift=4)
    return img
