// tde_kernels.cuh — the fused per-timestep kernel and its helpers (sm_100a).
//
// One warp owns one environment for the whole step: lanes are agents during the physics phases
// (bicycle step / NPC replay, all-pairs SAT, corner-to-lane-mesh offroad, red-light stop lines,
// wrong-way, reward + termination + bookkeeping) and primitives / image rows during the birdview
// phase.  Nothing is exchanged between warps, so there is no block-level synchronisation on the
// step path except the one-time staging of the lane mesh into shared memory.
#pragma once
#include "tde_device.cuh"
#include "../../include/tde_b200.h"

#define TDE_WARPS_PER_BLOCK 8
#define TDE_SPAN_STRIDE 33

struct MapDev {
    const float4* tri;          // 3 float4 per road triangle (see tde_point_tri_dist2)
    const float2* mark;         // 3 float2 per lane-marking triangle
    const float4* stop;         // 2 float4 per stop line: [x y hl hw] [c s 0 0]
    const uint8_t* lights;      // [period][nstop]
    const int* cell_start;      // [gnx*gny+1]
    const uint16_t* cell_items; // nearest-candidate triangle ids per grid cell
    int ntri, nmark, nstop, period;
    float gx0, gy0, inv_cell;
    int gnx, gny;
};

struct ScenDev {
    const float2* wp;
    const float4* init;        // [A] x y psi v
    const float4* attr;        // [A] length width lr 0
    const float4* rep_states;  // [T][A]
    const uint8_t* rep_mask;   // [T][A]
    int map, W, nag, rep_T;
    float start_heading;
    int pad[3];
};

struct StepParams {
    tde_config cfg;
    int E, A, phases, num_scen;
    unsigned long long seed;
    const MapDev* maps;
    const ScenDev* scens;
    float4* state;
    float4* attr;
    float4* infr;
    int* vars;
    float* ep_return;
    const int* scen_lo;
    const int* scen_hi;
    const float* actions;
    uint8_t* obs;
    float* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    float* info;
    double* stats;
    const uint8_t* reset_mask;
    uint32_t pal[3][4];  // per channel: 16 class bytes
    float ppm, ppmy;
};

struct Cam { float ex, ey, ce, se, ppm, ppmy; };

struct WarpScratch {
    float4 box[TDE_MAX_AGENTS * 2];                  // Box as two float4
    unsigned long long span[32 * TDE_SPAN_STRIDE];   // one 32-row band, [row][lane]
    unsigned long long planes[64 * 4];               // [row][bit-plane] class-index image
};

__device__ __forceinline__ Box ld_box(const float4* sb, int a) {
    float4 u = sb[2 * a], v = sb[2 * a + 1];
    Box b; b.x = u.x; b.y = u.y; b.hl = u.z; b.hw = u.w; b.c = v.x; b.s = v.y; b.present = v.z; b.r = v.w;
    return b;
}
__device__ __forceinline__ void st_box(float4* sb, int a, const Box& b) {
    sb[2 * a] = make_float4(b.x, b.y, b.hl, b.hw);
    sb[2 * a + 1] = make_float4(b.c, b.s, b.present, b.r);
}

// ---------------------------------------------------------------- lane-mesh queries

// squared distance from p to the road mesh of map M (0 on the road); exact: the grid cell lists hold
// every triangle that can be nearest for any point of the cell, points off the grid scan all triangles
__device__ __forceinline__ float mesh_dist2(const MapDev& M, float px, float py) {
    float best = INFINITY;
    float fx = floorf((px - M.gx0) * M.inv_cell), fy = floorf((py - M.gy0) * M.inv_cell);
    bool in_grid = fx >= 0.0f && fy >= 0.0f && fx < (float)M.gnx && fy < (float)M.gny;
    bool ins; float dc, ds;
    if (in_grid) {
        int cell = (int)fy * M.gnx + (int)fx;
        int i0 = M.cell_start[cell], i1 = M.cell_start[cell + 1];
        for (int i = i0; i < i1; ++i) {
            float d2 = tde_point_tri_dist2(M.tri + 3 * (int)M.cell_items[i], px, py, ins, dc, ds);
            best = fminf(best, d2);
        }
    } else {
        for (int t = 0; t < M.ntri; ++t) {
            float d2 = tde_point_tri_dist2(M.tri + 3 * t, px, py, ins, dc, ds);
            best = fminf(best, d2);
        }
    }
    return best;
}

// compute_offroad (gym_env.py:142,415,427): sum over corners of max(dist - threshold, 0)
__device__ __forceinline__ float offroad_box(const MapDev& M, const Box& b, float thr) {
    if (M.ntri <= 0) return 0.0f;
    float sum = 0.0f;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        float px, py;
        tde_box_corner(b, k, px, py);
        float d = sqrtf(mesh_dist2(M, px, py));
        sum = sum + fmaxf(d - thr, 0.0f);
    }
    return sum;
}

// compute_wrong_way: max(-cos(psi - lane_dir), 0), min over the triangles under the centre
__device__ __forceinline__ float wrong_way_box(const MapDev& M, const Box& b) {
    float best = INFINITY;
    float fx = floorf((b.x - M.gx0) * M.inv_cell), fy = floorf((b.y - M.gy0) * M.inv_cell);
    bool in_grid = fx >= 0.0f && fy >= 0.0f && fx < (float)M.gnx && fy < (float)M.gny;
    bool ins; float dc, ds;
    if (in_grid) {
        int cell = (int)fy * M.gnx + (int)fx;
        int i0 = M.cell_start[cell], i1 = M.cell_start[cell + 1];
        for (int i = i0; i < i1; ++i) {
            (void)tde_point_tri_dist2(M.tri + 3 * (int)M.cell_items[i], b.x, b.y, ins, dc, ds);
            if (ins) best = fminf(best, fmaxf(-(b.c * dc + b.s * ds), 0.0f));
        }
    } else {
        for (int t = 0; t < M.ntri; ++t) {
            (void)tde_point_tri_dist2(M.tri + 3 * t, b.x, b.y, ins, dc, ds);
            if (ins) best = fminf(best, fmaxf(-(b.c * dc + b.s * ds), 0.0f));
        }
    }
    return best == INFINITY ? 0.0f : best;
}

__device__ __forceinline__ int light_state_at(const MapDev& M, int step, int phase, int l) {
    if (M.period <= 0 || M.nstop <= 0) return TDE_LIGHT_GREEN;
    int t = (step + phase) % M.period;
    return M.lights[t * M.nstop + l];
}

// TrafficLightControl.compute_violation: rear strip of the agent box vs red stop lines
__device__ __forceinline__ float tl_violation_box(const MapDev& M, const Box& b, float rear_factor, int step, int phase) {
    float length = 2.0f * b.hl;  // exact: hl = 0.5f*length
    float len2 = length * rear_factor;
    float back = 0.5f * (length - len2);
    Box rear = b;
    rear.x = b.x - back * b.c;
    rear.y = b.y - back * b.s;
    rear.hl = 0.5f * len2;
    float viol = 0.0f;
    for (int l = 0; l < M.nstop; ++l) {
        if (light_state_at(M, step, phase, l) != TDE_LIGHT_RED) continue;
        float4 u = M.stop[2 * l], v = M.stop[2 * l + 1];
        Box sb; sb.x = u.x; sb.y = u.y; sb.hl = u.z; sb.hw = u.w; sb.c = v.x; sb.s = v.y; sb.present = 1.0f; sb.r = 0.0f;
        if (tde_overlap(rear, sb)) viol += 1.0f;
    }
    return viol;
}

// ---------------------------------------------------------------- reset (warp-cooperative)

// WaypointSuiteEnv.reset :319-349 / set_start_pos :351-367 / build_simulator init :192-198,241-247,275-283.
// Writes state, attributes, env vars and zeroed infractions of env e; returns the new uniform vars.
template <int AH>
__device__ __forceinline__ void reset_env_warp(const StepParams& p, int e, int lane, int& s, int& step, int& target,
                                               int& reached, int& lphase, int& episode, int& m, float4 (&st)[AH],
                                               float4 (&at)[AH]) {
    const tde_config& c = p.cfg;
    unsigned long long genv = (unsigned long long)(c.env_index_offset + e);
    unsigned long long ep = (unsigned long long)(unsigned int)episode;
    int lo = p.scen_lo[e], hi = p.scen_hi[e];
    int span = hi - lo; if (span < 1) span = 1;
    s = lo + (int)((tde_rng(p.seed, genv, ep, 0) >> 32) % (unsigned long long)span);
    const ScenDev& S = p.scens[s];
    m = S.map;
    float u_pos = tde_u01(tde_rng(p.seed, genv, ep, 1));
    float u_spd = tde_u01(tde_rng(p.seed, genv, ep, 2));
    float z = tde_normal8(tde_rng(p.seed, genv, ep, 3), tde_rng(p.seed, genv, ep, 4));
    float2 p0 = S.wp[0];
    float2 p1 = S.W > 1 ? S.wp[1] : p0;
#pragma unroll
    for (int h = 0; h < AH; ++h) {
        int a = h * 32 + lane;
        if (a < p.A) {
            float4 init = S.init[a];
            if (a >= 1 && S.rep_T > 0 && S.rep_mask[a]) init = S.rep_states[a];
            float4 attr = S.attr[a];
            attr.w = a < S.nag ? 1.0f : 0.0f;
            if (a == 0) {
                init.x = p0.x + u_pos * (p1.x - p0.x);
                init.y = p0.y + u_pos * (p1.y - p0.y);
                init.z = S.start_heading + c.start_heading_sigma * z;
                init.w = u_spd * c.start_speed_max;
                if (c.randomize_ego_attributes) {
                    attr.x = 4.8f + tde_u01(tde_rng(p.seed, genv, ep, 6)) * (5.5f - 4.8f);
                    attr.y = 1.8f + tde_u01(tde_rng(p.seed, genv, ep, 7)) * (2.2f - 1.8f);
                    attr.z = 0.82f + tde_u01(tde_rng(p.seed, genv, ep, 8)) * (0.97f - 0.82f);
                }
            }
            st[h] = init; at[h] = attr;
            p.state[(size_t)e * p.A + a] = init;
            p.attr[(size_t)e * p.A + a] = attr;
            p.infr[(size_t)e * p.A + a] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    int P = p.maps[m].period;
    step = 0; target = 1; reached = 0;
    lphase = (int)((tde_rng(p.seed, genv, ep, 5) >> 32) % (unsigned long long)(P > 0 ? P : 1));
    episode = (int)((unsigned int)episode + 1u);
    if (lane == 0) p.ep_return[e] = 0.0f;
}

__device__ __forceinline__ void store_vars(const StepParams& p, int e, int lane, int s, int step, int target,
                                           int reached, int lphase, int episode, int m) {
    int v = lane == 0 ? s : lane == 1 ? step : lane == 2 ? target : lane == 3 ? reached : lane == 4 ? lphase
          : lane == 5 ? episode : lane == 6 ? m : 0;
    if (lane < 8) p.vars[(size_t)e * 8 + lane] = v;
}

// ---------------------------------------------------------------- birdview rasteriser

struct LanePrim {
    int x[4], y[4];
    int n;    // 0 = nothing to draw
    int cls;
};

__device__ __forceinline__ int snap16(float f) {
    float r = rintf(f * 16.0f);
    r = fminf(fmaxf(r, -8191.0f), 8191.0f);
    return (int)r;
}

// world -> pixel (translate to ego, rotate by -psi, scale) + viewport test + snap to 1/16 px
template <int N>
__device__ __forceinline__ LanePrim make_prim(const Cam& cam, const float (&wx)[4], const float (&wy)[4], int cls, bool valid) {
    LanePrim pr;
    float minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < N) {
            float dx = wx[k] - cam.ex, dy = wy[k] - cam.ey;
            float cx = dx * cam.ce + dy * cam.se;
            float cy = dy * cam.ce - dx * cam.se;
            float fx = cx * cam.ppm + 0.5f * (float)TDE_OBS_W;
            float fy = cy * cam.ppmy + 0.5f * (float)TDE_OBS_H;
            minx = fminf(minx, fx); maxx = fmaxf(maxx, fx);
            miny = fminf(miny, fy); maxy = fmaxf(maxy, fy);
            pr.x[k] = snap16(fx); pr.y[k] = snap16(fy);
        } else {
            pr.x[k] = 0; pr.y[k] = 0;
        }
    }
    bool vis = valid && (maxx >= -1.0f && minx <= (float)TDE_OBS_W + 1.0f && maxy >= -1.0f && miny <= (float)TDE_OBS_H + 1.0f);
    pr.n = vis ? N : 0;
    pr.cls = cls;
    return pr;
}

// Rasterise up to 32 convex primitives (one per lane) into the warp's class-index bit-planes.
//   phase 1 (lane = primitive): exact integer edge stepping, one 64-bit coverage span per row,
//            written to the shared band buffer;
//   phase 2 (lane = row): OR the spans of each class present (ascending = painter's order) and
//            update the four bit-planes of the class index.
// Pixel-centre sampling on the 1/16-px grid with the top-left rule: identical to the oracle's
// per-pixel edge-function test.
__device__ __forceinline__ void raster_chunk(const LanePrim& pr, unsigned long long (&P)[2][4],
                                             unsigned long long* span, int lane) {
    // orientation
    int n = pr.n;
    int X[4], Y[4];
    {
        int area2 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int k1 = (k + 1 == n) ? 0 : k + 1;
            if (k < n) area2 += pr.x[k] * pr.y[k1] - pr.x[k1] * pr.y[k];
        }
        if (area2 == 0) n = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            int src = area2 > 0 ? k : (n - 1 - k);
            src = src < 0 ? 0 : src;
            // reversed order for negative area; entries k >= n are unused
            int xs = pr.x[0], ys = pr.y[0];
#pragma unroll
            for (int q = 1; q < 4; ++q) { if (src == q) { xs = pr.x[q]; ys = pr.y[q]; } }
            X[k] = xs; Y[k] = ys;
        }
    }
    int ymin = Y[0], ymax = Y[0], xmin = X[0], xmax = X[0];
#pragma unroll
    for (int k = 1; k < 4; ++k) {
        if (k < n) { ymin = min(ymin, Y[k]); ymax = max(ymax, Y[k]); xmin = min(xmin, X[k]); xmax = max(xmax, X[k]); }
    }
    int j0 = max(0, (ymin - 8 + 15) >> 4);        // ceil((ymin-8)/16)
    int j1 = min(TDE_OBS_H - 1, (ymax - 8) >> 4); // floor((ymax-8)/16)
    bool active = n >= 3 && j0 <= j1 && xmax >= 8 && xmin <= 16 * (TDE_OBS_W - 1) + 8;
    if (!__any_sync(FULL_MASK, active)) return;

    // edge setup: type 0 none, 1 left (dy<0), 2 right (dy>0), 3 horizontal
    int etype[4], F[4], rem[4], qS[4], rS[4], D[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        etype[k] = 0; F[k] = 0; rem[k] = 0; qS[k] = 0; rS[k] = 0; D[k] = 1;
        if (active && k < n) {
            int k1 = (k + 1 == n) ? 0 : k + 1;
            int ax = X[k], ay = Y[k];
            int bx = X[0], by = Y[0];
#pragma unroll
            for (int q = 1; q < 4; ++q) { if (k1 == q) { bx = X[q]; by = Y[q]; } }
            int dx = bx - ax, dy = by - ay;
            if (dx != 0 || dy != 0) {
                int bias = (dy < 0 || (dy == 0 && dx > 0)) ? 0 : -1;
                int C0 = dx * (8 - ay) - dy * (8 - ax) + bias;
                int S = 16 * dx;
                int K = C0 + S * j0;
                if (dy == 0) {
                    etype[k] = 3; F[k] = K; qS[k] = S;
                } else {
                    int d = dy > 0 ? 16 * dy : -16 * dy;
                    etype[k] = dy < 0 ? 1 : 2;
                    D[k] = d;
                    F[k] = tde_floordiv(K, d);
                    rem[k] = K - F[k] * d;
                    qS[k] = tde_floordiv(S, d);
                    rS[k] = S - qS[k] * d;
                }
            }
        }
    }

    int j = j0;
#pragma unroll 1
    for (int band = 0; band < 2; ++band) {
        int b0 = band * 32, b1 = b0 + 31;
        bool part = active && j0 <= b1 && j1 >= b0;
        unsigned bm = __ballot_sync(FULL_MASK, part);
        if (bm == 0) continue;
        // clear the band buffer
#pragma unroll 1
        for (int i = lane; i < 32 * TDE_SPAN_STRIDE; i += 32) span[i] = 0ull;
        __syncwarp();
        if (part) {
            int jend = min(j1, b1);
#pragma unroll 1
            for (; j <= jend; ++j) {
                int xl = 0, xr = TDE_OBS_W;
                bool ok = true;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (etype[k] == 1) xl = max(xl, -F[k]);
                    else if (etype[k] == 2) xr = min(xr, F[k] + 1);
                    else if (etype[k] == 3) ok = ok && (F[k] >= 0);
                }
                if (ok && xl < xr) {
                    unsigned long long mk = (~0ull >> (64 - xr)) & (~0ull << xl);
                    span[(j - b0) * TDE_SPAN_STRIDE + lane] = mk;
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (etype[k] == 3) F[k] += qS[k];
                    else if (etype[k] != 0) {
                        F[k] += qS[k];
                        rem[k] += rS[k];
                        if (rem[k] >= D[k]) { rem[k] -= D[k]; F[k] += 1; }
                    }
                }
            }
        }
        __syncwarp();
        // phase 2: lane = row (b0 + lane); classes ascending
        unsigned remaining = bm;
        while (remaining) {
            int mycls = ((remaining >> lane) & 1u) ? pr.cls : 0x7fffffff;
            int c = __reduce_min_sync(FULL_MASK, mycls);
            unsigned mc = __ballot_sync(FULL_MASK, mycls == c);
            unsigned long long acc = 0ull;
            unsigned it = mc;
            while (it) {
                int k = __ffs(it) - 1;
                it &= it - 1;
                acc |= span[lane * TDE_SPAN_STRIDE + k];
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                unsigned long long keep = ((c >> b) & 1) ? ~0ull : 0ull;
                P[band][b] = (P[band][b] & ~acc) | (acc & keep);
            }
            remaining &= ~mc;
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void box_quad(const Box& b, float (&wx)[4], float (&wy)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) tde_box_corner(b, k, wx[k], wy[k]);
}
__device__ __forceinline__ void box_dirtri(const Box& b, float (&wx)[4], float (&wy)[4]) {
    float ox[3] = {b.hl, 0.5f * b.hl, 0.5f * b.hl};
    float oy[3] = {0.0f, b.hw, -b.hw};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        wx[k] = b.x + (ox[k] * b.c - oy[k] * b.s);
        wy[k] = b.y + (ox[k] * b.s + oy[k] * b.c);
    }
    wx[3] = 0.f; wy[3] = 0.f;
}

// spread the 8 bits of b to the low bit of 8 nibbles
__device__ __forceinline__ uint32_t spread8(uint32_t b) {
    uint32_t x = b & 0xffu;
    x = (x | (x << 12)) & 0x000f000fu;
    x = (x | (x << 6)) & 0x03030303u;
    x = (x | (x << 3)) & 0x11111111u;
    return x;
}

// simulator.render_egocentric() (gym_env.py:122-124) for one env, by one warp.
template <int AH>
__device__ __forceinline__ void render_env_warp(const StepParams& p, const MapDev& M, const ScenDev& S, int e, int lane,
                                                WarpScratch* ws, int step, int lphase, int target) {
    const float4* sb = ws->box;
    Box ego = ld_box(sb, 0);
    Cam cam;
    cam.ex = ego.x; cam.ey = ego.y; cam.ce = ego.c; cam.se = ego.s; cam.ppm = p.ppm; cam.ppmy = p.ppmy;
    unsigned long long P[2][4];
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int q = 0; q < 4; ++q) P[b][q] = 0ull;
    float wx[4], wy[4];

    // conservative world-space reach of the viewport around the ego (half diagonal + 1 px)
    float reach = (0.70710678f * (float)(TDE_OBS_W + TDE_OBS_H) * 0.5f + 2.0f) / p.ppm;

    // level 1: road triangles
#pragma unroll 1
    for (int base = 0; base < M.ntri; base += 32) {
        int t = base + lane;
        bool valid = t < M.ntri;
        float4 t0 = make_float4(0.f, 0.f, 0.f, 0.f), t1 = t0;
        if (valid) { t0 = M.tri[3 * t]; t1 = M.tri[3 * t + 1]; }
        // cheap world-space reject (conservative): triangle bbox vs the viewport's bounding square
        float lox = fminf(fminf(t0.x, t0.z), t1.x), hix = fmaxf(fmaxf(t0.x, t0.z), t1.x);
        float loy = fminf(fminf(t0.y, t0.w), t1.y), hiy = fmaxf(fmaxf(t0.y, t0.w), t1.y);
        valid = valid && hix >= cam.ex - reach && lox <= cam.ex + reach && hiy >= cam.ey - reach && loy <= cam.ey + reach;
        if (!__any_sync(FULL_MASK, valid)) continue;
        wx[0] = t0.x; wy[0] = t0.y; wx[1] = t0.z; wy[1] = t0.w; wx[2] = t1.x; wy[2] = t1.y; wx[3] = 0.f; wy[3] = 0.f;
        LanePrim pr = make_prim<3>(cam, wx, wy, TDE_CLS_ROAD, valid);
        raster_chunk(pr, P, ws->span, lane);
    }
    // level 2: lane markings
#pragma unroll 1
    for (int base = 0; base < M.nmark; base += 32) {
        int t = base + lane;
        bool valid = t < M.nmark;
        float2 a = make_float2(0.f, 0.f), b = a, c = a;
        if (valid) { a = M.mark[3 * t]; b = M.mark[3 * t + 1]; c = M.mark[3 * t + 2]; }
        float lox = fminf(fminf(a.x, b.x), c.x), hix = fmaxf(fmaxf(a.x, b.x), c.x);
        float loy = fminf(fminf(a.y, b.y), c.y), hiy = fmaxf(fmaxf(a.y, b.y), c.y);
        valid = valid && hix >= cam.ex - reach && lox <= cam.ex + reach && hiy >= cam.ey - reach && loy <= cam.ey + reach;
        if (!__any_sync(FULL_MASK, valid)) continue;
        wx[0] = a.x; wy[0] = a.y; wx[1] = b.x; wy[1] = b.y; wx[2] = c.x; wy[2] = c.y; wx[3] = 0.f; wy[3] = 0.f;
        LanePrim pr = make_prim<3>(cam, wx, wy, TDE_CLS_LANE_MARKING, valid);
        raster_chunk(pr, P, ws->span, lane);
    }
    // levels 3-6: stop lines coloured by light state, then the goal waypoint
    {
        int nq = M.nstop + 1;
#pragma unroll 1
        for (int base = 0; base < nq; base += 32) {
            int q = base + lane;
            bool valid = false;
            int cls = TDE_CLS_TL_GREEN;
            wx[0] = wx[1] = wx[2] = wx[3] = 0.f; wy[0] = wy[1] = wy[2] = wy[3] = 0.f;
            if (q < M.nstop) {
                float4 u = M.stop[2 * q], v = M.stop[2 * q + 1];
                Box b; b.x = u.x; b.y = u.y; b.hl = u.z; b.hw = u.w; b.c = v.x; b.s = v.y; b.present = 1.f; b.r = 0.f;
                box_quad(b, wx, wy);
                int ls = light_state_at(M, step, lphase, q);
                cls = ls == TDE_LIGHT_RED ? TDE_CLS_TL_RED : (ls == TDE_LIGHT_YELLOW ? TDE_CLS_TL_YELLOW : TDE_CLS_TL_GREEN);
                valid = true;
            } else if (q == M.nstop && target < S.W) {
                float2 w = S.wp[target];
                float r = 2.0f;
                wx[0] = w.x + r; wy[0] = w.y; wx[1] = w.x; wy[1] = w.y + r;
                wx[2] = w.x - r; wy[2] = w.y; wx[3] = w.x; wy[3] = w.y - r;
                cls = TDE_CLS_WAYPOINT;
                valid = true;
            }
            if (!__any_sync(FULL_MASK, valid)) continue;
            LanePrim pr = make_prim<4>(cam, wx, wy, cls, valid);
            raster_chunk(pr, P, ws->span, lane);
        }
    }
    // levels 7-8: agent rectangles (ego highlighted), levels 9-10: direction triangles.
    // The chunk holding the ego goes last so that classes stay ascending across chunks.
#pragma unroll 1
    for (int h = AH - 1; h >= 0; --h) {
        int a = h * 32 + lane;
        Box b = ld_box(sb, a < p.A ? a : 0);
        bool valid = a < p.A && b.present != 0.0f;
        box_quad(b, wx, wy);
        LanePrim pr = make_prim<4>(cam, wx, wy, a == 0 ? TDE_CLS_EGO : TDE_CLS_VEHICLE, valid);
        raster_chunk(pr, P, ws->span, lane);
    }
#pragma unroll 1
    for (int h = AH - 1; h >= 0; --h) {
        int a = h * 32 + lane;
        Box b = ld_box(sb, a < p.A ? a : 0);
        bool valid = a < p.A && b.present != 0.0f;
        box_dirtri(b, wx, wy);
        LanePrim pr = make_prim<3>(cam, wx, wy, a == 0 ? TDE_CLS_EGO_DIRECTION : TDE_CLS_DIRECTION, valid);
        raster_chunk(pr, P, ws->span, lane);
    }

    // class-index planes -> shared, then palette lookup with byte permutes and 128-bit stores
#pragma unroll
    for (int band = 0; band < 2; ++band)
#pragma unroll
        for (int b = 0; b < 4; ++b) ws->planes[(band * 32 + lane) * 4 + b] = P[band][b];
    __syncwarp();
    const unsigned short* pl16 = reinterpret_cast<const unsigned short*>(ws->planes);
    uint8_t* out = p.obs + (size_t)e * (TDE_OBS_C * TDE_OBS_H * TDE_OBS_W);
#pragma unroll 1
    for (int it = 0; it < TDE_OBS_H / 8; ++it) {
        int row = it * 8 + (lane >> 2), q = lane & 3;
        uint32_t b0 = pl16[(row * 4 + 0) * 4 + q], b1 = pl16[(row * 4 + 1) * 4 + q];
        uint32_t b2 = pl16[(row * 4 + 2) * 4 + q], b3 = pl16[(row * 4 + 3) * 4 + q];
        uint32_t idx[2];
        idx[0] = spread8(b0) | (spread8(b1) << 1) | (spread8(b2) << 2) | (spread8(b3) << 3);
        idx[1] = spread8(b0 >> 8) | (spread8(b1 >> 8) << 1) | (spread8(b2 >> 8) << 2) | (spread8(b3 >> 8) << 3);
        uint32_t sel7[4], himask[4];
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t sel = (idx[g >> 1] >> (16 * (g & 1))) & 0xffffu;
            sel7[g] = sel & 0x7777u;
            himask[g] = __byte_perm(0x0000ff00u, 0u, (sel >> 3) & 0x1111u);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) {
            uint32_t w[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                uint32_t lo = __byte_perm(p.pal[ch][0], p.pal[ch][1], sel7[g]);
                uint32_t hi = __byte_perm(p.pal[ch][2], p.pal[ch][3], sel7[g]);
                w[g] = (lo & ~himask[g]) | (hi & himask[g]);
            }
            uint4 v = make_uint4(w[0], w[1], w[2], w[3]);
            *reinterpret_cast<uint4*>(out + ch * (TDE_OBS_H * TDE_OBS_W) + row * TDE_OBS_W + q * 16) = v;
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------- the fused step kernel

template <int AH>
__global__ void __launch_bounds__(TDE_WARPS_PER_BLOCK * 32) tde_step_kernel(const StepParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch* ws = reinterpret_cast<WarpScratch*>(smem_raw) + warp;
    const int warps_total = gridDim.x * TDE_WARPS_PER_BLOCK;
    const tde_config& c = p.cfg;
    double st_acc = 0.0;  // lane k accumulates statistic k

#pragma unroll 1
    for (int e = blockIdx.x * TDE_WARPS_PER_BLOCK + warp; e < p.E; e += warps_total) {
        int myvar = lane < 8 ? p.vars[(size_t)e * 8 + lane] : 0;
        int s = __shfl_sync(FULL_MASK, myvar, 0), step = __shfl_sync(FULL_MASK, myvar, 1);
        int target = __shfl_sync(FULL_MASK, myvar, 2), reached = __shfl_sync(FULL_MASK, myvar, 3);
        int lphase = __shfl_sync(FULL_MASK, myvar, 4), episode = __shfl_sync(FULL_MASK, myvar, 5);
        int m = __shfl_sync(FULL_MASK, myvar, 6);
        float4 st[AH], at[AH];
#pragma unroll
        for (int h = 0; h < AH; ++h) {
            int a = h * 32 + lane;
            st[h] = make_float4(0.f, 0.f, 0.f, 0.f); at[h] = make_float4(1.f, 1.f, 1.f, 0.f);
            if (a < p.A) { st[h] = p.state[(size_t)e * p.A + a]; at[h] = p.attr[(size_t)e * p.A + a]; }
        }
        // snapshot of the ego before the step (gym_env.py:371-375)
        float lx = __shfl_sync(FULL_MASK, st[0].x, 0), ly = __shfl_sync(FULL_MASK, st[0].y, 0);
        float lpsi = __shfl_sync(FULL_MASK, st[0].z, 0), lv = __shfl_sync(FULL_MASK, st[0].w, 0);

        if (p.phases & TDE_PH_KINEMATICS) {
            const ScenDev& S = p.scens[s];
            int t = step + 1;
            float act_a = p.actions[2 * (size_t)e], act_b = p.actions[2 * (size_t)e + 1];
#pragma unroll
            for (int h = 0; h < AH; ++h) {
                int a = h * 32 + lane;
                if (a < p.A && at[h].w != 0.0f) {
                    if (a == 0) {
                        st[h] = tde_bicycle(st[h], act_a, act_b, at[h].z, c.dt);
                    } else if (t < S.rep_T && S.rep_mask[(size_t)t * p.A + a]) {
                        st[h] = S.rep_states[(size_t)t * p.A + a];
                    } else {
                        st[h] = tde_bicycle(st[h], 0.0f, 0.0f, at[h].z, c.dt);
                    }
                    p.state[(size_t)e * p.A + a] = st[h];
                }
            }
            step = t;
        }
        // boxes of this env -> shared
#pragma unroll
        for (int h = 0; h < AH; ++h) {
            int a = h * 32 + lane;
            if (a < p.A) st_box(ws->box, a, tde_make_box(st[h].x, st[h].y, st[h].z, at[h].x, at[h].y, at[h].w));
        }
        __syncwarp();

        float4 inf0 = make_float4(0.f, 0.f, 0.f, 0.f);  // ego's infractions, valid on lane 0
        if (p.phases & TDE_PH_INFRACTIONS) {
            const MapDev& M = p.maps[m];
#pragma unroll
            for (int h = 0; h < AH; ++h) {
                int a = h * 32 + lane;
                float4 inf = make_float4(0.f, 0.f, 0.f, 0.f);
                bool mine = a < p.A && at[h].w != 0.0f;
                Box me = ld_box(ws->box, a < p.A ? a : 0);
                // all-pairs SAT: every lane tests its box against box j (broadcast from shared)
                float cnt = 0.0f;
#pragma unroll 1
                for (int j = 0; j < p.A; ++j) {
                    Box o = ld_box(ws->box, j);
                    bool cand = mine && j != a && o.present != 0.0f && !tde_far_apart(me, o);
                    if (__any_sync(FULL_MASK, cand)) {
                        if (cand && tde_overlap(me, o)) cnt += 1.0f;
                    }
                }
                if (mine) {
                    inf.x = cnt;
                    inf.y = offroad_box(M, me, c.offroad_threshold);
                    inf.z = tl_violation_box(M, me, c.tl_rear_factor, step, lphase);
                    inf.w = wrong_way_box(M, me);
                }
                if (a < p.A) p.infr[(size_t)e * p.A + a] = inf;
                if (h == 0) inf0 = inf;
            }
        } else if (p.phases & TDE_PH_REWARD) {
            if (lane == 0) inf0 = p.infr[(size_t)e * p.A];
        }

        if (p.phases & TDE_PH_REWARD) {
            const ScenDev& S = p.scens[s];
            float x = __shfl_sync(FULL_MASK, st[0].x, 0), y = __shfl_sync(FULL_MASK, st[0].y, 0);
            float psi = __shfl_sync(FULL_MASK, st[0].z, 0), spd = __shfl_sync(FULL_MASK, st[0].w, 0);
            float i_col = __shfl_sync(FULL_MASK, inf0.x, 0), i_off = __shfl_sync(FULL_MASK, inf0.y, 0);
            float i_tl = __shfl_sync(FULL_MASK, inf0.z, 0), i_ww = __shfl_sync(FULL_MASK, inf0.w, 0);
            float dx = x - lx, dy = y - ly;
            float d = sqrtf(dx * dx + dy * dy);                                      // :401
            float dist_reward = d > c.distance_cutoff ? c.distance_bonus : 0.0f;     // :402
            float sd, cd;
            tde_sincosf(psi - lpsi, sd, cd);
            float psi_reward = (1.0f - cd) * (-c.heading_penalty);                   // :403
            bool hit = false;
            if (target < S.W) {                                                      // :391-394
                float2 w = S.wp[target];
                float tx = x - w.x, ty = y - w.y;
                hit = sqrtf(tx * tx + ty * ty) < c.reach_radius;
            }
            float reach_reward = hit ? c.waypoint_bonus : 0.0f;                      // :404-408
            if (hit) reached += 1;                                                   // :406
            float r = (reach_reward + dist_reward) + psi_reward;                     // :410
            bool term = c.terminated_at_infraction && (i_off > 0.0f || i_col > 0.0f || i_tl > 0.0f);  // :413-417
            bool trunc = step >= c.max_environment_steps;                            // :134-135
            float ep_ret = p.ep_return[e] + r;
            bool done = term || trunc;
            // info row (get_info :419-437): lane k writes column k
            float col = 0.0f;
            switch (lane) {
                case TDE_INFO_OFFROAD: col = i_off; break;
                case TDE_INFO_COLLISION: col = i_col; break;
                case TDE_INFO_TL_VIOLATION: col = i_tl; break;
                case TDE_INFO_IS_SUCCESS: col = trunc ? 1.0f : 0.0f; break;
                case TDE_INFO_REACHED_WAYPOINT_NUM: col = (float)reached; break;
                case TDE_INFO_PSI_SMOOTHNESS: col = fabsf((lpsi - psi) / c.dt); break;
                case TDE_INFO_PSI_REWARD: col = psi_reward; break;
                case TDE_INFO_DIST_REWARD: col = dist_reward; break;
                case TDE_INFO_SPEED_SMOOTHNESS: col = fabsf((lv - spd) / c.dt); break;
                case TDE_INFO_WRONG_WAY: col = i_ww; break;
                case TDE_INFO_EPISODE_RETURN: col = ep_ret; break;
                case TDE_INFO_EPISODE_LENGTH: col = (float)step; break;
                case TDE_INFO_SCENARIO: col = (float)s; break;
                case TDE_INFO_DID_RESET: col = (done && c.auto_reset) ? 1.0f : 0.0f; break;
                default: break;
            }
            if (lane < TDE_INFO_STRIDE) p.info[(size_t)e * TDE_INFO_STRIDE + lane] = col;
            if (lane == 0) {
                p.reward[e] = r;
                p.terminated[e] = term ? 1 : 0;
                p.truncated[e] = trunc ? 1 : 0;
                p.ep_return[e] = ep_ret;
            }
            if (hit) target += 1;                                                    // :378-383
            // episode statistics: lane k owns statistic k
            double add = 0.0;
            if (lane == TDE_STAT_STEPS) add = 1.0;
            if (done) {
                switch (lane) {
                    case TDE_STAT_EPISODES: add = 1.0; break;
                    case TDE_STAT_RETURN_SUM: add = (double)ep_ret; break;
                    case TDE_STAT_LENGTH_SUM: add = (double)step; break;
                    case TDE_STAT_OFFROAD: add = i_off > 0.0f ? 1.0 : 0.0; break;
                    case TDE_STAT_COLLISION: add = i_col > 0.0f ? 1.0 : 0.0; break;
                    case TDE_STAT_TL_VIOLATION: add = i_tl > 0.0f ? 1.0 : 0.0; break;
                    case TDE_STAT_SUCCESS: add = trunc ? 1.0 : 0.0; break;
                    case TDE_STAT_REACHED_WAYPOINTS: add = (double)reached; break;
                    default: break;
                }
            }
            st_acc += add;
            if (done && c.auto_reset) {
                __syncwarp();
                reset_env_warp<AH>(p, e, lane, s, step, target, reached, lphase, episode, m, st, at);
#pragma unroll
                for (int h = 0; h < AH; ++h) {
                    int a = h * 32 + lane;
                    if (a < p.A) st_box(ws->box, a, tde_make_box(st[h].x, st[h].y, st[h].z, at[h].x, at[h].y, at[h].w));
                }
                __syncwarp();
            }
        }
        if (p.phases & (TDE_PH_KINEMATICS | TDE_PH_REWARD)) store_vars(p, e, lane, s, step, target, reached, lphase, episode, m);

        if ((p.phases & TDE_PH_RENDER) && p.obs != nullptr)
            render_env_warp<AH>(p, p.maps[m], p.scens[s], e, lane, ws, step, lphase, target);
        __syncwarp();
    }
    if ((p.phases & TDE_PH_REWARD) && lane < TDE_NUM_STATS && st_acc != 0.0) atomicAdd(&p.stats[lane], st_acc);
}

template <int AH>
__global__ void __launch_bounds__(TDE_WARPS_PER_BLOCK * 32) tde_reset_kernel(const StepParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_total = gridDim.x * TDE_WARPS_PER_BLOCK;
    for (int e = blockIdx.x * TDE_WARPS_PER_BLOCK + warp; e < p.E; e += warps_total) {
        if (p.reset_mask != nullptr && p.reset_mask[e] == 0) continue;
        int myvar = lane < 8 ? p.vars[(size_t)e * 8 + lane] : 0;
        int s, step, target, reached, lphase, m;
        int episode = __shfl_sync(FULL_MASK, myvar, 5);
        float4 st[AH], at[AH];
        reset_env_warp<AH>(p, e, lane, s, step, target, reached, lphase, episode, m, st, at);
        store_vars(p, e, lane, s, step, target, reached, lphase, episode, m);
    }
}

// ---------------------------------------------------------------- stateless micro-benchmark kernels (config C4)

// All-pairs oriented-box collision counts on caller-provided boxes: warp per env, boxes tiled in
// shared memory, every unordered pair tested once (the SAT is bitwise symmetric) along the
// "diagonals" j = i + k, hits exchanged with warp shuffles / shared counters.
template <int AH>
__global__ void __launch_bounds__(TDE_WARPS_PER_BLOCK * 32) tde_collision_kernel(const float4* __restrict__ state,
                                                                                  const float4* __restrict__ attr, int E, int A,
                                                                                  float* __restrict__ out) {
    __shared__ float4 sbox[TDE_WARPS_PER_BLOCK][TDE_MAX_AGENTS * 2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_total = gridDim.x * TDE_WARPS_PER_BLOCK;
    float4* sb = sbox[warp];
    for (int e = blockIdx.x * TDE_WARPS_PER_BLOCK + warp; e < E; e += warps_total) {
        Box me[AH];
#pragma unroll
        for (int h = 0; h < AH; ++h) {
            int a = h * 32 + lane;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = make_float4(1.f, 1.f, 1.f, 0.f);
            if (a < A) { s4 = state[(size_t)e * A + a]; a4 = attr[(size_t)e * A + a]; }
            me[h] = tde_make_box(s4.x, s4.y, s4.z, a4.x, a4.y, a4.w);
            if (a < A) st_box(sb, a, me[h]);
        }
        __syncwarp();
#pragma unroll
        for (int h = 0; h < AH; ++h) {
            int a = h * 32 + lane;
            bool mine = a < A && me[h].present != 0.0f;
            float cnt = 0.0f;
#pragma unroll 1
            for (int j = 0; j < A; ++j) {
                Box o = ld_box(sb, j);
                bool cand = mine && j != a && o.present != 0.0f && !tde_far_apart(me[h], o);
                if (__any_sync(FULL_MASK, cand)) {
                    if (cand && tde_overlap(me[h], o)) cnt += 1.0f;
                }
            }
            if (a < A) out[(size_t)e * A + a] = cnt;
        }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256) tde_offroad_kernel(const MapDev* maps, int map_id, float thr,
                                                          const float4* __restrict__ state, const float4* __restrict__ attr,
                                                          int n, float* __restrict__ out) {
    const MapDev& M = maps[map_id];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        float4 s4 = state[i], a4 = attr[i];
        float val = 0.0f;
        if (a4.w != 0.0f) {
            Box b = tde_make_box(s4.x, s4.y, s4.z, a4.x, a4.y, a4.w);
            val = offroad_box(M, b, thr);
        }
        out[i] = val;
    }
}
