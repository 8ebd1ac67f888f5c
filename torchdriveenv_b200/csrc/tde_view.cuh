// tde_view.cuh — the recording view (sm_100a): one env drawn at any resolution from a free camera.
//
// Replaces BirdviewRecordingWrapper(simulator, res=Resolution(video_res, video_res), fov=video_fov)
// (gym_env.py:52-53, :295-297) + simulator.get_birdviews() (:174).  Not on the per-step hot path: it runs
// once per recorded frame for ONE env, so it is written for clarity and exact agreement with the oracle's
// orc_render_view, not for the last microsecond: a first kernel transforms and snaps every primitive of
// the env (raw road / marking triangles, stop lines, goal diamond, vehicles, direction triangles), a second
// one walks 32x8-pixel tiles, compacts the primitives whose bounding box meets the tile into shared
// memory and evaluates the 64-bit edge functions per pixel centre (top-left rule, highest class wins).
#pragma once
#include "tde_kernels.cuh"

#define TDE_VIEW_SNAP_MAX 1048575.0f
#define TDE_VIEW_MAX_RES 4096

struct ViewParams {
    int env, W, H;
    float ex, ey, psi, ce, se, ppm, ppmy;
    uint8_t pal[TDE_NUM_CLASSES * 3];
};

struct ViewPrim {            // 48 bytes; n = 0: rejected
    int x[4], y[4];          // snapped vertices, oriented so that the signed area is positive
    int n, cls;
    short i0, i1, j0, j1;    // pixel range that can hold covered centres
};

__device__ __forceinline__ void view_emit(const ViewParams& v, const float (&wx)[4], const float (&wy)[4], int n, int cls, ViewPrim* out) {
    ViewPrim P;
    P.n = 0; P.cls = cls; P.i0 = P.j0 = 0; P.i1 = P.j1 = -1;
#pragma unroll
    for (int k = 0; k < 4; ++k) { P.x[k] = 0; P.y[k] = 0; }
    float fx[4], fy[4];
    float minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k < n) {
            float dx = wx[k] - v.ex, dy = wy[k] - v.ey;
            float cx = dx * v.ce + dy * v.se;
            float cy = dy * v.ce - dx * v.se;
            fx[k] = cx * v.ppm + 0.5f * (float)v.W;
            fy[k] = cy * v.ppmy + 0.5f * (float)v.H;
            minx = fminf(minx, fx[k]); maxx = fmaxf(maxx, fx[k]);
            miny = fminf(miny, fy[k]); maxy = fmaxf(maxy, fy[k]);
        }
    }
    const bool in_view = maxx >= -1.0f && minx <= (float)v.W + 1.0f && maxy >= -1.0f && miny <= (float)v.H + 1.0f;
    if (in_view) {
        long long X[4], Y[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            X[k] = 0; Y[k] = 0;
            if (k < n) {
                float rx = rintf(fx[k] * 16.0f), ry = rintf(fy[k] * 16.0f);
                rx = fminf(fmaxf(rx, -TDE_VIEW_SNAP_MAX), TDE_VIEW_SNAP_MAX);
                ry = fminf(fmaxf(ry, -TDE_VIEW_SNAP_MAX), TDE_VIEW_SNAP_MAX);
                X[k] = (long long)(int)rx; Y[k] = (long long)(int)ry;
            }
        }
        long long area2 = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (k < n) { const int k1 = (k + 1 == n) ? 0 : k + 1; area2 += X[k] * Y[k1] - X[k1] * Y[k]; }
        if (area2 != 0) {
            long long xmin = X[0], xmax = X[0], ymin = Y[0], ymax = Y[0];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (k < n) { xmin = min(xmin, X[k]); xmax = max(xmax, X[k]); ymin = min(ymin, Y[k]); ymax = max(ymax, Y[k]); }
            long long i0 = (xmin - 8) >= 0 ? (xmin - 8 + 15) / 16 : 0, i1 = (xmax - 8) >= 0 ? (xmax - 8) / 16 : -1;
            long long j0 = (ymin - 8) >= 0 ? (ymin - 8 + 15) / 16 : 0, j1 = (ymax - 8) >= 0 ? (ymax - 8) / 16 : -1;
            i1 = min(i1, (long long)v.W - 1); j1 = min(j1, (long long)v.H - 1);
            if (i0 <= i1 && j0 <= j1) {
                P.n = n;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (k < n) { const int src = area2 > 0 ? k : n - 1 - k; P.x[k] = (int)X[src]; P.y[k] = (int)Y[src]; }
                P.i0 = (short)i0; P.i1 = (short)i1; P.j0 = (short)j0; P.j1 = (short)j1;
            }
        }
    }
    *out = P;
}

// primitive i of env v.env: [0, ntri) road triangles, [.., +nmark) marking triangles, [.., +nstop) stop lines,
// one goal waypoint slot, then (rectangle, direction triangle) per agent slot
__global__ void __launch_bounds__(128) tde_view_prims_kernel(const StepParams p, ViewParams v, ViewPrim* __restrict__ out, int n_total) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_total) return;
    tde_sincosf(v.psi, v.se, v.ce);
    const int* vars = p.vars + (size_t)v.env * 8;
    const int s = vars[0], step = vars[1], target = vars[2], lphase = vars[4], m = vars[6];
    const MapDev& M = p.maps[m];
    const ScenDev& S = p.scens[s];
    float wx[4] = {0.f, 0.f, 0.f, 0.f}, wy[4] = {0.f, 0.f, 0.f, 0.f};
    int n = 0, cls = 0;
    int k = i;
    if (k < M.ntri) {
        const float4 t0 = M.tri[3 * k], t1 = M.tri[3 * k + 1];
        wx[0] = t0.x; wy[0] = t0.y; wx[1] = t0.z; wy[1] = t0.w; wx[2] = t1.x; wy[2] = t1.y;
        n = 3; cls = TDE_CLS_ROAD;
    } else if ((k -= M.ntri) < M.nmark) {
        const float* t = M.mark_raw + 6 * (size_t)k;
        wx[0] = t[0]; wy[0] = t[1]; wx[1] = t[2]; wy[1] = t[3]; wx[2] = t[4]; wy[2] = t[5];
        n = 3; cls = TDE_CLS_LANE_MARKING;
    } else if ((k -= M.nmark) < M.nstop) {
        const float4 u = M.stop[2 * k], w = M.stop[2 * k + 1];
        Box b; b.x = u.x; b.y = u.y; b.hl = u.z; b.hw = u.w; b.c = w.x; b.s = w.y; b.present = 1.0f; b.r = 0.0f;
#pragma unroll
        for (int c = 0; c < 4; ++c) tde_box_corner(b, c, wx[c], wy[c]);
        n = 4; cls = TDE_CLS_TL_GREEN + light_state_at(M, step, lphase, k);
    } else if ((k -= M.nstop) < 1) {
        if (target < S.W) {
            const float2 w = S.wp[target];
            const float r = 2.0f;
            wx[0] = w.x + r; wy[0] = w.y; wx[1] = w.x; wy[1] = w.y + r; wx[2] = w.x - r; wy[2] = w.y; wx[3] = w.x; wy[3] = w.y - r;
            n = 4; cls = TDE_CLS_WAYPOINT;
        }
    } else {
        k -= 1;
        const int a = k >> 1;
        float4 st = make_float4(0.f, 0.f, 0.f, 0.f), at = st;
        if (a < p.A) { st = p.state[(size_t)v.env * p.A + a]; at = p.attr[(size_t)v.env * p.A + a]; }   // the scratch is sized for the largest map
        if (at.w != 0.0f) {
            const Box b = tde_make_box(st.x, st.y, st.z, at.x, at.y, at.w);
            if (k & 1) {
                const float ox[3] = {b.hl, 0.5f * b.hl, 0.5f * b.hl}, oy[3] = {0.0f, b.hw, -b.hw};
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    wx[c] = b.x + (ox[c] * b.c - oy[c] * b.s);
                    wy[c] = b.y + (ox[c] * b.s + oy[c] * b.c);
                }
                n = 3; cls = a == 0 ? TDE_CLS_EGO_DIRECTION : TDE_CLS_DIRECTION;
            } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) tde_box_corner(b, c, wx[c], wy[c]);
                n = 4; cls = a == 0 ? TDE_CLS_EGO : TDE_CLS_VEHICLE;
            }
        }
    }
    if (n == 0) { ViewPrim P; memset(&P, 0, sizeof(P)); P.i1 = P.j1 = -1; out[i] = P; return; }
    view_emit(v, wx, wy, n, cls, &out[i]);
}

#define TDE_VIEW_TX 32
#define TDE_VIEW_TY 8
__global__ void __launch_bounds__(TDE_VIEW_TX * TDE_VIEW_TY) tde_view_raster_kernel(const ViewPrim* __restrict__ prims, int n_total,
                                                                                    const ViewParams v, uint8_t* __restrict__ out) {
    __shared__ ViewPrim sp[TDE_VIEW_TX * TDE_VIEW_TY];
    __shared__ int cnt;
    const int tid = threadIdx.y * TDE_VIEW_TX + threadIdx.x;
    const int tx0 = blockIdx.x * TDE_VIEW_TX, ty0 = blockIdx.y * TDE_VIEW_TY;
    const int i = tx0 + threadIdx.x, j = ty0 + threadIdx.y;
    const long long px = 16ll * i + 8, py = 16ll * j + 8;
    int best = 0;
    for (int base = 0; base < n_total; base += TDE_VIEW_TX * TDE_VIEW_TY) {
        if (tid == 0) cnt = 0;
        __syncthreads();
        const int q = base + tid;
        if (q < n_total) {
            const ViewPrim P = prims[q];
            if (P.n >= 3 && P.i0 <= tx0 + TDE_VIEW_TX - 1 && P.i1 >= tx0 && P.j0 <= ty0 + TDE_VIEW_TY - 1 && P.j1 >= ty0)
                sp[atomicAdd(&cnt, 1)] = P;
        }
        __syncthreads();
        const int c = cnt;
        for (int k = 0; k < c; ++k) {
            const ViewPrim& P = sp[k];
            if (P.cls <= best || i < P.i0 || i > P.i1 || j < P.j0 || j > P.j1) continue;
            bool in = true;
            for (int e = 0; e < P.n && in; ++e) {
                const int e1 = (e + 1 == P.n) ? 0 : e + 1;
                const long long dx = (long long)P.x[e1] - P.x[e], dy = (long long)P.y[e1] - P.y[e];
                if (dx == 0 && dy == 0) continue;
                const long long E = dx * (py - P.y[e]) - dy * (px - P.x[e]);
                const bool incl = dy < 0 || (dy == 0 && dx > 0);
                if (E < 0 || (E == 0 && !incl)) in = false;
            }
            if (in) best = P.cls;
        }
        __syncthreads();
    }
    if (i < v.W && j < v.H) {
        const size_t plane = (size_t)v.W * v.H, o = (size_t)j * v.W + i;
        out[o] = v.pal[3 * best]; out[plane + o] = v.pal[3 * best + 1]; out[2 * plane + o] = v.pal[3 * best + 2];
    }
}
