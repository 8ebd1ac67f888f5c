// tde_b200.cu — host side of libtde_b200.so: handle, scenario upload (incl. the nearest-candidate
// grid over the lane mesh), launches.  C ABI declared in include/tde_b200.h.
#ifndef TDE_HOST_EMU
#include <cuda_runtime.h>
#endif

#include <algorithm>
#include <array>
#include <map>
#include <memory>
#include <tuple>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include <immintrin.h>
#ifdef __linux__
#include <sched.h>
#endif

#include "tde_kernels.cuh"
#include "tde_view.cuh"

static thread_local std::string g_create_error;

// Everything tde_upload_scenarios builds on the device.  Handles made by tde_clone share one set (reference-counted).
struct ScenTables {
    MapDev* maps_dev = nullptr;
    ScenDev* scens_dev = nullptr;
    std::vector<MapDev> maps_host;
    std::vector<ScenDev> scens_host;
    std::vector<void*> scenario_allocs;
    std::vector<std::array<int, 8>> map_info;  // ntri nmark nstop gnx gny items safe_cells render_prims
    int num_maps = 0, num_scen = 0, device = 0;
    // the staged blob of map tables (nullptr: they do not fit beside the SAT scratch, or staging is switched off -> the
    // physics kernel reads them from global memory)
    unsigned char* stage_blob = nullptr;
    unsigned int stage_bytes = 0, stage_maps_off = 0;
    ~ScenTables() {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(device);
        for (void* p : scenario_allocs) cudaFree(p);
        if (maps_dev) cudaFree(maps_dev);
        if (scens_dev) cudaFree(scens_dev);
        if (stage_blob) cudaFree(stage_blob);
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct ExpandPool;
static void delete_pool(ExpandPool* pool);

struct tde_handle {
    tde_config cfg;
    std::shared_ptr<ScenTables> tab;
    int E = 0, A = 0, device = 0, sm_count = 0;
    float4 *state = nullptr, *attr = nullptr, *infr = nullptr;
    int* vars = nullptr;
    float* ep_return = nullptr;
    int *scen_lo = nullptr, *scen_hi = nullptr;
    double* stats = nullptr;
    uint8_t* restart = nullptr;
    unsigned int* tickets = nullptr;
    uint8_t* done_mask = nullptr;   // [E] envs that finished in the current tde_step_terminal
    uint8_t palette[TDE_NUM_CLASSES * 3];
    unsigned long long seed = 0;
    bool uploaded = false, was_reset = false;
    long long launches = 0;
    int grid_phys = 0, grid_render = 0;
    size_t smem_render = 0;
    // physics launch: warps per CTA, CTAs per SM, dynamic shared memory
    int phys_wpb = TDE_WARPS_PER_BLOCK, phys_per_sm = 1;
    size_t smem_phys = 0;
    // device staging for tde_step_host
    float* h_actions = nullptr; uint8_t* h_obs = nullptr; float* h_reward = nullptr;
    uint8_t *h_term = nullptr, *h_trunc = nullptr; float* h_info = nullptr;
    ViewPrim* view_prims = nullptr; int view_cap = 0;   // tde_render_view scratch
    cudaStream_t copy_stream = nullptr;          // tde_step_host: observation chunks go back while later chunks are computed
    cudaEvent_t chunk_done[16] = {}, copies_done = nullptr;
    // tde_step_host, compact mode: the 4-bit class image crosses PCIe and is expanded to RGB planes by host threads
    uint8_t *h_nib = nullptr, *pin_nib = nullptr;   // [E][64][32] on the device / in pinned host memory
    cudaEvent_t chunk_copied[16] = {};
    ExpandPool* pool = nullptr;
    bool pdl_physics = true;  // TDE_PDL_PHYSICS=0 switches the programmatic dependent launch of render-less steps off
    std::string err;
};

static const uint8_t k_default_palette[TDE_NUM_CLASSES * 3] = {
    0, 0, 0, 128, 128, 128, 255, 255, 255, 0, 200, 0, 230, 200, 0, 220, 0, 0,
    0, 170, 255, 60, 90, 220, 250, 120, 0, 200, 220, 255, 255, 230, 150,
};

static int fail(tde_handle* h, int code, const std::string& msg) {
    if (h) h->err = msg; else g_create_error = msg;
    return code;
}
#define CUDA_TRY(h, expr)                                                                         \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(h, TDE_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
    } while (0)

// Every export runs on the handle's GPU and leaves the caller's current device as it found it (PyTorch reads the
// current device through cudaGetDevice: a handle on another GPU must not switch it behind the caller's back).
struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        ok = prev == dev || cudaSetDevice(dev) == cudaSuccess;
        if (prev == dev) prev = -1;   // nothing to restore
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define TDE_ON_DEVICE(h)                                                              \
    DeviceGuard _guard((h)->device);                                                  \
    if (!_guard.ok) return fail(h, TDE_E_CUDA, "cudaSetDevice failed for the handle's device")

// ------------------------------------------------------------------ small prep kernels

// raw (M,8) triangles -> 3 float4 records with per-edge 1/len^2 (same binary32 ops as the oracle)
__global__ void prep_tris_kernel(const float* __restrict__ raw, int n, float4* __restrict__ out) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n) return;
    const float* r = raw + 8 * (size_t)t;
    float ax = r[0], ay = r[1], bx = r[2], by = r[3], cx = r[4], cy = r[5];
    auto il = [](float ax, float ay, float bx, float by) {
        float abx = bx - ax, aby = by - ay;
        float l2 = abx * abx + aby * aby;
        return l2 > 0.0f ? 1.0f / l2 : 0.0f;
    };
    out[3 * t] = make_float4(ax, ay, bx, by);
    out[3 * t + 1] = make_float4(cx, cy, il(ax, ay, bx, by), il(bx, by, cx, cy));
    out[3 * t + 2] = make_float4(il(cx, cy, ax, ay), r[6], r[7], 0.0f);
}
// raw (L,5) stop lines -> [x y hl hw][c s rr 0], rr = circumradius bound + 5 mm (broad phase)
__global__ void prep_stops_kernel(const float* __restrict__ raw, int n, float4* __restrict__ out) {
    int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= n) return;
    const float* r = raw + 5 * (size_t)l;
    Box b = tde_make_box(r[0], r[1], r[4], r[2], r[3], 1.0f);
    out[2 * l] = make_float4(b.x, b.y, b.hl, b.hw);
    out[2 * l + 1] = make_float4(b.c, b.s, b.hl + b.hw + 0.005f, 0.0f);
}

// ------------------------------------------------------------------ nearest-candidate grid (host, float64)

#ifndef TDE_GRID_MIN_CELL
#define TDE_GRID_MIN_CELL 0.7
#endif
#ifndef TDE_GRID_MAX_CELLS
#define TDE_GRID_MAX_CELLS 32768.0
#endif
namespace {
struct P2 { double x, y; };
inline double seg_d2(P2 p, P2 a, P2 b) {
    double abx = b.x - a.x, aby = b.y - a.y, apx = p.x - a.x, apy = p.y - a.y;
    double l2 = abx * abx + aby * aby;
    double t = l2 > 0 ? (apx * abx + apy * aby) / l2 : 0.0;
    t = std::min(1.0, std::max(0.0, t));
    double qx = apx - t * abx, qy = apy - t * aby;
    return qx * qx + qy * qy;
}
inline double tri_dist(P2 p, const float* t) {
    P2 a{t[0], t[1]}, b{t[2], t[3]}, c{t[4], t[5]};
    double c0 = (b.x - a.x) * (p.y - a.y) - (b.y - a.y) * (p.x - a.x);
    double c1 = (c.x - b.x) * (p.y - b.y) - (c.y - b.y) * (p.x - b.x);
    double c2 = (a.x - c.x) * (p.y - c.y) - (a.y - c.y) * (p.x - c.x);
    if ((c0 >= 0 && c1 >= 0 && c2 >= 0) || (c0 <= 0 && c1 <= 0 && c2 <= 0)) return 0.0;
    return std::sqrt(std::min(seg_d2(p, a, b), std::min(seg_d2(p, b, c), seg_d2(p, c, a))));
}

struct Grid {
    float gx0 = 0, gy0 = 0, inv_cell = 1;
    int nx = 0, ny = 0;
    std::vector<int> cell_start;
    std::vector<uint16_t> items;
    std::vector<uint16_t> meta;  // n_overlapping | TDE_CELL_SAFE
};

// exact distance between the square [c +- half] and a triangle (0 when they intersect)
double box_tri_mindist(P2 c, double half, const float* t) {
    P2 v[3] = {{t[0], t[1]}, {t[2], t[3]}, {t[4], t[5]}};
    bool sep = false;
    double mnx = std::min(v[0].x, std::min(v[1].x, v[2].x)), mxx = std::max(v[0].x, std::max(v[1].x, v[2].x));
    double mny = std::min(v[0].y, std::min(v[1].y, v[2].y)), mxy = std::max(v[0].y, std::max(v[1].y, v[2].y));
    if (mxx < c.x - half || mnx > c.x + half || mxy < c.y - half || mny > c.y + half) sep = true;
    for (int k = 0; k < 3 && !sep; ++k) {
        P2 a = v[k], b = v[(k + 1) % 3];
        double nx = -(b.y - a.y), ny = b.x - a.x;
        double r = half * (std::fabs(nx) + std::fabs(ny)), pc = c.x * nx + c.y * ny;
        double lo = 1e300, hi = -1e300;
        for (int q = 0; q < 3; ++q) { double pr = v[q].x * nx + v[q].y * ny; lo = std::min(lo, pr); hi = std::max(hi, pr); }
        if (hi < pc - r || lo > pc + r) sep = true;
    }
    if (!sep) return 0.0;
    double d = 1e300;
    P2 corner[4] = {{c.x - half, c.y - half}, {c.x + half, c.y - half}, {c.x + half, c.y + half}, {c.x - half, c.y + half}};
    for (int k = 0; k < 4; ++k) d = std::min(d, tri_dist(corner[k], t));
    for (int k = 0; k < 3; ++k) {
        double ex = std::max(std::fabs(v[k].x - c.x) - half, 0.0), ey = std::max(std::fabs(v[k].y - c.y) - half, 0.0);
        d = std::min(d, std::hypot(ex, ey));
    }
    return d;
}

// upper bound on the distance to the mesh over a square: 0 where a single triangle covers the
// square, otherwise refined by quartering down to `depth`
double cover_bound(P2 c, double half, const float* tris, const std::vector<int>& cand, int depth) {
    P2 corner[4] = {{c.x - half, c.y - half}, {c.x + half, c.y - half}, {c.x + half, c.y + half}, {c.x - half, c.y + half}};
    double best = 1e300;
    for (int t : cand) {
        double mx = 0;
        for (int k = 0; k < 4; ++k) mx = std::max(mx, tri_dist(corner[k], tris + 8 * (size_t)t));
        best = std::min(best, mx);
        if (best == 0.0) return 0.0;
    }
    if (depth == 0) return best;
    double worst = 0;
    for (int q = 0; q < 4; ++q) {
        P2 cc{c.x + ((q & 1) ? 0.5 : -0.5) * half, c.y + ((q & 2) ? 0.5 : -0.5) * half};
        worst = std::max(worst, cover_bound(cc, 0.5 * half, tris, cand, depth - 1));
        if (worst >= best) return best;
    }
    return std::min(best, worst);
}

// Nearest-candidate grid.  For every cell (dilated a little, to absorb the binary32 rounding of the
// device's cell index) keep (a) the triangles that intersect it, first, and (b) every other triangle
// that can be the nearest one for some point of the cell: with U = min_t max_{p in cell} dist(p, t) an
// upper bound of the nearest distance, a triangle further than U from the whole cell can never win.
// A cell is SAFE when every one of its points is provably closer to the mesh than `threshold`.
Grid build_grid(const float* tris, int M, double threshold) {
    Grid g;
    if (M <= 0) { g.nx = g.ny = 0; g.cell_start.assign(1, 0); return g; }
    double lox = 1e300, loy = 1e300, hix = -1e300, hiy = -1e300;
    std::vector<P2> cen(M);
    std::vector<double> rad(M);
    for (int t = 0; t < M; ++t) {
        const float* r = tris + 8 * (size_t)t;
        for (int k = 0; k < 3; ++k) {
            lox = std::min(lox, (double)r[2 * k]); hix = std::max(hix, (double)r[2 * k]);
            loy = std::min(loy, (double)r[2 * k + 1]); hiy = std::max(hiy, (double)r[2 * k + 1]);
        }
        cen[t] = P2{(r[0] + r[2] + r[4]) / 3.0, (r[1] + r[3] + r[5]) / 3.0};
        double rr = 0;
        for (int k = 0; k < 3; ++k) rr = std::max(rr, std::hypot(r[2 * k] - cen[t].x, r[2 * k + 1] - cen[t].y));
        rad[t] = rr;
    }
    const double margin = 8.0;
    lox -= margin; loy -= margin; hix += margin; hiy += margin;
    double w = hix - lox, hgt = hiy - loy;
    double cell = std::max(TDE_GRID_MIN_CELL, std::sqrt(w * hgt / TDE_GRID_MAX_CELLS));
    g.nx = std::max(1, (int)std::ceil(w / cell));
    g.ny = std::max(1, (int)std::ceil(hgt / cell));
    g.gx0 = (float)lox; g.gy0 = (float)loy;
    g.inv_cell = (float)(1.0 / cell);
    // the device maps a point with floorf((p - gx0) * inv_cell) in binary32: use the same constants
    double x0 = g.gx0, y0 = g.gy0, cs = 1.0 / (double)g.inv_cell;
    double dil = 1e-3 * cs + 1e-3;  // dilation covering binary32 rounding of the cell index
    double half = 0.5 * cs + dil, hd = half * std::sqrt(2.0);
    g.cell_start.assign((size_t)g.nx * g.ny + 1, 0);
    g.meta.assign((size_t)g.nx * g.ny, 0);
    std::vector<int> plaus, over, near, both;
    std::vector<double> mind;
    for (int iy = 0; iy < g.ny; ++iy) {
        for (int ix = 0; ix < g.nx; ++ix) {
            P2 c{x0 + (ix + 0.5) * cs, y0 + (iy + 0.5) * cs};
            P2 corner[4] = {{c.x - half, c.y - half}, {c.x + half, c.y - half}, {c.x + half, c.y + half}, {c.x - half, c.y + half}};
            double Uub = 1e300;  // cheap upper bound on U from centroids
            for (int t = 0; t < M; ++t) Uub = std::min(Uub, std::hypot(c.x - cen[t].x, c.y - cen[t].y) + hd);
            plaus.clear(); mind.clear();
            double U = 1e300;
            for (int t = 0; t < M; ++t) {
                double dc = std::hypot(c.x - cen[t].x, c.y - cen[t].y);
                if (dc - rad[t] - hd > Uub) continue;
                const float* r = tris + 8 * (size_t)t;
                double md = box_tri_mindist(c, half, r);
                if (md > Uub) continue;
                double mx = 0;
                for (int k = 0; k < 4; ++k) mx = std::max(mx, tri_dist(corner[k], r));
                U = std::min(U, mx);
                plaus.push_back(t); mind.push_back(md);
            }
            double lim = U * (1.0 + 1e-4) + 1e-3;
            over.clear(); near.clear();
            double dmin = 1e300;
            for (size_t k = 0; k < plaus.size(); ++k) {
                dmin = std::min(dmin, mind[k]);
                if (mind[k] <= 1e-9) over.push_back(plaus[k]);
                else if (mind[k] <= lim) near.push_back(plaus[k]);
            }
            bool safe = false;
            if (threshold > 2e-3 && dmin < threshold) {
                both = over; both.insert(both.end(), near.begin(), near.end());
                double bound = U + 1e-3 < threshold ? U : cover_bound(c, half, tris, both, 4);
                safe = bound + 1e-3 < threshold;
            }
            // overlapping triangles that cover most of the cell first: a containment loop then ends early
            if (over.size() > 1) {
                std::vector<std::pair<int, int>> cov;
                for (int t : over) {
                    int hits = 0;
                    for (int sy = 0; sy < 4; ++sy)
                        for (int sx = 0; sx < 4; ++sx)
                            hits += tri_dist(P2{c.x + (sx - 1.5) * 0.25 * cs, c.y + (sy - 1.5) * 0.25 * cs}, tris + 8 * (size_t)t) == 0.0;
                    cov.push_back({-hits, t});
                }
                std::stable_sort(cov.begin(), cov.end());
                for (size_t k = 0; k < over.size(); ++k) over[k] = cov[k].second;
            }
            size_t nover = std::min<size_t>(over.size(), 0x7fff);
            for (int t : over) g.items.push_back((uint16_t)t);
            for (int t : near) g.items.push_back((uint16_t)t);
            g.meta[(size_t)iy * g.nx + ix] = (uint16_t)(nover | (safe ? TDE_CELL_SAFE : 0));
            g.cell_start[(size_t)iy * g.nx + ix + 1] = (int)g.items.size();
        }
    }
    return g;
}

// Morton (Z-order) key of a point, used to order the static triangles so that runs of 32 are compact
uint32_t morton_key(double x, double y, double lox, double loy, double inv) {
    auto spread = [](uint32_t v) { v &= 0xffff; v = (v | (v << 8)) & 0x00ff00ff; v = (v | (v << 4)) & 0x0f0f0f0f; v = (v | (v << 2)) & 0x33333333; v = (v | (v << 1)) & 0x55555555; return v; };
    uint32_t ix = (uint32_t)std::min(65535.0, std::max(0.0, (x - lox) * inv)), iy = (uint32_t)std::min(65535.0, std::max(0.0, (y - loy) * inv));
    return spread(ix) | (spread(iy) << 1);
}
// Render-only primitives: pairs of triangles that share an edge and form a strictly convex
// quadrilateral are merged (the fixed-point fill rule is watertight, so a quad covers exactly the
// union of its two triangles); the rest stay triangles (fourth vertex = first).  8 floats each.
std::vector<float> merge_into_quads(const float* tris, int n, int stride, double max_extent) {
    struct Key { float ax, ay, bx, by; bool operator<(const Key& o) const { return std::tie(ax, ay, bx, by) < std::tie(o.ax, o.ay, o.bx, o.by); } };
    auto mk = [](float ax, float ay, float bx, float by) { return std::tie(ax, ay) < std::tie(bx, by) ? Key{ax, ay, bx, by} : Key{bx, by, ax, ay}; };
    std::map<Key, std::vector<int>> edges;  // undirected edge -> triangle*3 + edge index
    for (int t = 0; t < n; ++t) {
        const float* r = tris + (size_t)t * stride;
        for (int k = 0; k < 3; ++k) { int k1 = (k + 1) % 3; edges[mk(r[2 * k], r[2 * k + 1], r[2 * k1], r[2 * k1 + 1])].push_back(t * 3 + k); }
    }
    std::vector<char> used(n, 0);
    std::vector<float> out;
    auto convex = [](const double (&q)[4][2]) {
        int sgn = 0;
        for (int k = 0; k < 4; ++k) {
            const double* a = q[k]; const double* b = q[(k + 1) % 4]; const double* c = q[(k + 2) % 4];
            double ux = b[0] - a[0], uy = b[1] - a[1], vx = c[0] - b[0], vy = c[1] - b[1];
            double cr = ux * vy - uy * vx, lu = std::hypot(ux, uy), lv = std::hypot(vx, vy);
            if (lu < 1e-6 || lv < 1e-6) return false;
            double sn = cr / (lu * lv);  // sine of the turn: need a clear turn (interior angle <= ~162 deg)
            if (std::fabs(sn) < 0.3) return false;
            int s = sn > 0 ? 1 : -1;
            if (sgn == 0) sgn = s; else if (s != sgn) return false;
        }
        return true;
    };
    for (int t = 0; t < n; ++t) {
        if (used[t]) continue;
        const float* r = tris + (size_t)t * stride;
        bool merged = false;
        for (int k = 0; k < 3 && !merged; ++k) {
            int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
            for (int ref : edges[mk(r[2 * k], r[2 * k + 1], r[2 * k1], r[2 * k1 + 1])]) {
                int u = ref / 3, ue = ref % 3;
                if (u == t || used[u]) continue;
                const float* w = tris + (size_t)u * stride;
                int uo = (ue + 2) % 3;  // vertex of u opposite the shared edge
                // quad: t's opposite vertex, shared a, u's opposite vertex, shared b
                double q[4][2] = {{r[2 * k2], r[2 * k2 + 1]}, {r[2 * k], r[2 * k + 1]}, {w[2 * uo], w[2 * uo + 1]}, {r[2 * k1], r[2 * k1 + 1]}};
                if (!convex(q)) continue;
                // the far diagonal of a merged pair is not an edge of either triangle: keep the quad within the extent the
                // rasteriser's fixed-point range is sized for (the device skips the oracle's coordinate clamp on that basis)
                if (std::hypot(q[0][0] - q[2][0], q[0][1] - q[2][1]) > max_extent) continue;
                for (int v = 0; v < 4; ++v) { out.push_back((float)q[v][0]); out.push_back((float)q[v][1]); }
                used[t] = used[u] = 1; merged = true;
                break;
            }
        }
        if (!merged) {
            used[t] = 1;
            for (int v = 0; v < 3; ++v) { out.push_back(r[2 * v]); out.push_back(r[2 * v + 1]); }
            out.push_back(r[0]); out.push_back(r[1]);
        }
    }
    return out;
}

// Tile index over the static render primitives (both layers).  A primitive belongs to the tile of its
// bbox min corner; the device finds the tiles of [ego - reach - maxext, ego + reach] and takes, per tile
// row, one contiguous run.  Tile indices use the same binary32 expression on both sides.
struct StaticIndex {
    std::vector<float> prims;      // 8 floats each: oversized first, then tile-sorted
    std::vector<uint8_t> cls;
    std::vector<int> tile_start;   // [ny*nx + 1], indices into prims
    int n_big = 0, nx = 1, ny = 1;
    float gx0 = 0, gy0 = 0, inv = 1, maxext = 0;
};
StaticIndex build_static_index(const std::vector<float>& road, const std::vector<float>& mark, float reach) {
    StaticIndex si;
    struct Item { const float* v; uint8_t cls; float lox, loy, ext; int tile; };
    std::vector<Item> items;
    auto add = [&](const std::vector<float>& src, uint8_t cls) {
        for (size_t k = 0; k + 8 <= src.size(); k += 8) {
            const float* v = &src[k];
            float lx = std::min(std::min(v[0], v[2]), std::min(v[4], v[6])), hx = std::max(std::max(v[0], v[2]), std::max(v[4], v[6]));
            float ly = std::min(std::min(v[1], v[3]), std::min(v[5], v[7])), hy = std::max(std::max(v[1], v[3]), std::max(v[5], v[7]));
            items.push_back(Item{v, cls, lx, ly, std::max(hx - lx, hy - ly), -1});
        }
    };
    add(road, TDE_CLS_ROAD);
    add(mark, TDE_CLS_LANE_MARKING);
    if (items.empty()) { si.tile_start.assign(2, 0); return si; }
    // tile size: 4 m unless the viewport is large (measured at fov 35: 8 / 6 / 4 / 3 / 2 m -> render 106.8 / 105.7 / 105.4 /
    // 104.9 / 105.0 us); primitives longer than 16 m (and than two tiles) stay out of the tiles.  TDE_TILE_M overrides.
    double T0 = 4.0;
    if (const char* v = std::getenv("TDE_TILE_M")) T0 = std::max(0.5, std::atof(v));
    double T = std::max(T0, (2.0 * reach + 16.0) / 20.0 * (T0 / 8.0));
    for (int attempt = 0; attempt < 8; ++attempt) {
        const float big = (float)std::max(2.0 * T, 16.0);
        float lox = INFINITY, loy = INFINITY, hix = -INFINITY, hiy = -INFINITY, maxext = 0.f;
        for (auto& it : items)
            if (it.ext <= big) { lox = std::min(lox, it.lox); loy = std::min(loy, it.loy); hix = std::max(hix, it.lox); hiy = std::max(hiy, it.loy); maxext = std::max(maxext, it.ext); }
        if (!(lox <= hix)) { lox = loy = hix = hiy = 0.f; }
        si.gx0 = lox; si.gy0 = loy; si.inv = (float)(1.0 / T);
        si.maxext = maxext * 1.0001f + 1e-3f;
        si.nx = (int)std::floor((hix - si.gx0) * si.inv) + 1;
        si.ny = (int)std::floor((hiy - si.gy0) * si.inv) + 1;
        double rows = (2.0 * reach + si.maxext) / T + 2.0;
        if ((double)si.nx * si.ny <= 4.0e6 && rows <= 24.0) break;
        T *= 1.5;
    }
    const float big = (float)std::max(2.0 / (double)si.inv, 16.0);
    for (auto& it : items) {
        if (it.ext > big) { it.tile = -1; continue; }
        int tx = (int)std::floor((it.lox - si.gx0) * si.inv), ty = (int)std::floor((it.loy - si.gy0) * si.inv);
        tx = std::min(std::max(tx, 0), si.nx - 1); ty = std::min(std::max(ty, 0), si.ny - 1);
        it.tile = ty * si.nx + tx;
    }
    std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.tile < b.tile; });
    si.tile_start.assign((size_t)si.nx * si.ny + 1, 0);
    for (auto& it : items) {
        si.prims.insert(si.prims.end(), it.v, it.v + 8);
        si.cls.push_back(it.cls);
        if (it.tile < 0) si.n_big++; else si.tile_start[(size_t)it.tile + 1]++;
    }
    si.tile_start[0] = si.n_big;
    for (size_t k = 1; k < si.tile_start.size(); ++k) si.tile_start[k] += si.tile_start[k - 1];
    return si;
}

std::vector<float> morton_sorted(const float* tris, int n, int stride) {
    std::vector<float> out((size_t)n * stride);
    if (n == 0) return out;
    double lox = 1e300, loy = 1e300, hix = -1e300, hiy = -1e300;
    std::vector<P2> cen(n);
    for (int t = 0; t < n; ++t) {
        const float* r = tris + (size_t)t * stride;
        cen[t] = P2{(r[0] + r[2] + r[4]) / 3.0, (r[1] + r[3] + r[5]) / 3.0};
        lox = std::min(lox, cen[t].x); hix = std::max(hix, cen[t].x); loy = std::min(loy, cen[t].y); hiy = std::max(hiy, cen[t].y);
    }
    double inv = 65535.0 / std::max(1e-9, std::max(hix - lox, hiy - loy));
    std::vector<std::pair<uint32_t, int>> key(n);
    for (int t = 0; t < n; ++t) key[t] = {morton_key(cen[t].x, cen[t].y, lox, loy, inv), t};
    std::stable_sort(key.begin(), key.end());
    for (int k = 0; k < n; ++k) std::memcpy(&out[(size_t)k * stride], tris + (size_t)key[k].second * stride, sizeof(float) * stride);
    return out;
}
}  // namespace

// ------------------------------------------------------------------ handle

template <typename T>
static int dev_alloc(tde_handle* h, T** p, size_t n) {
    CUDA_TRY(h, cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    CUDA_TRY(h, cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return TDE_OK;
}
template <typename T>
static int dev_upload(tde_handle* h, T** p, const T* src, size_t n) {
    int rc = dev_alloc(h, p, n);
    if (rc) return rc;
    if (n) CUDA_TRY(h, cudaMemcpy(*p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    h->tab->scenario_allocs.push_back((void*)*p);
    return TDE_OK;
}

extern "C" int tde_version(void) { return TDE_VERSION; }

extern "C" const char* tde_last_error(const tde_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

extern "C" int tde_default_config(tde_config* c) {
    if (!c) return TDE_E_INVAL;
    std::memset(c, 0, sizeof(*c));
    c->num_envs = 1; c->max_agents = 1; c->env_index_offset = 0;
    c->max_environment_steps = 200; c->terminated_at_infraction = 1; c->left_handed_coordinates = 1;
    c->auto_reset = 0; c->randomize_ego_attributes = 0; c->device = 0;
    c->dt = 0.1f; c->waypoint_bonus = 100.f; c->heading_penalty = 25.f; c->distance_bonus = 1.f;
    c->distance_cutoff = 0.5f; c->reach_radius = 3.f; c->offroad_threshold = 0.5f; c->tl_rear_factor = 0.1f;
    c->fov = 35.f; c->start_speed_max = 10.f; c->start_heading_sigma = 0.1f;
    return TDE_OK;
}

static void free_scenarios(tde_handle* h) {
    h->tab.reset();   // the tables go when their last handle lets go
    h->uploaded = false;
}

// persistent grids: a multiple of the SM count (resident blocks per SM from the occupancy calculator)
template <int AH>
static int configure_render(tde_handle* h) {
    size_t smem = TDE_RENDER_SMEM_BYTES;   // groups + warps + the spread and span-mask tables
    CUDA_TRY(h, cudaFuncSetAttribute(tde_render_kernel<AH, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(h, cudaFuncSetAttribute(tde_render_kernel<AH, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(h, cudaFuncSetAttribute(tde_render_kernel<AH, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(h, cudaFuncSetAttribute(tde_render_kernel<AH, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tde_render_kernel<AH, 0>, TDE_WARPS_PER_BLOCK * 32, smem));
    if (per_sm < 1) per_sm = 1;
    // tuning knobs for co-residency experiments (tools/coresident.py): cap the resident blocks per SM of either kernel
    if (const char* v = std::getenv("TDE_RENDER_BLOCKS_CAP")) per_sm = std::max(1, std::min(per_sm, std::atoi(v)));
    h->smem_render = smem;
    h->grid_render = std::max(1, std::min((h->E + TDE_RENDER_GROUPS - 1) / TDE_RENDER_GROUPS, per_sm * h->sm_count));
    return TDE_OK;
}

// The physics launch, chosen once the scenario tables are known (end of tde_upload_scenarios).
//  * STAGED: the map tables fit beside the SAT scratch -> every CTA copies them into shared memory with bulk-async copies
//    (north-star item 3).  Fat CTAs (up to 32 warps, one per SM) so the copy is made once per SM; fewer warps per CTA for
//    small batches so that every SM still gets a CTA.
//  * otherwise 4-warp CTAs that read the tables through L1: ask for no more shared memory than the resident blocks need,
//    the rest of the SM's 228 KB stays L1 (round 1: 54.3 -> 52.5 us at 30 %, 57.7 us at 60 %, 73.3 us at 100 %).
template <int AH>
static int configure_physics(tde_handle* h) {
    const size_t scratch = sizeof(SatScratch<AH>);
    const size_t smem_max = 227 * 1024;
    if (h->tab->stage_blob) {
        int wpb = TDE_PHYS_STAGED_MAX_WARPS;
        while (wpb > 4 && (h->E + wpb - 1) / wpb < h->sm_count) wpb = std::max(4, wpb / 2);
        if (const char* v = std::getenv("TDE_PHYS_WARPS")) wpb = std::max(1, std::min(TDE_PHYS_STAGED_MAX_WARPS, std::atoi(v)));
        const size_t smem = (size_t)h->tab->stage_bytes + 16 + (size_t)wpb * scratch;
        if (smem <= smem_max) {
            CUDA_TRY(h, cudaFuncSetAttribute(tde_physics_kernel<AH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            int per_sm = 0;
            CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tde_physics_kernel<AH, true>, wpb * 32, smem));
            if (per_sm >= 1) {
                h->phys_wpb = wpb; h->phys_per_sm = per_sm; h->smem_phys = smem;
                h->grid_phys = std::max(1, std::min((h->E + wpb - 1) / wpb, per_sm * h->sm_count));
                return TDE_OK;
            }
        }
        cudaFree(h->tab->stage_blob); h->tab->stage_blob = nullptr; h->tab->stage_bytes = 0;   // does not fit: global path
    }
    const int wpb = TDE_WARPS_PER_BLOCK;
    const size_t smem = 16 + (size_t)wpb * scratch;
    int per_sm = 0;
    CUDA_TRY(h, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tde_physics_kernel<AH, false>, wpb * 32, smem));
    if (per_sm < 1) per_sm = 1;
    {
        const size_t need = (size_t)per_sm * (smem + 1024);
        int pct = (int)((need * 100 + 233471) / 233472) + 3;
        if (const char* v = std::getenv("TDE_PHYS_CARVEOUT")) pct = std::atoi(v);
        CUDA_TRY(h, cudaFuncSetAttribute(tde_physics_kernel<AH, false>, cudaFuncAttributePreferredSharedMemoryCarveout, std::max(0, std::min(pct, 100))));
    }
    if (const char* v = std::getenv("TDE_PHYS_BLOCKS_CAP")) per_sm = std::max(1, std::min(per_sm, std::atoi(v)));
    h->phys_wpb = wpb; h->phys_per_sm = per_sm; h->smem_phys = smem;
    h->grid_phys = std::max(1, std::min((h->E + wpb - 1) / wpb, per_sm * h->sm_count));
    return TDE_OK;
}

extern "C" int tde_create(const tde_config* cfg, tde_handle** out) {
    if (!cfg || !out) return fail(nullptr, TDE_E_INVAL, "tde_create: null argument");
    if (cfg->num_envs < 1) return fail(nullptr, TDE_E_INVAL, "tde_create: num_envs must be >= 1");
    if (cfg->max_agents < 1 || cfg->max_agents > TDE_MAX_AGENTS)
        return fail(nullptr, TDE_E_SHAPE, "tde_create: max_agents must be in [1, 64]");
    if (!(cfg->fov > 0.f) || !(cfg->dt > 0.f)) return fail(nullptr, TDE_E_INVAL, "tde_create: fov and dt must be > 0");
    if (cfg->max_environment_steps < 1) return fail(nullptr, TDE_E_INVAL, "tde_create: max_environment_steps must be >= 1");
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev == 0)
        return fail(nullptr, TDE_E_CUDA, std::string("tde_create: no CUDA device (") + cudaGetErrorString(ce) + "); there is no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(nullptr, TDE_E_INVAL, "tde_create: bad device ordinal");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(nullptr, TDE_E_CUDA, "tde_create: cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(nullptr, TDE_E_ARCH, "tde_create: device is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                             ", this library is built for sm_100a only");
    tde_handle* h = new tde_handle();
    h->cfg = *cfg; h->E = cfg->num_envs; h->A = cfg->max_agents; h->device = cfg->device;
    h->sm_count = prop.multiProcessorCount;
    std::memcpy(h->palette, k_default_palette, sizeof(k_default_palette));
    int rc = TDE_OK;
    auto bail = [&](int code) { g_create_error = h->err; tde_destroy(h); return code; };
    DeviceGuard guard(h->device);
    if (!guard.ok) return bail(fail(h, TDE_E_CUDA, "cudaSetDevice failed"));
    size_t EA = (size_t)h->E * h->A;
    if ((rc = dev_alloc(h, &h->state, EA)) || (rc = dev_alloc(h, &h->attr, EA)) || (rc = dev_alloc(h, &h->infr, EA)) ||
        (rc = dev_alloc(h, &h->vars, (size_t)h->E * 8)) || (rc = dev_alloc(h, &h->ep_return, (size_t)h->E)) ||
        (rc = dev_alloc(h, &h->scen_lo, (size_t)h->E)) || (rc = dev_alloc(h, &h->scen_hi, (size_t)h->E)) ||
        (rc = dev_alloc(h, &h->stats, (size_t)TDE_NUM_STATS)) || (rc = dev_alloc(h, &h->restart, (size_t)h->E)) ||
        (rc = dev_alloc(h, &h->tickets, (size_t)4)) || (rc = dev_alloc(h, &h->done_mask, (size_t)h->E)))
        return bail(rc);
    rc = h->A <= 32 ? configure_render<1>(h) : configure_render<2>(h);
    if (rc) return bail(rc);
    if (const char* v = std::getenv("TDE_PDL_PHYSICS")) h->pdl_physics = std::atoi(v) != 0;
    *out = h;
    return TDE_OK;
}

extern "C" int tde_destroy(tde_handle* h) {
    if (!h) return TDE_OK;
    DeviceGuard guard(h->device);
    free_scenarios(h);
    cudaFree(h->state); cudaFree(h->attr); cudaFree(h->infr); cudaFree(h->vars); cudaFree(h->ep_return);
    cudaFree(h->scen_lo); cudaFree(h->scen_hi); cudaFree(h->stats); cudaFree(h->restart); cudaFree(h->tickets); cudaFree(h->done_mask);
    cudaFree(h->view_prims);
    if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
    for (cudaEvent_t ev : h->chunk_done) if (ev) cudaEventDestroy(ev);
    if (h->copies_done) cudaEventDestroy(h->copies_done);
    for (cudaEvent_t ev : h->chunk_copied) if (ev) cudaEventDestroy(ev);
    delete_pool(h->pool);
    cudaFree(h->h_nib); if (h->pin_nib) cudaFreeHost(h->pin_nib);
    cudaFree(h->h_actions); cudaFree(h->h_obs); cudaFree(h->h_reward); cudaFree(h->h_term); cudaFree(h->h_trunc); cudaFree(h->h_info);
    delete h;
    return TDE_OK;
}

extern "C" int tde_set_palette(tde_handle* h, const uint8_t* rgb) {
    if (!h || !rgb) return fail(h, TDE_E_INVAL, "tde_set_palette: null argument");
    std::memcpy(h->palette, rgb, TDE_NUM_CLASSES * 3);
    return TDE_OK;
}

extern "C" int tde_upload_scenarios(tde_handle* h, const tde_scenario_set* s) {
    if (!h || !s) return fail(h, TDE_E_INVAL, "tde_upload_scenarios: null argument");
    if (s->num_maps < 1 || s->num_scenarios < 1) return fail(h, TDE_E_INVAL, "tde_upload_scenarios: need >= 1 map and >= 1 scenario");
    TDE_ON_DEVICE(h);
    free_scenarios(h);
    h->tab = std::make_shared<ScenTables>();
    h->tab->device = h->device;
    const int A = h->A;
    const float ppm = (float)TDE_OBS_W / h->cfg.fov;
    const double max_edge_m = 430.0 / (double)ppm;  // keeps snapped vertices of drawn primitives inside +-511 px
    h->tab->maps_host.resize(s->num_maps);
    // pieces of the staged blob (physics kernel): per map the device triangle / stop-line records, the light schedule and
    // the per-cell summary (u16: SAFE flag | the overlapping triangle covering most of the cell)
    struct StagePiece { const float4* tri; int ntri; const float4* stop; int nstop; const uint8_t* lights; size_t nlights; std::vector<uint16_t> cells; };
    std::vector<StagePiece> pieces((size_t)s->num_maps);
    bool stage_ok = true;
    for (int m = 0; m < s->num_maps; ++m) {
        MapDev& M = h->tab->maps_host[m];
        std::memset(&M, 0, sizeof(M));
        int t0 = s->map_tri_offset[m], nt = s->map_tri_offset[m + 1] - t0;
        int k0 = s->map_mark_offset[m], nk = s->map_mark_offset[m + 1] - k0;
        int l0 = s->map_stop_offset[m], nl = s->map_stop_offset[m + 1] - l0;
        if (nt < 0 || nk < 0 || nl < 0) return fail(h, TDE_E_INVAL, "tde_upload_scenarios: offsets must be non-decreasing");
        if (nt > 65535) return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: more than 65535 road triangles in one map");
        if (nl > TDE_MAX_STOPLINES) return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: more than 32 stop lines in one map");
        auto edge_ok = [&](const float* v, int stride_pts) {
            for (int k = 0; k < 3; ++k) {
                int k1 = (k + 1) % 3;
                double d = std::hypot((double)v[2 * k] - v[2 * k1], (double)v[2 * k + 1] - v[2 * k1 + 1]);
                if (!(d <= max_edge_m)) return false;
            }
            (void)stride_pts;
            return true;
        };
        for (int t = 0; t < nt; ++t)
            if (!edge_ok(s->road_tris + 8 * (size_t)(t0 + t), 0))
                return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: road triangle edge too long for the rasteriser's fixed-point range; subdivide the mesh");
        for (int t = 0; t < nk; ++t)
            if (!edge_ok(s->mark_tris + 6 * (size_t)(k0 + t), 0))
                return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: marking triangle edge too long; subdivide the mesh");
        // static triangles in Morton order: runs of 32 are spatially compact (render-time culling); the
        // order within a layer does not change any result
        std::vector<float> road_sorted = morton_sorted(s->road_tris + 8 * (size_t)t0, nt, 8);
        // road triangles -> device records
        float* raw = nullptr; float4* rec = nullptr;
        int rc;
        if ((rc = dev_upload(h, &raw, road_sorted.data(), (size_t)nt * 8))) return rc;
        if ((rc = dev_alloc(h, &rec, (size_t)nt * 3))) return rc;
        h->tab->scenario_allocs.push_back(rec);
        if (nt) TDE_LAUNCH((nt + 127) / 128, 128, 0, 0, prep_tris_kernel)(raw, nt, rec);
        M.tri = rec; M.ntri = nt;
        M.nmark = nk;
        float* mraw = nullptr;
        if ((rc = dev_upload(h, &mraw, s->mark_tris + 6 * (size_t)k0, (size_t)nk * 6))) return rc;
        M.mark_raw = mraw;
        // render-only static primitives: merged quads / leftover triangles of both layers, indexed by tile
        {
            std::vector<float> rp_road = merge_into_quads(s->road_tris + 8 * (size_t)t0, nt, 8, max_edge_m);
            std::vector<float> rp_mark = merge_into_quads(s->mark_tris + 6 * (size_t)k0, nk, 6, max_edge_m);
            const float reach = (0.70710678f * (float)(TDE_OBS_W + TDE_OBS_H) * 0.5f + 2.0f) / ppm;  // as in tde_render_kernel
            StaticIndex si = build_static_index(rp_road, rp_mark, reach);
            float* rpd = nullptr; uint8_t* clsd = nullptr; int* tsd = nullptr;
            if ((rc = dev_upload(h, &rpd, si.prims.data(), si.prims.size()))) return rc;
            if ((rc = dev_upload(h, &clsd, si.cls.data(), si.cls.size()))) return rc;
            if ((rc = dev_upload(h, &tsd, si.tile_start.data(), si.tile_start.size()))) return rc;
            M.rp = (const float4*)rpd; M.rp_cls = clsd; M.tile_start = tsd;
            M.n_rp = (int)si.cls.size(); M.n_big = si.n_big;
            M.tgx0 = si.gx0; M.tgy0 = si.gy0; M.tinv = si.inv; M.maxext = si.maxext; M.tnx = si.nx; M.tny = si.ny;
        }
        for (int l = 0; l < nl; ++l) {
            const float* sl = s->stoplines + 5 * (size_t)(l0 + l);
            if (!(std::hypot((double)sl[2], (double)sl[3]) <= max_edge_m))
                return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: stop line larger than the rasteriser's fixed-point range (430 px)");
        }
        float* sraw = nullptr; float4* srec = nullptr;
        if ((rc = dev_upload(h, &sraw, s->stoplines + 5 * (size_t)l0, (size_t)nl * 5))) return rc;
        if ((rc = dev_alloc(h, &srec, (size_t)nl * 2))) return rc;
        h->tab->scenario_allocs.push_back(srec);
        if (nl) TDE_LAUNCH(1, 64, 0, 0, prep_stops_kernel)(sraw, nl, srec);
        M.stop = srec; M.nstop = nl;
        int P = s->map_light_period[m];
        int lo = s->map_light_offset[m], ln = s->map_light_offset[m + 1] - lo;
        if (nl > 0 && (P < 1 || ln != P * nl)) return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: light schedule must hold period x stoplines states");
        uint8_t* lights = nullptr;
        if ((rc = dev_upload(h, &lights, s->light_states + lo, (size_t)std::max(ln, 0)))) return rc;
        M.lights = lights; M.period = nl > 0 ? P : 0;
        Grid g = build_grid(road_sorted.data(), nt, (double)h->cfg.offroad_threshold);
        std::vector<int2> crec(g.meta.size());
        for (size_t c = 0; c < g.meta.size(); ++c) {
            int cnt = g.cell_start[c + 1] - g.cell_start[c];
            if (cnt > 65535) return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: more than 65535 candidate triangles in one grid cell");
            crec[c] = make_int2(g.cell_start[c], (int)(((unsigned)cnt << 16) | g.meta[c]));
        }
        int2* crd = nullptr; uint16_t* items = nullptr;
        if ((rc = dev_upload(h, &crd, crec.data(), crec.size()))) return rc;
        if ((rc = dev_upload(h, &items, g.items.data(), g.items.size()))) return rc;
        M.cell_rec = crd; M.cell_items = items;
        {
            size_t safe = 0, nover = 0;
            for (uint16_t m : g.meta) { safe += (m & TDE_CELL_SAFE) ? 1 : 0; nover += m & 0x7fff; }
            h->tab->map_info.push_back({nt, nk, nl, g.nx, g.ny, (int)g.items.size(), (int)safe, M.n_rp});
        }
        M.gx0 = g.gx0; M.gy0 = g.gy0; M.inv_cell = g.inv_cell; M.gnx = g.nx; M.gny = g.ny;
        {
            StagePiece& sp = pieces[(size_t)m];
            sp.tri = rec; sp.ntri = nt; sp.stop = srec; sp.nstop = nl;
            sp.lights = s->light_states + lo; sp.nlights = (size_t)std::max(ln, 0);
            sp.cells.resize(g.meta.size());
            if (nt >= TDE_CELL16_NONE) stage_ok = false;   // triangle ids must fit 15 bits
            for (size_t c = 0; c < g.meta.size(); ++c) {
                const int nover = g.meta[c] & 0x7fff;
                const uint16_t first = nover > 0 ? g.items[(size_t)g.cell_start[c]] : (uint16_t)TDE_CELL16_NONE;
                sp.cells[c] = (uint16_t)((g.meta[c] & TDE_CELL_SAFE) | (first & 0x7fff));
            }
        }
    }
    h->tab->scens_host.resize(s->num_scenarios);
    for (int k = 0; k < s->num_scenarios; ++k) {
        ScenDev& S = h->tab->scens_host[k];
        std::memset(&S, 0, sizeof(S));
        S.map = s->scen_map[k];
        if (S.map < 0 || S.map >= s->num_maps) return fail(h, TDE_E_INVAL, "tde_upload_scenarios: scen_map out of range");
        int w0 = s->scen_wp_offset[k];
        S.W = s->scen_wp_offset[k + 1] - w0;
        if (S.W < 1) return fail(h, TDE_E_INVAL, "tde_upload_scenarios: a scenario needs >= 1 waypoint");
        S.nag = s->scen_num_agents[k];
        if (S.nag < 1 || S.nag > A) return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: scen_num_agents must be in [1, max_agents]");
        S.start_heading = s->scen_start_heading[k];
        int rc;
        float2* wp = nullptr;
        if ((rc = dev_upload(h, (float**)&wp, s->waypoints + 2 * (size_t)w0, (size_t)S.W * 2))) return rc;
        S.wp = wp;
        float4* init = nullptr;
        if ((rc = dev_upload(h, (float**)&init, s->agent_init + (size_t)k * A * 4, (size_t)A * 4))) return rc;
        S.init = init;
        std::vector<float> at4((size_t)A * 4, 0.f);
        for (int a = 0; a < A; ++a)
            for (int q = 0; q < 3; ++q) at4[4 * a + q] = s->agent_attr[((size_t)k * A + a) * 3 + q];
        float4* attr = nullptr;
        if ((rc = dev_upload(h, (float**)&attr, at4.data(), at4.size()))) return rc;
        S.attr = attr;
        S.rep_T = s->scen_replay_T[k];
        int r0 = s->scen_replay_offset[k];
        if (S.rep_T < 0 || s->scen_replay_offset[k + 1] - r0 != S.rep_T) return fail(h, TDE_E_SHAPE, "tde_upload_scenarios: replay offsets do not match scen_replay_T");
        float4* rs = nullptr; uint8_t* rm = nullptr;
        if ((rc = dev_upload(h, (float**)&rs, s->replay_states + (size_t)r0 * A * 4, (size_t)S.rep_T * A * 4))) return rc;
        if ((rc = dev_upload(h, &rm, s->replay_mask + (size_t)r0 * A, (size_t)S.rep_T * A))) return rc;
        S.rep_states = rs; S.rep_mask = rm;
    }
    h->tab->num_maps = s->num_maps; h->tab->num_scen = s->num_scenarios;
    CUDA_TRY(h, cudaMalloc((void**)&h->tab->maps_dev, sizeof(MapDev) * h->tab->num_maps));
    CUDA_TRY(h, cudaMemcpy(h->tab->maps_dev, h->tab->maps_host.data(), sizeof(MapDev) * h->tab->num_maps, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMalloc((void**)&h->tab->scens_dev, sizeof(ScenDev) * h->tab->num_scen));
    CUDA_TRY(h, cudaMemcpy(h->tab->scens_dev, h->tab->scens_host.data(), sizeof(ScenDev) * h->tab->num_scen, cudaMemcpyHostToDevice));
    std::vector<int> lo((size_t)h->E, 0), hi((size_t)h->E, h->tab->num_scen);
    CUDA_TRY(h, cudaMemcpy(h->scen_lo, lo.data(), sizeof(int) * h->E, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->scen_hi, hi.data(), sizeof(int) * h->E, cudaMemcpyHostToDevice));
    // the staged blob: [16 B header per map] [MapDev per map] [per map: triangle records, stop lines, lights, cell summary]
    // Measured at C3 (one map, 85 KB of tables): the staged launch is SLOWER than the one that reads the tables through L1
    // (57.4 vs 53.7 us at 64 registers): after the first wave of envs the tables are L1-resident
    // anyway.  A small batch never gets there - at most 8 envs per SM, every launch starts with a cold L1 and the table walks
    // of a lone warp go to L2 one after the other - and there staging wins (C2: 11.5 vs 12.0 us per step).  So:
    // cfg.stage_map_tables 0 = staged for handles of at most 8 envs per SM, 1 = always (when the tables fit), 2 = never;
    // TDE_PHYS_STAGE=0|1 overrides.
    {
        const char* v = std::getenv("TDE_PHYS_STAGE");
        const bool small_batch = h->E <= 8 * h->sm_count;
        const bool want = v ? std::atoi(v) != 0 : h->cfg.stage_map_tables == 1 || (h->cfg.stage_map_tables == 0 && small_batch);
        stage_ok = stage_ok && want;
    }
    if (stage_ok) {
        auto al16 = [](size_t x) { return (x + 15) & ~(size_t)15; };
        std::vector<uint32_t> hdr((size_t)h->tab->num_maps * 4);
        const size_t maps_off = al16((size_t)h->tab->num_maps * 16);
        size_t cur = al16(maps_off + sizeof(MapDev) * h->tab->num_maps);
        for (int m = 0; m < h->tab->num_maps; ++m) {
            const StagePiece& sp = pieces[(size_t)m];
            hdr[4 * m + 0] = (uint32_t)cur; cur = al16(cur + (size_t)sp.ntri * 48);
            hdr[4 * m + 1] = (uint32_t)cur; cur = al16(cur + (size_t)sp.nstop * 32);
            hdr[4 * m + 2] = (uint32_t)cur; cur = al16(cur + sp.nlights);
            hdr[4 * m + 3] = (uint32_t)cur; cur = al16(cur + sp.cells.size() * 2);
        }
        const size_t scratch_min = 16 + 4 * sizeof(SatScratch<1>);
        if (cur + scratch_min <= 227 * 1024) {
            CUDA_TRY(h, cudaMalloc((void**)&h->tab->stage_blob, cur));
            CUDA_TRY(h, cudaMemset(h->tab->stage_blob, 0, cur));
            CUDA_TRY(h, cudaMemcpy(h->tab->stage_blob, hdr.data(), hdr.size() * 4, cudaMemcpyHostToDevice));
            CUDA_TRY(h, cudaMemcpy(h->tab->stage_blob + maps_off, h->tab->maps_host.data(), sizeof(MapDev) * h->tab->num_maps, cudaMemcpyHostToDevice));
            for (int m = 0; m < h->tab->num_maps; ++m) {
                const StagePiece& sp = pieces[(size_t)m];
                if (sp.ntri) CUDA_TRY(h, cudaMemcpy(h->tab->stage_blob + hdr[4 * m + 0], sp.tri, (size_t)sp.ntri * 48, cudaMemcpyDeviceToDevice));
                if (sp.nstop) CUDA_TRY(h, cudaMemcpy(h->tab->stage_blob + hdr[4 * m + 1], sp.stop, (size_t)sp.nstop * 32, cudaMemcpyDeviceToDevice));
                if (sp.nlights) CUDA_TRY(h, cudaMemcpy(h->tab->stage_blob + hdr[4 * m + 2], sp.lights, sp.nlights, cudaMemcpyHostToDevice));
                if (!sp.cells.empty()) CUDA_TRY(h, cudaMemcpy(h->tab->stage_blob + hdr[4 * m + 3], sp.cells.data(), sp.cells.size() * 2, cudaMemcpyHostToDevice));
            }
            h->tab->stage_bytes = (unsigned int)cur; h->tab->stage_maps_off = (unsigned int)maps_off;
        }
    }
    {
        int rc = h->A <= 32 ? configure_physics<1>(h) : configure_physics<2>(h);
        if (rc) return rc;
    }
    CUDA_TRY(h, cudaDeviceSynchronize());
    h->uploaded = true; h->was_reset = false;
    return TDE_OK;
}

extern "C" int tde_set_env_scenario_range(tde_handle* h, const int32_t* lo, const int32_t* hi) {
    if (!h || !lo || !hi) return fail(h, TDE_E_INVAL, "tde_set_env_scenario_range: null argument");
    if (!h->uploaded) return fail(h, TDE_E_STATE, "tde_set_env_scenario_range: upload scenarios first");
    for (int e = 0; e < h->E; ++e)
        if (lo[e] < 0 || hi[e] > h->tab->num_scen || lo[e] >= hi[e]) return fail(h, TDE_E_INVAL, "tde_set_env_scenario_range: need 0 <= lo < hi <= num_scenarios");
    TDE_ON_DEVICE(h);
    CUDA_TRY(h, cudaMemcpy(h->scen_lo, lo, sizeof(int) * h->E, cudaMemcpyHostToDevice));
    CUDA_TRY(h, cudaMemcpy(h->scen_hi, hi, sizeof(int) * h->E, cudaMemcpyHostToDevice));
    return TDE_OK;
}

static StepParams make_params(tde_handle* h) {
    StepParams p;
    std::memset(&p, 0, sizeof(p));
    p.cfg = h->cfg; p.E = h->E; p.A = h->A; p.num_scen = h->tab->num_scen; p.seed = h->seed;
    p.maps = h->tab->maps_dev; p.scens = h->tab->scens_dev;
    p.state = h->state; p.attr = h->attr; p.infr = h->infr; p.vars = h->vars; p.ep_return = h->ep_return;
    p.scen_lo = h->scen_lo; p.scen_hi = h->scen_hi; p.stats = h->stats; p.restart = h->restart; p.tickets = h->tickets; p.n_stack = 1;
    p.e_begin = 0; p.e_end = h->E;
    p.stage_blob = h->tab->stage_blob; p.stage_bytes = h->tab->stage_bytes; p.stage_maps_off = h->tab->stage_maps_off;
    for (int ch = 0; ch < 3; ++ch)
        for (int w = 0; w < 4; ++w) {
            uint32_t v = 0;
            for (int b = 0; b < 4; ++b) {
                int cls = 4 * w + b;
                uint32_t byte = cls < TDE_NUM_CLASSES ? h->palette[3 * cls + ch] : 0;
                v |= byte << (8 * b);
            }
            p.pal[ch][w] = v;
        }
    p.ppm = (float)TDE_OBS_W / h->cfg.fov;
    p.ppmy = h->cfg.left_handed_coordinates ? p.ppm : -p.ppm;
    return p;
}

extern "C" int tde_reset(tde_handle* h, const uint8_t* env_mask_dev, uint64_t seed, void* stream) {
    if (!h) return TDE_E_INVAL;
    if (!h->uploaded) return fail(h, TDE_E_STATE, "tde_reset: upload scenarios first");
    TDE_ON_DEVICE(h);
    // the seed belongs to the handle: a full reset sets it, a masked reset keeps it (the in-kernel auto-resets of the
    // other envs must not change their draw stream because a few envs were reset by hand)
    if (env_mask_dev == nullptr) h->seed = seed;
    StepParams p = make_params(h);
    p.reset_mask = env_mask_dev;
    cudaStream_t st = (cudaStream_t)stream;
    int grid = std::max(1, std::min((h->E + TDE_WARPS_PER_BLOCK - 1) / TDE_WARPS_PER_BLOCK, h->sm_count * 8));
    if (h->A <= 32) TDE_LAUNCH(grid, TDE_WARPS_PER_BLOCK * 32, 0, st, tde_reset_kernel<1>)(p);
    else TDE_LAUNCH(grid, TDE_WARPS_PER_BLOCK * 32, 0, st, tde_reset_kernel<2>)(p);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    h->was_reset = true;
    return TDE_OK;
}

// pdl: launch with programmatic stream serialization (the kernel may be scheduled while the previous kernel on the stream
// drains; it waits for that kernel's completion itself, tde_pdl_wait)
template <typename Kernel>
static cudaError_t launch_step(Kernel k, int grid, int threads, size_t smem, cudaStream_t st, const StepParams& p, bool pdl = false) {
#ifndef TDE_HOST_EMU
    if (pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, k, p);
    }
#endif
    TDE_LAUNCH(grid, threads, smem, st, k)(p);
    return cudaGetLastError();
}

static int step_impl(tde_handle* h, int32_t phases, const float* actions, uint8_t* obs, int32_t n_stack, float* reward,
                     uint8_t* terminated, uint8_t* truncated, float* info, void* stream, const uint8_t* obs_prev = nullptr,
                     uint8_t* terminal_obs = nullptr, int e_begin = 0, int e_end = -1, long long slot_stride = 0, int slots_ahead = 0,
                     bool class_nibbles = false, int ring_slots = 0, int ring_pos = 0) {
    if (!h) return TDE_E_INVAL;
    if (n_stack < 1 || n_stack > 8) return fail(h, TDE_E_INVAL, "n_stack must be in 1..8");
    if (!h->uploaded || !h->was_reset) return fail(h, TDE_E_STATE, "tde_step: upload scenarios and reset first");
    if ((phases & ~TDE_PH_ALL) || phases == 0) return fail(h, TDE_E_INVAL, "tde_step_phases: bad phase mask");
    if ((phases & TDE_PH_KINEMATICS) && !actions) return fail(h, TDE_E_INVAL, "tde_step: actions is null");
    if ((phases & TDE_PH_REWARD) && (!reward || !terminated || !truncated || !info))
        return fail(h, TDE_E_INVAL, "tde_step: reward/terminated/truncated/info must be non-null");
    if (phases == TDE_PH_RENDER && !obs) return fail(h, TDE_E_INVAL, "tde_render: obs is null");
    // the kernels move rows with 128-bit (obs, info) and 64-bit (actions) accesses
    if (((uintptr_t)obs | (uintptr_t)obs_prev | (uintptr_t)terminal_obs | (uintptr_t)info) & 15)
        return fail(h, TDE_E_INVAL, "tde_step: obs / terminal_obs / info must be 16-byte aligned");
    if ((uintptr_t)actions & 7) return fail(h, TDE_E_INVAL, "tde_step: actions must be 8-byte aligned");
    TDE_ON_DEVICE(h);
    StepParams p = make_params(h);
    p.phases = phases; p.actions = actions; p.obs = obs; p.reward = reward; p.terminated = terminated;
    p.truncated = truncated; p.info = info; p.n_stack = n_stack;
    p.obs_prev = obs_prev ? obs_prev : obs;
    p.slot_stride = slot_stride; p.slots_ahead = slots_ahead;
    for (int j = 0; j < 8; ++j) {   // copy j of the new frame: j slots ahead (modulo the ring size in a ring of slots), one channel group down
        const long long slots = ring_slots > 0 ? (long long)((ring_pos + j) % ring_slots) - ring_pos : (long long)j;
        p.copy_off[j] = slots * slot_stride - (long long)j * (TDE_OBS_C * TDE_OBS_H * TDE_OBS_W);
    }
    const bool scatter = slot_stride != 0;
    if (e_end >= 0) { p.e_begin = e_begin; p.e_end = e_end; }   // a slice of the envs (tde_step_host's chunks)
    cudaStream_t st = (cudaStream_t)stream;
    const int threads = TDE_WARPS_PER_BLOCK * 32;
    const bool physics = phases & (TDE_PH_KINEMATICS | TDE_PH_INFRACTIONS | TDE_PH_REWARD);
    const bool render = (phases & TDE_PH_RENDER) && obs;
    const int want = (p.e_end - p.e_begin + TDE_WARPS_PER_BLOCK - 1) / TDE_WARPS_PER_BLOCK;
    const int want_render = (p.e_end - p.e_begin + TDE_RENDER_GROUPS - 1) / TDE_RENDER_GROUPS;   // one env per warp group
    // terminal observations: finished envs are only flagged by the physics kernel; after the frame of their last
    // state is out it is copied to terminal_obs, the flagged envs are re-initialised and rendered again
    const bool deferred = terminal_obs && physics && render && h->cfg.auto_reset;
    if (deferred) {
        CUDA_TRY(h, cudaMemsetAsync(h->done_mask, 0, (size_t)h->E, st));
        p.done_mask = h->done_mask;
    }
    if (physics) {
        const int wpb = h->phys_wpb, pthreads = wpb * 32;
        const int grid = std::min(h->grid_phys, (p.e_end - p.e_begin + wpb - 1) / wpb);
        // a step without observations follows a physics kernel in a stepping loop: let it be scheduled while that one drains
        // (13.7 against 14.9 us per step at 1,024 envs x 16 agents; inside a CUDA graph the plain edge is faster: 12.0 against 12.4)
        bool pdl = !render && !deferred && h->pdl_physics;
        if (pdl) {
            cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(st, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) pdl = false;
        }
        if (h->tab->stage_blob) {
            if (h->A <= 32) CUDA_TRY(h, launch_step(tde_physics_kernel<1, true>, grid, pthreads, h->smem_phys, st, p, pdl));
            else CUDA_TRY(h, launch_step(tde_physics_kernel<2, true>, grid, pthreads, h->smem_phys, st, p, pdl));
        } else {
            if (h->A <= 32) CUDA_TRY(h, launch_step(tde_physics_kernel<1, false>, grid, pthreads, h->smem_phys, st, p, pdl));
            else CUDA_TRY(h, launch_step(tde_physics_kernel<2, false>, grid, pthreads, h->smem_phys, st, p, pdl));
        }
        h->launches++;
    }
    if (render) {
        const int grid = std::min(h->grid_render, want_render);
        if (class_nibbles) {   // obs is the 4-bit class image [E][64][32]
            if (h->A <= 32) CUDA_TRY(h, launch_step(tde_render_kernel<1, 3>, grid, threads, h->smem_render, st, p));
            else CUDA_TRY(h, launch_step(tde_render_kernel<2, 3>, grid, threads, h->smem_render, st, p));
        } else if (scatter) {
            if (h->A <= 32) CUDA_TRY(h, launch_step(tde_render_kernel<1, 2>, grid, threads, h->smem_render, st, p));
            else CUDA_TRY(h, launch_step(tde_render_kernel<2, 2>, grid, threads, h->smem_render, st, p));
        } else if (n_stack > 1) {
            if (h->A <= 32) CUDA_TRY(h, launch_step(tde_render_kernel<1, 1>, grid, threads, h->smem_render, st, p));
            else CUDA_TRY(h, launch_step(tde_render_kernel<2, 1>, grid, threads, h->smem_render, st, p));
        } else {
            if (h->A <= 32) CUDA_TRY(h, launch_step(tde_render_kernel<1, 0>, grid, threads, h->smem_render, st, p));
            else CUDA_TRY(h, launch_step(tde_render_kernel<2, 0>, grid, threads, h->smem_render, st, p));
        }
        h->launches++;
    }
    if (deferred) {
        const int u4_per_env = n_stack * (TDE_OBS_C * TDE_OBS_H * TDE_OBS_W / 16);
        const int cgrid = std::max(1, std::min((h->E + 7) / 8, h->sm_count * 8));
        TDE_LAUNCH(cgrid, 256, 0, st, tde_copy_rows_kernel)(h->done_mask, (const uint4*)obs, (uint4*)terminal_obs, h->E, u4_per_env);
        CUDA_TRY(h, cudaGetLastError());
        StepParams q = p;
        q.done_mask = nullptr;
        q.reset_mask = h->done_mask;
        const int rgrid = std::max(1, std::min(want, h->sm_count * 8));
        if (h->A <= 32) TDE_LAUNCH(rgrid, threads, 0, st, tde_reset_kernel<1>)(q);
        else TDE_LAUNCH(rgrid, threads, 0, st, tde_reset_kernel<2>)(q);
        CUDA_TRY(h, cudaGetLastError());
        q.render_mask = h->done_mask;
        q.obs_prev = obs;   // the stack was shifted by the first pass; a re-initialised env only keeps zeros anyway
        const int grid = std::min(h->grid_render, want_render);
        if (n_stack > 1) {
            if (h->A <= 32) TDE_LAUNCH(grid, threads, h->smem_render, st, tde_render_kernel<1, 1>)(q);
            else TDE_LAUNCH(grid, threads, h->smem_render, st, tde_render_kernel<2, 1>)(q);
        } else {
            if (h->A <= 32) TDE_LAUNCH(grid, threads, h->smem_render, st, tde_render_kernel<1, 0>)(q);
            else TDE_LAUNCH(grid, threads, h->smem_render, st, tde_render_kernel<2, 0>)(q);
        }
        CUDA_TRY(h, cudaGetLastError());
        h->launches += 3;
    }
    return TDE_OK;
}

extern "C" int tde_step_terminal(tde_handle* h, const float* actions, uint8_t* obs, int32_t n_stack, uint8_t* terminal_obs,
                                 float* reward, uint8_t* terminated, uint8_t* truncated, float* info, void* stream) {
    if (h && (!obs || !terminal_obs)) return fail(h, TDE_E_INVAL, "tde_step_terminal: obs / terminal_obs is null");
    if (h && !h->cfg.auto_reset) return fail(h, TDE_E_STATE, "tde_step_terminal: needs auto_reset (without it nothing is re-initialised)");
    return step_impl(h, TDE_PH_ALL, actions, obs, n_stack, reward, terminated, truncated, info, stream, nullptr, terminal_obs);
}

extern "C" int tde_step_phases(tde_handle* h, int32_t phases, const float* actions, uint8_t* obs, float* reward,
                               uint8_t* terminated, uint8_t* truncated, float* info, void* stream) {
    return step_impl(h, phases, actions, obs, 1, reward, terminated, truncated, info, stream);
}
extern "C" int tde_step_stacked(tde_handle* h, const float* actions, uint8_t* stack, int32_t n_stack, float* reward,
                                uint8_t* terminated, uint8_t* truncated, float* info, void* stream) {
    if (h && !stack) return fail(h, TDE_E_INVAL, "tde_step_stacked: stack is null");
    return step_impl(h, TDE_PH_ALL, actions, stack, n_stack, reward, terminated, truncated, info, stream);
}
extern "C" int tde_render_stacked(tde_handle* h, uint8_t* stack, int32_t n_stack, void* stream) {
    return step_impl(h, TDE_PH_RENDER, nullptr, stack, n_stack, nullptr, nullptr, nullptr, nullptr, stream);
}
extern "C" int tde_step_rollout(tde_handle* h, const float* actions, const uint8_t* stack_prev, uint8_t* stack_next, int32_t n_stack,
                                float* reward, uint8_t* terminated, uint8_t* truncated, float* info, void* stream) {
    if (h && (!stack_prev || !stack_next)) return fail(h, TDE_E_INVAL, "tde_step_rollout: stack_prev / stack_next is null");
    if (h && n_stack < 2) return fail(h, TDE_E_INVAL, "tde_step_rollout: n_stack must be in 2..8");
    if (h && stack_prev != stack_next) {
        // the two slots must not overlap unless they are the same buffer (then the shift is done in place)
        const size_t bytes = (size_t)h->E * (size_t)n_stack * TDE_OBS_C * TDE_OBS_H * TDE_OBS_W;
        const uint8_t *a = stack_prev, *b = stack_next;
        if (a < b + bytes && b < a + bytes) return fail(h, TDE_E_INVAL, "tde_step_rollout: stack_prev and stack_next overlap");
    }
    return step_impl(h, TDE_PH_ALL, actions, stack_next, n_stack, reward, terminated, truncated, info, stream, stack_prev);
}

extern "C" int tde_step_rollout_scatter(tde_handle* h, const float* actions, uint8_t* stack_next, int64_t slot_stride_bytes, int32_t slots_ahead,
                                        int32_t n_stack, float* reward, uint8_t* terminated, uint8_t* truncated, float* info, void* stream) {
    if (h && !stack_next) return fail(h, TDE_E_INVAL, "tde_step_rollout_scatter: stack_next is null");
    if (h && n_stack < 2) return fail(h, TDE_E_INVAL, "tde_step_rollout_scatter: n_stack must be in 2..8");
    if (h && (slots_ahead < 1 || slots_ahead > n_stack)) return fail(h, TDE_E_INVAL, "tde_step_rollout_scatter: slots_ahead must be in 1..n_stack");
    if (h) {
        const long long slot = (long long)h->E * n_stack * TDE_OBS_C * TDE_OBS_H * TDE_OBS_W;
        if (slot_stride_bytes < slot || (slot_stride_bytes & 15)) return fail(h, TDE_E_INVAL, "tde_step_rollout_scatter: slot_stride_bytes must be >= one slot and a multiple of 16");
    }
    return step_impl(h, TDE_PH_ALL, actions, stack_next, n_stack, reward, terminated, truncated, info, stream, nullptr, nullptr, 0, -1,
                     (long long)slot_stride_bytes, slots_ahead);
}

extern "C" int tde_step_stacked_ring(tde_handle* h, const float* actions, uint8_t* ring, int32_t ring_slots, int32_t ring_pos, int32_t n_stack,
                                     float* reward, uint8_t* terminated, uint8_t* truncated, float* info, void* stream) {
    if (h && !ring) return fail(h, TDE_E_INVAL, "tde_step_stacked_ring: ring is null");
    if (h && (n_stack < 2 || n_stack > 8)) return fail(h, TDE_E_INVAL, "tde_step_stacked_ring: n_stack must be in 2..8");
    if (h && (ring_slots < n_stack || ring_slots > 64)) return fail(h, TDE_E_INVAL, "tde_step_stacked_ring: ring_slots must be in n_stack..64");
    if (h && (ring_pos < 0 || ring_pos >= ring_slots)) return fail(h, TDE_E_INVAL, "tde_step_stacked_ring: ring_pos out of range");
    if (!h) return TDE_E_INVAL;
    const long long slot = (long long)h->E * n_stack * TDE_OBS_C * TDE_OBS_H * TDE_OBS_W;
    return step_impl(h, TDE_PH_ALL, actions, ring + (size_t)ring_pos * (size_t)slot, n_stack, reward, terminated, truncated, info, stream, nullptr, nullptr,
                     0, -1, slot, n_stack, false, ring_slots, ring_pos);
}

extern "C" int tde_step(tde_handle* h, const float* actions, uint8_t* obs, float* reward, uint8_t* terminated,
                        uint8_t* truncated, float* info, void* stream) {
    int phases = obs ? TDE_PH_ALL : (TDE_PH_ALL & ~TDE_PH_RENDER);
    return tde_step_phases(h, phases, actions, obs, reward, terminated, truncated, info, stream);
}

extern "C" int tde_kinematics(tde_handle* h, const float* actions, void* stream) {
    return tde_step_phases(h, TDE_PH_KINEMATICS, actions, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}
extern "C" int tde_render(tde_handle* h, uint8_t* obs, void* stream) {
    return tde_step_phases(h, TDE_PH_RENDER, nullptr, obs, nullptr, nullptr, nullptr, nullptr, stream);
}
extern "C" int tde_render_classes(tde_handle* h, uint8_t* nibbles, void* stream) {
    return step_impl(h, TDE_PH_RENDER, nullptr, nibbles, 1, nullptr, nullptr, nullptr, nullptr, stream, nullptr, nullptr, 0, -1, 0, 0, true);
}
extern "C" int tde_compute_infractions(tde_handle* h, void* stream) {
    return tde_step_phases(h, TDE_PH_INFRACTIONS, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
}

// ---- tde_step_host, compact mode: the frames cross PCIe as the 4-bit class image (2 KB per env instead of 12 KB) and
// host threads expand them to the caller's RGB planes with the handle's palette while later chunks are on their way.
static void expand_env_scalar(const uint8_t* nib, uint8_t* rgb, const uint8_t* lut) {
    constexpr int PX = TDE_OBS_H * TDE_OBS_W;
    for (int b = 0; b < PX / 2; ++b) {
        const int c0 = nib[b] & 15, c1 = nib[b] >> 4;
        for (int ch = 0; ch < 3; ++ch) {
            rgb[ch * PX + 2 * b] = lut[16 * ch + c0];
            rgb[ch * PX + 2 * b + 1] = lut[16 * ch + c1];
        }
    }
}
__attribute__((target("avx2"))) static void expand_env_avx2(const uint8_t* nib, uint8_t* rgb, const uint8_t* lut) {
    constexpr int PX = TDE_OBS_H * TDE_OBS_W;
    const __m256i low4 = _mm256_set1_epi8(0x0f);
    __m256i L[3];
    for (int ch = 0; ch < 3; ++ch) L[ch] = _mm256_broadcastsi128_si256(_mm_loadu_si128((const __m128i*)(lut + 16 * ch)));
    const bool aligned = ((uintptr_t)rgb & 31) == 0;   // non-temporal stores: the planes are not read again by these threads
    for (int r = 0; r < TDE_OBS_H; ++r) {
        const __m256i v = _mm256_loadu_si256((const __m256i*)(nib + 32 * r));
        const __m256i lo = _mm256_and_si256(v, low4), hi = _mm256_and_si256(_mm256_srli_epi16(v, 4), low4);
        const __m256i a = _mm256_unpacklo_epi8(lo, hi), b = _mm256_unpackhi_epi8(lo, hi);
        const __m256i px0 = _mm256_permute2x128_si256(a, b, 0x20), px1 = _mm256_permute2x128_si256(a, b, 0x31);   // pixels 0..31, 32..63
        for (int ch = 0; ch < 3; ++ch) {
            __m256i* o = (__m256i*)(rgb + ch * PX + r * TDE_OBS_W);
            const __m256i w0 = _mm256_shuffle_epi8(L[ch], px0), w1 = _mm256_shuffle_epi8(L[ch], px1);
            if (aligned) { _mm256_stream_si256(o, w0); _mm256_stream_si256(o + 1, w1); }
            else { _mm256_storeu_si256(o, w0); _mm256_storeu_si256(o + 1, w1); }
        }
    }
}
__attribute__((target("avx512f,avx512bw"))) static void expand_env_avx512(const uint8_t* nib, uint8_t* rgb, const uint8_t* lut) {
    constexpr int PX = TDE_OBS_H * TDE_OBS_W;
    const __m512i low4 = _mm512_set1_epi16(0x0f0f);
    __m512i L[3];
    for (int ch = 0; ch < 3; ++ch) L[ch] = _mm512_broadcast_i32x4(_mm_loadu_si128((const __m128i*)(lut + 16 * ch)));
    const bool aligned = ((uintptr_t)rgb & 63) == 0;   // a row of a plane is one cache line: one non-temporal store each
    for (int r = 0; r < TDE_OBS_H; ++r) {
        // byte b of the row = pixels 2 b (low nibble) and 2 b + 1 (high nibble): widen to 16 bits and pull the high nibble up
        const __m512i x = _mm512_cvtepu8_epi16(_mm256_loadu_si256((const __m256i*)(nib + 32 * r)));
        const __m512i px = _mm512_and_si512(_mm512_or_si512(x, _mm512_slli_epi16(x, 4)), low4);
        for (int ch = 0; ch < 3; ++ch) {
            __m512i* o = (__m512i*)(rgb + ch * PX + r * TDE_OBS_W);
            const __m512i w = _mm512_shuffle_epi8(L[ch], px);
            if (aligned) _mm512_stream_si512(o, w); else _mm512_storeu_si512(o, w);
        }
    }
}
// widest instruction set the expansion may use: what the CPU has, capped by TDE_HOST_SIMD=scalar|avx2|avx512 (TDE_HOST_NO_SIMD=1 = scalar)
static int host_simd_level() {
    int level = __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512f") ? 2 : __builtin_cpu_supports("avx2") ? 1 : 0;
    if (std::getenv("TDE_HOST_NO_SIMD")) level = 0;
    if (const char* v = std::getenv("TDE_HOST_SIMD")) level = std::min(level, !std::strcmp(v, "scalar") ? 0 : !std::strcmp(v, "avx2") ? 1 : 2);
    return level;
}

// The host threads of the expansion.  One job per step: the threads take blocks of envs in order and wait (yielding,
// never spinning hard) until the chunk a block belongs to has landed in pinned memory; between steps they sleep on a
// condition variable.  The calling thread works too, so `threads` = 1 means no pool at all.
struct ExpandPool {
    static constexpr int BLOCK = 8;   // envs per grab: 16 KB read, 96 KB written
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    unsigned long long job = 0;
    int pending = 0;
    bool stop = false;
    // the job
    const uint8_t* nib = nullptr; uint8_t* rgb = nullptr; int E = 0; int simd = 0;
    uint8_t lut[48] = {};
    std::atomic<int> next{0}, ready{0};

    explicit ExpandPool(int threads) {
        for (int t = 1; t < threads; ++t) workers.emplace_back([this] { loop(); });
    }
    ~ExpandPool() {
        { std::lock_guard<std::mutex> lk(m); stop = true; }
        cv_job.notify_all();
        for (std::thread& t : workers) t.join();
    }
    void work() {
        constexpr int PX = TDE_OBS_H * TDE_OBS_W;
        for (;;) {
            const int b0 = next.fetch_add(BLOCK, std::memory_order_relaxed);
            if (b0 >= E) break;
            const int b1 = std::min(E, b0 + BLOCK);
            while (ready.load(std::memory_order_acquire) < b1) std::this_thread::yield();
            for (int e = b0; e < b1; ++e) {
                if (simd == 2) expand_env_avx512(nib + (size_t)e * (PX / 2), rgb + (size_t)e * 3 * PX, lut);
                else if (simd == 1) expand_env_avx2(nib + (size_t)e * (PX / 2), rgb + (size_t)e * 3 * PX, lut);
                else expand_env_scalar(nib + (size_t)e * (PX / 2), rgb + (size_t)e * 3 * PX, lut);
            }
        }
        if (simd) _mm_sfence();
    }
    void loop() {
        unsigned long long seen = 0;
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv_job.wait(lk, [&] { return stop || job != seen; });
            if (stop) return;
            seen = job;
            lk.unlock();
            work();
            lk.lock();
            if (--pending == 0) cv_done.notify_one();
        }
    }
    // post the job; the caller then publishes `ready` as the chunks arrive and finally calls finish()
    void begin(const uint8_t* nib_, uint8_t* rgb_, const uint8_t* palette, int E_) {
        {
            std::lock_guard<std::mutex> lk(m);
            nib = nib_; rgb = rgb_; E = E_;
            simd = host_simd_level();   // read per step: the narrower loops stay testable
            std::memset(lut, 0, sizeof(lut));
            for (int ch = 0; ch < 3; ++ch)
                for (int c = 0; c < TDE_NUM_CLASSES; ++c) lut[16 * ch + c] = palette[3 * c + ch];
            next.store(0); ready.store(0);
            pending = (int)workers.size();
            ++job;
        }
        cv_job.notify_all();
    }
    void arrived(int envs) { ready.store(envs, std::memory_order_release); }
    // the caller expands everything itself (all of it has arrived): no worker is woken
    void begin_serial(const uint8_t* nib_, uint8_t* rgb_, const uint8_t* palette, int E_) {
        std::lock_guard<std::mutex> lk(m);
        nib = nib_; rgb = rgb_; E = E_;
        simd = host_simd_level();
        std::memset(lut, 0, sizeof(lut));
        for (int ch = 0; ch < 3; ++ch)
            for (int c = 0; c < TDE_NUM_CLASSES; ++c) lut[16 * ch + c] = palette[3 * c + ch];
        next.store(0); ready.store(E_);
        work();
    }
    void finish() {
        work();
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return pending == 0; });
    }
};

static void delete_pool(ExpandPool* pool) { delete pool; }

// threads of the expansion: TDE_HOST_THREADS, else the CPUs this process may run on shared between the ranks of the
// box (torchrun's LOCAL_WORLD_SIZE), at most 16 (one GPU's frames saturate the host memory well before that)
static int host_thread_count() {
    if (const char* v = std::getenv("TDE_HOST_THREADS")) return std::max(1, std::min(64, std::atoi(v)));
    int cpus = (int)std::max(1u, std::thread::hardware_concurrency());
#ifdef __linux__
    cpu_set_t set;
    if (sched_getaffinity(0, sizeof(set), &set) == 0) cpus = std::max(1, CPU_COUNT(&set));
#endif
    int ranks = 1;
    if (const char* w = std::getenv("LOCAL_WORLD_SIZE")) ranks = std::max(1, std::atoi(w));
    return std::max(1, std::min(cpus / ranks, 16));
}

static int ensure_copy_stream(tde_handle* h) {
    if (h->copy_stream && h->copies_done) return TDE_OK;   // copies_done is created last: everything else exists then
    if (!h->copy_stream) CUDA_TRY(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (cudaEvent_t& ev : h->chunk_done) if (!ev) CUDA_TRY(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (cudaEvent_t& ev : h->chunk_copied) if (!ev) CUDA_TRY(h, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_TRY(h, cudaEventCreateWithFlags(&h->copies_done, cudaEventDisableTiming));
    return TDE_OK;
}

// chunks of envs whose frames travel while later chunks are computed (TDE_HOST_CHUNKS overrides the choice, 1..16)
static inline int host_chunks(const tde_handle* h, int few, int some) {
    int n = h->E < few ? 1 : h->E < some ? 8 : 16;
    if (const char* v = std::getenv("TDE_HOST_CHUNKS")) n = std::max(1, std::min(16, std::atoi(v)));
    return std::min(n, h->E);
}
static inline int compact_chunks(const tde_handle* h) { return host_chunks(h, 2048, 8192); }

static int step_host_compact_launch(tde_handle* h, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    const size_t E = (size_t)h->E, nib_env = TDE_OBS_H * TDE_OBS_W / 2;
    // each piece on its own: a failed allocation leaves the others to be retried by the next call
    if (!h->h_nib) CUDA_TRY(h, cudaMalloc((void**)&h->h_nib, E * nib_env));
    if (!h->pin_nib) CUDA_TRY(h, cudaMallocHost((void**)&h->pin_nib, E * nib_env));
    if (!h->pool) h->pool = new ExpandPool(host_thread_count());
    if (int rc = ensure_copy_stream(h)) return rc;
    const int chunks = compact_chunks(h);
    for (int c = 0; c < chunks; ++c) {
        const int e0 = (int)((long long)h->E * c / chunks), e1 = (int)((long long)h->E * (c + 1) / chunks);
        int rc = step_impl(h, TDE_PH_ALL, h->h_actions, h->h_nib, 1, h->h_reward, h->h_term, h->h_trunc, h->h_info, stream, nullptr, nullptr,
                           e0, e1, 0, 0, true);
        if (rc) return rc;
        CUDA_TRY(h, cudaEventRecord(h->chunk_done[c], st));
        CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->chunk_done[c], 0));
        CUDA_TRY(h, cudaMemcpyAsync(h->pin_nib + e0 * nib_env, h->h_nib + e0 * nib_env, (size_t)(e1 - e0) * nib_env, cudaMemcpyDeviceToHost, h->copy_stream));
        CUDA_TRY(h, cudaEventRecord(h->chunk_copied[c], h->copy_stream));
    }
    return TDE_OK;
}

static int step_host_compact_expand(tde_handle* h, uint8_t* obs) {
    const int chunks = compact_chunks(h);
    if (h->E < 256) {   // a handful of envs (one, behind the gym.Env API): waking the pool costs more than expanding them here
        for (int c = 0; c < chunks; ++c) CUDA_TRY(h, cudaEventSynchronize(h->chunk_copied[c]));
        h->pool->begin_serial(h->pin_nib, obs, h->palette, h->E);
        return TDE_OK;
    }
    h->pool->begin(h->pin_nib, obs, h->palette, h->E);
    cudaError_t err = cudaSuccess;
    for (int c = 0; c < chunks && err == cudaSuccess; ++c) {
        err = cudaEventSynchronize(h->chunk_copied[c]);
        if (err == cudaSuccess) h->pool->arrived((int)((long long)h->E * (c + 1) / chunks));
    }
    if (err != cudaSuccess) h->pool->arrived(h->E);   // release the threads; the caller gets the error, not the frames
    h->pool->finish();
    CUDA_TRY(h, err);
    return TDE_OK;
}

extern "C" int tde_step_host(tde_handle* h, const float* actions, uint8_t* obs, float* reward, uint8_t* terminated,
                             uint8_t* truncated, float* info, void* stream) {
    if (!h) return TDE_E_INVAL;
    if (!actions || !reward || !terminated || !truncated || !info) return fail(h, TDE_E_INVAL, "tde_step_host: null argument");
    TDE_ON_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t E = (size_t)h->E, obs_bytes = E * TDE_OBS_C * TDE_OBS_H * TDE_OBS_W;
    if (!h->h_actions) CUDA_TRY(h, cudaMalloc((void**)&h->h_actions, E * 2 * sizeof(float)));
    if (!h->h_reward) CUDA_TRY(h, cudaMalloc((void**)&h->h_reward, E * sizeof(float)));
    if (!h->h_term) CUDA_TRY(h, cudaMalloc((void**)&h->h_term, E));
    if (!h->h_trunc) CUDA_TRY(h, cudaMalloc((void**)&h->h_trunc, E));
    if (!h->h_info) CUDA_TRY(h, cudaMalloc((void**)&h->h_info, E * TDE_INFO_STRIDE * sizeof(float)));
    // the observation crosses PCIe as the class image unless cfg.host_obs_rgb / TDE_HOST_OBS=rgb asks for the planes
    const char* mode = std::getenv("TDE_HOST_OBS");
    const bool compact = obs && (mode ? std::strcmp(mode, "rgb") != 0 : h->cfg.host_obs_rgb == 0);
    if (obs && !compact && !h->h_obs) CUDA_TRY(h, cudaMalloc((void**)&h->h_obs, obs_bytes));
    CUDA_TRY(h, cudaMemcpyAsync(h->h_actions, actions, E * 2 * sizeof(float), cudaMemcpyHostToDevice, st));
    if (compact) {
        // the small rows travel while the host threads expand the frames
        int rc = step_host_compact_launch(h, stream);
        if (rc) return rc;
        CUDA_TRY(h, cudaMemcpyAsync(reward, h->h_reward, E * sizeof(float), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(h, cudaMemcpyAsync(terminated, h->h_term, E, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(h, cudaMemcpyAsync(truncated, h->h_trunc, E, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(h, cudaMemcpyAsync(info, h->h_info, E * TDE_INFO_STRIDE * sizeof(float), cudaMemcpyDeviceToHost, st));
        rc = step_host_compact_expand(h, obs);
        if (rc) return rc;
        CUDA_TRY(h, cudaStreamSynchronize(st));
        return TDE_OK;
    }
    // With observations the step is bound by the 12 KB per env that cross PCIe: the envs are stepped in
    // chunks, and the frames of a finished chunk travel on a second stream while the next chunk is computed.
    const int chunks = !obs ? 1 : host_chunks(h, 4096, 8192);   // only the first chunk's compute is not hidden behind the copies
    if (chunks == 1) {
        int rc = tde_step(h, h->h_actions, obs ? h->h_obs : nullptr, h->h_reward, h->h_term, h->h_trunc, h->h_info, stream);
        if (rc) return rc;
        if (obs) CUDA_TRY(h, cudaMemcpyAsync(obs, h->h_obs, obs_bytes, cudaMemcpyDeviceToHost, st));
    } else {
        if (int rc = ensure_copy_stream(h)) return rc;
        const size_t frame = (size_t)TDE_OBS_C * TDE_OBS_H * TDE_OBS_W;
        for (int c = 0; c < chunks; ++c) {
            const int e0 = (int)((long long)h->E * c / chunks), e1 = (int)((long long)h->E * (c + 1) / chunks);
            int rc = step_impl(h, TDE_PH_ALL, h->h_actions, h->h_obs, 1, h->h_reward, h->h_term, h->h_trunc, h->h_info, stream, nullptr, nullptr, e0, e1);
            if (rc) return rc;
            CUDA_TRY(h, cudaEventRecord(h->chunk_done[c], st));
            CUDA_TRY(h, cudaStreamWaitEvent(h->copy_stream, h->chunk_done[c], 0));
            CUDA_TRY(h, cudaMemcpyAsync(obs + e0 * frame, h->h_obs + e0 * frame, (size_t)(e1 - e0) * frame, cudaMemcpyDeviceToHost, h->copy_stream));
        }
        CUDA_TRY(h, cudaEventRecord(h->copies_done, h->copy_stream));
        CUDA_TRY(h, cudaStreamWaitEvent(st, h->copies_done, 0));
    }
    CUDA_TRY(h, cudaMemcpyAsync(reward, h->h_reward, E * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaMemcpyAsync(terminated, h->h_term, E, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaMemcpyAsync(truncated, h->h_trunc, E, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaMemcpyAsync(info, h->h_info, E * TDE_INFO_STRIDE * sizeof(float), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return TDE_OK;
}

extern "C" int tde_render_view(tde_handle* h, int32_t env, float cam_x, float cam_y, float cam_psi, float fov, int32_t width,
                               int32_t height, uint8_t* out, void* stream) {
    if (!h) return TDE_E_INVAL;
    if (!out) return fail(h, TDE_E_INVAL, "tde_render_view: out is null");
    if (!h->uploaded || !h->was_reset) return fail(h, TDE_E_STATE, "tde_render_view: upload scenarios and reset first");
    if (env < 0 || env >= h->E) return fail(h, TDE_E_INVAL, "tde_render_view: env out of range");
    if (width < 1 || height < 1 || width > TDE_VIEW_MAX_RES || height > TDE_VIEW_MAX_RES) return fail(h, TDE_E_INVAL, "tde_render_view: resolution must be in 1..4096");
    if (!(fov > 0.0f)) return fail(h, TDE_E_INVAL, "tde_render_view: fov must be positive");
    TDE_ON_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    int nmax = 0;   // the env's map is only known on the device: size the scratch for the largest map
    for (const MapDev& M : h->tab->maps_host) nmax = std::max(nmax, M.ntri + M.nmark + M.nstop);
    nmax += 1 + 2 * h->A;
    if (nmax > h->view_cap) {
        CUDA_TRY(h, cudaStreamSynchronize(st));
        cudaFree(h->view_prims); h->view_prims = nullptr; h->view_cap = 0;
        CUDA_TRY(h, cudaMalloc((void**)&h->view_prims, sizeof(ViewPrim) * (size_t)nmax));
        h->view_cap = nmax;
    }
    StepParams p = make_params(h);
    ViewParams v;
    v.env = env; v.W = width; v.H = height; v.ex = cam_x; v.ey = cam_y;
    v.psi = cam_psi; v.ce = 1.0f; v.se = 0.0f;   // sin / cos are taken on the device (tde_sincosf, as the oracle's twin does)
    v.ppm = (float)width / fov;
    v.ppmy = h->cfg.left_handed_coordinates ? v.ppm : -v.ppm;
    std::memcpy(v.pal, h->palette, sizeof(v.pal));
    // slots beyond the env's own map (a smaller map than the largest) are emitted as rejected primitives
    TDE_LAUNCH((nmax + 127) / 128, 128, 0, st, tde_view_prims_kernel)(p, v, h->view_prims, nmax);
    CUDA_TRY(h, cudaGetLastError());
    dim3 grid((width + TDE_VIEW_TX - 1) / TDE_VIEW_TX, (height + TDE_VIEW_TY - 1) / TDE_VIEW_TY), block(TDE_VIEW_TX, TDE_VIEW_TY);
    TDE_LAUNCH(grid, block, 0, st, tde_view_raster_kernel)(h->view_prims, nmax, v, out);
    CUDA_TRY(h, cudaGetLastError());
    h->launches += 2;
    return TDE_OK;
}

static int copy_dev(tde_handle* h, void* dst, const void* src, size_t bytes, void* stream, const char* what) {
    if (!h) return TDE_E_INVAL;
    if (!dst || !src) return fail(h, TDE_E_INVAL, std::string(what) + ": null argument");
    TDE_ON_DEVICE(h);
    CUDA_TRY(h, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
    return TDE_OK;
}
extern "C" int tde_get_state(tde_handle* h, float* out, void* stream) {
    return copy_dev(h, out, h ? h->state : nullptr, h ? (size_t)h->E * h->A * 16 : 0, stream, "tde_get_state");
}
extern "C" int tde_set_state(tde_handle* h, const float* in, void* stream) {
    return copy_dev(h, h ? h->state : nullptr, in, h ? (size_t)h->E * h->A * 16 : 0, stream, "tde_set_state");
}
extern "C" int tde_get_attributes(tde_handle* h, float* out, void* stream) {
    return copy_dev(h, out, h ? h->attr : nullptr, h ? (size_t)h->E * h->A * 16 : 0, stream, "tde_get_attributes");
}
extern "C" int tde_set_attributes(tde_handle* h, const float* in, void* stream) {
    return copy_dev(h, h ? h->attr : nullptr, in, h ? (size_t)h->E * h->A * 16 : 0, stream, "tde_set_attributes");
}
extern "C" int tde_get_infractions(tde_handle* h, float* out, void* stream) {
    return copy_dev(h, out, h ? h->infr : nullptr, h ? (size_t)h->E * h->A * 16 : 0, stream, "tde_get_infractions");
}
extern "C" int tde_get_env_vars(tde_handle* h, int32_t* out, void* stream) {
    return copy_dev(h, out, h ? h->vars : nullptr, h ? (size_t)h->E * 32 : 0, stream, "tde_get_env_vars");
}
extern "C" int tde_set_env_vars(tde_handle* h, const int32_t* in, void* stream) {
    return copy_dev(h, h ? h->vars : nullptr, in, h ? (size_t)h->E * 32 : 0, stream, "tde_set_env_vars");
}

extern "C" int tde_collision_boxes(const float* state, const float* attr, int32_t E, int32_t A, float* out, void* stream) {
    if (!state || !attr || !out || E < 1 || A < 1 || A > TDE_MAX_AGENTS) return fail(nullptr, TDE_E_INVAL, "tde_collision_boxes: bad argument");
    // runs on the GPU that owns the boxes (not on whatever device happens to be current)
    int dev = 0, sms = 0;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, state) != cudaSuccess || pa.type != cudaMemoryTypeDevice) {
        cudaGetLastError();
        return fail(nullptr, TDE_E_INVAL, "tde_collision_boxes: state_dev is not a device pointer");
    }
    dev = pa.device;
    DeviceGuard guard(dev);
    if (!guard.ok || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
        return fail(nullptr, TDE_E_CUDA, "tde_collision_boxes: no CUDA device");
    // resident CTAs per SM from the occupancy calculator (the 64-agent kernel holds 13, shared memory bound)
    int per_sm = 8;
    if (A <= 32) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tde_collision_kernel<1>, TDE_WARPS_PER_BLOCK * 32, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tde_collision_kernel<2>, TDE_WARPS_PER_BLOCK * 32, 0);
    int grid = std::max(1, std::min((E + TDE_WARPS_PER_BLOCK - 1) / TDE_WARPS_PER_BLOCK, sms * std::max(per_sm, 1)));
    cudaStream_t st = (cudaStream_t)stream;
    if (A <= 32) TDE_LAUNCH(grid, TDE_WARPS_PER_BLOCK * 32, 0, st, tde_collision_kernel<1>)((const float4*)state, (const float4*)attr, E, A, out);
    else TDE_LAUNCH(grid, TDE_WARPS_PER_BLOCK * 32, 0, st, tde_collision_kernel<2>)((const float4*)state, (const float4*)attr, E, A, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(nullptr, TDE_E_CUDA, std::string("tde_collision_boxes: ") + cudaGetErrorString(e));
    return TDE_OK;
}

extern "C" int tde_offroad_boxes(tde_handle* h, int32_t map_id, const float* state, const float* attr, int32_t E, int32_t A,
                                 float* out, void* stream) {
    if (!h) return TDE_E_INVAL;
    if (!h->uploaded) return fail(h, TDE_E_STATE, "tde_offroad_boxes: upload scenarios first");
    if (!state || !attr || !out || E < 1 || A < 1 || map_id < 0 || map_id >= h->tab->num_maps) return fail(h, TDE_E_INVAL, "tde_offroad_boxes: bad argument");
    TDE_ON_DEVICE(h);
    int n = E * A;
    // 6 CTAs of 8 warps per SM: the kernel lives on L1 hits of the cell records and triangle tables, and every resident
    // CTA takes 22 KB of shared memory away from L1 (measured: 232 us against 266 us with 8 CTAs per SM, 243 with 4)
    int per_sm = 6;
    if (const char* v = std::getenv("TDE_OFFROAD_CTAS")) per_sm = std::max(1, std::min(8, std::atoi(v)));
    int grid = std::max(1, std::min((n + 255) / 256, h->sm_count * per_sm));
    TDE_LAUNCH(grid, TDE_OFFROAD_WARPS * 32, 0, (cudaStream_t)stream, tde_offroad_kernel)(h->tab->maps_dev, map_id, h->cfg.offroad_threshold, (const float4*)state,
                                                               (const float4*)attr, n, out);
    CUDA_TRY(h, cudaGetLastError());
    h->launches++;
    return TDE_OK;
}

#ifdef TDE_TRACE
extern "C" int tde_debug_set_trace(unsigned long long* dev_ptr) {
    return cudaMemcpyToSymbol(g_trace, &dev_ptr, sizeof(dev_ptr)) == cudaSuccess ? TDE_OK : TDE_E_CUDA;
}
#endif

// simulator.copy() (gym_env.py:110): an independent handle with the same envs.  The per-env arrays are copied device
// to device on `stream`; the scenario tables are shared (reference-counted, read-only after the upload).
extern "C" int tde_clone(tde_handle* h, tde_handle** out, void* stream) {
    if (!h || !out) return fail(h, TDE_E_INVAL, "tde_clone: null argument");
    if (!h->uploaded) return fail(h, TDE_E_STATE, "tde_clone: upload scenarios first");
    tde_handle* c = nullptr;
    int rc = tde_create(&h->cfg, &c);
    if (rc) return fail(h, rc, std::string("tde_clone: ") + g_create_error);
    TDE_ON_DEVICE(h);
    c->tab = h->tab;
    c->uploaded = true; c->was_reset = h->was_reset; c->seed = h->seed;
    std::memcpy(c->palette, h->palette, sizeof(h->palette));
    c->phys_wpb = h->phys_wpb; c->phys_per_sm = h->phys_per_sm; c->smem_phys = h->smem_phys; c->grid_phys = h->grid_phys;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t EA = (size_t)h->E * h->A, E = (size_t)h->E;
    cudaError_t e = cudaSuccess;
    auto cp = [&](void* d, const void* s, size_t n) { if (e == cudaSuccess) e = cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st); };
    cp(c->state, h->state, EA * 16); cp(c->attr, h->attr, EA * 16); cp(c->infr, h->infr, EA * 16);
    cp(c->vars, h->vars, E * 32); cp(c->ep_return, h->ep_return, E * 4); cp(c->restart, h->restart, E);
    cp(c->scen_lo, h->scen_lo, E * 4); cp(c->scen_hi, h->scen_hi, E * 4);
    if (e != cudaSuccess) { tde_destroy(c); return fail(h, TDE_E_CUDA, std::string("tde_clone: ") + cudaGetErrorString(e)); }
    *out = c;   // episode statistics start at zero in the copy
    return TDE_OK;
}

extern "C" int tde_get_episode_stats(tde_handle* h, double* out, int32_t reset_after, void* stream) {
    if (!h || !out) return fail(h, TDE_E_INVAL, "tde_get_episode_stats: null argument");
    TDE_ON_DEVICE(h);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(h, cudaMemcpyAsync(out, h->stats, sizeof(double) * TDE_NUM_STATS, cudaMemcpyDeviceToHost, st));
    if (reset_after) CUDA_TRY(h, cudaMemsetAsync(h->stats, 0, sizeof(double) * TDE_NUM_STATS, st));
    CUDA_TRY(h, cudaStreamSynchronize(st));
    return TDE_OK;
}

extern "C" int tde_num_kernel_launches(const tde_handle* h, int64_t* out) {
    if (!h || !out) return TDE_E_INVAL;
    *out = h->launches;
    return TDE_OK;
}
extern "C" int tde_get_map_info(const tde_handle* h, int32_t map_id, int32_t* out8) {
    if (!h || !out8 || !h->tab || map_id < 0 || map_id >= (int)h->tab->map_info.size()) return TDE_E_INVAL;
    for (int k = 0; k < 8; ++k) out8[k] = h->tab->map_info[map_id][k];
    return TDE_OK;
}
extern "C" int tde_device_sm_count(const tde_handle* h, int32_t* out) {
    if (!h || !out) return TDE_E_INVAL;
    *out = h->sm_count;
    return TDE_OK;
}
