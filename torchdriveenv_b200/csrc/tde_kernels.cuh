// tde_kernels.cuh — the fused per-timestep kernel and its helpers (sm_100a).
//
// Two kernels per env step, one warp per environment in both:
//   tde_physics_kernel  lanes = agents: bicycle step / NPC replay, all-pairs SAT, corner-to-lane-mesh
//                       offroad, red-light stop lines, wrong-way, reward + termination + bookkeeping,
//                       episode statistics, in-kernel auto-reset;
//   tde_render_kernel   lanes = primitives, then image rows: the egocentric 3x64x64 birdview.
// Nothing is exchanged between warps, so there is no block-level synchronisation on the step path.
#pragma once
#include <climits>

#include "tde_device.cuh"
#include "../../include/tde_b200.h"

#ifdef TDE_HOST_EMU
#define TDE_DYN_SMEM(name) unsigned char* const name = emu::dyn_smem()
#else
#define TDE_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

// Small blocks (4 warps), so that the slots of a block are handed to the next launch as soon as its few warps have run
// dry.  Both step kernels run 6 blocks = 24 warps per SM at up to 80 registers per thread: at the 64 registers that 8 blocks
// per SM allow both spill (1.3 M local-memory requests per render launch, 1.4 M per physics launch at C3), and the spills
// cost more than the eight extra warps hide (render 117.9 -> 108.0 us, physics 53.8 -> 51.4 us; 5 blocks / 96 registers:
// 113.4 / 52.6 us).
#ifndef TDE_WARPS_PER_BLOCK
#define TDE_WARPS_PER_BLOCK 4
#endif
#ifndef TDE_RENDER_BLOCKS_PER_SM
#define TDE_RENDER_BLOCKS_PER_SM 6
#endif
#ifndef TDE_PHYS_BLOCKS_PER_SM
#define TDE_PHYS_BLOCKS_PER_SM 6
#endif
// the staged physics launch is one fat CTA per SM: up to 24 warps at the same register budget
#ifndef TDE_PHYS_STAGED_MAX_WARPS
#define TDE_PHYS_STAGED_MAX_WARPS 24
#endif

struct MapDev {
    const float4* tri;          // 3 float4 per road triangle (see tde_point_tri_dist2)
    const float4* rp;           // static render primitives (road + lane markings): 2 float4 = 4 vertices (triangle: v3 == v0);
                                // the n_big oversized ones first, the rest sorted by tile (row-major) of their bbox min corner
    const uint8_t* rp_cls;      // class of each static render primitive
    const int* tile_start;      // [tny*tnx + 1] index into rp of the first primitive of each tile
    const float4* stop;         // 2 float4 per stop line: [x y hl hw] [c s rr 0]
    const uint8_t* lights;      // [period][nstop]
    const int2* cell_rec;       // [gnx*gny] per cell: x = first item, y = n_items << 16 | TDE_CELL_SAFE | n_overlapping
    const uint16_t* cell_items; // per cell: overlapping triangles first, then the other nearest candidates
    int ntri, nmark, nstop, period;
    int n_rp, n_big;
    float gx0, gy0, inv_cell;
    int gnx, gny;
    float tgx0, tgy0, tinv, maxext;  // tile grid of the render primitives; maxext = largest bbox extent of a tiled primitive
    int tnx, tny;
    const float* mark_raw;      // [nmark][6] lane-marking triangles as uploaded (recording view, tde_view.cuh)
};

static_assert(sizeof(MapDev) % 16 == 0, "MapDev rows are copied into shared memory with 16-byte granularity");

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
    const uint8_t* obs_prev;  // stacked render: the stack the older frames are taken from (== obs: shifted in place)
    float* reward;
    uint8_t* terminated;
    uint8_t* truncated;
    float* info;
    double* stats;
    const uint8_t* reset_mask;
    unsigned int* tickets;   // render kernel of this launch: [0] next ticket, [1] finished warps
    int e_begin, e_end;      // envs [e_begin, e_end) of this launch
    uint8_t* done_mask;  // non-null: the physics kernel flags finished envs here instead of re-initialising them (tde_step_terminal)
    const uint8_t* render_mask;  // non-null: only envs with a non-zero byte are rendered (strided, no tickets)
    uint8_t* restart;    // [E] 1 = the env was reset since its last stacked frame
    int n_stack;         // frames per env in obs (1 = plain observation)
    int slots_ahead;     // scatter mode (rollout buffer): slots from the one being written to the end of the buffer, 1..n_stack
    long long slot_stride;   // scatter mode: bytes between consecutive time slots of the buffer
    long long copy_off[8];   // scatter mode: where copy j of the new frame goes, in bytes from copy 0 (one slot ahead and one
                             // channel group down per copy; the slots are consecutive in a rollout buffer, modulo the ring size in a ring)
    uint32_t pal[3][4];  // per channel: 16 class bytes
    float ppm, ppmy;
    // physics kernel, STAGED launches: the map tables as one blob (16 B header per map: offsets of its triangle records,
    // stop lines, light schedule and cell summary; then the MapDev array; then the tables), copied to shared memory once
    // per CTA with bulk-async copies
    const unsigned char* stage_blob;
    unsigned int stage_bytes, stage_maps_off;
};

struct Cam { float ex, ey, ce, se, ppm, ppmy; };

#ifdef TDE_TRACE   // debug builds only (tools/build_variant.sh trace -DTDE_TRACE): per-env start/end timestamps
__device__ unsigned long long* g_trace = nullptr;   // [E][8]: render start, render end, physics start, physics end (ns), static queued, dynamic items, queued total
__device__ __forceinline__ unsigned long long tde_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#define TDE_TRACE_MARK(e, k) do { if (g_trace && lane == 0) g_trace[(size_t)(e) * 8 + (k)] = tde_now(); } while (0)
#else
#define TDE_TRACE_MARK(e, k) do { } while (0)
#endif

// Programmatic dependent launch: a step kernel lets the next kernel on the stream be scheduled once all its warps have
// run dry, and itself waits for the previous kernel's completion and memory flush before it reads anything that
// kernel may have written.  No-ops unless the next kernel is launched with the attribute: the physics kernel of a step
// without observations is (small batches step in 10 us, the gap between two launches is 2 of them); between the physics
// and the render kernel of a full step it loses (different shared-memory carve-outs).
#ifndef TDE_HOST_EMU
__device__ __forceinline__ void tde_pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void tde_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#else
__device__ __forceinline__ void tde_pdl_launch_dependents() {}
__device__ __forceinline__ void tde_pdl_wait() {}
#endif

// Envs are handed to warps one at a time from a device-side ticket counter: their costs differ (what
// is in view, how many agents are off the road), so a fixed stride leaves most warps waiting for the
// unluckiest one.  ctr[0] = next ticket, ctr[1] = warps that have run dry; the last of them re-arms
// both for the next launch, so the kernels stay self-contained (and CUDA-graph replayable).
__device__ __forceinline__ int next_env(unsigned int* ctr, int lane) {
    unsigned int t = 0;
    if (lane == 0) t = atomicAdd(&ctr[0], 1u);
    return (int)__shfl_sync(FULL_MASK, t, 0);
}
__device__ __forceinline__ void envs_done(unsigned int* ctr, int lane, int warps_total) {
    if (lane == 0) {
        __threadfence();
        if (atomicAdd(&ctr[1], 1u) == (unsigned)(warps_total - 1)) { ctr[0] = 0u; ctr[1] = 0u; }
    }
}

#define TDE_PAIR_CAP 256
template <int AH>   // sized for 32 * AH agents: the physics kernel lives on L1 hits, every KB not carved out for shared memory counts
struct SatScratch {  // per warp: staged boxes, candidate pairs and hit counters of the all-pairs SAT
    float4 pos[32 * AH];                   // x y rr -   (rr = circumradius bound + 5 mm, NaN for an absent agent)
    float4 ext[32 * AH];                   // c s hl hw
    int cnt[32 * AH];                      // overlaps found per agent
    unsigned short pairs[TDE_PAIR_CAP];    // a | j << 8, a < j: pairs that survived the broad phase
};

template <int AH>
__device__ __forceinline__ Box sat_ld(const SatScratch<AH>* ws, int a) {
    float4 u = ws->pos[a], v = ws->ext[a];
    Box b; b.x = u.x; b.y = u.y; b.hl = v.z; b.hw = v.w; b.c = v.x; b.s = v.y; b.present = 1.0f; b.r = u.z;
    return b;
}

// compute_collision (gym_env.py:143,415,428): for every agent the number of other present agents whose
// box overlaps its own.  Broad phase: lane = agent a, box j broadcast from shared memory, circle test
// with the circumradius bounds (+1 cm, +0.1 %), only pairs a < j are kept as a bit per lane.  The
// surviving pairs are compacted into a list and the exact 4-axis SAT (bitwise symmetric in its
// arguments) runs once per pair with all lanes busy; a hit counts for both agents.  The broad phase only
// has to be conservative, not reproducible: it uses fused multiply-adds and radii scaled by sqrt(1.001) up front.
template <int AH>
__device__ __forceinline__ void sat_counts(SatScratch<AH>* ws, const Box (&me)[AH], int A, int lane, float (&cnt_out)[AH]) {
    float rr[AH];
#pragma unroll
    for (int h = 0; h < AH; ++h) {
        int a = h * 32 + lane;
        rr[h] = (a < A && me[h].present != 0.0f) ? (me[h].r + 0.005f) * 1.0005f : __int_as_float(0x7fc00000);   // 1.0005^2 > 1.001
        if (a < A) {
            ws->pos[a] = make_float4(me[h].x, me[h].y, rr[h], 0.0f);
            ws->ext[a] = make_float4(me[h].c, me[h].s, me[h].hl, me[h].hw);
            ws->cnt[a] = 0;
        }
    }
    __syncwarp();
    unsigned cm[AH][AH];  // [my slot][j / 32]
#pragma unroll
    for (int h = 0; h < AH; ++h)
#pragma unroll
        for (int g = 0; g < AH; ++g) cm[h][g] = 0u;
    if (AH == 1) {
        // one slot per lane: rotate instead of broadcasting.  At step r lane a meets lane (a + r) mod 32, so
        // r = 1..15 visits every unordered pair once and r = 16 twice (kept for a < 16): half the tests.
        unsigned m = 0u;   // bit r: the agent r lanes further on (mod 32) is a candidate
#pragma unroll 4
        for (int r = 1; r <= 16; ++r) {
            const int src = (lane + r) & 31;
            const float ox = __shfl_sync(FULL_MASK, me[0].x, src), oy = __shfl_sync(FULL_MASK, me[0].y, src);
            const float orr = __shfl_sync(FULL_MASK, rr[0], src);
            const float dx = ox - me[0].x, dy = oy - me[0].y, R = rr[0] + orr;
            if (__fmaf_rn(dx, dx, dy * dy) <= R * R) m |= 1u << r;   // NaN radius: never
        }
        if (lane >= 16) m &= 0xffffu;                          // r = 16 meets every pair twice
        cm[0][0] = __funnelshift_l(m, m, lane);                // bit r -> bit (lane + r) mod 32
    } else
#pragma unroll
    for (int g = 0; g < AH; ++g) {
        const int jn = min(A - g * 32, 32);
#pragma unroll 2
        for (int jj = 0; jj < jn; ++jj) {
            const float4 o = ws->pos[g * 32 + jj];
#pragma unroll
            for (int h = 0; h < AH; ++h) {
                if (h > g) continue;  // only pairs a < j
                const float dx = o.x - me[h].x, dy = o.y - me[h].y, R = rr[h] + o.z;
                if (__fmaf_rn(dx, dx, dy * dy) <= R * R) cm[h][g] |= 1u << jj;   // NaN radius: never
            }
        }
        cm[g][g] &= ~((2u << lane) - 1u);   // within a group of 32 only j > a (the test above also met j <= a)
    }
    int n = 0;
#pragma unroll
    for (int h = 0; h < AH; ++h)
#pragma unroll
        for (int g = 0; g < AH; ++g) n += __popc(cm[h][g]);
    int incl = n;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int t = __shfl_up_sync(FULL_MASK, incl, d);
        if (lane >= d) incl += t;
    }
    const int total = __shfl_sync(FULL_MASK, incl, 31);
    if (total > 0) {
        if (total <= TDE_PAIR_CAP) {
            int pos = incl - n;
#pragma unroll
            for (int h = 0; h < AH; ++h)
#pragma unroll
                for (int g = 0; g < AH; ++g) {
                    unsigned m = cm[h][g];
                    while (m) {
                        int j = g * 32 + __ffs(m) - 1;
                        m &= m - 1;
                        ws->pairs[pos++] = (unsigned short)((h * 32 + lane) | (j << 8));
                    }
                }
            __syncwarp();
#pragma unroll 1
            for (int i = lane; i < total; i += 32) {
                int pr = ws->pairs[i], a = pr & 255, j = pr >> 8;
                if (tde_overlap(sat_ld(ws, a), sat_ld(ws, j))) { atomicAdd(&ws->cnt[a], 1); atomicAdd(&ws->cnt[j], 1); }
            }
        } else {  // more candidate pairs than the list holds (agents piled up): each lane walks its own bits
#pragma unroll
            for (int h = 0; h < AH; ++h)
#pragma unroll
                for (int g = 0; g < AH; ++g) {
                    unsigned m = cm[h][g];
                    while (m) {
                        int j = g * 32 + __ffs(m) - 1;
                        m &= m - 1;
                        if (tde_overlap(me[h], sat_ld(ws, j))) { atomicAdd(&ws->cnt[h * 32 + lane], 1); atomicAdd(&ws->cnt[j], 1); }
                    }
                }
        }
    }
    __syncwarp();
#pragma unroll
    for (int h = 0; h < AH; ++h) {
        int a = h * 32 + lane;
        cnt_out[h] = a < A ? (float)ws->cnt[a] : 0.0f;
    }
    __syncwarp();
}

// ---------------------------------------------------------------- lane-mesh queries

// Distance queries against the road mesh of map M are exact: every grid cell lists the triangles
// overlapping it first (n_over of them, enough to decide containment) and then every other triangle
// that can be nearest to some point of the cell; points off the grid scan all triangles.  Cells whose
// every point is provably closer to the road than the offroad threshold are flagged SAFE at upload:
// there the offroad term is 0 without looking at a triangle.
#define TDE_CELL_SAFE 0x8000
#define TDE_CELL16_NONE 0x7fff

// What the physics reads of a map: `g` is the map descriptor (the global one, or its staged copy); with STAGED the
// triangle records, stop lines, light schedule and the per-cell summary live in shared memory (bulk-async copies made
// once per CTA, tde_physics_kernel) and tri_s / stop_s / lights_s / cells_s are their shared-window addresses.
// cells_s: one u16 per grid cell - bit 15 = SAFE, bits 0..14 = the overlapping triangle covering most of the cell
// (TDE_CELL16_NONE: no triangle overlaps the cell).  The cell records and candidate lists stay in global memory: an
// agent that drives on its lane needs neither.
struct MapRef { const MapDev* g; uint32_t tri_s, stop_s, lights_s, cells_s; };

// compute_offroad (gym_env.py:142,415,427): sum over corners of max(dist - threshold, 0), and
// compute_wrong_way: max(-cos(psi - lane_dir), 0), min over the triangles under the centre.  Called by
// the whole warp (lane = agent, `mine` = has a box).  The cell records of the four corners and the
// centre are fetched together; each lane then settles its corners alone (safe cell, or inside one of
// the cell's overlapping triangles); the corners that still need a distance are taken one at a time by
// the whole warp, lanes striding over the candidate triangles, min by redux.
struct MeshInfr { float offroad, wrong_way; };
template <bool WRONG_WAY, bool STAGED>
__device__ __noinline__ MeshInfr mesh_infractions_warp(const MapRef R, const Box& b, bool mine, float thr, int lane) {
    const MapDev& M = *R.g;
    MeshInfr out; out.offroad = 0.0f; out.wrong_way = 0.0f;
    if (M.ntri <= 0) return out;
    int2 rec[5];   // x = first item (-1: off the grid), y = n_items << 16 | SAFE | n_overlapping
    float cx[4], cy[4];
    float dc, ds;
    bool ww_settled = !mine;   // STAGED: the centre's summary triangle contains the centre and the agent follows it
#pragma unroll
    for (int k = 0; k < (WRONG_WAY ? 5 : 4); ++k) {
        float px = b.x, py = b.y;
        if (k < 4) { tde_box_corner(b, k, px, py); cx[k] = px; cy[k] = py; }
        float fx = floorf((px - M.gx0) * M.inv_cell), fy = floorf((py - M.gy0) * M.inv_cell);
        bool in_grid = fx >= 0.0f && fy >= 0.0f && fx < (float)M.gnx && fy < (float)M.gny;
        rec[k] = make_int2(in_grid ? 0 : -1, in_grid ? TDE_CELL_SAFE : 0);
        if (mine && in_grid) {
            const int cell = (int)fy * M.gnx + (int)fx;
            bool fetch = true;
            // staged: the per-cell summary (2 bytes in shared memory) settles a SAFE corner and a centre that sits on the lane
            // the agent follows without touching the 8-byte cell records.  (The same summary as a 64 KB global table for the
            // launch that does not stage was measured and dropped: physics 56.0 vs 54.2 us, C4 offroad 366 vs 347 us.)
            if (STAGED) {
                const uint32_t c16 = tde_lds_u16(R.cells_s + 2u * (uint32_t)cell);
                if (k < 4) fetch = !(c16 & TDE_CELL_SAFE);       // a SAFE corner needs nothing else
                else {
                    const int t0 = (int)(c16 & 0x7fffu);
                    if (t0 == TDE_CELL16_NONE) { ww_settled = true; fetch = false; }   // nothing under the centre: 0
                    else {
                        const Tri3 T = tde_load_tri<STAGED>(M.tri, R.tri_s, t0);
                        if (tde_tri_contains(T, b.x, b.y, dc, ds) && fmaxf(-(b.c * dc + b.s * ds), 0.0f) == 0.0f) { ww_settled = true; fetch = false; }
                    }
                }
            }
            if (fetch) rec[k] = __ldg(&M.cell_rec[cell]);
        }
        if (!mine) rec[k] = make_int2(0, TDE_CELL_SAFE);
    }
    float sum = 0.0f;
#pragma unroll 1
    for (int k = 0; k < 4; ++k) {
        const float px = k == 0 ? cx[0] : k == 1 ? cx[1] : k == 2 ? cx[2] : cx[3];
        const float py = k == 0 ? cy[0] : k == 1 ? cy[1] : k == 2 ? cy[2] : cy[3];
        const int2 rc = k == 0 ? rec[0] : k == 1 ? rec[1] : k == 2 ? rec[2] : rec[3];
        float d2 = 0.0f;
        int i0 = rc.x, i1 = rc.x + (int)((unsigned)rc.y >> 16);
        bool need = false;
        if (rc.x < 0) {
            need = true; i0 = 0; i1 = -M.ntri;  // off the grid: every triangle, containment included
        } else if (!(rc.y & TDE_CELL_SAFE)) {
            const int nover = rc.y & 0x7fff;
            need = true;
            for (int i = i0; i < i0 + nover; ++i)
                if (tde_tri_contains(tde_load_tri<STAGED>(M.tri, R.tri_s, (int)__ldg(&M.cell_items[i])), px, py, dc, ds)) { need = false; break; }
        }
        unsigned nm = __ballot_sync(FULL_MASK, need);
        if (nm == 0u) continue;
        if (__popc(nm) > 4) {   // many lanes are off the road (scattered boxes): each walks its own candidates
            if (need) {
                float best = INFINITY;
                bool inside = false;
                if (i1 < 0) {
                    for (int t = 0; t < -i1; ++t) {
                        const Tri3 T = tde_load_tri<STAGED>(M.tri, R.tri_s, t);
                        inside = inside || tde_tri_contains(T, px, py, dc, ds);
                        best = fminf(best, tde_tri_segdist2(T, px, py));
                    }
                } else {
                    for (int i = i0; i < i1; ++i) best = fminf(best, tde_tri_segdist2(tde_load_tri<STAGED>(M.tri, R.tri_s, (int)__ldg(&M.cell_items[i])), px, py));
                }
                d2 = inside ? 0.0f : best;
            }
            nm = 0u;
        }
        while (nm) {
            const int src = __ffs(nm) - 1;
            nm &= nm - 1;
            const float qx = __shfl_sync(FULL_MASK, px, src), qy = __shfl_sync(FULL_MASK, py, src);
            const int a0 = __shfl_sync(FULL_MASK, i0, src), a1 = __shfl_sync(FULL_MASK, i1, src);
            float best = INFINITY;
            bool inside = false;
            if (a1 < 0) {
                for (int t = lane; t < -a1; t += 32) {
                    const Tri3 T = tde_load_tri<STAGED>(M.tri, R.tri_s, t);
                    inside = inside || tde_tri_contains(T, qx, qy, dc, ds);
                    best = fminf(best, tde_tri_segdist2(T, qx, qy));
                }
                inside = __any_sync(FULL_MASK, inside);
            } else {
                for (int i = a0 + lane; i < a1; i += 32) best = fminf(best, tde_tri_segdist2(tde_load_tri<STAGED>(M.tri, R.tri_s, (int)__ldg(&M.cell_items[i])), qx, qy));
            }
            // non-negative binary32 values order like their bit patterns
            const unsigned ub = __reduce_min_sync(FULL_MASK, __float_as_uint(best));
            if (lane == src) d2 = inside ? 0.0f : __uint_as_float(ub);
        }
        float d = sqrtf(d2);
        sum = sum + fmaxf(d - thr, 0.0f);
    }
    out.offroad = mine ? sum : 0.0f;
    if (!WRONG_WAY) return out;
    // wrong way: the triangles under the centre
    float best = INFINITY;
    if (!ww_settled) {
        if (rec[4].x >= 0) {
            const int i0 = rec[4].x, nover = rec[4].y & 0x7fff;
            // the value is a minimum of non-negative terms: a containing triangle the agent follows (term 0) settles it
            for (int i = i0; i < i0 + nover; ++i)
                if (tde_tri_contains(tde_load_tri<STAGED>(M.tri, R.tri_s, (int)__ldg(&M.cell_items[i])), b.x, b.y, dc, ds)) {
                    best = fminf(best, fmaxf(-(b.c * dc + b.s * ds), 0.0f));
                    if (best == 0.0f) break;
                }
        } else {
            for (int t = 0; t < M.ntri; ++t)
                if (tde_tri_contains(tde_load_tri<STAGED>(M.tri, R.tri_s, t), b.x, b.y, dc, ds)) best = fminf(best, fmaxf(-(b.c * dc + b.s * ds), 0.0f));
        }
    }
    out.wrong_way = best == INFINITY ? 0.0f : best;
    return out;
}

__device__ __forceinline__ int light_state_at(const MapDev& M, int step, int phase, int l) {
    if (M.period <= 0 || M.nstop <= 0) return TDE_LIGHT_GREEN;
    int t = (step + phase) % M.period;
    return __ldg(&M.lights[t * M.nstop + l]);
}
// bit l set = stop line l shows red at this env time (lane l looks its own light up)
template <bool STAGED>
__device__ __forceinline__ unsigned red_lights_mask(const MapRef R, int step, int phase, int lane) {
    const MapDev& M = *R.g;
    bool red = false;
    if (lane < M.nstop && M.period > 0) {
        const int t = (step + phase) % M.period;
        const int st = STAGED ? (int)tde_lds_u8(R.lights_s + (uint32_t)(t * M.nstop + lane)) : (int)__ldg(&M.lights[t * M.nstop + lane]);
        red = st == TDE_LIGHT_RED;
    }
    return __ballot_sync(FULL_MASK, red);
}

// TrafficLightControl.compute_violation: rear strip of the agent box vs the stop lines showing red.
// Called by the whole warp; a circle test votes before the 4-axis SAT.
template <bool STAGED>
__device__ __noinline__ float tl_violation_warp(const MapRef R, const Box& b, bool mine, float rear_factor, unsigned red) {
    const MapDev& M = *R.g;
    float length = 2.0f * b.hl;  // exact: hl = 0.5f*length
    float len2 = length * rear_factor;
    float back = 0.5f * (length - len2);
    Box rear = b;
    rear.x = b.x - back * b.c;
    rear.y = b.y - back * b.s;
    rear.hl = 0.5f * len2;
    const float rr = rear.hl + rear.hw + 0.005f;
    float viol = 0.0f;
    while (red) {
        const int l = __ffs(red) - 1;
        red &= red - 1;
        float4 u, v;
        if (STAGED) { u = tde_lds_f4(R.stop_s + 32u * (uint32_t)l); v = tde_lds_f4(R.stop_s + 32u * (uint32_t)l + 16u); }
        else { u = __ldg(&M.stop[2 * l]); v = __ldg(&M.stop[2 * l + 1]); }
        float dx = u.x - rear.x, dy = u.y - rear.y, Rr = rr + v.z;
        bool cand = mine && dx * dx + dy * dy <= Rr * Rr * 1.001f;
        if (__any_sync(FULL_MASK, cand)) {
            Box sb; sb.x = u.x; sb.y = u.y; sb.hl = u.z; sb.hw = u.w; sb.c = v.x; sb.s = v.y; sb.present = 1.0f; sb.r = 0.0f;
            if (cand && tde_overlap(rear, sb)) viol += 1.0f;
        }
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
    float2 p0 = __ldg(&S.wp[0]);
    float2 p1 = S.W > 1 ? __ldg(&S.wp[1]) : p0;
#pragma unroll
    for (int h = 0; h < AH; ++h) {
        int a = h * 32 + lane;
        if (a < p.A) {
            float4 init = __ldg(&S.init[a]);
            if (a >= 1 && S.rep_T > 0 && __ldg(&S.rep_mask[a])) init = __ldg(&S.rep_states[a]);
            float4 attr = __ldg(&S.attr[a]);
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
    if (lane == 0) { p.ep_return[e] = 0.0f; p.restart[e] = 1; }
}

// the eight env variables are warp-uniform: lane 0 writes the row as two 128-bit stores
struct EnvVars { int s, step, target, reached, lphase, episode, m; };
// BCAST: every lane loads the row itself (one broadcast transaction); otherwise lanes 0..7 load one variable each and
// shuffle.  Measured (C3): the broadcast is faster in the physics kernel (54.0 vs 54.5 us), the shuffles in the render
// kernel (135.4 vs 137.6 us).
template <bool BCAST>
__device__ __forceinline__ EnvVars load_vars(const StepParams& p, int e, int lane) {
    EnvVars v;
    if (BCAST) {
        const int4* row = reinterpret_cast<const int4*>(p.vars) + (size_t)e * 2;
        const int4 a = row[0], b = row[1];
        v.s = a.x; v.step = a.y; v.target = a.z; v.reached = a.w; v.lphase = b.x; v.episode = b.y; v.m = b.z;
    } else {
        const int myvar = lane < 8 ? p.vars[(size_t)e * 8 + lane] : 0;
        v.s = __shfl_sync(FULL_MASK, myvar, 0); v.step = __shfl_sync(FULL_MASK, myvar, 1); v.target = __shfl_sync(FULL_MASK, myvar, 2);
        v.reached = __shfl_sync(FULL_MASK, myvar, 3); v.lphase = __shfl_sync(FULL_MASK, myvar, 4);
        v.episode = __shfl_sync(FULL_MASK, myvar, 5); v.m = __shfl_sync(FULL_MASK, myvar, 6);
    }
    return v;
}
__device__ __forceinline__ void store_vars(const StepParams& p, int e, int lane, int s, int step, int target,
                                           int reached, int lphase, int episode, int m) {
    if (lane == 0) {
        int4* row = reinterpret_cast<int4*>(p.vars) + (size_t)e * 2;
        row[0] = make_int4(s, step, target, reached);
        row[1] = make_int4(lphase, episode, m, 0);
    }
}

#include "tde_render.cuh"

// terminal observations: rows of the envs flagged in `mask` are copied out before those envs are re-initialised
__global__ void __launch_bounds__(256) tde_copy_rows_kernel(const uint8_t* __restrict__ mask, const uint4* __restrict__ src,
                                                            uint4* __restrict__ dst, int E, int u4_per_env) {
    const int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int warps_total = (gridDim.x * blockDim.x) >> 5;
    for (int e = warp; e < E; e += warps_total) {
        if (mask[e] == 0) continue;
        const uint4* s = src + (size_t)e * u4_per_env;
        uint4* d = dst + (size_t)e * u4_per_env;
        for (int i = lane; i < u4_per_env; i += 32) d[i] = s[i];
    }
}

// ---------------------------------------------------------------- physics kernel

// WaypointSuiteEnv.step :369-389 minus the observation: bicycle step / NPC replay, all-pairs SAT,
// offroad / red-light / wrong-way against the lane mesh, reward, termination, truncation, info,
// waypoint progress, episode statistics and (optionally) the in-kernel auto-reset.
template <int AH, bool STAGED>
__device__ __forceinline__ void physics_env(const StepParams& p, const int e, const int lane, SatScratch<AH>* ws, double& st_acc, int& n_steps,
                                            unsigned char* smem_raw, unsigned long long* mbar, bool& staged_ready) {
    const tde_config& c = p.cfg;
    {
        const EnvVars ev = load_vars<true>(p, e, lane);
        int s = ev.s, step = ev.step, target = ev.target, reached = ev.reached, lphase = ev.lphase, episode = ev.episode, m = ev.m;
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
            const float2 act = reinterpret_cast<const float2*>(p.actions)[e];
            const float act_a = act.x, act_b = act.y;
#pragma unroll
            for (int h = 0; h < AH; ++h) {
                int a = h * 32 + lane;
                if (a < p.A && at[h].w != 0.0f) {
                    bool replay = a != 0 && t < S.rep_T && __ldg(&S.rep_mask[(size_t)t * p.A + a]);
                    if (replay) {
                        st[h] = __ldg(&S.rep_states[(size_t)t * p.A + a]);
                    } else {
                        st[h] = tde_bicycle(st[h], a == 0 ? act_a : 0.0f, a == 0 ? act_b : 0.0f, at[h].z, c.dt);
                    }
                    p.state[(size_t)e * p.A + a] = st[h];
                }
            }
            step = t;
        }

        float4 inf0 = make_float4(0.f, 0.f, 0.f, 0.f);  // ego's infractions, valid on lane 0
        if (p.phases & TDE_PH_INFRACTIONS) {
            Box me[AH];
            float cnt[AH];
#pragma unroll
            for (int h = 0; h < AH; ++h) me[h] = tde_make_box(st[h].x, st[h].y, st[h].z, at[h].x, at[h].y, at[h].w);
            sat_counts<AH>(ws, me, p.A, lane, cnt);
            MapRef M;
            if (STAGED) {
                // the bulk copies were issued when the CTA started; the first env of a warp waits for them here, after
                // its bicycle step and SAT
                if (!staged_ready) { tde_mbar_wait(mbar, 0u); staged_ready = true; }
                const uint4 hd = reinterpret_cast<const uint4*>(smem_raw)[m];
                const uint32_t base = tde_smem_addr(smem_raw);
                M.g = reinterpret_cast<const MapDev*>(smem_raw + p.stage_maps_off) + m;
                M.tri_s = base + hd.x; M.stop_s = base + hd.y; M.lights_s = base + hd.z; M.cells_s = base + hd.w;
            } else {
                M.g = &p.maps[m];
                M.tri_s = M.stop_s = M.lights_s = M.cells_s = 0u;
            }
            const unsigned red = red_lights_mask<STAGED>(M, step, lphase, lane);
#pragma unroll
            for (int h = 0; h < AH; ++h) {
                int a = h * 32 + lane;
                bool mine = a < p.A && at[h].w != 0.0f;
                float4 inf = make_float4(0.f, 0.f, 0.f, 0.f);
                MeshInfr mi = mesh_infractions_warp<true, STAGED>(M, me[h], mine, c.offroad_threshold, lane);
                float tl = tl_violation_warp<STAGED>(M, me[h], mine, c.tl_rear_factor, red);
                if (mine) {
                    inf.x = cnt[h];
                    inf.y = mi.offroad;
                    inf.z = tl;
                    inf.w = mi.wrong_way;
                }
                if (a < p.A) p.infr[(size_t)e * p.A + a] = inf;
                if (h == 0) inf0 = inf;
            }
            __syncwarp();
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
                float2 w = __ldg(&S.wp[target]);
                float tx = x - w.x, ty = y - w.y;
                hit = sqrtf(tx * tx + ty * ty) < c.reach_radius;
            }
            float reach_reward = hit ? c.waypoint_bonus : 0.0f;                      // :404-408
            if (hit) reached += 1;                                                   // :406
            float r = (reach_reward + dist_reward) + psi_reward;                     // :410
            bool term = c.terminated_at_infraction && (i_off > 0.0f || i_col > 0.0f || i_tl > 0.0f);  // :413-417
            bool trunc = step >= c.max_environment_steps;                            // :134-135
            // lane 0 owns the env's return accumulator (it also writes it back below): the other lanes get the value by shuffle
            const float ep_ret = __shfl_sync(FULL_MASK, lane == 0 ? p.ep_return[e] + r : 0.0f, 0);
            bool done = term || trunc;
            // info row (get_info :419-437): every value is warp-uniform, lane 0 writes the row as four 128-bit stores
            const float did_reset = (done && c.auto_reset) ? 1.0f : 0.0f;
            if (lane == 0) {
                float4* row = reinterpret_cast<float4*>(p.info + (size_t)e * TDE_INFO_STRIDE);
                static_assert(TDE_INFO_OFFROAD == 0 && TDE_INFO_COLLISION == 1 && TDE_INFO_TL_VIOLATION == 2 && TDE_INFO_IS_SUCCESS == 3, "info layout");
                static_assert(TDE_INFO_REACHED_WAYPOINT_NUM == 4 && TDE_INFO_PSI_SMOOTHNESS == 5 && TDE_INFO_PSI_REWARD == 6 && TDE_INFO_DIST_REWARD == 7, "info layout");
                static_assert(TDE_INFO_SPEED_SMOOTHNESS == 8 && TDE_INFO_WRONG_WAY == 9 && TDE_INFO_EPISODE_RETURN == 10 && TDE_INFO_EPISODE_LENGTH == 11, "info layout");
                static_assert(TDE_INFO_SCENARIO == 12 && TDE_INFO_DID_RESET == 13 && TDE_INFO_STRIDE == 16, "info layout");
                row[0] = make_float4(i_off, i_col, i_tl, trunc ? 1.0f : 0.0f);
                row[1] = make_float4((float)reached, fabsf((lpsi - psi) / c.dt), psi_reward, dist_reward);
                row[2] = make_float4(fabsf((lv - spd) / c.dt), i_ww, ep_ret, (float)step);
                row[3] = make_float4((float)s, did_reset, 0.0f, 0.0f);
                p.reward[e] = r;
                p.terminated[e] = term ? 1 : 0;
                p.truncated[e] = trunc ? 1 : 0;
                p.ep_return[e] = ep_ret;
            }
            if (hit) target += 1;                                                    // :378-383
            // episode statistics: lane k owns statistic k; steps are counted as an integer and added at the end
            n_steps += 1;
            if (done) {
                const double add = lane == TDE_STAT_EPISODES ? 1.0
                                 : lane == TDE_STAT_RETURN_SUM ? (double)ep_ret
                                 : lane == TDE_STAT_LENGTH_SUM ? (double)step
                                 : lane == TDE_STAT_OFFROAD ? (i_off > 0.0f ? 1.0 : 0.0)
                                 : lane == TDE_STAT_COLLISION ? (i_col > 0.0f ? 1.0 : 0.0)
                                 : lane == TDE_STAT_TL_VIOLATION ? (i_tl > 0.0f ? 1.0 : 0.0)
                                 : lane == TDE_STAT_SUCCESS ? (trunc ? 1.0 : 0.0)
                                 : lane == TDE_STAT_REACHED_WAYPOINTS ? (double)reached : 0.0;
                st_acc += add;
            }
            if (done && c.auto_reset) {
                __syncwarp();
                if (p.done_mask != nullptr) { if (lane == 0) p.done_mask[e] = 1; }   // reset deferred until the terminal frame is out
                else reset_env_warp<AH>(p, e, lane, s, step, target, reached, lphase, episode, m, st, at);
            }
        }
        if (p.phases & (TDE_PH_KINEMATICS | TDE_PH_REWARD)) store_vars(p, e, lane, s, step, target, reached, lphase, episode, m);
        __syncwarp();
    }
}

// Shared memory of the physics kernel: [staged map tables (STAGED only)] [mbarrier, 16 B] [SatScratch per warp].  The
// number of warps per CTA is the launch's choice (blockDim): 4-warp CTAs when the tables stay in global memory (the rest
// of the SM is L1 for them), up to 32-warp CTAs when they are staged (one copy per SM).
#define TDE_PHYS_STAGE_CHUNK 32768u
template <int AH, bool STAGED>
__global__ void __launch_bounds__(STAGED ? TDE_PHYS_STAGED_MAX_WARPS * 32 : TDE_WARPS_PER_BLOCK * 32, STAGED ? 1 : TDE_PHYS_BLOCKS_PER_SM) tde_physics_kernel(const StepParams p) {
    TDE_DYN_SMEM(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wpb = blockDim.x >> 5;
    const unsigned int stage_bytes = STAGED ? p.stage_bytes : 0u;
    unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(smem_raw + stage_bytes);
    SatScratch<AH>* ws = reinterpret_cast<SatScratch<AH>*>(smem_raw + stage_bytes + 16) + warp;
    bool staged_ready = !STAGED;
    if (STAGED) {
        // per-scenario lane mesh & co staged once per CTA by the TMA engine: cp.async.bulk + mbarrier (complete_tx)
        if (threadIdx.x == 0) tde_mbar_init(mbar, 1);
        __syncthreads();
        if (threadIdx.x == 0) {
            tde_mbar_expect_tx(mbar, stage_bytes);
            for (unsigned int off = 0; off < p.stage_bytes; off += TDE_PHYS_STAGE_CHUNK)
                tde_bulk_g2s(smem_raw + off, p.stage_blob + off, min(TDE_PHYS_STAGE_CHUNK, p.stage_bytes - off), mbar);
        }
    }
    const int warps_total = gridDim.x * wpb;
    tde_pdl_wait();
    double st_acc = 0.0;  // lane k accumulates statistic k
    int n_steps = 0;      // env steps taken by this warp
#pragma unroll 1
    for (int e = p.e_begin + blockIdx.x * wpb + warp; e < p.e_end; e += warps_total) {
        TDE_TRACE_MARK(e, 2);
        physics_env<AH, STAGED>(p, e, lane, ws, st_acc, n_steps, smem_raw, mbar, staged_ready);
        TDE_TRACE_MARK(e, 3);
    }
    tde_pdl_launch_dependents();
    if (lane == TDE_STAT_STEPS) st_acc += (double)n_steps;
    if ((p.phases & TDE_PH_REWARD) && lane < TDE_NUM_STATS && st_acc != 0.0) atomicAdd(&p.stats[lane], st_acc);
    // a CTA must not exit while its bulk copies are in flight (a warp without envs, or one that skipped the infractions)
    if (STAGED && !staged_ready) tde_mbar_wait(mbar, 0u);
}

template <int AH>
__global__ void __launch_bounds__(TDE_WARPS_PER_BLOCK * 32) tde_reset_kernel(const StepParams p) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_total = gridDim.x * TDE_WARPS_PER_BLOCK;
    for (int e = blockIdx.x * TDE_WARPS_PER_BLOCK + warp; e < p.E; e += warps_total) {
        if (p.reset_mask != nullptr && p.reset_mask[e] == 0) continue;
        int s, step, target, reached, lphase, m;
        int episode = load_vars<true>(p, e, lane).episode;
        __syncwarp();   // every lane has read the row before lane 0 rewrites it below
        float4 st[AH], at[AH];
        reset_env_warp<AH>(p, e, lane, s, step, target, reached, lphase, episode, m, st, at);
        store_vars(p, e, lane, s, step, target, reached, lphase, episode, m);
    }
}

// ---------------------------------------------------------------- stateless micro-benchmark kernels (config C4)

// All-pairs oriented-box collision counts on caller-provided boxes: warp per env, the same broad phase
// + compacted exact SAT as the step kernel.
template <int AH>
__global__ void __launch_bounds__(TDE_WARPS_PER_BLOCK * 32) tde_collision_kernel(const float4* __restrict__ state,
                                                                                  const float4* __restrict__ attr, int E, int A,
                                                                                  float* __restrict__ out) {
    __shared__ SatScratch<AH> scratch[TDE_WARPS_PER_BLOCK];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int warps_total = gridDim.x * TDE_WARPS_PER_BLOCK;
    SatScratch<AH>* ws = &scratch[warp];
    for (int e = blockIdx.x * TDE_WARPS_PER_BLOCK + warp; e < E; e += warps_total) {
        Box me[AH];
        float cnt[AH];
#pragma unroll
        for (int h = 0; h < AH; ++h) {
            int a = h * 32 + lane;
            float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = make_float4(1.f, 1.f, 1.f, 0.f);
            if (a < A) { s4 = state[(size_t)e * A + a]; a4 = attr[(size_t)e * A + a]; }
            me[h] = tde_make_box(s4.x, s4.y, s4.z, a4.x, a4.y, a4.w);
        }
        sat_counts<AH>(ws, me, A, lane, cnt);
#pragma unroll
        for (int h = 0; h < AH; ++h) {
            int a = h * 32 + lane;
            if (a < A) out[(size_t)e * A + a] = me[h].present != 0.0f ? cnt[h] : 0.0f;
        }
    }
}

// Stateless offroad (config C4: scattered boxes, half of the corners off the road).  Walking each corner's candidate
// list in its own lane leaves 12 of 32 lanes busy (the lists differ in length and half the lanes have none), so the
// (corner, candidate) pairs of a warp's 32 boxes are numbered consecutively and taken 32 at a time with every lane
// busy, the way the rasteriser takes (primitive, row) items: a scan over the list lengths, a binary search for the
// corner an item belongs to, a shared-memory atomicMin per corner (non-negative binary32 values order like their bit
// patterns).  Same candidates, same per-candidate arithmetic, a minimum and a sum in the same order: same bits as
// mesh_infractions_warp.
struct OffroadScratch {        // per warp; corner k of lane l: start[4 l + k] (one 128-bit row per lane), the other arrays [32 k + l] (bank = lane)
    uint32_t start[128];       // candidates before the corner
    int base[128];             // first item of the corner's list, or -1: every triangle of the map, containment included
    float px[128], py[128];
    uint32_t best[128];        // min squared distance so far (bit pattern)
    unsigned short nover[128]; // the first nover candidates of the corner overlap its cell: the ones that can contain it
    uint32_t inside[4];        // corners found inside a triangle, bit l of word k
};
#define TDE_OFFROAD_WARPS 8
__global__ void __launch_bounds__(TDE_OFFROAD_WARPS * 32) tde_offroad_kernel(const MapDev* maps, int map_id, float thr,
                                                                              const float4* __restrict__ state, const float4* __restrict__ attr,
                                                                              int n, float* __restrict__ out) {
    __shared__ OffroadScratch scratch[TDE_OFFROAD_WARPS];
    const MapDev& M = maps[map_id];
    const int lane = threadIdx.x & 31;
    OffroadScratch* ws = &scratch[threadIdx.x >> 5];
    const int stride = gridDim.x * blockDim.x;
    for (int first = blockIdx.x * blockDim.x + (threadIdx.x & ~31); first < n; first += stride) {  // warp-uniform trip count
        const int i = first + lane;
        float4 s4 = make_float4(0.f, 0.f, 0.f, 0.f), a4 = make_float4(1.f, 1.f, 1.f, 0.f);
        if (i < n) { s4 = state[i]; a4 = attr[i]; }
        const Box b = tde_make_box(s4.x, s4.y, s4.z, a4.x, a4.y, a4.w);
        const bool mine = i < n && a4.w != 0.0f && M.ntri > 0;
        uint32_t cnt[4];
        unsigned need = 0u;
        if (lane < 4) ws->inside[lane] = 0u;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float px, py;
            tde_box_corner(b, k, px, py);
            cnt[k] = 0u;
            int bs = 0, nov = 0;
            if (mine) {
                const float fx = floorf((px - M.gx0) * M.inv_cell), fy = floorf((py - M.gy0) * M.inv_cell);
                if (!(fx >= 0.0f && fy >= 0.0f && fx < (float)M.gnx && fy < (float)M.gny)) {
                    need |= 1u << k; cnt[k] = (uint32_t)M.ntri; bs = -1;     // off the grid: every triangle, containment included
                } else {
                    const int2 rec = __ldg(&M.cell_rec[(int)fy * M.gnx + (int)fx]);
                    // a cell that is not SAFE: all its candidates become items; the overlapping ones (listed first) are
                    // also tested for containment, and a corner found inside ignores its distance
                    if (!(rec.y & TDE_CELL_SAFE)) { need |= 1u << k; cnt[k] = (uint32_t)rec.y >> 16; bs = rec.x; nov = rec.y & 0x7fff; }
                }
            }
            const int c = 32 * k + lane;
            ws->base[c] = bs; ws->px[c] = px; ws->py[c] = py; ws->best[c] = 0x7f800000u; ws->nover[c] = (unsigned short)nov;
        }
        const uint32_t mine_total = cnt[0] + cnt[1] + cnt[2] + cnt[3];
        uint32_t incl = mine_total;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(FULL_MASK, incl, d);
            if (lane >= d) incl += t;
        }
        const uint32_t total = __shfl_sync(FULL_MASK, incl, 31);
        const uint32_t excl = incl - mine_total;
        *reinterpret_cast<uint4*>(&ws->start[4 * lane]) = make_uint4(excl, excl + cnt[0], excl + cnt[0] + cnt[1], excl + cnt[0] + cnt[1] + cnt[2]);
        __syncwarp();
        for (uint32_t w0 = 0; w0 < total; w0 += 32) {   // warp-uniform trip count: the search shuffles
            const uint32_t w = min(w0 + lane, total - 1u);   // surplus lanes of the last pass repeat its last item (min is idempotent)
            // the last corner that starts at or before item w (corners without candidates share their start with the next
            // one): the box by a binary search over the lanes' running totals, then one of its four corners
            int o = 0;
#pragma unroll
            for (int st = 16; st >= 1; st >>= 1) {
                const uint32_t v = __shfl_sync(FULL_MASK, excl, (o + st) & 31);
                if (v <= w) o += st;
            }
            const uint4 s4c = *reinterpret_cast<const uint4*>(&ws->start[4 * o]);
            const int kk = (w >= s4c.y) + (w >= s4c.z) + (w >= s4c.w);
            const int c = 32 * kk + o;
            const int j = (int)(w - (kk == 0 ? s4c.x : kk == 1 ? s4c.y : kk == 2 ? s4c.z : s4c.w)), bs = ws->base[c];
            const float qx = ws->px[c], qy = ws->py[c];
            const Tri3 T = tde_load_tri<false>(M.tri, 0u, bs < 0 ? j : (int)__ldg(&M.cell_items[bs + j]));
            float dc, ds;
            if ((bs < 0 || j < (int)ws->nover[c]) && tde_tri_contains(T, qx, qy, dc, ds)) atomicOr(&ws->inside[c >> 5], 1u << (c & 31));
            atomicMin(&ws->best[c], __float_as_uint(tde_tri_segdist2(T, qx, qy)));
        }
        __syncwarp();
        float sum = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = 32 * k + lane;
            float d2 = 0.0f;
            if (need & (1u << k)) d2 = (ws->inside[c >> 5] >> (c & 31)) & 1u ? 0.0f : __uint_as_float(ws->best[c]);
            sum = sum + fmaxf(sqrtf(d2) - thr, 0.0f);
        }
        if (i < n) out[i] = mine ? sum : 0.0f;
        __syncwarp();   // the scratch is rewritten by the next round
    }
}
