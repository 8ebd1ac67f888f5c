/*
 * tde_oracle.c — CPU restatement of TorchDriveEnv's per-timestep simulation path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (torchdriveenv_b200/) may import, link or call
 * this file; it is used by tests/, by __graft_entry__.smoke() and by bench.py's cpu_baseline /
 * --impl reference legs as the checker and as the timed CPU baseline.
 *
 * PARITY: partly pinned.  The arithmetic of rows a2-a8 lives in the third-party package torchdrivesim
 * (pinned >=0.2.1, pyproject.toml:30; git 6c7957c780404980d9f69a00b40cb98eab0a87d5,
 * requirements.txt:74) which is neither vendored under /root/reference nor installable offline,
 * and the reference ships no tests or golden vectors.  What is restated here:
 *   - exactly, from the reference's own source: step ordering, reward, termination, truncation,
 *     info, waypoint progress, reset sampling and replay conventions (torchdriveenv/gym_env.py,
 *     cited per function below).  PINNED: the reference's own gym_env.py, imported unmodified and
 *     run over a SimulatorInterface-level surface backed by this oracle, produced the golden
 *     vectors tests/golden/ref_*.npz (tests/golden/make_reference_golden.py); the oracle and the
 *     CUDA path are compared with them in tests/test_reference_golden.py and
 *     tests/test_gpu_reference_golden.py;
 *   - from the published torchdrivesim algorithms as recalled in SURVEY.md §8a (a2-a8), with every
 *     open decision fixed in DESIGN.md §SPEC: kinematic bicycle, oriented-box overlap, corner-to-
 *     mesh offroad distance, wrong-way, red-light stop-line overlap, egocentric birdview.
 *     PARITY UNPINNED for these rows; they are pinned instead by independent property tests
 *     (tests/test_oracle_*.py: float64 closed forms, cv2.rotatedRectangleIntersection,
 *     cv2.fillPoly, brute-force float64 distances).
 *
 * Arithmetic contract (what makes CPU/GPU comparisons bit-exact): IEEE binary32, one rounding per
 * written operation, no contraction (-ffp-contract=off here, -fmad=false in nvcc) except where
 * fmaf() is written explicitly; sin/cos come from orc_sincosf below, not libm.
 *
 * Build: gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -mfma -mavx2 (see oracle/Makefile).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/tde_b200.h"

#define ORC_PI_F 3.14159274101257324219f
#define ORC_TWO_PI_F 6.28318548202514648438f

/* ------------------------------------------------------------------ math */

/* Deterministic sin/cos: 3-term Cody-Waite reduction by pi/2 and the cephes single-precision
   minimax polynomials, every step an explicit fmaf so the GPU can reproduce it bit for bit.
   Max error vs float64: < 1.5 ulp for |x| <= 1e4 (tests/test_oracle_math.py).
   Stands in for torch.cos/torch.sin in KinematicBicycle.step and math.cos in get_reward :403. */
void orc_sincosf(float x, float* s_out, float* c_out) {
    float kf = rintf(x * 0.636619772367581343f);
    if (!(fabsf(kf) < 1.0e9f)) kf = 0.0f; /* inf/nan/huge: no reduction */
    int32_t k = (int32_t)kf;
    float r = fmaf(kf, -1.57079625129699707031e+00f, x);
    r = fmaf(kf, -7.54978941586159635335e-08f, r);
    r = fmaf(kf, -5.39030285815811905290e-15f, r);
    float z = r * r;
    float sp = fmaf(-1.9515295891e-4f, z, 8.3321608736e-3f);
    sp = fmaf(sp, z, -1.6666654611e-1f);
    float sr = fmaf(sp * z, r, r);
    float cp = fmaf(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    cp = fmaf(cp, z, 4.166664568298827e-2f);
    float cr = fmaf(cp * z, z, fmaf(-0.5f, z, 1.0f));
    float s, c;
    switch (k & 3) {
        case 0: s = sr; c = cr; break;
        case 1: s = cr; c = -sr; break;
        case 2: s = -sr; c = -cr; break;
        default: s = -cr; c = sr; break;
    }
    *s_out = s;
    *c_out = c;
}

/* psi <- ((pi + psi) mod 2pi) - pi with floored (Python/torch.remainder) modulo. */
float orc_wrap_pi(float psi) {
    float t = psi + ORC_PI_F;
    float m = fmodf(t, ORC_TWO_PI_F);
    if (m < 0.0f) m += ORC_TWO_PI_F;
    return m - ORC_PI_F;
}

/* KinematicBicycle.step (torchdrivesim, called through simulator.step gym_env.py:117; model built
   with defaults :245 and lr = rear-axis offset :246).  [EXT-RECALLED] update order: new speed first,
   position with the new speed along psi+beta, yaw rate v/lr*sin(beta), then the wrap. */
void orc_bicycle_step(float* st, float a, float beta, float lr, float dt) {
    float x = st[0], y = st[1], psi = st[2], v = st[3];
    float s1, c1, sb, cb;
    v = v + a * dt;
    orc_sincosf(psi + beta, &s1, &c1);
    orc_sincosf(beta, &sb, &cb);
    x = x + (v * c1) * dt;
    y = y + (v * s1) * dt;
    psi = psi + ((v / lr) * sb) * dt;
    psi = orc_wrap_pi(psi);
    st[0] = x; st[1] = y; st[2] = psi; st[3] = v;
}

typedef struct { float x, y, hl, hw, c, s; } orc_box;

static orc_box make_box(float x, float y, float psi, float length, float width) {
    orc_box b;
    b.x = x; b.y = y; b.hl = 0.5f * length; b.hw = 0.5f * width;
    orc_sincosf(psi, &b.s, &b.c);
    return b;
}

/* Oriented-rectangle overlap with positive area: separating-axis test over the four face normals.
   CollisionMetric.nograd (gym_env.py:48) [EXT-RECALLED]: exact, touching is not a collision. */
int orc_box_overlap(const orc_box* A, const orc_box* B) {
    float dx = B->x - A->x, dy = B->y - A->y;
    float cc = A->c * B->c + A->s * B->s;
    float ss = A->c * B->s - A->s * B->c;
    float acc = fabsf(cc), ass = fabsf(ss);
    float p, lim;
    p = fabsf(dx * A->c + dy * A->s); lim = A->hl + (B->hl * acc + B->hw * ass);
    if (!(p < lim)) return 0;
    p = fabsf(dy * A->c - dx * A->s); lim = A->hw + (B->hl * ass + B->hw * acc);
    if (!(p < lim)) return 0;
    p = fabsf(dx * B->c + dy * B->s); lim = B->hl + (A->hl * acc + A->hw * ass);
    if (!(p < lim)) return 0;
    p = fabsf(dy * B->c - dx * B->s); lim = B->hw + (A->hl * ass + A->hw * acc);
    if (!(p < lim)) return 0;
    return 1;
}

/* smallest slack over the four axes (>0 overlap, <0 separated): the test harness' epsilon band */
float orc_box_margin(const orc_box* A, const orc_box* B) {
    float dx = B->x - A->x, dy = B->y - A->y;
    float cc = A->c * B->c + A->s * B->s;
    float ss = A->c * B->s - A->s * B->c;
    float acc = fabsf(cc), ass = fabsf(ss);
    float m = A->hl + (B->hl * acc + B->hw * ass) - fabsf(dx * A->c + dy * A->s);
    float t = A->hw + (B->hl * ass + B->hw * acc) - fabsf(dy * A->c - dx * A->s); if (t < m) m = t;
    t = B->hl + (A->hl * acc + A->hw * ass) - fabsf(dx * B->c + dy * B->s); if (t < m) m = t;
    t = B->hw + (A->hl * ass + A->hw * acc) - fabsf(dy * B->c - dx * B->s); if (t < m) m = t;
    return m;
}

/* corner k of a box: (+hl,+hw) (+hl,-hw) (-hl,-hw) (-hl,+hw) */
static void box_corner(const orc_box* b, int k, float* px, float* py) {
    float ox = (k == 0 || k == 1) ? b->hl : -b->hl;
    float oy = (k == 0 || k == 3) ? b->hw : -b->hw;
    *px = b->x + (ox * b->c - oy * b->s);
    *py = b->y + (ox * b->s + oy * b->c);
}

/* one triangle edge: squared distance from p to segment a->b and the edge function */
static void edge_terms(float ax, float ay, float bx, float by, float il, float px, float py, float* d2,
                       float* cr) {
    float abx = bx - ax, aby = by - ay, apx = px - ax, apy = py - ay;
    *cr = abx * apy - aby * apx;
    float t = (apx * abx + apy * aby) * il;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    float qx = apx - t * abx, qy = apy - t * aby;
    *d2 = qx * qx + qy * qy;
}

static float inv_len2(float ax, float ay, float bx, float by) {
    float abx = bx - ax, aby = by - ay;
    float l2 = abx * abx + aby * aby;
    return l2 > 0.0f ? 1.0f / l2 : 0.0f;
}

/* squared distance from p to a triangle (0 inside or on the boundary); *inside reports containment */
static float point_tri_dist2(const float* t, float px, float py, int* inside) {
    float d0, d1, d2, c0, c1, c2;
    edge_terms(t[0], t[1], t[2], t[3], inv_len2(t[0], t[1], t[2], t[3]), px, py, &d0, &c0);
    edge_terms(t[2], t[3], t[4], t[5], inv_len2(t[2], t[3], t[4], t[5]), px, py, &d1, &c1);
    edge_terms(t[4], t[5], t[0], t[1], inv_len2(t[4], t[5], t[0], t[1]), px, py, &d2, &c2);
    int in = (c0 >= 0.0f && c1 >= 0.0f && c2 >= 0.0f) || (c0 <= 0.0f && c1 <= 0.0f && c2 <= 0.0f);
    *inside = in;
    if (in) return 0.0f;
    return fminf(fminf(d0, d1), d2);
}

/* compute_offroad (gym_env.py:142,415,427) [EXT-RECALLED]: sum over the four box corners of
   max(distance to the road mesh - threshold, 0); brute force over every triangle of the map. */
float orc_offroad_box(const orc_box* b, const float* tris, int ntri, float threshold) {
    if (ntri <= 0) return 0.0f;
    float sum = 0.0f;
    for (int k = 0; k < 4; ++k) {
        float px, py;
        box_corner(b, k, &px, &py);
        float best = INFINITY;
        for (int t = 0; t < ntri; ++t) {
            int in;
            float d2 = point_tri_dist2(tris + 8 * t, px, py, &in);
            if (d2 < best) best = d2;
        }
        float d = sqrtf(best);
        sum = sum + fmaxf(d - threshold, 0.0f);
    }
    return sum;
}

/* compute_wrong_way (SimulatorInterface; not called by the reference env) [EXT-RECALLED]:
   max(-cos(psi - lane_dir), 0) for the lane triangle under the agent centre (min over overlapping
   triangles, 0 when the centre is on no triangle): > 0 iff heading is > 90 deg off the lane. */
float orc_wrong_way_box(const orc_box* b, const float* tris, int ntri) {
    float best = INFINITY;
    for (int t = 0; t < ntri; ++t) {
        int in;
        (void)point_tri_dist2(tris + 8 * t, b->x, b->y, &in);
        if (in) {
            float cosd = b->c * tris[8 * t + 6] + b->s * tris[8 * t + 7];
            float loss = fmaxf(-cosd, 0.0f);
            if (loss < best) best = loss;
        }
    }
    return best == INFINITY ? 0.0f : best;
}

/* ------------------------------------------------------------------ RNG */

static uint64_t mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
/* counter-based: (seed, global env, episode, k) -> 64 random bits */
uint64_t orc_rng(uint64_t seed, uint64_t genv, uint64_t episode, uint64_t k) {
    uint64_t x = seed + 0x9E3779B97F4A7C15ULL * (genv + 1);
    x = mix64(x);
    x = mix64(x + 0x9E3779B97F4A7C15ULL * (episode + 1));
    x = mix64(x + 0x9E3779B97F4A7C15ULL * (k + 1));
    return x;
}
static float u01(uint64_t r) { return (float)(uint32_t)(r >> 40) * 5.9604644775390625e-08f; } /* 2^-24 */
/* approx N(0,1): scaled Irwin-Hall sum of eight 16-bit uniforms (two draws) */
static float normal8(uint64_t r0, uint64_t r1) {
    uint32_t sum = 0;
    for (int i = 0; i < 4; ++i) { sum += (uint32_t)((r0 >> (16 * i)) & 0xFFFF); sum += (uint32_t)((r1 >> (16 * i)) & 0xFFFF); }
    float u = ((float)sum + 4.0f) * 1.52587890625e-05f - 4.0f; /* mean 0, var 8/12 */
    return u * 1.22474487139158894f;                           /* sqrt(12/8) */
}

/* ------------------------------------------------------------------ env container */

typedef struct orc_env_set {
    tde_config cfg;
    int A, E;
    /* maps */
    int num_maps; int32_t *tri_off, *mark_off, *stop_off, *light_period, *light_off;
    float *road_tris, *mark_tris, *stoplines; uint8_t* light_states;
    /* scenarios */
    int num_scen; int32_t *scen_map, *wp_off, *scen_nag, *rep_T, *rep_off;
    float *waypoints, *start_heading, *agent_init, *agent_attr, *replay_states; uint8_t* replay_mask;
    int32_t *scen_lo, *scen_hi;
    /* per env */
    float* state;   /* E*A*4 */
    float* attr;    /* E*A*4 */
    float* infr;    /* E*A*4 */
    int32_t* vars;  /* E*8: scenario, step, target, reached, light_phase, episode, map, reserved */
    uint8_t* terminal_obs; /* optional E*3*64*64: frame of the final state of an env, rendered before its auto-reset
                              (what SB3's VecEnv puts in info["terminal_observation"]); not owned */
    float* ep_return; /* E */
    uint64_t seed;
    double stats[TDE_NUM_STATS];
    uint8_t palette[TDE_NUM_CLASSES * 3];
} orc_env_set;

static void* dup_mem(const void* p, size_t n) {
    void* q = malloc(n ? n : 1);
    if (p && n) memcpy(q, p, n);
    return q;
}

static const uint8_t k_default_palette[TDE_NUM_CLASSES * 3] = {
    0, 0, 0,        /* background */
    128, 128, 128,  /* road */
    255, 255, 255,  /* lane marking */
    0, 200, 0,      /* tl green */
    230, 200, 0,    /* tl yellow */
    220, 0, 0,      /* tl red */
    0, 170, 255,    /* waypoint */
    60, 90, 220,    /* vehicle */
    250, 120, 0,    /* ego */
    200, 220, 255,  /* direction */
    255, 230, 150,  /* ego direction */
};

orc_env_set* orc_create(const tde_config* cfg, const tde_scenario_set* s) {
    orc_env_set* o = (orc_env_set*)calloc(1, sizeof(orc_env_set));
    o->cfg = *cfg;
    int A = o->A = cfg->max_agents, E = o->E = cfg->num_envs;
    int Nm = o->num_maps = s->num_maps, Ns = o->num_scen = s->num_scenarios;
    o->tri_off = dup_mem(s->map_tri_offset, sizeof(int32_t) * (Nm + 1));
    o->mark_off = dup_mem(s->map_mark_offset, sizeof(int32_t) * (Nm + 1));
    o->stop_off = dup_mem(s->map_stop_offset, sizeof(int32_t) * (Nm + 1));
    o->light_period = dup_mem(s->map_light_period, sizeof(int32_t) * Nm);
    o->light_off = dup_mem(s->map_light_offset, sizeof(int32_t) * (Nm + 1));
    o->road_tris = dup_mem(s->road_tris, sizeof(float) * 8 * o->tri_off[Nm]);
    o->mark_tris = dup_mem(s->mark_tris, sizeof(float) * 6 * o->mark_off[Nm]);
    o->stoplines = dup_mem(s->stoplines, sizeof(float) * 5 * o->stop_off[Nm]);
    o->light_states = dup_mem(s->light_states, o->light_off[Nm]);
    o->scen_map = dup_mem(s->scen_map, sizeof(int32_t) * Ns);
    o->wp_off = dup_mem(s->scen_wp_offset, sizeof(int32_t) * (Ns + 1));
    o->waypoints = dup_mem(s->waypoints, sizeof(float) * 2 * o->wp_off[Ns]);
    o->start_heading = dup_mem(s->scen_start_heading, sizeof(float) * Ns);
    o->scen_nag = dup_mem(s->scen_num_agents, sizeof(int32_t) * Ns);
    o->agent_init = dup_mem(s->agent_init, sizeof(float) * 4 * A * Ns);
    o->agent_attr = dup_mem(s->agent_attr, sizeof(float) * 3 * A * Ns);
    o->rep_T = dup_mem(s->scen_replay_T, sizeof(int32_t) * Ns);
    o->rep_off = dup_mem(s->scen_replay_offset, sizeof(int32_t) * (Ns + 1));
    o->replay_states = dup_mem(s->replay_states, sizeof(float) * 4 * A * o->rep_off[Ns]);
    o->replay_mask = dup_mem(s->replay_mask, (size_t)A * o->rep_off[Ns]);
    o->scen_lo = calloc(E, sizeof(int32_t));
    o->scen_hi = malloc(sizeof(int32_t) * (E ? E : 1));
    for (int e = 0; e < E; ++e) o->scen_hi[e] = Ns;
    o->state = calloc((size_t)E * A * 4, sizeof(float));
    o->attr = calloc((size_t)E * A * 4, sizeof(float));
    o->infr = calloc((size_t)E * A * 4, sizeof(float));
    o->vars = calloc((size_t)E * 8, sizeof(int32_t));
    o->ep_return = calloc(E, sizeof(float));
    memcpy(o->palette, k_default_palette, sizeof(k_default_palette));
    return o;
}

void orc_destroy(orc_env_set* o) {
    if (!o) return;
    free(o->tri_off); free(o->mark_off); free(o->stop_off); free(o->light_period); free(o->light_off);
    free(o->road_tris); free(o->mark_tris); free(o->stoplines); free(o->light_states);
    free(o->scen_map); free(o->wp_off); free(o->waypoints); free(o->start_heading); free(o->scen_nag);
    free(o->agent_init); free(o->agent_attr); free(o->rep_T); free(o->rep_off); free(o->replay_states);
    free(o->replay_mask); free(o->scen_lo); free(o->scen_hi);
    free(o->state); free(o->attr); free(o->infr); free(o->vars); free(o->ep_return);
    free(o);
}

void orc_set_env_scenario_range(orc_env_set* o, const int32_t* lo, const int32_t* hi) {
    memcpy(o->scen_lo, lo, sizeof(int32_t) * o->E);
    memcpy(o->scen_hi, hi, sizeof(int32_t) * o->E);
}
void orc_set_terminal_buffer(orc_env_set* o, uint8_t* buf) { o->terminal_obs = buf; }
void orc_set_palette(orc_env_set* o, const uint8_t* rgb) { memcpy(o->palette, rgb, TDE_NUM_CLASSES * 3); }
float* orc_state(orc_env_set* o) { return o->state; }
float* orc_attr(orc_env_set* o) { return o->attr; }
float* orc_infractions(orc_env_set* o) { return o->infr; }
int32_t* orc_vars(orc_env_set* o) { return o->vars; }
double* orc_stats(orc_env_set* o) { return o->stats; }

/* WaypointSuiteEnv.reset :319-349, set_start_pos :351-367, build_simulator state init :192-198,
   241-247, replay tensors :275-283.  Randomness is the counter-based generator above instead of
   numpy's global MT19937 (the reference ignores reset(seed=) :107-109,319); N(0, 0.1) heading noise
   :361 is the Irwin-Hall approximation. */
static void reset_env(orc_env_set* o, int e) {
    const tde_config* c = &o->cfg;
    int A = o->A;
    int32_t* v = o->vars + 8 * e;
    uint64_t genv = (uint64_t)(c->env_index_offset + e);
    uint64_t ep = (uint64_t)(uint32_t)v[5];
    int lo = o->scen_lo[e], hi = o->scen_hi[e];
    int span = hi - lo; if (span < 1) span = 1;
    int s = lo + (int)((orc_rng(o->seed, genv, ep, 0) >> 32) % (uint64_t)span);
    int m = o->scen_map[s];
    const float* wp = o->waypoints + 2 * o->wp_off[s];
    int W = o->wp_off[s + 1] - o->wp_off[s];
    float u_pos = u01(orc_rng(o->seed, genv, ep, 1));
    float u_spd = u01(orc_rng(o->seed, genv, ep, 2));
    float z = normal8(orc_rng(o->seed, genv, ep, 3), orc_rng(o->seed, genv, ep, 4));
    float p0x = wp[0], p0y = wp[1];
    float p1x = W > 1 ? wp[2] : p0x, p1y = W > 1 ? wp[3] : p0y;
    float* st = o->state + (size_t)e * A * 4;
    float* at = o->attr + (size_t)e * A * 4;
    int nag = o->scen_nag[s];
    int T = o->rep_T[s];
    for (int a = 0; a < A; ++a) {
        const float* init = o->agent_init + ((size_t)s * A + a) * 4;
        const float* attr = o->agent_attr + ((size_t)s * A + a) * 3;
        if (a >= 1 && T > 0 && o->replay_mask[(size_t)o->rep_off[s] * A + a]) init = o->replay_states + ((size_t)o->rep_off[s] * A + a) * 4;
        for (int k = 0; k < 4; ++k) st[4 * a + k] = init[k];
        at[4 * a + 0] = attr[0]; at[4 * a + 1] = attr[1]; at[4 * a + 2] = attr[2];
        at[4 * a + 3] = a < nag ? 1.0f : 0.0f;
    }
    st[0] = p0x + u_pos * (p1x - p0x);
    st[1] = p0y + u_pos * (p1y - p0y);
    st[2] = o->start_heading[s] + c->start_heading_sigma * z;
    st[3] = u_spd * c->start_speed_max;
    if (c->randomize_ego_attributes) {
        at[0] = 4.8f + u01(orc_rng(o->seed, genv, ep, 6)) * (5.5f - 4.8f);
        at[1] = 1.8f + u01(orc_rng(o->seed, genv, ep, 7)) * (2.2f - 1.8f);
        at[2] = 0.82f + u01(orc_rng(o->seed, genv, ep, 8)) * (0.97f - 0.82f);
    }
    int P = o->light_period[m];
    v[0] = s; v[1] = 0; v[2] = 1; v[3] = 0;
    v[4] = (int32_t)((orc_rng(o->seed, genv, ep, 5) >> 32) % (uint64_t)(P > 0 ? P : 1));
    v[5] = (int32_t)((uint32_t)v[5] + 1u);
    v[6] = m; v[7] = 0;
    o->ep_return[e] = 0.0f;
    memset(o->infr + (size_t)e * A * 4, 0, sizeof(float) * 4 * A);
}

void orc_reset(orc_env_set* o, const uint8_t* mask, uint64_t seed) {
    if (!mask) o->seed = seed;   /* the seed belongs to the env set: a masked reset keeps it (tde_reset) */
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o->E; ++e)
        if (!mask || mask[e]) reset_env(o, e);
}

static int light_state(const orc_env_set* o, int m, int step, int phase, int l) {
    int P = o->light_period[m];
    int L = o->stop_off[m + 1] - o->stop_off[m];
    if (P <= 0 || L <= 0) return TDE_LIGHT_GREEN;
    int t = (step + phase) % P;
    return o->light_states[o->light_off[m] + t * L + l];
}

/* simulator.step(action) gym_env.py:117 [EXT-RECALLED]: ego takes the action; every NPC slot is
   overwritten from replay_states[:, a, t] where replay_mask (:275-294), otherwise (IAI DRIVE being
   unreachable offline) advances at constant velocity = bicycle step with zero action. */
static void kinematics_env(orc_env_set* o, int e, const float* action) {
    int A = o->A;
    const tde_config* c = &o->cfg;
    int32_t* v = o->vars + 8 * e;
    int s = v[0];
    int t = v[1] + 1; /* time index after this step */
    float* st = o->state + (size_t)e * A * 4;
    const float* at = o->attr + (size_t)e * A * 4;
    int T = o->rep_T[s];
    for (int a = 0; a < A; ++a) {
        if (at[4 * a + 3] == 0.0f) continue; /* absent agents do not move */
        if (a == 0) {
            orc_bicycle_step(st, action[0], action[1], at[2], c->dt);
        } else if (t < T && o->replay_mask[((size_t)o->rep_off[s] + t) * A + a]) {
            const float* r = o->replay_states + (((size_t)o->rep_off[s] + t) * A + a) * 4;
            for (int k = 0; k < 4; ++k) st[4 * a + k] = r[k];
        } else {
            orc_bicycle_step(st + 4 * a, 0.0f, 0.0f, at[4 * a + 2], c->dt);
        }
    }
}

/* compute_collision / compute_offroad / compute_traffic_lights_violations / compute_wrong_way for
   every agent of env e at its current state and light time index `step`. */
static void infractions_env(orc_env_set* o, int e, int step) {
    int A = o->A;
    const tde_config* c = &o->cfg;
    const int32_t* v = o->vars + 8 * e;
    int m = v[6];
    const float* st = o->state + (size_t)e * A * 4;
    const float* at = o->attr + (size_t)e * A * 4;
    float* inf = o->infr + (size_t)e * A * 4;
    const float* tris = o->road_tris + 8 * (size_t)o->tri_off[m];
    int ntri = o->tri_off[m + 1] - o->tri_off[m];
    int L = o->stop_off[m + 1] - o->stop_off[m];
    orc_box box[TDE_MAX_AGENTS];
    for (int a = 0; a < A; ++a) box[a] = make_box(st[4 * a], st[4 * a + 1], st[4 * a + 2], at[4 * a], at[4 * a + 1]);
    for (int a = 0; a < A; ++a) {
        float* out = inf + 4 * a;
        out[0] = out[1] = out[2] = out[3] = 0.0f;
        if (at[4 * a + 3] == 0.0f) continue;
        float cnt = 0.0f;
        for (int b = 0; b < A; ++b)
            if (b != a && at[4 * b + 3] != 0.0f && orc_box_overlap(&box[a], &box[b])) cnt += 1.0f;
        out[TDE_INFR_COLLISION] = cnt;
        out[TDE_INFR_OFFROAD] = orc_offroad_box(&box[a], tris, ntri, c->offroad_threshold);
        /* TrafficLightControl.compute_violation [EXT-RECALLED]: the rear `tl_rear_factor` strip of the
           agent box against stop-line boxes whose light is red */
        float f = c->tl_rear_factor;
        float len2 = at[4 * a] * f;
        float back = 0.5f * (at[4 * a] - len2);
        orc_box rear = box[a];
        rear.x = box[a].x - back * box[a].c;
        rear.y = box[a].y - back * box[a].s;
        rear.hl = 0.5f * len2;
        float viol = 0.0f;
        for (int l = 0; l < L; ++l) {
            if (light_state(o, m, step, v[4], l) != TDE_LIGHT_RED) continue;
            const float* sl = o->stoplines + 5 * ((size_t)o->stop_off[m] + l);
            orc_box sb = make_box(sl[0], sl[1], sl[4], sl[2], sl[3]);
            if (orc_box_overlap(&rear, &sb)) viol += 1.0f;
        }
        out[TDE_INFR_TL_VIOLATION] = viol;
        out[TDE_INFR_WRONG_WAY] = orc_wrong_way_box(&box[a], tris, ntri);
    }
}

/* ------------------------------------------------------------------ birdview */

typedef struct { int32_t x[4], y[4]; int n; int cls; } orc_prim;

typedef struct { float ex, ey, ce, se, ppm, ppmy; } orc_cam;

static int32_t snap16(float f) {
    float r = rintf(f * 16.0f);
    if (r > 8191.0f) r = 8191.0f;
    if (r < -8191.0f) r = -8191.0f;
    return (int32_t)r;
}

/* world -> pixel: translate to the ego, rotate by -psi (ego faces +x = right of the image),
   scale by W/fov; left-handed (CARLA) coordinates keep +y down the image, right-handed flip it. */
static void to_pixel(const orc_cam* cam, float wx, float wy, float* fx, float* fy) {
    float dx = wx - cam->ex, dy = wy - cam->ey;
    float cx = dx * cam->ce + dy * cam->se;
    float cy = dy * cam->ce - dx * cam->se;
    *fx = cx * cam->ppm + 0.5f * (float)TDE_OBS_W;
    *fy = cy * cam->ppmy + 0.5f * (float)TDE_OBS_H;
}

/* returns 0 when the primitive's float bbox misses the viewport grown by one pixel */
static int make_prim(const orc_cam* cam, const float* wx, const float* wy, int n, int cls, orc_prim* p) {
    float minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
    float fx[4], fy[4];
    for (int k = 0; k < n; ++k) {
        to_pixel(cam, wx[k], wy[k], &fx[k], &fy[k]);
        minx = fminf(minx, fx[k]); maxx = fmaxf(maxx, fx[k]);
        miny = fminf(miny, fy[k]); maxy = fmaxf(maxy, fy[k]);
    }
    if (!(maxx >= -1.0f && minx <= (float)TDE_OBS_W + 1.0f && maxy >= -1.0f && miny <= (float)TDE_OBS_H + 1.0f)) return 0;
    for (int k = 0; k < n; ++k) { p->x[k] = snap16(fx[k]); p->y[k] = snap16(fy[k]); }
    p->n = n; p->cls = cls;
    return 1;
}

/* pixel-centre sampling, integer edge functions on the 1/16-pixel grid, top-left tie rule */
static void paint_prim(const orc_prim* p, uint8_t* cls_img) {
    int n = p->n;
    int64_t area2 = 0;
    for (int k = 0; k < n; ++k) {
        int k1 = (k + 1) % n;
        area2 += (int64_t)p->x[k] * p->y[k1] - (int64_t)p->x[k1] * p->y[k];
    }
    if (area2 == 0) return;
    int32_t X[4], Y[4];
    for (int k = 0; k < n; ++k) {
        int src = area2 > 0 ? k : (n - 1 - k);
        X[k] = p->x[src]; Y[k] = p->y[src];
    }
    int32_t xmin = X[0], xmax = X[0], ymin = Y[0], ymax = Y[0];
    for (int k = 1; k < n; ++k) {
        if (X[k] < xmin) xmin = X[k];
        if (X[k] > xmax) xmax = X[k];
        if (Y[k] < ymin) ymin = Y[k];
        if (Y[k] > ymax) ymax = Y[k];
    }
    /* only pixel centres inside the snapped bounding box can be covered */
    int i0 = (xmin - 8) >= 0 ? (xmin - 8) / 16 : 0, i1 = (xmax - 8) >= 0 ? (xmax - 8) / 16 : -1;
    int j0 = (ymin - 8) >= 0 ? (ymin - 8) / 16 : 0, j1 = (ymax - 8) >= 0 ? (ymax - 8) / 16 : -1;
    if (i1 > TDE_OBS_W - 1) i1 = TDE_OBS_W - 1;
    if (j1 > TDE_OBS_H - 1) j1 = TDE_OBS_H - 1;
    for (int j = j0; j <= j1; ++j) {
        for (int i = i0; i <= i1; ++i) {
            int64_t px = 16 * i + 8, py = 16 * j + 8;
            int in = 1;
            for (int k = 0; k < n && in; ++k) {
                int k1 = (k + 1) % n;
                int64_t dx = X[k1] - X[k], dy = Y[k1] - Y[k];
                if (dx == 0 && dy == 0) continue;
                int64_t E = dx * (py - Y[k]) - dy * (px - X[k]);
                int incl = (dy < 0) || (dy == 0 && dx > 0);
                if (E < 0 || (E == 0 && !incl)) in = 0;
            }
            if (in && p->cls > cls_img[j * TDE_OBS_W + i]) cls_img[j * TDE_OBS_W + i] = (uint8_t)p->cls;
        }
    }
}

static void box_world_quad(const orc_box* b, float* wx, float* wy) {
    for (int k = 0; k < 4; ++k) box_corner(b, k, &wx[k], &wy[k]);
}
static void box_world_dirtri(const orc_box* b, float* wx, float* wy) {
    float ox[3] = {b->hl, 0.5f * b->hl, 0.5f * b->hl};
    float oy[3] = {0.0f, b->hw, -b->hw};
    for (int k = 0; k < 3; ++k) {
        wx[k] = b->x + (ox[k] * b->c - oy[k] * b->s);
        wy[k] = b->y + (ox[k] * b->s + oy[k] * b->c);
    }
}

/* simulator.render_egocentric() gym_env.py:122-124 [EXT-RECALLED]: 3x64x64 uint8 RGB centred on the
   ego; painter's levels road < lane markings < stop lines by light state < goal waypoint <
   vehicles < direction triangles < ego (highlighted) */
void orc_render_env(orc_env_set* o, int e, uint8_t* obs /* 3*64*64 */, uint8_t* cls_out /* 64*64 or NULL */) {
    int A = o->A;
    const tde_config* c = &o->cfg;
    const int32_t* v = o->vars + 8 * e;
    int s = v[0], m = v[6];
    const float* st = o->state + (size_t)e * A * 4;
    const float* at = o->attr + (size_t)e * A * 4;
    uint8_t cls_img[TDE_OBS_H * TDE_OBS_W];
    memset(cls_img, 0, sizeof(cls_img));
    orc_cam cam;
    cam.ex = st[0]; cam.ey = st[1];
    orc_sincosf(st[2], &cam.se, &cam.ce);
    cam.ppm = (float)TDE_OBS_W / c->fov;
    cam.ppmy = c->left_handed_coordinates ? cam.ppm : -cam.ppm;
    orc_prim p;
    float wx[4], wy[4];
    for (int t = o->tri_off[m]; t < o->tri_off[m + 1]; ++t) {
        const float* tr = o->road_tris + 8 * (size_t)t;
        wx[0] = tr[0]; wy[0] = tr[1]; wx[1] = tr[2]; wy[1] = tr[3]; wx[2] = tr[4]; wy[2] = tr[5];
        if (make_prim(&cam, wx, wy, 3, TDE_CLS_ROAD, &p)) paint_prim(&p, cls_img);
    }
    for (int t = o->mark_off[m]; t < o->mark_off[m + 1]; ++t) {
        const float* tr = o->mark_tris + 6 * (size_t)t;
        wx[0] = tr[0]; wy[0] = tr[1]; wx[1] = tr[2]; wy[1] = tr[3]; wx[2] = tr[4]; wy[2] = tr[5];
        if (make_prim(&cam, wx, wy, 3, TDE_CLS_LANE_MARKING, &p)) paint_prim(&p, cls_img);
    }
    int L = o->stop_off[m + 1] - o->stop_off[m];
    for (int l = 0; l < L; ++l) {
        const float* sl = o->stoplines + 5 * ((size_t)o->stop_off[m] + l);
        orc_box sb = make_box(sl[0], sl[1], sl[4], sl[2], sl[3]);
        box_world_quad(&sb, wx, wy);
        int ls = light_state(o, m, v[1], v[4], l);
        int cls = ls == TDE_LIGHT_RED ? TDE_CLS_TL_RED : (ls == TDE_LIGHT_YELLOW ? TDE_CLS_TL_YELLOW : TDE_CLS_TL_GREEN);
        if (make_prim(&cam, wx, wy, 4, cls, &p)) paint_prim(&p, cls_img);
    }
    int W = o->wp_off[s + 1] - o->wp_off[s];
    if (v[2] < W) { /* current target waypoint: a diamond of circumradius 2 m */
        const float* wp = o->waypoints + 2 * ((size_t)o->wp_off[s] + v[2]);
        float r = 2.0f;
        wx[0] = wp[0] + r; wy[0] = wp[1]; wx[1] = wp[0]; wy[1] = wp[1] + r;
        wx[2] = wp[0] - r; wy[2] = wp[1]; wx[3] = wp[0]; wy[3] = wp[1] - r;
        if (make_prim(&cam, wx, wy, 4, TDE_CLS_WAYPOINT, &p)) paint_prim(&p, cls_img);
    }
    for (int a = 0; a < A; ++a) {
        if (at[4 * a + 3] == 0.0f) continue;
        orc_box b = make_box(st[4 * a], st[4 * a + 1], st[4 * a + 2], at[4 * a], at[4 * a + 1]);
        box_world_quad(&b, wx, wy);
        if (make_prim(&cam, wx, wy, 4, a == 0 ? TDE_CLS_EGO : TDE_CLS_VEHICLE, &p)) paint_prim(&p, cls_img);
        box_world_dirtri(&b, wx, wy);
        if (make_prim(&cam, wx, wy, 3, a == 0 ? TDE_CLS_EGO_DIRECTION : TDE_CLS_DIRECTION, &p)) paint_prim(&p, cls_img);
    }
    if (obs)
        for (int ch = 0; ch < 3; ++ch)
            for (int k = 0; k < TDE_OBS_H * TDE_OBS_W; ++k) obs[ch * TDE_OBS_H * TDE_OBS_W + k] = o->palette[3 * cls_img[k] + ch];
    if (cls_out) memcpy(cls_out, cls_img, sizeof(cls_img));
}

/* ------------------------------------------------------------------ recording view

   BirdviewRecordingWrapper(simulator, res=Resolution(video_res, video_res), fov=video_fov) gym_env.py:52-53,
   :295-297 [EXT-RECALLED]: a W x H frame of ONE env from a free camera (centre, heading, field of view in
   metres across the width).  Same primitives, classes and fill rule as the observation (raw triangles, no
   quad merging), but the snapped coordinates are kept in a wider range (+-2^20 sixteenths of a pixel, 64-bit
   edge functions), so any resolution up to 4096 and any zoom can be drawn. */
#define VIEW_SNAP_MAX 1048575.0f
typedef struct { float ex, ey, ce, se, ppm, ppmy; int W, H; } view_cam;

static int view_make_prim(const view_cam* cam, const float* wx, const float* wy, int n, int cls, orc_prim* p) {
    float minx = INFINITY, maxx = -INFINITY, miny = INFINITY, maxy = -INFINITY;
    float fx[4], fy[4];
    for (int k = 0; k < n; ++k) {
        float dx = wx[k] - cam->ex, dy = wy[k] - cam->ey;
        float cx = dx * cam->ce + dy * cam->se;
        float cy = dy * cam->ce - dx * cam->se;
        fx[k] = cx * cam->ppm + 0.5f * (float)cam->W;
        fy[k] = cy * cam->ppmy + 0.5f * (float)cam->H;
        minx = fminf(minx, fx[k]); maxx = fmaxf(maxx, fx[k]);
        miny = fminf(miny, fy[k]); maxy = fmaxf(maxy, fy[k]);
    }
    if (!(maxx >= -1.0f && minx <= (float)cam->W + 1.0f && maxy >= -1.0f && miny <= (float)cam->H + 1.0f)) return 0;
    for (int k = 0; k < n; ++k) {
        float rx = rintf(fx[k] * 16.0f), ry = rintf(fy[k] * 16.0f);
        rx = fminf(fmaxf(rx, -VIEW_SNAP_MAX), VIEW_SNAP_MAX);
        ry = fminf(fmaxf(ry, -VIEW_SNAP_MAX), VIEW_SNAP_MAX);
        p->x[k] = (int32_t)rx; p->y[k] = (int32_t)ry;
    }
    p->n = n; p->cls = cls;
    return 1;
}

static void view_paint_prim(const orc_prim* p, uint8_t* cls_img, int W, int H) {
    int n = p->n;
    int64_t area2 = 0;
    for (int k = 0; k < n; ++k) {
        int k1 = (k + 1) % n;
        area2 += (int64_t)p->x[k] * p->y[k1] - (int64_t)p->x[k1] * p->y[k];
    }
    if (area2 == 0) return;
    int64_t X[4], Y[4];
    for (int k = 0; k < n; ++k) {
        int src = area2 > 0 ? k : (n - 1 - k);
        X[k] = p->x[src]; Y[k] = p->y[src];
    }
    int64_t xmin = X[0], xmax = X[0], ymin = Y[0], ymax = Y[0];
    for (int k = 1; k < n; ++k) {
        if (X[k] < xmin) xmin = X[k];
        if (X[k] > xmax) xmax = X[k];
        if (Y[k] < ymin) ymin = Y[k];
        if (Y[k] > ymax) ymax = Y[k];
    }
    int64_t i0 = (xmin - 8) >= 0 ? (xmin - 8 + 15) / 16 : 0, i1 = (xmax - 8) >= 0 ? (xmax - 8) / 16 : -1;
    int64_t j0 = (ymin - 8) >= 0 ? (ymin - 8 + 15) / 16 : 0, j1 = (ymax - 8) >= 0 ? (ymax - 8) / 16 : -1;
    if (i1 > W - 1) i1 = W - 1;
    if (j1 > H - 1) j1 = H - 1;
    for (int64_t j = j0; j <= j1; ++j)
        for (int64_t i = i0; i <= i1; ++i) {
            int64_t px = 16 * i + 8, py = 16 * j + 8;
            int in = 1;
            for (int k = 0; k < n && in; ++k) {
                int k1 = (k + 1) % n;
                int64_t dx = X[k1] - X[k], dy = Y[k1] - Y[k];
                if (dx == 0 && dy == 0) continue;
                int64_t E = dx * (py - Y[k]) - dy * (px - X[k]);
                int incl = (dy < 0) || (dy == 0 && dx > 0);
                if (E < 0 || (E == 0 && !incl)) in = 0;
            }
            if (in && p->cls > cls_img[j * W + i]) cls_img[j * W + i] = (uint8_t)p->cls;
        }
}

/* out: 3*H*W uint8 (planar RGB).  Returns 0, or -1 for a bad argument. */
int orc_render_view(orc_env_set* o, int e, float cam_x, float cam_y, float cam_psi, float fov, int W, int H, uint8_t* out) {
    if (e < 0 || e >= o->E || W < 1 || H < 1 || W > 4096 || H > 4096 || !(fov > 0.0f) || !out) return -1;
    int A = o->A;
    const int32_t* v = o->vars + 8 * e;
    int s = v[0], m = v[6];
    const float* st = o->state + (size_t)e * A * 4;
    const float* at = o->attr + (size_t)e * A * 4;
    uint8_t* cls_img = (uint8_t*)calloc((size_t)W * H, 1);
    view_cam cam;
    cam.ex = cam_x; cam.ey = cam_y; cam.W = W; cam.H = H;
    orc_sincosf(cam_psi, &cam.se, &cam.ce);
    cam.ppm = (float)W / fov;
    cam.ppmy = o->cfg.left_handed_coordinates ? cam.ppm : -cam.ppm;
    orc_prim p;
    float wx[4], wy[4];
    for (int t = o->tri_off[m]; t < o->tri_off[m + 1]; ++t) {
        const float* tr = o->road_tris + 8 * (size_t)t;
        wx[0] = tr[0]; wy[0] = tr[1]; wx[1] = tr[2]; wy[1] = tr[3]; wx[2] = tr[4]; wy[2] = tr[5];
        if (view_make_prim(&cam, wx, wy, 3, TDE_CLS_ROAD, &p)) view_paint_prim(&p, cls_img, W, H);
    }
    for (int t = o->mark_off[m]; t < o->mark_off[m + 1]; ++t) {
        const float* tr = o->mark_tris + 6 * (size_t)t;
        wx[0] = tr[0]; wy[0] = tr[1]; wx[1] = tr[2]; wy[1] = tr[3]; wx[2] = tr[4]; wy[2] = tr[5];
        if (view_make_prim(&cam, wx, wy, 3, TDE_CLS_LANE_MARKING, &p)) view_paint_prim(&p, cls_img, W, H);
    }
    int L = o->stop_off[m + 1] - o->stop_off[m];
    for (int l = 0; l < L; ++l) {
        const float* sl = o->stoplines + 5 * ((size_t)o->stop_off[m] + l);
        orc_box sb = make_box(sl[0], sl[1], sl[4], sl[2], sl[3]);
        box_world_quad(&sb, wx, wy);
        int ls = light_state(o, m, v[1], v[4], l);
        int cls = ls == TDE_LIGHT_RED ? TDE_CLS_TL_RED : (ls == TDE_LIGHT_YELLOW ? TDE_CLS_TL_YELLOW : TDE_CLS_TL_GREEN);
        if (view_make_prim(&cam, wx, wy, 4, cls, &p)) view_paint_prim(&p, cls_img, W, H);
    }
    int NW = o->wp_off[s + 1] - o->wp_off[s];
    if (v[2] < NW) {
        const float* wp = o->waypoints + 2 * ((size_t)o->wp_off[s] + v[2]);
        float r = 2.0f;
        wx[0] = wp[0] + r; wy[0] = wp[1]; wx[1] = wp[0]; wy[1] = wp[1] + r;
        wx[2] = wp[0] - r; wy[2] = wp[1]; wx[3] = wp[0]; wy[3] = wp[1] - r;
        if (view_make_prim(&cam, wx, wy, 4, TDE_CLS_WAYPOINT, &p)) view_paint_prim(&p, cls_img, W, H);
    }
    for (int a = 0; a < A; ++a) {
        if (at[4 * a + 3] == 0.0f) continue;
        orc_box b = make_box(st[4 * a], st[4 * a + 1], st[4 * a + 2], at[4 * a], at[4 * a + 1]);
        box_world_quad(&b, wx, wy);
        if (view_make_prim(&cam, wx, wy, 4, a == 0 ? TDE_CLS_EGO : TDE_CLS_VEHICLE, &p)) view_paint_prim(&p, cls_img, W, H);
        box_world_dirtri(&b, wx, wy);
        if (view_make_prim(&cam, wx, wy, 3, a == 0 ? TDE_CLS_EGO_DIRECTION : TDE_CLS_DIRECTION, &p)) view_paint_prim(&p, cls_img, W, H);
    }
    for (int ch = 0; ch < 3; ++ch)
        for (size_t k = 0; k < (size_t)W * H; ++k) out[(size_t)ch * W * H + k] = o->palette[3 * cls_img[k] + ch];
    free(cls_img);
    return 0;
}

void orc_render(orc_env_set* o, uint8_t* obs) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int e = 0; e < o->E; ++e) orc_render_env(o, e, obs + (size_t)e * 3 * TDE_OBS_H * TDE_OBS_W, NULL);
}
void orc_render_classes(orc_env_set* o, uint8_t* cls) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int e = 0; e < o->E; ++e) orc_render_env(o, e, NULL, cls + (size_t)e * TDE_OBS_H * TDE_OBS_W);
}

void orc_compute_infractions(orc_env_set* o) {
#pragma omp parallel for schedule(dynamic, 8)
    for (int e = 0; e < o->E; ++e) infractions_env(o, e, o->vars[8 * e + 1]);
}

void orc_kinematics(orc_env_set* o, const float* actions) {
#pragma omp parallel for schedule(static)
    for (int e = 0; e < o->E; ++e) {
        kinematics_env(o, e, actions + 2 * e);
        o->vars[8 * e + 1] += 1; /* the simulator's time index advances with every simulator.step */
    }
}

/* WaypointSuiteEnv.step :369-389 with GymEnv.step :115-120 inlined, for all E envs.
   `phases` selects sub-steps exactly as tde_step_phases does. */
void orc_step_phases(orc_env_set* o, int phases, const float* actions, uint8_t* obs, float* reward,
                     uint8_t* terminated, uint8_t* truncated, float* info, double* stats_per_env /* unused */) {
    (void)stats_per_env;
    int A = o->A, E = o->E;
    const tde_config* c = &o->cfg;
    double st_acc[TDE_NUM_STATS];
    memset(st_acc, 0, sizeof(st_acc));
#pragma omp parallel
    {
        double loc[TDE_NUM_STATS];
        memset(loc, 0, sizeof(loc));
#pragma omp for schedule(dynamic, 8)
        for (int e = 0; e < E; ++e) {
            int32_t* v = o->vars + 8 * e;
            float* st = o->state + (size_t)e * A * 4;
            float lx = st[0], ly = st[1], lpsi = st[2], lv = st[3]; /* :371-375 snapshot */
            if (phases & TDE_PH_KINEMATICS) {
                kinematics_env(o, e, actions + 2 * e); /* :117 */
                v[1] += 1;                             /* :116 */
            }
            if (phases & TDE_PH_INFRACTIONS) infractions_env(o, e, v[1]);
            if (phases & TDE_PH_REWARD) {
                int s = v[0];
                const float* inf = o->infr + (size_t)e * A * 4;
                float x = st[0], y = st[1], psi = st[2], spd = st[3];
                float dx = x - lx, dy = y - ly;
                float d = sqrtf(dx * dx + dy * dy);                              /* :401 */
                float dist_reward = d > c->distance_cutoff ? c->distance_bonus : 0.0f; /* :402 */
                float sd, cd;
                orc_sincosf(psi - lpsi, &sd, &cd);
                float psi_reward = (1.0f - cd) * (-c->heading_penalty);          /* :403 */
                int W = o->wp_off[s + 1] - o->wp_off[s];
                int reached = 0;
                if (v[2] < W) { /* check_reach_target :391-394 */
                    const float* wp = o->waypoints + 2 * ((size_t)o->wp_off[s] + v[2]);
                    float tx = x - wp[0], ty = y - wp[1];
                    reached = sqrtf(tx * tx + ty * ty) < c->reach_radius;
                }
                float reach_reward = reached ? c->waypoint_bonus : 0.0f;          /* :404-408 */
                if (reached) v[3] += 1;                                           /* :406 */
                float r = (reach_reward + dist_reward) + psi_reward;              /* :410 */
                int term = c->terminated_at_infraction &&
                           (inf[TDE_INFR_OFFROAD] > 0.0f || inf[TDE_INFR_COLLISION] > 0.0f || inf[TDE_INFR_TL_VIOLATION] > 0.0f); /* :413-417 */
                int trunc = v[1] >= c->max_environment_steps;                     /* :134-135 */
                o->ep_return[e] = o->ep_return[e] + r;
                float* io = info + (size_t)e * TDE_INFO_STRIDE;
                for (int k = 0; k < TDE_INFO_STRIDE; ++k) io[k] = 0.0f;
                io[TDE_INFO_OFFROAD] = inf[TDE_INFR_OFFROAD];                     /* :427 */
                io[TDE_INFO_COLLISION] = inf[TDE_INFR_COLLISION];                 /* :428 */
                io[TDE_INFO_TL_VIOLATION] = inf[TDE_INFR_TL_VIOLATION];           /* :429 */
                io[TDE_INFO_IS_SUCCESS] = trunc ? 1.0f : 0.0f;                    /* :430 */
                io[TDE_INFO_REACHED_WAYPOINT_NUM] = (float)v[3];                  /* :431 */
                io[TDE_INFO_PSI_SMOOTHNESS] = fabsf((lpsi - psi) / c->dt);        /* :432 */
                io[TDE_INFO_PSI_REWARD] = psi_reward;                             /* :433 */
                io[TDE_INFO_DIST_REWARD] = dist_reward;                           /* :434 */
                io[TDE_INFO_SPEED_SMOOTHNESS] = fabsf((lv - spd) / c->dt);        /* :435 */
                io[TDE_INFO_WRONG_WAY] = inf[TDE_INFR_WRONG_WAY];
                io[TDE_INFO_EPISODE_RETURN] = o->ep_return[e];
                io[TDE_INFO_EPISODE_LENGTH] = (float)v[1];
                io[TDE_INFO_SCENARIO] = (float)s;
                reward[e] = r;
                terminated[e] = (uint8_t)term;
                truncated[e] = (uint8_t)trunc;
                if (reached) v[2] += 1; /* :378-383 */
                loc[TDE_STAT_STEPS] += 1.0;
                if (term || trunc) {
                    loc[TDE_STAT_EPISODES] += 1.0;
                    loc[TDE_STAT_RETURN_SUM] += (double)o->ep_return[e];
                    loc[TDE_STAT_LENGTH_SUM] += (double)v[1];
                    loc[TDE_STAT_OFFROAD] += inf[TDE_INFR_OFFROAD] > 0.0f;
                    loc[TDE_STAT_COLLISION] += inf[TDE_INFR_COLLISION] > 0.0f;
                    loc[TDE_STAT_TL_VIOLATION] += inf[TDE_INFR_TL_VIOLATION] > 0.0f;
                    loc[TDE_STAT_SUCCESS] += trunc ? 1.0 : 0.0;
                    loc[TDE_STAT_REACHED_WAYPOINTS] += (double)v[3];
                    if (c->auto_reset) {
                        if (o->terminal_obs && (phases & TDE_PH_RENDER) && obs)
                            orc_render_env(o, e, o->terminal_obs + (size_t)e * 3 * TDE_OBS_H * TDE_OBS_W, NULL);
                        reset_env(o, e);
                        io[TDE_INFO_DID_RESET] = 1.0f;
                    }
                }
            }
            if ((phases & TDE_PH_RENDER) && obs) orc_render_env(o, e, obs + (size_t)e * 3 * TDE_OBS_H * TDE_OBS_W, NULL);
        }
#pragma omp critical
        for (int k = 0; k < TDE_NUM_STATS; ++k) st_acc[k] += loc[k];
    }
    for (int k = 0; k < TDE_NUM_STATS; ++k) o->stats[k] += st_acc[k];
}

void orc_step(orc_env_set* o, const float* actions, uint8_t* obs, float* reward, uint8_t* terminated,
              uint8_t* truncated, float* info) {
    orc_step_phases(o, TDE_PH_ALL, actions, obs, reward, terminated, truncated, info, NULL);
}

/* ------------------------------------------------------------------ stateless helpers for tests */

void orc_collision_boxes(const float* state, const float* attr, int E, int A, float* out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int e = 0; e < E; ++e) {
        orc_box box[TDE_MAX_AGENTS];
        const float* st = state + (size_t)e * A * 4;
        const float* at = attr + (size_t)e * A * 4;
        for (int a = 0; a < A; ++a) box[a] = make_box(st[4 * a], st[4 * a + 1], st[4 * a + 2], at[4 * a], at[4 * a + 1]);
        for (int a = 0; a < A; ++a) {
            float cnt = 0.0f;
            if (at[4 * a + 3] != 0.0f)
                for (int b = 0; b < A; ++b)
                    if (b != a && at[4 * b + 3] != 0.0f && orc_box_overlap(&box[a], &box[b])) cnt += 1.0f;
            out[(size_t)e * A + a] = cnt;
        }
    }
}

/* smallest |margin| over the pairs of each agent: lets the tests report the epsilon band */
void orc_collision_margins(const float* state, const float* attr, int E, int A, float* out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int e = 0; e < E; ++e) {
        orc_box box[TDE_MAX_AGENTS];
        const float* st = state + (size_t)e * A * 4;
        const float* at = attr + (size_t)e * A * 4;
        for (int a = 0; a < A; ++a) box[a] = make_box(st[4 * a], st[4 * a + 1], st[4 * a + 2], at[4 * a], at[4 * a + 1]);
        for (int a = 0; a < A; ++a) {
            float best = INFINITY;
            if (at[4 * a + 3] != 0.0f)
                for (int b = 0; b < A; ++b)
                    if (b != a && at[4 * b + 3] != 0.0f) { float m = fabsf(orc_box_margin(&box[a], &box[b])); if (m < best) best = m; }
            out[(size_t)e * A + a] = best;
        }
    }
}

void orc_offroad_boxes(const float* tris, int ntri, float threshold, const float* state, const float* attr,
                       int E, int A, float* out) {
#pragma omp parallel for schedule(dynamic, 16)
    for (int e = 0; e < E; ++e) {
        const float* st = state + (size_t)e * A * 4;
        const float* at = attr + (size_t)e * A * 4;
        for (int a = 0; a < A; ++a) {
            float val = 0.0f;
            if (at[4 * a + 3] != 0.0f) {
                orc_box b = make_box(st[4 * a], st[4 * a + 1], st[4 * a + 2], at[4 * a], at[4 * a + 1]);
                val = orc_offroad_box(&b, tris, ntri, threshold);
            }
            out[(size_t)e * A + a] = val;
        }
    }
}

void orc_sincos_array(const float* x, int n, float* s, float* c) {
    for (int i = 0; i < n; ++i) orc_sincosf(x[i], &s[i], &c[i]);
}
void orc_bicycle_array(float* state, const float* action, const float* lr, float dt, int n) {
    for (int i = 0; i < n; ++i) orc_bicycle_step(state + 4 * i, action[2 * i], action[2 * i + 1], lr[i], dt);
}
int orc_overlap_raw(const float* a5, const float* b5) { /* x y psi length width */
    orc_box A = make_box(a5[0], a5[1], a5[2], a5[3], a5[4]);
    orc_box B = make_box(b5[0], b5[1], b5[2], b5[3], b5[4]);
    return orc_box_overlap(&A, &B);
}
float orc_point_mesh_distance(const float* tris, int ntri, float px, float py) {
    float best = INFINITY;
    for (int t = 0; t < ntri; ++t) { int in; float d2 = point_tri_dist2(tris + 8 * t, px, py, &in); if (d2 < best) best = d2; }
    return sqrtf(best);
}
int orc_num_threads(void) {
    int n = 1;
#ifdef _OPENMP
    extern int omp_get_max_threads(void);
    n = omp_get_max_threads();
#endif
    return n;
}
/* torch.distributed.run exports OMP_NUM_THREADS=1 to its workers: the timing legs set the thread count themselves */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
    extern void omp_set_num_threads(int);
    if (n >= 1) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
