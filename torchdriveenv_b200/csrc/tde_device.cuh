// tde_device.cuh — device-side arithmetic of the simulation hot path (sm_100a).
//
// Arithmetic contract: IEEE binary32, one rounding per written operation (the translation unit is
// compiled with -fmad=false), FMA only where __fmaf_rn is written.  This is what lets the CPU oracle
// be compared bit for bit; see DESIGN.md §SPEC.
#pragma once
#ifdef TDE_HOST_EMU
#include "cuda_emu.h"   // tests/emu: the same source compiled for the host-side lockstep emulator (test infrastructure only)
#else
#include <cuda_runtime.h>
#define TDE_LAUNCH(g, b, s, st, ...) __VA_ARGS__<<<g, b, s, st>>>
#endif
#include <math.h>
#include <stdint.h>

// ---------------------------------------------------------------- bulk-async staging into shared memory (TMA engine)
//
// cp.async.bulk global -> shared::cta, completion signalled on an mbarrier (SASS: UBLKCP + SYNCS).  One thread arms the
// barrier with the byte count and issues the copies; every thread that reads the staged tables waits on the barrier's
// phase first.  Sizes and addresses are multiples of 16 bytes.  Shared-memory data is then read through explicit
// ld.shared wrappers taking 32-bit shared-window addresses (the helpers that use them are not inlined, so a C++ pointer
// would degrade to generic loads).  The emulator build copies immediately and reads through the block's buffer.
#ifdef TDE_HOST_EMU
__device__ __forceinline__ uint32_t tde_smem_addr(const void* p) { return (uint32_t)((const unsigned char*)p - emu::dyn_smem()); }
// emulated mbarrier word: low half = bytes still expected, high half = completed phases (one arrival per phase)
__device__ __forceinline__ void tde_mbar_init(unsigned long long* bar, int) { *bar = 0ull; }
__device__ __forceinline__ void tde_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) { *bar += bytes; }
__device__ __forceinline__ void tde_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    memcpy(dst, src, bytes);
    *bar -= bytes;
    if ((uint32_t)*bar == 0u) *bar += 1ull << 32;   // the last byte landed: the phase completes
    emu::rt().progress++;
}
__device__ __forceinline__ void tde_mbar_wait(unsigned long long* bar, uint32_t parity) {
    emu::wait_phase(bar, parity);
}
__device__ __forceinline__ float4 tde_lds_f4(uint32_t a) { return *reinterpret_cast<const float4*>(emu::dyn_smem() + a); }
__device__ __forceinline__ uint32_t tde_lds_u16(uint32_t a) { return *reinterpret_cast<const uint16_t*>(emu::dyn_smem() + a); }
__device__ __forceinline__ uint32_t tde_lds_u8(uint32_t a) { return *reinterpret_cast<const uint8_t*>(emu::dyn_smem() + a); }
#else
__device__ __forceinline__ uint32_t tde_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tde_mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tde_smem_addr(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");   // make the init visible to the async proxy
}
__device__ __forceinline__ void tde_mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(tde_smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tde_bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tde_smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(tde_smem_addr(bar))
                 : "memory");
}
__device__ __forceinline__ void tde_mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TDE_MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TDE_MBAR_DONE;\n"
        "bra TDE_MBAR_WAIT;\n"
        "TDE_MBAR_DONE:\n"
        "}\n" ::"r"(tde_smem_addr(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ float4 tde_lds_f4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t tde_lds_u16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t tde_lds_u8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
#endif

#define TDE_PI_F 3.14159274101257324219f
#define TDE_TWO_PI_F 6.28318548202514648438f
#define FULL_MASK 0xffffffffu

// Deterministic sin/cos (Cody-Waite by pi/2 + cephes minimax polynomials, all explicit FMAs).
// Replaces torch.sin/torch.cos of KinematicBicycle.step and math.cos of get_reward (gym_env.py:403).
__device__ __forceinline__ void tde_sincosf(float x, float& s, float& c) {
    float kf = rintf(x * 0.636619772367581343f);
    if (!(fabsf(kf) < 1.0e9f)) kf = 0.0f;
    int k = (int)kf;
    float r = __fmaf_rn(kf, -1.57079625129699707031e+00f, x);
    r = __fmaf_rn(kf, -7.54978941586159635335e-08f, r);
    r = __fmaf_rn(kf, -5.39030285815811905290e-15f, r);
    float z = r * r;
    float sp = __fmaf_rn(-1.9515295891e-4f, z, 8.3321608736e-3f);
    sp = __fmaf_rn(sp, z, -1.6666654611e-1f);
    float sr = __fmaf_rn(sp * z, r, r);
    float cp = __fmaf_rn(2.443315711809948e-5f, z, -1.388731625493765e-3f);
    cp = __fmaf_rn(cp, z, 4.166664568298827e-2f);
    float cr = __fmaf_rn(cp * z, z, __fmaf_rn(-0.5f, z, 1.0f));
    float ss = (k & 1) ? cr : sr;
    float cc = (k & 1) ? sr : cr;
    // quadrant signs: k&3 = 0:(s,c) 1:(c,-s) 2:(-s,-c) 3:(-c,s)
    s = (k & 2) ? -ss : ss;
    c = (((k + 1) & 2) ? -cc : cc);
}

// psi <- ((pi + psi) mod 2pi) - pi, floored modulo (torch.remainder semantics)
__device__ __forceinline__ float tde_wrap_pi(float psi) {
    float t = psi + TDE_PI_F;
    float m = fmodf(t, TDE_TWO_PI_F);
    if (m < 0.0f) m += TDE_TWO_PI_F;
    return m - TDE_PI_F;
}

// KinematicBicycle.step (simulator.step, gym_env.py:117): v, then x/y with the new v, then psi, wrap.
__device__ __forceinline__ float4 tde_bicycle(float4 st, float a, float beta, float lr, float dt) {
    float s1, c1, sb, cb;
    float v = st.w + a * dt;
    tde_sincosf(st.z + beta, s1, c1);
    tde_sincosf(beta, sb, cb);
    float x = st.x + (v * c1) * dt;
    float y = st.y + (v * s1) * dt;
    float psi = st.z + ((v / lr) * sb) * dt;
    psi = tde_wrap_pi(psi);
    return make_float4(x, y, psi, v);
}

struct Box {  // 32 bytes, two float4 in shared memory
    float x, y, hl, hw, c, s, present, r;
};

__device__ __forceinline__ Box tde_make_box(float x, float y, float psi, float length, float width, float present) {
    Box b;
    b.x = x; b.y = y; b.hl = 0.5f * length; b.hw = 0.5f * width;
    tde_sincosf(psi, b.s, b.c);
    b.present = present;
    b.r = b.hl + b.hw;  // bound on the circumradius, only used for conservative rejection
    return b;
}

// Oriented-box overlap with positive area (CollisionMetric.nograd, gym_env.py:48): SAT over the four
// face normals; bitwise symmetric in (A, B).
__device__ __forceinline__ bool tde_overlap(const Box& A, const Box& B) {
    float dx = B.x - A.x, dy = B.y - A.y;
    float cc = A.c * B.c + A.s * B.s;
    float ss = A.c * B.s - A.s * B.c;
    float acc = fabsf(cc), ass = fabsf(ss);
    bool ok = fabsf(dx * A.c + dy * A.s) < A.hl + (B.hl * acc + B.hw * ass);
    ok = ok && (fabsf(dy * A.c - dx * A.s) < A.hw + (B.hl * ass + B.hw * acc));
    ok = ok && (fabsf(dx * B.c + dy * B.s) < B.hl + (A.hl * acc + A.hw * ass));
    ok = ok && (fabsf(dy * B.c - dx * B.s) < B.hw + (A.hl * ass + A.hw * acc));
    return ok;
}

// conservative "cannot overlap": centres further apart than the sum of circumradius bounds + 1 cm
__device__ __forceinline__ bool tde_far_apart(const Box& A, const Box& B) {
    float dx = B.x - A.x, dy = B.y - A.y;
    float R = A.r + B.r + 0.01f;
    return dx * dx + dy * dy > R * R * 1.001f;
}

__device__ __forceinline__ void tde_box_corner(const Box& b, int k, float& px, float& py) {
    float ox = (k == 0 || k == 1) ? b.hl : -b.hl;
    float oy = (k == 0 || k == 3) ? b.hw : -b.hw;
    px = b.x + (ox * b.c - oy * b.s);
    py = b.y + (ox * b.s + oy * b.c);
}

__device__ __forceinline__ void tde_edge_terms(float ax, float ay, float bx, float by, float il, float px,
                                               float py, float& d2, float& cr) {
    float abx = bx - ax, aby = by - ay, apx = px - ax, apy = py - ay;
    cr = abx * apy - aby * apx;
    float t = (apx * abx + apy * aby) * il;
    t = fminf(fmaxf(t, 0.0f), 1.0f);
    float qx = apx - t * abx, qy = apy - t * aby;
    d2 = qx * qx + qy * qy;
}

// Triangle record: 3 float4  [ax ay bx by] [cx cy il_ab il_bc] [il_ca dir_cos dir_sin 0], read from the global table
// (read-only path) or from its copy staged in shared memory
struct Tri3 { float4 t0, t1, t2; };
template <bool STAGED>
__device__ __forceinline__ Tri3 tde_load_tri(const float4* __restrict__ g, uint32_t s, int t) {
    Tri3 r;
    if (STAGED) {
        const uint32_t a = s + (uint32_t)t * 48u;
        r.t0 = tde_lds_f4(a); r.t1 = tde_lds_f4(a + 16u); r.t2 = tde_lds_f4(a + 32u);
    } else {
        const float4* q = g + 3 * t;
        r.t0 = __ldg(q); r.t1 = __ldg(q + 1); r.t2 = __ldg(q + 2);
    }
    return r;
}
// containment (edge functions with a fixed operand order), lane direction of the triangle in (dc, ds)
__device__ __forceinline__ bool tde_tri_contains(const Tri3& T, float px, float py, float& dc, float& ds) {
    const float4 t0 = T.t0, t1 = T.t1;
    float c0 = (t0.z - t0.x) * (py - t0.y) - (t0.w - t0.y) * (px - t0.x);
    float c1 = (t1.x - t0.z) * (py - t0.w) - (t1.y - t0.w) * (px - t0.z);
    float c2 = (t0.x - t1.x) * (py - t1.y) - (t0.y - t1.y) * (px - t1.x);
    dc = T.t2.y; ds = T.t2.z;
    return (c0 >= 0.0f && c1 >= 0.0f && c2 >= 0.0f) || (c0 <= 0.0f && c1 <= 0.0f && c2 <= 0.0f);
}
// min squared distance to the three edges (the distance of a point outside the triangle)
__device__ __forceinline__ float tde_tri_segdist2(const Tri3& T, float px, float py) {
    const float4 t0 = T.t0, t1 = T.t1, t2 = T.t2;
    float d0, d1, d2, c;
    tde_edge_terms(t0.x, t0.y, t0.z, t0.w, t1.z, px, py, d0, c);
    tde_edge_terms(t0.z, t0.w, t1.x, t1.y, t1.w, px, py, d1, c);
    tde_edge_terms(t1.x, t1.y, t0.x, t0.y, t2.x, px, py, d2, c);
    return fminf(fminf(d0, d1), d2);
}

// ---- counter-based RNG shared (as a specification) with the oracle
__host__ __device__ __forceinline__ uint64_t tde_mix64(uint64_t x) {
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ULL;
    x ^= x >> 27; x *= 0x94D049BB133111EBULL;
    x ^= x >> 31;
    return x;
}
__host__ __device__ __forceinline__ uint64_t tde_rng(uint64_t seed, uint64_t genv, uint64_t episode, uint64_t k) {
    uint64_t x = seed + 0x9E3779B97F4A7C15ULL * (genv + 1);
    x = tde_mix64(x);
    x = tde_mix64(x + 0x9E3779B97F4A7C15ULL * (episode + 1));
    x = tde_mix64(x + 0x9E3779B97F4A7C15ULL * (k + 1));
    return x;
}
__device__ __forceinline__ float tde_u01(uint64_t r) { return (float)(uint32_t)(r >> 40) * 5.9604644775390625e-08f; }
__device__ __forceinline__ float tde_normal8(uint64_t r0, uint64_t r1) {
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        sum += (uint32_t)((r0 >> (16 * i)) & 0xFFFF);
        sum += (uint32_t)((r1 >> (16 * i)) & 0xFFFF);
    }
    float u = ((float)sum + 4.0f) * 1.52587890625e-05f - 4.0f;
    return u * 1.22474487139158894f;
}

// exact floor(a / d) and remainder for d > 0 from a float estimate refined twice in integers
// (inv_d ~ 1/d, any approximation good to ~2^-20 relative)
__device__ __forceinline__ int tde_floordiv(int a, int d, float inv_d, int& rem) {
    int q = __float2int_rd((float)a * inv_d);
    int r = a - q * d;                       // |r| <= ~17 d after the first estimate
    int q2 = __float2int_rd((float)r * inv_d);
    q += q2; r -= q2 * d;                    // now off by at most one
    if (r < 0) { q -= 1; r += d; }
    else if (r >= d) { q += 1; r -= d; }
    rem = r;
    return q;
}
