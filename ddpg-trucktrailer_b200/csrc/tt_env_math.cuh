// tt_env_math.cuh -- per-environment arithmetic of the fused env kernel (one thread = one environment).
//
// What it computes = truck_trailer_sim/simv2.py step()/reset() + reward_functionv1.py, see the line
// references on each block below (paths relative to the reference root).  HOW it computes it is new:
//
//   * The reference integrates the 6-state kinematic model with scipy RK45 in float64.  For this ODE and
//     dt = 0.08 s that is one Dormand-Prince-5 step in ~90 % of calls; the kernel
//     always takes exactly one DP5 step (difference to the adaptive solver <= 1.2e-8 m over an episode).
//   * Reversing a trailer is an UNSTABLE system: an error e in the hitch angle theta = psi1 - psi2 grows like
//     exp(|v|/L2 * t) = up to ~5e4 over a 15 s episode.  Float32 rounding (1e-7) in the theta dynamics would
//     therefore end up far outside the 1e-4 parity bar, while errors in the position integrals are not fed
//     back at all.  So the 1-D theta ODE (psi1' is constant within a step) is integrated in float64 with
//     short polynomial rotations of (sin theta0, cos theta0) -- no libm call, ~215 DFMA-class ops -- and
//     the four position integrals use float32 stage values; the positions themselves are kept as 32-bit FIXED
//     POINT (2^-25 m = 3e-8 m steps over +-64 m): uniformly finer than float32 at map scale, half the bytes of
//     float64, exact integer differences/comparisons, and no float64 arithmetic on the position path.
//   * All sines/cosines of the stage angles and of the new state come from rotating one base pair per
//     angle by the (small) stage increment; the observation needs no further trig and no atan2:
//     sin/cos(heading_error) = -dy_local/d, -dx_local/d and the orientation tests compare cosines.
//   * reward_functionv1 is evaluated in float32 on float64-derived distances; every threshold the
//     reference evaluates on float64 state (jackknife 70/85/90 deg, map bounds, past-goal) is compared in
//     float64 here as well.
//
// The file is plain C++ that also compiles for the host: tests/host_math builds it with g++ to measure
// the numerical error against the float64 oracle on the CPU (test infrastructure; the product only ever
// runs the CUDA build).
#pragma once
#include <math.h>
#include <stdint.h>
#if !defined(__CUDACC__)
#include <algorithm>
using std::max;
using std::min;
#endif

#if defined(__CUDACC__)
#define TT_HD __host__ __device__ __forceinline__
#else
#define TT_HD static inline
#endif

namespace ttm {

// ---------------------------------------------------------------------------------------------------
// float64 literals.  As immediates every double costs two UMOV (32-bit halves) in front of the DFMA that uses
// it, which doubled the instruction count of the float64 polynomial sections; from __constant__ memory the
// DFMA reads them as a constant-bank operand for free.  The host build uses the same values from a plain array.
// ---------------------------------------------------------------------------------------------------
#define TTM_F64_TABLE                                                                                                  \
    /* 0..5   fdlibm __kernel_sin S1..S6 */                                                                             \
    -1.66666666666666324348e-01, 8.33333333332248946124e-03, -1.98412698298579493134e-04, 2.75573137070700676789e-06,  \
    -2.50507602534068634195e-08, 1.58969099521155010221e-10,                                                           \
    /* 6..11  fdlibm __kernel_cos C1..C6 */                                                                             \
    4.16666666666666019037e-02, -1.38888888888741095749e-03, 2.48015872894767294178e-05, -2.75573143513906633035e-07,  \
    2.08757232129817482790e-09, -1.13596475577881948265e-11,                                                           \
    /* 12..14 2/pi, pi/2 hi, pi/2 lo */                                                                                 \
    6.36619772367581382433e-01, 1.57079632679489655800e+00, 6.12323399573676603587e-17,                                \
    /* 15..18 Taylor sin: -1/3!, 1/5!, -1/7!, 1/9! */                                                                   \
    -1.6666666666666666e-01, 8.3333333333333332e-03, -1.9841269841269841e-04, 2.7557319223985893e-06,                  \
    /* 19..22 Taylor cos: -1/2!, 1/4!, -1/6!, 1/8! */                                                                   \
    -0.5, 4.1666666666666664e-02, -1.3888888888888889e-03, 2.4801587301587302e-05,                                     \
    /* 23..37 Dormand-Prince a21 | a31 a32 | a41 a42 a43 | a51..a54 | a61..a65 (scipy rk.py:541-549) */                 \
    1.0 / 5, 3.0 / 40, 9.0 / 40, 44.0 / 45, -56.0 / 15, 32.0 / 9, 19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561,     \
    -212.0 / 729, 9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656,                              \
    /* 38..42 c2..c6 */                                                                                                  \
    1.0 / 5, 3.0 / 10, 4.0 / 5, 8.0 / 9, 1.0,                                                                            \
    /* 43..48 b1..b6 (rk.py:550) */                                                                                      \
    35.0 / 384, 0.0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84
static const double kF64Host[] = {TTM_F64_TABLE};
#if defined(__CUDACC__)
static __constant__ double kF64Dev[] = {TTM_F64_TABLE};
#endif
#if defined(__CUDA_ARCH__)
#define TTM_K(i) kF64Dev[i]
#else
#define TTM_K(i) kF64Host[i]
#endif
enum { K_S = 0, K_C = 6, K_2OPI = 12, K_PIO2H = 13, K_PIO2L = 14, K_TS = 15, K_TC = 19, K_A = 23, K_CN = 38, K_B = 43 };

// ---------------------------------------------------------------------------------------------------
// constants derived once on the host from tt_env_cfg (passed to kernels by value -> constant bank)
// ---------------------------------------------------------------------------------------------------
struct StepConsts {
    double vL1, vL2;            // v1x / L1, v1x / L2                      simv2.py:284,291
    double h;                   // dt                                       simv2.py:40
    double steer_max;           // clip bound                               simv2.py:504
    double max_hitch;           // 90 deg                                   simv2.py:305-310
    double jk_major, jk_minor;  // 85 / 70 deg                              reward_functionv1.py:381-387
    double map_min, map_max;    //                                          simv2.py:315-326
    double step_len;            // 0.40096                                  simv2.py:265
    double L2;
    double sx_lo, sx_w, sy_lo, sy_w, syaw_lo, syaw_w;   // start-pose box   simv2.py:331-333
    double gx, gy, gyaw;        // default goal                             simv2.py:335-337
    float  hv_fix;              // dt * v1x * 2^25   (position increment in fixed-point units)
    float  pos_inv;             // 2^-25
    double pos_scale_d;         // 2^25
    int32_t map_min_fix, map_max_fix, maj_lo_fix, maj_hi_fix;   // map bounds and the +-2 m "major" band, fixed point
    int32_t gx_fix, gy_fix;     // default goal, fixed point
    float  sgy0, cgy0;          // sin / cos of the default goal yaw
    float  mid, inv_half;       // map centre, 2 / width                    simv2.py:109-114
    float  inv_maxd;            // 1 / max_expected_distance                simv2.py:57
    float  pos_thr;             // 0.5                                      simv2.py:97
    float  cos_ori_thr;         // cos(15 deg)   (|atan2(s,c)| <= a  <=>  c >= cos a, a < 90 deg)
    float  cos_45;              //                                          reward_functionv1.py:357
};

// packed per-env word: steps | emax | (rmax - emax) | stage2 | stage3 | finished
enum : uint32_t {
    PK_STEPS_MASK = 0xFFFu, PK_EMAX_SHIFT = 12, PK_EMAX_MASK = 0xFFFu, PK_RMAX_EXTRA = 1u << 24,
    PK_ST2 = 1u << 25, PK_ST3 = 1u << 26, PK_FINISHED = 1u << 27
};

// per-env persistent state, as held in registers
struct EnvRegs {
    double psi1, psi2;                         // simv2.py:489 state: headings (float64) ...
    int32_t x1, y1, x2, y2;                    // ... and positions (fixed point, 2^-25 m)
    int32_t gx, gy;                            // goal position (fixed point)
    float  sgy, cgy;                           // sin/cos(goal yaw)
    float  d0;                                 // hypot(goal - start), simv2.py:264 / reward_functionv1.py:34
    float  closest, cum, first_steer;          // reward_functionv1.py:99-109
    // distance_history / previous_distance (reward_functionv1.py:40-67) are kept as the last three per-step
    // distance DECREMENTS g_k = d_{k-1} - d_k (g1 = most recent): the reward only ever uses differences of
    // distances, and a float32 decrement (~0.4 m) is 100x more precise than a float32 distance (~60 m).
    float  g1, g2, g3;
    float  ep_ret;                             // running episode return (trainv2.py:529 `score`)
    uint32_t packed;
    // trailer length of THIS env and v1x / L2 (heatmap.py:89 assigns env.L2 per trial).  Filled from StepConsts unless
    // per-env values were injected (tt_env_set_l2); they are episode-independent and survive reset_from_pose.
    double L2, vL2;
};

TT_HD void use_default_l2(const StepConsts &k, EnvRegs &e) { e.L2 = k.L2; e.vL2 = k.vL2; }

struct StepOut {
    float obs[23];
    float reward;
    float comps[10];     // distance, progress, heading, orientation, staged, safety, exploration, final, backward, smoothness
    uint32_t flags;      // TT_F_* bits
    uint32_t viol;       // TT_V_*
    bool done, success;
};

// ---------------------------------------------------------------------------------------------------
// polynomial kernels (no libm: identical algorithm on host and device)
// ---------------------------------------------------------------------------------------------------
// float64 sin/cos on |r| <= pi/4 (fdlibm __kernel_sin/__kernel_cos minimax coefficients, < 1 ulp)
TT_HD void sincos_pio4_f64(double r, double &s, double &c) {
    const double z = r * r;
    double ps = fma(z, TTM_K(K_S + 5), TTM_K(K_S + 4));
    ps = fma(z, ps, TTM_K(K_S + 3));
    ps = fma(z, ps, TTM_K(K_S + 2));
    ps = fma(z, ps, TTM_K(K_S + 1));
    ps = fma(z, ps, TTM_K(K_S + 0));
    s = fma(r * z, ps, r);
    double pc = fma(z, TTM_K(K_C + 5), TTM_K(K_C + 4));
    pc = fma(z, pc, TTM_K(K_C + 3));
    pc = fma(z, pc, TTM_K(K_C + 2));
    pc = fma(z, pc, TTM_K(K_C + 1));
    pc = fma(z, pc, TTM_K(K_C + 0));
    c = fma(z * z, pc, fma(z, -0.5, 1.0));
}

// reduce x to r in [-pi/4, pi/4], quadrant q (x = r + q*pi/2); |x| up to ~1e5 keeps < 1e-15 error
TT_HD double reduce_pio2(double x, int &q) {
    const double k = rint(x * TTM_K(K_2OPI));
    q = (int)k;
    double r = fma(-k, TTM_K(K_PIO2H), x);
    return fma(-k, TTM_K(K_PIO2L), r);
}

TT_HD void sincos_f64(double x, double &s, double &c) {
    int q;
    const double r = reduce_pio2(x, q);
    double sr, cr;
    sincos_pio4_f64(r, sr, cr);
    const double a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    s = (q & 2) ? -a : a;
    c = ((q + 1) & 2) ? -b : b;
}

// float32 sin/cos of a float64 angle: float64 range reduction, float32 minimax polynomials (cephes)
TT_HD void sincos_f32_of_f64(double x, float &s, float &c) {
    int q;
    const float r = (float)reduce_pio2(x, q);
    const float z = r * r;
    float ps = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
    ps = fmaf(z, ps, -1.6666654611e-1f);
    const float sr = fmaf(r * z, ps, r);
    float pc = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
    pc = fmaf(z, pc, 4.166664568298827e-2f);
    const float cr = fmaf(z * z, pc, fmaf(z, -0.5f, 1.0f));
    const float a = (q & 1) ? cr : sr, b = (q & 1) ? sr : cr;
    s = (q & 2) ? -a : a;
    c = ((q + 1) & 2) ? -b : b;
}

// Taylor sin/cos of a SMALL float64 angle (|d| <~ 0.3; truncation < 4e-14 there, < 2e-17 at 0.15)
TT_HD void sincos_small_f64(double d, double &s, double &c) {
    const double z = d * d;
    double ps = fma(z, TTM_K(K_TS + 3), TTM_K(K_TS + 2));
    ps = fma(z, ps, TTM_K(K_TS + 1));
    ps = fma(z, ps, TTM_K(K_TS + 0));
    s = fma(d * z, ps, d);
    double pc = fma(z, TTM_K(K_TC + 3), TTM_K(K_TC + 2));
    pc = fma(z, pc, TTM_K(K_TC + 1));
    pc = fma(z, pc, TTM_K(K_TC + 0));
    c = fma(z, pc, 1.0);
}

// Taylor sin/cos of a small float32 angle (|d| <~ 0.3: truncation < 2e-9)
TT_HD void sincos_small_f32(float d, float &s, float &c) {
    const float z = d * d;
    float ps = fmaf(z, -1.9841270e-4f, 8.3333333e-3f);
    ps = fmaf(z, ps, -1.6666667e-1f);
    s = fmaf(d * z, ps, d);
    float pc = fmaf(z, -1.3888889e-3f, 4.1666667e-2f);
    pc = fmaf(z, pc, -0.5f);
    c = fmaf(z, pc, 1.0f);
}

TT_HD float fast_tanhf(float x) {   // 1 - 2/(e^{2x}+1); abs error ~1e-7, exact limits at +-inf
#if defined(__CUDA_ARCH__)
    const float e = __expf(2.0f * x);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
#else
    const float e = expf(2.0f * x);
    return 1.0f - 2.0f / (e + 1.0f);
#endif
}

TT_HD float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }

// fixed-point positions
constexpr double kPosScale = 33554432.0;        // 2^25
TT_HD int32_t pos_from_double(double x) {
    const double v = rint(x * kPosScale);
    return v > 2147483520.0 ? 2147483520 : (v < -2147483520.0 ? -2147483520 : (int32_t)v);
}
TT_HD double pos_to_double(int32_t x) { return (double)x * (1.0 / kPosScale); }
TT_HD int32_t f2i_rn(float x) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(x);
#else
    return (int32_t)lrintf(x);
#endif
}

// ---------------------------------------------------------------------------------------------------
// observation packing: simv2.py:103-181 (index map in SURVEY.md section 8a row a5)
// trig inputs: s1/c1 = sin/cos psi1, s2/c2 = psi2, sth/cth = hitch, sdl/cdl = steering
// ---------------------------------------------------------------------------------------------------
struct ObsAux { float d, dx, dy; };

TT_HD ObsAux pack_obs(const StepConsts &k, const EnvRegs &e, float s1, float c1, float s2, float c2, float sth,
                      float cth, float sdl, float cdl, float *o) {
    const float dx = (float)(e.gx - e.x2) * k.pos_inv, dy = (float)(e.gy - e.y2) * k.pos_inv;   // exact integer differences
    const float d = sqrtf(fmaf(dx, dx, dy * dy));
    const float dxl = fmaf(dx, c2, dy * s2), dyl = fmaf(dy, c2, -dx * s2);
    o[0] = ((float)e.x1 * k.pos_inv - k.mid) * k.inv_half;  o[1] = ((float)e.y1 * k.pos_inv - k.mid) * k.inv_half;
    o[2] = s1;  o[3] = c1;
    o[4] = ((float)e.x2 * k.pos_inv - k.mid) * k.inv_half;  o[5] = ((float)e.y2 * k.pos_inv - k.mid) * k.inv_half;
    o[6] = s2;  o[7] = c2;
    o[8] = sth; o[9] = cth; o[10] = sdl; o[11] = cdl;
    o[12] = ((float)e.gx * k.pos_inv - k.mid) * k.inv_half; o[13] = ((float)e.gy * k.pos_inv - k.mid) * k.inv_half;
    o[14] = e.sgy; o[15] = e.cgy;
    o[16] = clampf(d * k.inv_maxd, 0.0f, 1.0f);
    o[17] = clampf(dxl * k.inv_maxd, -1.0f, 1.0f);
    o[18] = clampf(dyl * k.inv_maxd, -1.0f, 1.0f);
    o[19] = fmaf(e.sgy, c2, -e.cgy * s2);      // sin(goal_yaw - psi2)
    o[20] = fmaf(e.cgy, c2, e.sgy * s2);       // cos(goal_yaw - psi2)
    // heading_error = atan2(dy, dx) - (psi2 + pi): sin = -dy_local/d, cos = -dx_local/d (d == 0: atan2(0,0) = 0)
    if (d > 0.0f) {
        const float inv = 1.0f / d;
        o[21] = -dyl * inv; o[22] = -dxl * inv;
    } else { o[21] = s2; o[22] = -c2; }
    ObsAux a; a.d = d; a.dx = dx; a.dy = dy;
    return a;
}

// ---------------------------------------------------------------------------------------------------
// episode start: simv2.py:481-496 (pose -> float32 state), :263-267 (max steps), reward_functionv1.py:38
// ---------------------------------------------------------------------------------------------------
TT_HD uint32_t pack_limits(const StepConsts &k, double d0) {
    int emax = (int)(d0 / k.step_len) + 75;
    int rmax = (int)((d0 + 1e-6) / k.step_len) + 75;
    if (emax > 4094) emax = 4094;
    if (emax < 0) emax = 0;
    return ((uint32_t)emax << PK_EMAX_SHIFT) | (rmax > emax ? PK_RMAX_EXTRA : 0u);
}

// start a fresh episode from an explicit state (set_state) -- also the tail of reset
TT_HD void begin_episode(const StepConsts &k, EnvRegs &e, double sx, double sy, double gx, double gy, double gyaw,
                         float *obs) {
    float sg, cg;
    sincos_f32_of_f64(gyaw, sg, cg);
    e.gx = pos_from_double(gx); e.gy = pos_from_double(gy); e.sgy = sg; e.cgy = cg;
    const double ddx = gx - sx, ddy = gy - sy;
    const double d0 = sqrt(ddx * ddx + ddy * ddy);
    e.d0 = (float)d0;
    e.packed = pack_limits(k, d0);
    e.closest = e.cum = e.first_steer = e.g1 = e.g2 = e.g3 = 0.0f;
    e.ep_ret = 0.0f;
    if (obs) {
        float s2, c2, s1, c1, sth, cth;
        sincos_f32_of_f64(e.psi2, s2, c2);
        sincos_f32_of_f64(e.psi1, s1, c1);
        sth = fmaf(s1, c2, -c1 * s2); cth = fmaf(c1, c2, s1 * s2);
        pack_obs(k, e, s1, c1, s2, c2, sth, cth, 0.0f, 1.0f, obs);     // steering 0, simv2.py:493
    }
}

// reset from a start pose (simv2.py:481-489): truck ahead of the trailer along its heading, float32 rounding
TT_HD void reset_from_pose(const StepConsts &k, EnvRegs &e, double sx, double sy, double syaw, double gx, double gy,
                           double gyaw, float *obs) {
    double sn, cs;
    sincos_f64(syaw, sn, cs);
    e.psi1 = e.psi2 = (double)(float)syaw;
    e.x1 = pos_from_double((double)(float)fma(e.L2, cs, sx)); e.y1 = pos_from_double((double)(float)fma(e.L2, sn, sy));
    e.x2 = pos_from_double((double)(float)sx); e.y2 = pos_from_double((double)(float)sy);
    begin_episode(k, e, sx, sy, gx, gy, gyaw, obs);
}

// Philox4x32-10; spec shared with oracle/tt_oracle.c (tto_philox4x32_10, tto_rng_pose, tto_rng_normal)
TT_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
TT_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t *out) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

TT_HD void rng_pose(const StepConsts &k, uint64_t seed, uint32_t gid, uint32_t t, double &sx, double &sy, double &syaw) {
    uint32_t w[4];
    philox4x32_10(gid, t, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    const double s = 1.0 / 4294967296.0;
    const double u0 = ((double)w[0] + 0.5) * s, u1 = ((double)w[1] + 0.5) * s, u2 = ((double)w[2] + 0.5) * s;
    sx = fma(k.sx_w, u0, k.sx_lo); sy = fma(k.sy_w, u1, k.sy_lo); syaw = fma(k.syaw_w, u2, k.syaw_lo);
}

// The same generator with the 10 round keys (k + r * Weyl constant) expanded once on the host: passed by value as a kernel
// parameter they are constant-bank operands of the round's XOR, which saves 20 integer adds per call (OU noise kernel).
struct PhiloxKeys { uint32_t k0[10], k1[10]; };
inline PhiloxKeys philox_expand_key(uint64_t seed) {
    PhiloxKeys ks;
    uint32_t a = (uint32_t)seed, b = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; r++) { ks.k0[r] = a; ks.k1[r] = b; a += 0x9E3779B9u; b += 0xBB67AE85u; }
    return ks;
}
TT_HD float rng_normal_ks(const PhiloxKeys &ks, uint32_t gid, uint32_t t) {
    uint32_t c0 = gid, c1 = t, c2 = 1u, c3 = 0u;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int r = 0; r < 10; r++) {
        const uint32_t h0 = mulhi32(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        const uint32_t h1 = mulhi32(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ ks.k0[r]; c1 = l1; c2 = h0 ^ c3 ^ ks.k1[r]; c3 = l0;
    }
    const float u1 = (float)((c0 >> 8) + 1u) * (1.0f / 16777216.0f), u2 = (float)(c1 >> 8) * (1.0f / 16777216.0f);
#if defined(__CUDA_ARCH__)
    return sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2);
#else
    return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
#endif
}

TT_HD float rng_normal(uint64_t seed, uint32_t gid, uint32_t t) {
    uint32_t w[4];
    philox4x32_10(gid, t, 1u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), w);
    const float u1 = (float)((w[0] >> 8) + 1u) * (1.0f / 16777216.0f), u2 = (float)(w[1] >> 8) * (1.0f / 16777216.0f);
#if defined(__CUDA_ARCH__)
    return sqrtf(-2.0f * __logf(u1)) * cospif(2.0f * u2);
#else
    return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
#endif
}

// OUActionNoise.__call__ (DDPG/noise.py:12-17) with theta = 0.2, mu = 0, dt = 1e-2, sigma = 0.15:
// x <- x + theta (mu - x) dt + sigma sqrt(dt) N(0, 1).  Explicitly un-contracted (the same roundings wherever it is inlined:
// the stand-alone noise kernel and the actor kernels' fused output stage must agree bit for bit, and so does the oracle).
TT_HD float ou_advance(float xp, float nrm) {
#if defined(__CUDA_ARCH__)
    return __fadd_rn(__fadd_rn(xp, __fmul_rn(__fmul_rn(0.2f, __fsub_rn(0.0f, xp)), 0.01f)), __fmul_rn(0.15f * 0.1f, nrm));
#else
    const float t = (0.2f * (0.0f - xp)) * 0.01f;
    const float u = xp + t;
    return u + (0.15f * 0.1f) * nrm;
#endif
}

// ---------------------------------------------------------------------------------------------------
// one env step: simv2.py:499-545 + reward_functionv1.py (restated in SURVEY.md appendix A)
// `action` = scaled steering angle as passed to env.step (trainv2.py:516-520)
// ---------------------------------------------------------------------------------------------------
template <bool kWantComps>
TT_HD void env_step(const StepConsts &k, EnvRegs &e, float action, StepOut &out) {
    // ---- simv2.py:504-505 clip (float64 bound: a saturated float32 action becomes float64 pi/4) ----
    double delta = (double)action;
    delta = delta < -k.steer_max ? -k.steer_max : (delta > k.steer_max ? k.steer_max : delta);
    double sdl, cdl;
    sincos_pio4_f64(delta, sdl, cdl);                        // |delta| <= steer_max <= pi/4: no range reduction
    // tan(delta) = sdl / cdl without a float64 division: float32 quotient refined by one Newton step (err ~1e-14)
    const float cf = (float)cdl;
    const float t0f = (float)sdl / cf;
    const double t0 = (double)t0f;
    const double tand = fma(fma(-t0, cdl, sdl), (double)(1.0f / cf), t0);
    const double hw = k.h * (k.vL1 * tand);                  // psi1 increment: psi1' = (v/L1) tan(delta) is constant

    // ---- base trig ----
    double S0, C0;
    sincos_f64(e.psi1 - e.psi2, S0, C0);                     // hitch angle theta0
    float s2b, c2b;
    sincos_f32_of_f64(e.psi2, s2b, c2b);
    const float S0f = (float)S0, C0f = (float)C0;

    // ---- Dormand-Prince 5 (scipy rk.py:541-550 tableau, TTM_K), theta/psi2 in float64, positions in float32 ----
    double u[6];                                             // psi2' at the stages = (v/L2) sin(theta_j)
    u[0] = e.vL2 * S0;
    float ax1, ay1, ax2, ay2;                                // sum_j b_j * (unit velocity components)
    {
        const float s1b = fmaf(S0f, c2b, C0f * s2b), c1b = fmaf(C0f, c2b, -S0f * s2b);
        const float b = (float)(35.0 / 384);
        ax1 = b * c1b; ay1 = b * s1b; ax2 = b * (C0f * c2b); ay2 = b * (C0f * s2b);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int j = 1; j < 6; j++) {
        double acc = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int m = 0; m < j; m++) acc = fma(TTM_K(K_A + j * (j - 1) / 2 + m), u[m], acc);
        const double dpsi2 = k.h * acc;
        const double dth = fma(TTM_K(K_CN + j - 1), hw, -dpsi2);
        double sd, cd;
        sincos_small_f64(dth, sd, cd);
        const double sth = fma(S0, cd, C0 * sd);
        u[j] = e.vL2 * sth;
        if (j != 1) {                                        // b2 = 0: stage 2 does not enter the position quadrature
            const float cthf = (float)fma(C0, cd, -S0 * sd), sthf = (float)sth;
            float sp, cp;
            sincos_small_f32((float)dpsi2, sp, cp);
            const float s2j = fmaf(s2b, cp, c2b * sp), c2j = fmaf(c2b, cp, -s2b * sp);
            const float s1j = fmaf(sthf, c2j, cthf * s2j), c1j = fmaf(cthf, c2j, -sthf * s2j);
            const float b = j == 2 ? (float)(500.0 / 1113) : j == 3 ? (float)(125.0 / 192) : j == 4 ? (float)(-2187.0 / 6784) : (float)(11.0 / 84);
            ax1 = fmaf(b, c1j, ax1); ay1 = fmaf(b, s1j, ay1);
            ax2 = fmaf(b, cthf * c2j, ax2); ay2 = fmaf(b, cthf * s2j, ay2);
        }
    }
    double accb = 0.0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int m = 0; m < 6; m++) if (m != 1) accb = fma(TTM_K(K_B + m), u[m], accb);
    const double dpsi2 = k.h * accb;
    // goal offset of the trailer BEFORE the move (for the distance decrement below)
    const float dxp = (float)(e.gx - e.x2) * k.pos_inv, dyp = (float)(e.gy - e.y2) * k.pos_inv;
    const int32_t ix2i = f2i_rn(k.hv_fix * ax2), iy2i = f2i_rn(k.hv_fix * ay2);
    const float ix2 = (float)ix2i * k.pos_inv, iy2 = (float)iy2i * k.pos_inv;      // the increment actually applied
    e.psi1 += hw; e.psi2 += dpsi2;                           // simv2.py:516-517
    e.x1 += f2i_rn(k.hv_fix * ax1); e.y1 += f2i_rn(k.hv_fix * ay1);
    e.x2 += ix2i; e.y2 += iy2i;

    // ---- trig of the new state (for the observation) ----
    double sdn, cdn;
    sincos_small_f64(hw - dpsi2, sdn, cdn);
    const double thn_s = fma(S0, cdn, C0 * sdn), thn_c = fma(C0, cdn, -S0 * sdn);
    const float sth = (float)thn_s, cth = (float)thn_c;
    float sp, cp;
    sincos_small_f32((float)dpsi2, sp, cp);
    const float s2 = fmaf(s2b, cp, c2b * sp), c2 = fmaf(c2b, cp, -s2b * sp);
    const float s1 = fmaf(sth, c2, cth * s2), c1 = fmaf(cth, c2, -sth * s2);

    // ---- observation, simv2.py:519 ----
    const ObsAux oa = pack_obs(k, e, s1, c1, s2, c2, sth, cth, (float)sdl, (float)cdl, out.obs);
    const float d = oa.d;

    // ---- reward_functionv1.__init__ :6-97 ----
    uint32_t steps = (e.packed & PK_STEPS_MASK) + 1u;        // simv2.py:523
    const uint32_t emax = (e.packed >> PK_EMAX_SHIFT) & PK_EMAX_MASK;
    const uint32_t rmax = emax + ((e.packed & PK_RMAX_EXTRA) ? 1u : 0u);
    const float steer = (float)delta;                        // == arctan2(obs[10], obs[11]) to float32 rounding (:37)
    // g = previous_distance - current_distance = (dp^2 - d^2) / (dp + d), with dp^2 - d^2 expanded through the
    // exact float32 position increment: no cancellation, abs error ~1e-7 instead of ~8e-6
    float g;
    if (steps == 1u) {                                       // :40-76 first step of the episode: previous = current
        g = 0.0f; e.first_steer = steer; e.cum = 0.0f; e.g1 = e.g2 = e.g3 = 0.0f; e.closest = d;
        e.packed &= ~(PK_ST2 | PK_ST3);
    } else {
        const float dp = sqrtf(fmaf(dxp, dxp, dyp * dyp));
        const float num = fmaf(ix2, dxp + oa.dx, iy2 * (dyp + oa.dy));
        g = (dp + d) > 0.0f ? num / (dp + d) : 0.0f;
        e.closest = fminf(e.closest, d);                     // :70-74
    }
    // ---- compute_dynamic_weights :189-238 ----
    const float d0 = e.d0 + 1e-6f;
    const float jp = clampf((d0 - d) / d0, 0.0f, 1.0f);
    const float w_final = 0.5f * (fast_tanhf(7.0f * (jp - 0.3f)) + 1.0f), w_head = 1.0f - w_final;
    // ---- calculate_progress_reward :144-187 ----
    float prog = fast_tanhf(g);                              // instant_progress = previous - current
    prog = g > 0.0f ? prog : 0.5f * prog;
    prog += 0.5f * fast_tanhf(0.5f * (((g + e.g1) + e.g2) + e.g3));   // history[0] - current = d_{t-4} - d_t
    prog += (e.g1 >= 0.0f && g >= 0.0f) ? 0.2f : 0.0f;      // history[2] >= history[3] >= history[4]
    // ---- heading / orientation :285-324 ----
    const float heading = out.obs[22], orient = out.obs[20];
    // ---- staged success :338-367 (|atan2(o19,o20)| <= a  <=>  o20 >= cos a) ----
    float staged = d <= 5.0f ? 10.0f : 0.0f;
    if (d <= 2.0f && orient >= k.cos_45 && !(e.packed & PK_ST2)) { staged += 25.0f; e.packed |= PK_ST2; }
    const bool success = d <= k.pos_thr && orient >= k.cos_ori_thr;            // :466-470, simv2.py:533-536
    if (success && !(e.packed & PK_ST3)) { staged += 100.0f; e.packed |= PK_ST3; }
    // ---- safety penalties :369-421 (float64 comparisons on float64 state, like the reference) ----
    float saf = 0.0f; uint32_t viol = 0u;
    const double th = fabs(e.psi1 - e.psi2);
    if (th > k.jk_major) { saf += -500.0f; viol = 1u; }
    else if (th > k.jk_minor) { saf += -50.0f; viol = 2u; }
    const int32_t mn = min(min(e.x1, e.y1), min(e.x2, e.y2)), mx = max(max(e.x1, e.y1), max(e.x2, e.y2));
    const bool oom = mn < k.map_min_fix || mx > k.map_max_fix;
    if (mn < k.maj_lo_fix || mx > k.maj_hi_fix) { saf += -500.0f; viol = 3u; }
    else if (oom) { saf += -50.0f; viol = 4u; }
    const bool passed = e.gy > e.y2;
    if (passed) { saf += -500.0f; viol = 5u; }
    if (steps >= rmax) { saf += -500.0f; viol = 6u; }
    const bool exb = d > e.closest + 6.0f;                   // :120-124
    if (exb) { saf += -500.0f; viol = 7u; }
    // ---- exploration bonus :423-439 (float64 products like the reference) ----
    const double rm = (double)rmax, sd_ = (double)steps;
    const float expl = sd_ < rm * 0.5 ? 4.0f : (sd_ < rm * 0.8 ? 2.0f : 0.0f);
    // ---- backward movement penalty :240-283 ----
    e.cum += fmaxf(0.0f, -g);
    const float budget = 5.0f * fminf(1.0f, (float)steps * 0.02f);
    const float ex = fmaxf(0.0f, e.cum - budget);
    const float back = -0.5f * ex * sqrtf(ex);
    // ---- steering smoothness :326-335 (previous_steering is frozen at the first step, :45-48,:108) ----
    const float smooth = fabsf(steer - e.first_steer) * 0.63661977236758134f;
    const float fin = success ? 200.0f : 0.0f;
    // ---- history / previous distance :166-171, :472 ----
    e.g3 = e.g2; e.g2 = e.g1; e.g1 = g;
    // ---- total :475-486 ----
    const float c_prog = 15.0f * prog, c_head = 15.0f * heading * w_head, c_ori = 15.0f * orient * w_final,
                c_smooth = -25.0f * smooth;
    out.reward = c_prog + c_head + c_ori + staged + saf + expl + back + c_smooth + fin;
    if (kWantComps) {
        out.comps[0] = 0.0f; out.comps[1] = c_prog; out.comps[2] = c_head; out.comps[3] = c_ori; out.comps[4] = staged;
        out.comps[5] = saf; out.comps[6] = expl; out.comps[7] = fin; out.comps[8] = back; out.comps[9] = c_smooth;
    }
    // ---- termination simv2.py:528-541 ----
    uint32_t fl = 0u;
    fl |= th > k.max_hitch ? 1u : 0u;
    fl |= oom ? 2u : 0u;
    fl |= steps >= emax ? 4u : 0u;
    fl |= success ? 8u : 0u;
    fl |= passed ? 16u : 0u;
    fl |= exb ? 32u : 0u;
    out.flags = fl; out.viol = viol; out.done = fl != 0u; out.success = success;
    if (steps > PK_STEPS_MASK) steps = PK_STEPS_MASK;
    e.packed = (e.packed & ~PK_STEPS_MASK) | steps;
    e.ep_ret += out.reward;
}

}  // namespace ttm
