/*
 * tt_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement (plain C, float64 like the reference) of the
 * hot path of pain7576/ddpg-trucktrailer: simv2 env reset/step, reward_functionv1, OU noise, the actor
 * forward and ReplayBuffer.store_transition.  It is the checker for the CUDA kernels in
 * ddpg-trucktrailer_b200/csrc; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load it.  The product package never does.
 *
 * PARITY PINNING: the reference ships no tests.  This file is pinned against
 *   (1) the recorded episode DDPG/episode_replays/episode_10579_reward_4792.pkl (193 steps; committed as
 *       tests/golden/episode_10579.npz by oracle/make_golden.py), and
 *   (2) outputs of the untouched reference imported in the build container (oracle/ref_harness.py ->
 *       tests/golden/ref_rollouts.npz, tests/golden/ref_actor.npz).
 * See tests/test_oracle_golden.py and tests/test_oracle_vs_reference.py.
 *
 * Third-party arithmetic restated here: scipy.integrate.solve_ivp(method='RK45') -- scipy is an unpinned
 * dependency of the reference (README.md:76); the build container has scipy 1.18.1 and this file follows
 * scipy/integrate/_ivp/rk.py (rk_step :14-71, RungeKutta.__init__ :85-98, _step_impl :108-182, RK45 tableau
 * :541-550), common.py (norm :63-65, select_initial_step :68-133) and ivp.py's step loop.
 *
 * Every function cites the reference file:line it follows (paths relative to the reference root).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TTO_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ config ------------------------------ */
typedef struct {
    double L1, L2, v1x, dt;           /* simv2.py:34-40 */
    double map_min, map_max;          /* simv2.py:25-28 */
    double max_hitch;                 /* simv2.py:49 (90 deg) */
    double steer_max;                 /* simv2.py:51-52 (45 deg) */
    double pos_thr, ori_thr;          /* simv2.py:97-98 */
    double step_len;                  /* 0.40096, simv2.py:265 / reward_functionv1.py:38 */
    double max_expected_distance;     /* simv2.py:57 */
    int    integrator;                /* 0 = scipy adaptive RK45 (reference), 1 = one fixed DP5 step */
} tto_cfg;

TTO_API void tto_default_cfg(tto_cfg *c) {
    c->L1 = 5.0; c->L2 = 7.0; c->v1x = -5.012; c->dt = 0.08;
    c->map_min = -40.0; c->map_max = 40.0;
    c->max_hitch = 90.0 * M_PI / 180.0;
    c->steer_max = 45.0 * M_PI / 180.0;
    c->pos_thr = 0.5; c->ori_thr = 15.0 * M_PI / 180.0;
    c->step_len = 0.40096;
    c->max_expected_distance = sqrt(80.0 * 80.0 + 80.0 * 80.0);
    c->integrator = 0;
}

typedef struct {
    double st[6];                      /* psi1, psi2, x1, y1, x2, y2  (simv2.py:489) */
    double startx, starty, startyaw;   /* simv2.py:465 */
    double gx, gy, gyaw;
    int    steps, emax;                /* episode_steps, max_episode_steps (simv2.py:491-494) */
    /* reward_functionv1 persistent state, reward_functionv1.py:99-109 */
    int    has_rs;
    double prev, cum, closest, hist[5];
    float  first_steer;                /* "previous_steering": frozen at step 1 (reward_functionv1.py:45-48,108) */
    int    bt_steps;
    int    stage[3];
} tto_env;

/* ------------------------------------------------------------------ kinematics -------------------------- */
/* simv2.py:269-303 kinematic_model (hitch_offset = 0.0 kept symbolic-free: terms with it vanish) */
static void kin(const tto_cfg *c, const double *x, double delta, double *xd) {
    double psi1 = x[0], psi2 = x[1];
    double hitch = psi1 - psi2;
    double dpsi1 = (c->v1x / c->L1) * tan(delta);
    double v2x = c->v1x * cos(hitch);
    double dpsi2 = (c->v1x / c->L2) * sin(hitch);
    xd[0] = dpsi1; xd[1] = dpsi2;
    xd[2] = c->v1x * cos(psi1); xd[3] = c->v1x * sin(psi1);
    xd[4] = v2x * cos(psi2);    xd[5] = v2x * sin(psi2);
}

/* scipy rk.py:541-550 */
static const double RK_A[6][5] = {
    {0, 0, 0, 0, 0},
    {1.0 / 5, 0, 0, 0, 0},
    {3.0 / 40, 9.0 / 40, 0, 0, 0},
    {44.0 / 45, -56.0 / 15, 32.0 / 9, 0, 0},
    {19372.0 / 6561, -25360.0 / 2187, 64448.0 / 6561, -212.0 / 729, 0},
    {9017.0 / 3168, -355.0 / 33, 46732.0 / 5247, 49.0 / 176, -5103.0 / 18656}};
static const double RK_B[6] = {35.0 / 384, 0, 500.0 / 1113, 125.0 / 192, -2187.0 / 6784, 11.0 / 84};
static const double RK_E[7] = {-71.0 / 57600, 0, 71.0 / 16695, -71.0 / 1920, 17253.0 / 339200, -22.0 / 525, 1.0 / 40};

/* scipy common.py:63-65 */
static double rms_norm6(const double *x) {
    double s = 0;
    for (int i = 0; i < 6; i++) s += x[i] * x[i];
    return sqrt(s) / sqrt(6.0);
}

/* scipy rk.py:14-71 rk_step; K has 7 rows, K[0] = f on entry */
static void rk_step(const tto_cfg *c, double delta, const double *y, double h, double K[7][6], double *ynew) {
    double yt[6];
    for (int s = 1; s < 6; s++) {
        for (int i = 0; i < 6; i++) {
            double dy = 0;
            for (int j = 0; j < s; j++) dy += K[j][i] * RK_A[s][j];
            yt[i] = y[i] + dy * h;
        }
        kin(c, yt, delta, K[s]);
    }
    for (int i = 0; i < 6; i++) {
        double acc = 0;
        for (int j = 0; j < 6; j++) acc += K[j][i] * RK_B[j];
        ynew[i] = y[i] + h * acc;
    }
    kin(c, ynew, delta, K[6]);
}

/* solve_ivp(fun, [0, dt], y0, method='RK45') as called at simv2.py:510-515; returns y(dt) in ynew.
 * nsteps/nfev are diagnostics (accepted steps, RHS evaluations). */
TTO_API void tto_rk45(const tto_cfg *c, const double *y0, double delta, double *ynew, int *nsteps, int *nfev) {
    const double rtol = 1e-3, atol = 1e-6, SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0;
    const double t_bound = c->dt, err_exp = -1.0 / 5.0;
    double t = 0.0, y[6], f[6], K[7][6];
    int fev = 0, steps = 0;
    memcpy(y, y0, sizeof y);
    kin(c, y, delta, f); fev++;
    /* select_initial_step, common.py:68-133 (order = error_estimator_order = 4, direction = +1) */
    double h_abs;
    {
        double scale[6], a[6], b[6], y1[6], f1[6];
        for (int i = 0; i < 6; i++) { scale[i] = atol + fabs(y[i]) * rtol; a[i] = y[i] / scale[i]; b[i] = f[i] / scale[i]; }
        double d0 = rms_norm6(a), d1 = rms_norm6(b), h0, h1;
        h0 = (d0 < 1e-5 || d1 < 1e-5) ? 1e-6 : 0.01 * d0 / d1;
        if (h0 > t_bound) h0 = t_bound;
        for (int i = 0; i < 6; i++) y1[i] = y[i] + h0 * f[i];
        kin(c, y1, delta, f1); fev++;
        for (int i = 0; i < 6; i++) a[i] = (f1[i] - f[i]) / scale[i];
        double d2 = rms_norm6(a) / h0;
        if (d1 <= 1e-15 && d2 <= 1e-15) h1 = fmax(1e-6, h0 * 1e-3);
        else h1 = pow(0.01 / fmax(d1, d2), 1.0 / 5.0);
        h_abs = fmin(fmin(100 * h0, h1), t_bound);   /* max_step = inf */
    }
    /* ivp.py loop: step until t == t_bound; rk.py:108-182 _step_impl */
    while (t < t_bound) {
        double min_step = 10 * fabs(nextafter(t, INFINITY) - t);
        if (h_abs < min_step) h_abs = min_step;
        int rejected = 0;
        for (;;) {
            double h = h_abs, t_new = t + h;
            if (t_new - t_bound > 0) t_new = t_bound;
            h = t_new - t; h_abs = fabs(h);
            memcpy(K[0], f, sizeof f);
            double yn[6];
            rk_step(c, delta, y, h, K, yn); fev += 6;
            double e[6];
            for (int i = 0; i < 6; i++) {
                double scale = atol + fmax(fabs(y[i]), fabs(yn[i])) * rtol, acc = 0;
                for (int j = 0; j < 7; j++) acc += K[j][i] * RK_E[j];
                e[i] = acc * h / scale;
            }
            double en = rms_norm6(e);
            if (en < 1) {
                double factor = (en == 0) ? MAX_FACTOR : fmin(MAX_FACTOR, SAFETY * pow(en, err_exp));
                if (rejected) factor = fmin(1.0, factor);
                h_abs *= factor;
                t = t_new; memcpy(y, yn, sizeof y); memcpy(f, K[6], sizeof f);
                steps++;
                break;
            }
            h_abs *= fmax(MIN_FACTOR, SAFETY * pow(en, err_exp));
            rejected = 1;
        }
    }
    memcpy(ynew, y, sizeof y);
    if (nsteps) *nsteps = steps;
    if (nfev) *nfev = fev;
}

/* one fixed Dormand-Prince-5 step of length dt (what the CUDA kernel executes) */
TTO_API void tto_dp5_fixed(const tto_cfg *c, const double *y0, double delta, double *ynew) {
    double K[7][6];
    kin(c, y0, delta, K[0]);
    rk_step(c, delta, y0, c->dt, K, ynew);
}

/* ------------------------------------------------------------------ observation ------------------------- */
static double clipd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* simv2.py:103-181 compute_observation (== compute_observation1 :183-261); float64 math, cast at the end */
TTO_API void tto_obs(const tto_cfg *c, const double *st, double steer, double gx, double gy, double gyaw, float *o) {
    double psi1 = st[0], psi2 = st[1], x1 = st[2], y1 = st[3], x2 = st[4], y2 = st[5];
    double mid = (c->map_max + c->map_min) / 2, half = (c->map_max - c->map_min) / 2;
    double hitch = psi1 - psi2;
    double dist = sqrt((x2 - gx) * (x2 - gx) + (y2 - gy) * (y2 - gy));
    double dist_n = clipd(dist / c->max_expected_distance, 0, 1);
    double ang_to_goal = atan2(gy - y2, gx - x2);
    double ori_err = gyaw - psi2;
    double dxg = gx - x2, dyg = gy - y2;
    double dxl = dxg * cos(psi2) + dyg * sin(psi2);
    double dyl = -dxg * sin(psi2) + dyg * cos(psi2);
    double head_err = ang_to_goal - (psi2 + M_PI);
    double v[23] = {
        (x1 - mid) / half, (y1 - mid) / half, sin(psi1), cos(psi1),
        (x2 - mid) / half, (y2 - mid) / half, sin(psi2), cos(psi2),
        sin(hitch), cos(hitch), sin(steer), cos(steer),
        (gx - mid) / half, (gy - mid) / half, sin(gyaw), cos(gyaw),
        dist_n, clipd(dxl / c->max_expected_distance, -1, 1), clipd(dyl / c->max_expected_distance, -1, 1),
        sin(ori_err), cos(ori_err), sin(head_err), cos(head_err)};
    for (int i = 0; i < 23; i++) o[i] = (float)v[i];
}

/* ------------------------------------------------------------------ reset / state injection ------------- */
static int max_steps_of(const tto_cfg *c, double d0) { return (int)(d0 / c->step_len) + 75; }   /* simv2.py:263-267 */

/* simv2.py:481-496: state from the trailer pose, stored as float32; obs with steering 0 */
TTO_API void tto_reset_pose(const tto_cfg *c, tto_env *e, double sx, double sy, double syaw,
                            double gx, double gy, double gyaw, float *obs) {
    memset(e, 0, sizeof *e);
    e->startx = sx; e->starty = sy; e->startyaw = syaw; e->gx = gx; e->gy = gy; e->gyaw = gyaw;
    double x1 = sx + c->L2 * cos(syaw), y1 = sy + c->L2 * sin(syaw);
    double s[6] = {syaw, syaw, x1, y1, sx, sy};
    for (int i = 0; i < 6; i++) e->st[i] = (double)(float)s[i];
    e->emax = max_steps_of(c, sqrt((gx - sx) * (gx - sx) + (gy - sy) * (gy - sy)));
    e->steps = 0; e->has_rs = 0;
    /* NOTE: with numpy>=2 scalar promotion the reference evaluates this reset observation partly in
     * float32 (the state array is float32 until the first solve_ivp); here it is float64 on the
     * float32-rounded state.  The two agree to float32 rounding (<=2.4e-7, tests/test_oracle_vs_reference). */
    if (obs) tto_obs(c, e->st, 0.0, gx, gy, gyaw, obs);
}

/* state injection as DDPG/test.py:96-115, heatmap.py:119, episode_replay_collectorv2.py:258 do it */
TTO_API void tto_set_state(const tto_cfg *c, tto_env *e, const double *st, double sx, double sy, double syaw,
                           double gx, double gy, double gyaw, float *obs) {
    memset(e, 0, sizeof *e);
    memcpy(e->st, st, sizeof e->st);
    e->startx = sx; e->starty = sy; e->startyaw = syaw; e->gx = gx; e->gy = gy; e->gyaw = gyaw;
    e->emax = max_steps_of(c, sqrt((gx - sx) * (gx - sx) + (gy - sy) * (gy - sy)));
    if (obs) tto_obs(c, e->st, 0.0, gx, gy, gyaw, obs);
}

/* ------------------------------------------------------------------ reward ------------------------------ */
/* component order used everywhere (oracle, kernels, fixtures) */
enum { C_TOTAL = 0, C_DISTANCE, C_PROGRESS, C_HEADING, C_ORIENT, C_STAGED, C_SAFETY, C_EXPLORE, C_FINAL,
       C_BACKWARD, C_SMOOTH, C_NCOMP };
/* violation_type codes, reward_functionv1.py:378-419 (last writer wins in this order) */
enum { V_NONE = 0, V_JACKKNIFE, V_JACKKNIFE_WARNING, V_MAJOR_BOUNDARY, V_MINOR_BOUNDARY, V_PAST_GOAL, V_MAX_STEP,
       V_EXCESSIVE_BACKWARD };
/* termination bits, simv2.py:528-541 */
enum { F_JACKKNIFE = 0, F_OUT_OF_MAP, F_MAX_STEPS, F_GOAL_REACHED, F_GOAL_PASSED, F_EXCESSIVE_BACKWARD, F_NFLAGS };

/* simv2.py:499-545 step(); reward_functionv1.py:6-109 (__init__), :442-506 (compute_reward) and parts.
 * action = scaled steering angle (float32, as DDPG/trainv2.py:516 passes it).  Returns done. */
/* tto_step_m additionally reports, for the termination tests, how far the float64 state is from each threshold
 * (`margins`, may be NULL; the sign is "positive = flag raised"):  [0] |theta| - 90 deg (rad), [1] distance of the closest
 * of x1,y1,x2,y2 beyond the map edge (m), [2] goal_y - y2 (m), [3] d - (closest + 6) (m), [4] pos_thr - d (m),
 * [5] ori_thr - |orientation error| (rad), [6] steps - max_episode_steps.  The parity tests accept a termination-step
 * mismatch against the CUDA path only where the differing flag's margin is inside its documented epsilon. */
enum { M_JACKKNIFE = 0, M_MAP, M_PAST, M_BACKWARD, M_GOAL_POS, M_GOAL_ORI, M_STEPS, M_NMARGINS };
TTO_API int tto_step_m(const tto_cfg *c, tto_env *e, float action, float *obs, double *comps, int *viol,
                       int *flags, int *success_out, double *margins) {
    /* simv2.py:504-505 */
    double delta = clipd((double)action, -c->steer_max, c->steer_max);
    double yn[6];
    if (c->integrator == 0) tto_rk45(c, e->st, delta, yn, 0, 0); else tto_dp5_fixed(c, e->st, delta, yn);
    memcpy(e->st, yn, sizeof yn);                                   /* :516-517 */
    tto_obs(c, e->st, delta, e->gx, e->gy, e->gyaw, obs);           /* :519-520 */
    e->steps += 1;                                                  /* :523 */
    const double *st = e->st;
    const int episode_steps = e->steps;

    /* ---- RewardFunction.__init__, reward_functionv1.py:6-97 ---- */
    double cur = sqrt((st[4] - e->gx) * (st[4] - e->gx) + (st[5] - e->gy) * (st[5] - e->gy));       /* :111-114 */
    double d0 = sqrt((e->gx - e->startx) * (e->gx - e->startx) + (e->gy - e->starty) * (e->gy - e->starty)) + 1e-6; /* :34-35 */
    float cur_steer = atan2f(obs[10], obs[11]);                     /* :37 (float32 obs -> float32 arctan2) */
    int rmax = (int)(d0 / c->step_len) + 75;                        /* :38 */
    if (!e->has_rs) {                                               /* :40-76, None branch */
        e->prev = cur; e->first_steer = cur_steer; e->cum = 0.0; e->bt_steps = 0;
        for (int i = 0; i < 5; i++) e->hist[i] = cur;
        e->stage[0] = e->stage[1] = e->stage[2] = 0;
        e->closest = cur;
        e->has_rs = 1;
    } else if (cur < e->closest) e->closest = cur;                  /* :70-74 */

    /* ---- compute_dynamic_weights :189-238 ---- */
    double jp = clipd((d0 - cur) / d0, 0.0, 1.0);
    double w_final = (tanh(7.0 * (jp - 0.3)) + 1) / 2.0, w_head = 1.0 - w_final;
    /* ---- calculate_exponential_distance_reward :126-142 (weight 0) ---- */
    double dist_r = exp(-2.0 * fmin(cur / c->max_expected_distance, 1.0));
    /* ---- calculate_progress_reward :144-187 ---- */
    double inst = e->prev - cur;
    double prog = inst > 0 ? tanh(inst / 1.0) : tanh(inst / 1.0) * 0.5;
    for (int i = 0; i < 4; i++) e->hist[i] = e->hist[i + 1];
    e->hist[4] = cur;
    double net = e->hist[0] - cur;
    double netr = tanh(net / 2.0) * 0.5;
    double eff = (e->hist[2] >= e->hist[3] && e->hist[3] >= e->hist[4]) ? 0.2 : 0.0;
    double progress = prog + netr + eff;
    /* ---- calculate_orientation_alignment_reward :285-309 ---- */
    double desired = atan2(e->gy - st[5], e->gx - st[4]);
    float cur_ori = atan2f(obs[6], obs[7]);
    double adiff = desired - ((double)cur_ori + M_PI);
    adiff = atan2(sin(adiff), cos(adiff));
    double heading = cos(adiff);
    /* ---- calculate_trailer_goal_orientation_reward :311-324 ---- */
    double orient = (double)obs[20];
    /* ---- calculate_staged_success_rewards :338-367 ---- */
    double oerr = (double)fabsf(atan2f(obs[19], obs[20]));
    double staged = 0;
    if (cur <= 5.0) { staged += 10; e->stage[0] = 1; }
    if (cur <= 2.0 && oerr <= 45.0 * M_PI / 180.0 && !e->stage[1]) { staged += 25; e->stage[1] = 1; }
    if (cur <= c->pos_thr && oerr <= c->ori_thr && !e->stage[2]) { staged += 100; e->stage[2] = 1; }
    /* ---- calculate_safety_penalties :369-421 ---- */
    double saf = 0; int vt = V_NONE;
    double th = fabs(st[0] - st[1]);
    if (th > 85.0 * M_PI / 180.0) { saf += -500.0; vt = V_JACKKNIFE; }
    else if (th > 70.0 * M_PI / 180.0) { saf += -50.0; vt = V_JACKKNIFE_WARNING; }
    double lo = c->map_min, hi = c->map_max;
    if (st[2] < lo - 2 || st[2] > hi + 2 || st[3] < lo - 2 || st[3] > hi + 2 ||
        st[4] < lo - 2 || st[4] > hi + 2 || st[5] < lo - 2 || st[5] > hi + 2) { saf += -500.0; vt = V_MAJOR_BOUNDARY; }
    else if (st[2] < lo || st[2] > hi || st[3] < lo || st[3] > hi ||
             st[4] < lo || st[4] > hi || st[5] < lo || st[5] > hi) { saf += -50.0; vt = V_MINOR_BOUNDARY; }
    if (e->gy > st[5]) { saf += -500.0; vt = V_PAST_GOAL; }
    if (episode_steps >= rmax) { saf += -500.0; vt = V_MAX_STEP; }
    int exb = cur > e->closest + 6.0;                               /* :120-124 */
    if (exb) { saf += -500.0; vt = V_EXCESSIVE_BACKWARD; }
    /* ---- calculate_exploration_bonus :423-439 ---- */
    double expl = episode_steps < rmax * 0.5 ? 4.0 : (episode_steps < rmax * 0.8 ? 2.0 : 0.0);
    /* ---- calculate_backward_movement_penalty :240-283 ---- */
    e->cum += fmax(0.0, cur - e->prev);
    e->bt_steps += 1;
    double budget = 5.0 * fmin(1.0, e->bt_steps / 50.0);
    double excess = fmax(0.0, e->cum - budget);
    double back = excess > 0 ? -(pow(excess, 1.5) * 0.5) : 0.0;
    /* ---- calculate_steering_smoothness_penalty :326-335 ---- */
    float steer_change = cur_steer - e->first_steer;                /* float32 - float32 */
    double smooth = (double)fabsf(steer_change) / (90.0 * M_PI / 180.0);
    /* ---- compute_reward :462-486 ---- */
    int success = cur <= c->pos_thr && oerr <= c->ori_thr;
    double fin = success ? 200.0 : 0.0;
    e->prev = cur;                                                  /* :472 */
    double c_dist = dist_r * 0.0 * 1.0, c_prog = progress * 15.0 * 1.0, c_head = heading * 15.0 * w_head,
           c_ori = orient * 15.0 * w_final, c_back = back * 1.0, c_smooth = smooth * -25.0;
    double total = c_dist + c_prog + c_head + c_ori + staged + saf + expl + c_back + c_smooth + fin;
    if (comps) {
        comps[C_TOTAL] = total; comps[C_DISTANCE] = c_dist; comps[C_PROGRESS] = c_prog; comps[C_HEADING] = c_head;
        comps[C_ORIENT] = c_ori; comps[C_STAGED] = staged; comps[C_SAFETY] = saf; comps[C_EXPLORE] = expl;
        comps[C_FINAL] = fin; comps[C_BACKWARD] = c_back; comps[C_SMOOTH] = c_smooth;
    }
    if (viol) *viol = vt;
    if (success_out) *success_out = success;

    /* ---- termination, simv2.py:528-541 ---- */
    int fl[F_NFLAGS];
    fl[F_JACKKNIFE] = fabs(st[0] - st[1]) > c->max_hitch;           /* :305-310 (deg2rad(90)) */
    fl[F_OUT_OF_MAP] = st[2] < lo || st[2] > hi || st[3] < lo || st[3] > hi ||
                       st[4] < lo || st[4] > hi || st[5] < lo || st[5] > hi;   /* :315-326 */
    fl[F_MAX_STEPS] = e->steps >= e->emax;                          /* :312-313 */
    fl[F_GOAL_PASSED] = e->gy > st[5];                              /* :341-345 */
    fl[F_GOAL_REACHED] = cur <= c->pos_thr && oerr <= c->ori_thr;   /* :533-536 */
    fl[F_EXCESSIVE_BACKWARD] = exb;                                 /* :538 */
    int done = 0;
    for (int i = 0; i < F_NFLAGS; i++) { done |= fl[i]; if (flags) flags[i] = fl[i]; }
    if (margins) {
        double out = -1e30;
        for (int i = 2; i < 6; i++) { out = fmax(out, lo - st[i]); out = fmax(out, st[i] - hi); }
        margins[M_JACKKNIFE] = fabs(st[0] - st[1]) - c->max_hitch; margins[M_MAP] = out; margins[M_PAST] = e->gy - st[5];
        margins[M_BACKWARD] = cur - (e->closest + 6.0); margins[M_GOAL_POS] = c->pos_thr - cur;
        margins[M_GOAL_ORI] = c->ori_thr - oerr; margins[M_STEPS] = (double)(e->steps - e->emax);
    }
    return done;
}

TTO_API int tto_step(const tto_cfg *c, tto_env *e, float action, float *obs, double *comps, int *viol,
                     int *flags, int *success_out) {
    return tto_step_m(c, e, action, obs, comps, viol, flags, success_out, 0);
}

/* Replays T actions from the env's current state; arrays are [T,...] row-major; stops after `done`
 * (later rows untouched).  Returns the number of steps executed. */
TTO_API int tto_replay(const tto_cfg *c, tto_env *e, const float *actions, int T, double *state, float *obs,
                       double *comps, uint8_t *viol, uint8_t *flags, uint8_t *done, uint8_t *success) {
    int n = 0;
    for (int t = 0; t < T; t++) {
        float o[23]; double cp[C_NCOMP]; int v, fl[F_NFLAGS], su;
        int d = tto_step(c, e, actions[t], o, cp, &v, fl, &su);
        if (state) memcpy(state + 6 * t, e->st, 6 * sizeof(double));
        if (obs) memcpy(obs + 23 * t, o, sizeof o);
        if (comps) memcpy(comps + C_NCOMP * t, cp, sizeof cp);
        if (viol) viol[t] = (uint8_t)v;
        if (flags) for (int i = 0; i < F_NFLAGS; i++) flags[F_NFLAGS * t + i] = (uint8_t)fl[i];
        if (done) done[t] = (uint8_t)d;
        if (success) success[t] = (uint8_t)su;
        n = t + 1;
        if (d) break;
    }
    return n;
}

/* tto_replay plus the per-step threshold margins [T, M_NMARGINS] (see tto_step_m) */
TTO_API int tto_replay_m(const tto_cfg *c, tto_env *e, const float *actions, int T, double *state, float *obs,
                         double *comps, uint8_t *flags, uint8_t *done, double *margins) {
    int n = 0;
    for (int t = 0; t < T; t++) {
        float o[23]; double cp[C_NCOMP]; int v, fl[F_NFLAGS], su;
        int d = tto_step_m(c, e, actions[t], o, cp, &v, fl, &su, margins ? margins + M_NMARGINS * t : 0);
        if (state) memcpy(state + 6 * t, e->st, 6 * sizeof(double));
        if (obs) memcpy(obs + 23 * t, o, sizeof o);
        if (comps) memcpy(comps + C_NCOMP * t, cp, sizeof cp);
        if (flags) for (int i = 0; i < F_NFLAGS; i++) flags[F_NFLAGS * t + i] = (uint8_t)fl[i];
        if (done) done[t] = (uint8_t)d;
        n = t + 1;
        if (d) break;
    }
    return n;
}

TTO_API int tto_sizeof_env(void) { return (int)sizeof(tto_env); }

/* ------------------------------------------------------------------ actor (networks.py:138-147) --------- */
/* y = tanh(W3 . relu(LN2(W2 . relu(LN1(W1 s + b1)) + b2)) + b3); LayerNorm eps 1e-5, biased variance (torch).
 * Weights in the reference state_dict layout: fc1.weight[H1,IN] fc1.bias[H1] bn1.weight[H1] bn1.bias[H1]
 * fc2.weight[H2,H1] fc2.bias[H2] bn2.weight[H2] bn2.bias[H2] mu.weight[1,H2] mu.bias[1].  float64 math. */
typedef struct {
    int in_dim, h1, h2;
    const float *w1, *b1, *g1, *be1, *w2, *b2, *g2, *be2, *w3, *b3;
} tto_actor;

static void layer_norm_relu(double *x, int n, const float *g, const float *b) {
    double m = 0, v = 0;
    for (int i = 0; i < n; i++) m += x[i];
    m /= n;
    for (int i = 0; i < n; i++) v += (x[i] - m) * (x[i] - m);
    v /= n;
    double r = 1.0 / sqrt(v + 1e-5);
    for (int i = 0; i < n; i++) { double y = (x[i] - m) * r * g[i] + b[i]; x[i] = y > 0 ? y : 0; }
}

TTO_API void tto_actor_forward(const tto_actor *a, const float *obs, int ld_obs, int B, float *out) {
    {
        double *h1 = malloc(sizeof(double) * a->h1), *h2 = malloc(sizeof(double) * a->h2);
        for (int n = 0; n < B; n++) {
            const float *s = obs + (size_t)n * ld_obs;
            for (int j = 0; j < a->h1; j++) {
                double acc = a->b1[j];
                for (int k = 0; k < a->in_dim; k++) acc += (double)a->w1[j * a->in_dim + k] * s[k];
                h1[j] = acc;
            }
            layer_norm_relu(h1, a->h1, a->g1, a->be1);
            for (int j = 0; j < a->h2; j++) {
                double acc = a->b2[j];
                for (int k = 0; k < a->h1; k++) acc += (double)a->w2[j * a->h1 + k] * h1[k];
                h2[j] = acc;
            }
            layer_norm_relu(h2, a->h2, a->g2, a->be2);
            double acc = a->b3[0];
            for (int k = 0; k < a->h2; k++) acc += (double)a->w3[k] * h2[k];
            out[n] = (float)tanh(acc);
        }
        free(h1); free(h2);
    }
}

/* float32 variant used by the CPU-baseline port (what torch CPU fp32 does, modulo summation order) */
static void layer_norm_relu_f(float *x, int n, const float *g, const float *b) {
    float m = 0, v = 0;
    for (int i = 0; i < n; i++) m += x[i];
    m /= n;
    for (int i = 0; i < n; i++) v += (x[i] - m) * (x[i] - m);
    v /= n;
    float r = 1.0f / sqrtf(v + 1e-5f);
    for (int i = 0; i < n; i++) { float y = (x[i] - m) * r * g[i] + b[i]; x[i] = y > 0 ? y : 0; }
}

static float actor_forward_one_f(const tto_actor *a, const float *s, float *h1, float *h2) {
    for (int j = 0; j < a->h1; j++) {
        float acc = 0.f;
        const float *w = a->w1 + j * a->in_dim;
#pragma omp simd reduction(+ : acc)
        for (int k = 0; k < a->in_dim; k++) acc += w[k] * s[k];
        h1[j] = acc + a->b1[j];
    }
    layer_norm_relu_f(h1, a->h1, a->g1, a->be1);
    for (int j = 0; j < a->h2; j++) {
        float acc = 0.f;
        const float *w = a->w2 + j * a->h1;
#pragma omp simd reduction(+ : acc)
        for (int k = 0; k < a->h1; k++) acc += w[k] * h1[k];
        h2[j] = acc + a->b2[j];
    }
    layer_norm_relu_f(h2, a->h2, a->g2, a->be2);
    float acc = a->b3[0];
    for (int k = 0; k < a->h2; k++) acc += a->w3[k] * h2[k];
    return tanhf(acc);
}

/* ------------------------------------------------------------------ RNG spec (new; no reference analogue) */
/* Philox4x32-10 (Salmon et al., SC'11).  The reference uses the global MT19937 (simv2.py:460-462,
 * noise.py:14); bit parity with it is neither possible nor required.  The spec the
 * CUDA kernels and this oracle share:
 *   key     = (seed_lo, seed_hi)
 *   counter = (global_env_id, iteration t, stream, 0)   stream 0 = reset pose, 1 = OU noise
 *   pose:   u_k = (w_k + 0.5) * 2^-32 (double);  start_x = -27 + 54 u0;  start_y = 27 u1;
 *           yaw = deg2rad(45) + (deg2rad(120) - deg2rad(45)) u2          (simv2.py:331-333)
 *   normal: u1 = ((w0 >> 8) + 1) * 2^-24, u2 = (w1 >> 8) * 2^-24 (float);  n = sqrt(-2 ln u1) cos(2 pi u2)
 */
TTO_API void tto_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

TTO_API void tto_rng_pose(uint64_t seed, uint32_t gid, uint32_t t, double *sx, double *sy, double *syaw) {
    uint32_t ctr[4] = {gid, t, 0u, 0u}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
    tto_philox4x32_10(ctr, key, w);
    const double s = 1.0 / 4294967296.0;
    double u0 = ((double)w[0] + 0.5) * s, u1 = ((double)w[1] + 0.5) * s, u2 = ((double)w[2] + 0.5) * s;
    const double a45 = 45.0 * M_PI / 180.0, a120 = 120.0 * M_PI / 180.0;
    *sx = fma(54.0, u0, -27.0); *sy = 27.0 * u1; *syaw = fma(a120 - a45, u2, a45);
}

TTO_API float tto_rng_normal(uint64_t seed, uint32_t gid, uint32_t t) {
    uint32_t ctr[4] = {gid, t, 1u, 0u}, key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)}, w[4];
    tto_philox4x32_10(ctr, key, w);
    float u1 = (float)((w[0] >> 8) + 1u) * (1.0f / 16777216.0f), u2 = (float)(w[1] >> 8) * (1.0f / 16777216.0f);
    return sqrtf(-2.0f * logf(u1)) * cosf(6.283185307179586f * u2);
}

/* noise.py:12-17: x <- x + theta (mu - x) dt + sigma sqrt(dt) N(0,1); theta .2 sigma .15 dt 1e-2 mu 0 */
TTO_API void tto_ou_step(float *x, float *action, const uint8_t *reset_mask, int N, uint64_t seed,
                         uint32_t gid0, uint32_t t) {
    for (int i = 0; i < N; i++) {
        float xp = (reset_mask && reset_mask[i]) ? 0.0f : x[i];      /* trainv2.py:492 noise.reset() */
        float n = tto_rng_normal(seed, gid0 + (uint32_t)i, t);
        float xn = xp + 0.2f * (0.0f - xp) * 0.01f + 0.15f * 0.1f * n;
        x[i] = xn;
        if (action) action[i] += xn;                                 /* DDPG_agent.py:41-43 */
    }
}

/* Open-loop replay of a recorded rollout (tests): envs [0, n) of a shard with global ids gid0 + i, T iterations of the
 * training-loop body (DDPG/trainv2.py:489-531) driven by the RAW actions a[T, n] that the rollout stored (what
 * agent.remember keeps): scaled = clip(a, -1, 1) * float32(pi / 4) (trainv2.py:516), env.step, and on `done` a reset from
 * the env's Philox pose stream at that iteration (iter0 + t) plus noise.reset().  Outputs are [T, n, ...] row-major like the
 * ring: s (observation before the step), s2 (after, terminal when done), reward, done, termination bits, state after the
 * step, threshold margins, and the OU state x that choose_action added at that iteration.  Initial poses: the full
 * reset's salted counter `pose_t0`. */
TTO_API void tto_rollout_replay(const tto_cfg *c, uint64_t seed, uint32_t gid0, uint32_t pose_t0, uint32_t iter0, int64_t n, int T,
                                const float *a, float *s, float *s2, double *rew, uint8_t *done, uint8_t *flags,
                                double *state, double *margins, float *ou) {
    for (int64_t i = 0; i < n; i++) {
        tto_env e; float o[23]; double sx, sy, syaw; float x = 0.0f;
        const uint32_t gid = gid0 + (uint32_t)i;
        tto_rng_pose(seed, gid, pose_t0, &sx, &sy, &syaw);
        tto_reset_pose(c, &e, sx, sy, syaw, 0.0, -30.0, 90.0 * M_PI / 180.0, o);
        for (int t = 0; t < T; t++) {
            const int64_t r = (int64_t)t * n + i;
            x = x + 0.2f * (0.0f - x) * 0.01f + 0.15f * 0.1f * tto_rng_normal(seed, gid, iter0 + (uint32_t)t);
            if (ou) ou[r] = x;
            if (s) memcpy(s + r * 23, o, sizeof o);
            const float scaled = fminf(fmaxf(a[r], -1.0f), 1.0f) * 0.78539819f;
            float o2[23]; double cp[C_NCOMP]; int v, fl[F_NFLAGS], su;
            const int d = tto_step_m(c, &e, scaled, o2, cp, &v, fl, &su, margins ? margins + r * M_NMARGINS : 0);
            if (s2) memcpy(s2 + r * 23, o2, sizeof o2);
            if (rew) rew[r] = cp[C_TOTAL];
            if (done) done[r] = (uint8_t)d;
            if (flags) { uint8_t b = 0; for (int k = 0; k < F_NFLAGS; k++) b |= (uint8_t)(fl[k] << k); flags[r] = b; }
            if (state) memcpy(state + r * 6, e.st, sizeof e.st);
            if (d) {
                tto_rng_pose(seed, gid, iter0 + (uint32_t)t, &sx, &sy, &syaw);
                tto_reset_pose(c, &e, sx, sy, syaw, 0.0, -30.0, 90.0 * M_PI / 180.0, o);
                x = 0.0f;
            } else memcpy(o, o2, sizeof o);
        }
    }
}

/* ------------------------------------------------------------------ replay (replay_buffer.py:13-21) ----- */
/* N sequential store_transition calls in env order: index (cntr+i) % cap, last writer wins. float32 rows. */
TTO_API void tto_replay_store(float *S, float *A, float *R, float *S2, uint8_t *D, int64_t cap, int64_t cntr,
                              const float *s, int ld_s, const float *a, const float *r, const float *s2, int ld_s2,
                              const uint8_t *d, int64_t N) {
    for (int64_t i = 0; i < N; i++) {
        int64_t idx = (cntr + i) % cap;
        memcpy(S + idx * 23, s + i * ld_s, 23 * sizeof(float));
        A[idx] = a[i]; R[idx] = r[i];
        memcpy(S2 + idx * 23, s2 + i * ld_s2, 23 * sizeof(float));
        D[idx] = d[i];
    }
}

/* ------------------------------------------------------------------ CPU-baseline port ------------------- */
/* One rollout iteration over envs [i0, i1) (the caller runs one slice per host thread; ctypes releases
 * the GIL): choose_action (actor fp32 + OU noise,
 * DDPG_agent.py:36-49) -> clip * pi/4 (trainv2.py:516) -> env.step (simv2.py:499-545) -> store_transition
 * (replay_buffer.py:13-21) -> reset on done (trainv2.py:489-492).  Used only by bench.py's cpu_baseline /
 * --impl reference legs (kind "port").  Returns the number of episodes finished. */
TTO_API int64_t tto_rollout_port(const tto_cfg *c, tto_env *envs, float *obs /*[N,23] in/out*/, float *ou,
                                 const tto_actor *a, int64_t i0, int64_t i1, uint64_t seed, uint32_t gid0, uint32_t t,
                                 float *S, float *A, float *R, float *S2, uint8_t *D, int64_t cap, int64_t cntr,
                                 double *reward_sum) {
    int64_t finished = 0; double rsum = 0;
    {
        float *h1 = malloc(sizeof(float) * a->h1), *h2 = malloc(sizeof(float) * a->h2);
        for (int64_t i = i0; i < i1; i++) {
            float *o = obs + i * 23;
            float mu = actor_forward_one_f(a, o, h1, h2);
            float xn = ou[i] + 0.2f * (0.0f - ou[i]) * 0.01f + 0.15f * 0.1f * tto_rng_normal(seed, gid0 + (uint32_t)i, t);
            ou[i] = xn;
            float act = mu + xn;
            float scaled = fminf(fmaxf(act, -1.0f), 1.0f) * 0.78539819f;
            float o2[23]; double cp[C_NCOMP]; int v, fl[F_NFLAGS], su;
            int d = tto_step(c, &envs[i], scaled, o2, cp, &v, fl, &su);
            int64_t idx = (cntr + i) % cap;
            memcpy(S + idx * 23, o, sizeof o2); A[idx] = act; R[idx] = (float)cp[C_TOTAL];
            memcpy(S2 + idx * 23, o2, sizeof o2); D[idx] = (uint8_t)d;
            rsum += cp[C_TOTAL];
            if (d) {
                double sx, sy, syaw;
                tto_rng_pose(seed, gid0 + (uint32_t)i, t, &sx, &sy, &syaw);
                tto_reset_pose(c, &envs[i], sx, sy, syaw, 0.0, -30.0, 90.0 * M_PI / 180.0, o);
                ou[i] = 0.0f; finished++;
            } else memcpy(o, o2, sizeof o2);
        }
        free(h1); free(h2);
    }
    if (reward_sum) *reward_sum = rsum;
    return finished;
}

