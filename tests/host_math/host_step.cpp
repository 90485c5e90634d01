// TEST INFRASTRUCTURE ONLY: host (g++) build of the kernel arithmetic in csrc/tt_env_math.cuh, so that
// its numerical error against the float64 oracle can be measured on the CPU (tests/test_host_math.py).
// The product never loads this; it only ever runs the CUDA build of the same header.
#include "../../ddpg-trucktrailer_b200/csrc/tt_env_math.cuh"
#include "../../ddpg-trucktrailer_b200/csrc/tt_consts.h"
#include <string.h>

extern "C" {

// replay T actions from an injected state; outputs [T,...]; stops after done; returns steps executed
int hm_replay(const double *state0, const double *start, const double *goal, const float *actions, int T,
              double *state, float *obs, float *comps /*[T,11] total first*/, unsigned char *viol,
              unsigned char *flags, unsigned char *done, unsigned char *success, float *obs0) {
    tt_env_cfg cfg; tt_fill_default_cfg(&cfg);
    ttm::StepConsts k = tt_make_consts(cfg);
    ttm::EnvRegs e;
    ttm::use_default_l2(k, e);
    e.psi1 = state0[0]; e.psi2 = state0[1];
    e.x1 = ttm::pos_from_double(state0[2]); e.y1 = ttm::pos_from_double(state0[3]);
    e.x2 = ttm::pos_from_double(state0[4]); e.y2 = ttm::pos_from_double(state0[5]);
    ttm::begin_episode(k, e, start[0], start[1], goal[0], goal[1], goal[2], obs0);
    int n = 0;
    for (int t = 0; t < T; t++) {
        ttm::StepOut o;
        ttm::env_step<true>(k, e, actions[t], o);
        double s[6] = {e.psi1, e.psi2, ttm::pos_to_double(e.x1), ttm::pos_to_double(e.y1), ttm::pos_to_double(e.x2), ttm::pos_to_double(e.y2)};
        memcpy(state + 6 * t, s, sizeof s);
        memcpy(obs + 23 * t, o.obs, sizeof o.obs);
        comps[11 * t] = o.reward;
        memcpy(comps + 11 * t + 1, o.comps, sizeof o.comps);
        viol[t] = (unsigned char)o.viol; flags[t] = (unsigned char)o.flags; done[t] = o.done; success[t] = o.success;
        n = t + 1;
        if (o.done) break;
    }
    return n;
}

int hm_max_steps(const double *start, const double *goal) {
    tt_env_cfg cfg; tt_fill_default_cfg(&cfg);
    ttm::StepConsts k = tt_make_consts(cfg);
    double dx = goal[0] - start[0], dy = goal[1] - start[1];
    return (int)((ttm::pack_limits(k, sqrt(dx * dx + dy * dy)) >> ttm::PK_EMAX_SHIFT) & ttm::PK_EMAX_MASK);
}

void hm_reset_pose(double sx, double sy, double syaw, double *state, float *obs) {
    tt_env_cfg cfg; tt_fill_default_cfg(&cfg);
    ttm::StepConsts k = tt_make_consts(cfg);
    ttm::EnvRegs e;
    ttm::use_default_l2(k, e);
    ttm::reset_from_pose(k, e, sx, sy, syaw, cfg.goal_x, cfg.goal_y, cfg.goal_yaw, obs);
    double s[6] = {e.psi1, e.psi2, ttm::pos_to_double(e.x1), ttm::pos_to_double(e.y1), ttm::pos_to_double(e.x2), ttm::pos_to_double(e.y2)};
    memcpy(state, s, sizeof s);
}

void hm_rng_pose(unsigned long long seed, unsigned gid, unsigned t, double *out3) {
    tt_env_cfg cfg; tt_fill_default_cfg(&cfg);
    ttm::StepConsts k = tt_make_consts(cfg);
    ttm::rng_pose(k, seed, gid, t, out3[0], out3[1], out3[2]);
}
float hm_rng_normal(unsigned long long seed, unsigned gid, unsigned t) { return ttm::rng_normal(seed, gid, t); }
void hm_sincos_f64(double x, double *s, double *c) { ttm::sincos_f64(x, *s, *c); }
void hm_sincos_f32(double x, float *s, float *c) { ttm::sincos_f32_of_f64(x, *s, *c); }
}
