// tt_consts.h -- host-side derivation of the kernel constants from tt_env_cfg (reference literals:
// truck_trailer_sim/simv2.py:23-101, :331-337; reward_functionv1.py:381-387).
#pragma once
#include <math.h>
#include "../../include/tt_b200.h"
#include "tt_env_math.cuh"

static inline void tt_fill_default_cfg(tt_env_cfg *c) {
    const double deg = 3.14159265358979323846 / 180.0;
    c->L1 = 5.0; c->L2 = 7.0; c->v1x = -5.012; c->dt = 0.08;
    c->map_min = -40.0; c->map_max = 40.0;
    c->max_hitch = 90.0 * deg; c->steer_max = 45.0 * deg;
    c->pos_thr = 0.5; c->ori_thr = 15.0 * deg;
    c->step_len = 0.40096;
    c->start_x_lo = -27.0; c->start_x_hi = 27.0; c->start_y_lo = 0.0; c->start_y_hi = 27.0;
    c->start_yaw_lo = 45.0 * deg; c->start_yaw_hi = 120.0 * deg;
    c->goal_x = 0.0; c->goal_y = -30.0; c->goal_yaw = 90.0 * deg;
}

static inline ttm::StepConsts tt_make_consts(const tt_env_cfg &c) {
    const double deg = 3.14159265358979323846 / 180.0;
    ttm::StepConsts k;
    k.vL1 = c.v1x / c.L1; k.vL2 = c.v1x / c.L2; k.h = c.dt;
    k.steer_max = c.steer_max; k.max_hitch = c.max_hitch;
    k.jk_major = 85.0 * deg; k.jk_minor = 70.0 * deg;
    k.map_min = c.map_min; k.map_max = c.map_max; k.step_len = c.step_len; k.L2 = c.L2;
    k.sx_lo = c.start_x_lo; k.sx_w = c.start_x_hi - c.start_x_lo;
    k.sy_lo = c.start_y_lo; k.sy_w = c.start_y_hi - c.start_y_lo;
    k.syaw_lo = c.start_yaw_lo; k.syaw_w = c.start_yaw_hi - c.start_yaw_lo;
    k.gx = c.goal_x; k.gy = c.goal_y; k.gyaw = c.goal_yaw;
    k.hv_fix = (float)(c.dt * c.v1x * ttm::kPosScale);
    k.pos_inv = (float)(1.0 / ttm::kPosScale); k.pos_scale_d = ttm::kPosScale;
    k.map_min_fix = ttm::pos_from_double(c.map_min); k.map_max_fix = ttm::pos_from_double(c.map_max);
    k.maj_lo_fix = ttm::pos_from_double(c.map_min - 2.0); k.maj_hi_fix = ttm::pos_from_double(c.map_max + 2.0);
    k.gx_fix = ttm::pos_from_double(c.goal_x); k.gy_fix = ttm::pos_from_double(c.goal_y);
    k.sgy0 = (float)sin(c.goal_yaw); k.cgy0 = (float)cos(c.goal_yaw);
    const double w = c.map_max - c.map_min;
    k.mid = (float)((c.map_max + c.map_min) / 2.0); k.inv_half = (float)(2.0 / w);
    k.inv_maxd = (float)(1.0 / sqrt(w * w + w * w));
    k.pos_thr = (float)c.pos_thr;
    k.cos_ori_thr = (float)cos(c.ori_thr); k.cos_45 = (float)cos(45.0 * deg);
    return k;
}
