// tron1_params.h -- public params (include/mpc_b200.h) -> derived device constants.
#pragma once
#include <math.h>
#include <string.h>

#include "../../include/mpc_b200.h"
#include "tron1_core.cuh"

namespace mpcb200 {

inline void tron1_default_params(mpc_b200_tron1_params& p) {
    // constants: reference include/mpcQP.h:18-22,54-56 ; include/MPCParam.h:44-49
    static const double q[13] = {1, 1, 10, 100, 100, 100, 50, 50, 50, 100, 100, 100, 0.1};
    static const double I[9] = {140110.479E-06, 534.939E-06, 28184.116E-06,
                                534.939E-06, 110641.449E-06, -27.278E-06,
                                28184.116E-06, -27.278E-06, 98944.542E-06};
    memset(&p, 0, sizeof(p));
    p.Ts = 0.005;
    p.mass = 9.585;
    memcpy(p.inertia, I, sizeof(I));
    memcpy(p.q, q, sizeof(q));
    p.r = 0.1;
    p.p_scale = 20.0;
    p.mu = 0.5;
    p.f_max = 2.0 * 9.585 * 9.8;
    p.ltv = 1;
    p.per_step_feet = 0;
    p.gait_dt = 0.001f;
    p.gait_mpc_step = 5;
    p.gait_swing_time = 0.5f;
    p.gait_stance_time = 0.5f;
    p.max_newton = 12;
    p.max_admm = 2000;
    p.tol = 1e-9;
    // reference include/MPCParam.h:13-38,64-73
    const double ox = 0.05556 - 0.077 - 0.15 + 0.145 + 0.0, oz = -0.2602 + 0.0 - 0.25981 - 0.2598 - 0.032;
    p.foot_offset_left[0] = ox; p.foot_offset_left[1] = -0.105 - 0.0205 - (-0.0205) + 0.0 + 0.0; p.foot_offset_left[2] = oz;
    p.foot_offset_right[0] = ox; p.foot_offset_right[1] = 0.105 + 0.0205 + (-0.0205) + 0.0 + 0.0; p.foot_offset_right[2] = oz;
}

// returns non-zero on invalid parameters
inline int make_tron1_const(const mpc_b200_tron1_params& p, Tron1Const& c) {
    if (!(p.Ts > 0.0) || !(p.mass > 0.0) || !(p.r > 0.0) || !(p.mu > 0.0) || !(p.f_max > 0.0)) return 1;
    if (!(p.p_scale >= 0.0) || p.max_newton < 1 || p.max_admm < 0 || !(p.tol > 0.0)) return 1;
    for (int i = 0; i < 13; ++i) if (!(p.q[i] >= 0.0)) return 1;
    memset(&c, 0, sizeof(c));
    c.Ts = p.Ts;
    c.inv_m = 1.0 / p.mass;
    const double* M = p.inertia;  // symmetric: row/column-major agree
    double a = M[0], b = M[1], cc = M[2], d = M[3], e = M[4], f = M[5], g = M[6], h = M[7], i = M[8];
    double det = a * (e * i - f * h) - b * (d * i - f * g) + cc * (d * h - e * g);
    if (!(fabs(det) > 0.0)) return 1;
    double id = 1.0 / det;
    c.Iinv[0] = (e * i - f * h) * id; c.Iinv[1] = (cc * h - b * i) * id; c.Iinv[2] = (b * f - cc * e) * id;
    c.Iinv[3] = (f * g - d * i) * id; c.Iinv[4] = (a * i - cc * g) * id; c.Iinv[5] = (cc * d - a * f) * id;
    c.Iinv[6] = (d * h - e * g) * id; c.Iinv[7] = (b * g - a * h) * id; c.Iinv[8] = (a * e - b * d) * id;
    memcpy(c.q, p.q, sizeof(c.q));
    c.r = p.r;
    c.p_scale = p.p_scale;
    c.mu = p.mu;
    c.f_max = p.f_max;
    c.tol = p.tol;
    c.gamma = 0.5;
    c.admm_alpha = 1.6;
    c.ltv = p.ltv;
    c.per_step_feet = p.per_step_feet;
    c.max_newton = p.max_newton;
    c.max_admm = p.max_admm;
    c.gait_dt = p.gait_dt;
    c.gait_swing = p.gait_swing_time;
    c.gait_stance = p.gait_stance_time;
    c.gait_mpc_step = p.gait_mpc_step;
    {   // cycle = swing + stance as ONE float add (reference include/MPCController.h:63 with include/MPCParam.h:48-49)
        volatile float cyc = p.gait_swing_time + p.gait_stance_time;
        c.gait_cycle = (double)cyc;
        c.gait_inv_cycle = 1.0 / c.gait_cycle;
    }
    for (int k = 0; k < 3; ++k) { c.foot_off_l[k] = p.foot_offset_left[k]; c.foot_off_r[k] = p.foot_offset_right[k]; }
    return 0;
}

}  // namespace mpcb200
