// MPCParam.h -- configuration surface of the hot path (reference include/MPCParam.h:13-86):
// kinematic offsets, gait clock, nominal foot offsets.  Values are the reference's literals.
#pragma once
#include <array>

namespace mpcb200 {
namespace host {

struct kinematicValues {   // reference include/MPCParam.h:13-38
    double abad_offset_x = 0.05556, abad_offset_y = 0.105, abad_offset_z = -0.2602;
    double hip_offset_x = -0.077, hip_offset_y = 0.02050, hip_offset_z = 0.0;
    double knee_offset_x = -0.1500, knee_offset_y = -0.02050, knee_offset_z = -0.25981;
    double foot_offset_x = 0.145, foot_offset_y = 0.0, foot_offset_z = -0.2598;
    double contact_offset_x = 0.0, contact_offset_y = 0.0, contact_offset_z = -0.032;
};

class MPCParam {
public:
    MPCParam() {   // reference include/MPCParam.h:64-73
        const kinematicValues& k = KinematicValues;
        const double x = k.abad_offset_x + k.hip_offset_x + k.knee_offset_x + k.foot_offset_x + k.contact_offset_x;
        const double z = k.abad_offset_z + k.hip_offset_z + k.knee_offset_z + k.foot_offset_z + k.contact_offset_z;
        static_foot_offset_left = {x, -k.abad_offset_y - k.hip_offset_y - k.knee_offset_y + k.foot_offset_y + k.contact_offset_y, z};
        static_foot_offset_right = {x, k.abad_offset_y + k.hip_offset_y + k.knee_offset_y + k.foot_offset_y + k.contact_offset_y, z};
    }
    float dt = 0.001f;                                          // :44
    int milliseconds_per_step = static_cast<int>(1 / dt);       // :45 (evaluates to 999)
    int mpcStep = 5;                                            // :46
    float dtMPC = dt * mpcStep;                                 // :47
    float swing_time = 0.5f;                                    // :48
    float stance_time = 0.5f;                                   // :49
    float gait_height = 0.1f;                                   // :51
    float givenErrorRate = 0.1f;                                // :53
    kinematicValues KinematicValues;
    std::array<double, 3> static_foot_offset_right;
    std::array<double, 3> static_foot_offset_left;
};

}  // namespace host
}  // namespace mpcb200
