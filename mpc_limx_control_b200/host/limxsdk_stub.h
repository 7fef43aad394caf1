// limxsdk_stub.h -- POD stand-ins for the limX SDK types the controller signatures use
// (limxsdk/datatypes.h is a vendor binary dependency, reference CMakeLists.txt:55-57).  Field lists
// follow the uses in the reference: RobotState q/dq/tau (include/stateEstimator.h:57-64),
// RobotCmd mode/q/dq/tau/Kp/Kd (src/pf_controller_base.cpp:47-51), ImuData acc/gyro/quat.
#pragma once
#include <cstdint>
#include <vector>

namespace limxsdk {
struct RobotState {
    uint64_t stamp = 0;
    std::vector<float> tau = std::vector<float>(6, 0.f);
    std::vector<float> q = std::vector<float>(6, 0.f);
    std::vector<float> dq = std::vector<float>(6, 0.f);
};
struct RobotCmd {
    uint64_t stamp = 0;
    std::vector<uint8_t> mode = std::vector<uint8_t>(6, 0);
    std::vector<float> q = std::vector<float>(6, 0.f);
    std::vector<float> dq = std::vector<float>(6, 0.f);
    std::vector<float> tau = std::vector<float>(6, 0.f);
    std::vector<float> Kp = std::vector<float>(6, 0.f);
    std::vector<float> Kd = std::vector<float>(6, 0.f);
};
struct ImuData {
    uint64_t stamp = 0;
    float acc[3] = {0, 0, 0};
    float gyro[3] = {0, 0, 0};
    float quat[4] = {1, 0, 0, 0};
};
}  // namespace limxsdk
