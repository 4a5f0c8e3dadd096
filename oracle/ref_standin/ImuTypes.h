// stand-in for include/ImuTypes.h (inertial path: dead code on the video sensors): only the type names
// include/Frame.h mentions.
#pragma once
namespace MOV_SLAM {
namespace IMU {
class Bias {};
class Calib {};
class Preintegrated {};
}  // namespace IMU
}  // namespace MOV_SLAM
