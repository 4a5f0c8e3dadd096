/* forwards to the fake libav declarations shared with the shim tests (mov-slam_b200/shim/standin/libav_standin.h) */
#pragma once
#include "../../mov-slam_b200/shim/standin/libav_standin.h"
