#pragma once
// stand-in: forwards to the minimal Eigen/Sophus subset (geom_standin.h)
#include "../geom_standin.h"
