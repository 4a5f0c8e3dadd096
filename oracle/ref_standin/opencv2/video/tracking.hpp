#pragma once
// stand-in: forwards to the minimal cv:: subset (cv_standin.h), which declares calcOpticalFlowPyrLK
#include "../../cv_standin.h"
