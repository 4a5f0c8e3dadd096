#pragma once
// stand-in: forwards to the minimal cv:: subset this repo wrote for building the reference's front-end sources
#include "../../cv_standin.h"
