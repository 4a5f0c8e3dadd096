#pragma once
/* stand-in: forwards to the fake libav declarations (libav_standin.h) */
#include "../libav_standin.h"
