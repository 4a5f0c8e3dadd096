// stand-in for include/Settings.h: included by include/Frame.h, nothing of it is used on the front-end path
#pragma once
