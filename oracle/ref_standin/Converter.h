// stand-in for include/Converter.h: included by include/Frame.h, nothing of it is used on the front-end path
#pragma once
