// forwarding header: callers inside this repository include "grgsm_vitac.h"; the reference's callers include
// <grgsm_vitac/grgsm_vitac.h> (utils/va-test/burst-gen.cpp:46, ms/ms_upper.cpp)
#pragma once
#include "grgsm_vitac/grgsm_vitac.h"
