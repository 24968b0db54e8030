// advection.hpp — forwards to csim_dropin.hpp, which declares the reference's include/advection.hpp interface
// on top of the B200 C ABI (see that file's header for the file:line map).
#pragma once
#include "csim_dropin.hpp"
