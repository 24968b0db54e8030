// io.hpp — forwards to csim_driver.hpp, which declares the reference's include/io.hpp interface
// for this build (see that file's header for the file:line map).
#pragma once
#include "csim_driver.hpp"
