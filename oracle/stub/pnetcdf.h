/* oracle/stub/pnetcdf.h — empty stand-in so the reference's src/init.cpp (which includes
 * <pnetcdf.h> at init.cpp:3 but calls nothing from it) compiles unmodified.  TEST INFRASTRUCTURE. */
#pragma once
