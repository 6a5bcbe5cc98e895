/* stand-in for lib/cfl/include/cfl/cfl_version.h.in (lib/cfl/CMakeLists.txt:7-10) */
#ifndef CFL_VERSION_H
#define CFL_VERSION_H
#define CFL_VERSION_MAJOR 0
#define CFL_VERSION_MINOR 6
#define CFL_VERSION_PATCH 1
#define CFL_VERSION (CFL_VERSION_MAJOR * 10000 + CFL_VERSION_MINOR * 100 + CFL_VERSION_PATCH)
#define CFL_VERSION_STR "0.6.1"
#endif
