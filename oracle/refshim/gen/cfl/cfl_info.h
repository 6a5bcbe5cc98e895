/* stand-in for lib/cfl/include/cfl/cfl_info.h.in */
#ifndef CFL_INFO_H
#define CFL_INFO_H
#define CFL_SOURCE_DIR "/root/reference/lib/cfl"
#define CFL_HAVE_TIMESPEC_GET
#define CFL_HAVE_GMTIME_R
#endif
