/* stand-in for lib/cprofiles/include/cprofiles/cprof_info.h.in */
#ifndef CPROF_INFO_H
#define CPROF_INFO_H
#define CPROF_HAVE_TIMESPEC_GET
#define CPROF_HAVE_GMTIME_R
#define CPROF_HAVE_CFL
#endif
