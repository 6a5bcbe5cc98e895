/* stand-in for lib/ctraces/include/ctraces/ctr_info.h.in */
#ifndef CTR_INFO_H
#define CTR_INFO_H
#define CTR_HAVE_TIMESPEC_GET
#define CTR_HAVE_GMTIME_R
#define CTR_HAVE_CFL
#endif
