/* stand-in for lib/cmetrics/include/cmetrics/cmt_info.h.in */
#ifndef CMT_INFO_H
#define CMT_INFO_H
#define CMT_SOURCE_DIR "/root/reference/lib/cmetrics"
#define CMT_HAVE_TIMESPEC_GET
#define CMT_HAVE_GMTIME_R
#define CMT_HAVE_CFL
#define CMT_HAVE_CFL_INTERNAL
#endif
