/* stand-in for lib/cmetrics/include/cmetrics/cmt_version.h.in (lib/cmetrics/CMakeLists.txt:7-10) */
#ifndef CMT_VERSION_H
#define CMT_VERSION_H
#define CMT_VERSION_MAJOR 2
#define CMT_VERSION_MINOR 1
#define CMT_VERSION_PATCH 1
#define CMT_VERSION (CMT_VERSION_MAJOR * 10000 + CMT_VERSION_MINOR * 100 + CMT_VERSION_PATCH)
#define CMT_VERSION_STR "2.1.1"
#endif
