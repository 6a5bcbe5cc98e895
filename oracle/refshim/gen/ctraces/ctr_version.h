/* stand-in for lib/ctraces/include/ctraces/ctr_version.h.in */
#ifndef CTR_VERSION_H
#define CTR_VERSION_H
#define CTR_VERSION_MAJOR 0
#define CTR_VERSION_MINOR 0
#define CTR_VERSION_PATCH 0
#define CTR_VERSION 0
#define CTR_VERSION_STR "0.0.0"
#endif
