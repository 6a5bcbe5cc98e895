/* stand-in for lib/cprofiles/include/cprofiles/cprof_version.h.in */
#ifndef CPROF_VERSION_H
#define CPROF_VERSION_H
#define CPROF_VERSION_MAJOR 0
#define CPROF_VERSION_MINOR 0
#define CPROF_VERSION_PATCH 0
#define CPROF_VERSION 0
#define CPROF_VERSION_STR "0.0.0"
#endif
