/* stand-in for include/fluent-bit/flb_version.h.in: version constants only */
#ifndef FLB_VERSION_H
#define FLB_VERSION_H
#include <fluent-bit/flb_info.h>
#define FLB_VERSION_MAJOR 5
#define FLB_VERSION_MINOR 0
#define FLB_VERSION_PATCH 2
#define FLB_VERSION (FLB_VERSION_MAJOR * 10000 + FLB_VERSION_MINOR * 100 + FLB_VERSION_PATCH)
#define FLB_VERSION_STR "5.0.2"
#define FLB_GIT_HASH "3e414ac"
#endif
