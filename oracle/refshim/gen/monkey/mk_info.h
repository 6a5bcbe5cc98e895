/* stand-in for lib/monkey/include/monkey/mk_info.h.in */
#ifndef MK_INFO_H
#define MK_INFO_H
#include <monkey/mk_core.h>
#define MK_VERSION_MAJOR 1
#define MK_VERSION_MINOR 8
#define MK_VERSION_PATCH 0
#define MK_VERSION 10800
#define MK_VERSION_STR "1.8.0"
#define MK_BUILD_OS "Linux"
#define MK_BUILD_UNAME "Linux"
#define MK_BUILD_CMD ""
#define MK_PATH_CONF ""
#define MK_PLUGIN_DIR ""
#define MK_HAVE_ACCEPT4
#endif
