/* Hand-written stand-in for the header Onigmo's cmake generates from lib/onigmo/config.h.cmake
 * (x86-64 Linux, glibc; the options of lib/onigmo/CMakeLists.txt:85-86 at their defaults: off).
 * Test infrastructure only (oracle/_ref). */
#ifndef FLBREF_ONIGMO_CONFIG_H
#define FLBREF_ONIGMO_CONFIG_H
#define HAVE_ALLOCA 1
#define HAVE_ALLOCA_H 1
#define HAVE_DLFCN_H 1
#define HAVE_INTTYPES_H 1
#define HAVE_MEMORY_H 1
#define HAVE_STDINT_H 1
#define HAVE_STDLIB_H 1
#define HAVE_STRINGS_H 1
#define HAVE_STRING_H 1
#define HAVE_SYS_STAT_H 1
#define HAVE_SYS_TIMES_H 1
#define HAVE_SYS_TIME_H 1
#define HAVE_SYS_TYPES_H 1
#define HAVE_UNISTD_H 1
#define SIZEOF_INT 4
#define SIZEOF_LONG 8
#define SIZEOF_LONG_LONG 8
#define SIZEOF_SHORT 2
#define SIZEOF_VOIDP 8
#define STDC_HEADERS 1
#define TIME_WITH_SYS_TIME 1
#endif
