/* stand-in for lib/chunkio/include/chunkio/cio_version.h.in */
#ifndef CIO_VERSION_H
#define CIO_VERSION_H
#define CIO_VERSION_MAJOR 0
#define CIO_VERSION_MINOR 0
#define CIO_VERSION_PATCH 0
#define CIO_VERSION 0
#define CIO_VERSION_STR "0.0.0"
#endif
