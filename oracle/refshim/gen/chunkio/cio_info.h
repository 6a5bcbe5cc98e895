/* stand-in for lib/chunkio/include/chunkio/cio_info.h.in */
#ifndef CIO_INFO_H
#define CIO_INFO_H
#define CIO_HAVE_TIMESPEC_GET
#define CIO_HAVE_GMTIME_R
#define CIO_HAVE_CFL
#endif
