/* Hand-written stand-in for the header the reference's build normally generates
 * from include/fluent-bit/flb_info.h.in (reference file:line 27-33).  Only the
 * feature switches the parse->filter path needs are turned on. */
#ifndef FLB_INFO_H
#define FLB_INFO_H
#define STR_HELPER(s)      #s
#define STR(s)             STR_HELPER(s)
#define FLB_SOURCE_DIR "/root/reference"
#define FLB_HAVE_PARSER
#define FLB_HAVE_REGEX
#define FLB_HAVE_RECORD_ACCESSOR
#define FLB_HAVE_YYJSON
#define FLB_HAVE_SIMD
#define FLB_HAVE_METRICS
#define FLB_HAVE_GMTOFF
#define FLB_HAVE_C_TLS
#define FLB_HAVE_TIMESPEC_GET
#define FLB_HAVE_LITTLE_ENDIAN_SYSTEM
#define FLB_HAVE_ATTRIBUTE_ALLOC_SIZE
#define FLB_HAVE_FORK
#define FLB_HAVE_UNIX_SOCKET
#define FLB_HAVE_ACCEPT4
#define FLB_EVENT_LOOP_EPOLL
#define JSMN_PARENT_LINKS
#define JSMN_STRICT
#ifndef FLB_MSGPACK_TO_JSON_INIT_BUFFER_SIZE
#define FLB_MSGPACK_TO_JSON_INIT_BUFFER_SIZE 2.0
#endif
#ifndef FLB_MSGPACK_TO_JSON_REALLOC_BUFFER_SIZE
#define FLB_MSGPACK_TO_JSON_REALLOC_BUFFER_SIZE 0.10
#endif
#ifndef FLB_CORO_STACK_SIZE
#define FLB_CORO_STACK_SIZE 24576
#endif
#define FLB_INFO_FLAGS "oracle-subset"
#endif
