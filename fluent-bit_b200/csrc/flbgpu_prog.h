/* flbgpu_prog.h -- plain-old-data layouts shared by the host-side compilers (C)
 * and the device interpreters (CUDA).  Everything a kernel needs at run time
 * (regex programs, strptime programs, the filter-chain program, the constant
 * pool) lives in ONE relocatable byte blob; every cross reference is a byte
 * offset from the blob base, so the blob is uploaded to HBM with one memcpy.
 */
#ifndef FLBGPU_PROG_H
#define FLBGPU_PROG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ regex */
/* Instruction word: low 8 bits opcode, high 24 bits argument.  Semantics are
 * those of Onigmo's backtracking matcher (lib/onigmo/regexec.c:1431 match_at):
 * leftmost start, alternatives tried in priority order, captures restored on
 * backtrack. */
enum {
    RX_MATCH = 0,      /* success */
    RX_FAIL,           /* force backtrack */
    RX_BYTE,           /* arg = byte value */
    RX_STR,            /* arg = n bytes; bytes packed in following ceil(n/4) words */
    RX_CLASS,          /* arg = class index; consumes one character */
    RX_ANY,            /* any character except '\n' (OP_ANYCHAR) */
    RX_ANY_NL,         /* any character (OP_ANYCHAR_ML) */
    RX_JMP,            /* arg = absolute pc */
    RX_SPLIT,          /* continue at pc+1, alternative = arg (absolute pc) */
    RX_SPLIT_LAZY,     /* continue at arg, alternative = pc+1 */
    RX_SAVE,           /* arg = capture slot (2*g or 2*g+1) */
    RX_CSTAR_POSS,     /* arg = class: cls* with no alternatives left behind */
    RX_CSTAR_BT,       /* arg = class: greedy cls*, one back-off frame */
    RX_CSTAR_LAZY,     /* arg = class: lazy cls*? */
    RX_BOL,            /* ^  (ANCHOR_BEGIN_LINE) */
    RX_EOL,            /* $  (ANCHOR_END_LINE) */
    RX_BEGIN_BUF,      /* \A */
    RX_END_BUF,        /* \z */
    RX_SEMI_END_BUF,   /* \Z */
    RX_WORD_B,         /* \b */
    RX_NOT_WORD_B,     /* \B */
    RX_NULL_START,     /* arg = loop id: remember position (OP_NULL_CHECK_START) */
    RX_NULL_END,       /* arg = loop id: if iteration was empty skip next insn */
    RX_MARK,           /* arg = kind: push a cut mark (look-ahead / atomic) */
    RX_CUT_POS,        /* positive look-ahead end: cut to mark, restore position */
    RX_CUT_ATOMIC,     /* atomic group end: cut to mark, keep position */
    RX_CUT_NEG,        /* negative look-ahead body matched: cut to mark, drop the
                          alternative below it, fail */
    RX_BACKREF,        /* arg = group number */
    RX_OPCOUNT
};

#define RX_OP(w)   ((w) & 0xffu)
#define RX_ARG(w)  ((w) >> 8)
#define RX_MK(op, arg) (((uint32_t)(arg) << 8) | (uint32_t)(op))

/* class.mb_mode */
enum { RX_MB_NONE = 0, RX_MB_ALL = 1, RX_MB_RANGES = 2, RX_MB_NOT_RANGES = 3 };

struct rx_class {
    uint32_t bits[8];      /* membership of single-byte characters (incl. invalid high bytes) */
    uint32_t mb_mode;      /* how valid multi-byte characters are decided */
    uint32_t n_ranges;     /* code point ranges [lo,hi] for RX_MB_RANGES / RX_MB_NOT_RANGES */
    uint32_t ranges_off;   /* byte offset from the rx_prog header to uint32 pairs */
    uint32_t pad;
};

#define RX_MAX_GROUPS 47   /* named/numbered groups, group 0 excluded */
#define RX_F_ANCHOR_BOL   1u   /* every match starts at a line start   */
#define RX_F_ANCHOR_BUF   2u   /* every match starts at offset 0       */
#define RX_F_HAS_FIRSTSET 4u   /* first[] is a sound first-byte filter */
#define RX_F_NULLABLE     8u   /* the pattern can match the empty string */

struct rx_prog {
    uint32_t total_bytes;  /* size of this program incl. header, code, classes, ranges */
    uint32_t n_code;       /* 32-bit words of code */
    uint32_t code_off;     /* byte offset of code[] from this header */
    uint32_t n_classes;
    uint32_t class_off;    /* byte offset of struct rx_class[] */
    uint32_t n_groups;     /* capture groups (group 0 excluded) */
    uint32_t flags;
    uint32_t n_null;       /* number of null-check loop ids */
    uint32_t first[8];     /* first-byte filter (valid when RX_F_HAS_FIRSTSET) */
};

/* run-time status of one match attempt */
#define RX_R_NOMATCH   0
#define RX_R_MATCH     1
#define RX_R_ESTACK   -2   /* backtrack stack exhausted: rerun with a bigger stack */
#define RX_R_EBUDGET  -3   /* step budget exhausted (catastrophic backtracking guard) */

#ifdef __cplusplus
}
#endif
#endif
