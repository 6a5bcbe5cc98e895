/* dev_time.cuh -- time-string -> epoch on the device.
 *
 * Mirrors, step for step:
 *   flb_parser_time_lookup ... reference src/flb_parser.c:1159-1278
 *   parse_subseconds ......... src/flb_parser.c:1136-1157
 *   flb_strptime ............. src/flb_strptime.c:248-838 (OpenBSD-derived; C locale)
 *   flb_parser_tm2time ....... include/fluent-bit/flb_parser.h:78-92 (timegm - gmtoff)
 * The format text is interpreted directly (it lives in the program blob); the host
 * rejects conversions this file does not implement (%c %x %X %E* %O* and the
 * system-tzname fallback of %Z) when the parser is created.
 */
#ifndef FLBGPU_DEV_TIME_CUH
#define FLBGPU_DEV_TIME_CUH

#include <stdint.h>

#include "flbgpu_prog.h"
#ifndef FLB_HD
#ifdef __CUDACC__
#define FLB_HD __host__ __device__ __forceinline__
#ifdef FLB_INLINE_ALL
#define FLB_HDN __host__ __device__ __forceinline__
#else
#define FLB_HDN __host__ __device__ __noinline__
#endif
#else
#define FLB_HD static inline
#define FLB_HDN static
#endif
#endif

#ifdef __CUDA_ARCH__
#define DT_CONST __device__ static const
#else
#define DT_CONST static const
#endif

struct dt_tm {
    int sec, min, hour, mday, mon, year, wday, yday, isdst;
    long gmtoff;
};

struct dt_state { int century, relyear, fields; };

#define DT_F_MON  1
#define DT_F_MDAY 2
#define DT_F_WDAY 4
#define DT_F_YDAY 8
#define DT_F_YEAR 16

FLB_HD int dt_isspace(int c) { return c == ' ' || (c >= 9 && c <= 13); }
FLB_HD int dt_isdigit(int c) { return c >= '0' && c <= '9'; }
FLB_HD int dt_isalnum(int c) { return dt_isdigit(c) || ((c | 0x20) >= 'a' && (c | 0x20) <= 'z'); }
FLB_HD int dt_lower(int c) { return (c >= 'A' && c <= 'Z') ? c + 32 : c; }

/* strncasecmp(name, bp, strlen(name)) == 0 */
FLB_HD int dt_prefix_ci(const char *name, const unsigned char *bp, int *len)
{
    int i = 0;
    while (name[i]) {
        if (dt_lower((unsigned char) name[i]) != dt_lower(bp[i])) return 0;
        i++;
    }
    *len = i;
    return 1;
}

FLB_HD int dt_isleap(int y) { return (y % 4) == 0 && ((y % 100) != 0 || (y % 400) == 0); }
FLB_HD int dt_leaps_thru_end_of(int y)
{
    if (y >= 0) return y / 4 - y / 100 + y / 400;
    y = -(y + 1);
    return -((y / 4 - y / 100 + y / 400) + 1);
}

/* days since 1970-01-01 of the proleptic Gregorian date y-m-d (m in 1..12) */
FLB_HD int64_t dt_days_from_civil(int64_t y, int m, int d)
{
    int64_t era, yoe, doy, doe;
    y -= m <= 2;
    era = (y >= 0 ? y : y - 399) / 400;
    yoe = y - era * 400;
    doy = (153 * (m + (m > 2 ? -3 : 9)) + 2) / 5 + d - 1;
    doe = yoe * 365 + yoe / 4 - yoe / 100 + doy;
    return era * 146097 + doe - 719468;
}

/* timegm(): fields may be out of range; normalisation is linear like glibc's */
FLB_HD int64_t dt_timegm(const struct dt_tm *t)
{
    int64_t year = (int64_t) t->year + 1900, mon = t->mon, days;
    year += mon / 12;
    mon %= 12;
    if (mon < 0) { mon += 12; year--; }
    days = dt_days_from_civil(year, (int) mon + 1, 1) + (t->mday - 1);
    return days * 86400 + (int64_t) t->hour * 3600 + (int64_t) t->min * 60 + t->sec;
}

/* gmtime_r() for %s */
FLB_HD void dt_gmtime(int64_t t, struct dt_tm *tm)
{
    int64_t days = t / 86400, rem = t % 86400, z, era, doe, yoe, y, doy, mp, m, d;
    int k;
    if (rem < 0) { rem += 86400; days--; }
    tm->hour = (int) (rem / 3600); tm->min = (int) ((rem % 3600) / 60); tm->sec = (int) (rem % 60);
    tm->wday = (int) ((days + 4) % 7); if (tm->wday < 0) tm->wday += 7;
    z = days + 719468;
    era = (z >= 0 ? z : z - 146096) / 146097;
    doe = z - era * 146097;
    yoe = (doe - doe / 1460 + doe / 36524 - doe / 146096) / 365;
    y = yoe + era * 400;
    doy = doe - (365 * yoe + yoe / 4 - yoe / 100);
    mp = (5 * doy + 2) / 153;
    d = doy - (153 * mp + 2) / 5 + 1;
    m = mp + (mp < 10 ? 3 : -9);
    if (m <= 2) y++;
    tm->year = (int) (y - 1900); tm->mon = (int) (m - 1); tm->mday = (int) d;
    {
        const int cum[12] = { 0, 31, 59, 90, 120, 151, 181, 212, 243, 273, 304, 334 };
        k = cum[tm->mon] + tm->mday - 1;
        if (tm->mon > 1 && dt_isleap((int) y)) k++;
        tm->yday = k;
    }
    tm->isdst = 0;
}

FLB_HD int dt_conv_num(const unsigned char **buf, int *dest, int llim, int ulim)
{
    int result = 0, rulim = ulim;
    if (**buf < '0' || **buf > '9') return 0;
    do {
        result *= 10;
        result += *(*buf)++ - '0';
        rulim /= 10;
    } while ((result * 10 <= ulim) && rulim && **buf >= '0' && **buf <= '9');
    if (result < llim || result > ulim) return 0;
    *dest = result;
    return 1;
}

FLB_HD int dt_conv_num64(const unsigned char **buf, int64_t *dest, int64_t llim, int64_t ulim)
{
    int64_t result = 0, rulim = ulim;
    if (**buf < '0' || **buf > '9') return 0;
    do {
        if (result > 922337203685477580ll) return 0;
        result *= 10;
        if (result > 9223372036854775760ll) return 0;
        result += *(*buf)++ - '0';
        rulim /= 10;
        if (result >= 922337203685477580ll) return 0;
    } while ((result * 10 <= ulim) && rulim && **buf >= '0' && **buf <= '9');
    if (result < llim || result > ulim) return 0;
    *dest = result;
    return 1;
}

struct dt_tz { const char *abbr; int off; int dst; };

#define DT_H 3600
/* src/flb_strptime.c:100-196 flb_known_timezones (same order: first hit wins) */
#define DT_TZ_TABLE(X) \
  X("GMT",0,0) X("UTC",0,0) X("Z",0,0) X("UT",0,0) \
  X("EST",-5*DT_H,0) X("EDT",-4*DT_H,1) X("CST",-6*DT_H,0) X("CDT",-5*DT_H,1) X("MST",-7*DT_H,0) X("MDT",-6*DT_H,1) \
  X("PST",-8*DT_H,0) X("PDT",-7*DT_H,1) X("AKST",-9*DT_H,0) X("AKDT",-8*DT_H,1) X("HST",-10*DT_H,0) X("HADT",-9*DT_H,1) \
  X("AST",-4*DT_H,0) X("ADT",-3*DT_H,1) X("NST",-12600,0) X("NDT",-9000,1) \
  X("WET",0,0) X("WEST",1*DT_H,1) X("CET",1*DT_H,0) X("CEST",2*DT_H,1) X("EET",2*DT_H,0) X("EEST",3*DT_H,1) X("MSK",3*DT_H,0) \
  X("ART",-3*DT_H,0) X("BRT",-3*DT_H,0) X("BRST",-2*DT_H,1) X("CLT",-4*DT_H,0) X("CLST",-3*DT_H,1) \
  X("AEST",10*DT_H,0) X("AEDT",11*DT_H,1) X("ACST",34200,0) X("ACDT",37800,1) X("AWST",8*DT_H,0) X("NZST",12*DT_H,0) X("NZDT",13*DT_H,1) \
  X("JST",9*DT_H,0) X("KST",9*DT_H,0) X("SGT",8*DT_H,0) X("IST",19800,0) X("GST",4*DT_H,0) X("ICT",7*DT_H,0) X("WIB",7*DT_H,0) \
  X("WITA",8*DT_H,0) X("WIT",9*DT_H,0) X("MYT",8*DT_H,0) X("BDT",6*DT_H,0) X("NPT",20700,0) \
  X("WAT",1*DT_H,0) X("CAT",2*DT_H,0) X("EAT",3*DT_H,0) X("SAST",2*DT_H,0) \
  X("A",1*DT_H,0) X("B",2*DT_H,0) X("C",3*DT_H,0) X("D",4*DT_H,0) X("E",5*DT_H,0) X("F",6*DT_H,0) X("G",7*DT_H,0) X("H",8*DT_H,0) \
  X("I",9*DT_H,0) X("K",10*DT_H,0) X("L",11*DT_H,0) X("M",12*DT_H,0) X("N",-1*DT_H,0) X("O",-2*DT_H,0) X("P",-3*DT_H,0) X("Q",-4*DT_H,0) \
  X("R",-5*DT_H,0) X("S",-6*DT_H,0) X("T",-7*DT_H,0) X("U",-8*DT_H,0) X("V",-9*DT_H,0) X("W",-10*DT_H,0) X("X",-11*DT_H,0) X("Y",-12*DT_H,0)

/* returns 1 and advances *bp when a known abbreviation (followed by a non-alnum) matches */
FLB_HD int dt_known_tz(const unsigned char **bp, struct dt_tm *tm)
{
    int len;
#define X(name, o, d) if (dt_prefix_ci(name, *bp, &len) && !dt_isalnum((*bp)[len])) { tm->isdst = d; tm->gmtoff = (o); *bp += len; return 1; }
    DT_TZ_TABLE(X)
#undef X
    return 0;
}

/* 0..6 / 0..11 by full then abbreviated English name, -1 when none; *len = matched length */
FLB_HD int dt_find_day(const unsigned char *bp, int *len)
{
#define D(i, full, ab) if (dt_prefix_ci(full, bp, len)) return i; if (dt_prefix_ci(ab, bp, len)) return i;
    D(0, "Sunday", "Sun") D(1, "Monday", "Mon") D(2, "Tuesday", "Tue") D(3, "Wednesday", "Wed")
    D(4, "Thursday", "Thu") D(5, "Friday", "Fri") D(6, "Saturday", "Sat")
    return -1;
}
FLB_HD int dt_find_mon(const unsigned char *bp, int *len)
{
    D(0, "January", "Jan") D(1, "February", "Feb") D(2, "March", "Mar") D(3, "April", "Apr")
    D(4, "May", "May") D(5, "June", "Jun") D(6, "July", "Jul") D(7, "August", "Aug")
    D(8, "September", "Sep") D(9, "October", "Oct") D(10, "November", "Nov") D(11, "December", "Dec")
#undef D
    return -1;
}

/* tail of _flb_strptime(): resolve %y/%C and derive yday/wday/mon/mday when possible
 * (src/flb_strptime.c:765-812); runs at the end of every (nested) format */
FLB_HD void dt_finalize(struct dt_tm *tm, struct dt_state *st)
{
    int i;
    if (st->relyear != -1) {
        if (st->century == 1900) {
            if (st->relyear <= 68) tm->year = st->relyear + 2000 - 1900;
            else tm->year = st->relyear + 1900 - 1900;
        }
        else tm->year = st->relyear + st->century - 1900;
        st->fields |= DT_F_YEAR;
    }
    if (st->fields & DT_F_YEAR) {
        const int year = (int) ((unsigned) tm->year + 1900u);
        const int ml[2][12] = { { 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31 },
                                { 31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31 } };
        const int *mon_lens = ml[dt_isleap(year)];
        if (!(st->fields & DT_F_YDAY) && (st->fields & DT_F_MON) && (st->fields & DT_F_MDAY)) {
            tm->yday = tm->mday - 1;
            for (i = 0; i < tm->mon; i++) tm->yday += mon_lens[i];
            st->fields |= DT_F_YDAY;
        }
        if (st->fields & DT_F_YDAY) {
            int days = tm->yday;
            if (!(st->fields & DT_F_WDAY)) {
                tm->wday = 4 + ((year - 1970) % 7) * (365 % 7) + dt_leaps_thru_end_of(year - 1) -
                           dt_leaps_thru_end_of(1969) + tm->yday;
                tm->wday %= 7;
                if (tm->wday < 0) tm->wday += 7;
            }
            if (!(st->fields & DT_F_MON)) {
                tm->mon = 0;
                while (tm->mon < 12 && days >= mon_lens[tm->mon]) days -= mon_lens[tm->mon++];
            }
            if (!(st->fields & DT_F_MDAY)) tm->mday = days + 1;
        }
    }
}

/* _flb_strptime(): returns the position after the last consumed byte, or NULL.
 * buf must be NUL terminated.  The reference recurses for %D %R %r %T %F; their
 * expansions contain no nested composite, so one saved return pointer replaces the
 * recursion (no device call stack needed). */
FLB_HDN const unsigned char *dt_strptime(const unsigned char *bp, const char *fmt, struct dt_tm *tm,
                                         struct dt_state *st, int initialize)
{
    unsigned char c;
    int i, len, offs, neg;
    const char *sub, *ret_fmt = 0;

    if (initialize) {
        st->century = 1900; st->relyear = -1; st->fields = 0;
        tm->gmtoff = 0; tm->isdst = -1;
    }
    for (;;) {
        c = (unsigned char) *fmt;
        if (c == '\0') {
            if (!ret_fmt) break;
            dt_finalize(tm, st);            /* end of the nested format */
            fmt = ret_fmt; ret_fmt = 0;
            continue;
        }
        if (dt_isspace(c)) {
            while (dt_isspace(*bp)) bp++;
            fmt++;
            continue;
        }
        if (*bp == '\0') return 0;
        if ((c = (unsigned char) *fmt++) != '%') goto literal;
again:
        switch (c = (unsigned char) *fmt++) {
        case '%':
literal:
            if (c != *bp++) return 0;
            break;
        case 'E': case 'O':            /* alternative modifiers are accepted and ignored */
            goto again;
        case 'D': sub = "%m/%d/%y"; goto recurse;
        case 'R': sub = "%H:%M"; goto recurse;
        case 'r': sub = "%I:%M:%S %p"; goto recurse;
        case 'T': sub = "%H:%M:%S"; goto recurse;
        case 'F': sub = "%Y-%m-%d";
recurse:
            if (ret_fmt) return 0;          /* cannot happen: expansions hold no composites */
            ret_fmt = fmt; fmt = sub;
            continue;
        case 'A': case 'a':
            i = dt_find_day(bp, &len);
            if (i < 0) return 0;
            tm->wday = i; bp += len; st->fields |= DT_F_WDAY;
            break;
        case 'B': case 'b': case 'h':
            i = dt_find_mon(bp, &len);
            if (i < 0) return 0;
            tm->mon = i; bp += len; st->fields |= DT_F_MON;
            break;
        case 'C':
            if (!dt_conv_num(&bp, &i, 0, 99)) return 0;
            st->century = i * 100;
            break;
        case 'e':
            if (dt_isspace(*bp)) bp++;
            /* fall through */
        case 'd':
            if (!dt_conv_num(&bp, &tm->mday, 1, 31)) return 0;
            st->fields |= DT_F_MDAY;
            break;
        case 'k': case 'H':
            if (!dt_conv_num(&bp, &tm->hour, 0, 23)) return 0;
            break;
        case 'l': case 'I':
            if (!dt_conv_num(&bp, &tm->hour, 1, 12)) return 0;
            break;
        case 'j':
            if (!dt_conv_num(&bp, &tm->yday, 1, 366)) return 0;
            tm->yday--;
            st->fields |= DT_F_YDAY;
            break;
        case 'M':
            if (!dt_conv_num(&bp, &tm->min, 0, 59)) return 0;
            break;
        case 'm':
            if (!dt_conv_num(&bp, &tm->mon, 1, 12)) return 0;
            tm->mon--;
            st->fields |= DT_F_MON;
            break;
        case 'p':
            if (dt_prefix_ci("AM", bp, &len)) {
                if (tm->hour > 12) return 0;
                else if (tm->hour == 12) tm->hour = 0;
                bp += len;
                break;
            }
            if (dt_prefix_ci("PM", bp, &len)) {
                if (tm->hour > 12) return 0;
                else if (tm->hour < 12) tm->hour += 12;
                bp += len;
                break;
            }
            return 0;
        case 'S':
            if (!dt_conv_num(&bp, &tm->sec, 0, 60)) return 0;
            break;
        case 's': {
            int64_t i64;
            if (!dt_conv_num64(&bp, &i64, 0, 9223372036854775807ll)) return 0;
            dt_gmtime(i64, tm);
            tm->gmtoff = 0;
            tm->isdst = 0;
            st->fields = 0xffff;
            break;
        }
        case 'U': case 'W':
            if (!dt_conv_num(&bp, &i, 0, 53)) return 0;
            break;
        case 'w':
            if (!dt_conv_num(&bp, &tm->wday, 0, 6)) return 0;
            st->fields |= DT_F_WDAY;
            break;
        case 'u':
            if (!dt_conv_num(&bp, &i, 1, 7)) return 0;
            tm->wday = i % 7;
            st->fields |= DT_F_WDAY;
            continue;
        case 'g':
            if (!dt_conv_num(&bp, &i, 0, 99)) return 0;
            continue;
        case 'G':
            do bp++; while (dt_isdigit(*bp));
            continue;
        case 'V':
            if (!dt_conv_num(&bp, &i, 0, 53)) return 0;
            continue;
        case 'Y':
            if (!dt_conv_num(&bp, &i, 0, 9999)) return 0;
            st->relyear = -1;
            tm->year = i - 1900;
            st->fields |= DT_F_YEAR;
            break;
        case 'y':
            if (!dt_conv_num(&bp, &st->relyear, 0, 99)) return 0;
            break;
        case 'Z':
            if (dt_known_tz(&bp, tm)) continue;
            if (bp[0] == 'G' && bp[1] == 'M' && bp[2] == 'T') { tm->isdst = 0; tm->gmtoff = 0; bp += 3; continue; }
            if (bp[0] == 'U' && bp[1] == 'T' && bp[2] == 'C') { tm->isdst = 0; tm->gmtoff = 0; bp += 3; continue; }
            return 0;    /* system tzname[] fallback is not reproduced (environment dependent) */
        case 'z':
            while (dt_isspace(*bp)) bp++;
            neg = 0;
            switch (*bp++) {
            case 'G':
                if (*bp++ != 'M') return 0;
                if (*bp++ != 'T') return 0;
                tm->isdst = 0; tm->gmtoff = 0;
                continue;
            case 'U':
                if (*bp++ != 'T') return 0;
                if (*bp == 'C') bp++;
                tm->isdst = 0; tm->gmtoff = 0;
                continue;
            case 'Z':
                tm->isdst = 0; tm->gmtoff = 0;
                continue;
            case '+': neg = 0; break;
            case '-': neg = 1; break;
            default:
                --bp;
                if (dt_prefix_ci("EST", bp, &len)) i = 0;
                else if (dt_prefix_ci("CST", bp, &len)) i = 1;
                else if (dt_prefix_ci("MST", bp, &len)) i = 2;
                else if (dt_prefix_ci("PST", bp, &len)) i = 3;
                else i = -1;
                if (i >= 0) { tm->gmtoff = (-5 - i) * 3600; tm->isdst = 0; bp += len; continue; }
                if (dt_prefix_ci("EDT", bp, &len)) i = 0;
                else if (dt_prefix_ci("CDT", bp, &len)) i = 1;
                else if (dt_prefix_ci("MDT", bp, &len)) i = 2;
                else if (dt_prefix_ci("PDT", bp, &len)) i = 3;
                else i = -1;
                if (i >= 0) { tm->isdst = 1; tm->gmtoff = (-4 - i) * 3600; bp += len; continue; }
                return 0;
            }
            if (!dt_isdigit(bp[0]) || !dt_isdigit(bp[1])) return 0;
            offs = ((bp[0] - '0') * 10 + (bp[1] - '0')) * 3600;
            bp += 2;
            if (*bp == ':') bp++;
            if (dt_isdigit(*bp)) {
                offs += (*bp++ - '0') * 10 * 60;
                if (!dt_isdigit(*bp)) return 0;
                offs += (*bp++ - '0') * 60;
            }
            if (neg) offs = -offs;
            tm->isdst = 0;
            tm->gmtoff = offs;
            continue;
        case 'n': case 't':
            while (dt_isspace(*bp)) bp++;
            break;
        default:
            return 0;
        }
    }
    dt_finalize(tm, st);
    return bp;
}

/* device-side view of the time part of struct flb_parser */
/* "%d/%b/%Y:%H:%M:%S %z" (apache / nginx access logs) on a value of the canonical shape
 * DD/Mon/YYYY:HH:MM:SS +hhmm: the same fields flb_strptime() produces for it, without the directive
 * interpreter and the copy into a NUL-terminated buffer.  Anything else (other lengths, one-digit
 * days, out-of-range numbers, a zone in another form) returns 0 and takes the general path. */
FLB_HD int dt_fast_apache(const uint8_t *s, uint32_t n, struct dt_tm *tm)
{
    static const char mon3[] = "janfebmaraprmayjunjulaugsepoctnovdec";
    int d[14], i, m;
    /* digit positions of DD/Mon/YYYY:HH:MM:SS +hhmm */
    const int pos[14] = { 0, 1, 7, 8, 9, 10, 12, 13, 15, 16, 18, 19, 22, 23 };
    if (n != 26 || s[2] != '/' || s[6] != '/' || s[11] != ':' || s[14] != ':' || s[17] != ':' || s[20] != ' ') return 0;
    if (s[21] != '+' && s[21] != '-') return 0;
    for (i = 0; i < 14; i++) { d[i] = s[pos[i]] - '0'; if ((unsigned) d[i] > 9u) return 0; }
    if ((unsigned) (s[24] - '0') > 9u || (unsigned) (s[25] - '0') > 9u) return 0;
    for (m = 0; m < 12; m++)
        if (dt_lower(s[3]) == mon3[3 * m] && dt_lower(s[4]) == mon3[3 * m + 1] && dt_lower(s[5]) == mon3[3 * m + 2]) break;
    if (m == 12) return 0;
    {   /* nothing is written unless every field is in range: the general path starts from the caller's zeroed tm,
         * and in non-strict mode what it leaves half-filled is what the reference reports */
        const int mday = d[0] * 10 + d[1], hour = d[6] * 10 + d[7], min = d[8] * 10 + d[9], sec = d[10] * 10 + d[11];
        if (mday < 1 || mday > 31 || hour > 23 || min > 59 || sec > 60) return 0;
        tm->mday = mday; tm->mon = m; tm->hour = hour; tm->min = min; tm->sec = sec;
    }
    tm->year = d[2] * 1000 + d[3] * 100 + d[4] * 10 + d[5] - 1900;
    tm->gmtoff = ((d[12] * 10 + d[13]) * 3600 + ((s[24] - '0') * 10 + (s[25] - '0')) * 60) * (s[21] == '-' ? -1 : 1);
    tm->isdst = 0;
    return 1;
}

/* Fixed-shape time program (TF_* ops, compiled by runtime.c:time_fast_compile): 1 = tm / ns filled
 * exactly as the general path would, 0 = not this shape, let the general path decide. */
FLB_HD int dt_fast_prog(const uint8_t *prog, const uint8_t *s, uint32_t n, struct dt_tm *tm, double *ns)
{
    static const char mon3[] = "janfebmaraprmayjunjulaugsepoctnovdec";
    uint32_t p = 0;
    for (;;) {
        const uint32_t op = *prog++;
        switch (op) {
        case TF_END:
            return 1;
        case TF_D2: {
            const uint32_t f = prog[0], lo = prog[1], hi = prog[2];
            uint32_t d0, d1, v;
            prog += 3;
            if (p + 2 > n) return 0;
            d0 = (uint32_t) s[p] - '0'; d1 = (uint32_t) s[p + 1] - '0';
            if (d0 > 9u || d1 > 9u) return 0;
            v = d0 * 10 + d1;
            if (v < lo || v > hi) return 0;
            p += 2;
            if (f == TFF_MDAY) tm->mday = (int) v; else if (f == TFF_MON) tm->mon = (int) v - 1;
            else if (f == TFF_HOUR) tm->hour = (int) v; else if (f == TFF_MIN) tm->min = (int) v; else tm->sec = (int) v;
            break;
        }
        case TF_Y4: {
            uint32_t k, v = 0;
            if (p + 4 > n) return 0;
            for (k = 0; k < 4; k++) { const uint32_t d = (uint32_t) s[p + k] - '0'; if (d > 9u) return 0; v = v * 10 + d; }
            p += 4;
            tm->year = (int) v - 1900;
            break;
        }
        case TF_LIT:
            if (p >= n || s[p] != *prog) return 0;
            prog++; p++;
            break;
        case TF_SPACE:
            if (p >= n || s[p] != ' ') return 0;
            p++;
            if (p < n && dt_isspace(s[p])) return 0;
            break;
        case TF_MON3: {
            int m;
            if (p + 3 > n) return 0;
            for (m = 0; m < 12; m++)
                if (dt_lower(s[p]) == mon3[3 * m] && dt_lower(s[p + 1]) == mon3[3 * m + 1] && dt_lower(s[p + 2]) == mon3[3 * m + 2]) break;
            if (m == 12) return 0;
            p += 3;
            if (p < n && ((s[p] | 0x20) >= 'a' && (s[p] | 0x20) <= 'z')) return 0;     /* maybe a full month name */
            tm->mon = m;
            break;
        }
        case TF_TZ: {
            uint32_t c0 = p < n ? s[p] : 0, h0, h1, offs;
            if (c0 == 'Z') { p++; tm->isdst = 0; tm->gmtoff = 0; break; }
            if (c0 != '+' && c0 != '-') return 0;
            p++;
            h0 = (p < n ? (uint32_t) s[p] : 0u) - '0'; h1 = (p + 1 < n ? (uint32_t) s[p + 1] : 0u) - '0';
            if (h0 > 9u || h1 > 9u) return 0;
            offs = (h0 * 10 + h1) * 3600;
            p += 2;
            if (p < n && s[p] == ':') p++;
            if (p < n && dt_isdigit(s[p])) {
                const uint32_t m0 = (uint32_t) s[p] - '0', m1 = (p + 1 < n ? (uint32_t) s[p + 1] : 0u) - '0';
                if (m1 > 9u) return 0;
                offs += (m0 * 10 + m1) * 60;
                p += 2;
            }
            tm->isdst = 0;
            tm->gmtoff = c0 == '-' ? -(long) offs : (long) offs;
            break;
        }
        case TF_FRAC: {
            uint32_t avail = n - p, digits = avail < 9 ? avail : 9, nd = 0;
            uint64_t num = 0;
            double den = 1.0;
            while (nd < digits && dt_isdigit(s[p + nd])) { num = num * 10 + (s[p + nd] - '0'); den *= 10.0; nd++; }
            if (nd == 0) return 0;
            if (nd == 9 && p + 9 < n && dt_isdigit(s[p + 9])) return 0;
            *ns = (double) num / den;
            p += nd;
            tm->gmtoff = 0; tm->isdst = -1;          /* the part after %L is a fresh flb_strptime() call */
            break;
        }
        default:
            return 0;
        }
    }
}

struct dt_parser {
    const char *fmt;        /* format up to %L (already prefixed with "%Y " when !with_year) */
    const char *frac_fmt;   /* format after %L, or NULL when the format has no %L */
    int with_year, with_tz, strict;
    int offset;             /* Time_Offset in seconds */
    int fast_apache;        /* the format is exactly "%d/%b/%Y:%H:%M:%S %z" */
    const uint8_t *tfast;   /* fixed-shape program for the format, or NULL */
};

/* flb_parser_time_lookup(): 0 ok (tm/frac filled, maybe partially), -1 error */
FLB_HDN int dt_time_lookup(const uint8_t *s, uint32_t tsize, int64_t now, const struct dt_parser *p,
                           struct dt_tm *tm, double *ns)
{
    unsigned char tmp[64];
    const unsigned char *q;
    struct dt_state st;
    int time_len = (int) tsize, i;

    *ns = 0;
    if (p->fast_apache && dt_fast_apache(s, tsize, tm)) return 0;
    if (p->tfast && tsize <= 63) {
        struct dt_tm t2 = *tm;
        double ns2 = 0;
        if (dt_fast_prog(p->tfast, s, tsize, &t2, &ns2)) {
            *tm = t2; *ns = ns2;
            if (!p->with_tz) tm->gmtoff = p->offset;
            return 0;
        }
    }
    if (tsize > sizeof(tmp) - 1) return -1;
    if (!p->with_year) {
        struct dt_tm tmy;
        unsigned y;
        if (time_len + 6 >= (int) sizeof(tmp)) return -1;
        dt_gmtime(now, &tmy);
        tm->mon = tmy.mon;
        tm->mday = tmy.mday;
        y = (unsigned) (tmy.year + 1900);
        tmp[0] = '0' + (y / 1000) % 10; tmp[1] = '0' + (y / 100) % 10;
        tmp[2] = '0' + (y / 10) % 10; tmp[3] = '0' + y % 10;
        tmp[4] = ' ';
        for (i = 0; i < time_len; i++) tmp[5 + i] = s[i];
        tmp[5 + time_len] = 0;
    }
    else {
        for (i = 0; i < time_len; i++) tmp[i] = s[i];
        tmp[time_len] = 0;
    }
    time_len = 0;
    while (tmp[time_len]) time_len++;              /* strlen(): stops at an embedded NUL */

    q = dt_strptime(tmp, p->fmt, tm, &st, 1);
    if (!q) return p->strict ? -1 : 0;
    if (p->frac_fmt) {
        /* parse_subseconds(): strtod("0." + up to 9 following bytes) */
        int avail = time_len - (int) (q - tmp), digits = avail < 9 ? avail : 9, nd = 0;
        uint64_t num = 0;
        double den = 1.0;
        while (nd < digits && dt_isdigit(q[nd])) { num = num * 10 + (q[nd] - '0'); den *= 10.0; nd++; }
        if (nd <= 0) return p->strict ? -1 : 0;
        *ns = (double) num / den;                  /* both exact, one correctly rounded division */
        q += nd;
        q = dt_strptime(q, p->frac_fmt, tm, &st, 1);
        if (!q) return p->strict ? -1 : 0;
    }
    if (!p->with_tz) tm->gmtoff = p->offset;
    return 0;
}

#endif
