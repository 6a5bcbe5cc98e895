/* oracle: flb_strptime() restated.  TEST INFRASTRUCTURE (see orc.h).
 * Follows src/flb_strptime.c:248-838 (OpenBSD-derived, C locale) directive by directive, including
 * the function-static century/relyear/fields state, _conv_num's digit budget and the table of
 * known timezone abbreviations (:92-196). */
#define _GNU_SOURCE
#include <ctype.h>
#include <string.h>
#include <strings.h>
#include <time.h>
#include "orc.h"

#define F_WDAY 1
#define F_MDAY 2
#define F_MON 4
#define F_YDAY 8
#define F_YEAR 16

static const struct { const char *abbr; long off; int dst; } zones[] = {
    { "GMT", 0, 0 }, { "UTC", 0, 0 }, { "Z", 0, 0 }, { "UT", 0, 0 },
    { "EST", -5 * 3600, 0 }, { "EDT", -4 * 3600, 1 }, { "CST", -6 * 3600, 0 }, { "CDT", -5 * 3600, 1 },
    { "MST", -7 * 3600, 0 }, { "MDT", -6 * 3600, 1 }, { "PST", -8 * 3600, 0 }, { "PDT", -7 * 3600, 1 },
    { "AKST", -9 * 3600, 0 }, { "AKDT", -8 * 3600, 1 }, { "HST", -10 * 3600, 0 }, { "HADT", -9 * 3600, 1 },
    { "AST", -4 * 3600, 0 }, { "ADT", -3 * 3600, 1 }, { "NST", -12600, 0 }, { "NDT", -9000, 1 },
    { "WET", 0, 0 }, { "WEST", 3600, 1 }, { "CET", 3600, 0 }, { "CEST", 7200, 1 }, { "EET", 7200, 0 }, { "EEST", 10800, 1 },
    { "MSK", 10800, 0 },
    { "ART", -3 * 3600, 0 }, { "BRT", -3 * 3600, 0 }, { "BRST", -2 * 3600, 1 }, { "CLT", -4 * 3600, 0 }, { "CLST", -3 * 3600, 1 },
    { "AEST", 36000, 0 }, { "AEDT", 39600, 1 }, { "ACST", 34200, 0 }, { "ACDT", 37800, 1 }, { "AWST", 28800, 0 },
    { "NZST", 43200, 0 }, { "NZDT", 46800, 1 },
    { "JST", 32400, 0 }, { "KST", 32400, 0 }, { "SGT", 28800, 0 }, { "IST", 19800, 0 }, { "GST", 14400, 0 }, { "ICT", 25200, 0 },
    { "WIB", 25200, 0 }, { "WITA", 28800, 0 }, { "WIT", 32400, 0 }, { "MYT", 28800, 0 }, { "BDT", 21600, 0 }, { "NPT", 20700, 0 },
    { "WAT", 3600, 0 }, { "CAT", 7200, 0 }, { "EAT", 10800, 0 }, { "SAST", 7200, 0 },
    { "A", 3600, 0 }, { "B", 7200, 0 }, { "C", 10800, 0 }, { "D", 14400, 0 }, { "E", 18000, 0 }, { "F", 21600, 0 },
    { "G", 25200, 0 }, { "H", 28800, 0 }, { "I", 32400, 0 }, { "K", 36000, 0 }, { "L", 39600, 0 }, { "M", 43200, 0 },
    { "N", -3600, 0 }, { "O", -7200, 0 }, { "P", -10800, 0 }, { "Q", -14400, 0 }, { "R", -18000, 0 }, { "S", -21600, 0 },
    { "T", -25200, 0 }, { "U", -28800, 0 }, { "V", -32400, 0 }, { "W", -36000, 0 }, { "X", -39600, 0 }, { "Y", -43200, 0 },
    { 0, 0, 0 }
};

static const char *days[7] = { "Sunday", "Monday", "Tuesday", "Wednesday", "Thursday", "Friday", "Saturday" };
static const char *abdays[7] = { "Sun", "Mon", "Tue", "Wed", "Thu", "Fri", "Sat" };
static const char *mons[12] = { "January", "February", "March", "April", "May", "June", "July", "August", "September",
                                "October", "November", "December" };
static const char *abmons[12] = { "Jan", "Feb", "Mar", "Apr", "May", "Jun", "Jul", "Aug", "Sep", "Oct", "Nov", "Dec" };
static const int mon_lengths[2][12] = { { 31, 28, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31 }, { 31, 29, 31, 30, 31, 30, 31, 31, 30, 31, 30, 31 } };

static int isleap(int y) { return (y % 4) == 0 && ((y % 100) != 0 || (y % 400) == 0); }
static int leaps_thru_end_of(int y) { return y >= 0 ? (y / 4 - y / 100 + y / 400) : -(leaps_thru_end_of(-(y + 1)) + 1); }

/* src/flb_strptime.c:790-812 */
static int conv_num(const unsigned char **buf, int *dest, int llim, int ulim)
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

/* src/flb_strptime.c:814-850 */
static int conv_num64(const unsigned char **buf, int64_t *dest, int64_t llim, int64_t ulim)
{
    int64_t result = 0, rulim = ulim;
    if (**buf < '0' || **buf > '9') return 0;
    do {
        if (result > 922337203685477580ll) return 0;
        result *= 10;
        if (result > 9223372036854775760ll) return 0;
        result += *(*buf)++ - '0';
        rulim /= 10;
    } while (rulim && **buf >= '0' && **buf <= '9');
    if (result < llim || result > ulim) return 0;
    *dest = result;
    return 1;
}

struct sp_state { int century, relyear, fields; };

static const unsigned char *find_string(const unsigned char *bp, int *tgt, const char *const *n1, int c)
{
    int i;
    for (i = 0; i < c; i++) {
        size_t len = strlen(n1[i]);
        if (strncasecmp(n1[i], (const char *) bp, len) == 0) { *tgt = i; return bp + len; }
    }
    return 0;
}

static const char *sp(const char *buf, const char *fmt, struct orc_tm *tm, struct sp_state *st, int initialize)
{
    unsigned char c;
    const unsigned char *bp, *ep;
    size_t len = 0;
    int i, offs, neg;
    static const char *const nast[4] = { "EST", "CST", "MST", "PST" };
    static const char *const nadt[4] = { "EDT", "CDT", "MDT", "PDT" };

    if (initialize) {
        st->century = 1900; st->relyear = -1; st->fields = 0;
        tm->gmtoff = 0; tm->isdst = -1;
    }
    bp = (const unsigned char *) buf;
    while ((c = (unsigned char) *fmt) != '\0') {
        if (isspace(c)) { while (isspace(*bp)) bp++; fmt++; continue; }
        if (*bp == '\0') return 0;
        if ((c = (unsigned char) *fmt++) != '%') goto literal;
again:
        switch (c = (unsigned char) *fmt++) {
        case '%':
literal:
            if (c != *bp++) return 0;
            break;
        case 'E': case 'O': goto again;              /* alternative modifiers: flag only */
        case 'c': if (!(bp = (const unsigned char *) sp((const char *) bp, "%a %b %e %H:%M:%S %Y", tm, st, 0))) return 0; break;
        case 'D': if (!(bp = (const unsigned char *) sp((const char *) bp, "%m/%d/%y", tm, st, 0))) return 0; break;
        case 'F': if (!(bp = (const unsigned char *) sp((const char *) bp, "%Y-%m-%d", tm, st, 0))) return 0; continue;
        case 'R': if (!(bp = (const unsigned char *) sp((const char *) bp, "%H:%M", tm, st, 0))) return 0; break;
        case 'r': if (!(bp = (const unsigned char *) sp((const char *) bp, "%I:%M:%S %p", tm, st, 0))) return 0; break;
        case 'T': if (!(bp = (const unsigned char *) sp((const char *) bp, "%H:%M:%S", tm, st, 0))) return 0; break;
        case 'X': if (!(bp = (const unsigned char *) sp((const char *) bp, "%H:%M:%S", tm, st, 0))) return 0; break;
        case 'x': if (!(bp = (const unsigned char *) sp((const char *) bp, "%m/%d/%y", tm, st, 0))) return 0; break;
        case 'A': case 'a':
            for (i = 0; i < 7; i++) {
                len = strlen(days[i]);
                if (strncasecmp(days[i], (const char *) bp, len) == 0) break;
                len = strlen(abdays[i]);
                if (strncasecmp(abdays[i], (const char *) bp, len) == 0) break;
            }
            if (i == 7) return 0;
            tm->wday = i; bp += len; st->fields |= F_WDAY;
            break;
        case 'B': case 'b': case 'h':
            for (i = 0; i < 12; i++) {
                len = strlen(mons[i]);
                if (strncasecmp(mons[i], (const char *) bp, len) == 0) break;
                len = strlen(abmons[i]);
                if (strncasecmp(abmons[i], (const char *) bp, len) == 0) break;
            }
            if (i == 12) return 0;
            tm->mon = i; bp += len; st->fields |= F_MON;
            break;
        case 'C': if (!conv_num(&bp, &i, 0, 99)) return 0; st->century = i * 100; break;
        case 'e': if (isspace(*bp)) bp++; /* FALLTHROUGH */
        case 'd': if (!conv_num(&bp, &tm->mday, 1, 31)) return 0; st->fields |= F_MDAY; break;
        case 'k': case 'H': if (!conv_num(&bp, &tm->hour, 0, 23)) return 0; break;
        case 'l': case 'I': if (!conv_num(&bp, &tm->hour, 1, 12)) return 0; break;
        case 'j': if (!conv_num(&bp, &tm->yday, 1, 366)) return 0; tm->yday--; st->fields |= F_YDAY; break;
        case 'M': if (!conv_num(&bp, &tm->min, 0, 59)) return 0; break;
        case 'm': if (!conv_num(&bp, &tm->mon, 1, 12)) return 0; tm->mon--; st->fields |= F_MON; break;
        case 'p':
            if (strncasecmp("AM", (const char *) bp, 2) == 0) {
                if (tm->hour > 12) return 0;
                else if (tm->hour == 12) tm->hour = 0;
                bp += 2; break;
            }
            if (strncasecmp("PM", (const char *) bp, 2) == 0) {
                if (tm->hour > 12) return 0;
                else if (tm->hour < 12) tm->hour += 12;
                bp += 2; break;
            }
            return 0;
        case 'S': if (!conv_num(&bp, &tm->sec, 0, 60)) return 0; break;
        case 's': {
            int64_t v;
            struct tm g;
            time_t t;
            if (!conv_num64(&bp, &v, 0, INT64_MAX)) return 0;
            t = (time_t) v;
            if (!gmtime_r(&t, &g)) return 0;
            tm->sec = g.tm_sec; tm->min = g.tm_min; tm->hour = g.tm_hour; tm->mday = g.tm_mday; tm->mon = g.tm_mon;
            tm->year = g.tm_year; tm->wday = g.tm_wday; tm->yday = g.tm_yday;
            tm->gmtoff = 0; tm->isdst = 0;
            st->fields = 0xffff;
            break;
        }
        case 'U': case 'W': if (!conv_num(&bp, &i, 0, 53)) return 0; break;
        case 'w': if (!conv_num(&bp, &tm->wday, 0, 6)) return 0; st->fields |= F_WDAY; break;
        case 'u': if (!conv_num(&bp, &i, 1, 7)) return 0; tm->wday = i % 7; st->fields |= F_WDAY; continue;
        case 'g': if (!conv_num(&bp, &i, 0, 99)) return 0; continue;
        case 'G': do bp++; while (isdigit(*bp)); continue;
        case 'V': if (!conv_num(&bp, &i, 0, 53)) return 0; continue;
        case 'Y':
            if (!conv_num(&bp, &i, 0, 9999)) return 0;
            st->relyear = -1; tm->year = i - 1900; st->fields |= F_YEAR;
            break;
        case 'y': if (!conv_num(&bp, &st->relyear, 0, 99)) return 0; break;
        case 'Z': {
            int found = 0, z;
            for (z = 0; zones[z].abbr; z++) {
                size_t al = strlen(zones[z].abbr);
                if (strncasecmp(zones[z].abbr, (const char *) bp, al) == 0 && !isalnum(bp[al])) {
                    tm->isdst = zones[z].dst; tm->gmtoff = zones[z].off; bp += al; found = 1;
                    break;
                }
            }
            if (!found) {
                /* falls back to the system's tzname[]; the harness runs with TZ=UTC */
                if (strncmp((const char *) bp, "GMT", 3) == 0 || strncmp((const char *) bp, "UTC", 3) == 0) {
                    tm->isdst = 0; tm->gmtoff = 0; bp += 3;
                }
                else return 0;
            }
            continue;
        }
        case 'z':
            while (isspace(*bp)) bp++;
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
            case 'Z': tm->isdst = 0; tm->gmtoff = 0; continue;
            case '+': neg = 0; break;
            case '-': neg = 1; break;
            default:
                --bp;
                ep = find_string(bp, &i, nast, 4);
                if (ep) { tm->gmtoff = (-5 - i) * 3600; tm->isdst = 0; bp = ep; continue; }
                ep = find_string(bp, &i, nadt, 4);
                if (ep) { tm->isdst = 1; tm->gmtoff = (-4 - i) * 3600; bp = ep; continue; }
                return 0;
            }
            if (!isdigit(bp[0]) || !isdigit(bp[1])) return 0;
            offs = ((bp[0] - '0') * 10 + (bp[1] - '0')) * 3600;
            bp += 2;
            if (*bp == ':') bp++;
            if (isdigit(*bp)) {
                offs += (*bp++ - '0') * 10 * 60;
                if (!isdigit(*bp)) return 0;
                offs += (*bp++ - '0') * 60;
            }
            if (neg) offs = -offs;
            tm->isdst = 0; tm->gmtoff = offs;
            continue;
        case 'n': case 't': while (isspace(*bp)) bp++; break;
        default: return 0;
        }
    }

    if (st->relyear != -1) {
        if (st->century == 1900) tm->year = st->relyear <= 68 ? st->relyear + 2000 - 1900 : st->relyear;
        else tm->year = st->relyear + st->century - 1900;
        st->fields |= F_YEAR;
    }
    if (st->fields & F_YEAR) {
        const int year = (int) ((unsigned) tm->year + 1900u);
        const int *ml = mon_lengths[isleap(year)];
        if (!(st->fields & F_YDAY) && (st->fields & F_MON) && (st->fields & F_MDAY)) {
            tm->yday = tm->mday - 1;
            for (i = 0; i < tm->mon; i++) tm->yday += ml[i];
            st->fields |= F_YDAY;
        }
        if (st->fields & F_YDAY) {
            int d = tm->yday;
            if (!(st->fields & F_WDAY)) {
                tm->wday = 4 + ((year - 1970) % 7) * (365 % 7) + leaps_thru_end_of(year - 1) - leaps_thru_end_of(1969) + tm->yday;
                tm->wday %= 7;
                if (tm->wday < 0) tm->wday += 7;
            }
            if (!(st->fields & F_MON)) { tm->mon = 0; while (tm->mon < 12 && d >= ml[tm->mon]) d -= ml[tm->mon++]; }
            if (!(st->fields & F_MDAY)) tm->mday = d + 1;
        }
    }
    return (const char *) bp;
}

/* The reference keeps century/relyear/fields in function statics that only a top-level call
 * resets; a second top-level call (the part after %L) re-initialises them.  `state` carries them. */
int orc_strptime(const char *buf, const char *fmt, struct orc_tm *tm, int64_t unused)
{
    struct sp_state st;
    const char *p = sp(buf, fmt, tm, &st, 1);
    (void) unused;
    return p ? (int) (p - buf) : -1;
}

int64_t orc_timegm(const struct orc_tm *tm)
{
    struct tm t;
    memset(&t, 0, sizeof(t));
    t.tm_sec = tm->sec; t.tm_min = tm->min; t.tm_hour = tm->hour; t.tm_mday = tm->mday; t.tm_mon = tm->mon;
    t.tm_year = tm->year; t.tm_wday = tm->wday; t.tm_yday = tm->yday; t.tm_isdst = tm->isdst;
    return (int64_t) timegm(&t);
}
