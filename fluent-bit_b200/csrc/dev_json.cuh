/* dev_json.cuh -- JSON text -> msgpack on the device, one lane per document.
 *
 * Mirrors what the reference's default JSON back end produces
 * (flb_pack_json_recs -> pack_json_to_msgpack_yyjson, src/flb_pack.c:389-508, with
 * yyjson 0.12 read flags STOP_WHEN_DONE | INSITU | ALLOW_INVALID_UNICODE |
 * REPLACE_INVALID_UNICODE, :441-442) and yyjson_val_to_msgpack (:317-373):
 *   object -> map (count = members, duplicates kept, order kept), array -> array,
 *   string -> str (escapes decoded as yyjson's read_str_opt / read_uni_esc do),
 *   non-negative integer -> uint minimal, negative integer -> int minimal,
 *   integer overflow / fraction / exponent -> float64 (correctly rounded),
 *   true/false/null.
 * Grammar is RFC 8259 as yyjson enforces it without extension flags: no trailing
 * commas, no comments, no leading zeros, no NaN/Inf (an overflowing real is an error),
 * whitespace = space \t \n \r.  Raw control characters and invalid UTF-8 inside strings
 * are accepted and copied (ALLOW_INVALID_UNICODE); unknown escapes are errors.
 *
 * Decimal -> double: Clinger's exact fast path, then Eisel-Lemire with the 128-bit
 * power-of-five table (dev_pow5_table.h).  When Eisel-Lemire cannot decide (a few
 * halfway cases, >19 significant digits straddling a boundary) DJ_E_FLOAT is reported
 * instead of guessing.
 */
#ifndef FLBGPU_DEV_JSON_CUH
#define FLBGPU_DEV_JSON_CUH

#include <stdint.h>
#include "dev_msgpack.cuh"

#ifdef __CUDA_ARCH__
#define DJ_TABLE_QUAL __device__ static const
#else
#define DJ_TABLE_QUAL static const
#endif
#include "dev_pow5_table.h"

#define DJ_MAX_DEPTH 31        /* nesting the reference's msgpack_unpack_next() still accepts */
#define DJ_E_FLOAT   1
/* A value nested this deep still parses, but as part of a record body -- one or two containers further down -- it is at or
 * near the depth where msgpack-c's unpacker gives up (MSGPACK_EMBED_STACK_SIZE, dev_msgpack.cuh: mp_skip_lim): every filter
 * that decodes the parser's result then stops at this record.  A chain that hands fields from filter to filter cannot show that,
 * so the transcoder reports it and the chain is run filter by filter for this call (runtime.c: chain_do_one_by_one). */
#define DJ_E_DEEP    2
#define DJ_DEEP_HINT 28

FLB_HD int dj_ws(uint32_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
FLB_HD int dj_hex(uint32_t c)
{
    if (c >= '0' && c <= '9') return (int) c - '0';
    if (c >= 'a' && c <= 'f') return (int) c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return (int) c - 'A' + 10;
    return -1;
}

FLB_HD void dj_mul64(uint64_t a, uint64_t b, uint64_t *hi, uint64_t *lo)
{
#ifdef __CUDA_ARCH__
    *lo = a * b;
    *hi = __umul64hi(a, b);
#else
    unsigned __int128 p = (unsigned __int128) a * b;
    *lo = (uint64_t) p;
    *hi = (uint64_t) (p >> 64);
#endif
}

FLB_HD int dj_clz64(uint64_t x)
{
#ifdef __CUDA_ARCH__
    return __clzll((long long) x);
#else
    return __builtin_clzll(x);
#endif
}

/* w * 10^q -> IEEE double bits.  Returns 0 ok, 1 overflow (infinity), 2 undecided. */
FLB_HDN int dj_eisel_lemire(uint64_t w, int64_t q, uint64_t *bits)
{
    uint64_t hi, lo, mant;
    int lz, upperbit;
    int64_t exponent;
    if (w == 0 || q < DJ_POW5_MIN) { *bits = 0; return 0; }
    if (q > DJ_POW5_MAX) return 1;
    lz = dj_clz64(w);
    w <<= lz;
    {
        const uint64_t t_hi = dj_pow5[2 * (q - DJ_POW5_MIN)], t_lo = dj_pow5[2 * (q - DJ_POW5_MIN) + 1];
        dj_mul64(w, t_hi, &hi, &lo);
        if ((hi & 0x1FF) == 0x1FF) {
            uint64_t shi, slo;
            dj_mul64(w, t_lo, &shi, &slo);
            lo += shi;
            if (shi > lo) hi++;
            (void) slo;        /* "the computed product is always sufficient" (Lemire 2021, section 6) */
        }
    }
    upperbit = (int) (hi >> 63);
    mant = hi >> (upperbit + 64 - 52 - 3);
    /* power2 = power(q) + upperbit - lz - minimum_exponent, power(q) = ((217706*q) >> 16) + 63 */
    exponent = ((((int64_t) (152170 + 65536) * q) >> 16) + 63) + upperbit - lz + 1023;
    if (exponent <= 0) {                                               /* subnormal */
        if (-exponent + 1 >= 64) { *bits = 0; return 0; }
        mant >>= -exponent + 1;
        mant += (mant & 1);
        mant >>= 1;
        *bits = mant;                                                  /* bit 52, when set, is exponent 1 */
        return 0;
    }
    if (lo <= 1 && q >= -4 && q <= 23 && (mant & 3) == 1) {
        if ((mant << (upperbit + 64 - 52 - 3)) == hi) mant &= ~(uint64_t) 1;
    }
    mant += (mant & 1);
    mant >>= 1;
    if (mant >= ((uint64_t) 2 << 52)) { mant = (uint64_t) 1 << 52; exponent++; }
    mant &= ~((uint64_t) 1 << 52);
    if (exponent >= 0x7FF) return 1;
    *bits = mant | ((uint64_t) exponent << 52);
    return 0;
}

/* ---- exact decision for the rare number whose first 19 digits leave two candidate doubles ----
 * (more than 19 significant digits and the Eisel-Lemire results for w and w+1 differ).  The decimal
 * value D x 10^E is compared, in exact integer arithmetic, with the midpoint (2M+1) x 2^(k-1) of the two
 * adjacent candidates M x 2^k and (M+1) x 2^k; ties go to the even mantissa, as every correctly
 * rounding reader (yyjson included) does.  Returns 0 = lower candidate, 1 = upper, -1 = cannot tell
 * (more digits / exponent than the 160-limb integers hold). */
#define DJ_BIG_LIMBS 160
#define DJ_BIG_DIGITS 780      /* every midpoint of two adjacent doubles has at most 767 significant digits */
struct dj_big { uint32_t n; uint32_t v[DJ_BIG_LIMBS]; };

FLB_HD int djb_mul_small(struct dj_big *b, uint32_t m, uint32_t add)
{
    uint64_t carry = add;
    uint32_t i;
    for (i = 0; i < b->n; i++) { carry += (uint64_t) b->v[i] * m; b->v[i] = (uint32_t) carry; carry >>= 32; }
    if (carry) { if (b->n >= DJ_BIG_LIMBS) return -1; b->v[b->n++] = (uint32_t) carry; }
    return 0;
}
FLB_HD int djb_mul_pow5(struct dj_big *b, uint32_t e)
{
    while (e >= 13) { if (djb_mul_small(b, 1220703125u, 0)) return -1; e -= 13; }
    while (e--) if (djb_mul_small(b, 5u, 0)) return -1;
    return 0;
}
FLB_HD int djb_shl(struct dj_big *b, uint32_t bits)
{
    const uint32_t words = bits >> 5, sh = bits & 31;
    int i;
    if (b->n == 0) return 0;
    if (b->n + words + 1 > DJ_BIG_LIMBS) return -1;
    if (sh) {
        uint32_t carry = 0, k;
        for (k = 0; k < b->n; k++) { const uint32_t x = b->v[k]; b->v[k] = (x << sh) | carry; carry = x >> (32 - sh); }
        if (carry) b->v[b->n++] = carry;
    }
    if (words) {
        for (i = (int) b->n - 1; i >= 0; i--) b->v[i + words] = b->v[i];
        for (i = 0; i < (int) words; i++) b->v[i] = 0;
        b->n += words;
    }
    return 0;
}
FLB_HD int djb_cmp(const struct dj_big *a, const struct dj_big *b)
{
    int i;
    if (a->n != b->n) return a->n > b->n ? 1 : -1;
    for (i = (int) a->n - 1; i >= 0; i--) if (a->v[i] != b->v[i]) return a->v[i] > b->v[i] ? 1 : -1;
    return 0;
}

FLB_HDN int dj_big_decide(const uint8_t *s, int n, int pos, uint64_t bits_lo)
{
    struct dj_big L, R;
    int p = pos, nd = 0, sticky = 0, seen_nz = 0;
    int64_t E = 0;
    uint64_t M;
    int k, c;
    L.n = 0;
    if (p < n && s[p] == '-') p++;
    for (; p < n && s[p] >= '0' && s[p] <= '9'; p++) {
        const uint32_t d = s[p] - '0';
        if (!seen_nz && !d) continue;
        seen_nz = 1;
        if (nd < DJ_BIG_DIGITS) { if (L.n == 0) { if (d) { L.v[0] = d; L.n = 1; } } else if (djb_mul_small(&L, 10u, d)) return -1; nd++; }
        else { E++; if (d) sticky = 1; }
    }
    if (p < n && s[p] == '.') {
        for (p++; p < n && s[p] >= '0' && s[p] <= '9'; p++) {
            const uint32_t d = s[p] - '0';
            if (!seen_nz && !d) { E--; continue; }
            seen_nz = 1;
            if (nd < DJ_BIG_DIGITS) { if (L.n == 0) { L.v[0] = d; L.n = 1; } else if (djb_mul_small(&L, 10u, d)) return -1; nd++; E--; }
            else if (d) sticky = 1;
        }
    }
    if (p < n && (s[p] == 'e' || s[p] == 'E')) {
        int eneg = 0;
        int64_t ev = 0;
        p++;
        if (p < n && (s[p] == '+' || s[p] == '-')) { eneg = s[p] == '-'; p++; }
        for (; p < n && s[p] >= '0' && s[p] <= '9'; p++) if (ev < 100000) ev = ev * 10 + (s[p] - '0');
        E += eneg ? -ev : ev;
    }
    if (L.n == 0 || E > 400 || E < -1200) return -1;
    /* lower candidate M x 2^k */
    {
        const uint32_t ef = (uint32_t) ((bits_lo >> 52) & 0x7ff);
        M = bits_lo & (((uint64_t) 1 << 52) - 1);
        if (ef) { M |= (uint64_t) 1 << 52; k = (int) ef - 1075; } else k = -1074;
    }
    {
        const uint64_t m2 = 2 * M + 1;
        R.v[0] = (uint32_t) m2; R.v[1] = (uint32_t) (m2 >> 32); R.n = R.v[1] ? 2 : 1;
    }
    if (E >= 0) { if (djb_mul_pow5(&L, (uint32_t) E) || djb_shl(&L, (uint32_t) E)) return -1; }
    else { if (djb_mul_pow5(&R, (uint32_t) -E) || djb_shl(&R, (uint32_t) -E)) return -1; }
    if (k - 1 >= 0) { if (djb_shl(&R, (uint32_t) (k - 1))) return -1; }
    else { if (djb_shl(&L, (uint32_t) (1 - k))) return -1; }
    c = djb_cmp(&L, &R);
    if (c > 0) return 1;
    if (c < 0) return 0;
    if (sticky) return 1;
    return (M & 1) ? 1 : 0;                       /* exactly half way: to the even mantissa */
}

/* Parse the JSON number at s[pos].  Returns the position after it or -1.
 * kind: 0 uint (u), 1 sint (value in u as two's complement), 2 real (bits in u). */
/* exact powers of ten (Clinger fast path) */
DJ_TABLE_QUAL double dj_p10[23] = { 1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15,
                                    1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22 };

FLB_HDN int dj_number(const uint8_t *s, int n, int pos, int *kind, uint64_t *u, uint32_t *err)
{
    int neg = 0, p = pos, nd = 0, dropped = 0, is_real = 0, int_overflow = 0, truncated = 0;
    uint64_t w = 0, iw = 0;
    int64_t exp10 = 0;
    const double *p10 = dj_p10;
    if (p < n && s[p] == '-') { neg = 1; p++; }
    if (p >= n || s[p] < '0' || s[p] > '9') return -1;
    if (s[p] == '0') {
        p++;
        if (p < n && s[p] >= '0' && s[p] <= '9') return -1;           /* leading zero */
    }
    else {
        while (p < n && s[p] >= '0' && s[p] <= '9') {
            uint32_t d = s[p] - '0';
            if (iw > (0xffffffffffffffffull - d) / 10) int_overflow = 1; else iw = iw * 10 + d;
            if (nd < 19) { w = w * 10 + d; nd++; } else { dropped++; if (d) truncated = 1; }
            p++;
        }
    }
    if (p < n && s[p] == '.') {
        int fd = 0;
        p++;
        while (p < n && s[p] >= '0' && s[p] <= '9') {
            uint32_t d = s[p] - '0';
            if (nd < 19) { if (w || d) { w = w * 10 + d; nd++; } exp10--; }
            else if (d) truncated = 1;
            fd++; p++;
        }
        if (fd == 0) return -1;
        is_real = 1;
    }
    if (p < n && (s[p] == 'e' || s[p] == 'E')) {
        int eneg = 0, ed = 0;
        int64_t ev = 0;
        p++;
        if (p < n && (s[p] == '+' || s[p] == '-')) { eneg = s[p] == '-'; p++; }
        while (p < n && s[p] >= '0' && s[p] <= '9') { if (ev < 100000) ev = ev * 10 + (s[p] - '0'); ed++; p++; }
        if (ed == 0) return -1;
        exp10 += eneg ? -ev : ev;
        is_real = 1;
    }
    if (!is_real) {
        if (!int_overflow) {
            if (!neg) { *kind = 0; *u = iw; return p; }
            if (iw <= 9223372036854775808ull) { *kind = 1; *u = (uint64_t) (0 - iw); return p; }
        }
    }
    exp10 += dropped;
    *kind = 2;
    {
        uint64_t bits = 0;
        if (w == 0) bits = 0;
        else if (!truncated && w <= ((uint64_t) 1 << 53) && exp10 >= -22 && exp10 <= 22) {
            union { double d; uint64_t u; } cv;
            cv.d = (double) w;
            if (exp10 >= 0) cv.d = cv.d * p10[exp10]; else cv.d = cv.d / p10[-exp10];
            bits = cv.u;
        }
        else {
            int r = dj_eisel_lemire(w, exp10, &bits);
            if (r == 1) return -1;                                     /* infinity: yyjson rejects the document */
            if (r == 0 && truncated) {
                uint64_t b2 = 0;
                int r2 = dj_eisel_lemire(w + 1, exp10, &b2);
                if (r2 != 0 || b2 != bits || (bits >> 52) == 0) {
                    /* the dropped digits matter (or the result is subnormal, where the 128-bit product is
                     * rounded twice): decide exactly between the two neighbours */
                    const int up = (r2 == 1 || b2 == bits + 1 || b2 == bits) ? dj_big_decide(s, n, pos, bits) : -1;
                    if (up < 0) r = 2;
                    else if (up == 1) {
                        if (r2 == 1) return -1;                        /* rounds to infinity: rejected like any overflow */
                        bits = bits + 1;
                    }
                }
            }
            if (r == 2) { *err |= DJ_E_FLOAT; bits = 0; }
        }
        if (neg) bits |= (uint64_t) 1 << 63;
        *u = bits;
    }
    return p;
}

/* strtod() of the C locale for the texts `Types key:float` (flb_parser.c:flb_parser_typecast -> atof)
 * and log_to_metrics' sscanf("%lf") hand it: white space, sign, digits [. digits] [e[+-]digits], "inf" /
 * "infinity" / "nan"; the longest valid prefix converts, nothing convertible gives +0.0.  The result is
 * correctly rounded (nearest, ties to even) like glibc's: Clinger's exact products, Eisel-Lemire on the
 * first 19 digits, and the big-integer comparison when the digits beyond them decide.  *ok = 0 for the
 * forms that are not restated (hex floats, nan(payload)) and for texts beyond dj_big_decide's reach. */
FLB_HDN uint64_t dj_strtod(const uint8_t *s, int n, int *ok)
{
    int p = 0, neg = 0, nd = 0, dropped = 0, truncated = 0, seen = 0, start;
    uint64_t w = 0, bits = 0;
    int64_t exp10 = 0;
    const uint64_t sign = (uint64_t) 1 << 63, inf = (uint64_t) 0x7ff << 52;
    *ok = 1;
    while (p < n && (s[p] == ' ' || (s[p] >= 9 && s[p] <= 13))) p++;
    if (p < n && (s[p] == '+' || s[p] == '-')) { neg = s[p] == '-'; p++; }
    start = p;
    if (p + 2 < n && (s[p] | 0x20) == 'i' && (s[p + 1] | 0x20) == 'n' && (s[p + 2] | 0x20) == 'f') return neg ? inf | sign : inf;
    if (p + 2 < n && (s[p] | 0x20) == 'n' && (s[p + 1] | 0x20) == 'a' && (s[p + 2] | 0x20) == 'n') {
        if (p + 3 < n && s[p + 3] == '(') *ok = 0;
        return (inf | (uint64_t) 1 << 51) | (neg ? sign : 0);
    }
    if (p + 1 < n && s[p] == '0' && (s[p + 1] | 0x20) == 'x') { *ok = 0; return 0; }
    for (; p < n && s[p] >= '0' && s[p] <= '9'; p++) {
        const uint32_t d = s[p] - '0';
        seen = 1;
        if (w || d) { if (nd < 19) { w = w * 10 + d; nd++; } else { dropped++; if (d) truncated = 1; } }
    }
    if (p < n && s[p] == '.' && (seen || (p + 1 < n && s[p + 1] >= '0' && s[p + 1] <= '9'))) {
        for (p++; p < n && s[p] >= '0' && s[p] <= '9'; p++) {
            const uint32_t d = s[p] - '0';
            seen = 1;
            if (w || d) { if (nd < 19) { w = w * 10 + d; nd++; exp10--; } else if (d) truncated = 1; }
            else exp10--;
        }
    }
    if (!seen) return 0;                                               /* no conversion */
    if (p < n && (s[p] | 0x20) == 'e') {
        int q = p + 1, eneg = 0;
        int64_t ev = 0;
        if (q < n && (s[q] == '+' || s[q] == '-')) { eneg = s[q] == '-'; q++; }
        if (q < n && s[q] >= '0' && s[q] <= '9') {
            for (; q < n && s[q] >= '0' && s[q] <= '9'; q++) if (ev < 100000) ev = ev * 10 + (s[q] - '0');
            exp10 += eneg ? -ev : ev;
        }
    }
    exp10 += dropped;
    if (w == 0) bits = 0;
    else if (!truncated && w <= ((uint64_t) 1 << 53) && exp10 >= -22 && exp10 <= 22) {
        union { double d; uint64_t u; } cv;
        cv.d = (double) w;
        if (exp10 >= 0) cv.d = cv.d * dj_p10[exp10]; else cv.d = cv.d / dj_p10[-exp10];
        bits = cv.u;
    }
    else {
        const int r = dj_eisel_lemire(w, exp10, &bits);
        if (r == 1) bits = inf;                                        /* HUGE_VAL */
        else if (truncated) {
            uint64_t b2 = 0;
            const int r2 = dj_eisel_lemire(w + 1, exp10, &b2);
            if (r2 != 0 || b2 != bits || (bits >> 52) == 0) {
                const int up = (r2 == 1 || b2 == bits + 1 || b2 == bits) ? dj_big_decide(s, n, start, bits) : -1;
                if (up < 0) { *ok = 0; bits = 0; }
                else if (up == 1) bits = bits + 1;                     /* the largest finite double + 1 ulp is infinity's pattern */
            }
        }
    }
    return neg ? bits | sign : bits;
}

/* printf("%f") of a double, as glibc prints it: the exact binary value times 10^6 rounded to an integer
 * (nearest, ties to even), six digits after the point, "inf" / "nan" with their sign.  At most `cap`
 * bytes are written (snprintf's truncation); returns how many.  log_to_metrics renders FLOAT label
 * values this way (log_to_metrics.c:1028-1031). */
FLB_HD uint32_t djb_divmod_small(struct dj_big *b, uint32_t d)
{
    uint64_t rem = 0;
    int i;
    for (i = (int) b->n - 1; i >= 0; i--) {
        const uint64_t cur = (rem << 32) | b->v[i];
        b->v[i] = (uint32_t) (cur / d);
        rem = cur % d;
    }
    while (b->n && b->v[b->n - 1] == 0) b->n--;
    return (uint32_t) rem;
}

FLB_HDN uint32_t dj_fmt_f6(uint64_t bits, uint8_t *dst, uint32_t cap)
{
    struct dj_big N;
    uint8_t tmp[336];                                  /* 309 integer digits + 6 + slack, least significant first */
    uint32_t nt = 0, n = 0, i;
    const uint32_t ef = (uint32_t) ((bits >> 52) & 0x7ff);
    uint64_t M = bits & (((uint64_t) 1 << 52) - 1);
    int k;
#define DJ_FPUT(c) do { if (n < cap) dst[n++] = (uint8_t) (c); } while (0)
    if (bits >> 63) DJ_FPUT('-');
    if (ef == 0x7ff) {
        if (M) { DJ_FPUT('n'); DJ_FPUT('a'); DJ_FPUT('n'); } else { DJ_FPUT('i'); DJ_FPUT('n'); DJ_FPUT('f'); }
        return n;
    }
    if (ef) { M |= (uint64_t) 1 << 52; k = (int) ef - 1075; } else k = -1074;
    N.n = 0;
    if (M) {
        if (k >= 0) {                                   /* an integer: M * 2^k * 10^6, exactly */
            N.v[0] = (uint32_t) M; N.v[1] = (uint32_t) (M >> 32); N.n = N.v[1] ? 2 : 1;
            djb_shl(&N, (uint32_t) k);
            djb_mul_small(&N, 1000000u, 0);
        }
        else {                                          /* X = M * 10^6 < 2^73;  N = round(X / 2^s) */
            const uint32_t sft = (uint32_t) -k;
            uint64_t hi, lo, qhi = 0, qlo = 0;
            int up = 0;
            dj_mul64(M, 1000000ull, &hi, &lo);
            if (sft <= 74) {
                uint64_t rhi, rlo, hhi, hlo;            /* remainder and half of the divisor */
                if (sft >= 64) { qlo = hi >> (sft - 64); qhi = 0; rhi = sft == 64 ? 0 : hi & (((uint64_t) 1 << (sft - 64)) - 1); rlo = lo; }
                else { qlo = (lo >> sft) | (hi << (64 - sft)); qhi = hi >> sft; rhi = 0; rlo = lo & (((uint64_t) 1 << sft) - 1); }
                if (sft - 1 >= 64) { hhi = (uint64_t) 1 << (sft - 1 - 64); hlo = 0; } else { hhi = 0; hlo = (uint64_t) 1 << (sft - 1); }
                if (rhi > hhi || (rhi == hhi && rlo > hlo)) up = 1;
                else if (rhi == hhi && rlo == hlo) up = (int) (qlo & 1);
            }                                           /* else X < 2^73 <= half of 2^s: rounds to zero */
            if (up) { qlo++; if (qlo == 0) qhi++; }
            N.v[0] = (uint32_t) qlo; N.v[1] = (uint32_t) (qlo >> 32); N.v[2] = (uint32_t) qhi; N.v[3] = (uint32_t) (qhi >> 32);
            N.n = 4;
            while (N.n && N.v[N.n - 1] == 0) N.n--;
        }
    }
    while (N.n) {                                       /* nine digits at a time */
        uint32_t r = djb_divmod_small(&N, 1000000000u), j;
        for (j = 0; j < 9 && (N.n || r); j++) { tmp[nt++] = (uint8_t) ('0' + r % 10); r /= 10; }
    }
    while (nt < 7) tmp[nt++] = '0';                     /* at least "0" before the point */
    for (i = nt; i > 6; i--) DJ_FPUT(tmp[i - 1]);
    DJ_FPUT('.');
    for (i = 6; i > 0; i--) DJ_FPUT(tmp[i - 1]);
#undef DJ_FPUT
    return n;
}

/* Decode the string whose opening quote is at s[pos].  Writes the decoded bytes to o
 * (when o != NULL), returns the position after the closing quote or -1; *olen = length. */
FLB_HDN int dj_string(const uint8_t *s, int n, int pos, uint8_t *o, uint32_t *olen)
{
    int p = pos + 1;
    uint32_t k = 0;
#define DJ_PUT(b) do { if (o) o[k] = (uint8_t) (b); k++; } while (0)
    for (;;) {
        uint32_t c;
        if (p >= n) return -1;                                         /* unclosed string */
        c = s[p];
        if (c == '"') { *olen = k; return p + 1; }
        if (c != '\\') { DJ_PUT(c); p++; continue; }
        if (p + 1 >= n) return -1;
        c = s[p + 1];
        switch (c) {
        case '"': DJ_PUT('"'); p += 2; continue;
        case '\\': DJ_PUT('\\'); p += 2; continue;
        case '/': DJ_PUT('/'); p += 2; continue;
        case 'b': DJ_PUT(8); p += 2; continue;
        case 'f': DJ_PUT(12); p += 2; continue;
        case 'n': DJ_PUT(10); p += 2; continue;
        case 'r': DJ_PUT(13); p += 2; continue;
        case 't': DJ_PUT(9); p += 2; continue;
        case 'u': break;
        default: return -1;                                            /* invalid escaped sequence */
        }
        {
            /* read_uni_esc(), lib/yyjson-0.12.0/src/yyjson.c:4659-4820 with REPLACE_INVALID_UNICODE */
            int q = p + 2, cnt = 0, h;
            uint32_t hi = 0, lo = 0, uni;
            while (cnt < 4 && q + cnt < n && (h = dj_hex(s[q + cnt])) >= 0) { hi = hi * 16 + (uint32_t) h; cnt++; }
            if (cnt < 4) {
                uint32_t ch = (q + cnt < n) ? s[q + cnt] : 0;
                int i;
                DJ_PUT('\\'); DJ_PUT('u');
                for (i = 0; i < cnt; i++) DJ_PUT(s[q + i]);
                p = q + cnt;
                if (ch && ch != '"' && ch != '\'') p++;
                continue;
            }
            q += 4;
            if ((hi & 0xF800) != 0xD800) {
                if (hi >= 0x800) { DJ_PUT(0xE0 | (hi >> 12)); DJ_PUT(0x80 | ((hi >> 6) & 0x3F)); DJ_PUT(0x80 | (hi & 0x3F)); }
                else if (hi >= 0x80) { DJ_PUT(0xC0 | (hi >> 6)); DJ_PUT(0x80 | (hi & 0x3F)); }
                else DJ_PUT(hi);
                p = q;
                continue;
            }
            if ((hi & 0xFC00) != 0xD800) { DJ_PUT(0xEF); DJ_PUT(0xBF); DJ_PUT(0xBD); p = q; continue; }   /* lone low surrogate */
            if (!(q + 1 < n && s[q] == '\\' && s[q + 1] == 'u')) { DJ_PUT(0xEF); DJ_PUT(0xBF); DJ_PUT(0xBD); p = q; continue; }
            cnt = 0;
            while (cnt < 4 && q + 2 + cnt < n && (h = dj_hex(s[q + 2 + cnt])) >= 0) { lo = lo * 16 + (uint32_t) h; cnt++; }
            if (cnt < 4) { DJ_PUT(0xEF); DJ_PUT(0xBF); DJ_PUT(0xBD); p = q + 2 + cnt; continue; }
            if ((lo & 0xFC00) != 0xDC00) { DJ_PUT(0xEF); DJ_PUT(0xBF); DJ_PUT(0xBD); p = q + 6; continue; }
            uni = (((hi - 0xD800) << 10) | (lo - 0xDC00)) + 0x10000;
            DJ_PUT(0xF0 | (uni >> 18)); DJ_PUT(0x80 | ((uni >> 12) & 0x3F)); DJ_PUT(0x80 | ((uni >> 6) & 0x3F)); DJ_PUT(0x80 | (uni & 0x3F));
            p = q + 6;
        }
    }
#undef DJ_PUT
}

/* members of the container whose opening bracket is at s[pos] (pairs for objects);
 * -1 when it never closes.  Purely structural: the real parse validates -- a wrong count
 * cannot survive it (the member loop expects exactly `count` values before the closing
 * bracket).  exact=0 skips strings by looking only at backslashes, which disagrees with
 * yyjson's lexer in its invalid-\u corner; dj_parse_record() retries with exact=1 (the
 * real string lexer) whenever the fast attempt fails. */
FLB_HD int dj_count(const uint8_t *s, int n, int pos, int exact)
{
    int p = pos + 1, depth = 0, count = 0, seen = 0;
    while (p < n) {
        uint32_t c = s[p];
        if (c == '"' && !exact) {
            p++;
            while (p < n && s[p] != '"') p += (s[p] == '\\') ? 2 : 1;
            if (p >= n) return -1;
            p++; seen = 1;
            continue;
        }
        if (c == '"') { uint32_t sl; p = dj_string(s, n, p, 0, &sl); if (p < 0) return -1; seen = 1; continue; }   /* the same lexer the real parse uses (\\u quirks) */
        if (c == '[' || c == '{') { depth++; seen = 1; }
        else if (c == ']' || c == '}') { if (depth == 0) return count + (seen ? 1 : 0); depth--; }
        else if (c == ',' && depth == 0) { count++; seen = 0; }
        else if (!dj_ws(c)) seen = 1;
        p++;
    }
    return -1;
}

/* string at s[pos] -> msgpack str at o+k; returns the position after it or -1.  Strings
 * without a backslash (nearly all) are located by one scan and copied once. */
FLB_HD int dj_emit_string(const uint8_t *s, int n, int pos, uint8_t *o, uint32_t *k)
{
    int p = pos + 1;
    uint32_t sl = 0;
    while (p < n && s[p] != '"' && s[p] != '\\') p++;
    if (p >= n) return -1;
    if (s[p] == '"') {
        sl = (uint32_t) (p - pos - 1);
        if (o) { uint32_t h = mp_put_str_hdr(o + *k, sl); mp_copy(o + *k + h, s + pos + 1, sl); }
        *k += mp_str_hdr_size(sl) + sl;
        return p + 1;
    }
    p = dj_string(s, n, pos, 0, &sl);
    if (p < 0) return -1;
    if (o) { uint32_t h = mp_put_str_hdr(o + *k, sl); dj_string(s, n, pos, o + *k + h, &sl); }
    *k += mp_str_hdr_size(sl) + sl;
    return p;
}

/* One JSON value at s[pos] -> msgpack at o (NULL = only measure and validate).
 * Returns the position after the value, or -1 on any syntax error.  *olen = msgpack bytes. */
FLB_HDN int dj_value(const uint8_t *s, int n, int pos, uint8_t *o, uint32_t *olen, uint32_t *err, int exact)
{
    uint32_t k = 0, rem[DJ_MAX_DEPTH + 1];
    uint8_t isobj[DJ_MAX_DEPTH + 1];
    int sp = 0, p = pos;

    for (;;) {
        /* ---- a value starts at p (whitespace already skipped by the caller of this state) ---- */
        uint32_t c;
        while (p < n && dj_ws(s[p])) p++;
        if (p >= n) return -1;
        c = s[p];
        if (c == '{' || c == '[') {
            int cnt = dj_count(s, n, p, exact);
            if (cnt < 0) return -1;
            if (sp > DJ_MAX_DEPTH) return -1;              /* a 33rd open container, empty or not (unpack_template.h:140-144) */
            if (sp >= DJ_DEEP_HINT) *err |= DJ_E_DEEP;
            if (c == '{') { if (o) mp_put_map_hdr(o + k, (uint32_t) cnt); }
            else if (o) mp_put_array_hdr(o + k, (uint32_t) cnt);
            k += mp_cnt_hdr_size((uint32_t) cnt);
            p++;
            if (cnt == 0) {
                while (p < n && dj_ws(s[p])) p++;
                if (p >= n || s[p] != (c == '{' ? '}' : ']')) return -1;
                p++;
                goto value_done;
            }
            rem[sp] = (uint32_t) cnt; isobj[sp] = (c == '{'); sp++;
            goto next_member;
        }
        if (c == '"') {
            p = dj_emit_string(s, n, p, o, &k);
            if (p < 0) return -1;
            goto value_done;
        }
        if (c == '-' || (c >= '0' && c <= '9')) {
            int kind = 0;
            uint64_t u = 0;
            int e = dj_number(s, n, p, &kind, &u, err);
            if (e < 0) return -1;
            if (kind == 0) { if (o) mp_put_uint(o + k, u); k += mp_uint_size(u); }
            else if (kind == 1) { if (o) mp_put_int(o + k, (int64_t) u); k += mp_int_size((int64_t) u); }
            else { if (o) { o[k] = 0xcb; mp_put_be64(o + k + 1, u); } k += 9; }
            p = e;
            goto value_done;
        }
        if (c == 't' && p + 4 <= n && s[p + 1] == 'r' && s[p + 2] == 'u' && s[p + 3] == 'e') { if (o) o[k] = 0xc3; k++; p += 4; goto value_done; }
        if (c == 'f' && p + 5 <= n && s[p + 1] == 'a' && s[p + 2] == 'l' && s[p + 3] == 's' && s[p + 4] == 'e') { if (o) o[k] = 0xc2; k++; p += 5; goto value_done; }
        if (c == 'n' && p + 4 <= n && s[p + 1] == 'u' && s[p + 2] == 'l' && s[p + 3] == 'l') { if (o) o[k] = 0xc0; k++; p += 4; goto value_done; }
        return -1;

value_done:
        if (sp == 0) { *olen = k; return p; }
        rem[sp - 1]--;
        while (p < n && dj_ws(s[p])) p++;
        if (p >= n) return -1;
        if (rem[sp - 1] == 0) {
            if (s[p] != (isobj[sp - 1] ? '}' : ']')) return -1;
            p++;
            sp--;
            goto value_done;
        }
        if (s[p] != ',') return -1;
        p++;
next_member:
        if (isobj[sp - 1]) {
            while (p < n && dj_ws(s[p])) p++;
            if (p >= n || s[p] != '"') return -1;
            p = dj_emit_string(s, n, p, o, &k);
            if (p < 0) return -1;
            while (p < n && dj_ws(s[p])) p++;
            if (p >= n || s[p] != ':') return -1;
            p++;
        }
        /* loop: parse the member value */
    }
}

/* flb_pack_json_recs() for one line: exactly one document, which must be an object.
 * Returns 1 and the msgpack map at o (length *olen), or 0. */
FLB_HDN int dj_parse_record(const uint8_t *s, int n, uint8_t *o, uint32_t *olen, uint32_t *err, int *consumed)
{
    int p = 0, e;
    uint32_t dummy = 0, derr = 0;
    while (p < n && dj_ws(s[p])) p++;
    if (p >= n) return 0;
    if (s[p] != '{') {                                                 /* only objects can succeed; still must be valid to count as a record */
        return 0;
    }
    e = dj_value(s, n, p, o, olen, err, 0);
    if (e < 0) { *err = 0; e = dj_value(s, n, p, o, olen, err, 1); }
    if (e < 0) return 0;
    if (s[p] != '{') return 0;                                         /* root must be a map (src/flb_parser_json.c:70-85) */
    /* a second parsable document means records == 2 -> rejected; trailing junk is fine */
    p = e;
    while (p < n && dj_ws(s[p])) p++;
    if (p < n && dj_value(s, n, p, 0, &dummy, &derr, 1) >= 0) return 0;
    *consumed = p;                                                     /* pack_json_to_msgpack_yyjson(): start - insitu_buf when the loop ends */
    return 1;
}

#endif
