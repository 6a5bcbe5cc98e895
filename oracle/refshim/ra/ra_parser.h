/* Hand-written stand-in for the bison-generated header of
 * src/record_accessor/ra.y (flex/bison are absent in this image). */
#ifndef ORACLE_RA_PARSER_H
#define ORACLE_RA_PARSER_H
struct flb_ra_parser;
typedef void *yyscan_t;
int flb_ra_parse(struct flb_ra_parser *rp, const char *query, void *scanner);
#endif
