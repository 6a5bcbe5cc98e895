/* Hand-written stand-in for the flex-generated header of src/record_accessor/ra.l. */
#ifndef ORACLE_RA_LEX_H
#define ORACLE_RA_LEX_H
#ifndef ORACLE_RA_PARSER_H
typedef void *yyscan_t;
#endif
typedef struct ora_buf *YY_BUFFER_STATE;
int flb_ra_lex_init(yyscan_t *scanner);
int flb_ra_lex_destroy(yyscan_t scanner);
YY_BUFFER_STATE flb_ra__scan_string(const char *str, yyscan_t scanner);
void flb_ra__delete_buffer(YY_BUFFER_STATE b, yyscan_t scanner);
#endif
