/* Symbols the reference sources reference but that are never reached on the
 * parse->filter path; they abort loudly if called.  Part of oracle/_ref (test
 * infrastructure only).  Deliberately compiled without the reference headers so
 * the K&R-style definitions do not clash with the prototypes. */
#include <stdio.h>
#include <stdlib.h>
#define DEAD(name) void *name() { fprintf(stderr, "oracle/_ref: unexpected call to " #name "\n"); abort(); return NULL; }
DEAD(flb_cf_section_property_get_string) DEAD(flb_cf_destroy) DEAD(flb_cf_create_from_file)
DEAD(flb_file_read) DEAD(flb_condition_evaluate)

DEAD(mk_print_dead)
int mk_print() { return 0; }
/* src/flb_router.c is compiled for flb_router_match(); its output-routing half (never reached on the filter path)
 * refers to these */
int flb_router_apply_config(void *config) { (void) config; return 0; }
void flb_routes_empty_mask_destroy(void *config) { (void) config; }
/* filter_rewrite_tag's emitter set-up and its processor-stage shortcut (ingest_inline needs parent_processor, which a
 * filter instance never has) */
int flb_metrics_title(const char *title, void *metrics) { (void) title; (void) metrics; return 0; }
DEAD(flb_input_instance_exit) DEAD(flb_input_instance_destroy) DEAD(flb_input_log_append_skip_processor_stages)
