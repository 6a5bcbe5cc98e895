/* tests/hostsim -- CPU-only emulation of the DEVICE code paths.
 *
 * The .cuh files under fluent-bit_b200/csrc are written as host+device functions.
 * This file compiles them with g++ so the `-m "not gpu"` tests can check the exact
 * device algorithms (regex VM, msgpack walkers, chain interpreter) against the
 * oracle on a box without a GPU.  It is TEST INFRASTRUCTURE: libflbgpu.so (the
 * product) never links or loads it and has no CPU path of any kind.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "../../fluent-bit_b200/csrc/flbgpu_prog.h"
#include "../../fluent-bit_b200/csrc/rx_compile.h"
#include "../../fluent-bit_b200/csrc/dev_regex.cuh"

extern "C" {

void *sim_rx_compile(const char *pattern, char *err, int errcap)
{
    struct rx_compiled *c = (struct rx_compiled *) calloc(1, sizeof(*c));
    if (rx_compile(pattern, c) != 0) {
        if (err) { strncpy(err, c->err, errcap - 1); err[errcap - 1] = 0; }
        free(c);
        return NULL;
    }
    return c;
}

void sim_rx_free(void *h) { if (h) { rx_compiled_free((struct rx_compiled *) h); free(h); } }

int sim_rx_ngroups(void *h) { return (int) ((struct rx_compiled *) h)->prog->n_groups; }
int sim_rx_ncode(void *h) { return (int) ((struct rx_compiled *) h)->prog->n_code; }
const uint32_t *sim_rx_code(void *h)
{
    struct rx_prog *p = ((struct rx_compiled *) h)->prog;
    return (const uint32_t *) ((const char *) p + p->code_off);
}

int sim_rx_names(void *h, char *out, int cap)
{
    struct rx_compiled *c = (struct rx_compiled *) h;
    int i, len = 0;
    for (i = 0; i < c->n_names; i++) {
        int l = (int) strlen(c->names[i].name);
        if (len + l + 1 >= cap) break;
        memcpy(out + len, c->names[i].name, l);
        len += l;
        out[len++] = '\n';
    }
    out[len] = 0;
    return c->n_names;
}

/* returns RX_R_*; caps gets 2*(ngroups+1) ints */
int sim_rx_search(void *h, const char *s, int len, int *caps, int stack_words, unsigned budget)
{
    struct rx_compiled *c = (struct rx_compiled *) h;
    uint32_t *stk = (uint32_t *) malloc(sizeof(uint32_t) * (stack_words > 0 ? stack_words : 1));
    uint32_t b = budget;
    int r = rx_search(c->prog, (const uint8_t *) s, len, caps, stk, stack_words, &b);
    free(stk);
    return r;
}

}
