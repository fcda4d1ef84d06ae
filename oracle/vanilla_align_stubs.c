/*
 * TEST INFRASTRUCTURE ONLY.  Externals the reference's vanillaAlign.c needs beyond the five hot-path sources, so that
 * the UNMODIFIED vanillaAlign.c links into oracle/_ref/vanillaAlign (golden generator for the batched sibling
 * tools/cpecan_align.cpp): two sonLib string helpers, checkPairwiseAlignment, and the HDP entry points (out of scope:
 * they abort if ever reached).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "sonLib.h"
#include "pairwiseAlignment.h"

static char comp(char c) {
    switch (c) {
        case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
        case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a';
        default: return c;
    }
}
char *stString_reverseComplementString(const char *s) {
    size_t n = strlen(s);
    char *r = st_malloc(n + 1);
    for (size_t i = 0; i < n; i++) r[i] = comp(s[n - 1 - i]);
    r[n] = 0;
    return r;
}
char *stString_replace(const char *s, const char *what, const char *with) {
    size_t nw = strlen(what), nr = strlen(with), n = strlen(s), cap = n * (nr > nw ? nr : nw) / (nw ? nw : 1) + n + 1;
    char *out = st_malloc(cap + 1), *o = out;
    while (*s) {
        if (nw && strncmp(s, what, nw) == 0) { memcpy(o, with, nr); o += nr; s += nw; }
        else *o++ = *s++;
    }
    *o = 0;
    return out;
}
void checkPairwiseAlignment(struct PairwiseAlignment *pA) {
    /* sonLib asserts that the operations add up to the two intervals */
    int64_t lx = 0, ly = 0;
    for (int64_t i = 0; i < pA->operationList->length; i++) {
        struct AlignmentOperation *op = pA->operationList->list[i];
        if (op->opType != PAIRWISE_INDEL_Y) lx += op->length;
        if (op->opType != PAIRWISE_INDEL_X) ly += op->length;
    }
    int64_t ex = pA->end1 - pA->start1, ey = pA->end2 - pA->start2;
    if (llabs(ex) != lx || llabs(ey) != ly) st_errAbort("checkPairwiseAlignment: cigar operations do not match the intervals");
}
#ifndef STANDIN_REAL_HDP
#define HDP_STUB(sig) sig { st_errAbort("HDP is out of scope for the oracle build"); return 0; }
HDP_STUB(void *deserialize_nhdp(const char *f))
HDP_STUB(int serialize_nhdp(void *h, const char *f))
HDP_STUB(int destroy_nanopore_hdp(void *h))
HDP_STUB(int execute_nhdp_gibbs_sampling(void *h, int64_t a, int64_t b, int64_t c, int d))
HDP_STUB(int finalize_nhdp_distributions(void *h))
HDP_STUB(int nanoporeHdp_buildNanoporeHdpFromAlignment(int t, const char *a, const char *b, const char *c, const char *d, const char *e))
#endif
