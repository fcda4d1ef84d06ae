/*
 * TEST INFRASTRUCTURE ONLY.  The HDP entry points of the reference's vanillaAlign.c that this build does not provide:
 * BUILDING and Gibbs-sampling HDPs, serialising them, and feeding assignment files back into the sampler (SURVEY.md
 * 8(f) N4 covers reading an HDP, aligning with it and collecting its assignment lists: deserialize_nhdp,
 * destroy_nanopore_hdp, getHdpStateMachine3 and the HdpHmm container come from libcpecan_host.so).  They let the UNMODIFIED vanillaAlign.c link (oracle/_ref/vanillaAlign_dropin,
 * tests/test_vanilla_align_drop_in.py) and abort if ever reached.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#define HDP_STUB(sig) sig { fprintf(stderr, "HDP is out of scope of this build\n"); abort(); }
HDP_STUB(int serialize_nhdp(void *h, const char *f))
HDP_STUB(int execute_nhdp_gibbs_sampling(void *h, int64_t a, int64_t b, int64_t c, int d))
HDP_STUB(int finalize_nhdp_distributions(void *h))
HDP_STUB(int nanoporeHdp_buildNanoporeHdpFromAlignment(int t, const char *a, const char *b, const char *c, const char *d, const char *e))
HDP_STUB(void *hdpHmm_loadFromFile(const char *fileName, void *nhdp))
