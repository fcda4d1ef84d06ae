/*
 * TEST INFRASTRUCTURE ONLY.  The HDP entry points the reference's vanillaAlign.c references (threeStateHdp, SURVEY.md
 * 8(f) N4: out of scope).  They let the UNMODIFIED vanillaAlign.c link against libcpecan_host.so
 * (oracle/_ref/vanillaAlign_dropin, tests/test_vanilla_align_drop_in.py) and abort if ever reached.
 */
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#define HDP_STUB(sig) sig { fprintf(stderr, "HDP is out of scope of this build\n"); abort(); }
HDP_STUB(void *deserialize_nhdp(const char *f))
HDP_STUB(int serialize_nhdp(void *h, const char *f))
HDP_STUB(int destroy_nanopore_hdp(void *h))
HDP_STUB(int execute_nhdp_gibbs_sampling(void *h, int64_t a, int64_t b, int64_t c, int d))
HDP_STUB(int finalize_nhdp_distributions(void *h))
HDP_STUB(int nanoporeHdp_buildNanoporeHdpFromAlignment(int t, const char *a, const char *b, const char *c, const char *d, const char *e))
HDP_STUB(void *getHdpStateMachine3(void *hdp, const char *modelFile))
HDP_STUB(void *hdpHmm_loadFromFile(const char *fileName, void *nhdp))
