/* TEST INFRASTRUCTURE ONLY -- sonLib header stand-in: impl/nanopore_hdp.c includes it and uses nothing from it. */
#ifndef FASTCMATHS_STANDIN_H_
#define FASTCMATHS_STANDIN_H_
#endif
