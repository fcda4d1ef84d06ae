/* TEST INFRASTRUCTURE ONLY -- stand-in for sonLib's pairwiseAlignment.h (cigar records). */
#ifndef SONLIB_STANDIN_PAIRWISE_ALIGNMENT_H_
#define SONLIB_STANDIN_PAIRWISE_ALIGNMENT_H_
#include <stdio.h>
#include <stdint.h>

#define PAIRWISE_INDEL_X 0
#define PAIRWISE_INDEL_Y 1
#define PAIRWISE_MATCH 2

struct List {
    void **list;
    int64_t length;
    int64_t maxLength;
    void (*destroyElement)(void *);
};

struct AlignmentOperation {
    int64_t opType;
    int64_t length;
    float score;
};

struct PairwiseAlignment {
    char *contig1;
    int64_t start1;
    int64_t end1;
    int64_t strand1;
    char *contig2;
    int64_t start2;
    int64_t end2;
    int64_t strand2;
    float score;
    struct List *operationList;
};

struct PairwiseAlignment *cigarRead(FILE *fileHandle);
void destructPairwiseAlignment(struct PairwiseAlignment *pA);
void checkPairwiseAlignment(struct PairwiseAlignment *pA);
/* sonLib string helpers the reference's vanillaAlign.c calls (it reaches sonLib.h only through this header) */
char *stString_reverseComplementString(const char *s);
char *stString_replace(const char *s, const char *what, const char *with);

#endif
