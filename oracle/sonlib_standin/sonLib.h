/*
 * TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * Minimal stand-in for benedictpaten/sonLib, which the reference depends on
 * (include.mk:2 of the reference: sonLibRootPath=${rootPath}../sonLib) but does
 * not vendor and does not pin.  Only the symbols the hot-path sources
 * (impl/pairwiseAligner.c, impl/stateMachine.c, impl/nanopore.c,
 * impl/continuousHmm.c, impl/discreteHmm.c) and, for the threeStateHdp goldens,
 * the HDP sources (impl/hdp.c, impl/nanopore_hdp.c, impl/hdp_math_utils.c)
 * reference are provided; they are containers and string/file helpers, none of
 * the DP arithmetic lives here.
 * Used exclusively to compile the UNMODIFIED reference sources into
 * oracle/_ref/ so the oracle restatement can be pinned against them.
 */
#ifndef SONLIB_STANDIN_H_
#define SONLIB_STANDIN_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>
#include <stdbool.h>
#include <inttypes.h>
#include <assert.h>
#include <math.h>
#include "sonLibTypes.h"

#ifdef __cplusplus
extern "C" {
#endif

void *st_malloc(size_t n);
void *st_calloc(size_t n, size_t sz);
void st_errAbort(const char *fmt, ...);
void st_uglyf(const char *fmt, ...);
void st_logDebug(const char *fmt, ...);
void st_logInfo(const char *fmt, ...);
int64_t st_system(const char *fmt, ...);
double st_random(void);
int64_t st_randomInt(int64_t lo, int64_t hi);

/* stList: growable array of void* with optional element destructor */
stList *stList_construct(void);
stList *stList_construct3(int64_t length, void (*destructElement)(void *));
void stList_destruct(stList *l);
int64_t stList_length(stList *l);
void *stList_get(stList *l, int64_t i);
void stList_set(stList *l, int64_t i, void *item);
void stList_append(stList *l, void *item);
void stList_appendAll(stList *l, stList *other);
void *stList_pop(stList *l);
void stList_sort(stList *l, int (*cmp)(const void *, const void *));
void stList_setDestructor(stList *l, void (*destructElement)(void *));
double *stList_toDoublePtr(stList *l, int64_t *lengthOut);
int64_t *stList_toIntPtr(stList *l, int64_t *lengthOut);
void stList_removeItem(stList *l, void *item);
stListIterator *stList_getIterator(stList *l);
void *stList_getNext(stListIterator *it);
void stList_destructIterator(stListIterator *it);

/* stSet: hash set of pointers (identity), iteration in table order */
stSet *stSet_construct(void);
stSet *stSet_construct2(void (*destructElement)(void *));
void stSet_destruct(stSet *s);
void stSet_insert(stSet *s, void *item);
void *stSet_search(stSet *s, void *item);
void *stSet_remove(stSet *s, void *item);
int64_t stSet_size(stSet *s);
stSetIterator *stSet_getIterator(stSet *s);
void *stSet_getNext(stSetIterator *it);
void stSet_destructIterator(stSetIterator *it);

/* stIntTuple: fixed small tuple of int64 */
stIntTuple *stIntTuple_construct2(int64_t a, int64_t b);
stIntTuple *stIntTuple_construct3(int64_t a, int64_t b, int64_t c);
stIntTuple *stIntTuple_construct4(int64_t a, int64_t b, int64_t c, int64_t d);
void stIntTuple_destruct(stIntTuple *t);
int64_t stIntTuple_get(stIntTuple *t, int64_t i);
int64_t stIntTuple_length(stIntTuple *t);
int stIntTuple_cmpFn(const void *a, const void *b);

/* stSortedSet: only construct3/insert/search/destruct are used on the hot path */
stSortedSet *stSortedSet_construct3(int (*cmp)(const void *, const void *), void (*destructElement)(void *));
void stSortedSet_destruct(stSortedSet *s);
void stSortedSet_insert(stSortedSet *s, void *item);
void *stSortedSet_search(stSortedSet *s, void *item);

char *stString_print(const char *fmt, ...);
char *stString_copy(const char *s);
char *stString_getSubString(const char *s, int64_t start, int64_t length);
stList *stString_split(const char *s);
stList *stString_splitByString(const char *s, const char *delim);
char *stFile_getLineFromFile(FILE *f);
/* used by the reference's vanillaAlign.c only (oracle/vanilla_align_stubs.c) */
char *stString_reverseComplementString(const char *s);
char *stString_replace(const char *s, const char *what, const char *with);

/* exceptions: the hot path throws only from diagonal_construct on invalid coordinates */
void stThrowNew(const char *id, const char *fmt, ...);

#ifndef TRUE
#define TRUE 1
#endif
#ifndef FALSE
#define FALSE 0
#endif

#ifdef __cplusplus
}
#endif
#endif
