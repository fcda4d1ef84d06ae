/*
 * TEST INFRASTRUCTURE ONLY -- see sonLib.h in this directory.
 * Containers / string helpers / cigar parser needed to compile the unmodified
 * reference hot-path sources into oracle/_ref/.  None of this is reachable from
 * the product library.
 */
#define _GNU_SOURCE
#include <stdarg.h>
#include <unistd.h>
#include <ctype.h>
#include "sonLib.h"
#include "bioioC.h"
#include "pairwiseAlignment.h"

/* ---------------------------------------------------------------- basics */
void *st_malloc(size_t n) {
    /* +64: the reference under-allocates its model tables by 7 bytes and writes one byte past its 6-byte k-mer
     * buffers; with slack the unmodified sources run deterministically instead of corrupting the heap */
    void *p = calloc((n ? n : 1) + 64, 1);   /* zeroed: vanillaAlign.c:63-66 takes strlen() of an unterminated 6-byte k-mer */
    if (!p) { fprintf(stderr, "st_malloc: out of memory\n"); abort(); }
    return p;
}
void *st_calloc(size_t n, size_t sz) {
    void *p = calloc(n ? n : 1, sz ? sz : 1);
    if (!p) { fprintf(stderr, "st_calloc: out of memory\n"); abort(); }
    return p;
}
void st_errAbort(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
    fputc('\n', stderr);
    abort();
}
void st_uglyf(const char *fmt, ...) {
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
void st_logDebug(const char *fmt, ...) { (void) fmt; }
void st_logInfo(const char *fmt, ...) { (void) fmt; }
int64_t st_system(const char *fmt, ...) {
    char *cmd = NULL; va_list ap; va_start(ap, fmt);
    if (vasprintf(&cmd, fmt, ap) < 0) abort();
    va_end(ap);
    int rc = system(cmd); free(cmd); return rc;
}
double st_random(void) { return (double) rand() / ((double) RAND_MAX + 1.0); }
int64_t st_randomInt(int64_t lo, int64_t hi) { return lo + (int64_t) (st_random() * (double) (hi - lo)); }

/* ---------------------------------------------------------------- stList */
struct _stList { void **items; int64_t n, cap; void (*destructElement)(void *); };

stList *stList_construct3(int64_t length, void (*destructElement)(void *)) {
    stList *l = st_malloc(sizeof(*l));
    l->cap = length > 8 ? length : 8;
    l->items = st_calloc(l->cap, sizeof(void *));
    l->n = length; l->destructElement = destructElement;
    return l;
}
stList *stList_construct(void) { return stList_construct3(0, NULL); }
void stList_destruct(stList *l) {
    if (!l) return;
    if (l->destructElement) for (int64_t i = 0; i < l->n; i++) if (l->items[i]) l->destructElement(l->items[i]);
    free(l->items); free(l);
}
int64_t stList_length(stList *l) { return l ? l->n : 0; }
void *stList_get(stList *l, int64_t i) { assert(i >= 0 && i < l->n); return l->items[i]; }
void stList_set(stList *l, int64_t i, void *item) { assert(i >= 0 && i < l->n); l->items[i] = item; }
void stList_append(stList *l, void *item) {
    if (l->n == l->cap) { l->cap *= 2; l->items = realloc(l->items, l->cap * sizeof(void *)); if (!l->items) abort(); }
    l->items[l->n++] = item;
}
void stList_appendAll(stList *l, stList *o) { for (int64_t i = 0; i < o->n; i++) stList_append(l, o->items[i]); }
void *stList_pop(stList *l) { assert(l->n > 0); return l->items[--l->n]; }
void stList_setDestructor(stList *l, void (*d)(void *)) { l->destructElement = d; }
static int sortShim(const void *a, const void *b, void *cmp) {
    int (*f)(const void *, const void *) = (int (*)(const void *, const void *)) cmp;
    return f(*(void *const *) a, *(void *const *) b);
}
void stList_sort(stList *l, int (*cmp)(const void *, const void *)) {
    qsort_r(l->items, l->n, sizeof(void *), sortShim, (void *) cmp);
}
#ifndef STANDIN_REAL_HDP   /* impl/hdp_math_utils.c defines this one itself */
double *stList_toDoublePtr(stList *l, int64_t *lengthOut) {
    double *d = st_malloc(sizeof(double) * (l->n ? l->n : 1));
    for (int64_t i = 0; i < l->n; i++) d[i] = *(double *) l->items[i];
    *lengthOut = l->n; return d;
}
#endif

/* ------------------------------------------------------------ stIntTuple */
struct _stIntTuple { int64_t n; int64_t v[4]; };
static stIntTuple *tupleN(int64_t n, int64_t a, int64_t b, int64_t c, int64_t d) {
    stIntTuple *t = st_malloc(sizeof(*t)); t->n = n; t->v[0] = a; t->v[1] = b; t->v[2] = c; t->v[3] = d; return t;
}
stIntTuple *stIntTuple_construct2(int64_t a, int64_t b) { return tupleN(2, a, b, 0, 0); }
stIntTuple *stIntTuple_construct3(int64_t a, int64_t b, int64_t c) { return tupleN(3, a, b, c, 0); }
stIntTuple *stIntTuple_construct4(int64_t a, int64_t b, int64_t c, int64_t d) { return tupleN(4, a, b, c, d); }
void stIntTuple_destruct(stIntTuple *t) { free(t); }
int64_t stIntTuple_get(stIntTuple *t, int64_t i) { assert(i >= 0 && i < t->n); return t->v[i]; }
int64_t stIntTuple_length(stIntTuple *t) { return t->n; }
int stIntTuple_cmpFn(const void *a, const void *b) {
    const stIntTuple *s = a, *t = b;
    int64_t n = s->n < t->n ? s->n : t->n;
    for (int64_t i = 0; i < n; i++) if (s->v[i] != t->v[i]) return s->v[i] < t->v[i] ? -1 : 1;
    return s->n == t->n ? 0 : (s->n < t->n ? -1 : 1);
}

/* ------------------------------------------------------------ stSortedSet
 * Hash-free, order-free: a flat array searched linearly would be O(n^2) on
 * large anchor sets, so keep it sorted lazily and bsearch. */
struct _stSortedSet { void **items; int64_t n, cap; int sorted; int (*cmp)(const void *, const void *); void (*destructElement)(void *); };
stSortedSet *stSortedSet_construct3(int (*cmp)(const void *, const void *), void (*destructElement)(void *)) {
    stSortedSet *s = st_malloc(sizeof(*s)); s->cap = 16; s->n = 0; s->sorted = 1;
    s->items = st_malloc(sizeof(void *) * s->cap); s->cmp = cmp; s->destructElement = destructElement; return s;
}
void stSortedSet_destruct(stSortedSet *s) {
    if (s->destructElement) for (int64_t i = 0; i < s->n; i++) s->destructElement(s->items[i]);
    free(s->items); free(s);
}
void stSortedSet_insert(stSortedSet *s, void *item) {
    if (s->n == s->cap) { s->cap *= 2; s->items = realloc(s->items, s->cap * sizeof(void *)); if (!s->items) abort(); }
    s->items[s->n++] = item; s->sorted = 0;
}
void *stSortedSet_search(stSortedSet *s, void *item) {
    if (!s->sorted) { qsort_r(s->items, s->n, sizeof(void *), sortShim, (void *) s->cmp); s->sorted = 1; }
    int64_t lo = 0, hi = s->n - 1;
    while (lo <= hi) {
        int64_t mid = (lo + hi) / 2; int c = s->cmp(item, s->items[mid]);
        if (c == 0) return s->items[mid];
        if (c < 0) hi = mid - 1; else lo = mid + 1;
    }
    return NULL;
}

/* --------------------------------------------------------------- strings */
char *stString_print(const char *fmt, ...) {
    char *out = NULL; va_list ap; va_start(ap, fmt);
    if (vasprintf(&out, fmt, ap) < 0) abort();
    va_end(ap); return out;
}
char *stString_copy(const char *s) { char *c = st_malloc(strlen(s) + 1); strcpy(c, s); return c; }
char *stString_getSubString(const char *s, int64_t start, int64_t length) {
    char *c = st_malloc(length + 1); memcpy(c, s + start, length); c[length] = '\0'; return c;
}
stList *stString_split(const char *s) {
    stList *tokens = stList_construct3(0, free);
    const char *p = s;
    while (*p) {
        while (*p && isspace((unsigned char) *p)) p++;
        if (!*p) break;
        const char *q = p;
        while (*q && !isspace((unsigned char) *q)) q++;
        stList_append(tokens, stString_getSubString(p, 0, q - p));
        p = q;
    }
    return tokens;
}
char *stFile_getLineFromFile(FILE *f) {
    char *line = NULL; size_t cap = 0;
    ssize_t n = getline(&line, &cap, f);
    if (n < 0) { free(line); return NULL; }
    while (n > 0 && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = '\0';
    return line;
}

void stThrowNew(const char *id, const char *fmt, ...) {
    fprintf(stderr, "uncaught exception %s: ", id);
    va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
    fputc('\n', stderr);
    abort();
}

/* ------------------------------------------------------------ bioio bits */
void fastaWrite(char *sequence, char *header, FILE *file) {
    fprintf(file, ">%s\n", header);
    size_t l = strlen(sequence);
    for (size_t i = 0; i < l; i += 100) fprintf(file, "%.100s\n", sequence + i);
}
char *getTempFile(void) {
    char tmpl[] = "/tmp/cpecan_ref_XXXXXX";
    int fd = mkstemp(tmpl);
    if (fd < 0) st_errAbort("mkstemp failed");
    close(fd);
    return stString_copy(tmpl);
}

/* exonerate-style cigar: "cigar: <query> qs qe q± <target> ts te t± score (op len)*".
 * The query lands in contig2/start2/end2, the target in contig1 (as sonLib does);
 * M -> match, D -> indel in X (target only), I -> indel in Y (query only). */
struct PairwiseAlignment *cigarRead(FILE *fh) {
    char *line;
    while ((line = stFile_getLineFromFile(fh)) != NULL) {
        if (strncmp(line, "cigar:", 6) != 0) { free(line); continue; }
        stList *tok = stString_split(line + 6);
        free(line);
        if (stList_length(tok) < 9) { stList_destruct(tok); continue; }
        struct PairwiseAlignment *pA = st_calloc(1, sizeof(*pA));
        pA->contig2 = stString_copy(stList_get(tok, 0));
        pA->start2 = atoll(stList_get(tok, 1));
        pA->end2 = atoll(stList_get(tok, 2));
        pA->strand2 = ((char *) stList_get(tok, 3))[0] == '+';
        pA->contig1 = stString_copy(stList_get(tok, 4));
        pA->start1 = atoll(stList_get(tok, 5));
        pA->end1 = atoll(stList_get(tok, 6));
        pA->strand1 = ((char *) stList_get(tok, 7))[0] == '+';
        pA->score = (float) atof(stList_get(tok, 8));
        struct List *ops = st_calloc(1, sizeof(*ops));
        int64_t nOps = (stList_length(tok) - 9) / 2;
        ops->list = st_calloc(nOps ? nOps : 1, sizeof(void *));
        ops->length = nOps; ops->maxLength = nOps;
        for (int64_t i = 0; i < nOps; i++) {
            struct AlignmentOperation *op = st_calloc(1, sizeof(*op));
            char c = ((char *) stList_get(tok, 9 + 2 * i))[0];
            op->opType = c == 'M' ? PAIRWISE_MATCH : (c == 'D' ? PAIRWISE_INDEL_X : PAIRWISE_INDEL_Y);
            op->length = atoll(stList_get(tok, 10 + 2 * i));
            ops->list[i] = op;
        }
        pA->operationList = ops;
        stList_destruct(tok);
        return pA;
    }
    return NULL;
}
void destructPairwiseAlignment(struct PairwiseAlignment *pA) {
    for (int64_t i = 0; i < pA->operationList->length; i++) free(pA->operationList->list[i]);
    free(pA->operationList->list); free(pA->operationList);
    free(pA->contig1); free(pA->contig2); free(pA);
}

/* ------------------------------------------ containers the HDP sources use */
#ifndef STANDIN_REAL_HDP
int64_t *stList_toIntPtr(stList *l, int64_t *lengthOut) {
    int64_t *v = st_malloc(sizeof(int64_t) * (l->n ? l->n : 1));
    for (int64_t i = 0; i < l->n; i++) if (sscanf(l->items[i], "%" SCNd64, &v[i]) != 1) st_errAbort("stList_toIntPtr: not an integer");
    *lengthOut = l->n;
    return v;
}
#endif
void stList_removeItem(stList *l, void *item) {
    for (int64_t i = 0; i < l->n; i++)
        if (l->items[i] == item) { memmove(l->items + i, l->items + i + 1, sizeof(void *) * (l->n - i - 1)); l->n--; return; }
}
struct _stListIterator { stList *l; int64_t i; };
stListIterator *stList_getIterator(stList *l) { stListIterator *it = st_malloc(sizeof(*it)); it->l = l; it->i = 0; return it; }
void *stList_getNext(stListIterator *it) { return it->i < it->l->n ? it->l->items[it->i++] : NULL; }
void stList_destructIterator(stListIterator *it) { free(it); }

#define SET_TOMB ((void *) 1)
struct _stSet { void **slot; int64_t cap, n, used; void (*destructElement)(void *); };
/* the iterator walks a snapshot: impl/hdp.c destroys sets (and the factors in them) while iterating over them */
struct _stSetIterator { void **items; int64_t n, i; };
static uint64_t ptrHash(void *p) { uint64_t x = (uint64_t) (uintptr_t) p; x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; return x; }
stSet *stSet_construct2(void (*d)(void *)) {
    stSet *s = st_calloc(1, sizeof(*s));
    s->cap = 8; s->slot = st_calloc(s->cap, sizeof(void *)); s->destructElement = d;
    return s;
}
stSet *stSet_construct(void) { return stSet_construct2(NULL); }
void stSet_destruct(stSet *s) {
    if (s->destructElement) for (int64_t i = 0; i < s->cap; i++) if (s->slot[i] && s->slot[i] != SET_TOMB) s->destructElement(s->slot[i]);
    free(s->slot); free(s);
}
static void setPut(stSet *s, void *item) {
    int64_t i = (int64_t) (ptrHash(item) & (uint64_t) (s->cap - 1)), tomb = -1;
    while (s->slot[i]) {
        if (s->slot[i] == item) return;
        if (s->slot[i] == SET_TOMB && tomb < 0) tomb = i;
        i = (i + 1) & (s->cap - 1);
    }
    if (tomb >= 0) i = tomb; else s->used++;
    s->slot[i] = item; s->n++;
}
void stSet_insert(stSet *s, void *item) {
    if (item == NULL || item == SET_TOMB) st_errAbort("stSet_insert: reserved pointer");
    if (2 * (s->used + 1) > s->cap) {
        void **old = s->slot; int64_t oc = s->cap;
        s->cap = 2 * s->n + 2 > s->cap ? 2 * s->cap : s->cap;      /* grow, or only clear the tombstones */
        s->slot = st_calloc(s->cap, sizeof(void *)); s->n = 0; s->used = 0;
        for (int64_t i = 0; i < oc; i++) if (old[i] && old[i] != SET_TOMB) setPut(s, old[i]);
        free(old);
    }
    setPut(s, item);
}
static int64_t setFind(stSet *s, void *item) {
    int64_t i = (int64_t) (ptrHash(item) & (uint64_t) (s->cap - 1));
    while (s->slot[i]) { if (s->slot[i] == item) return i; i = (i + 1) & (s->cap - 1); }
    return -1;
}
void *stSet_search(stSet *s, void *item) { return setFind(s, item) >= 0 ? item : NULL; }
void *stSet_remove(stSet *s, void *item) {
    int64_t i = setFind(s, item);
    if (i < 0) return NULL;
    s->slot[i] = SET_TOMB; s->n--;
    return item;
}
int64_t stSet_size(stSet *s) { return s->n; }
stSetIterator *stSet_getIterator(stSet *s) {
    stSetIterator *it = st_malloc(sizeof(*it));
    it->items = st_malloc(sizeof(void *) * (s->n ? s->n : 1)); it->n = 0; it->i = 0;
    for (int64_t k = 0; k < s->cap; k++) if (s->slot[k] && s->slot[k] != SET_TOMB) it->items[it->n++] = s->slot[k];
    return it;
}
void *stSet_getNext(stSetIterator *it) { return it->i < it->n ? it->items[it->i++] : NULL; }
void stSet_destructIterator(stSetIterator *it) { free(it->items); free(it); }

stList *stString_splitByString(const char *s, const char *delim) {
    stList *l = stList_construct3(0, free);
    size_t nd = strlen(delim);
    const char *p = s;
    for (;;) {
        const char *q = nd ? strstr(p, delim) : NULL;
        size_t n = q ? (size_t) (q - p) : strlen(p);
        char *tok = st_malloc(n + 1); memcpy(tok, p, n); tok[n] = 0;
        stList_append(l, tok);
        if (!q) break;
        p = q + nd;
    }
    return l;
}

#ifndef STANDIN_REAL_HDP
/* ----------------------------------------------- HDP symbols (stubs)
 * stateMachine.c / continuousHmm.c reference four HDP entry points for the
 * threeStateHdp variant; builds without the reference's HDP sources
 * (libcpecan_ref.so, oracle/_ref/vanillaAlign) get aborting stubs, the
 * HDP golden generator (oracle/_ref/vanillaAlign_hdp, -DSTANDIN_REAL_HDP)
 * links the real ones. */
typedef struct _nanoporeHDP NanoporeHDP;
double get_nanopore_kmer_density(NanoporeHDP *nhdp, void *kmer, void *event) {
    (void) nhdp; (void) kmer; (void) event; st_errAbort("HDP emissions are out of scope"); return 0.0;
}
int64_t kmer_id(char *kmer, char *alphabet, int64_t alphabet_size, int64_t kmer_length) {
    (void) kmer; (void) alphabet; (void) alphabet_size; (void) kmer_length; st_errAbort("HDP is out of scope"); return 0;
}
void pass_data_to_hdp(void *hdp, double *data, int64_t *dp_ids, int64_t length) {
    (void) hdp; (void) data; (void) dp_ids; (void) length; st_errAbort("HDP is out of scope");
}
void reset_hdp_data(void *hdp) { (void) hdp; st_errAbort("HDP is out of scope"); }
#endif
