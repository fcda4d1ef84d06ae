/* TEST INFRASTRUCTURE ONLY -- opaque sonLib type names (see sonLib.h in this directory). */
#ifndef SONLIB_STANDIN_TYPES_H_
#define SONLIB_STANDIN_TYPES_H_
#include <stdint.h>
#include <stdbool.h>
typedef struct _stList stList;
typedef struct _stIntTuple stIntTuple;
typedef struct _stSortedSet stSortedSet;
typedef struct _stSet stSet;
typedef struct _stSetIterator stSetIterator;
typedef struct _stListIterator stListIterator;
typedef struct _stHash stHash;
#endif
