/* TEST INFRASTRUCTURE ONLY -- stand-in for sonLib's bioioC.h (fasta writer + temp files). */
#ifndef SONLIB_STANDIN_BIOIOC_H_
#define SONLIB_STANDIN_BIOIOC_H_
#include <stdio.h>
#include "pairwiseAlignment.h"
void fastaWrite(char *sequence, char *header, FILE *file);
char *getTempFile(void);
#endif
