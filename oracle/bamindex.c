/* oracle/bamindex.c -- TEST INFRASTRUCTURE: builds <file>.bam.bai with the reference's own
 * vendored htslib (thirdparty/samtools/samtools-1.3.1/htslib-1.3.1, sam.c `sam_index_build`)
 * so that the reference binary reads an index produced by the library it was written against. */
#include <stdio.h>
#include "htslib/sam.h"
int main(int argc, char **argv)
{
    if (argc != 2) { fprintf(stderr, "usage: bamindex in.bam\n"); return 2; }
    int r = sam_index_build(argv[1], 0);
    if (r < 0) { fprintf(stderr, "bamindex: sam_index_build failed (%d)\n", r); return 1; }
    return 0;
}
