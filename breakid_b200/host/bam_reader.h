// bam_reader.h -- host BAM decode -> bkid_batch (see bam_reader.cc)
#pragma once
#include "../../include/breakid_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct bkid_host_bam bkid_host_bam;

/* Decode a whole BAM into one record batch.  Returns NULL and fills `err` on failure. */
bkid_host_bam *bkid_host_read_bam(const char *path, int threads, char *err, int errlen);
const bkid_header *bkid_host_bam_header(const bkid_host_bam *h);
const bkid_batch *bkid_host_bam_batch(const bkid_host_bam *h);
void bkid_host_bam_free(bkid_host_bam *h);

#ifdef __cplusplus
}
#endif
