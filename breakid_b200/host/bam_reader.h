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
const bkid_batch *bkid_host_bam_batch(const bkid_host_bam *h);            /* wide columns */
const bkid_batch *bkid_host_bam_batch_narrow(const bkid_host_bam *h);     /* same batch, narrow encodings where the data fits (11 B/record) */
void bkid_host_bam_free(bkid_host_bam *h);


/* Host half of the device decode path (bkid_push_bgzf): mmap the file, walk the BGZF block headers (BSIZE / ISIZE)
 * into a block table and inflate just enough leading blocks (zlib) to parse the BAM header.  Nothing else is
 * decompressed on the host. */
typedef struct bkid_host_bgzf bkid_host_bgzf;
bkid_host_bgzf *bkid_host_bgzf_open(const char *path, char *err, int errlen);
const bkid_header *bkid_host_bgzf_header(const bkid_host_bgzf *h);
const uint8_t *bkid_host_bgzf_data(const bkid_host_bgzf *h);            /* the mapped file */
uint64_t bkid_host_bgzf_size(const bkid_host_bgzf *h);
const bkid_bgzf_block *bkid_host_bgzf_blocks(const bkid_host_bgzf *h);
int64_t bkid_host_bgzf_n_blocks(const bkid_host_bgzf *h);
uint64_t bkid_host_bgzf_first_record(const bkid_host_bgzf *h);          /* uncompressed offset of the first record */
uint64_t bkid_host_bgzf_usize(const bkid_host_bgzf *h);                 /* total uncompressed size */
int32_t bkid_host_bgzf_first_l_qseq(const bkid_host_bgzf *h);           /* read length of the first record (driver's _params.txt), -1 if none */
void bkid_host_bgzf_close(bkid_host_bgzf *h);

#ifdef __cplusplus
}
#endif
