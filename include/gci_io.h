/*
 * libgci_io.so — native decoders (host C++, zlib, threads) into the columnar record schema of
 * include/gci_cuda.h.  They replace what the reference obtains from pysam / htslib / Biopython:
 *   gci_bam_*    pysam.AlignmentFile header + per-record accessors        GCI.py:150-166, :201-208, :963-976
 *   gci_paf_*    the PAF column parse                                     GCI.py:216-229
 *   gci_fasta_*  SeqIO.parse ids and the (?i)N+ runs                      GCI.py:28-35, :939-941
 * Functions return 0 on success, -1 on error (message: gci_io_last_error, thread-local).
 */
#ifndef GCI_IO_H
#define GCI_IO_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct gci_interner gci_interner;   /* read name -> dense id, shared by all files of a read type */
typedef struct gci_bam gci_bam;
typedef struct gci_paf gci_paf;
typedef struct gci_fasta gci_fasta;

const char* gci_io_last_error(void);

gci_interner* gci_interner_create(void);
void gci_interner_destroy(gci_interner* it);
int64_t gci_interner_size(gci_interner* it);

/* decode a BAM file into columns: the file is streamed through a window of 256 MB of inflated data (BGZF blocks
   inflated in parallel on `threads` threads, record sizes walked in file order, fields extracted in parallel; a
   record cut by the window border is carried into the next window), so memory is bounded whatever the file size.
   GCI_IO_WINDOW_BYTES in the environment overrides the window (tests). */
int gci_bam_open(const char* path, int threads, gci_bam** out);
void gci_bam_close(gci_bam* b);
int32_t gci_bam_n_refs(gci_bam* b);
const char* gci_bam_ref_name(gci_bam* b, int32_t i);
int64_t gci_bam_ref_len(gci_bam* b, int32_t i);
int64_t gci_bam_n_records(gci_bam* b);
int64_t gci_bam_n_ops(gci_bam* b);          /* CIGAR ops after resolving CG:B,I long CIGARs */
/* fill caller-allocated columns (sizes from gci_bam_n_records / gci_bam_n_ops); nm = INT32_MIN when absent */
int gci_bam_fill(gci_bam* b, gci_interner* it, int32_t* ref_id, int32_t* ref_start, uint8_t* mapq, uint16_t* flag,
                 int32_t* nm, int32_t* qlen, uint32_t* read_id, uint64_t* cigar_off, uint32_t* cigar);

/* PAF columns 0,1,2,3,5,7,8,9,10,11; ref_id = index into contig_names or -1.  Lines are parsed in one chunk per
   host thread (GCI_IO_THREADS overrides the hardware count); read ids follow first appearance in file order and
   a malformed line is reported by its number. */
int gci_paf_open(const char* path, gci_interner* it, int32_t n_contigs, const char* const* contig_names,
                 gci_paf** out);
int64_t gci_paf_n_lines(gci_paf* p);
int gci_paf_fill(gci_paf* p, uint32_t* read_id, int32_t* qlen, int32_t* qstart, int32_t* qend, int32_t* ref_id,
                 int32_t* tstart, int32_t* tend, int32_t* nmatch, int32_t* alnlen, int32_t* mapq);
void gci_paf_close(gci_paf* p);

/* FASTA (plain or gzip): record ids in file order and all N/n runs as (record index, start, end) */
int gci_fasta_open(const char* path, gci_fasta** out);
int32_t gci_fasta_n_records(gci_fasta* f);
const char* gci_fasta_id(gci_fasta* f, int32_t i);
int64_t gci_fasta_n_runs(gci_fasta* f);
int gci_fasta_runs(gci_fasta* f, int32_t* rec, int64_t* start, int64_t* end);
void gci_fasta_close(gci_fasta* f);

/* .depth.gz (the resume point of utility/GCI_score.py:11-39): ">name" lines, then one decimal depth per line;
   multi-member gzip.  Inflated through a window, the numbers parsed on `threads` threads. */
typedef struct gci_depth gci_depth;
int gci_depth_open(const char* path, int threads, gci_depth** out);
int32_t gci_depth_n_contigs(gci_depth* d);
const char* gci_depth_name(gci_depth* d, int32_t i);
int64_t gci_depth_len(gci_depth* d, int32_t i);
int gci_depth_fill(gci_depth* d, int32_t i, int32_t* out);
void gci_depth_close(gci_depth* d);

#ifdef __cplusplus
}
#endif
#endif
