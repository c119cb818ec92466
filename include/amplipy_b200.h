/*
 * amplipy_b200.h -- C ABI of the B200-native trim -> pileup -> call path (libamplipy_b200.so).
 *
 * The reference (Niema-Lab/AmpliPy, /root/reference/AmpliPy.py) has no FFI: its hot path is three
 * Python functions called per read / per position.  Each entry point below states which of them it
 * replaces; INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions: every function returns 0 on success or a negative AMP_ERR_* code (message from
 * amp_last_error(), thread-local); nothing throws; all buffers are caller-owned; no torch types.
 * "device" pointers are CUDA device pointers on the context's device; `stream` is a cudaStream_t
 * passed as void* (NULL = default stream).
 *
 * Batch layout (struct-of-arrays, one entry per read unless noted; amplipy_b200/batch.py):
 *   pos i32, flag u16, tlen i32, cig_off u32[N+1], cigar u32[sumC] (BAM: len<<4|op),
 *   seq_off u32[N+1] (bytes), seq u8 (BAM 4-bit, (l_seq+1)/2 bytes per read),
 *   qual_off u32[N+1] (bytes; l_seq = diff), qual u8 (phred).
 *   `seq` and `qual` base addresses must be 16-byte aligned, and device copies of them must be readable up to the
 *   next 16-byte boundary past their last byte (the kernels move them in whole 16-byte pieces).
 * Trim output layout: row i of out_cigar starts at cig_off[i] + 3*i (capacity n_cigar[i] + 3).
 */
#ifndef AMPLIPY_B200_H
#define AMPLIPY_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AMP_ABI_VERSION 1

/* error codes */
#define AMP_OK 0
#define AMP_ERR_CUDA (-1)
#define AMP_ERR_ARG (-2)
#define AMP_ERR_NOMEM (-3)
#define AMP_ERR_STATE (-4)
#define AMP_ERR_DATA (-5)     /* malformed input data (message says what) */

/* out_flags bits per read (AmpliPy.py:443-447 return tuple of trim_read + write gate 910 + skip 902) */
#define AMP_FLAG_TRIM_START 1
#define AMP_FLAG_TRIM_END 2
#define AMP_FLAG_TRIM_QUAL 4
#define AMP_FLAG_KEEP 8
#define AMP_FLAG_SKIPPED 16
#define AMP_FLAG_ERROR 32

/* device error word bits (amp_error_flags): inputs for which the reference itself crashes */
#define AMP_DEVERR_COORD 1
#define AMP_DEVERR_BASE 2
#define AMP_DEVERR_INS_END 4
#define AMP_DEVERR_CIGAR 8
#define AMP_DEVERR_TABLE_FULL 16
#define AMP_DEVERR_ARENA_FULL 32

/* mode bits for amp_process_* */
#define AMP_MODE_TRIM 1      /* trim_read, AmpliPy.py:426-687 (+ gate 902/910) */
#define AMP_MODE_PILEUP 2    /* update_base_counts, AmpliPy.py:690-753 */

typedef struct amp_ctx amp_ctx;

typedef struct {
    int32_t device;             /* CUDA device ordinal */
    int32_t ref_len;            /* L: length of the reference genome */
    int32_t n_samples;          /* number of count matrices kept on the device (1; plate mode: many) */
    int32_t min_quality;        /* AmpliPy.py:27  */
    int32_t sliding_window;     /* AmpliPy.py:29  */
    int32_t min_length;         /* AmpliPy.py:26  */
    int32_t include_no_primer;  /* -e */
    int64_t ins_slots;          /* insertion hash table capacity (power of two; 0 = default) */
    int64_t ins_arena_bytes;    /* insertion string arena (0 = default) */
} amp_config;

typedef struct {
    int64_t first;              /* global index of the first read to process */
    int64_t n_reads;            /* reads to process: [first, first + n_reads) */
    const int32_t* pos; const uint16_t* flag; const int32_t* tlen;
    const uint32_t* cig_off; const uint32_t* cigar;
    const uint32_t* seq_off; const uint8_t* seq;
    const uint32_t* qual_off; const uint8_t* qual;
} amp_batch;

typedef struct {
    int32_t* pos;               /* [N] new reference_start (AmpliPy.py:514) */
    uint16_t* ncig;             /* [N] number of ops in the rewritten CIGAR */
    uint8_t* flags;             /* [N] AMP_FLAG_* */
    uint32_t* cigar;            /* [sumC + 3N] rows at cig_off[i] + 3*i */
} amp_trim_out;

typedef struct {
    int32_t min_depth_consensus; double min_freq_consensus;   /* AmpliPy.py:22,24 */
    int32_t min_depth_variants;  double min_freq_variants;    /* AmpliPy.py:23,25 */
} amp_call_params;

/* per-(sample, position) outputs of amp_call, host arrays of n_samples*L entries (x6 where noted) */
typedef struct {
    int32_t* depth;             /* total depth incl. insertion alleles (AmpliPy.py:767) */
    int32_t* top_id;            /* top allele: 0..5 = A C G T N '-', 6+k = insertion allele k of amp_ins_export order, -1 none */
    int32_t* top_count;
    uint8_t* pos_flags;         /* bit0 consensus passes (928); bit1 variant record emitted (940); bit2 GT includes ref (948) */
    int32_t* ref_count;         /* REF_DP (937) */
    double*  fixed_freq;        /* [.. x6] count/total as IEEE float64 */
    int32_t* fixed_rank;        /* [.. x6] index in the reference's sorted allele list (771), -1 if count 0 */
    uint8_t* alt_mask;          /* bit ch set: fixed symbol ch is an ALT allele (938) */
} amp_call_out;

const char* amp_last_error(void);
int amp_abi_version(void);

/* context: device tables + count matrices + insertion-allele table.  Primer tables are the two lists
 * returned by find_overlapping_primers (AmpliPy.py:174-209) with None encoded as -1; pass NULL for
 * pileup-only use (variants / consensus subcommands). */
int amp_create(const amp_config* cfg, const int32_t* min_primer_start, const int32_t* max_primer_end,
               int32_t max_primer_len, amp_ctx** out);
int amp_destroy(amp_ctx* ctx);
int amp_reset(amp_ctx* ctx);                          /* zero counts + insertion table + error word */
int amp_reset_async(amp_ctx* ctx, void* stream);      /* same, enqueued on `stream` without synchronising */
int amp_error_flags(amp_ctx* ctx, uint32_t* flags);   /* AMP_DEVERR_* accumulated since amp_reset */
int amp_lpad(amp_ctx* ctx);                           /* row stride of the count matrices (L rounded up to 32) */
int amp_sm_count(amp_ctx* ctx);

/* Replaces the per-read loop AmpliPy.py:896-915 for reads [first, first+n): trim_read + write gate,
 * and/or update_base_counts into count matrix `sample`.  Device-pointer variant: inputs already
 * resident in HBM; runs asynchronously on `stream`.  out may be NULL without AMP_MODE_TRIM.
 * sum_cigar_ops / sum_qual_bytes = totals over the processed range (known to the host that built the
 * batch); they size the long-CIGAR scratch and pick the shared-memory tile shape. */
int amp_process_device(amp_ctx* ctx, const amp_batch* dev_batch, int64_t sum_cigar_ops, int64_t sum_qual_bytes, int mode,
                       int sample, const amp_trim_out* dev_out, void* stream);
/* Host-buffer variant (the call a reference-side binding makes): chunks the batch, overlaps H2D copy,
 * kernel and D2H of the trim outputs on internal streams; returns when everything has landed. */
int amp_process_host(amp_ctx* ctx, const amp_batch* host_batch, int mode, int sample, const amp_trim_out* host_out);
int amp_last_launches(amp_ctx* ctx);                  /* kernels launched by the last amp_process_* / amp_call */

/* count matrices: device pointer [n_samples][6][lpad] int32 (channel order A C G T N '-'); copy to host */
int amp_counts_device(amp_ctx* ctx, int32_t** dev_counts);
int amp_bind_counts(amp_ctx* ctx, int32_t* dev_counts);          /* use a caller-owned device buffer (e.g. a torch tensor for NCCL) */
int amp_counts_host(amp_ctx* ctx, int sample, int32_t* host_counts /* [6][L] */);
int amp_counts_upload(amp_ctx* ctx, int sample, const int32_t* host_counts /* [6][L] */);   /* replace a sample's matrix */

/* insertion alleles (the non-fixed keys of symbol_counts_at_ref_pos, AmpliPy.py:745-748) */
int amp_ins_count(amp_ctx* ctx, int64_t* n_alleles, int64_t* n_chars);
int amp_ins_export(amp_ctx* ctx, int32_t* sample, int32_t* pos, int32_t* count, int64_t* str_off /* [n+1] */, char* chars);
/* add alleles produced elsewhere (another rank's amp_ins_export) into this context's table */
int amp_ins_merge(amp_ctx* ctx, int64_t n, const int32_t* sample, const int32_t* pos, const int32_t* count,
                  const int64_t* str_off, const char* chars);

/* Replaces alleles_from_counts + the calling loop (AmpliPy.py:756-771, 921-951) for every sample.
 * ref_seq: L raw FASTA characters.  Per-insertion-allele outputs are indexed like amp_ins_export.
 * The outputs are copied back as one queue of asynchronous copies followed by a single wait: destinations from
 * amp_host_alloc (page-locked) receive them at full PCIe speed, pageable ones work but are staged by the driver. */
int amp_call(amp_ctx* ctx, const char* ref_seq, const amp_call_params* p, const amp_call_out* host_out,
             double* ins_freq, int32_t* ins_rank, uint8_t* ins_alt);
/* the same kernels without the device->host copies: reference characters are uploaded once, results stay
 * in context-owned HBM buffers; asynchronous on `stream` */
int amp_set_reference(amp_ctx* ctx, const char* ref_seq);
int amp_call_device(amp_ctx* ctx, const amp_call_params* p, void* stream);

/* ---- deep-sample mode: one sample's reads sharded over ranks (the reference has no counterpart, its loop
 * AmpliPy.py:896-915 is sequential).  After every rank has processed its read range:
 *   amp_allreduce_counts  one ncclAllReduce(sum, int32) of the [n_samples][6][lpad] matrices, in place
 *   amp_ins_pack_device   this rank's insertion table -> one fixed-size slot in device memory
 *                         { u64 n_alleles, u64 arena_words, {u64 arena offset, u64 count}[cap_entries], arena[cap_arena_bytes] }
 *   (all-gather of the slots: amp_nccl_allgather, or any collective the host prefers)
 *   amp_ins_merge_packed  everybody else's alleles added to this rank's table by one kernel
 * Everything is asynchronous on `stream`; nothing allocates.  A table that does not fit its slot raises
 * AMP_DEVERR_TABLE_FULL / AMP_DEVERR_ARENA_FULL (amp_error_flags) and packs as empty.
 * NCCL is resolved with dlopen at the first call (libnccl.so.2 as already mapped by the process, else the system's;
 * AMP_NCCL_LIB overrides); `comm` is an ncclComm_t, from amp_nccl_comm_init or the caller's own ncclCommInitRank. */
int amp_nccl_unique_id(uint8_t* id128 /* [128] */);
int amp_nccl_comm_init(int device, int n_ranks, int rank, const uint8_t* id128, void** comm);
int amp_nccl_comm_destroy(void* comm);
int amp_nccl_allgather(void* comm, const void* dev_send, void* dev_recv, int64_t bytes_per_rank, void* stream);
int amp_allreduce_counts(amp_ctx* ctx, void* comm, void* stream);
int64_t amp_ins_slot_bytes(int64_t cap_entries, int64_t cap_arena_bytes);
int amp_ins_pack_device(amp_ctx* ctx, void* dev_slot, int64_t cap_entries, int64_t cap_arena_bytes, void* stream);
int amp_ins_merge_packed(amp_ctx* ctx, const void* dev_slots, int n_ranks, int my_rank, int64_t cap_entries,
                         int64_t cap_arena_bytes, void* stream);

/* pre-size what amp_process_device would otherwise grow on demand (it synchronises `stream` when it has to allocate) */
int amp_reserve(amp_ctx* ctx, int64_t max_reads, int64_t max_cigar_ops);
/* device-to-device copy of the count matrices into a caller-owned buffer of the same shape */
int amp_counts_copy_device(amp_ctx* ctx, int32_t* dev_dst, void* stream);

/* ---- BAM on the device (replaces pysam's record iteration, AmpliPy.py:296-360 + 896, in front of the path).
 * The compressed file crosses PCIe as it is (about a fifth of the decoded arrays) and is decoded in HBM: one warp inflates one
 * BGZF block (RFC 1951), one thread per block walks its BAM record chain, one warp per block scatters the records into the
 * struct-of-arrays batch the kernels read.  Requires what htslib guarantees for the files it writes: no record straddles a BGZF
 * block boundary; otherwise AMP_ERR_DATA is returned and the caller uses its host decoder.  CRC32 of the blocks is not checked
 * on the device (ISIZE is).
 *   bgzf, n_bytes      the file's bytes (page-locked memory makes the copy asynchronous)
 *   block_off[n_blocks], block_isize[n_blocks]   start of every BGZF block in `bgzf` and its ISIZE (a host-side header scan)
 *   body_off           offset of the first alignment record in the inflated stream (= size of the BAM header)
 * amp_process_decoded runs the fused kernel on the decoded batch and copies the trim outputs (rows as in amp_trim_out,
 * sized from amp_bam_info) to host_out; host_out may be NULL.  amp_decoded_copy_host hands the decoded arrays back. */
typedef struct { int64_t n_reads, sum_cigar_ops, sum_seq_bytes, sum_qual_bytes, raw_bytes; } amp_bam_info;
typedef struct {
    int32_t* pos; uint16_t* flag; int32_t* tlen; uint32_t* cig_off; uint32_t* cigar; uint32_t* seq_off; uint8_t* seq;
    uint32_t* qual_off; uint8_t* qual;
} amp_batch_out;
int amp_bam_decode_host(amp_ctx* ctx, const uint8_t* bgzf, int64_t n_bytes, const int64_t* block_off, const uint32_t* block_isize,
                        int64_t n_blocks, int64_t body_off, amp_bam_info* info);
int amp_process_decoded(amp_ctx* ctx, int mode, int sample, const amp_trim_out* host_out);
int amp_decoded_copy_host(amp_ctx* ctx, const amp_batch_out* host_arrays, uint64_t* rec_off);

/* ---- BGZF deflate on the device (replaces the zlib deflate behind out_aln.write, AmpliPy.py:911).
 * host data -> BGZF blocks: block k = in[bstart[k], bstart[k + 1]) (bstart has n_blocks + 1 entries, bstart[0] = 0, the last one =
 * n_bytes, every block at most 0xff00 bytes -- the caller cuts at record boundaries as htslib does), followed by the EOF block.
 * One warp per block: LZ77 (4-byte hash + run candidate), a Huffman code built for every 8 k tokens (or the fixed code where that is
 * shorter), CRC-32 and ISIZE in the footer; a block that does not shrink is stored.  Any inflater reads the result; it is a little
 * smaller than zlib level 1 writes.
 * out must hold n_bytes + 31 * n_blocks + 28 bytes.  Returns the number of bytes written, or a negative AMP_ERR_* code. */
int64_t amp_bgzf_deflate_host(amp_ctx* ctx, const uint8_t* in, int64_t n_bytes, const int64_t* bstart, int64_t n_blocks, uint8_t* out,
                              int64_t out_cap);

/* The trimmed BAM file of the decoded batch, built on the device (replaces the cigartuples / reference_start assignments and
 * out_aln.write of AmpliPy.py:463-658, 910-911 for BAM input): after amp_bam_decode_host + amp_process_decoded(AMP_MODE_TRIM) the
 * inflated input and the trim outputs are both in HBM; the records of the reads that pass the write gate are rebuilt there (new
 * block_size / pos / bin / n_cigar / CIGAR, the rest byte for byte), cut into BGZF blocks at record boundaries, compressed
 * (amp_bgzf_deflate_host's kernels) and copied to the host once.  header = the BAM header the file starts with (magic, l_text,
 * text, references).  out must hold header_bytes + the inflated input + 31 bytes per 64 KB + 28.  Returns the size of the file
 * or a negative AMP_ERR_* code; *n_records = the number of records written. */
int64_t amp_decoded_write_bam(amp_ctx* ctx, const uint8_t* header, int64_t header_bytes, uint8_t* out, int64_t out_cap, int64_t* n_records);

/* ---- heterogeneous plates: a primer scheme and / or a reference of its own for one sample of the context.
 * amp_set_scheme builds the two per-position tables of find_overlapping_primers (AmpliPy.py:174-209) on the device from the
 * sorted primer list (start, end) and the offset; ref_len <= the context's ref_len.  amp_process_* of that sample then trims
 * against these tables, amp_call calls it against its own reference; one calling launch still covers the whole plate. */
int amp_set_scheme(amp_ctx* ctx, int sample, int32_t ref_len, const int32_t* primer_start, const int32_t* primer_end,
                   int32_t n_primers, int32_t offset);
int amp_get_scheme(amp_ctx* ctx, int sample, int32_t* min_primer_start, int32_t* max_primer_end, int32_t* max_primer_len);
int amp_set_sample_reference(amp_ctx* ctx, int sample, const char* ref_seq, int32_t len);

/* pinned host memory helpers for callers that want full-speed amp_process_host */
int amp_host_alloc(void** p, int64_t bytes);
int amp_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
