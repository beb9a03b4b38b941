/* sahara_b200.h — C ABI of the B200-native search path of sahara.
 *
 * The reference (seqan/sahara) has no FFI: the seam this library fills is the set of library calls made
 * by `runSearch` in /root/reference/src/sahara/search.cpp (all file:line below are relative to
 * /root/reference/).  A maintainer swaps those calls for the entry points declared here; INTEGRATION.md
 * shows the patch.  Everything is plain C: pointers and sizes, no C++ or torch types.
 *
 * Conventions
 *   - every function returns 0 on success and non-zero on failure; sb200_last_error() returns the
 *     message of the last failure on the calling thread (the reference throws error_fmt instead,
 *     src/sahara/utils/error_fmt.h:11-22).
 *   - host memory passed in stays owned by the caller; arrays returned through `**out` parameters are
 *     allocated by the library and released with sb200_free().
 *   - a context owns one GPU.  There is no CPU fallback: without a usable CUDA device
 *     sb200_create() fails.
 *   - ranks: 0 = '$' delimiter, 1..4 = A,C,G,T, 5 = N (only when sigma == 6).
 *
 * Limits the reference does not have (it computes in size_t, src/sahara/search.cpp:63-68); each is refused with a
 * clear error, none is silently truncated:
 *   - the index has fewer than 2^32 - 8192 rows (text + delimiters; a 3.1 Gbp genome has 3.1e9) — rows are u32 on the GPU;
 *   - at most 4 errors (search schemes with u <= 4), queries of at most 1000 characters, at most 255 searches per scheme;
 *   - fewer than 2^32 (query, search) pairs, cursors and hits per call, and fewer than 2^32 bytes of delta-coded hit records
 *     per batch (about 700 M hits; split the batch);
 *   - the 16-byte and packed hit formats (sb200_search_reads, sb200_submit_reads) need bits_for_position <= 32;
 *     sb200_search returns the reference's full 64-bit tuples.
 */
#ifndef SAHARA_B200_H
#define SAHARA_B200_H

#include <stddef.h>
#include <stdint.h>

#include "sahara_policy.h"

#ifdef __cplusplus
extern "C" {
#endif

#define SB200_ABI_VERSION 3

typedef struct sb200_ctx sb200_ctx;

/* ---- library / context -------------------------------------------------------------------------- */

int sb200_abi_version(void);
const char* sb200_last_error(void);
/* number of CUDA devices visible; fails when the driver is not usable */
int sb200_device_count(int* count);
/* one context per GPU (the reference is a single-threaded CPU program and has no equivalent) */
int sb200_create(int device, sb200_ctx** out);
int sb200_destroy(sb200_ctx* ctx);
/* launch all subsequent work on this cudaStream_t (NULL = the context's own stream) */
int sb200_set_stream(sb200_ctx* ctx, void* cuda_stream);
/* block until all work queued by this context has finished */
int sb200_synchronize(sb200_ctx* ctx);

/* ---- index ---------------------------------------------------------------------------------------
 * In-memory image of fmc::BiFMIndex<Sigma, fmc::string::InterleavedBitvector16> exactly as
 * `archive(index)` stores it (src/sahara/index.cpp:96-100, read back at src/sahara/search.cpp:162-169):
 * per 64 BWT rows one block of Sigma u16 in-superblock counters followed by Sigma one-hot u64 bitplanes
 * (10*Sigma bytes, unpadded); one superblock row of Sigma u64 every 65536 rows. */
typedef struct sb200_index_view {
    uint64_t sigma;              /* 5 (d_dna4) or 6 (d_dna5), src/sahara/search.cpp:284-287 */
    uint64_t n_rows;             /* text length including one delimiter per sequence */
    uint64_t n_blocks;           /* n_rows / 64 + 1 */
    const void* bwt_blocks;      /* n_blocks * 10*sigma bytes */
    const uint64_t* bwt_super;   /* ceil(n_blocks / 1024) * sigma */
    const void* bwtrev_blocks;   /* same for the BWT of the reversed text */
    const uint64_t* bwtrev_super;
    const uint64_t* C;           /* sigma + 1 entries */
    const uint64_t* ssa;         /* sampled suffix array: (seqId << bits_for_position) | seqPos, row order */
    uint64_t n_ssa;
    const uint64_t* mark_bits;   /* n_rows / 64 + 1 words, bit r set <=> row r is sampled */
    uint64_t sampling_rate;      /* 16 in the reference, src/sahara/index.cpp:87 */
    uint64_t bits_for_position;
} sb200_index_view;

/* replaces `archive(index)` into an fmc::BiFMIndex (src/sahara/search.cpp:162-169): copies the image to
 * the GPU and re-lays it into the device layout; verifies the image and fails with
 * "index layout not understood: ..." on any inconsistency. */
int sb200_index_upload(sb200_ctx* ctx, const sb200_index_view* view);

/* replaces `fmc::BiFMIndex<Sigma, InterleavedBitvector16>{ref, samplingRate, threads}`
 * (src/sahara/index.cpp:87): builds the index on the GPU from the concatenated ranks of all sequences
 * (no delimiters; seq_lens[n_seqs] gives the lengths). */
int sb200_index_build(sb200_ctx* ctx, const uint8_t* seq_ranks, const uint64_t* seq_lens, uint64_t n_seqs,
                      uint32_t sigma, uint32_t sampling_rate);

/* same, for a text that already lives on the GPU (device pointer to concatenated ranks) */
int sb200_index_build_device(sb200_ctx* ctx, const uint8_t* d_seq_ranks, const uint64_t* seq_lens, uint64_t n_seqs,
                             uint32_t sigma, uint32_t sampling_rate);

/* copies the index back in the reference image (for `archive(index)` when writing X.idx,
 * src/sahara/index.cpp:96-100).  All arrays of *out are library-allocated: release with
 * sb200_index_view_free(). */
int sb200_index_download(sb200_ctx* ctx, sb200_index_view* out);
void sb200_index_view_free(sb200_index_view* view);

typedef struct sb200_index_info {
    uint64_t sigma, n_rows, n_ssa, sampling_rate, bits_for_position;
    uint64_t device_sampling_rate; /* sampling rate of the device-side suffix array (<= sampling_rate) */
    uint64_t device_bytes;         /* HBM used by the index */
    uint64_t C[8];
} sb200_index_info;
int sb200_index_info_get(sb200_ctx* ctx, sb200_index_info* out);

/* re-samples the device suffix array to a denser rate (16, 8, 4, 2 or 1) by walking the LF mapping once;
 * results of locate are unchanged, LF steps per located row shrink. */
int sb200_index_densify(sb200_ctx* ctx, uint32_t device_sampling_rate);

/* in-text verification: builds the complete suffix array, its inverse and the packed text on the device
 * (about 17 bytes per row) so that cursors holding a single row are extended by reading the text instead of
 * probing the occurrence tables.  Results are unchanged.  enable = 0 releases the tables. */
int sb200_index_enable_text(sb200_ctx* ctx, int enable);

/* builds the q-gram jump table: the cursor of every string of `q` symbols over A,C,G,T, used to skip
 * the first error-free steps of a search (q = 0 removes it; q <= 15, 16 bytes x 4^q of device memory). */
int sb200_index_build_qgram(sb200_ctx* ctx, uint32_t q);

/* copies the index of `src` — occurrence tables, samples, and whatever derived tables it holds (verification tables,
 * q-gram table, densified samples) — into `dst`, GPU to GPU (cudaMemcpyPeerAsync: NVLink where the GPUs are peers).
 * One process driving several GPUs loads / derives the index once and replicates it; the reference loads its index once
 * per process too (/root/reference/src/sahara/search.cpp:162-169).  `src` must not change during the call; several
 * destinations may clone from the same source at the same time. */
int sb200_index_clone(sb200_ctx* dst, sb200_ctx* src);

/* ---- search scheme ---------------------------------------------------------------------------------
 * expanded scheme, one entry per query character: pi/l/u are [n_searches][len] row-major — the
 * `Scheme` handed to search_ng24::search (src/sahara/search.cpp:222-231).  edit != 0 selects
 * search<true> (Levenshtein), 0 selects search<false> (Hamming; apply limitToHamming beforehand as
 * src/sahara/search.cpp:226 does). */
int sb200_set_scheme(sb200_ctx* ctx, uint32_t n_searches, uint32_t len, const uint16_t* pi, const uint8_t* l,
                     const uint8_t* u, int edit);

/* The rules of the recursion that are reconstructions of fmindex-collection's behaviour (which operation may follow
 * which, what may be reported, order of the children, expansion of the lower bounds) live in ONE table,
 * include/sahara_policy.h; this call replaces the table in force (default: SB200_POLICY_DEFAULT).  The CPU oracle
 * consumes the same struct.  Meant for pinning against a real sahara binary (tools/pin_against_sahara.sh). */
int sb200_set_policy(sb200_ctx* ctx, const sb200_policy* policy);
int sb200_get_policy(sb200_ctx* ctx, sb200_policy* out);

/* knobs of the host orchestration; none of them changes results (they select between equivalent kernels / sort paths
 * or tune launch geometry and chunking; tests use them to reach every path).  Names: bucket_sort, fused_sort, textpos,
 * ordered_only, debug, chunk, edge_div, pool_blocks_per_sm, pool_threads, run_rounds, items_blocks_per_sm,
 * ordered_blocks_per_sm, overlap (how the batches in flight share the GPU: 0 one after the other, 1 unconstrained streams,
 * 2 = default: own streams, the persistent verification kernel of a batch starts when its predecessor batch is complete).  The library reads no environment variables. */
int sb200_set_option(sb200_ctx* ctx, const char* name, int64_t value);

/* replaces the maxHits argument of fmc::search_ng24::search_n<Edit>(index, queries, scheme, maxHits, res_cb)
 * (src/sahara/search.cpp:228,231; `--max_hits`, src/sahara/search.cpp:91-96).  max_hits > 0: every following search
 * call delivers at most max_hits suffix-array rows per query — the first ones in the order of the reference's
 * recursion (searches in scheme order; match, then the symbols ascending with deletion before substitution,
 * then insertion); the cursor that crosses the limit is cut to its first rows.  0 (default) = unlimited,
 * the plain search<Edit>. */
int sb200_set_max_hits(sb200_ctx* ctx, uint64_t max_hits);

/* ---- search + locate ------------------------------------------------------------------------------- */

/* mirrors std::tuple<size_t, fmc::LeftBiFMIndexCursor, size_t> (src/sahara/search.cpp:214-220) */
typedef struct sb200_cursor {
    uint64_t query_id, lb, len, errors;
} sb200_cursor;

/* mirrors the result tuple (queryId, seqId, seqPos+offset, e) (src/sahara/search.cpp:244-250) */
typedef struct sb200_hit {
    uint64_t query_id, seq_id, pos, errors;
} sb200_hit;

/* replaces fmc::search_ng24::search<Edit>(index, queries, scheme, res_cb) (src/sahara/search.cpp:227-231).
 * queries: n_queries * len ranks, dense, in the reference's order ([2i] = read i, [2i+1] = its reverse
 * complement).  Returns every reported cursor (the multiset the reference's callback receives), sorted
 * by (query_id, lb, len, errors). */
int sb200_search_cursors(sb200_ctx* ctx, const uint8_t* queries, uint64_t n_queries, uint32_t len,
                         sb200_cursor** cursors, uint64_t* n_cursors);

/* replaces the LocateLinear loop (src/sahara/search.cpp:244-250).  Hits come back sorted by
 * (query_id, seq_id, pos, errors); multiplicities are preserved. */
int sb200_locate(sb200_ctx* ctx, const sb200_cursor* cursors, uint64_t n_cursors, sb200_hit** hits,
                 uint64_t* n_hits);

/* search + locate in one call with everything kept on the GPU in between — the call `sahara search` makes */
int sb200_search(sb200_ctx* ctx, const uint8_t* queries, uint64_t n_queries, uint32_t len, sb200_hit** hits,
                 uint64_t* n_hits);

/* compact variant for callers that hold only the reads: the reverse complements (src/sahara/search.cpp:121-123,
 * skipped with --no-reverse) are made on the device and the hits come back as 16-byte records, which halves the
 * PCIe traffic of sb200_search.  query_id counts queries as the reference does (2i = read i, 2i+1 = its reverse
 * complement when with_reverse != 0, else i = read i).  Needs bits_for_position <= 32. */
typedef struct sb200_hit32 {
    uint32_t query_id, seq_id, pos, errors;
} sb200_hit32;
int sb200_search_reads(sb200_ctx* ctx, const uint8_t* reads, uint64_t n_reads, uint32_t len, int with_reverse,
                       sb200_hit32** hits, uint64_t* n_hits);

/* ---- asynchronous batches ---------------------------------------------------------------------------------
 * The same search + locate as sb200_search_reads, split into submit and wait so that consecutive batches overlap: while
 * batch i computes, the reads of batch i+1 travel to the GPU and the hits of batch i-1 travel back; nothing inside a
 * batch waits for the host.  Up to SB200_MAX_IN_FLIGHT batches may be submitted before the oldest is waited for.
 *
 *   reads   host memory (page-locked memory from sb200_host_alloc makes the copy a single DMA); it must stay valid until
 *           sb200_wait_batch returned.  format SB200_READS_RANKS: n_reads * len ranks, one byte each (the reference's
 *           std::vector<uint8_t> per read, src/sahara/search.cpp:115-124).  SB200_READS_PACKED4: 4 bits per base, 8 bases
 *           per little-endian 32-bit word, (len + 7) / 8 words per read (what sbh_pack_reads4 of the host library and the
 *           CLI's FASTA reader produce): half the PCIe bytes.  SB200_READS_PACKED2: 2 bits per base (A, C, G, T = 0 .. 3), 16
 *           bases per little-endian 32-bit word, (len + 15) / 16 words per read (sbh_pack_reads2, which refuses reads with
 *           any other symbol — send those batches in one of the other formats): a quarter of the bytes.
 *   result  hits in CSR form: the hits of query q are records [q ? hit_end[q-1] : 0, hit_end[q]); a record is
 *           record_bytes little-endian bytes holding ((seq_id << bits_for_position | pos) << 4) | errors, sorted by that
 *           value within a query.  query ids count as in the reference (2i = read i, 2i+1 = its reverse complement when
 *           with_reverse != 0).  5 bytes per hit for a human-sized genome instead of the reference's 32-byte tuple.
 *           Delta coding (the default; sb200_set_option(ctx, "delta_records", 0) turns it off): the hits of a query lie close
 *           together, so only the first record of a query is stored as above; every further one is the difference of its
 *           value to its predecessor's as a variable-length integer (7 bits per byte, low bits first, top bit = another
 *           byte follows) — about 1.3 bytes per hit.  hit_end[q] is then the end of query q's BYTES in `records`, and the
 *           number of its hits follows from decoding.  sbh_decode_records (libsahara_host) turns either form into tuples.
 *           The arrays belong to the batch and stay valid until sb200_release_batch(ticket). */
#define SB200_MAX_IN_FLIGHT 3
#define SB200_READS_RANKS 0
#define SB200_READS_PACKED4 1
#define SB200_READS_PACKED2 2
typedef struct sb200_batch_result {
    uint64_t n_queries, n_hits, n_cursors;
    const uint32_t* hit_end;   /* [n_queries] */
    const uint8_t* records;    /* n_hits * record_bytes */
    uint32_t record_bytes, bits_for_position;
    uint64_t h2d_bytes, d2h_bytes;          /* bytes this batch moved over PCIe */
    float ms_search, ms_locate, ms_sort;    /* CUDA-event times of the batch's kernels */
    uint32_t delta_coded;                   /* 1: `records` is delta coded and hit_end[] holds BYTE offsets (see above) */
    uint64_t n_record_bytes;                /* size of `records` */
} sb200_batch_result;
int sb200_submit_reads(sb200_ctx* ctx, const void* reads, uint64_t n_reads, uint32_t len, int format, int with_reverse,
                       uint64_t* ticket);
/* the same for queries that already sit in HBM (both strands, one byte per rank; device pointer) */
int sb200_submit_device(sb200_ctx* ctx, const uint8_t* d_queries, uint64_t n_queries, uint32_t len, uint64_t* ticket);
/* blocks until the batch is finished; copy_to_host != 0 also brings hit_end / records to page-locked host memory
 * (0: counts only, the hits stay in HBM) */
int sb200_wait_batch(sb200_ctx* ctx, uint64_t ticket, int copy_to_host, sb200_batch_result* out);
int sb200_release_batch(sb200_ctx* ctx, uint64_t ticket);

/* same pipeline with the queries already in HBM (d_queries = device pointer) and the hits left on the
 * GPU; returns only the counts.  Used to time the kernels without PCIe transfers. */
int sb200_search_device(sb200_ctx* ctx, const uint8_t* d_queries, uint64_t n_queries, uint32_t len,
                        uint64_t* n_cursors, uint64_t* n_hits);
/* copies the hits of the last sb200_search_device call to the host */
int sb200_fetch_hits(sb200_ctx* ctx, sb200_hit** hits, uint64_t* n_hits);

void sb200_free(void* p);

/* page-locked host memory for query batches (makes the host->device copy of sb200_search a single DMA);
 * released with sb200_free. */
int sb200_host_alloc(uint64_t bytes, void** out);

/* ---- rank / occ probe (kernel 1) -------------------------------------------------------------------
 * all_ranks at BWT rows: out[i*sigma + c] = number of symbols c in bwt[0, positions[i]).
 * which: 0 = bwt, 1 = bwtRev.  Parity hook for String::all_ranks of InterleavedBitvector16. */
int sb200_rank_probe(sb200_ctx* ctx, int which, const uint64_t* positions, uint64_t n, uint64_t* out);

/* micro-benchmark of the same device function: n_chains independent chains of `iters` dependent probes
 * at pseudo-random rows (seeded); returns elapsed milliseconds and a checksum of the ranks. */
int sb200_rank_bench(sb200_ctx* ctx, int which, uint64_t n_chains, uint32_t iters, uint64_t seed, float* ms,
                     uint64_t* checksum);

/* ---- counters -------------------------------------------------------------------------------------- */
typedef struct sb200_counters {
    uint64_t nodes;        /* cursor extensions executed by the search kernel */
    uint64_t rank_ops;     /* BWT rows probed (2 per extension) */
    uint64_t cursors;      /* cursors reported */
    uint64_t lf_steps;     /* LF steps walked by the locate kernel */
    uint64_t hits;         /* located positions */
    uint64_t kernel_launches; /* kernels launched by this context */
    float ms_search, ms_locate, ms_sort, ms_h2d, ms_d2h; /* last call, CUDA events */
    float ms_fm, ms_text;     /* last call: the walk over the occurrence tables (fm_roots_kernel + fm_items_kernel), the in-text verification (text_pool_kernel) */
    uint64_t nodes_text;      /* extensions verified in the text instead of the occurrence tables (subset of nodes) */
    uint64_t batch_restarts;  /* batches started over because a work buffer was too small (buffers grow, then stay) */
} sb200_counters;
int sb200_get_counters(sb200_ctx* ctx, sb200_counters* out);
int sb200_reset_counters(sb200_ctx* ctx);

/* ---- synthetic workload ---------------------------------------------
 * Deterministic counter-based generators, bit-identical to sahara_b200/synth.py.
 * genome: n_bases ranks in 1..4 written to d_out (device) or out (host). */
int sb200_synth_genome_device(sb200_ctx* ctx, uint64_t n_bases, uint64_t seed, uint8_t* d_out);
/* reads: n_reads * 2 queries of length len (read, reverse complement) sampled from the device-resident
 * genome with up to k errors (substitutions only when edit == 0), 10 % random reads. */
int sb200_synth_reads_device(sb200_ctx* ctx, const uint8_t* d_genome, uint64_t n_bases, uint64_t n_reads,
                             uint32_t len, uint32_t k, int edit, uint64_t seed, uint64_t first_read,
                             uint8_t* d_out);

/* raw device memory helpers for callers that have no CUDA runtime of their own */
int sb200_device_alloc(sb200_ctx* ctx, uint64_t bytes, void** d_ptr);
int sb200_device_free(sb200_ctx* ctx, void* d_ptr);
int sb200_copy_to_host(sb200_ctx* ctx, void* dst, const void* d_src, uint64_t bytes);
int sb200_copy_to_device(sb200_ctx* ctx, void* d_dst, const void* src, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* SAHARA_B200_H */
