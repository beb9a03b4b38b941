/* sahara_host.h — C entry points of the host-side helpers (no CUDA inside): search-scheme generators,
 * index file reader/writer, FASTA + alphabet conversion.  They mirror the non-kernel calls `sahara search`
 * makes around the hot path (file:line relative to /root/reference/):
 *   generator lookup + expand + limitToHamming      src/sahara/search.cpp:174-212, 226
 *   archive(sigma); archive(index)                   src/sahara/search.cpp:162-169, src/sahara/index.cpp:96-100
 *   fasta reader, char->rank, reverse complement     src/sahara/search.cpp:115-124
 * Used by the C++ CLI directly (headers in sahara_b200/host/) and by the Python tests through ctypes.
 * All functions return 0 on success; sbh_last_error() holds the message otherwise. */
#ifndef SAHARA_HOST_H
#define SAHARA_HOST_H
#include <stddef.h>
#include <stdint.h>
#include "sahara_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

const char* sbh_last_error(void);

/* ---- search schemes ---- */
/* comma separated list of generator names */
const char* sbh_scheme_names(void);
/* generates scheme `name` for minK..maxK errors, expands it to `len` characters (len == 0: not expanded,
 * one entry per part), optionally applies limitToHamming.  Outputs are library-allocated
 * [n_searches][n_entries] tables released with sbh_free. */
int sbh_scheme_generate(const char* name, int minK, int maxK, uint32_t len, int limit_to_hamming, uint32_t* n_searches,
                        uint32_t* n_entries, uint16_t** pi, uint8_t** l, uint8_t** u);
/* `--dynamic_generator` (src/sahara/search.cpp:193-195, 203-205): the parts are sized by optimizeByWNCTopDown<Edit>(scheme,
 * len, sigma, ref_len, 1) instead of len / parts each; limitToHamming is applied when edit == 0 (search.cpp:226).
 * partition (optional, partition_cap entries) receives the part sizes, n_parts their number. */
int sbh_scheme_generate_dynamic(const char* name, int minK, int maxK, uint32_t len, int edit, uint64_t sigma, uint64_t ref_len,
                                uint32_t* n_searches, uint32_t* n_entries, uint16_t** pi, uint8_t** l, uint8_t** u, uint32_t* partition,
                                uint32_t partition_cap, uint32_t* n_parts);
/* same from a Columba-format text ("{pi} {L} {U}" per line, 0-based) */
int sbh_scheme_from_columba(const char* text, uint32_t len, int limit_to_hamming, uint32_t* n_searches, uint32_t* n_entries,
                            uint16_t** pi, uint8_t** l, uint8_t** u);
/* checks on an unexpanded scheme given as tables [n_searches][parts] */
int sbh_scheme_check(uint32_t n_searches, uint32_t parts, const uint16_t* pi, const uint8_t* l, const uint8_t* u, int minK, int maxK,
                     int* valid, int* complete, int* non_redundant);
/* node counts of an expanded scheme (printed by `sahara search`, src/sahara/search.cpp:197-198) */
int sbh_scheme_node_count(uint32_t n_searches, uint32_t len, const uint16_t* pi, const uint8_t* l, const uint8_t* u, int edit,
                          uint64_t sigma, uint64_t ref_len, double* node_count, double* weighted_node_count);

/* ---- index file ---- */
int sbh_idx_peek_sigma(const char* path, uint64_t* sigma);
/* loads X.idx; arrays of *out are owned by the returned handle (release with sbh_idx_free) */
int sbh_idx_load(const char* path, sb200_index_view* out, void** handle);
void sbh_idx_free(void* handle);
int sbh_idx_save(const char* path, const sb200_index_view* view);

/* ---- FASTA / alphabet ---- */
/* reads all records, converts to ranks (sigma 5: d_dna4, 6: d_dna5); concatenated ranks + lengths.
 * with_revcomp != 0: every record is followed by its reverse complement (the query order of
 * src/sahara/search.cpp:121-123).  Invalid characters fail with the reference's message. */
int sbh_fasta_load_ranks(const char* path, uint64_t sigma, int with_revcomp, uint8_t** ranks, uint64_t** lens, uint64_t* n_seqs);
/* the read set of `sahara search` (src/sahara/search.cpp:115-124) converted to ranks by `threads` host threads: all
 * records must have the length of the first one; fails with the reference's message on an invalid character
 * ("query '<id>' (<n>) has invalid character at position <p> '<c>'(<hex>)", n = 1-based record number) and with a
 * clear message on a record of another length — whichever comes first in the file.  ranks: n_reads * len bytes,
 * released with sbh_free. */
int sbh_fasta_load_reads(const char* path, uint64_t sigma, uint32_t threads, uint8_t** ranks, uint64_t* n_reads, uint64_t* len);
int sbh_revcomp_ranks(const uint8_t* in, uint64_t n, uint8_t* out);
/* ranks (one byte per base, values 0..15) -> SB200_READS_PACKED4 (include/sahara_b200.h): 4 bits per base, 8 bases per
 * little-endian 32-bit word, (len + 7) / 8 words per read, unused nibbles of a read's last word 0xF.  out: n_reads *
 * ((len + 7) / 8) words (page-locked memory from sb200_host_alloc makes the copy to the GPU a single DMA). */
int sbh_pack_reads4(const uint8_t* ranks, uint64_t n_reads, uint32_t len, uint32_t threads, uint32_t* out);
/* ranks -> SB200_READS_PACKED2: 2 bits per base (ranks 1 .. 4 = A, C, G, T -> 0 .. 3), 16 bases per little-endian 32-bit word,
 * (len + 15) / 16 words per read.  Fails (and names the read) when a read holds any other symbol. */
int sbh_pack_reads2(const uint8_t* ranks, uint64_t n_reads, uint32_t len, uint32_t threads, uint32_t* out);
/* CSR result of sb200_wait_batch (include/sahara_b200.h: hit_end, records, record_bytes, bits_for_position, delta_coded) ->
 * tuples: out[4 * i + 0..3] = queryId (first_query + q), seqId, pos, errors of hit i, in the order of the records; out holds
 * 4 * n_hits values.  Fails when the records do not decode to exactly n_hits hits. */
int sbh_decode_records(const uint32_t* hit_end, const uint8_t* records, uint64_t n_queries, uint64_t n_hits, uint32_t record_bytes,
                       uint32_t bits_for_position, int delta_coded, uint64_t first_query, uint64_t* out);
/* rule `expand_lower` of the policy table (include/sahara_policy.h) used by every expansion that follows (default 0) */
int sbh_set_expand_rule(uint32_t rule);

void sbh_free(void* p);

#ifdef __cplusplus
}
#endif
#endif
