/* sahara_policy.h — the ONE table of reconstructed search rules.
 *
 * The arithmetic of `fmc::search_ng24::search<Edit>` (called at /root/reference/src/sahara/search.cpp:227-231) lives in
 * fmindex-collection 1.1.0, which is not vendored in the reference (cpm.dependencies:20-24).  Every rule of that
 * recursion that decides WHICH of several equivalent alignments is reported — and therefore bit-exactness of the
 * edit-distance hit multiset — is a reconstruction ([RECALL]).  They are collected here, as data, and
 * consumed by
 *     oracle/sahara_oracle.cpp                      (the CPU checker: Searcher::dir / Searcher::next)
 *     sahara_b200/csrc/search.cuh                   (fm_node, fm_ordered_thread, text_states, text_run,
 *                                                    build_state_flags)
 *     sahara_b200/host/scheme.hpp                   (expand: rule `expand_lower`)
 * so that a correction found by tools/pin_against_sahara.sh is a change of ONE initialiser below (or one call of
 * sb200_set_policy / orc_set_policy / sbh_set_expand_rule at run time).  tests/test_policy.py flips every switch and
 * checks that oracle and kernels move together — and that the flip changes the result, i.e. the switch is live.
 *
 * Plain C: included from C, C++ and CUDA.
 */
#ifndef SAHARA_POLICY_H
#define SAHARA_POLICY_H

#include <stdint.h>

/* last operation at one end of the match ("LInfo" / "RInfo" of the recursion) */
enum { SB200_INFO_M = 0, SB200_INFO_S = 1, SB200_INFO_I = 2, SB200_INFO_D = 3 };
#define SB200_INFO_BIT(i) (1u << (i))

typedef struct sb200_policy {
    /* bit i set: a DELETION (a text symbol the query does not have) may follow operation i at the same end.
     * Reconstruction: after a match or another deletion — never next to an insertion or a substitution. */
    uint32_t del_after;
    /* bit i set: an INSERTION (a query symbol the text does not have) may follow operation i at the same end.
     * Reconstruction: after a match or another insertion. */
    uint32_t ins_after;
    /* bit i set: a cursor is reported when operation i was the last one at an end; BOTH ends are checked.
     * Reconstruction: an alignment must not end in a substitution or a deletion at either end (M or I only). */
    uint32_t end_ok;
    /* order of the children of a node, which decides the first n rows of search_n (--max_hits):
     * the match child is always first; then
     *   bit 0 clear: per mismatching symbol the deletion before the substitution; set: substitution first
     *   bit 1 clear: the insertion after all symbols;                              set: insertion before them */
    uint32_t child_order;
    /* expansion of a scheme to the query length (fmc::search_scheme::expand, search.cpp:191): lower bound of the
     * characters of part i that are not its last one.  0: the lower bound of the previous part (the part's own bound
     * is only demanded at its last character); 1: the part's own lower bound at every character. */
    uint32_t expand_lower;
    uint32_t reserved[3];
} sb200_policy;

#define SB200_CHILD_SUB_BEFORE_DEL 1u
#define SB200_CHILD_INS_BEFORE_SYMBOLS 2u

/* THE reconstruction in force */
#define SB200_POLICY_DEFAULT                                                                                  \
    {                                                                                                         \
        /* del_after    */ SB200_INFO_BIT(SB200_INFO_M) | SB200_INFO_BIT(SB200_INFO_D),                       \
        /* ins_after    */ SB200_INFO_BIT(SB200_INFO_M) | SB200_INFO_BIT(SB200_INFO_I),                       \
        /* end_ok       */ SB200_INFO_BIT(SB200_INFO_M) | SB200_INFO_BIT(SB200_INFO_I),                       \
        /* child_order  */ 0u,                                                                                \
        /* expand_lower */ 0u,                                                                                \
        {0u, 0u, 0u}                                                                                          \
    }

#if defined(__CUDACC__)
#define SB200_POLICY_FN __host__ __device__ __forceinline__
#else
#define SB200_POLICY_FN static inline
#endif

/* may a deletion / an insertion follow operation `info` at the end that is extended? */
SB200_POLICY_FN int sb200_pol_del(const sb200_policy* p, uint32_t info) { return (int)((p->del_after >> info) & 1u); }
SB200_POLICY_FN int sb200_pol_ins(const sb200_policy* p, uint32_t info) { return (int)((p->ins_after >> info) & 1u); }
/* may a cursor whose ends carry (left, right) be reported? */
SB200_POLICY_FN int sb200_pol_end1(const sb200_policy* p, uint32_t info) { return (int)((p->end_ok >> info) & 1u); }
SB200_POLICY_FN int sb200_pol_end(const sb200_policy* p, uint32_t left, uint32_t right) {
    return (int)((p->end_ok >> left) & (p->end_ok >> right) & 1u);
}
/* The kernels fold the deletion (step, e+1, D) and the substitution (step+1, e+1, S) of a symbol into one PAIR frame
 * and expand neither half's insertion child; that is only right when no insertion may follow a D or an S. */
SB200_POLICY_FN int sb200_pol_pairs(const sb200_policy* p) {
    return (p->ins_after & (SB200_INFO_BIT(SB200_INFO_S) | SB200_INFO_BIT(SB200_INFO_D))) == 0u;
}
SB200_POLICY_FN int sb200_pol_valid(const sb200_policy* p) {
    return p->del_after < 16u && p->ins_after < 16u && p->end_ok < 16u && p->child_order < 4u && p->expand_lower < 2u;
}

#endif /* SAHARA_POLICY_H */
