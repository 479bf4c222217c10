// TEST INFRASTRUCTURE ONLY.  C-ABI shim around the reference's own CPU suffix-array code
// (/root/reference/SuffixArray.c:196 suffixArrayConstruct -> :51 suffixArrayInt (DC3/skew),
// :143 buildLCPTable, :131 recursion_lcp).  Compiled together with that file, from where it
// lies, by oracle/build_ref.sh into oracle/_ref/libref_sa.so.  Used by tests/ as the bit-exact
// SA oracle and by bench.py as the "reference" CPU baseline for the SA-build metric.
#include "ComTypes.h"
#include "SuffixArray.h"
#include <time.h>

extern "C" {

// tokens: n+3 ints, the last three 0 (layout of Start.cu:321-327,354); `last` = largest symbol.
// sa_out: n ints.  aux_out (may be NULL): 4*n ints = lcpleft | lcpright | lcp | rank, the
// layout SuffixArray.c:145-171 packs into ref->buf.  Returns seconds spent.
double ref_sa_build(const int *tokens, int n, int last, int *sa_out, int *aux_out) {
    ref_t ref;
    memset(&ref, 0, sizeof(ref));
    int *tmp = (int *)malloc(sizeof(int) * ((size_t)n + 3));
    memcpy(tmp, tokens, sizeof(int) * ((size_t)n + 3));
    int *buf = aux_out ? aux_out : (int *)malloc(sizeof(int) * (size_t)n * 4 + 4);
    ref.buf = buf;
    ref.sa = sa_out;
    ref.str = (int *)malloc(sizeof(int) * ((size_t)n + 3));
    memcpy(ref.str, tokens, sizeof(int) * ((size_t)n + 3));
    ref.toklen = (unsigned int)n;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    suffixArrayConstruct(&ref, last, tmp);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    free(tmp);
    free(ref.str);
    if (!aux_out) free(buf);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

}  // extern "C"
