// TEST INFRASTRUCTURE ONLY.  Dump helper #included into an out-of-tree, instrumented copy of the
// reference (oracle/build_ref_dump.sh).  The instrumentation only adds fwrite calls after the
// reference has produced its arrays; no reference logic is changed.
#ifndef CGX_REF_DUMP_HOOKS_H
#define CGX_REF_DUMP_HOOKS_H
#include <stdio.h>
#include <stdlib.h>
static inline void cgx_dump(const char *name, const void *p, size_t bytes) {
    const char *dir = getenv("CGX_DUMP_DIR");
    if (!dir) return;
    char fn[4096];
    snprintf(fn, sizeof fn, "%s/%s.bin", dir, name);
    FILE *fh = fopen(fn, "wb");
    if (!fh) { fprintf(stderr, "cgx_dump: cannot write %s\n", fn); return; }
    if (bytes) fwrite(p, 1, bytes, fh);
    fclose(fh);
}
// wall-clock brackets (oracle/build_ref_dump.sh puts them around the reference's host aggregation calls)
#include <time.h>
static double cgx_tic_t_ = 0.0;
static inline double cgx_now_(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec; }
static inline void cgx_tic(void) { cgx_tic_t_ = cgx_now_(); }
static inline void cgx_toc(const char *name, long records) { fprintf(stderr, "cgx_ref_timer %s %.6f s %ld records\n", name, cgx_now_() - cgx_tic_t_, records); }
#endif
