/* TEST INFRASTRUCTURE ONLY -- command-line front end of the CPU oracle with the reference's six
 * positional arguments (Main.c:46-60).  Writes grammar.<qid>.s files and prints the stage counts the
 * reference prints on stderr, so the two logs can be diffed. */
#include "cgx_oracle.h"
#include <stdio.h>
#include <time.h>

int main(int argc, char **argv) {
    if (argc != 7) { fprintf(stderr, "usage: %s <source> <query> <target> <alignment> <lex> <outdir>\n", argv[0]); return 2; }
    struct timespec t0, t1, t2;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    orc_t *o = orc_create_from_files(argv[1], argv[3], argv[4], argv[5]);
    if (!o) return 1;
    orc_build_sa(o);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    int rc = orc_run_query_file(o, argv[2]);
    clock_gettime(CLOCK_MONOTONIC, &t2);
    if (rc) { fprintf(stderr, "oracle: run failed rc=%d\n", rc); return 1; }
    orc_counts_t c; orc_get_counts(o, &c);
    fprintf(stderr, "toklen %d target %d queries %d tokens %d\n", c.n, c.m, c.Q, c.T);
    fprintf(stderr, "Found %d pairs!!\n", c.precomp_count);
    fprintf(stderr, "Found %d pairs for one gap enumeration!!\nDistinct one gap pattern: %d\nFound %d for one gap on SA!\n", c.enu1, c.D1, c.hits1);
    fprintf(stderr, "Found %d pairs for two gap enumeration!!\nDisticnt two gap pattern: %d\nFound %d two gap on SA\n", c.enu2, c.D2, c.hits2);
    fprintf(stderr, "Found %d cotinous pairs!!\nFound %d 1gap gappy pairs!!\nFound %d 2gap gappy pairs!!\n", c.n_ab, c.n_1gap_contig, c.n_2gap_contig);
    fprintf(stderr, "Found %d two gap phrase extraction pairs!!\n", c.n_axbxc);
    fprintf(stderr, "Found %d one gap phrase extraction pairs!!\nFound %d two gap phrase extraction pairs!!\n", c.n_axb, c.n_2gap_from1);
    fprintf(stderr, "Lexicon count for aXb, Xab, abX is %d\nLexicon count for aXbXc, XabX, XaXb, aXbX is %d\nLexicon count for continous ab is %d\n", c.lex_1gap, c.lex_2gap, c.lex_ab);
    fprintf(stderr, "distinct contiguous phrases %d\n", c.G);
    fprintf(stderr, "index %.3f s, match+extract %.3f s\n", (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec), (t2.tv_sec - t1.tv_sec) + 1e-9 * (t2.tv_nsec - t1.tv_nsec));
    rc = orc_write_grammars(o, argv[6]);
    orc_destroy(o);
    return rc;
}
