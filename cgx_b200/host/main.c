/* strmatchcuda -- drop-in command line of the reference (Main.c:28-86): same getopt string "hl:t:s:",
 * exactly six positionals (source, query, target, alignment, lex file, output directory), help + exit(0)
 * otherwise.  Extensions that do not disturb the contract: -g <gpus>, -b <queries per batch>,
 * -w <writer threads>, -q (quiet), -i <index file> (persisted GPU index: loaded when the file exists -- alignment and
 * lexical files are then not parsed and no suffix array is built -- else built and saved there), -z <level> (gzip'd grammar
 * files, grammar.<qid>.s.gz), -S (server: after the command line's query file, "<query file> <output dir>" requests are read
 * from stdin and served against the resident index; one "done ..." line per request on stdout). */
#include "cgx_host.h"
#include <stdio.h>
#include <stdlib.h>
#include <unistd.h>

static void print_help(void) {
    printf("\nGPU source codes for gappy extraction. Please check your input arguments.\n\n"
           "usage: strmatchcuda [-l minmatchlen] [-t fingerlen] [-s timefile] [-g gpus] [-b batch] [-w threads] [-i index_file] [-z level] [-S] [-q]\n"
           "       <source_corpus> <query_file> <target_corpus> <alignment_file> <lex_file> <output_dir>\n");
    exit(0);
}

int main(int argc, char **argv) {
    cgxh_options_t o;
    int ch, errflg = 0;
    o.reffile = o.qryfile = o.reftargetfile = o.align = o.wordscdec = o.destinationDirectory = o.timefile = NULL;
    o.minmatchlen = 1; o.fingerlen = 10; o.n_gpus = 1; o.batch_queries = 0; o.writer_threads = 1; o.quiet = 0; o.index_file = NULL; o.serve = 0; o.gzip_level = 0;
    while (!errflg && (ch = getopt(argc, argv, "hl:t:s:g:b:w:i:z:Sq")) != -1) {
        switch (ch) {
        case 'h': print_help(); break;
        case 'l': o.minmatchlen = atoi(optarg); break;
        case 't': o.fingerlen = atoi(optarg); break;
        case 's': o.timefile = optarg; break;
        case 'g': o.n_gpus = atoi(optarg); break;
        case 'b': o.batch_queries = atoi(optarg); break;
        case 'w': o.writer_threads = atoi(optarg); break;
        case 'i': o.index_file = optarg; break;
        case 'q': o.quiet = 1; break;
        case 'z': o.gzip_level = atoi(optarg); break;
        case 'S': o.serve = 1; break;
        case '?': fprintf(stderr, "Unknown option %c\n", optopt); errflg = 1; break;
        default: errflg = 1; break;
        }
    }
    if (optind != argc - 6 || errflg) print_help();                           /* Main.c:46-48 */
    if (o.fingerlen > 10 || o.fingerlen <= 0) { fprintf(stderr, "finger length must be between 1 and 10\n"); exit(0); }   /* Main.c:50-53 */
    o.reffile = argv[optind++]; o.qryfile = argv[optind++]; o.reftargetfile = argv[optind++];
    o.align = argv[optind++]; o.wordscdec = argv[optind++]; o.destinationDirectory = argv[optind++];
    fprintf(stderr, "reference file: %s\nquery file: %s\nref target file: %s\nalign file: %s\nMinimum match length: %d\n", o.reffile, o.qryfile,
            o.reftargetfile, o.align, o.minmatchlen);                         /* Main.c:63-76 */
    return cgxh_run(&o);
}
