/* cgx-b200 host: the driver behind strmatchcuda -- load, index, match + extract per query batch, write.
 * Replaces start() (Start.cu:488-629).  Queries are independent, so they are cut
 * into batches and, with n_gpus > 1, the batches are dealt round-robin to the GPUs of the box; the index
 * is built once on GPU 0 and broadcast to the peers over NVLink (cgx_index_broadcast, NCCL). */
#define _GNU_SOURCE
#include "cgx_host.h"
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

static double now_s(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

typedef struct {
    cgx_ctx_t *ctx;
    const cgxh_options_t *opt;
    const cgxh_queries_t *qry;
    const cgxh_side_t *src, *tgt;
    int gpu, n_gpus, batch;
    double t_gpu, t_write;
    int64_t rules, launches;
    int rc;
} worker_t;

/* one finished batch handed to the writer thread: host views of its results (valid for two more cgx_extract_begin) */
typedef struct {
    pthread_t th;
    int active, rc;
    cgx_result_t res;
    int32_t *off, q0;
    const worker_t *w;
    double t_write;
} wtask_t;

static void *writer_main(void *arg) {
    wtask_t *t = (wtask_t *)arg;
    double t0 = now_s();
    t->rc = cgxh_write_grammars_ex(t->w->opt->destinationDirectory, &t->res, t->off, t->q0, t->w->src, t->w->tgt, t->w->opt->writer_threads, t->w->opt->gzip_level);
    t->t_write = now_s() - t0;
    return NULL;
}

static int writer_join(wtask_t *t, worker_t *w) {
    if (!t->active) return 0;
    pthread_join(t->th, NULL);
    t->active = 0;
    free(t->off); t->off = NULL;
    w->t_write += t->t_write;
    return t->rc;
}

/* Batches of this GPU run as a three-deep pipeline (cgx_extract_begin): while batch k computes, the results of batch
 * k-1 are still travelling over PCIe and the grammars of batch k-2 are being written (the reference's "IO step",
 * README.md:76-79, PrintResults.c:434-574, off the critical path).  A batch the library refuses as too large
 * (CGX_E_BATCH_TOO_LARGE: hit lists grow with corpus size x batch size) is cut in two and both halves queued again. */
static void *worker_main(void *arg) {
    worker_t *w = (worker_t *)arg;
    const cgxh_queries_t *q = w->qry;
    int32_t n_batches = (q->Q + w->batch - 1) / w->batch;
    wtask_t wt;
    memset(&wt, 0, sizeof wt);
    wt.w = w;
    int32_t *prev_off = NULL, prev_q0 = 0;
    int have_prev = 0;
    /* work list of [q0,q1) ranges, processed in ascending order (a refused range is replaced by its halves) */
    int32_t cap = 64, top = 0;
    int32_t (*stack)[2] = (int32_t (*)[2])malloc(sizeof(int32_t[2]) * (size_t)cap);
    for (int32_t bi = w->gpu; bi < n_batches; bi += w->n_gpus) {
        int32_t q0 = bi * w->batch, q1 = q0 + w->batch > q->Q ? q->Q : q0 + w->batch;
        stack[0][0] = q0; stack[0][1] = q1; top = 1;
        while (top > 0 && !w->rc) {
            top--;
            q0 = stack[top][0]; q1 = stack[top][1];
            int32_t fit = cgx_batch_advice(w->ctx, q1 - q0);          /* from the hits per query of the last batch: at most one refusal per stream */
            if (fit < q1 - q0) {
                if (top + 1 > cap) { cap *= 2; stack = (int32_t (*)[2])realloc(stack, sizeof(int32_t[2]) * (size_t)cap); }
                stack[top][0] = q0 + fit; stack[top][1] = q1; top++;
                q1 = q0 + fit;
            }
            int32_t nq = q1 - q0, base = q->off[q0];
            int32_t *off = (int32_t *)malloc(sizeof(int32_t) * ((size_t)nq + 1));
            for (int32_t i = 0; i <= nq; i++) off[i] = q->off[q0 + i] - base;
            double t0 = now_s();
            int rc = cgx_extract_begin(w->ctx, q->tok + base, off, nq);
            if (rc == CGX_E_BATCH_TOO_LARGE && nq > 1) {
                if (!w->opt->quiet) fprintf(stderr, "[gpu %d] queries %d..%d: %s -- splitting\n", w->gpu, q0, q1 - 1, cgx_last_error(w->ctx));
                free(off);
                if (top + 2 > cap) { cap *= 2; stack = (int32_t (*)[2])realloc(stack, sizeof(int32_t[2]) * (size_t)cap); }
                int32_t mid = q0 + nq / 2;
                stack[top][0] = mid; stack[top][1] = q1; top++;          /* second half below the first: ascending order */
                stack[top][0] = q0; stack[top][1] = mid; top++;
                w->t_gpu += now_s() - t0;
                continue;
            }
            if (rc) { fprintf(stderr, "cgx_extract_begin: %s\n", cgx_last_error(w->ctx)); w->rc = 1; free(off); break; }
            cgx_batch_info_t bi_info;
            cgx_batch_info(w->ctx, &bi_info);
            w->t_gpu += now_s() - t0;
            w->rules += (int64_t)bi_info.rules[0] + bi_info.rules[1] + bi_info.rules[2];
            w->launches += bi_info.launches;
            if (!w->opt->quiet)
                fprintf(stderr, "[gpu %d] queries %d..%d: phrases %d, aXb patterns %d (%lld hits), aXbXc patterns %d (%lld hits), rules %d/%d/%d, device %.3f ms\n",
                        w->gpu, q0, q1 - 1, bi_info.G, bi_info.D1, (long long)bi_info.hits1, bi_info.D2, (long long)bi_info.hits2, bi_info.rules[0],
                        bi_info.rules[1], bi_info.rules[2], bi_info.ms_total);
            if (writer_join(&wt, w)) { w->rc = 1; free(off); break; }        /* batch k-2 written (it ran beside this batch's kernels) */
            if (have_prev) {                                                  /* batch k-1 is on the host by now: hand it to the writer */
                if (cgx_result_at(w->ctx, 1, &wt.res)) { fprintf(stderr, "cgx_result_at: %s\n", cgx_last_error(w->ctx)); w->rc = 1; free(off); break; }
                wt.off = prev_off; wt.q0 = prev_q0; wt.active = 1; prev_off = NULL;
                pthread_create(&wt.th, NULL, writer_main, &wt);
            }
            prev_off = off; prev_q0 = q0; have_prev = 1;
        }
        if (w->rc) break;
    }
    free(stack);
    if (writer_join(&wt, w)) w->rc = 1;
    if (have_prev && prev_off && !w->rc) {                                /* the last batch */
        double t0 = now_s();
        if (cgx_result_at(w->ctx, 0, &wt.res)) { fprintf(stderr, "cgx_result_at: %s\n", cgx_last_error(w->ctx)); w->rc = 1; }
        w->t_gpu += now_s() - t0;
        t0 = now_s();
        if (!w->rc && cgxh_write_grammars_ex(w->opt->destinationDirectory, &wt.res, prev_off, prev_q0, w->src, w->tgt, w->opt->writer_threads, w->opt->gzip_level)) w->rc = 1;
        w->t_write += now_s() - t0;
    }
    free(prev_off);
    return NULL;
}

/* one query file against the resident index: load the queries, run the batches on the GPUs, write the grammars */
typedef struct {
    const cgxh_options_t *opt;
    cgx_ctx_t **ctx;
    int n_gpus;
    const cgxh_side_t *src, *tgt;
    double t_load, t_index, t_begin;
    int last_Q;
} serve_t;

static int serve_one(serve_t *sv, const char *qryfile, const char *outdir) {
    const cgxh_options_t *opt = sv->opt;
    cgxh_options_t o = *opt;                 /* the workers read the output directory from their options */
    o.destinationDirectory = outdir;
    const int n_gpus = sv->n_gpus;
    cgxh_queries_t qry;
    if (cgxh_queries_load(qryfile, sv->src, &qry)) return 1;
    fprintf(stderr, "\nMax length of queries is %d\n", qry.max_len);
    sv->last_Q = qry.Q;
    double t2 = now_s();
    fprintf(stderr, "Start Extract Pair\n");
    /* queries per batch: -b, else an even split over the GPUs capped at CGXH_DEFAULT_BATCH (the per-batch hit lists grow
     * with corpus size x batch size; the reference needs <= ~5 k queries per process at its preallocation sizes) */
    int batch = opt->batch_queries > 0 ? opt->batch_queries : (qry.Q > 0 ? (qry.Q + n_gpus - 1) / n_gpus : 1);
    if (opt->batch_queries <= 0 && batch > CGXH_DEFAULT_BATCH) batch = CGXH_DEFAULT_BATCH;
    worker_t *w = (worker_t *)calloc((size_t)n_gpus, sizeof(worker_t));
    pthread_t *th = (pthread_t *)calloc((size_t)n_gpus, sizeof(pthread_t));
    for (int g = 0; g < n_gpus; g++) {
        w[g].ctx = sv->ctx[g]; w[g].opt = &o; w[g].qry = &qry; w[g].src = sv->src; w[g].tgt = sv->tgt; w[g].gpu = g; w[g].n_gpus = n_gpus; w[g].batch = batch;
        if (n_gpus == 1) worker_main(&w[g]); else pthread_create(&th[g], NULL, worker_main, &w[g]);
    }
    int rc = 0;
    double t_gpu = 0, t_write = 0;
    int64_t rules = 0;
    for (int g = 0; g < n_gpus; g++) {
        if (n_gpus > 1) pthread_join(th[g], NULL);
        rc |= w[g].rc;
        if (w[g].t_gpu > t_gpu) t_gpu = w[g].t_gpu;
        if (w[g].t_write > t_write) t_write = w[g].t_write;
        rules += w[g].rules;
    }
    double t3 = now_s();
    fprintf(stderr, "Start Printing Gappy Phrases...\n");   /* the reference's completion marker (README.md:76-79) */
    fprintf(stderr, "loading %.3f s, index %.3f s, match+extract %.3f s (max over %d GPU%s), grammar writing %.3f s, total %.3f s; %lld rules; %.1f query sentences/s\n",
            sv->t_load, sv->t_index, t_gpu, n_gpus, n_gpus > 1 ? "s" : "", t_write, t3 - sv->t_begin, (long long)rules, qry.Q / (t3 - t2 > 0 ? t3 - t2 : 1e-9));
    if (opt->timefile) {
        FILE *fh = fopen(opt->timefile, "a");
        if (fh) {
            fprintf(fh, "total: %f , load: %f , index: %f , extract: %f , write: %f , gpus: %d , queries: %d\n", t3 - sv->t_begin, sv->t_load, sv->t_index, t_gpu, t_write, n_gpus, qry.Q);
            fclose(fh);
        }
    }
    free(w); free(th);
    cgxh_queries_free(&qry);
    return rc;
}

/* device side of the start-up, run beside the text loaders: the CUDA contexts (a few hundred ms of driver start-up) and, when a
 * persisted index exists, reading it into HBM */
typedef struct {
    const cgxh_options_t *opt;
    cgx_ctx_t **ctx;
    int n_gpus, have_index, rc;
    char err[600];
} gpu_open_t;
static void *gpu_open_main(void *arg) {
    gpu_open_t *g = (gpu_open_t *)arg;
    for (int d = 0; d < g->n_gpus; d++)
        if (cgx_create(d, &g->ctx[d])) { snprintf(g->err, sizeof g->err, "cgx_create(%d): %s", d, cgx_last_error(NULL)); g->rc = 1; return NULL; }
    if (g->have_index && cgx_index_load(g->ctx[0], g->opt->index_file)) {
        snprintf(g->err, sizeof g->err, "cgx_index_load: %s", cgx_last_error(g->ctx[0]));
        g->rc = 1;
    }
    return NULL;
}

int cgxh_run(const cgxh_options_t *opt) {
    cgxh_side_t src, tgt;
    cgxh_align_t al;
    cgxh_lex_t lex;
    double t0 = now_s();
    int have_index = 0;
    if (opt->index_file) { FILE *fh = fopen(opt->index_file, "rb"); if (fh) { have_index = 1; fclose(fh); } }
    int n_gpus = opt->n_gpus > 0 ? opt->n_gpus : 1;
    cgx_ctx_t **ctx = (cgx_ctx_t **)calloc((size_t)n_gpus, sizeof(cgx_ctx_t *));
    gpu_open_t go;
    memset(&go, 0, sizeof go);
    go.opt = opt; go.ctx = ctx; go.n_gpus = n_gpus; go.have_index = have_index;
    pthread_t go_th;
    const int go_threaded = pthread_create(&go_th, NULL, gpu_open_main, &go) == 0;
    fprintf(stderr, "\nLoading the reference\n");
    /* source || target, then alignment || lexical file; with a persisted index the last two are not parsed */
    const int load_rc = cgxh_load_files(opt->reffile, opt->reftargetfile, have_index ? NULL : opt->align, have_index ? NULL : opt->wordscdec, &src, &tgt, &al, &lex);
    double t1 = now_s();
    if (go_threaded) pthread_join(go_th, NULL); else gpu_open_main(&go);
    if (load_rc) return 1;
    fprintf(stderr, "Reference toklen number is %lld HASH_COUNT %d\n", (long long)src.n, cgxh_vocab_size(src.vocab) - 2);
    fprintf(stderr, "Target Reference toklen number is %lld HASH_COUNT TARGET %d\n", (long long)tgt.n, cgxh_vocab_size(tgt.vocab) - 2);
    if (!have_index) fprintf(stderr, "Lex File Word Possibility COUNTER: %lld\n", (long long)lex.count);
    if (go.rc) { fprintf(stderr, "%s\n", go.err); return 1; }
    if (have_index) {
        fprintf(stderr, "index loaded from %s\n", opt->index_file);
    } else {
        if (al.wide) fprintf(stderr, "sentences of 255 tokens and more: 16-bit alignment fields\n");
        if (al.wide ? cgx_index_build_wide(ctx[0], src.tok, src.n, tgt.tok, tgt.n, al.RLP64, al.L_tar16, al.R_tar16)
                    : cgx_index_build(ctx[0], src.tok, src.n, tgt.tok, tgt.n, al.RLP, al.L_tar, al.R_tar)) { fprintf(stderr, "cgx_index_build: %s\n", cgx_last_error(ctx[0])); return 1; }
        if (cgx_lex_load(ctx[0], lex.f, lex.e, lex.v1, lex.v2, lex.count)) { fprintf(stderr, "cgx_lex_load: %s\n", cgx_last_error(ctx[0])); return 1; }
        if (opt->index_file) {
            if (cgx_index_save(ctx[0], opt->index_file)) { fprintf(stderr, "cgx_index_save: %s\n", cgx_last_error(ctx[0])); return 1; }
            fprintf(stderr, "index saved to %s\n", opt->index_file);
        }
    }
    cgx_index_info_t ii;
    cgx_index_info(ctx[0], &ii);
    if (ii.n != src.n || ii.m != tgt.n || !cgx_index_matches(ctx[0], src.tok, src.n, tgt.tok, tgt.n)) { fprintf(stderr, "index file %s was built from another corpus (%lld / %lld tokens, corpus has %lld / %lld)\n", opt->index_file ? opt->index_file : "?", (long long)ii.n, (long long)ii.m, (long long)src.n, (long long)tgt.n); return 1; }
    fprintf(stderr, "SA Construction %.4f sec (GPU prefix doubling, %d rounds, %d-bit keys); auxiliary index %.4f sec; %.1f MB resident\n",
            ii.sa_build_ms / 1e3, ii.sa_rounds, ii.sa_key_bits, ii.aux_build_ms / 1e3, (double)ii.index_bytes / 1048576.0);
    if (n_gpus > 1 && cgx_index_broadcast(ctx, n_gpus)) { fprintf(stderr, "cgx_index_broadcast: %s\n", cgx_last_error(ctx[0])); return 1; }
    double t2 = now_s();

    /* the query file of the command line, then -- in server mode (-S) -- one request per line of stdin: "<query file> <output dir>",
     * answered on stdout when its grammar files are written.  The corpus, the suffix array and the lexical table stay resident in
     * HBM between requests (the reference reloads and rebuilds everything for every query file, Start.cu:488-629). */
    serve_t sv;
    sv.opt = opt; sv.ctx = ctx; sv.n_gpus = n_gpus; sv.src = &src; sv.tgt = &tgt; sv.t_load = t1 - t0; sv.t_index = t2 - t1; sv.t_begin = t0;
    int rc = serve_one(&sv, opt->qryfile, opt->destinationDirectory);
    if (opt->serve) {
        char *line = NULL;
        size_t cap_line = 0;
        sv.t_load = sv.t_index = 0.0;
        while (getline(&line, &cap_line, stdin) > 0) {
            char *save = NULL;
            char *qf = strtok_r(line, " \t\r\n", &save), *od = strtok_r(NULL, " \t\r\n", &save);
            if (!qf) continue;                                   /* empty line */
            if (!strcmp(qf, "quit")) break;
            if (!od) { printf("error %s: expected \"<query file> <output dir>\"\n", qf); fflush(stdout); continue; }
            sv.t_begin = now_s();
            int r1 = serve_one(&sv, qf, od);
            if (r1) printf("error %s\n", qf);
            else printf("done %s %d queries %.3f s\n", qf, sv.last_Q, now_s() - sv.t_begin);
            fflush(stdout);
        }
        free(line);
    }
    for (int g = 0; g < n_gpus; g++) cgx_destroy(ctx[g]);
    free(ctx);
    cgxh_side_free(&src); cgxh_side_free(&tgt); cgxh_align_free(&al); cgxh_lex_free(&lex);
    return rc;
}
