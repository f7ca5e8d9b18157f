/* Minimal C (not C++) client of the lambda_b200 C ABI: open a reference-built .lba index, create the device copy,
 * search one batch of already-encoded queries, print BLAST tabular lines.
 *   gcc -std=c99 -Iinclude examples/minimal.c -Llambda_b200 -llambda_b200 -Wl,-rpath,$PWD/lambda_b200 -o minimal
 *   ./minimal db.lba
 * Without a CUDA device lgpu_index_create() fails with LGPU_ERR_CUDA (there is no CPU fallback); the program reports
 * that and exits with 3. */
#include <stdio.h>
#include <string.h>

#include "lambda_b200.h"

int main(int argc, char ** argv)
{
    lgpu_lba *              lba = NULL;
    lgpu_index *            ix  = NULL;
    lgpu_ctx *              ctx = NULL;
    lgpu_index_desc const * d;
    lgpu_params             p;
    /* two protein queries as aa27 ranks (A=0 ... Z=25, *=26), concatenated, with offsets */
    static char const * const text = "MKVLAAGIVGLLLAQPSAHAMKTAYIAKQRQISFVKSHFSRQLEERLGLIEVQ";
    uint8_t                   residues[64];
    uint64_t                  offsets[3];
    lgpu_query_batch          qb;
    lgpu_hits                 hits;
    lgpu_stats                st;
    size_t                    i, n = strlen(text);
    int                       rc;

    if (argc < 2)
    {
        fprintf(stderr, "usage: %s INDEX.lba\n", argv[0]);
        return 2;
    }
    printf("lambda_b200 ABI version %d\n", lgpu_version());
    if (lgpu_lba_open(&lba, argv[1]) != LGPU_OK)
    {
        fprintf(stderr, "%s\n", lgpu_last_error(NULL));
        return 1;
    }
    d = lgpu_lba_desc(lba);
    printf("index: %llu subjects, %llu residues\n", (unsigned long long)d->n_seqs, (unsigned long long)d->n_residues);
    for (i = 0; i < n; ++i)
        residues[i] = (uint8_t)(text[i] - 'A');
    offsets[0] = 0;
    offsets[1] = 21;
    offsets[2] = n;

    rc = lgpu_index_create(&ix, d, 0);
    if (rc != LGPU_OK)
    {
        fprintf(stderr, "lgpu_index_create: %d: %s\n", rc, lgpu_last_error(NULL));
        lgpu_lba_close(lba);
        return rc == LGPU_ERR_CUDA ? 3 : 1;
    }
    lgpu_params_default(&p, LGPU_DOMAIN_PROTEIN, "none");
    if (lgpu_ctx_create(&ctx, ix, &p) != LGPU_OK)
    {
        fprintf(stderr, "%s\n", lgpu_last_error(NULL));
        return 1;
    }
    memset(&qb, 0, sizeof(qb));
    qb.residues = residues;
    qb.offsets  = offsets;
    qb.n_queries = 2;
    memset(&st, 0, sizeof(st));
    if (lgpu_search_batch(ctx, &qb, &hits, &st) != LGPU_OK)
    {
        fprintf(stderr, "%s\n", lgpu_last_error(ctx));
        return 1;
    }
    for (i = 0; i < hits.n; ++i)
    {
        char        line[512], qid[32];
        char const * sid_b = d->ids + d->id_delims[hits.hits[i].s_id];
        char        sid[256];
        size_t      len = (size_t)(d->id_delims[hits.hits[i].s_id + 1] - d->id_delims[hits.hits[i].s_id]);
        if (len >= sizeof(sid))
            len = sizeof(sid) - 1;
        memcpy(sid, sid_b, len);
        sid[len] = 0;
        sprintf(qid, "query%u", (unsigned)hits.hits[i].q_id);
        if (lgpu_format_m8(&p, &hits.hits[i], qid, sid, line, sizeof(line)) > 0)
            fputs(line, stdout);
    }
    printf("%llu hits\n", (unsigned long long)hits.n);
    lgpu_ctx_destroy(ctx);
    lgpu_index_destroy(ix);
    lgpu_lba_close(lba);
    return 0;
}
