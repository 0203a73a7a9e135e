/* csc.c -- CPU restatement of make_sysparse (TEST INFRASTRUCTURE, see oracle.h).
 *
 * make_sysparse.cpp:245-277 walks the kNN rows in order and, for the first k of the maxk entries of
 * row i, inserts the edge (min(i,j), max(i,j)) -> distance into a B-tree keyed by (from, to)
 * (compare_edge, mdsctk.cpp:567-577); Db::put overwrites, so the LAST insertion of an edge wins
 * (rows in increasing order, entries left to right; self edges i == j are skipped by both branches).
 * make_sysparse.cpp:287-329 then walks the tree in key order: column c holds the entries with
 * from == c (split_edges, mdsctk.cpp:580-604), rows ascending, and the output file is
 *     int n; int pcol[n+1]; int irow[nnz]; double val[nnz]
 * (the same layout CSC_matrix reads back, mdsctk.cpp:44-59).
 * Here the tree is replaced by a stable sort of (from, to, insertion order).
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

typedef struct { int from, to; long long order; double d; } csc_edge;

static int cmp_edge(const void *a, const void *b)
{
    const csc_edge *x = (const csc_edge *)a, *y = (const csc_edge *)b;
    if (x->from != y->from) return x->from < y->from ? -1 : 1;
    if (x->to != y->to) return x->to < y->to ? -1 : 1;
    return x->order < y->order ? -1 : (x->order > y->order ? 1 : 0);
}

/* idx [n][maxk], dist [n][maxk]; only the first k entries of a row are used.
 * pcol [n+1]; irow / val: capacity n*k entries.  Returns nnz, or -1. */
long long oracle_make_sysparse(const int *idx, const double *dist, long long n, int maxk, int k, int *pcol, int *irow,
                               double *val)
{
    if (n < 0 || k < 0 || k > maxk) return -1;
    csc_edge *e = (csc_edge *)malloc(sizeof(csc_edge) * (size_t)(n * k > 0 ? n * k : 1));
    if (!e) return -1;
    long long m = 0;
    for (long long i = 0; i < n; i++)
        for (int x = 0; x < k; x++) {
            const int j = idx[i * maxk + x];
            if (j == (int)i) continue;
            e[m].from = j < (int)i ? j : (int)i;
            e[m].to = j < (int)i ? (int)i : j;
            e[m].order = i * maxk + x;
            e[m].d = dist[i * maxk + x];
            m++;
        }
    qsort(e, (size_t)m, sizeof(csc_edge), cmp_edge);
    long long nnz = 0;
    memset(pcol, 0, sizeof(int) * (size_t)(n + 1));
    for (long long a = 0; a < m; a++) {
        if (a + 1 < m && e[a + 1].from == e[a].from && e[a + 1].to == e[a].to) continue; /* overwritten later */
        if (e[a].from < 0 || e[a].from >= n) continue; /* the reference's column walk never reaches such keys */
        irow[nnz] = e[a].to;
        val[nnz] = e[a].d;
        pcol[e[a].from + 1]++;
        nnz++;
    }
    for (long long c = 0; c < n; c++) pcol[c + 1] += pcol[c];
    free(e);
    return nnz;
}

/* make_gesparse.cpp:246-275: for the first k entries of row i, in order:
 *   put((i, j) -> d)                                   -- always, overwriting
 *   with -s: if (j, i) is absent, put((j, i) -> d)     -- db.get(...) == DB_NOTFOUND
 * then the same (from, to) walk.  The B-tree is simulated per key: the events of a key are replayed in
 * chronological order with exactly those put / put-if-absent semantics.  `order` doubles as the flag:
 * even = direct put, odd = fill attempt. */
long long oracle_make_gesparse(const int *idx, const double *dist, long long n, int maxk, int k, int symmetric, int *pcol,
                               int *irow, double *val)
{
    if (n < 0 || k < 0 || k > maxk) return -1;
    const long long cap = n * k * (symmetric ? 2 : 1);
    csc_edge *e = (csc_edge *)malloc(sizeof(csc_edge) * (size_t)(cap > 0 ? cap : 1));
    if (!e) return -1;
    long long m = 0;
    for (long long i = 0; i < n; i++)
        for (int x = 0; x < k; x++) {
            const int j = idx[i * maxk + x];
            e[m].from = (int)i; e[m].to = j; e[m].order = 2 * (i * maxk + x); e[m].d = dist[i * maxk + x];
            m++;
            if (symmetric) {
                e[m].from = j; e[m].to = (int)i; e[m].order = 2 * (i * maxk + x) + 1; e[m].d = dist[i * maxk + x];
                m++;
            }
        }
    qsort(e, (size_t)m, sizeof(csc_edge), cmp_edge);
    long long nnz = 0;
    memset(pcol, 0, sizeof(int) * (size_t)(n + 1));
    for (long long a = 0; a < m;) {
        long long b = a;
        int present = 0;
        double v = 0.0;
        for (; b < m && e[b].from == e[a].from && e[b].to == e[a].to; b++) {
            if ((e[b].order & 1) == 0) { present = 1; v = e[b].d; }          /* put: overwrite */
            else if (!present) { present = 1; v = e[b].d; }                   /* fill: only while absent */
        }
        if (present && e[a].from >= 0 && e[a].from < n) {
            irow[nnz] = e[a].to;
            val[nnz] = v;
            pcol[e[a].from + 1]++;
            nnz++;
        }
        a = b;
    }
    for (long long c = 0; c < n; c++) pcol[c + 1] += pcol[c];
    free(e);
    return nnz;
}
