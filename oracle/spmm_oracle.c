/*
 * oracle/spmm_oracle.c -- CPU restatement of the reference's CSR x dense SpMM path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (ge-spmm_b200/, include/) may
 * link, import or execute this file.  It is used by tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs, as the checker and as the
 * reported CPU baseline -- never as the thing shipped.
 *
 * Parity pin: validated against the reference's own code compiled from
 * /root/reference (oracle/_ref, see oracle/Makefile): the reference .mtx reader
 * (util/util.hpp readMtx) on the three bundled matrices and on hand-made edge
 * cases, and -- on the GPU box -- the reference kernels (pytorch-custom/
 * spmm_kernel.cu, spmm_test.cu) bit-for-bit.  Fixtures: tests/golden/.
 *
 * Every function cites the reference file:line it follows.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <ctype.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------
 * SpMM, valued.  Follows the CPU golden loop of the CLI (spmm_test.cu:595-605):
 *     acc = 0; for ptr in row: acc += A_data[ptr] * B[K*A_indices[ptr] + k]
 * and the per-element order of every valued GPU kernel (spmm_kernel.cu:210-379,
 * spmm_test.cu:64-454): one fp32 accumulator per output element, CSR order, start 0.
 * use_fma=1 restates what nvcc emits for `acc += val*B[..]` on the device (one FFMA);
 * use_fma=0 restates the host golden (separately rounded multiply, then add).
 * ldb/ldc generalise the row stride (reference: ldb = ldc = K).
 * ---------------------------------------------------------------------------------- */
void oracle_spmm_valued_f32(int64_t M, int64_t K, const int32_t *rowptr, const int32_t *colind,
                            const float *val, const float *B, int64_t ldb, float *C, int64_t ldc,
                            int use_fma, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256) if (nthreads != 1)
#endif
    for (int64_t i = 0; i < M; i++) {
        float *crow = C + i * ldc;
        for (int64_t k = 0; k < K; k++) crow[k] = 0.0f;
        for (int32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
            const float a = val[p];
            const float *brow = B + (int64_t)colind[p] * ldb;
            if (use_fma) {
                for (int64_t k = 0; k < K; k++) crow[k] = fmaf(a, brow[k], crow[k]);
            } else {
                /* built with -ffp-contract=off: multiply and add round separately */
                for (int64_t k = 0; k < K; k++) crow[k] = crow[k] + a * brow[k];
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * SpMM, unvalued (A treated as all-ones).  Follows sum_init()/sum_reduce()
 * (spmm_kernel.cu:23-29) as used by topoSimple/topoCache/topoCacheCoarsen
 * (spmm_kernel.cu:56-66, 123-129, 164-171): acc = 0; acc = acc + B[colind*k + cid].
 * ---------------------------------------------------------------------------------- */
void oracle_spmm_unvalued_f32(int64_t M, int64_t K, const int32_t *rowptr, const int32_t *colind,
                              const float *B, int64_t ldb, float *C, int64_t ldc, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256) if (nthreads != 1)
#endif
    for (int64_t i = 0; i < M; i++) {
        float *crow = C + i * ldc;
        for (int64_t k = 0; k < K; k++) crow[k] = 0.0f;
        for (int32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
            const float *brow = B + (int64_t)colind[p] * ldb;
            for (int64_t k = 0; k < K; k++) crow[k] = crow[k] + brow[k];
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Max-reduce variant.  Follows max_reduce()/max_init() and the kernel loops of the DGL patch
 * (dgl-custom/binary_reduce_max.cu:18-24, 26-168): acc = init; acc = acc > x ? acc : x over the row in
 * CSR order; empty rows yield init (the reference's init is -10000).  val == NULL: x = B[col, k].
 * ---------------------------------------------------------------------------------- */
void oracle_spmm_max_f32(int64_t M, int64_t K, const int32_t *rowptr, const int32_t *colind, const float *val,
                         const float *B, int64_t ldb, float *C, int64_t ldc, float init, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256) if (nthreads != 1)
#endif
    for (int64_t i = 0; i < M; i++) {
        float *crow = C + i * ldc;
        for (int64_t k = 0; k < K; k++) crow[k] = init;
        for (int32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
            const float *brow = B + (int64_t)colind[p] * ldb;
            for (int64_t k = 0; k < K; k++) {
                const float x = val ? val[p] * brow[k] : brow[k];
                crow[k] = crow[k] > x ? crow[k] : x;
            }
        }
    }
}

/* fp64 golden of the same product (not in the reference; used to bound fp32 error). */
void oracle_spmm_f64(int64_t M, int64_t K, const int32_t *rowptr, const int32_t *colind,
                     const float *val /* nullable */, const float *B, int64_t ldb, double *C,
                     double *rowabs /* nullable: sum |a||b| per element */, int nthreads)
{
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
#pragma omp parallel for schedule(dynamic, 256) if (nthreads != 1)
#endif
    for (int64_t i = 0; i < M; i++) {
        double *crow = C + i * K;
        double *arow = rowabs ? rowabs + i * K : NULL;
        for (int64_t k = 0; k < K; k++) { crow[k] = 0.0; if (arow) arow[k] = 0.0; }
        for (int32_t p = rowptr[i]; p < rowptr[i + 1]; p++) {
            const double a = val ? (double)val[p] : 1.0;
            const float *brow = B + (int64_t)colind[p] * ldb;
            for (int64_t k = 0; k < K; k++) {
                crow[k] += a * (double)brow[k];
                if (arow) arow[k] += fabs(a * (double)brow[k]);
            }
        }
    }
}

/* ------------------------------------------------------------------------------------
 * Which reference kernel a given K lands in (spmm_kernel.cu:186,193,200 / 437,444,451).
 * Returns 0 = simple (k<32), 1 = cache (32<=k<64), 2 = cacheCoarsen (k>=64).  Also the
 * launch shape, for documentation/tests of the dispatch table.
 * ---------------------------------------------------------------------------------- */
int oracle_ref_dispatch(int64_t m, int64_t k, int64_t grid[2], int64_t block[2], int64_t *smem_unvalued)
{
    if (k < 32) {
        int64_t rpb = 128 / k; /* k==0 divides by zero in the reference too (:187) */
        grid[0] = (m + rpb - 1) / rpb; grid[1] = 1; block[0] = k; block[1] = rpb; *smem_unvalued = 0;
        return 0;
    }
    if (k < 64) {
        grid[0] = (m + 3) / 4; grid[1] = (k + 31) / 32; block[0] = 32; block[1] = 4; *smem_unvalued = 128 * 4;
        return 1;
    }
    grid[0] = (m + 7) / 8; grid[1] = (k + 63) / 64; block[0] = 32; block[1] = 8; *smem_unvalued = 8 * 32 * 4;
    return 2;
}

/* ------------------------------------------------------------------------------------
 * COO (sorted by row, col) -> CSR by counting sort, values forced to 1.
 * Follows spmm_test.cu:557-581 (A_data[ptr] = 1 at :574).
 * ---------------------------------------------------------------------------------- */
void oracle_coo_to_csr(int64_t nrows, int64_t nnz, const int32_t *row, const int32_t *col,
                       int32_t *indptr /* nrows+1 */, int32_t *indices /* nnz */, float *data /* nnz */)
{
    for (int64_t i = 0; i < nrows + 1; i++) indptr[i] = 0;
    for (int64_t n = 0; n < nnz; n++) indptr[row[n] + 1]++;
    for (int64_t n = 1; n < nrows + 1; n++) indptr[n] += indptr[n - 1];
    for (int64_t n = 0; n < nnz; n++) {
        int32_t ptr = indptr[row[n]];
        indices[ptr] = col[n];
        data[ptr] = 1.0f;
        indptr[row[n]] = ptr + 1;
    }
    for (int64_t n = nrows - 1; n > 0; n--) indptr[n] = indptr[n - 1];
    indptr[0] = 0;
}

/* ------------------------------------------------------------------------------------
 * Dense operand of the CLI: B[i] = float(rand()%100 - 50)/100 (spmm_test.cu:592-594).
 * The reference seeds with time(0) (:586-588); here the seed is explicit.  glibc rand().
 * ---------------------------------------------------------------------------------- */
void oracle_fill_B_cli(float *B, int64_t n, unsigned seed)
{
    srand(seed);
    for (int64_t i = 0; i < n; i++) B[i] = (float)(rand() % 100 - 50) / 100;
}

/* ------------------------------------------------------------------------------------
 * Matrix-Market coordinate reader with readMtx's post-conditions.
 *   banner            mm_read_banner        util/mmio.hpp:215-298
 *   size line         mm_read_mtx_crd_size  util/mmio.hpp:308-336
 *   tuples            readTuples            util/util.hpp:104-216 (integer -> %d, real -> %f, pattern -> 1)
 *   symmetric         makeSymmetric         util/util.hpp:218-284 (mirror off-diagonals, sort,
 *                                           drop self-loops and duplicate (row,col))
 *   final order       customSort            util/util.hpp:75-102 (by row, then col)
 * Values: the reference compacts row/col but not values in makeSymmetric (util.hpp:268-283)
 * and the CLI overwrites them with 1 anyway (spmm_test.cu:574); this restatement returns the
 * value belonging to each kept (row,col) for general matrices and the first occurrence's value
 * for symmetric ones.
 * Returns 0 on success; negative on error.  Two-call protocol: pass row==NULL to get sizes.
 * ---------------------------------------------------------------------------------- */
typedef struct { int32_t r, c; float v; int64_t idx; } coo_t;

static int coo_cmp(const void *a, const void *b)
{
    const coo_t *x = (const coo_t *)a, *y = (const coo_t *)b;
    if (x->r != y->r) return x->r < y->r ? -1 : 1;
    if (x->c != y->c) return x->c < y->c ? -1 : 1;
    /* std::sort is not stable; ties are identical (row,col) so order only matters for values */
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

static void lower(char *s) { for (; *s; s++) *s = (char)tolower((unsigned char)*s); }

int oracle_read_mtx(const char *fname, int32_t *nrows, int32_t *ncols, int64_t *nvals,
                    int32_t *row, int32_t *col, float *val, int64_t capacity)
{
    FILE *f = fopen(fname, "r");
    if (!f) return -1;
    char line[1025], banner[65], mtx[65], crd[65], dtype[65], scheme[65];
    if (!fgets(line, sizeof line, f)) { fclose(f); return -2; }
    if (sscanf(line, "%64s %64s %64s %64s %64s", banner, mtx, crd, dtype, scheme) != 5) { fclose(f); return -2; }
    lower(mtx); lower(crd); lower(dtype); lower(scheme);
    if (strncmp(banner, "%%MatrixMarket", 14) != 0) { fclose(f); return -3; }
    if (strcmp(mtx, "matrix") != 0 || strcmp(crd, "coordinate") != 0) { fclose(f); return -4; }
    int is_int = !strcmp(dtype, "integer"), is_real = !strcmp(dtype, "real"), is_pat = !strcmp(dtype, "pattern");
    int is_cplx = !strcmp(dtype, "complex");
    if (!is_int && !is_real && !is_pat && !is_cplx) { fclose(f); return -4; }
    int is_sym = !strcmp(scheme, "symmetric");
    if (!is_sym && strcmp(scheme, "general") && strcmp(scheme, "hermitian") && strcmp(scheme, "skew-symmetric")) {
        fclose(f); return -4;
    }
    do { if (!fgets(line, sizeof line, f)) { fclose(f); return -2; } } while (line[0] == '%');
    int M = 0, N = 0, nz = 0;
    while (sscanf(line, "%d %d %d", &M, &N, &nz) != 3) {
        if (!fgets(line, sizeof line, f)) { fclose(f); return -2; }
    }
    *nrows = M; *ncols = N;
    int64_t cap = is_sym ? 2 * (int64_t)nz : (int64_t)nz;
    coo_t *t = (coo_t *)malloc((size_t)(cap > 0 ? cap : 1) * sizeof(coo_t));
    if (!t) { fclose(f); return -5; }
    int64_t n = 0;
    if (is_int || is_real || is_pat) { /* complex: readMtx reads nothing (util.hpp:315-320) */
        for (int i = 0; i < nz; i++) {
            int r, c; float v = 1.0f;
            if (fscanf(f, "%d", &r) == EOF) break; /* "not enough rows" */
            if (fscanf(f, "%d", &c) != 1) c = 0;
            if (is_int) { int iv = 0; if (fscanf(f, "%d", &iv) != 1) iv = 0; v = (float)iv; }
            else if (is_real) { if (fscanf(f, "%f", &v) != 1) v = 0.0f; }
            t[n].r = r - 1; t[n].c = c - 1; t[n].v = v; t[n].idx = n; n++;
        }
    }
    fclose(f);
    if (is_sym) {
        int64_t n0 = n;
        for (int64_t i = 0; i < n0; i++)
            if (t[i].c != t[i].r) { t[n].r = t[i].c; t[n].c = t[i].r; t[n].v = t[i].v; t[n].idx = n; n++; }
        qsort(t, (size_t)n, sizeof(coo_t), coo_cmp);
        int64_t w = 0;
        for (int64_t i = 0; i < n; i++) {
            if (t[i].r == t[i].c) continue;                                     /* self-loop */
            if (w > 0 && t[i].r == t[w - 1].r && t[i].c == t[w - 1].c) continue; /* duplicate */
            /* a duplicate of a dropped self-loop is itself a self-loop: already skipped */
            t[w++] = t[i];
        }
        n = w;
    }
    qsort(t, (size_t)n, sizeof(coo_t), coo_cmp);
    *nvals = n;
    int rc = 0;
    if (row) {
        if (capacity < n) rc = -6;
        else for (int64_t i = 0; i < n; i++) { row[i] = t[i].r; col[i] = t[i].c; val[i] = t[i].v; }
    }
    free(t);
    return rc;
}

int oracle_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
