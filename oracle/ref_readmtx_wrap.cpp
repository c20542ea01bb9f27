// oracle/_ref/libref_readmtx.so -- the reference's own Matrix-Market reader
// (/root/reference/util/util.hpp readMtx<float>, util/mmio.hpp), compiled from where it
// lies and exposed through one C entry point so the oracle restatement
// (oracle_read_mtx) can be checked against it.  TEST INFRASTRUCTURE ONLY.
#include "util/util.hpp"

extern "C" int ref_read_mtx(const char *fname, int *nrows, int *ncols, long long *nvals,
                            int *row, int *col, float *val, long long capacity)
{
    std::vector<int> r, c;
    std::vector<float> v;
    int nr = 0, nc = 0, nv = 0;
    readMtx<float>(fname, r, c, v, nr, nc, nv);   // exits the process on a missing file (util.hpp:300-303)
    *nrows = nr; *ncols = nc; *nvals = nv;
    if (row) {
        if (capacity < nv) return -6;
        for (int i = 0; i < nv; i++) { row[i] = r[i]; col[i] = c[i]; val[i] = v[i]; }
    }
    return 0;
}
