// LRU miss count of the B-row access stream of a CSR walked in row order, for several capacities (in rows).
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
int main(int argc, char **argv) {
    FILE *f = fopen(argv[1], "rb");
    int64_t n; if (fread(&n, 8, 1, f) != 1) return 1;
    int32_t N; if (fread(&N, 4, 1, f) != 1) return 1;
    int32_t *col = malloc(n * 4); if ((int64_t)fread(col, 4, n, f) != n) return 1; fclose(f);
    for (int a = 2; a < argc; a++) {
        long cap = atol(argv[a]);
        // LRU via timestamp array + lazy ring: exact LRU using a doubly linked list over row ids
        int32_t *prev = malloc((size_t)N * 4), *next = malloc((size_t)N * 4); char *in = calloc(N, 1);
        int32_t head = -1, tail = -1; long size = 0, miss = 0;
        for (int64_t i = 0; i < n; i++) {
            int32_t c = col[i];
            if (in[c]) { // move to head
                if (head != c) {
                    int32_t p = prev[c], q = next[c];
                    next[p] = q; if (q >= 0) prev[q] = p; else tail = p;
                    prev[c] = -1; next[c] = head; prev[head] = c; head = c;
                }
            } else {
                miss++;
                if (size == cap) { int32_t t = tail; tail = prev[t]; if (tail >= 0) next[tail] = -1; else head = -1; in[t] = 0; size--; }
                in[c] = 1; prev[c] = -1; next[c] = head; if (head >= 0) prev[head] = c; else tail = c; head = c; size++;
            }
        }
        printf("capacity %ld rows (%.1f MB at 512 B): misses %ld of %ld gathers (%.1f%%) -> %.3f GB of B reads\n", cap, cap * 512 / 1e6, miss, (long)n, 100.0 * miss / n, miss * 512 / 1e9);
        free(prev); free(next); free(in);
    }
    return 0;
}
