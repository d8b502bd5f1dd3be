/* LRU model of the L2 seen by the gathers of A_hat.H (scripts/l2_model.py): every access touches one row of H
 * (row_bytes each); returns the number of misses for a cache of cap_rows rows.  gcc -O2 -shared -fPIC. */
#include <stdint.h>
#include <stdlib.h>

int64_t lru_misses(const int32_t* acc, int64_t n_acc, int32_t n_rows, int32_t cap_rows) {
  int32_t* prev = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_rows);
  int32_t* next = (int32_t*)malloc(sizeof(int32_t) * (size_t)n_rows);
  uint8_t* in = (uint8_t*)calloc((size_t)n_rows, 1);
  int32_t head = -1, tail = -1, size = 0;   /* head = most recent */
  int64_t miss = 0;
  for (int64_t i = 0; i < n_acc; ++i) {
    const int32_t r = acc[i];
    if (in[r]) {
      if (r == head) continue;
      /* unlink */
      const int32_t p = prev[r], q = next[r];
      if (p >= 0) next[p] = q;
      if (q >= 0) prev[q] = p; else tail = p;
    } else {
      ++miss;
      if (size == cap_rows) {   /* evict the least recent */
        const int32_t t = tail;
        tail = prev[t];
        if (tail >= 0) next[tail] = -1; else head = -1;
        in[t] = 0;
        --size;
      }
      in[r] = 1;
      ++size;
    }
    prev[r] = -1;
    next[r] = head;
    if (head >= 0) prev[head] = r;
    head = r;
    if (tail < 0) tail = r;
  }
  free(prev); free(next); free(in);
  return miss;
}
