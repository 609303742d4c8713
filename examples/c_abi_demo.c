/* Minimal host program against the C ABI alone (no Python, no torch): build a small bank, search it
 * from host buffers, print the hits. The same calls are what a cgo / JNI / ctypes binding makes.
 *
 *   gcc -std=c99 -I include examples/c_abi_demo.c -L retrieval_augmented_mds_b200 -lmips_b200 \
 *       -Wl,-rpath,$PWD/retrieval_augmented_mds_b200 -o /tmp/c_abi_demo && /tmp/c_abi_demo
 *
 * Replaces, end to end: faiss.index_factory + index.add (reference sotasum/mips.py:333-340) and
 * faiss_index.search (mips.py:383-386). Needs a B200; there is no CPU compute path. */
#include <stdio.h>
#include <stdlib.h>

#include "mips_b200.h"

int main(void) {
  enum { D = 768, N = 4096, NQ = 4, K = 5 };
  float* xb = (float*)malloc(sizeof(float) * N * D);
  float* xq = (float*)malloc(sizeof(float) * NQ * D);
  float dist[NQ * K];
  int64_t ids[NQ * K];
  unsigned s = 12345u;
  int i, j;
  mips_handle h = NULL;
  if (!xb || !xq) return 2;
  for (i = 0; i < N * D; ++i) {
    s = s * 1664525u + 1013904223u;
    xb[i] = (float)((s >> 8) & 0xffff) / 32768.0f - 1.0f;
  }
  for (j = 0; j < NQ; ++j)            /* query j = bank row 100 * j: it must come back first */
    for (i = 0; i < D; ++i) xq[j * D + i] = xb[(100 * j) * D + i];

  if (mips_create(&h, D, MIPS_METRIC_IP, MIPS_DTYPE_F32, /*device=*/0, /*capacity_rows=*/N) != 0 ||
      mips_add(h, xb, N, /*x_on_device=*/0, /*normalize=*/0, /*stream=*/NULL) != 0 ||
      mips_search_host(h, xq, NQ, K, /*q_normalize=*/0, /*ignore_ids=*/NULL, MIPS_OUT_IP, dist, ids, NULL) != 0) {
    fprintf(stderr, "mips_b200: %s\n", mips_last_error());
    return 1;
  }
  for (j = 0; j < NQ; ++j) {
    printf("query %d:", j);
    for (i = 0; i < K; ++i) printf("  %lld (%.3f)", (long long)ids[j * K + i], dist[j * K + i]);
    printf("\n");
    if (ids[j * K] != 100 * j) return 3;
  }
  printf("kernel: %s, rows: %lld\n", mips_last_algo(h), (long long)mips_ntotal(h));
  mips_destroy(h);
  free(xb);
  free(xq);
  return 0;
}
