/* Exact K-nearest-neighbour oracle in plain C (TEST INFRASTRUCTURE - never linked into the product).
 *
 * Restates the algorithm of the reference's CPU KNN
 * (nerf_loc/models/ops/knn/src/knn_cpu.cpp:13-64, a vendored copy of pytorch3d's): for every query the
 * squared L2 distance to every support point is accumulated over d = 0..D-1 in that order, one rounding
 * per multiply and per add, and the K smallest are kept by (distance, index) - a candidate replaces the
 * current worst only if its distance is strictly smaller, and the worst among equal distances is the
 * larger index.  Output is ascending (knn_cpu.cpp:54-60 drains the max-heap back to front).
 *
 * Build: gcc -O2 -ffp-contract=off -pthread -shared -fPIC knn_oracle.c -o libknn_oracle.so
 * (-ffp-contract=off: no FMA, so distances are bit-identical to the reference build).
 * Queries are independent; `knn_oracle_mt` splits them over `threads` pthreads (no OpenMP in this image).
 */
#include <pthread.h>
#include <stdint.h>

static void knn_range(const float* p1, int64_t lo, int64_t hi, const float* p2, int64_t n2, int D, int K,
                      int64_t* idx, float* dist) {
  for (int64_t i = lo; i < hi; ++i) {
    float bd[64];
    int64_t bi[64];
    int cnt = 0;
    for (int64_t j = 0; j < n2; ++j) {
      float d = 0.f;
      for (int c = 0; c < D; ++c) {
        float diff = p1[i * D + c] - p2[j * D + c];
        d += diff * diff;
      }
      if (cnt < K) {            /* grow: insert keeping (dist, idx) ascending */
        int pos = cnt++;
        while (pos > 0 && bd[pos - 1] > d) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bd[pos] = d; bi[pos] = j;
      } else if (d < bd[K - 1]) { /* strictly better than the worst kept */
        int pos = K - 1;
        while (pos > 0 && bd[pos - 1] > d) { bd[pos] = bd[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
        bd[pos] = d; bi[pos] = j;
      }
    }
    for (int k = 0; k < K; ++k) {
      idx[i * K + k] = k < cnt ? bi[k] : 0;
      dist[i * K + k] = k < cnt ? bd[k] : 0.f;
    }
  }
}

void knn_oracle(const float* p1, int64_t n1, const float* p2, int64_t n2, int D, int K,
                int64_t* idx, float* dist) {
  knn_range(p1, 0, n1, p2, n2, D, K, idx, dist);
}

typedef struct { const float* p1; int64_t lo, hi; const float* p2; int64_t n2; int D, K; int64_t* idx; float* dist; } job_t;
static void* job_main(void* a) {
  job_t* j = (job_t*)a;
  knn_range(j->p1, j->lo, j->hi, j->p2, j->n2, j->D, j->K, j->idx, j->dist);
  return 0;
}

void knn_oracle_mt(const float* p1, int64_t n1, const float* p2, int64_t n2, int D, int K,
                   int64_t* idx, float* dist, int threads) {
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  pthread_t th[256];
  job_t jobs[256];
  int64_t per = (n1 + threads - 1) / threads;
  int started = 0;
  for (int t = 0; t < threads; ++t) {
    int64_t lo = t * per, hi = lo + per > n1 ? n1 : lo + per;
    if (lo >= hi) break;
    jobs[t] = (job_t){p1, lo, hi, p2, n2, D, K, idx, dist};
    pthread_create(&th[t], 0, job_main, &jobs[t]);
    ++started;
  }
  for (int t = 0; t < started; ++t) pthread_join(th[t], 0);
}
