/*
 * CPU ORACLE (test infrastructure only -- NOT part of the product).
 *
 * Plain-C restatement of the reference's alignment simulator walk:
 *   elynx-markov/src/ELynx/Simulate/MarkovProcessAlongTree.hs
 *     getRootStates :57-63, simulateAndFlatten' :88-98, simulateAndFlattenPar :101-125,
 *     getComponentsAndRootStates :163-173, simulateAndFlattenMixtureModel' :195-208,
 *     getChunks :210-215, parComp :221-235, simulateAndFlattenMixtureModelPar :238-262
 *   elynx-markov/src/ELynx/Simulate/MarkovProcess.hs   jump :48-49
 * plus the two third-party pieces that are not vendored under /root/reference, restated from
 * their published definitions:
 *   mwc-random `categorical`: cv = scanl1 (+) w; p = last cv * u; first i with cv[i] >= p
 *   random `StdGen` (SplitMix) with uniformDoublePositive01M.
 * Pinned by the reference's only known-answer test for this path (KAT-1,
 * test/ELynx/Simulate/MarkovProcessAlongTreeSpec.hs:45-48: mkStdGen 0 -> state 2); everything else
 * about the stream is PARITY UNPINNED and machine dependent in the reference itself (chunking by
 * getNumCapabilities), which is why the bit-exact contract is the injected-uniforms mode.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.
 *
 * Tree encoding: preorder arrays (root = node 0, children left to right), parent[n] < n,
 * parent[0] = -1, leaf_row[n] = DFS-left-to-right leaf index or -1.
 * Tables: p[N][C][K][K] probability matrices (row = from-state), as probMatrix returns them.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORACLE_MAX_K 64

/* mwc-random categorical on a K-vector with the uniform u in (0,1]. Returns -1 for "bad weights". */
static inline int categorical(const double *w, int k, double u) {
  double cv[ORACLE_MAX_K];
  double acc = w[0];
  cv[0] = acc;
  for (int i = 1; i < k; ++i) {
    acc = acc + w[i]; /* scanl1' (+): sequential, left to right */
    cv[i] = acc;
  }
  volatile double p = cv[k - 1] * u;
  for (int i = 0; i < k; ++i)
    if (cv[i] >= p) return i;
  return -1;
}

/* Same rule on a precomputed cumulative row (used by the timed CPU baseline only). */
static inline int categorical_cum(const double *cv, int k, double u) {
  double p = cv[k - 1] * u;
  for (int i = 0; i < k; ++i)
    if (cv[i] >= p) return i;
  return -1;
}

/*
 * Injected-uniforms walk.  U has (is_mixture ? 2 : 1) + N rows of S doubles:
 * [component draws (mixture only)] [root-state draws] [node 0 draws] ... [node N-1 draws]
 * (SURVEY.md Appendix A.1).  Returns 0, or -(1+node) when categorical fails at a node
 * (-(N+1) for the component draw, -(N+2) for the root draw).
 */
int oracle_simulate_injected(int N, int K, int C, int is_mixture, const int32_t *parent, const int32_t *leaf_row,
                             const double *ws, const double *ds, const double *p, int64_t S, const double *U,
                             uint8_t *leaves, int32_t *comps_out, uint8_t *internal_out) {
  if (K > ORACLE_MAX_K || K < 1 || C < 1) return -1000000;
  const int64_t BLK = 2048;
  uint8_t *st = (uint8_t *)malloc((size_t)N * BLK);
  int32_t *comp = (int32_t *)malloc(sizeof(int32_t) * BLK);
  uint8_t *root = (uint8_t *)malloc(BLK);
  if (!st || !comp || !root) return -1000001;
  int rc = 0;
  const int64_t KK = (int64_t)K * K;
  for (int64_t s0 = 0; s0 < S && rc == 0; s0 += BLK) {
    int64_t nb = S - s0 < BLK ? S - s0 : BLK;
    const double *u = U + s0;
    int64_t row = 0;
    if (is_mixture) {
      for (int64_t i = 0; i < nb; ++i) {
        int c = categorical(ws, C, u[i]);
        if (c < 0) { rc = -(N + 1); break; }
        comp[i] = c;
      }
      row = 1;
    } else {
      for (int64_t i = 0; i < nb; ++i) comp[i] = 0;
    }
    if (rc) break;
    for (int64_t i = 0; i < nb; ++i) {
      int r = categorical(ds + (int64_t)comp[i] * K, K, u[row * S + i]);
      if (r < 0) { rc = -(N + 2); break; }
      root[i] = (uint8_t)r;
    }
    if (rc) break;
    row += 1;
    for (int n = 0; n < N && rc == 0; ++n) {
      const uint8_t *par = parent[n] < 0 ? root : st + (int64_t)parent[n] * BLK;
      uint8_t *me = st + (int64_t)n * BLK;
      const double *pn = p + (int64_t)n * C * KK;
      const double *un = u + (row + n) * S;
      for (int64_t i = 0; i < nb; ++i) {
        int j = categorical(pn + (int64_t)comp[i] * KK + (int64_t)par[i] * K, K, un[i]);
        if (j < 0) { rc = -(1 + n); break; }
        me[i] = (uint8_t)j;
      }
      if (leaf_row[n] >= 0) memcpy(leaves + (int64_t)leaf_row[n] * S + s0, me, (size_t)nb);
      if (internal_out) memcpy(internal_out + (int64_t)n * S + s0, me, (size_t)nb);
    }
    if (comps_out)
      for (int64_t i = 0; i < nb; ++i) comps_out[s0 + i] = comp[i];
  }
  free(st);
  free(comp);
  free(root);
  return rc;
}

/* ------------------------------------------------------------------------------------------- */
/* SplitMix (random's StdGen)                                                                  */
/* ------------------------------------------------------------------------------------------- */
typedef struct { uint64_t seed, gamma; } smgen;

static inline uint64_t mix64(uint64_t z) {
  z ^= z >> 33; z *= 0xff51afd7ed558ccdULL; z ^= z >> 33; z *= 0xc4ceb9fe1a85ec53ULL; z ^= z >> 33;
  return z;
}
static inline uint64_t mix64v13(uint64_t z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ULL; z ^= z >> 27; z *= 0x94d049bb133111ebULL; z ^= z >> 31;
  return z;
}
static inline uint64_t mix_gamma(uint64_t z) {
  z = mix64v13(z) | 1ULL;
  return __builtin_popcountll(z ^ (z >> 1)) >= 24 ? z : z ^ 0xaaaaaaaaaaaaaaaaULL;
}
static smgen mk_std_gen(uint64_t s) {
  smgen g = {mix64(s), mix_gamma(s + 0x9e3779b97f4a7c15ULL)};
  return g;
}
static inline uint64_t next_word64(smgen *g) {
  g->seed += g->gamma;
  return mix64(g->seed);
}
/* split g = (g1, g2): g1 is handed out, g2 is kept */
static smgen split_gen(smgen *g) {
  uint64_t s1 = g->seed + g->gamma, s2 = s1 + g->gamma;
  smgen first = {s2, g->gamma};
  smgen second = {mix64(s1), mix_gamma(s2)};
  *g = second;
  return first;
}
static inline double uniform_double_positive01(smgen *g) {
  return (double)next_word64(g) / 18446744073709551616.0 + 0x1p-65;
}

uint64_t oracle_splitmix_first_word(uint64_t seed) {
  smgen g = mk_std_gen(seed);
  return next_word64(&g);
}

typedef struct {
  int N, K, C, is_mixture, precomputed_cum;
  const int32_t *parent, *leaf_row;
  const double *ws, *ds, *p;
  int64_t S, s0, n;
  smgen gen;
  uint8_t *leaves;
  int32_t *comps;
  uint8_t *internal;
  int rc;
} chunk_job;

static void *run_chunk(void *arg) {
  chunk_job *j = (chunk_job *)arg;
  const int N = j->N, K = j->K, C = j->C;
  const int64_t KK = (int64_t)K * K, n = j->n, S = j->S;
  j->rc = 0;
  if (n == 0) return 0;
  /* live-state buffers: one per tree depth level would do; keep it simple: N rows of n bytes,
     allocated lazily and freed when the last child has consumed them. */
  uint8_t **st = (uint8_t **)calloc((size_t)N, sizeof(uint8_t *));
  int *pending = (int *)calloc((size_t)N, sizeof(int));
  int32_t *comp = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
  uint8_t *root = (uint8_t *)malloc((size_t)n);
  for (int i = 1; i < N; ++i) pending[j->parent[i]]++;
  double cw[ORACLE_MAX_K * 64];
  smgen g = j->gen;
  /* all component draws, then all root draws (MarkovProcessAlongTree.hs:170-172) */
  if (j->is_mixture) {
    double *cvw = (double *)malloc(sizeof(double) * (size_t)C);
    double acc = 0;
    for (int c = 0; c < C; ++c) { acc = c ? acc + j->ws[c] : j->ws[0]; cvw[c] = acc; }
    for (int64_t i = 0; i < n; ++i) {
      double u = uniform_double_positive01(&g), pp = cvw[C - 1] * u;
      int c = 0;
      while (c < C && !(cvw[c] >= pp)) ++c;
      if (c == C) { j->rc = -(N + 1); c = C - 1; }
      comp[i] = c;
    }
    free(cvw);
  } else {
    for (int64_t i = 0; i < n; ++i) comp[i] = 0;
  }
  (void)cw;
  for (int64_t i = 0; i < n; ++i) {
    int r = categorical(j->ds + (int64_t)comp[i] * K, K, uniform_double_positive01(&g));
    if (r < 0) { j->rc = -(N + 2); r = 0; }
    root[i] = (uint8_t)r;
  }
  for (int nd = 0; nd < N; ++nd) {
    const uint8_t *par = j->parent[nd] < 0 ? root : st[j->parent[nd]];
    uint8_t *me = (uint8_t *)malloc((size_t)n);
    st[nd] = me;
    const double *pn = j->p + (int64_t)nd * C * KK;
    if (j->precomputed_cum) {
      for (int64_t i = 0; i < n; ++i) {
        int c = categorical_cum(pn + (int64_t)comp[i] * KK + (int64_t)par[i] * K, K, uniform_double_positive01(&g));
        if (c < 0) { j->rc = -(1 + nd); c = 0; }
        me[i] = (uint8_t)c;
      }
    } else {
      for (int64_t i = 0; i < n; ++i) {
        int c = categorical(pn + (int64_t)comp[i] * KK + (int64_t)par[i] * K, K, uniform_double_positive01(&g));
        if (c < 0) { j->rc = -(1 + nd); c = 0; }
        me[i] = (uint8_t)c;
      }
    }
    if (j->leaf_row[nd] >= 0) memcpy(j->leaves + (int64_t)j->leaf_row[nd] * S + j->s0, me, (size_t)n);
    if (j->internal) memcpy(j->internal + (int64_t)nd * S + j->s0, me, (size_t)n);
    if (pending[nd] == 0) { free(me); st[nd] = 0; }
    int pa = j->parent[nd];
    if (pa >= 0 && --pending[pa] == 0) { free(st[pa]); st[pa] = 0; }
  }
  if (j->comps)
    for (int64_t i = 0; i < n; ++i) j->comps[j->s0 + i] = comp[i];
  for (int i = 0; i < N; ++i) free(st[i]);
  free(st); free(pending); free(comp); free(root);
  return 0;
}

/*
 * Free-running walk on the reference's own random stream.
 *   nchunks == 0 : the serial API (simulateAndFlatten / simulateAndFlattenMixtureModel, :71-83,:178-193):
 *                  one chunk, the generator is used directly.
 *   nchunks >= 1 : the *Par API (:101-125, :221-262): the generator is split nchunks times, chunk i of
 *                  size getChunks(nchunks, S)[i] runs on split i (the reference uses
 *                  nchunks = getNumCapabilities), chunks run on up to `nthreads` OS threads.
 * `p` holds probability matrices, or row-cumulative tables when precomputed_cum != 0 (timed baseline:
 * hoists the reference's per-call scanl1 out of the loop; identical results when cum = cumsum(p)).
 */
int oracle_simulate_splitmix(int N, int K, int C, int is_mixture, const int32_t *parent, const int32_t *leaf_row,
                             const double *ws, const double *ds, const double *p, int precomputed_cum, int64_t S,
                             uint64_t seed, int nchunks, int nthreads, uint8_t *leaves, int32_t *comps_out,
                             uint8_t *internal_out) {
  if (K > ORACLE_MAX_K || K < 1 || C < 1) return -1000000;
  smgen g = mk_std_gen(seed);
  int nc = nchunks > 0 ? nchunks : 1;
  chunk_job *jobs = (chunk_job *)calloc((size_t)nc, sizeof(chunk_job));
  int64_t q = S / nc, r = S % nc, s0 = 0;
  for (int i = 0; i < nc; ++i) {
    chunk_job *j = &jobs[i];
    j->N = N; j->K = K; j->C = C; j->is_mixture = is_mixture; j->precomputed_cum = precomputed_cum;
    j->parent = parent; j->leaf_row = leaf_row; j->ws = ws; j->ds = ds; j->p = p;
    j->S = S; j->s0 = s0; j->n = q + (i < r ? 1 : 0);
    j->gen = nchunks > 0 ? split_gen(&g) : g;
    j->leaves = leaves; j->comps = comps_out; j->internal = internal_out;
    s0 += j->n;
  }
  int rc = 0;
  if (nthreads <= 1 || nc == 1) {
    for (int i = 0; i < nc; ++i) { run_chunk(&jobs[i]); if (jobs[i].rc) rc = jobs[i].rc; }
  } else {
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int base = 0; base < nc; base += nthreads) {
      int m = nc - base < nthreads ? nc - base : nthreads;
      for (int i = 0; i < m; ++i) pthread_create(&th[i], 0, run_chunk, &jobs[base + i]);
      for (int i = 0; i < m; ++i) { pthread_join(th[i], 0); if (jobs[base + i].rc) rc = jobs[base + i].rc; }
    }
    free(th);
  }
  free(jobs);
  return rc;
}
