/*
 * TEST INFRASTRUCTURE ONLY -- executable specification of the PRODUCT's free-running draw
 * (DESIGN.md "free-running draw"), used to check the CUDA kernels bit for bit, including the rare
 * refinement path that no statistical test can see.  It restates elynx_b200's own spec, not the
 * reference: the reference-side semantics (inverse-CDF categorical on an injected uniform matrix) live
 * in oracle/sim.c.  The walk itself is the reference's (MarkovProcessAlongTree.hs:88-98,195-208):
 * every node draws from the row of its parent's state, root first.
 *
 * Spec:  key = (seed lo, seed hi); Philox4x32-R, R in {7,10}
 *   component (mixtures): x = philox(site lo, site hi, 0, 2); u = ((x1:x0 >> 11) + 1) 2^-53;
 *                         c = first i with cumw[i] >= cumw[C-1]*u      (mwc-random categorical)
 *   root state:           x = philox(site lo, site hi, 0, 3); same rule on cumpi[c]
 *   node n, site s:       x = philox(s>>3 lo, s>>3 hi, n, 0); x16 = 16-bit field (s & 7), low half of word (s&7)>>1 first
 *                         j = x16 & (KP-1); r = x16 >> A; e = e16[n][c][parent][j]
 *                         r < e>>A -> j ;  r > e>>A -> alias = e & (KP-1) ;
 *                         tie -> y = philox(s lo, s hi, n, 1).x0 ;  y < qlo[n][c][parent][j] ? j : alias
 *
 * Pair mode (models with K <= 4; elynx_b200/csrc/elx_pair.cu): the children of a node are drawn two at a time from the
 * joint row of record g = (nodeA, nodeB | parent), 16 outcomes o = state(nodeA) | state(nodeB) << 2:
 *   x16 as above with n = nodeA;  j = x16 & 15; r = x16 >> 4; e = pe16[g][c][parent state][j]
 *   r < e>>4 -> o = j ;  r > e>>4 -> o = e & 15 ;  tie -> philox(s lo, s hi, nodeA, 1).x0 < pqlo[g][c][parent state][j] ? j : e & 15
 * Records come in walk order (a record's parent is drawn by an earlier record; node_b = -1 for a group of one).
 *
 * Quad mode (models with K <= 4, leaves only; elynx_b200/csrc/elx_quad.cu): record g = (out[0..3] | r) draws the states of up to
 * four frontier nodes below r at once from a 256-outcome joint row (inner nodes summed out), o = sum_k state(out[k]) << 2k.
 * A draw is 24 bits; the 16 sites of a group s >> 4 share three Philox calls:
 *   B[0..47] = little-endian bytes of philox(s>>4 lo, s>>4 hi, out[0], 4 * call), call = 0, 1, 2;  d = s & 15
 *   X[0..11] = the 12 little-endian words of B; x24 = X[3(d>>2) + (d&3)] >> 8 for d&3 < 3, else the low bytes of X[3(d>>2)], X[3(d>>2)+1],
 *   X[3(d>>2)+2] (first = least significant);  j = x24 & 255; t = x24 >> 8
 *   t != 0:  e = qe32[g][c][state of r][j] = [q : 16][alias : 8][0 : 8];  o = x24 < (e >> 8) ? j : alias     (two-level alias rows:
 *   t == 0:  u = philox(s lo, s hi, out[0], 1).x0;  er = qres32[g][c][state of r][u & 255] = [thr : 24][alias : 8];     no ties)
 *            o = (u >> 8) < thr ? u & 255 : alias
 * Records come in walk order (r is the root or a frontier node of an earlier record); recs = 52 int32 per record in the
 * layout include/elynx_b200.h documents.
 *
 * Duo mode (models with 5..22 states, leaves only; elynx_b200/csrc/elx_duo.cu): record g = (out[0..1] | r) draws up to two frontier
 * nodes at once from a K*K-outcome joint row in 512 alias columns, o = state(out[0]) + K * state(out[1]).  Sites are bucketed by
 * component inside superblocks of 2^SB sites (SB = 20 for C <= 16, 22 for C <= 64, else 24): B = s >> SB, rank = number of earlier sites of B with the same component,
 * q = rank >> 4, d = rank & 15; the 16 sites of a group share three Philox calls:
 *   Z[0..47] = little-endian bytes of philox(q | c << 20, B lo, out[0], (B hi << 8) | 4 * call), call = 0, 1, 2
 *   x24 = draw d of Z as in quad mode;  j = x24 & 511; t = x24 >> 9
 *   t != 0:  e = de32[g][c][state of r][j] = [q : 15][alias : 9][0 : 8];  o = x24 < (e >> 8) ? j : alias
 *   t == 0:  u = philox(s lo, s hi, out[0], 1).x0;  er = dres32[g][c][state of r][u & 511] = [thr : 23][alias : 9];
 *            o = (u >> 9) < thr ? u & 511 : alias
 */
#include <stdint.h>
#include <stdlib.h>

static void philox(int rounds, uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < rounds; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

void spec_philox4x32(int rounds, const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
  philox(rounds, c, key[0], key[1]);
  for (int i = 0; i < 4; ++i) out[i] = c[i];
}

static int cat_cum(const double *cv, int k, double u) {
  volatile double p = cv[k - 1] * u;
  for (int i = 0; i < k; ++i)
    if (cv[i] >= p) return i;
  return -1;
}

int spec_freerun(int N, int K, int C, int is_mixture, int A, int rounds, const int32_t *parent, const int32_t *leaf_row,
                 const double *cumw, const double *cumpi, const uint16_t *e16, const uint32_t *qlo, uint64_t seed,
                 int64_t site0, int64_t nsites, uint8_t *leaves, int64_t leaves_pitch, int32_t *comps, uint8_t *internal,
                 int64_t internal_pitch) {
  const int KP = 1 << A;
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint8_t *st = (uint8_t *)malloc((size_t)N);
  for (int64_t col = 0; col < nsites; ++col) {
    uint64_t s = (uint64_t)(site0 + col);
    int c = 0;
    if (is_mixture) {
      uint32_t x[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0, 2};
      philox(rounds, x, k0, k1);
      uint64_t w = ((uint64_t)x[1] << 32) | x[0];
      c = cat_cum(cumw, C, (double)((w >> 11) + 1) * 0x1p-53);
      if (c < 0) { free(st); return -(N + 1); }
    }
    if (comps) comps[col] = c;
    uint32_t x[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0, 3};
    philox(rounds, x, k0, k1);
    uint64_t w = ((uint64_t)x[1] << 32) | x[0];
    int root = cat_cum(cumpi + (int64_t)c * K, K, (double)((w >> 11) + 1) * 0x1p-53);
    if (root < 0) { free(st); return -(N + 2); }
    for (int n = 0; n < N; ++n) {
      int par = parent[n] < 0 ? root : st[parent[n]];
      uint64_t g = s >> 3;
      uint32_t y[4] = {(uint32_t)g, (uint32_t)(g >> 32), (uint32_t)n, 0};
      philox(rounds, y, k0, k1);
      int f = (int)(s & 7);
      uint32_t x16 = (y[f >> 1] >> ((f & 1) * 16)) & 0xffffu;
      int j = x16 & (KP - 1);
      uint32_t r = x16 >> A;
      int64_t idx = (((int64_t)n * C + c) * K + par) * KP + j;
      uint32_t e = e16[idx], qh = e >> A;
      int al = e & (KP - 1), child;
      if (r < qh) child = j;
      else if (r > qh) child = al;
      else {
        uint32_t z[4] = {(uint32_t)s, (uint32_t)(s >> 32), (uint32_t)n, 1};
        philox(rounds, z, k0, k1);
        child = z[0] < qlo[idx] ? j : al;
      }
      st[n] = (uint8_t)child;
      if (leaf_row[n] >= 0) leaves[(int64_t)leaf_row[n] * leaves_pitch + col] = (uint8_t)child;
      if (internal) internal[(int64_t)n * internal_pitch + col] = (uint8_t)child;
    }
  }
  free(st);
  return 0;
}

int spec_freerun_pair(int N, int K, int C, int is_mixture, int rounds, const int32_t *parent, const int32_t *leaf_row,
                      const double *cumw, const double *cumpi, int G, const int32_t *node_a, const int32_t *node_b,
                      const uint16_t *pe16, const uint32_t *pqlo, uint64_t seed, int64_t site0, int64_t nsites, uint8_t *leaves,
                      int64_t leaves_pitch, int32_t *comps, uint8_t *internal, int64_t internal_pitch) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint8_t *st = (uint8_t *)malloc((size_t)N);
  uint8_t *done = (uint8_t *)malloc((size_t)N);
  for (int64_t col = 0; col < nsites; ++col) {
    uint64_t s = (uint64_t)(site0 + col);
    int c = 0;
    if (is_mixture) {
      uint32_t x[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0, 2};
      philox(rounds, x, k0, k1);
      uint64_t w = ((uint64_t)x[1] << 32) | x[0];
      c = cat_cum(cumw, C, (double)((w >> 11) + 1) * 0x1p-53);
      if (c < 0) { free(st); free(done); return -(N + 1); }
    }
    if (comps) comps[col] = c;
    uint32_t x[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0, 3};
    philox(rounds, x, k0, k1);
    uint64_t w = ((uint64_t)x[1] << 32) | x[0];
    int root = cat_cum(cumpi + (int64_t)c * K, K, (double)((w >> 11) + 1) * 0x1p-53);
    if (root < 0) { free(st); free(done); return -(N + 2); }
    for (int n = 0; n < N; ++n) done[n] = 0;
    for (int g = 0; g < G; ++g) {
      const int nA = node_a[g], nB = node_b[g];
      if (nA < 0 || nA >= N || nB >= N || done[nA] || (nB >= 0 && (done[nB] || parent[nB] != parent[nA]))) { free(st); free(done); return -1; }
      if (parent[nA] >= 0 && !done[parent[nA]]) { free(st); free(done); return -2; }  /* not in walk order */
      int par = parent[nA] < 0 ? root : st[parent[nA]];
      uint64_t gg = s >> 3;
      uint32_t y[4] = {(uint32_t)gg, (uint32_t)(gg >> 32), (uint32_t)nA, 0};
      philox(rounds, y, k0, k1);
      int f = (int)(s & 7);
      uint32_t x16 = (y[f >> 1] >> ((f & 1) * 16)) & 0xffffu;
      uint32_t j = x16 & 15u, r = x16 >> 4;
      int64_t idx = ((((int64_t)g * C + c) * K) + par) * 16 + j;
      uint32_t e = pe16[idx], qh = e >> 4, al = e & 15u, o;
      if (r < qh) o = j;
      else if (r > qh) o = al;
      else {
        uint32_t z[4] = {(uint32_t)s, (uint32_t)(s >> 32), (uint32_t)nA, 1};
        philox(rounds, z, k0, k1);
        o = z[0] < pqlo[idx] ? j : al;
      }
      st[nA] = (uint8_t)(o & 3u); done[nA] = 1;
      if (nB >= 0) { st[nB] = (uint8_t)(o >> 2); done[nB] = 1; }
    }
    for (int n = 0; n < N; ++n) {
      if (!done[n]) { free(st); free(done); return -3; }
      if (leaf_row[n] >= 0) leaves[(int64_t)leaf_row[n] * leaves_pitch + col] = st[n];
      if (internal) internal[(int64_t)n * internal_pitch + col] = st[n];
    }
  }
  free(st); free(done);
  return 0;
}

/* draw d (0..15) of the 48 random bytes = 12 little-endian words X[0..11]: draws 4k, 4k+1, 4k+2 are the upper 24 bits of
 * X[3k], X[3k+1], X[3k+2]; draw 4k+3 collects the three low bytes those leave over (X[3k] first) */
static uint32_t draw24(const uint8_t *B, int d) {
  const int w0 = 3 * (d >> 2), i = d & 3;
  if (i < 3) return (uint32_t)B[4 * (w0 + i) + 1] | ((uint32_t)B[4 * (w0 + i) + 2] << 8) | ((uint32_t)B[4 * (w0 + i) + 3] << 16);
  return (uint32_t)B[4 * w0] | ((uint32_t)B[4 * (w0 + 1)] << 8) | ((uint32_t)B[4 * (w0 + 2)] << 16);
}

#define QUAD_REC_INTS 52
static int64_t quad_ties = 0;
int64_t spec_quad_ties(void) { return quad_ties; } /* escape-level draws (t == 0) seen by the last spec_freerun_quad call */
int spec_freerun_quad(int N, int K, int C, int is_mixture, int rounds, const int32_t *leaf_row, const double *cumw,
                      const double *cumpi, int G, const int32_t *recs, const uint32_t *qe32, const uint32_t *qres32, uint64_t seed,
                      int64_t site0, int64_t nsites, uint8_t *leaves, int64_t leaves_pitch, int32_t *comps) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint8_t *st = (uint8_t *)malloc((size_t)N);
  uint8_t *done = (uint8_t *)malloc((size_t)N);
  int rc = 0;
  quad_ties = 0;
  for (int64_t col = 0; col < nsites && !rc; ++col) {
    uint64_t s = (uint64_t)(site0 + col);
    int c = 0;
    if (is_mixture) {
      uint32_t x[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0, 2};
      philox(rounds, x, k0, k1);
      uint64_t w = ((uint64_t)x[1] << 32) | x[0];
      c = cat_cum(cumw, C, (double)((w >> 11) + 1) * 0x1p-53);
      if (c < 0) { rc = -(N + 1); break; }
    }
    if (comps) comps[col] = c;
    uint32_t x[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0, 3};
    philox(rounds, x, k0, k1);
    uint64_t w = ((uint64_t)x[1] << 32) | x[0];
    int root = cat_cum(cumpi + (int64_t)c * K, K, (double)((w >> 11) + 1) * 0x1p-53);
    if (root < 0) { rc = -(N + 2); break; }
    for (int n = 0; n < N; ++n) done[n] = 0;
    st[0] = (uint8_t)root; done[0] = 1;
    for (int g = 0; g < G; ++g) {
      const int32_t *r = recs + (size_t)g * QUAD_REC_INTS;
      const int rt = r[0], n0 = r[1];
      if (rt < 0 || rt >= N || !done[rt] || n0 < 0 || n0 >= N) { rc = -2; break; }  /* not in walk order */
      uint64_t gg = s >> 4;
      uint8_t B[48];
      for (int call = 0; call < 3; ++call) {
        uint32_t y[4] = {(uint32_t)gg, (uint32_t)(gg >> 32), (uint32_t)n0, 4u * (uint32_t)call};
        philox(rounds, y, k0, k1);
        for (int b = 0; b < 16; ++b) B[16 * call + b] = (uint8_t)(y[b >> 2] >> (8 * (b & 3)));
      }
      const int d = (int)(s & 15);
      uint32_t x24 = draw24(B, d);
      uint32_t j = x24 & 255u, t = x24 >> 8;
      int64_t idx = (((int64_t)g * C + c) * 4 + st[rt]) * 256 + j;
      uint32_t o;
      if (t != 0) {
        uint32_t e = qe32[idx];
        o = x24 < (e >> 8) ? j : ((e >> 8) & 255u);
      } else { /* escape level: the row's residual masses, 32 fresh bits */
        uint32_t z[4] = {(uint32_t)s, (uint32_t)(s >> 32), (uint32_t)n0, 1};
        philox(rounds, z, k0, k1);
        uint32_t jr = z[0] & 255u, er = qres32[idx - j + jr];
        o = (z[0] >> 8) < (er >> 8) ? jr : (er & 255u);
        ++quad_ties;
      }
      for (int k = 0; k < 4; ++k) {
        const int n = r[1 + k];
        if (n < 0) continue;
        if (n >= N || done[n]) { rc = -1; break; }
        st[n] = (uint8_t)((o >> (2 * k)) & 3u); done[n] = 1;
      }
      if (rc) break;
    }
    if (rc) break;
    for (int n = 0; n < N; ++n)
      if (leaf_row[n] >= 0) {
        if (!done[n]) { rc = -3; break; }
        leaves[(int64_t)leaf_row[n] * leaves_pitch + col] = st[n];
      }
  }
  free(st); free(done);
  return rc;
}

static int duo_sb(int C) { return C <= 16 ? 20 : C <= 64 ? 22 : 24; }
static int64_t duo_ties = 0;
int64_t spec_duo_ties(void) { return duo_ties; } /* escape-level draws (t == 0) seen by the last spec_freerun_duo call */
int spec_freerun_duo(int N, int K, int C, int is_mixture, int rounds, const int32_t *leaf_row, const double *cumw,
                     const double *cumpi, int G, const int32_t *recs, const uint32_t *de32, const uint32_t *dres32, uint64_t seed,
                     int64_t site0, int64_t nsites, uint8_t *leaves, int64_t leaves_pitch, int32_t *comps) {
  const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint8_t *st = (uint8_t *)malloc((size_t)N);
  uint8_t *done = (uint8_t *)malloc((size_t)N);
  uint32_t *count = (uint32_t *)calloc((size_t)C, sizeof(uint32_t));
  int rc = 0;
  duo_ties = 0;
  const int DUO_SB = duo_sb(C);
  const uint64_t first = ((uint64_t)site0 >> DUO_SB) << DUO_SB;
  uint64_t cur_sb = (uint64_t)-1;
  for (uint64_t s = first; s < (uint64_t)(site0 + nsites) && !rc; ++s) {
    const uint64_t sb = s >> DUO_SB;
    if (sb != cur_sb) { for (int c = 0; c < C; ++c) count[c] = 0; cur_sb = sb; }
    int c = 0;
    if (is_mixture) {
      uint32_t x[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0, 2};
      philox(rounds, x, k0, k1);
      uint64_t w = ((uint64_t)x[1] << 32) | x[0];
      c = cat_cum(cumw, C, (double)((w >> 11) + 1) * 0x1p-53);
      if (c < 0) { rc = -(N + 1); break; }
    }
    const uint32_t rank = count[c]++;
    if (s < (uint64_t)site0) continue;
    const int64_t col = (int64_t)(s - (uint64_t)site0);
    if (comps) comps[col] = c;
    uint32_t x[4] = {(uint32_t)s, (uint32_t)(s >> 32), 0, 3};
    philox(rounds, x, k0, k1);
    uint64_t w = ((uint64_t)x[1] << 32) | x[0];
    int root = cat_cum(cumpi + (int64_t)c * K, K, (double)((w >> 11) + 1) * 0x1p-53);
    if (root < 0) { rc = -(N + 2); break; }
    for (int n = 0; n < N; ++n) done[n] = 0;
    st[0] = (uint8_t)root; done[0] = 1;
    const uint32_t q = rank >> 4;
    const int d = (int)(rank & 15);
    for (int g = 0; g < G; ++g) {
      const int32_t *r = recs + (size_t)g * QUAD_REC_INTS;
      const int rt = r[0], n0 = r[1];
      if (rt < 0 || rt >= N || !done[rt] || n0 < 0 || n0 >= N) { rc = -2; break; }  /* not in walk order */
      uint8_t B[48];
      for (int call = 0; call < 3; ++call) {
        uint32_t y[4] = {q | ((uint32_t)c << 20), (uint32_t)sb, (uint32_t)n0, ((uint32_t)(sb >> 32) << 8) | (4u * (uint32_t)call)};
        philox(rounds, y, k0, k1);
        for (int b = 0; b < 16; ++b) B[16 * call + b] = (uint8_t)(y[b >> 2] >> (8 * (b & 3)));
      }
      uint32_t x24 = draw24(B, d);
      uint32_t j = x24 & 511u, t = x24 >> 9;
      int64_t idx = (((int64_t)g * C + c) * K + st[rt]) * 512 + j;
      uint32_t o;
      if (t != 0) {
        uint32_t e = de32[idx];
        o = x24 < (e >> 8) ? j : ((e >> 8) & 511u);
      } else { /* escape level */
        uint32_t z[4] = {(uint32_t)s, (uint32_t)(s >> 32), (uint32_t)n0, 1};
        philox(rounds, z, k0, k1);
        uint32_t jr = z[0] & 511u, er = dres32[idx - j + jr];
        o = (z[0] >> 9) < (er >> 9) ? jr : (er & 511u);
        ++duo_ties;
      }
      const uint32_t sv[2] = {o % (uint32_t)K, o / (uint32_t)K};
      if (sv[1] >= (uint32_t)K) { rc = -4; break; }
      for (int k = 0; k < 2; ++k) {
        const int n = r[1 + k];
        if (n < 0) continue;
        if (n >= N || done[n]) { rc = -1; break; }
        st[n] = (uint8_t)sv[k]; done[n] = 1;
      }
      if (rc) break;
    }
    if (rc) break;
    for (int n = 0; n < N; ++n)
      if (leaf_row[n] >= 0) {
        if (!done[n]) { rc = -3; break; }
        leaves[(int64_t)leaf_row[n] * leaves_pitch + col] = st[n];
      }
  }
  free(st); free(done); free(count);
  return rc;
}
