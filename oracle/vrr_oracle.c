/*
 * vrr_oracle.c - plain-C restatement of the reference's tables and attention core.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): loaded by tests/ only, never by the product.
 * Citations are relative to /root/reference.  Double precision throughout; compared against the
 * numpy oracle and the golden fixtures generated from the unmodified reference
 * (tests/test_oracle_c.py).  Built by __graft_entry__.build_oracle() with gcc.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

/* RelativePositionalEncoding.__init__ - models/positional_encoding.py:67-75: idx[i][j] = i - j + L - 1 */
void vrr_c_relative_index(int L, int64_t* idx) {
  for (int i = 0; i < L; ++i)
    for (int j = 0; j < L; ++j) {
      int64_t v = (int64_t)i - j + (L - 1);
      if (v < 0) v = 0;
      if (v > 2 * L - 2) v = 2 * L - 2;
      idx[(size_t)i * L + j] = v;
    }
}

/* init_t_xy - models/positional_encoding.py:198-214: t_x = t % w, t_y = floor(t / w) */
void vrr_c_grid_coords(int h, int w, float* tx, float* ty) {
  for (int t = 0; t < h * w; ++t) {
    tx[t] = (float)(t % w);
    ty[t] = (float)(t / w);
  }
}

/* RoPEMixed.get_freqs_cis scramble - models/positional_encoding.py:337-342:
 * output [h'][n'] reads source head hs, source position ps with (hs, ps) = divmod(n' * H + h', N) */
void vrr_c_mixed_scramble(int H, int N, int64_t* hs, int64_t* ps) {
  for (int hp = 0; hp < H; ++hp)
    for (int n = 0; n < N; ++n) {
      int64_t flat = (int64_t)n * H + hp;
      hs[(size_t)hp * N + n] = flat / N;
      ps[(size_t)hp * N + n] = flat % N;
    }
}

/* PolynomialRPE.get_bias coordinates - models/positional_encoding.py:134-142: y = p % g, x = p / g */
void vrr_c_poly_l1(int g, int64_t* dist) {
  int np = g * g;
  for (int p = 0; p < np; ++p)
    for (int q = 0; q < np; ++q)
      dist[(size_t)p * np + q] = llabs((long long)(p % g) - (q % g)) + llabs((long long)(p / g) - (q / g));
}

/* bias[h][i][j] for the three modes of include/vrr.h (0 none, 1 table [H][2N-1], 2 poly [Hc][len]) */
static double bias_at(int mode, const double* param, int heads, int len, int grid, int N, int h, int i, int j) {
  if (mode == 1) return param[(size_t)h * len + (i - j + N - 1)];
  if (mode == 2) {
    if (i == 0 || j == 0) return 0.0;
    int pi = i - 1, pj = j - 1;
    double d = (double)(abs(pi % grid - pj % grid) + abs(pi / grid - pj / grid));
    const double* c = param + (size_t)(heads == 1 ? 0 : h) * len;
    double pw = 1.0, acc = 0.0;
    for (int k = 0; k < len; ++k) {
      acc += c[k] * pw;
      pw *= d;
    }
    return acc;
  }
  return 0.0;
}

/* Attention core forward - models/vit.py:71-88.  q,k,v [B][H][N][D] (already rotated for RoPE);
 * out [B][N][H*D]; lse [B][H][N].  scale applied after q.k, before the bias. */
void vrr_c_attn_fwd(const double* q, const double* k, const double* v, int B, int H, int N, int D, double scale,
                    int mode, const double* param, int heads, int len, int grid, double* out, double* lse) {
  double* s = (double*)malloc((size_t)N * sizeof(double));
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h) {
      const size_t base = ((size_t)b * H + h) * N * D;
      for (int i = 0; i < N; ++i) {
        double m = -INFINITY;
        for (int j = 0; j < N; ++j) {
          double acc = 0.0;
          for (int d = 0; d < D; ++d) acc += q[base + (size_t)i * D + d] * k[base + (size_t)j * D + d];
          s[j] = acc * scale + bias_at(mode, param, heads, len, grid, N, h, i, j);
          if (s[j] > m) m = s[j];
        }
        double l = 0.0;
        for (int j = 0; j < N; ++j) {
          s[j] = exp(s[j] - m);
          l += s[j];
        }
        for (int d = 0; d < D; ++d) {
          double acc = 0.0;
          for (int j = 0; j < N; ++j) acc += s[j] * v[base + (size_t)j * D + d];
          out[((size_t)b * N + i) * ((size_t)H * D) + (size_t)h * D + d] = acc / l;
        }
        lse[((size_t)b * H + h) * N + i] = m + log(l);
      }
    }
  free(s);
}

/* Attention core backward.  d_out [B][N][H*D]; dq,dk,dv [B][H][N][D];
 * d_bias [H][N][N] = sum_b dS (caller zero-fills). */
void vrr_c_attn_bwd(const double* q, const double* k, const double* v, const double* out, const double* d_out,
                    const double* lse, int B, int H, int N, int D, double scale, int mode, const double* param,
                    int heads, int len, int grid, double* dq, double* dk, double* dv, double* d_bias) {
  const size_t E = (size_t)H * D;
  for (size_t t = 0; t < (size_t)B * H * N * D; ++t) dq[t] = dk[t] = dv[t] = 0.0;
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h) {
      const size_t base = ((size_t)b * H + h) * N * D;
      for (int i = 0; i < N; ++i) {
        const double* go = d_out + ((size_t)b * N + i) * E + (size_t)h * D;
        const double* oo = out + ((size_t)b * N + i) * E + (size_t)h * D;
        double delta = 0.0;
        for (int d = 0; d < D; ++d) delta += go[d] * oo[d];
        const double li = lse[((size_t)b * H + h) * N + i];
        for (int j = 0; j < N; ++j) {
          double sc = 0.0, dp = 0.0;
          for (int d = 0; d < D; ++d) {
            sc += q[base + (size_t)i * D + d] * k[base + (size_t)j * D + d];
            dp += go[d] * v[base + (size_t)j * D + d];
          }
          const double p = exp(sc * scale + bias_at(mode, param, heads, len, grid, N, h, i, j) - li);
          const double ds = p * (dp - delta);
          d_bias[((size_t)h * N + i) * N + j] += ds;
          for (int d = 0; d < D; ++d) {
            dv[base + (size_t)j * D + d] += p * go[d];
            dq[base + (size_t)i * D + d] += scale * ds * k[base + (size_t)j * D + d];
            dk[base + (size_t)j * D + d] += scale * ds * q[base + (size_t)i * D + d];
          }
        }
      }
    }
}
