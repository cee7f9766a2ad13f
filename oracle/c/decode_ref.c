/* CPU restatement (plain C + OpenMP) of one decode step of the vllm.rs forward path.
 *
 * TEST INFRASTRUCTURE / CPU BASELINE (see oracle/__init__.py): this is what bench.py times as
 * "the reference's CPU path" (kind = "port": the reference itself has no CPU path for
 * AWQ/GPTQ, src/utils/gptq.rs:212-222, and cannot be built here).
 * Same op sequence and fp16 rounding points as oracle/model.py:
 *   llama.rs:107-131 (layer), attention.rs:648-839, mlp.rs:451-469, others.rs:11-66,
 *   rotary_emb.rs:88-104 (NeoX), kvcache_allocator.rs:150-173 (cache layout L2),
 *   runner.rs:1243-1388 (decode metadata), AWQ format per wna16.rs:127-152 / SURVEY A.4.
 *
 * Build: gcc -O3 -march=native -fopenmp -shared -fPIC decode_ref.c -o ../_build/libdecode_ref.so -lm
 */
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef _Float16 f16;

static const int AWQ_SHIFT[8] = {0, 16, 4, 20, 8, 24, 12, 28}; /* channel r -> bit offset */

/* y[M,N] = fp16( x[M,K] @ fp16((q - z) * s) + bias ), fp32 accumulation. AWQ packing. */
void ref_awq_gemm(const f16* x, const int32_t* qweight, const int32_t* qzeros, const f16* scales,
                  const f16* bias, f16* y, int M, int K, int N, int G) {
  const int NB = N / 8;
  float* xf = (float*)malloc(sizeof(float) * (size_t)M * K);
  for (size_t i = 0; i < (size_t)M * K; ++i) xf[i] = (float)x[i];
#pragma omp parallel for schedule(static)
  for (int nb = 0; nb < NB; ++nb) {
    float acc[64][8];
    for (int m0 = 0; m0 < M; m0 += 64) {
      const int mc = M - m0 < 64 ? M - m0 : 64;
      for (int m = 0; m < mc; ++m)
        for (int r = 0; r < 8; ++r) acc[m][r] = 0.f;
      float zf[8], sf[8];
      for (int k = 0; k < K; ++k) {
        if (k % G == 0) {
          const uint32_t zw = (uint32_t)qzeros[(size_t)(k / G) * NB + nb];
          for (int r = 0; r < 8; ++r) {
            zf[r] = (float)((zw >> AWQ_SHIFT[r]) & 0xF);
            sf[r] = (float)scales[(size_t)(k / G) * N + nb * 8 + r];
          }
        }
        const uint32_t w = (uint32_t)qweight[(size_t)k * NB + nb];
        float wf[8];
        for (int r = 0; r < 8; ++r) /* one rounding to fp16, as a register-dequant kernel does */
          wf[r] = (float)(f16)(((float)((w >> AWQ_SHIFT[r]) & 0xF) - zf[r]) * sf[r]);
        for (int m = 0; m < mc; ++m) {
          const float xv = xf[(size_t)(m0 + m) * K + k];
          for (int r = 0; r < 8; ++r) acc[m][r] += xv * wf[r];
        }
      }
      for (int m = 0; m < mc; ++m)
        for (int r = 0; r < 8; ++r) {
          float v = acc[m][r];
          if (bias) v += (float)bias[nb * 8 + r];
          y[(size_t)(m0 + m) * N + nb * 8 + r] = (f16)v;
        }
    }
  }
  free(xf);
}

/* y[M,N] (fp32 out) = x[M,K] @ W[N,K]^T, dense fp16 (lm_head); logits rounded to fp16 first */
void ref_dense_gemm(const f16* x, const f16* w, float* y, int M, int N, int K) {
#pragma omp parallel for schedule(static)
  for (int n = 0; n < N; ++n) {
    const f16* wr = w + (size_t)n * K;
    for (int m = 0; m < M; ++m) {
      const f16* xr = x + (size_t)m * K;
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += (float)xr[k] * (float)wr[k];
      y[(size_t)m * N + n] = (float)(f16)acc;
    }
  }
}

void ref_rmsnorm(const f16* x, const f16* w, f16* out, int T, int hidden, float eps) {
#pragma omp parallel for
  for (int t = 0; t < T; ++t) {
    float ss = 0.f;
    for (int i = 0; i < hidden; ++i) {
      const float v = (float)x[(size_t)t * hidden + i];
      ss += v * v;
    }
    const float inv = 1.f / sqrtf(ss / hidden + eps);
    for (int i = 0; i < hidden; ++i)
      out[(size_t)t * hidden + i] = (f16)((float)x[(size_t)t * hidden + i] * inv * (float)w[i]);
  }
}

/* NeoX rotary in place on q [T,H,D] and k [T,Hkv,D]; then write k,v into the paged cache */
void ref_rope_kv_write(f16* q, f16* k, const f16* v, const f16* cos, const f16* sin,
                       const int64_t* positions, const int64_t* slots, f16* k_cache, f16* v_cache,
                       int T, int H, int Hkv, int D) {
  const int half = D / 2;
#pragma omp parallel for
  for (int t = 0; t < T; ++t) {
    const f16* c = cos + positions[t] * half;
    const f16* s = sin + positions[t] * half;
    for (int hh = 0; hh < H + Hkv; ++hh) {
      f16* p = hh < H ? q + ((size_t)t * H + hh) * D : k + ((size_t)t * Hkv + (hh - H)) * D;
      for (int i = 0; i < half; ++i) {
        const float x0 = (float)p[i], x1 = (float)p[i + half];
        p[i] = (f16)(x0 * (float)c[i] - x1 * (float)s[i]);
        p[i + half] = (f16)(x1 * (float)c[i] + x0 * (float)s[i]);
      }
    }
    if (slots[t] >= 0) {
      memcpy(k_cache + (size_t)slots[t] * Hkv * D, k + (size_t)t * Hkv * D, sizeof(f16) * Hkv * D);
      memcpy(v_cache + (size_t)slots[t] * Hkv * D, v + (size_t)t * Hkv * D, sizeof(f16) * Hkv * D);
    }
  }
}

/* out[B,H,D]: one query token per sequence over its block table; cache [nb, bs, Hkv, D] */
void ref_paged_attn_decode(const f16* q, const f16* k_cache, const f16* v_cache,
                           const uint32_t* block_tables, const uint32_t* context_lens, f16* out,
                           int B, int H, int Hkv, int D, int bs, int max_blocks, float scale) {
  const int G = H / Hkv;
#pragma omp parallel for collapse(2) schedule(dynamic)
  for (int b = 0; b < B; ++b)
    for (int h = 0; h < H; ++h) {
      const int ctx = (int)context_lens[b];
      f16* o = out + ((size_t)b * H + h) * D;
      if (ctx == 0) {
        memset(o, 0, sizeof(f16) * D);
        continue;
      }
      const int hk = h / G;
      float qf[256], acc[256];
      for (int d = 0; d < D; ++d) {
        qf[d] = (float)q[((size_t)b * H + h) * D + d];
        acc[d] = 0.f;
      }
      float m = -INFINITY, l = 0.f;
      for (int t = 0; t < ctx; ++t) {
        const size_t row = ((size_t)block_tables[(size_t)b * max_blocks + t / bs] * bs + t % bs) * Hkv + hk;
        const f16* kr = k_cache + row * D;
        const f16* vr = v_cache + row * D;
        float s = 0.f;
        for (int d = 0; d < D; ++d) s += qf[d] * (float)kr[d];
        s *= scale;
        const float mn = s > m ? s : m;
        const float alpha = expf(m - mn), p = expf(s - mn);
        l = l * alpha + p;
        for (int d = 0; d < D; ++d) acc[d] = acc[d] * alpha + p * (float)vr[d];
        m = mn;
      }
      for (int d = 0; d < D; ++d) o[d] = (f16)(acc[d] / l);
    }
}

void ref_silu_mul(const f16* gate, const f16* up, f16* out, size_t n) {
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) {
    const float g = (float)gate[i];
    const f16 act = (f16)(g / (1.f + expf(-g)));
    out[i] = (f16)((float)act * (float)up[i]);
  }
}

void ref_add(f16* x, const f16* y, size_t n) {
#pragma omp parallel for
  for (size_t i = 0; i < n; ++i) x[i] = (f16)((float)x[i] + (float)y[i]);
}

typedef struct {
  const f16 *input_norm, *post_norm;
  const int32_t *q_w, *q_z, *k_w, *k_z, *v_w, *v_z, *o_w, *o_z, *g_w, *g_z, *u_w, *u_z, *d_w, *d_z;
  const f16 *q_s, *k_s, *v_s, *o_s, *g_s, *u_s, *d_s;
  f16 *k_cache, *v_cache;
} ref_layer;

/* x [B, hidden] is updated in place through `num_layers` decoder layers (7 separate
 * quantised GEMMs per layer, like the reference: SURVEY.md F6). */
void ref_decode_layers(f16* x, const ref_layer* layers, int num_layers, const f16* cos,
                       const f16* sin, const int64_t* positions, const int64_t* slots,
                       const uint32_t* block_tables, const uint32_t* context_lens, int B,
                       int hidden, int H, int Hkv, int D, int inter, int bs, int max_blocks,
                       int G, float eps) {
  f16* xn = malloc(sizeof(f16) * (size_t)B * hidden);
  f16* q = malloc(sizeof(f16) * (size_t)B * H * D);
  f16* k = malloc(sizeof(f16) * (size_t)B * Hkv * D);
  f16* v = malloc(sizeof(f16) * (size_t)B * Hkv * D);
  f16* a = malloc(sizeof(f16) * (size_t)B * H * D);
  f16* o = malloc(sizeof(f16) * (size_t)B * hidden);
  f16* g = malloc(sizeof(f16) * (size_t)B * inter);
  f16* u = malloc(sizeof(f16) * (size_t)B * inter);
  const float scale = 1.f / sqrtf((float)D);
  for (int l = 0; l < num_layers; ++l) {
    const ref_layer* w = &layers[l];
    ref_rmsnorm(x, w->input_norm, xn, B, hidden, eps);
    ref_awq_gemm(xn, w->q_w, w->q_z, w->q_s, 0, q, B, hidden, H * D, G);
    ref_awq_gemm(xn, w->k_w, w->k_z, w->k_s, 0, k, B, hidden, Hkv * D, G);
    ref_awq_gemm(xn, w->v_w, w->v_z, w->v_s, 0, v, B, hidden, Hkv * D, G);
    ref_rope_kv_write(q, k, v, cos, sin, positions, slots, w->k_cache, w->v_cache, B, H, Hkv, D);
    ref_paged_attn_decode(q, w->k_cache, w->v_cache, block_tables, context_lens, a, B, H, Hkv, D,
                          bs, max_blocks, scale);
    ref_awq_gemm(a, w->o_w, w->o_z, w->o_s, 0, o, B, H * D, hidden, G);
    ref_add(x, o, (size_t)B * hidden);
    ref_rmsnorm(x, w->post_norm, xn, B, hidden, eps);
    ref_awq_gemm(xn, w->g_w, w->g_z, w->g_s, 0, g, B, hidden, inter, G);
    ref_awq_gemm(xn, w->u_w, w->u_z, w->u_s, 0, u, B, hidden, inter, G);
    ref_silu_mul(g, u, g, (size_t)B * inter);
    ref_awq_gemm(g, w->d_w, w->d_z, w->d_s, 0, o, B, inter, hidden, G);
    ref_add(x, o, (size_t)B * hidden);
  }
  free(xn); free(q); free(k); free(v); free(a); free(o); free(g); free(u);
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the timed baseline asks for the host's cores itself */
void ref_set_num_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
}

int ref_num_threads(void) {
  int n = 1;
#pragma omp parallel
  {
#pragma omp master
    n = omp_get_num_threads();
  }
  return n;
}
