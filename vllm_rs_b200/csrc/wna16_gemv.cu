// W4A16 / W8A16: checkpoint -> B200 tile-major prepack, and the flat (CUDA-core) GEMV
// variant used for very small batches.
//
// Replaces attention_rs::kernels::ffi::{awq_repack, gptq_repack} and the small-M side of
// {marlin_awq_4bit_f16, marlin_4bit_f16, gemm_half_q_half_alt} as called from
// /root/reference/src/utils/gptq.rs:100-200,314-333.
//
// GEMV design: thread = one output channel, CTA = one 128-channel weight tile x a K split.
// A warp's 32 threads read 512 contiguous bytes of packed weights per step (16 B each =
// 32 int4 or 16 int8 K-values of their channel), dequantise in registers (lop3 / prmt magic
// numbers, exact (q - z), one fp16 rounding on the scale multiply) and accumulate x*w in
// fp32 with FHFMA against activations broadcast from shared memory.  K splits store fp32
// partials; the last CTA of a tile sums them in split order and applies bias / SiLU*mul.  HBM-bound only for M <= 4; larger M goes to the tcgen05 path.
#include <cstring>

#include <cstdlib>

#include <map>
#include <mutex>

#include "wna16.cuh"

namespace b200 {

// ------------------------------------------------------------------ prepack
template <int BITS, bool AWQ>
__global__ void prepack_kernel(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int K,
                               int N) {
  constexpr int SUB = BITS == 4 ? 2 : 4;          // 16-byte sub-rows per 64-K chunk
  constexpr int KPW = 32 / BITS;                  // K values per output word
  const size_t total = static_cast<size_t>(K) * N / KPW;
  const size_t idx = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int w = idx & 3;
  const int ch = (idx >> 2) & 127;
  size_t r = idx >> 9;
  const int sub = r % SUB;
  r /= SUB;
  const int KC = K / kChunkK;
  const int kc = r % KC;
  const int nt = r / KC;
  const int n = nt * kTileN + ch;
  const int k0 = kc * kChunkK + sub * (kChunkK / SUB) + w * KPW;
  uint32_t o = 0;
  if (BITS == 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = k0 + j;
      uint32_t qv;
      if (AWQ)
        qv = (in[static_cast<size_t>(k) * (N / 8) + n / 8] >> (4 * awq_nibble_of(n & 7))) & 0xF;
      else
        qv = (in[static_cast<size_t>(k / 8) * N + n] >> (4 * (k & 7))) & 0xF;
      const int p = (j >> 1) + ((j & 1) << 2);   // j = 2p -> p ; j = 2p+1 -> p+4
      o |= qv << (4 * p);
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = k0 + j;
      const uint32_t qv = (in[static_cast<size_t>(k / 4) * N + n] >> (8 * (k & 3))) & 0xFF;
      o |= qv << (8 * j);
    }
  }
  out[idx] = o;
}

// ------------------------------------------------------------------ flat GEMV
template <int BITS, int MB>
__global__ void __launch_bounds__(128)
wna16_gemv_kernel(const __half* __restrict__ a, const uint32_t* __restrict__ bp,
                  const __half* __restrict__ scales, const uint32_t* __restrict__ qzeros,
                  const SplitOut so, int M, int K, int N, int group_size, int qfmt,
                  int chunks_per_cta, int m_base) {
  constexpr int SUB = BITS == 4 ? 2 : 4;
  constexpr int KSUB = kChunkK / SUB;             // K values per 16-byte load
  pdl_trigger();
  pdl_wait();
  const int nt = blockIdx.x;
  const int ch = threadIdx.x;
  const int n = nt * kTileN + ch;
  const int KC = K / kChunkK;
  const int kc0 = blockIdx.y * chunks_per_cta;
  const int kc1 = min(kc0 + chunks_per_cta, KC);
  __shared__ __align__(16) __half xs[MB][kChunkK + 8];
  float acc[MB];
#pragma unroll
  for (int m = 0; m < MB; ++m) acc[m] = 0.f;
  const uint4* wp = reinterpret_cast<const uint4*>(bp) +
                    (static_cast<size_t>(nt) * KC * SUB) * kTileN + ch;
  int cur_g = -1;
  DqConst dq = make_dq(__float2half(0.f), 0);
  for (int kc = kc0; kc < kc1; ++kc) {
    __syncthreads();
    for (int i = threadIdx.x; i < MB * (kChunkK / 8); i += 128) {
      const int m = i / (kChunkK / 8), v = i % (kChunkK / 8);
      uint4 val = make_uint4(0, 0, 0, 0);
      if (m_base + m < M)
        val = *reinterpret_cast<const uint4*>(a + static_cast<size_t>(m_base + m) * K +
                                              kc * kChunkK + v * 8);
      *reinterpret_cast<uint4*>(&xs[m][v * 8]) = val;
    }
    __syncthreads();
#pragma unroll
    for (int sub = 0; sub < SUB; ++sub) {
      const uint4 pk = ld_nc_v4(wp + (static_cast<size_t>(kc) * SUB + sub) * kTileN);
      const int kbase = kc * kChunkK + sub * KSUB;
      const int g = kbase / group_size;
      if (g != cur_g) {
        cur_g = g;
        dq = make_dq(scales[static_cast<size_t>(g) * N + n], load_zero<BITS>(qzeros, qfmt, g, n, N));
      }
      uint32_t wv[KSUB / 2];
      const uint32_t words[4] = {pk.x, pk.y, pk.z, pk.w};
      if (BITS == 4) {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          uint32_t o4[4];
          dequant4_word(words[w], dq, o4);
#pragma unroll
          for (int e = 0; e < 4; ++e) wv[(w * 4 + e) % (KSUB / 2)] = o4[e];
        }
      } else {
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          uint32_t o2[2];
          dequant8_word(words[w], dq, o2);
          wv[(w * 2) % (KSUB / 2)] = o2[0];
          wv[(w * 2 + 1) % (KSUB / 2)] = o2[1];
        }
      }
#pragma unroll
      for (int m = 0; m < MB; ++m) {
        const uint32_t* xr = reinterpret_cast<const uint32_t*>(&xs[m][sub * KSUB]);
#pragma unroll
        for (int e = 0; e < KSUB / 2; ++e) {
          const uint32_t xv = xr[e];
          fma_f32_f16(acc[m], static_cast<unsigned short>(wv[e] & 0xFFFF),
                      static_cast<unsigned short>(xv & 0xFFFF));
          fma_f32_f16(acc[m], static_cast<unsigned short>(wv[e] >> 16),
                      static_cast<unsigned short>(xv >> 16));
        }
      }
    }
  }
  __shared__ int s_flag;
  const int m_hi = min(M, m_base + MB);
  if (so.splits == 1 && epi_is_plain(so)) {
    const float bv = so.bias ? __half2float(so.bias[n]) : 0.f;
#pragma unroll
    for (int m = 0; m < MB; ++m)
      if (m_base + m < M) so.c[static_cast<size_t>(m_base + m) * N + n] = __float2half_rn(acc[m] + bv);
    return;
  }
  float* part = so.partials + static_cast<size_t>(blockIdx.y) * M * N;
#pragma unroll
  for (int m = 0; m < MB; ++m)
    if (m_base + m < M) part[static_cast<size_t>(m_base + m) * N + n] = acc[m];
  split_arrive_and_reduce(so, nt, ch, 128, m_base, m_hi, 0, &s_flag, 2);
}

template <int BITS, int MB>
static void launch_gemv_mb(const WnaParams& p, cudaStream_t st, int m_base) {
  const int KC = p.K / kChunkK;
  const int tiles = p.N / kTileN;
  // enough CTAs for ~4 per SM; each CTA at least 2 chunks
  int splits = (sm_count() * 4 + tiles - 1) / tiles;
  if (splits > KC / 2) splits = KC / 2;
  if (splits < 1) splits = 1;
  while (splits > 1 && static_cast<long>(splits) * p.N > kMaxSplitCols) --splits;
  const int cpc = (KC + splits - 1) / splits;
  splits = (KC + cpc - 1) / cpc;
  SplitOut so;
  so.counters = reinterpret_cast<uint32_t*>(p.workspace);
  so.partials = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(p.workspace) + kCounterBytes);
  so.bias = p.bias; so.c = p.c; so.M = p.M; so.N = p.N; so.splits = splits; so.epilogue = p.epilogue;
  if (p.fe) so.fe = *p.fe; else memset(&so.fe, 0, sizeof(so.fe));
  dim3 grid(tiles, splits);
  launch_pdl(wna16_gemv_kernel<BITS, MB>, grid, dim3(128), 0, st, p.a, p.bp, p.scales, p.qzeros, so,
             p.M, p.K, p.N, p.group_size, p.qfmt, cpc, m_base);
  B200_CUDA_LAUNCH_CHECK();
}

template <int BITS>
static void launch_gemv_bits(const WnaParams& p, cudaStream_t st) {
  int m_base = 0;
  while (m_base < p.M) {
    const int rem = p.M - m_base;
    if (rem >= 16) {
      launch_gemv_mb<BITS, 16>(p, st, m_base);
      m_base += 16;
    } else if (rem > 4) {
      launch_gemv_mb<BITS, 8>(p, st, m_base);
      m_base += 8;
    } else if (rem > 2) {
      launch_gemv_mb<BITS, 4>(p, st, m_base);
      m_base += 4;
    } else if (rem == 2) {
      launch_gemv_mb<BITS, 2>(p, st, m_base);
      m_base += 2;
    } else {
      launch_gemv_mb<BITS, 1>(p, st, m_base);
      m_base += 1;
    }
  }
}

void launch_wna16_gemv(const WnaParams& p, cudaStream_t st) {
  if (p.bits == 4)
    launch_gemv_bits<4>(p, st);
  else
    launch_gemv_bits<8>(p, st);
}

static bool check_wna(const WnaParams& p, const char** why) {
  if (!(p.bits == 4 || p.bits == 8)) { *why = "wna16: bits must be 4 or 8"; return false; }
  if (p.M <= 0 || p.K <= 0 || p.N <= 0) { *why = "wna16: non-positive dims"; return false; }
  if (p.N % kTileN != 0) { *why = "wna16: N must be a multiple of 128"; return false; }
  if (p.K % kChunkK != 0) { *why = "wna16: K must be a multiple of 64"; return false; }
  const int g = p.group_size;
  if (g <= 0 || g % 32 != 0 || p.K % g != 0) { *why = "wna16: group_size must divide K and be a multiple of 32"; return false; }
  if (!p.a || !p.bp || !p.scales || !p.c || !p.workspace) { *why = "wna16: NULL pointer"; return false; }
  if (p.epilogue == B200_EPI_SILU_MUL && (p.N / kTileN) % 2 != 0) { *why = "wna16: SILU_MUL needs an even number of 128-channel tiles"; return false; }
  if (p.N / kTileN > kCounterBytes / 4) { *why = "wna16: N too large for the tile counters"; return false; }
  return true;
}

}  // namespace b200

namespace b200 {
int gemv_max_m() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("B200_GEMV_MAX_M");
    v = e ? atoi(e) : 0;
    if (v < 0) v = 0;
  }
  return v;
}
}  // namespace b200

extern "C" {
using namespace b200;

void b200_awq_prepack(const void* in, void* out, int32_t k, int32_t n_packed, int32_t bits,
                      int64_t stream) {
  B200_CHECK_ARG(bits == 4, "awq_prepack: only 4-bit AWQ is supported");
  const int n = n_packed * 8;
  B200_CHECK_ARG(k % kChunkK == 0 && n % kTileN == 0, "awq_prepack: need K % 64 == 0, N % 128 == 0");
  const size_t total = static_cast<size_t>(k) * n / 8;
  prepack_kernel<4, true><<<static_cast<unsigned>((total + 255) / 256), 256, 0,
                            reinterpret_cast<cudaStream_t>(stream)>>>(
      (const uint32_t*)in, (uint32_t*)out, k, n);
  B200_CUDA_LAUNCH_CHECK();
}

void b200_gptq_prepack(const void* in, void* out, int32_t k_packed, int32_t n, int32_t bits,
                       int64_t stream) {
  B200_CHECK_ARG(bits == 4 || bits == 8, "gptq_prepack: bits must be 4 or 8");
  const int k = k_packed * (32 / bits);
  B200_CHECK_ARG(k % kChunkK == 0 && n % kTileN == 0, "gptq_prepack: need K % 64 == 0, N % 128 == 0");
  const size_t total = static_cast<size_t>(k_packed) * n;
  const unsigned blocks = static_cast<unsigned>((total + 255) / 256);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (bits == 4)
    prepack_kernel<4, false><<<blocks, 256, 0, st>>>((const uint32_t*)in, (uint32_t*)out, k, n);
  else
    prepack_kernel<8, false><<<blocks, 256, 0, st>>>((const uint32_t*)in, (uint32_t*)out, k, n);
  B200_CUDA_LAUNCH_CHECK();
}

int64_t b200_wna16_workspace_bytes(int32_t max_m, int32_t max_n) {
  // [16 KB arrival counters][fp32 split-K partials]: launchers keep splits * N <= kMaxSplitCols
  // unless splits == 1, so max(max_n, kMaxSplitCols) columns always suffice.
  const int64_t m = max_m < 16 ? 16 : max_m;
  const int64_t cols = max_n > kSkSlotCols ? max_n : kSkSlotCols;
  return kCounterBytes + m * cols * 4 + 256;
}

void b200_wna16_gemm_ex(const void* a, const void* b_packed, const void* scales,
                        const void* qzeros, const void* bias, void* c, int32_t m, int32_t k,
                        int32_t n, int32_t bits, int32_t group_size, int32_t qfmt,
                        int32_t epilogue, int32_t variant, void* workspace, int64_t stream) {
  if (m <= 0) return;
  WnaParams p;
  p.a = (const __half*)a;
  p.bp = (const uint32_t*)b_packed;
  p.scales = (const __half*)scales;
  p.qzeros = (const uint32_t*)qzeros;
  p.bias = (const __half*)bias;
  p.c = (__half*)c;
  p.M = m; p.K = k; p.N = n; p.bits = bits;
  p.group_size = group_size <= 0 ? k : group_size;
  p.qfmt = qfmt; p.epilogue = epilogue; p.workspace = workspace;
  p.no_split = (variant == 3) ? 1 : 0;
  p.fe = nullptr;
  const char* why = nullptr;
  if (!check_wna(p, &why)) { set_error(1, why); return; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (variant == 5 || variant == 6) {                  // stream-K decode kernel: 5 = tcgen05 flavour, 6 = mma.sync flavour
    if (variant == 6 && bits != 4) { set_error(1, "wna16: the mma.sync stream-K flavour is 4-bit only"); return; }
    if (launch_wna16_sk(p, st, true, false, variant == 6 ? 1 : 0)) return;
    set_error(1, "wna16: shape unsupported by the stream-K decode kernel");
    return;
  }
  if (variant == 2 || variant == 3 || variant == 4 || (variant == 0 && m > gemv_max_m())) {
    if (variant == 0 && launch_wna16_sk(p, st)) return;        // persistent stream-K decode kernel (M <= 32)
    if (variant != 4 && launch_wna16_tc_dec(p, st)) return;    // round-1 decode kernel (M <= 64; A/B arm: variant 2)
    if (launch_wna16_tc(p, st)) return;
    if (variant >= 2) { set_error(1, "wna16: shape unsupported by the tcgen05 variant"); return; }
  }
  launch_wna16_gemv(p, st);
}

void b200_marlin_awq_4bit_f16(const void* a, const int32_t* b, const void* scales,
                              const void* qzeros, const void* g_idx, void* c, int32_t m,
                              int32_t k, int32_t n, const void* workspace, int32_t group_size,
                              int64_t stream) {
  (void)g_idx;   // desc_act is rejected by the reference loader (utils/mod.rs:1315-1317)
  b200_wna16_gemm_ex(a, b, scales, qzeros, nullptr, c, m, k, n, 4, group_size, B200_QFMT_AWQ,
                     B200_EPI_NONE, 0, const_cast<void*>(workspace), stream);
}

void b200_marlin_4bit_f16(const void* a, const int32_t* b, const void* scales,
                          const void* qzeros, const void* g_idx, void* c, int32_t m, int32_t k,
                          int32_t n, const void* workspace, int32_t group_size, int64_t stream) {
  (void)g_idx;
  b200_wna16_gemm_ex(a, b, scales, qzeros, nullptr, c, m, k, n, 4, group_size, B200_QFMT_GPTQ,
                     B200_EPI_NONE, 0, const_cast<void*>(workspace), stream);
}

// bf16 flavour of the Marlin entry points (gptq.rs:150-178; bf16 is the reference's default dtype,
// utils/mod.rs:1919-1925): A, scales and C are bf16.  The stream-K tcgen05 kernel (wna16_sk.cu) does the work:
// exact (q - z) as bf16 into TMEM, kind::f16 MMA with bf16 operands, fp32 group scale after the MMA, bf16 output.
// It takes 32 token rows per launch; larger m runs as consecutive 32-row launches on the stream (a prefill-sized
// bf16 call therefore re-streams the weight once per 32 rows: correct, not fast -- the fp16 entry points are the
// tuned path).
static void marlin_bf16(const void* a, const int32_t* b, const void* scales, const void* qzeros, void* c, int32_t m,
                        int32_t k, int32_t n, const void* workspace, int32_t group_size, int32_t qfmt,
                        int64_t stream) {
  if (m <= 0) return;
  WnaParams p;
  p.bp = (const uint32_t*)b;
  p.scales = (const __half*)scales;
  p.qzeros = (const uint32_t*)qzeros;
  p.bias = nullptr;
  p.K = k; p.N = n; p.bits = 4;
  p.group_size = group_size <= 0 ? k : group_size;
  p.qfmt = qfmt; p.epilogue = B200_EPI_NONE; p.workspace = const_cast<void*>(workspace);
  p.no_split = 0;
  p.fe = nullptr;
  p.a = (const __half*)a; p.c = (__half*)c; p.M = m < 32 ? m : 32;
  const char* why = nullptr;
  if (!check_wna(p, &why)) { set_error(1, why); return; }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  for (int m0 = 0; m0 < m; m0 += 32) {
    p.a = reinterpret_cast<const __half*>(a) + static_cast<size_t>(m0) * k;      // 2-byte elements either way
    p.c = reinterpret_cast<__half*>(c) + static_cast<size_t>(m0) * n;
    p.M = m - m0 < 32 ? m - m0 : 32;
    if (!launch_wna16_sk(p, st, true, true)) {
      set_error(1, "marlin bf16: needs K % 128 == 0 and group_size = 128 * 2^j (or one group per channel)");
      return;
    }
  }
}

void b200_marlin_awq_4bit_bf16(const void* a, const int32_t* b, const void* scales, const void* qzeros,
                               const void* g_idx, void* c, int32_t m, int32_t k, int32_t n, const void* workspace,
                               int32_t group_size, int64_t stream) {
  (void)g_idx;
  marlin_bf16(a, b, scales, qzeros, c, m, k, n, workspace, group_size, B200_QFMT_AWQ, stream);
}

void b200_marlin_4bit_bf16(const void* a, const int32_t* b, const void* scales, const void* qzeros,
                           const void* g_idx, void* c, int32_t m, int32_t k, int32_t n, const void* workspace,
                           int32_t group_size, int64_t stream) {
  (void)g_idx;
  marlin_bf16(a, b, scales, qzeros, c, m, k, n, workspace, group_size, B200_QFMT_GPTQ, stream);
}

void b200_gemm_half_q_half_8bit(const void* a, const uint32_t* qweight, const uint32_t* qzeros,
                                const void* scales, const int32_t* g_idx, void* c, int32_t m,
                                int32_t n, int32_t k, int32_t group_size, const void* workspace,
                                int64_t stream) {
  (void)g_idx;
  b200_wna16_gemm_ex(a, qweight, scales, qzeros, nullptr, c, m, k, n, 8, group_size,
                     B200_QFMT_GPTQ, B200_EPI_NONE, 0, const_cast<void*>(workspace), stream);
}

// ---- gemm_half_q_half_alt with the reference's EXACT signature (gptq.rs:183-195, 316-324): the caller hands over
// the checkpoint-format GPTQ tensor [K/pack, N] (no repack at load on this path), no group size and no workspace.
// The group size is what the reference kernel derives from g_idx (one 8-byte read, cached per tensor); the B200
// tile-major copy of the weight is made on first use and kept (the repack the Marlin path does at load,
// wna16.rs:127-152, deferred); the split-K workspace is library-owned and grows on demand.
namespace {
struct AltEntry { uint32_t* packed; int group_size; };
struct AltKey {
  const void* qw; int k, n, bits;
  bool operator<(const AltKey& o) const {
    if (qw != o.qw) return qw < o.qw;
    if (k != o.k) return k < o.k;
    if (n != o.n) return n < o.n;
    return bits < o.bits;
  }
};
std::mutex g_alt_mu;
std::map<AltKey, AltEntry> g_alt;
void* g_alt_ws = nullptr;
int64_t g_alt_ws_bytes = 0;
}  // namespace

void b200_gemm_half_q_half_alt(const void* a, const uint32_t* qweight, const uint32_t* qzeros, const void* scales,
                               const int32_t* g_idx, void* c, int32_t m, int32_t n, int32_t k, int32_t bits,
                               int64_t stream) {
  if (m <= 0) return;
  B200_CHECK_ARG(bits == 4 || bits == 8, "gemm_half_q_half_alt: bits must be 4 or 8");
  B200_CHECK_ARG(a && qweight && scales && c, "gemm_half_q_half_alt: NULL pointer");
  B200_CHECK_ARG(k % kChunkK == 0 && n % kTileN == 0, "gemm_half_q_half_alt: need K % 64 == 0, N % 128 == 0");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  std::lock_guard<std::mutex> lock(g_alt_mu);
  const AltKey key{qweight, k, n, bits};
  auto it = g_alt.find(key);
  if (it == g_alt.end()) {
    B200_CHECK_ARG(g_idx != nullptr, "gemm_half_q_half_alt: g_idx is required (the group size is derived from it)");
    int32_t first = 0, last = 0;
    if (cudaMemcpyAsync(&first, g_idx, 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaMemcpyAsync(&last, g_idx + (k - 1), 4, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess) {
      set_error(2, "gemm_half_q_half_alt: reading g_idx failed");
      return;
    }
    const int groups = last - first + 1;           // a TP shard's g_idx starts at its first global group
    if (groups <= 0 || k % groups != 0 || (k / groups) % 32 != 0) {
      set_error(1, "gemm_half_q_half_alt: g_idx is not k / group_size (act-order checkpoints are rejected by the loader)");
      return;
    }
    AltEntry e;
    e.group_size = k / groups;
    if (cudaMalloc(&e.packed, static_cast<size_t>(k) * n * bits / 8) != cudaSuccess) {
      set_error(2, "gemm_half_q_half_alt: cudaMalloc of the tile-major weight copy failed");
      return;
    }
    b200_gptq_prepack(qweight, e.packed, k * bits / 32, n, bits, stream);
    it = g_alt.emplace(key, e).first;
  }
  const int64_t need = b200_wna16_workspace_bytes(m, n);
  if (need > g_alt_ws_bytes) {
    cudaStreamSynchronize(st);                      // a previous call may still be using the old block
    if (g_alt_ws) cudaFree(g_alt_ws);
    g_alt_ws = nullptr;
    g_alt_ws_bytes = 0;
    if (cudaMalloc(&g_alt_ws, need) != cudaSuccess || cudaMemsetAsync(g_alt_ws, 0, need, st) != cudaSuccess) {
      set_error(2, "gemm_half_q_half_alt: workspace allocation failed");
      return;
    }
    g_alt_ws_bytes = need;
  }
  b200_wna16_gemm_ex(a, it->second.packed, scales, qzeros, nullptr, c, m, k, n, bits, it->second.group_size,
                     B200_QFMT_GPTQ, B200_EPI_NONE, 0, g_alt_ws, stream);
}

// Drop the tile-major copies and the workspace b200_gemm_half_q_half_alt made (model unload).
void b200_gemm_half_q_half_alt_release(void) {
  std::lock_guard<std::mutex> lock(g_alt_mu);
  cudaDeviceSynchronize();
  for (auto& kv : g_alt) cudaFree(kv.second.packed);
  g_alt.clear();
  if (g_alt_ws) cudaFree(g_alt_ws);
  g_alt_ws = nullptr;
  g_alt_ws_bytes = 0;
}

}  // extern "C"
