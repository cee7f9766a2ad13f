// Device sampler: temperature softmax -> top-k (k <= 256) -> top-p -> multinomial, one CTA per
// batch row, plus the frequency / presence penalties; and the host mirror of the reference's
// seeded step-seed stream.
//
// Replaces attention_rs::sampler::Sampler::sample_cuda as called from
//   LogitsProcessor::sample_with_strategy  /root/reference/src/utils/logits_processor.rs:199-231
// (k, p, t mapping of :208-216) with the semantics of the source-visible CPU branch
//   sample_topk_topp / sample_topk / sample_topp  logits_processor.rs:84-196
// i.e. probabilities are the FULL-vocabulary softmax, the top-p cut walks the un-renormalised top-k
// probabilities and zeroes everything after the cumulative mass first reaches p, and the draw is
// WeightedIndex over what is left.  Penalties: logits_processor.rs:286-343.
// The per-step seed is StdRng::seed_from_u64(seed).next_u64() (logits_processor.rs:29-30,222-225):
// rand 0.9 StdRng = ChaCha12 keyed by a PCG32 expansion of the u64 (restated from the published
// algorithm; the `rand` crate is not vendored).  How the reference kernel turns (seed, row) into a
// uniform is not source-visible: row_uniform() below is this library's definition.
//
// Bound: HBM / L2 (the row of fp32 logits is read 8 times: max, sum, 4 radix-select passes over the
// order-preserving integer image of the logits, 2 collection passes; 16 MB of logits at b = 32 stay
// in the 126 MB L2).  Removes the 16 MiB D2H of the logits per step in non-greedy mode.
#include <cstring>

#include "../../include/b200_vllm.h"
#include "common.cuh"

namespace b200 {
namespace {

constexpr int kSampThreads = 1024;
constexpr int kMaxK = 256;

__device__ __forceinline__ uint32_t sortable(float f) {          // larger float -> larger unsigned
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unsortable(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float row_uniform(uint64_t seed, int row) {
  const uint64_t h = splitmix64(seed ^ (static_cast<uint64_t>(row + 1) * 0xD1342543DE82EF95ull));
  return static_cast<float>(h >> 40) * (1.0f / 16777216.0f);
}

template <typename T, typename Op>
__device__ __forceinline__ T block_reduce(T v, Op op, T* scratch, T identity) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T r = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : identity;
  if (warp == 0) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r = op(r, __shfl_xor_sync(0xffffffffu, r, o));
    if (lane == 0) scratch[0] = r;
  }
  __syncthreads();
  r = scratch[0];
  __syncthreads();
  return r;
}

// out_ids[b] ~ top-k / top-p / temperature multinomial of logits[b, :]; seed read from device memory
// (graph replays change it with a copy, not a re-capture).
__global__ void __launch_bounds__(kSampThreads)
sample_topk_topp_kernel(uint32_t* __restrict__ out_ids, const float* __restrict__ logits, int V, int k,
                        float top_p, float inv_t, const uint64_t* __restrict__ seed_ptr) {
  const int row = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* x = logits + static_cast<size_t>(row) * V;
  __shared__ float red_f[32];
  __shared__ double red_d[32];
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_need;
  __shared__ uint32_t warp_gt[32], warp_eq[32];
  __shared__ uint64_t cand[kMaxK];                 // (sortable logit << 32) | ~index, sorted descending
  pdl_trigger();
  pdl_wait();

  // ---- pass 1, 2: max and softmax denominator of logits / T
  float mx = -INFINITY;
  for (int i = tid; i < V; i += kSampThreads) mx = fmaxf(mx, x[i]);
  mx = block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, red_f, -INFINITY);
  double sum = 0.0;
  for (int i = tid; i < V; i += kSampThreads) sum += static_cast<double>(__expf((x[i] - mx) * inv_t));
  sum = block_reduce(sum, [](double a, double b) { return a + b; }, red_d, 0.0);

  // ---- passes 3-6: radix select of the k-th largest key (8 bits per pass, MSB first)
  if (tid == 0) { s_prefix = 0; s_need = static_cast<uint32_t>(k); }
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    for (int i = tid; i < 256; i += kSampThreads) hist[i] = 0;
    __syncthreads();
    const uint32_t prefix = s_prefix;
    const uint32_t pmask = pass == 0 ? 0u : (0xFFFFFFFFu << (shift + 8));
    for (int i = tid; i < V; i += kSampThreads) {
      const uint32_t key = sortable(x[i]);
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 0xFF], 1u);
    }
    __syncthreads();
    if (tid == 0) {                                   // walk the 256 bins from the top
      uint32_t need = s_need, b = 255;
      for (;; --b) {
        if (hist[b] >= need || b == 0) break;
        need -= hist[b];
      }
      s_need = need;                                  // still wanted from inside bin b
      s_prefix = prefix | (b << shift);
    }
    __syncthreads();
  }
  const uint32_t kth = s_prefix;                      // key of the k-th largest logit
  const uint32_t need_eq = s_need;                    // how many keys == kth belong to the top k

  // ---- passes 7, 8: collect in INDEX order (deterministic; ties at the threshold keep the lowest
  // indices): a warp owns a contiguous span, lanes stride inside it
  const int span = (V + 31) / 32;
  const int w0 = warp * span, w1 = min(V, w0 + span);
  uint32_t cgt = 0, ceq = 0;
  for (int base = w0; base < w1; base += 32) {
    const int i = base + lane;
    const uint32_t key = i < w1 ? sortable(x[i]) : 0u;
    cgt += __popc(__ballot_sync(0xffffffffu, i < w1 && key > kth));
    ceq += __popc(__ballot_sync(0xffffffffu, i < w1 && key == kth));
  }
  if (lane == 0) { warp_gt[warp] = cgt; warp_eq[warp] = ceq; }
  __syncthreads();
  uint32_t off_gt = 0, off_eq = 0, tot_gt = 0;
  for (int w = 0; w < 32; ++w) {
    if (w < warp) { off_gt += warp_gt[w]; off_eq += warp_eq[w]; }
    tot_gt += warp_gt[w];
  }
  for (int base = w0; base < w1; base += 32) {
    const int i = base + lane;
    const uint32_t key = i < w1 ? sortable(x[i]) : 0u;
    const bool gt = i < w1 && key > kth, eq = i < w1 && key == kth;
    const uint32_t bg = __ballot_sync(0xffffffffu, gt), be = __ballot_sync(0xffffffffu, eq);
    const uint32_t below = (1u << lane) - 1;
    const uint64_t item = (static_cast<uint64_t>(key) << 32) | static_cast<uint32_t>(~static_cast<uint32_t>(i));
    if (gt) cand[off_gt + __popc(bg & below)] = item;
    if (eq) {
      const uint32_t r = off_eq + __popc(be & below);
      if (r < need_eq) cand[tot_gt + r] = item;
    }
    off_gt += __popc(bg);
    off_eq += __popc(be);
  }
  const int n = static_cast<int>(tot_gt + need_eq);   // == min(k, V)
  for (int i = n + tid; i < kMaxK; i += kSampThreads) cand[i] = 0;
  __syncthreads();

  // ---- bitonic sort of the 256 slots, descending (logit desc, index asc; empty slots sink)
  for (int size = 2; size <= kMaxK; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < kMaxK / 2) {
        const int lo = (tid / stride) * 2 * stride + (tid % stride), hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const uint64_t a = cand[lo], b = cand[hi];
        if ((a < b) == desc) { cand[lo] = b; cand[hi] = a; }
      }
      __syncthreads();
    }

  // ---- top-p cut and the draw (<= 256 candidates: one thread, double cumulative sums)
  if (tid == 0) {
    const float inv_sum = static_cast<float>(1.0 / sum);
    float sum_p = 0.f;
    for (int i = 0; i < n; ++i)
      sum_p += __expf((unsortable(static_cast<uint32_t>(cand[i] >> 32)) - mx) * inv_t) * inv_sum;
    const bool cut = !(top_p <= 0.f || top_p >= sum_p);
    // kept set = prefix [0, m): the walk zeroes entry i once the mass BEFORE it has reached top_p
    int m = n;
    double total = 0.0;
    if (cut) {
      float cumsum = 0.f;
      m = 0;
      for (int i = 0; i < n; ++i) {
        if (cumsum >= top_p) break;
        const float pi = __expf((unsortable(static_cast<uint32_t>(cand[i] >> 32)) - mx) * inv_t) * inv_sum;
        cumsum += pi;
        total += static_cast<double>(pi);
        m = i + 1;
      }
    } else {
      for (int i = 0; i < n; ++i)
        total += static_cast<double>(__expf((unsortable(static_cast<uint32_t>(cand[i] >> 32)) - mx) * inv_t) * inv_sum);
    }
    const double target = static_cast<double>(row_uniform(*seed_ptr, row)) * total;
    double cum = 0.0;
    int pick = m - 1;
    for (int i = 0; i < m; ++i) {
      cum += static_cast<double>(__expf((unsortable(static_cast<uint32_t>(cand[i] >> 32)) - mx) * inv_t) * inv_sum);
      if (cum > target) { pick = i; break; }
    }
    out_ids[row] = ~static_cast<uint32_t>(cand[pick < 0 ? 0 : pick]);
  }
}

// Sampling::All (logits_processor.rs:199-207: temperature only): multinomial over the FULL softmax(logits / T).
// A thread owns a contiguous span of the vocabulary; the span sums are scanned in index order, so the draw is
// the inverse CDF in token-id order (deterministic for a given seed).
__global__ void __launch_bounds__(kSampThreads)
sample_all_kernel(uint32_t* __restrict__ out_ids, const float* __restrict__ logits, int V, float inv_t,
                  const uint64_t* __restrict__ seed_ptr) {
  const int row = blockIdx.x, tid = threadIdx.x;
  const float* x = logits + static_cast<size_t>(row) * V;
  __shared__ float red_f[32];
  __shared__ double part[kSampThreads];
  __shared__ double s_total;
  __shared__ int s_pick;
  pdl_trigger();
  pdl_wait();
  float mx = -INFINITY;
  for (int i = tid; i < V; i += kSampThreads) mx = fmaxf(mx, x[i]);
  mx = block_reduce(mx, [](float a, float b) { return fmaxf(a, b); }, red_f, -INFINITY);
  const int span = (V + kSampThreads - 1) / kSampThreads;
  const int i0 = min(V, tid * span), i1 = min(V, i0 + span);
  double local = 0.0;
  for (int i = i0; i < i1; ++i) local += static_cast<double>(__expf((x[i] - mx) * inv_t));
  part[tid] = local;
  if (tid == 0) s_pick = V - 1;
  __syncthreads();
  if (tid == 0) {                                     // 1024 adds: exclusive prefix in place
    double run = 0.0;
    for (int t = 0; t < kSampThreads; ++t) {
      const double v = part[t];
      part[t] = run;
      run += v;
    }
    s_total = run;
  }
  __syncthreads();
  const double total = s_total;
  const double target = static_cast<double>(row_uniform(*seed_ptr, row)) * total;
  const double before = part[tid];
  if (i0 < i1 && target >= before && target < before + local) {
    double cum = before;
    int pick = i1 - 1;
    for (int i = i0; i < i1; ++i) {
      cum += static_cast<double>(__expf((x[i] - mx) * inv_t));
      if (cum > target) { pick = i; break; }
    }
    s_pick = pick;                                    // exactly one span contains the target
  }
  __syncthreads();
  if (tid == 0) out_ids[row] = static_cast<uint32_t>(s_pick);
}

// logits[b, tok] -= count * frequency + (count > 0) * presence over the row's penalty window
// (<= 128 tokens, logits_processor.rs:286-306); windows shorter than 2 tokens are no-ops (:325).
__global__ void __launch_bounds__(128)
penalties_kernel(float* __restrict__ logits, int V, const uint32_t* __restrict__ context, const int32_t* __restrict__ lens,
                 int cap, float frequency, float presence) {
  const int row = blockIdx.x, i = threadIdx.x;
  __shared__ uint32_t toks[128];
  pdl_trigger();
  pdl_wait();
  const int n = min(lens[row], min(cap, 128));
  if (n <= 1) return;
  if (i < n) toks[i] = context[static_cast<size_t>(row) * cap + i];
  __syncthreads();
  if (i >= n) return;
  const uint32_t t = toks[i];
  if (t >= static_cast<uint32_t>(V)) return;
  int count = 0;
  bool first = true;
  for (int j = 0; j < n; ++j) {
    if (toks[j] == t) {
      ++count;
      if (j < i) first = false;
    }
  }
  if (first) {
    // the reference's expression, left to right, without FMA contraction (logits_processor.rs:302-304)
    float* lp = logits + static_cast<size_t>(row) * V + t;
    *lp = __fsub_rn(__fsub_rn(*lp, __fmul_rn(static_cast<float>(count), frequency)), presence);
  }
}

}  // namespace

void launch_sample_topk_topp(uint32_t* out_ids, const float* logits, int B, int V, int k, float p, float t,
                             const uint64_t* seed_dev, cudaStream_t st) {
  if (k <= 0) {                                        // Sampling::All
    launch_pdl(sample_all_kernel, dim3(B), dim3(kSampThreads), 0, st, out_ids, logits, V, 1.f / t, seed_dev);
    B200_CUDA_LAUNCH_CHECK();
    return;
  }
  launch_pdl(sample_topk_topp_kernel, dim3(B), dim3(kSampThreads), 0, st, out_ids, logits, V, k, p, 1.f / t,
             seed_dev);
  B200_CUDA_LAUNCH_CHECK();
}

}  // namespace b200

extern "C" {

void b200_sample_topk_topp_f32(uint32_t* out_ids, const float* logits, int32_t batch, int32_t vocab,
                               int32_t top_k, float top_p, float temperature, const uint64_t* seed_dev,
                               int64_t stream) {
  using namespace b200;
  if (batch <= 0) return;
  B200_CHECK_ARG(out_ids && logits && seed_dev && vocab > 0, "sample: null pointer");
  B200_CHECK_ARG(top_k >= 0 && top_k <= kMaxK, "sample: top_k must be in [1, 256], or 0 = Sampling::All (temperature only)");
  B200_CHECK_ARG(temperature > 0.f, "sample: temperature must be positive (greedy is b200_argmax_f32)");
  launch_sample_topk_topp(out_ids, logits, batch, vocab, top_k < vocab ? top_k : vocab, top_p, temperature,
                          seed_dev, reinterpret_cast<cudaStream_t>(stream));
}

void b200_apply_penalties_f32(float* logits, int32_t batch, int32_t vocab, const uint32_t* context,
                              const int32_t* context_lens, int32_t context_cap, float frequency_penalty,
                              float presence_penalty, int64_t stream) {
  using namespace b200;
  if (batch <= 0) return;
  B200_CHECK_ARG(logits && context && context_lens && context_cap > 0, "penalties: bad arguments");
  // guards of apply_batch_repeat_penalty (logits_processor.rs:325-329)
  const bool f_on = frequency_penalty != 1.0f && frequency_penalty != 0.f;
  const bool p_on = presence_penalty != 1.0f && presence_penalty != 0.f;
  if (!f_on && !p_on) return;
  launch_pdl(penalties_kernel, dim3(batch), dim3(128), 0, reinterpret_cast<cudaStream_t>(stream), logits,
             vocab, context, context_lens, context_cap, frequency_penalty, presence_penalty);
  B200_CUDA_LAUNCH_CHECK();
}

// ------------------------------------------------------------------ StdRng mirror (host)
struct b200_rng {
  uint32_t key[8];
  uint64_t counter;
  uint32_t buf[16];
  int idx;
};

static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

static void chacha12_block(b200_rng* r) {
  uint32_t st[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u};
  for (int i = 0; i < 8; ++i) st[4 + i] = r->key[i];
  st[12] = static_cast<uint32_t>(r->counter);
  st[13] = static_cast<uint32_t>(r->counter >> 32);
  st[14] = st[15] = 0;
  uint32_t x[16];
  std::memcpy(x, st, sizeof(x));
#define B200_QR(a, b, c, d)                                                  \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 16); x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 12); \
  x[a] += x[b]; x[d] = rotl32(x[d] ^ x[a], 8);  x[c] += x[d]; x[b] = rotl32(x[b] ^ x[c], 7);
  for (int round = 0; round < 6; ++round) {
    B200_QR(0, 4, 8, 12) B200_QR(1, 5, 9, 13) B200_QR(2, 6, 10, 14) B200_QR(3, 7, 11, 15)
    B200_QR(0, 5, 10, 15) B200_QR(1, 6, 11, 12) B200_QR(2, 7, 8, 13) B200_QR(3, 4, 9, 14)
  }
#undef B200_QR
  for (int i = 0; i < 16; ++i) r->buf[i] = x[i] + st[i];
  r->counter += 1;
  r->idx = 0;
}

b200_rng* b200_rng_new(uint64_t seed) {
  auto* r = new b200_rng();
  uint64_t state = seed;
  for (int i = 0; i < 8; ++i) {            // rand_core seed_from_u64: PCG32 XSH-RR
    state = state * 6364136223846793005ull + 11634580027462260723ull;
    const uint32_t xs = static_cast<uint32_t>(((state >> 18) ^ state) >> 27);
    const uint32_t rot = static_cast<uint32_t>(state >> 59);
    r->key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
  }
  r->counter = 0;
  r->idx = 16;
  return r;
}
void b200_rng_free(b200_rng* r) { delete r; }
uint32_t b200_rng_next_u32(b200_rng* r) {
  if (r->idx >= 16) chacha12_block(r);
  return r->buf[r->idx++];
}
uint64_t b200_rng_next_u64(b200_rng* r) {
  const uint64_t lo = b200_rng_next_u32(r);
  return lo | (static_cast<uint64_t>(b200_rng_next_u32(r)) << 32);
}

/* LogitsProcessor::get_strategy + the device (k, p, t) of sample_with_strategy
 * (logits_processor.rs:48-70,208-216).  has_* = 0 means None.  Returns the strategy kind:
 * 0 ArgMax, 1 All, 2 TopK, 3 TopP, 4 TopKThenTopP; k/p/t receive the kernel parameters
 * (k = 0 for kinds the device sampler does not take).                                   */
int32_t b200_sampling_strategy(int32_t has_temperature, float temperature, int32_t has_top_k, int64_t top_k,
                               int32_t has_top_p, float top_p, int32_t* k, float* p, float* t) {
  const bool temp_ok = has_temperature && !(temperature < 1e-7f) && temperature > 0.f;
  const bool k_ok = has_top_k && top_k > 0;
  *k = 0; *p = 0.f; *t = 0.f;
  if (!temp_ok) return 0;
  *t = temperature;
  if (!k_ok && !has_top_p) return 1;
  if (k_ok && !has_top_p) { *k = static_cast<int32_t>(top_k); *p = 1.0f; return 2; }
  if (!k_ok) { *k = 256; *p = top_p; return 3; }
  *k = static_cast<int32_t>(top_k); *p = top_p;
  return 4;
}

}  // extern "C"
