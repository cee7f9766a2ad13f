// Tensor-parallel communicator: one process per GPU, one NCCL communicator, one collective
// (all-reduce SUM of the row-parallel GEMM output) -- exactly the reference's data plane:
//   Comm::from_rank(device, rank, num_shards, nccl_id)   /root/reference/src/runner/runner.rs:81-89
//   Id shipped as 128 raw bytes                           /root/reference/src/runner/mod.rs:33-58
//   AllReduce::cuda_fwd -> comm.all_reduce(SUM)           /root/reference/src/models/layers/distributed.rs:335-395
// NCCL is resolved with dlopen("libnccl.so.2") so single-GPU users need no NCCL at all.
//
// On top of the NCCL baseline there is a one-shot all-reduce for decode-sized messages
// (<= 1 MiB): every rank's partial sits in a peer-mapped buffer (cudaIpc over NVLink), a rank
// signals its peers with a release store, waits for theirs with acquire loads, then sums all
// partials in rank order straight from peer memory (775 GB/s peer loads, ~2 us latency) -- one
// kernel, CUDA-graph capturable, bit-identical result on every rank.
#include <dlfcn.h>

#include <cstring>
#include <vector>

#include "../../include/b200_vllm.h"
#include "common.cuh"
#include "wna16.cuh"

namespace b200 {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat16 = 6, ncclUint8 = 1, ncclInt32 = 2, ncclSum = 0, ncclMin = 3 };

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  bool ok = false;
};

static NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
      api.AllReduce = (decltype(api.AllReduce))dlsym(h, "ncclAllReduce");
      api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllReduce &&
               api.AllGather;
    }
  }
  return api;
}

constexpr int kMaxRanks = 8;
constexpr size_t kOneShotMaxBytes = 1 << 20;          // per-rank payload limit of the one-shot path
constexpr int kSlots = 2;                             // double-buffered payload slots

struct OneShotPeers {
  uint8_t* data[kMaxRanks];       // peer payload buffers  [kSlots][kOneShotMaxBytes]
  uint32_t* flags[kMaxRanks];     // peer flag arrays      [kSlots][kMaxRanks]
};

// Device-resident state so that a captured launch replays correctly: `state[0]` = number of
// completed one-shot all-reduces (sequence number), state[1] / state[2] = intra-grid arrival
// counters (self-resetting).  Launched as a single small wave (the kernel spins on peers).
__global__ void __launch_bounds__(512)
oneshot_allreduce_kernel(OneShotPeers peers, __half* __restrict__ buf, size_t count, int rank,
                         int world, uint32_t* __restrict__ state) {
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(&state[0]) + 1;   // same in every CTA
  const int slot = seq & 1;
  const size_t n8 = count / 8;                       // 16-byte vectors
  uint4* mine = reinterpret_cast<uint4*>(peers.data[rank] + static_cast<size_t>(slot) * kOneShotMaxBytes);
  const uint4* src = reinterpret_cast<const uint4*>(buf);
  const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
  // 1. publish my partial into my own symmetric buffer
  for (size_t i = tid; i < n8; i += nthreads) mine[i] = src[i];
  __threadfence_system();
  __syncthreads();
  // 2. intra-grid arrival; the last CTA signals every peer (release, system scope)
  __shared__ bool last;
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(&state[1], 1u);
    last = (prev == gridDim.x - 1);
    if (last) state[1] = 0;
  }
  __syncthreads();
  if (last && threadIdx.x < world) {
    uint32_t* f = peers.flags[threadIdx.x] + slot * kMaxRanks + rank;
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
  }
  // 3. wait until every peer's partial for this sequence number has landed
  if (threadIdx.x < world) {
    const uint32_t* f = peers.flags[rank] + slot * kMaxRanks + threadIdx.x;
    uint32_t v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (clock64() - t0 > 20000000000LL) __trap();      // ~10 s: a peer died or ran a different step count
    } while (static_cast<int32_t>(v - seq) < 0);
  }
  __syncthreads();
  // 4. sum all partials in rank order (same order on every rank => identical results)
  uint4* dst = reinterpret_cast<uint4*>(buf);
  for (size_t i = tid; i < n8; i += nthreads) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < world; ++r) {
      const uint4 v = __ldcv(reinterpret_cast<const uint4*>(
                                 peers.data[r] + static_cast<size_t>(slot) * kOneShotMaxBytes) + i);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    uint4 o;
    __half2* oh = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    dst[i] = o;
  }
  // 5. the last CTA to finish advances the sequence number for the next launch / replay
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(&state[2], 1u);
    if (prev == gridDim.x - 1) {
      state[2] = 0;
      *reinterpret_cast<volatile uint32_t*>(&state[0]) = seq;
    }
  }
}

// ---------------------------------------------------------------- fused GEMM -> all-reduce -> add + RMSNorm
// Receive side of the kEpiPushPeers GEMM epilogue (wna16.cuh): every rank's row-parallel GEMM stores
// its fp16 partial tiles straight into recv[buf][src] on ALL ranks while it is still computing other
// tiles, and bumps cnt[buf][src] once per finished tile.  This kernel -- the only one between the
// GEMM and the next layer's GEMM -- waits until every source has delivered `tiles` tiles of call
// number `seq`, sums the partials in rank order (fp32, rounded once to fp16: the all-reduce result of
// distributed.rs:367-369), adds the residual in fp16 (llama.rs:126-130) and applies the next RMSNorm.
// Replaces three launches (copy + one-shot all-reduce, add_rmsnorm) and one NVLink round trip.
// Two buffers alternate between the o_proj and down_proj call sites: a rank can only push into
// buffer A again after it passed the wait on buffer B, which every peer feeds only after it finished
// reading A -- so no slot is overwritten while it is being read.
constexpr size_t kPushSlotBytes = 1 << 20;            // [M <= 64][hidden <= 8192] fp16
constexpr int kPushBufs = 2;

struct PushState {
  uint32_t seq[kPushBufs];        // completed calls per buffer
  uint32_t done[kPushBufs];       // intra-grid arrival counters (self-resetting)
};

__global__ void __launch_bounds__(1024)
tp_reduce_add_rmsnorm_kernel(const __half* __restrict__ recv, const uint32_t* __restrict__ cnt,
                             PushState* __restrict__ state, int buf, int world, uint32_t tiles,
                             __half* __restrict__ resid, const __half* __restrict__ weight,
                             __half* __restrict__ out, int hidden, float eps) {
  const int m = blockIdx.x, tid = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(&state->seq[buf]) + 1;
  if (tid < world) {
    const uint32_t want = seq * tiles;
    const uint32_t* c = cnt + buf * kMaxRanks + tid;
    uint32_t v;
    const long long t0 = clock64();
    do {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(c) : "memory");
      if (clock64() - t0 > 20000000000LL) __trap();      // ~10 s: a peer died
    } while (static_cast<int32_t>(v - want) < 0);
  }
  __syncthreads();
  __shared__ float red[32];
  const int n8 = hidden / 8;
  float x[8];
  float sq = 0.f;
  const bool active = tid < n8;
  if (active) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int r = 0; r < world; ++r) {
      const uint4 v = __ldcv(reinterpret_cast<const uint4*>(
                                 recv + (static_cast<size_t>(buf) * kMaxRanks + r) * (kPushSlotBytes / 2) +
                                 static_cast<size_t>(m) * hidden) + tid);
      const __half2* h = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h[j]);
        acc[2 * j] += f.x;
        acc[2 * j + 1] += f.y;
      }
    }
    uint4 rv = *(reinterpret_cast<const uint4*>(resid + static_cast<size_t>(m) * hidden) + tid);
    __half2* rh = reinterpret_cast<__half2*>(&rv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(rh[j]);
      // all-reduce result rounded to fp16, then the fp16 residual add
      const float a0 = __half2float(__float2half_rn(acc[2 * j])), a1 = __half2float(__float2half_rn(acc[2 * j + 1]));
      x[2 * j] = __half2float(__float2half_rn(f.x + a0));
      x[2 * j + 1] = __half2float(__float2half_rn(f.y + a1));
      rh[j] = __floats2half2_rn(x[2 * j], x[2 * j + 1]);
      sq += x[2 * j] * x[2 * j] + x[2 * j + 1] * x[2 * j + 1];
    }
    *(reinterpret_cast<uint4*>(resid + static_cast<size_t>(m) * hidden) + tid) = rv;
  }
  sq = warp_sum(sq);
  if ((tid & 31) == 0) red[tid >> 5] = sq;
  __syncthreads();
  if (tid < 32) {
    float t = tid < (blockDim.x >> 5) ? red[tid] : 0.f;
    t = warp_sum(t);
    if (tid == 0) red[0] = t;
  }
  __syncthreads();
  const float rs = rsqrtf(red[0] / static_cast<float>(hidden) + eps);
  if (active) {
    const uint4 wv = *(reinterpret_cast<const uint4*>(weight) + tid);
    const __half2* wh = reinterpret_cast<const __half2*>(&wv);
    uint4 ov;
    __half2* oh = reinterpret_cast<__half2*>(&ov);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 w = __half22float2(wh[j]);
      oh[j] = __floats2half2_rn(x[2 * j] * rs * w.x, x[2 * j + 1] * rs * w.y);
    }
    *(reinterpret_cast<uint4*>(out + static_cast<size_t>(m) * hidden) + tid) = ov;
  }
  // the last row to finish advances the call number for the next launch / graph replay
  __syncthreads();
  if (tid == 0) {
    const uint32_t prev = atomicAdd(&state->done[buf], 1u);
    if (prev == gridDim.x - 1) {
      state->done[buf] = 0;
      *reinterpret_cast<volatile uint32_t*>(&state->seq[buf]) = seq;
    }
  }
}

}  // namespace b200

using namespace b200;

struct b200_comm {
  ncclComm_t nccl = nullptr;
  int rank = 0, world = 1;
  // one-shot state
  bool oneshot = false;
  uint8_t* my_data = nullptr;
  uint32_t* my_flags = nullptr;
  uint32_t* state = nullptr;
  OneShotPeers peers;
  std::vector<void*> opened;
  int force_nccl = 0;
  // fused push path (kEpiPushPeers): this rank's receive area and every peer's view of it
  size_t push_off = 0, cnt_off = 0, pstate_off = 0;
};

extern "C" {

int32_t b200_nccl_unique_id(uint8_t* out128) {
  NcclApi& a = nccl();
  if (!a.ok) {
    set_error(2, "comm: libnccl.so.2 not found");
    return -1;
  }
  ncclUniqueId id;
  if (a.GetUniqueId(&id) != 0) {
    set_error(2, "comm: ncclGetUniqueId failed");
    return -1;
  }
  std::memcpy(out128, id.internal, 128);
  return 0;
}

b200_comm* b200_comm_init_rank(const uint8_t* id128, int32_t rank, int32_t world) {
  NcclApi& a = nccl();
  if (!a.ok || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) {
    set_error(1, "comm_init_rank: NCCL unavailable or bad rank/world");
    return nullptr;
  }
  auto* c = new b200_comm();
  c->rank = rank;
  c->world = world;
  ncclUniqueId id;
  std::memcpy(id.internal, id128, 128);
  const ncclResult_t r = a.CommInitRank(&c->nccl, world, id, rank);
  if (r != 0) {
    set_error(2, a.GetErrorString ? a.GetErrorString(r) : "ncclCommInitRank failed");
    delete c;
    return nullptr;
  }
  return c;
}

/* Enable the one-shot path: allocate the symmetric buffer, exchange cudaIpc handles with
 * ncclAllGather, map every peer.  Every rank runs the SAME collectives whatever happened locally and the
 * outcome is agreed with an all-reduce(MIN) of the local ok flags: the path is on for all ranks or for
 * none (a rank that silently fell back to NCCL would leave its peers spinning on flags that never come).
 * Returns 0 on success; on failure everything is released and the communicator keeps working through
 * NCCL only. */
int32_t b200_comm_enable_oneshot(b200_comm* c, int64_t stream) {
  if (!c || c->world < 2) return -1;
  NcclApi& a = nccl();
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t data_bytes = kSlots * kOneShotMaxBytes;
  const size_t flag_bytes = kSlots * kMaxRanks * sizeof(uint32_t);
  uint8_t* base = nullptr;
  const size_t push_off = data_bytes + flag_bytes + 256;
  const size_t push_bytes = static_cast<size_t>(kPushBufs) * kMaxRanks * kPushSlotBytes;
  const size_t cnt_off = push_off + push_bytes;
  const size_t pstate_off = cnt_off + 256;
  const size_t total = pstate_off + 256;
  int ok = 1;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&base, total) != cudaSuccess) {
    cudaGetLastError();
    base = nullptr;
    ok = 0;
  } else {
    cudaMemset(base, 0, total);
    if (cudaIpcGetMemHandle(&mine, base) != cudaSuccess) {
      cudaGetLastError();
      ok = 0;
    }
  }
  // scratch for the two collectives: [world handles][1 int flag]
  uint8_t* d_x = nullptr;
  const size_t hbytes = sizeof(cudaIpcMemHandle_t) * static_cast<size_t>(c->world);
  if (cudaMalloc(&d_x, hbytes + 16) != cudaSuccess) {
    // without even this scratch the rank cannot take part in the agreement: a hard error, not a fallback
    cudaGetLastError();
    if (base) cudaFree(base);
    set_error(2, "comm_enable_oneshot: cudaMalloc of the exchange scratch failed");
    return -1;
  }
  cudaIpcMemHandle_t* d_all = reinterpret_cast<cudaIpcMemHandle_t*>(d_x);
  int* d_ok = reinterpret_cast<int*>(d_x + hbytes);
  cudaMemcpy(d_all + c->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice);
  bool coll_ok = a.AllGather(d_all + c->rank, d_all, sizeof(mine), ncclUint8, c->nccl, st) == 0;
  cudaStreamSynchronize(st);
  std::vector<cudaIpcMemHandle_t> all(c->world);
  cudaMemcpy(all.data(), d_all, hbytes, cudaMemcpyDeviceToHost);
  OneShotPeers peers;
  std::memset(&peers, 0, sizeof(peers));
  std::vector<void*> opened;
  if (ok && coll_ok) {
    for (int r = 0; r < c->world && ok; ++r) {
      uint8_t* p = base;
      if (r != c->rank) {
        void* q = nullptr;
        if (cudaIpcOpenMemHandle(&q, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          ok = 0;
          break;
        }
        opened.push_back(q);
        p = reinterpret_cast<uint8_t*>(q);
      }
      peers.data[r] = p;
      peers.flags[r] = reinterpret_cast<uint32_t*>(p + data_bytes);
    }
  }
  // agreement: enabled only if EVERY rank mapped every peer
  int all_ok = ok && coll_ok ? 1 : 0;
  cudaMemcpy(d_ok, &all_ok, sizeof(int), cudaMemcpyHostToDevice);
  if (a.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, c->nccl, st) != 0) all_ok = 0;
  cudaStreamSynchronize(st);
  int agreed = 0;
  cudaMemcpy(&agreed, d_ok, sizeof(int), cudaMemcpyDeviceToHost);
  cudaFree(d_x);
  if (!all_ok || !agreed) {
    for (void* q : opened) cudaIpcCloseMemHandle(q);
    if (base) cudaFree(base);
    c->oneshot = false;
    c->my_data = nullptr;
    c->my_flags = nullptr;
    c->state = nullptr;
    return -1;
  }
  c->push_off = push_off;
  c->cnt_off = cnt_off;
  c->pstate_off = pstate_off;
  c->my_data = base;
  c->my_flags = reinterpret_cast<uint32_t*>(base + data_bytes);
  c->state = reinterpret_cast<uint32_t*>(base + data_bytes + flag_bytes);
  c->peers = peers;
  c->opened = opened;
  c->oneshot = true;
  return 0;
}

void b200_comm_set_force_nccl(b200_comm* c, int32_t force) {
  if (c) c->force_nccl = force;
}

void b200_comm_free(b200_comm* c) {
  if (!c) return;
  for (void* p : c->opened) cudaIpcCloseMemHandle(p);
  if (c->my_data) cudaFree(c->my_data);
  if (c->nccl) nccl().CommDestroy(c->nccl);
  delete c;
}

/* In-place all-reduce SUM of `count` fp16 values (distributed.rs:367-369, ReduceOp::Sum). */
void b200_allreduce_f16(b200_comm* c, void* buf, int64_t count, int64_t stream) {
  if (!c || c->world == 1 || count <= 0) return;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t bytes = static_cast<size_t>(count) * 2;
  if (c->oneshot && !c->force_nccl && bytes <= kOneShotMaxBytes && count % 8 == 0) {
    const size_t n8 = count / 8;
    int blocks = static_cast<int>((n8 + 511) / 512);
    if (blocks > 32) blocks = 32;      // one wave by a wide margin: the kernel spins on peers
    if (blocks < 1) blocks = 1;
    oneshot_allreduce_kernel<<<blocks, 512, 0, st>>>(c->peers, reinterpret_cast<__half*>(buf),
                                                    static_cast<size_t>(count), c->rank, c->world,
                                                    c->state);
    B200_CUDA_LAUNCH_CHECK();
    return;
  }
  const ncclResult_t r = nccl().AllReduce(buf, buf, static_cast<size_t>(count), ncclFloat16, ncclSum,
                                          c->nccl, st);
  count_launch(1);
  if (r != 0) set_error(2, "allreduce: ncclAllReduce failed");
}

int32_t b200_comm_rank(const b200_comm* c) { return c ? c->rank : 0; }
int32_t b200_comm_world(const b200_comm* c) { return c ? c->world : 1; }

}  // extern "C"

// ---------------------------------------------------------------- internal (model_runner.cpp)
namespace b200 {

// Fill the kEpiPushPeers fields of a GEMM epilogue for call-site buffer `buf` (0 = o_proj, 1 = down_proj).
bool comm_push_targets(b200_comm* c, int buf, int M, int N, FusedEpi* fe) {
  if (!c || !c->oneshot || c->force_nccl || buf < 0 || buf >= kPushBufs) return false;
  if (static_cast<size_t>(M) * N * 2 > kPushSlotBytes || N % 8 != 0) return false;
  for (int r = 0; r < c->world; ++r) {
    uint8_t* base = c->peers.data[r];
    fe->push[r] = reinterpret_cast<__half*>(base + c->push_off +
                                            (static_cast<size_t>(buf) * kMaxRanks + c->rank) * kPushSlotBytes);
    fe->push_cnt[r] = reinterpret_cast<uint32_t*>(base + c->cnt_off) + buf * kMaxRanks + c->rank;
  }
  fe->push_world = c->world;
  return true;
}

// x += all-reduce(pushed partials); out = rmsnorm(x) * weight.  `tiles` = 128-channel tiles every rank pushes.
void comm_reduce_add_rmsnorm(b200_comm* c, int buf, int tiles, void* resid, const void* weight, void* out,
                             int M, int hidden, float eps, cudaStream_t st) {
  int threads = ((hidden / 8 + 31) / 32) * 32;
  if (threads < 32) threads = 32;
  launch_pdl(tp_reduce_add_rmsnorm_kernel, dim3(M), dim3(threads), 0, st,
             reinterpret_cast<const __half*>(c->my_data + c->push_off),
             reinterpret_cast<const uint32_t*>(c->my_data + c->cnt_off),
             reinterpret_cast<PushState*>(c->my_data + c->pstate_off), buf, c->world,
             static_cast<uint32_t>(tiles), reinterpret_cast<__half*>(resid),
             reinterpret_cast<const __half*>(weight), reinterpret_cast<__half*>(out), hidden, eps);
  B200_CUDA_LAUNCH_CHECK();
}

}  // namespace b200
