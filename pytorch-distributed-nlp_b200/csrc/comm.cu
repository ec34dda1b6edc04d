// Peer-HBM plumbing over NVLink 5 / NVSwitch: IPC-shared symmetric buffers and flag-based device barriers.
// One process per GPU; the host exchanges the 64-byte IPC handles through torch.distributed once at start-up,
// after which the step path never calls NCCL:
//   b2_peer_barrier          <- torch.distributed.barrier()            multi-gpu-distributed-cls.py:171,208,230
//   b2_allgather_rows        <- Trainer.output_reduce / all_gather     multi-gpu-distributed-cls.py:145-155
//   b2_scalar_allreduce_mean <- Trainer.loss_reduce  / all_reduce      multi-gpu-distributed-cls.py:139-143
// Flag protocol: rank r owns uint32 flags[B2_FLAG_SLOTS][world] in its own pad; peer q signals r by storing a
// monotonically increasing epoch into flags[slot][q] with a system-scope release; r spins (bounded) with
// system-scope acquire loads.  Epochs live in device memory so CUDA-graph replays stay in lock step.
#include "common.cuh"
#include <cstdlib>
#include <cstring>
#include "../../include/b2_ddp_bert.h"

namespace b2 {

constexpr int MAX_WORLD = 8;

struct PeerPtrs {
  void* p[MAX_WORLD];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// executed by threads 0..world-1 of ONE block; all earlier writes of this block must be fenced by the caller.
// Bounded by WALL-CLOCK time (timeout_ns, default 30 s, B2_BARRIER_TIMEOUT_S): a rank that died or diverged must
// surface as a trapped kernel (-> CUDA error -> RuntimeError on the host) within seconds, not hang the job; a peer
// that is legitimately late (rank 0 writing a checkpoint between two steps) gets that long.
__device__ __forceinline__ void barrier_signal_wait(const PeerPtrs& flags, int world, int rank, int slot,
                                                    uint32_t epoch, unsigned long long timeout_ns) {
  const int r = threadIdx.x;
  if (r < world) {
    uint32_t* remote = reinterpret_cast<uint32_t*>(flags.p[r]) + slot * world + rank;
    st_release_sys(remote, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(flags.p[rank]) + slot * world + r;
    unsigned long long t0 = 0;
    uint32_t spins = 0;
    while ((int32_t)(ld_acquire_sys(mine) - epoch) < 0) {
      if ((++spins & 0x3ffu) == 0) {          // look at the clock every 1024 polls
        const unsigned long long now = global_timer_ns();
        if (t0 == 0) t0 = now;
        if (now - t0 > timeout_ns) {
          printf("b2: peer barrier timed out after %.1f s (rank %d waiting for rank %d, slot %d, epoch %u)\n",
                 (double)(now - t0) * 1e-9, rank, r, slot, epoch);
          __trap();
        }
      }
    }
  }
}

__global__ void peer_barrier_kernel(PeerPtrs flags, int world, int rank, int slot, uint32_t* epoch_ctr,
                                    unsigned long long timeout_ns) {
  pdl_wait();               // PDL: predecessors complete + visible before any global access
  pdl_launch_dependents();  // let the next kernel in the stream begin launching
  __shared__ uint32_t epoch;
  if (threadIdx.x == 0) {
    epoch = *epoch_ctr + 1;
    *epoch_ctr = epoch;
  }
  __syncthreads();
  __threadfence_system();
  barrier_signal_wait(flags, world, rank, slot, epoch, timeout_ns);
}

// every rank stores `bytes` (multiple of 4) into slot `rank` of every peer's buffer, then barrier
__global__ void allgather_rows_kernel(const uint32_t* __restrict__ src, long long words, PeerPtrs dst,
                                      PeerPtrs flags, int world, int rank, int slot, uint32_t* epoch_ctr,
                                      unsigned long long timeout_ns) {
  pdl_wait();               // PDL: predecessors complete + visible before any global access
  pdl_launch_dependents();  // let the next kernel in the stream begin launching
  __shared__ uint32_t epoch;
  if (threadIdx.x == 0) {
    epoch = *epoch_ctr + 1;
    *epoch_ctr = epoch;
  }
  for (int r = 0; r < world; ++r) {
    uint32_t* d = reinterpret_cast<uint32_t*>(dst.p[r]) + (size_t)rank * words;
    for (long long i = threadIdx.x; i < words; i += blockDim.x) d[i] = src[i];
  }
  __threadfence_system();
  __syncthreads();
  barrier_signal_wait(flags, world, rank, slot, epoch, timeout_ns);
}

// mean of one fp32 scalar over ranks; scratch is float[2][world] on every rank (double-buffered by epoch parity)
__global__ void scalar_allreduce_mean_kernel(const float* __restrict__ src, float* __restrict__ dst, PeerPtrs scratch,
                                             PeerPtrs flags, int world, int rank, int slot, uint32_t* epoch_ctr,
                                      unsigned long long timeout_ns) {
  pdl_wait();               // PDL: predecessors complete + visible before any global access
  pdl_launch_dependents();  // let the next kernel in the stream begin launching
  __shared__ uint32_t epoch;
  if (threadIdx.x == 0) {
    epoch = *epoch_ctr + 1;
    *epoch_ctr = epoch;
  }
  __syncthreads();
  const int par = epoch & 1;
  if ((int)threadIdx.x < world)
    reinterpret_cast<float*>(scratch.p[threadIdx.x])[par * world + rank] = *src;
  __threadfence_system();
  __syncthreads();
  barrier_signal_wait(flags, world, rank, slot, epoch, timeout_ns);
  __syncthreads();
  if (threadIdx.x == 0) {
    const volatile float* mine = reinterpret_cast<const volatile float*>(scratch.p[rank]) + par * world;
    float s = 0.f;
    for (int r = 0; r < world; ++r) s += mine[r];  // fixed order: identical result on every rank
    *dst = s / (float)world;
  }
}

static unsigned long long barrier_timeout_ns() {
  static unsigned long long v = 0;
  if (v == 0) {
    const char* e = getenv("B2_BARRIER_TIMEOUT_S");
    double sec = e != nullptr ? atof(e) : 30.0;
    if (!(sec > 0.0)) sec = 30.0;
    v = (unsigned long long)(sec * 1e9);
  }
  return v;
}

static int32_t fill_peers(PeerPtrs* out, void* const* in, int world, const char* what) {
  for (int r = 0; r < MAX_WORLD; ++r) out->p[r] = nullptr;
  for (int r = 0; r < world; ++r) {
    if (in[r] == nullptr) {
      set_error("%s: null peer pointer for rank %d", what, r);
      return -2;
    }
    out->p[r] = in[r];
  }
  return 0;
}

}  // namespace b2

using namespace b2;

extern "C" int32_t b2_comm_alloc(int64_t bytes, void** ptr) {
  B2_REQUIRE(ptr && bytes > 0, "comm_alloc: bad args");
  B2_CUDA(cudaMalloc(ptr, (size_t)bytes));
  B2_CUDA(cudaMemset(*ptr, 0, (size_t)bytes));
  return 0;
}
extern "C" int32_t b2_comm_free(void* ptr) {
  if (ptr) B2_CUDA(cudaFree(ptr));
  return 0;
}
extern "C" int32_t b2_comm_export(void* ptr, uint8_t handle[B2_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == B2_IPC_HANDLE_BYTES, "IPC handle size");
  B2_REQUIRE(ptr && handle, "comm_export: bad args");
  cudaIpcMemHandle_t h;
  B2_CUDA(cudaIpcGetMemHandle(&h, ptr));
  memcpy(handle, &h, sizeof(h));
  return 0;
}
extern "C" int32_t b2_comm_import(const uint8_t handle[B2_IPC_HANDLE_BYTES], void** ptr) {
  B2_REQUIRE(ptr && handle, "comm_import: bad args");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  B2_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int32_t b2_comm_unimport(void* ptr) {
  if (ptr) B2_CUDA(cudaIpcCloseMemHandle(ptr));
  return 0;
}

extern "C" int32_t b2_peer_barrier(void* const* peer_flags, int32_t world, int32_t rank, int32_t slot, uint32_t* epoch,
                                   void* stream_) {
  B2_REQUIRE(peer_flags && epoch, "peer_barrier: null pointer");
  B2_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "peer_barrier: world=%d rank=%d", world,
             rank);
  B2_REQUIRE(slot >= 0 && slot < B2_FLAG_SLOTS, "peer_barrier: slot=%d", slot);
  PeerPtrs f;
  int32_t st = fill_peers(&f, peer_flags, world, "peer_barrier");
  if (st) return st;
  B2_LAUNCH(peer_barrier_kernel, 1, 32, 0, (cudaStream_t)stream_, f, world, rank, slot, epoch,
            barrier_timeout_ns());
  B2_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

extern "C" int32_t b2_allgather_rows(const void* src, int64_t bytes, void* const* peer_dst, void* const* peer_flags,
                                     int32_t world, int32_t rank, int32_t slot, uint32_t* epoch, void* stream_) {
  B2_REQUIRE(src && peer_dst && peer_flags && epoch, "allgather_rows: null pointer");
  B2_REQUIRE(bytes > 0 && bytes % 4 == 0, "allgather_rows: bytes=%lld must be a positive multiple of 4",
             (long long)bytes);
  B2_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "allgather_rows: world=%d rank=%d", world,
             rank);
  B2_REQUIRE(slot >= 0 && slot < B2_FLAG_SLOTS, "allgather_rows: slot=%d", slot);
  PeerPtrs d, f;
  int32_t st = fill_peers(&d, peer_dst, world, "allgather_rows(dst)");
  if (st) return st;
  st = fill_peers(&f, peer_flags, world, "allgather_rows(flags)");
  if (st) return st;
  B2_LAUNCH(allgather_rows_kernel, 1, 256, 0, (cudaStream_t)stream_, (const uint32_t*)src, bytes / 4, d, f, world, rank, slot,
                                                              epoch, barrier_timeout_ns());
  B2_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}

extern "C" int32_t b2_scalar_allreduce_mean(const float* src, float* dst, float* const* peer_scratch,
                                            void* const* peer_flags, int32_t world, int32_t rank, int32_t slot,
                                            uint32_t* epoch, void* stream_) {
  B2_REQUIRE(src && dst && peer_scratch && peer_flags && epoch, "scalar_allreduce_mean: null pointer");
  B2_REQUIRE(world >= 1 && world <= MAX_WORLD && rank >= 0 && rank < world, "scalar_allreduce_mean: world=%d rank=%d",
             world, rank);
  B2_REQUIRE(slot >= 0 && slot < B2_FLAG_SLOTS, "scalar_allreduce_mean: slot=%d", slot);
  PeerPtrs s, f;
  int32_t st = fill_peers(&s, (void* const*)peer_scratch, world, "scalar_allreduce_mean(scratch)");
  if (st) return st;
  st = fill_peers(&f, peer_flags, world, "scalar_allreduce_mean(flags)");
  if (st) return st;
  B2_LAUNCH(scalar_allreduce_mean_kernel, 1, 32, 0, (cudaStream_t)stream_, src, dst, s, f, world, rank, slot, epoch,
            barrier_timeout_ns());
  B2_CUDA(cudaGetLastError());
  count_launches(1);
  return 0;
}
