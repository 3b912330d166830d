// One-shot all-reduce of the predictive-moment buffers over NVLink / NVSwitch peer memory.
//
// The only exchange on the path is the sum over ranks of the [B,C] sample-sums (posteriormodel.py:94-101 sharded over
// posterior samples).  At the per-rank step times of the strong split (0.6 ms at 8 GPUs) the fixed cost of a library
// collective issued from Python matters, so the exchange is ONE kernel on the engine's stream:
//
//   every rank owns a "window" in its own HBM (cudaMalloc + cudaIpc handle, mapped by every peer):
//       slot[2][cap] floats   (double-buffered by epoch parity)     flags[world]   (written by the peers)
//   phase 0   copy the local sums into slot[epoch & 1] of the OWN window (plain stores, 16 bytes per thread)
//   signal    the last block to finish phase 0 (device-scope counter) publishes `epoch` into flag[rank] of every peer
//             (fence.sys + st.release.sys through the NVLink mapping)
//   wait      every block polls its own flags until all `world` of them show >= epoch (ld.acquire.sys)
//   phase 1   every rank reads slot[epoch & 1] of ALL windows in rank order 0..world-1 and writes the sum in place:
//             the same order on every rank, so all ranks hold bit-identical results and the result does not depend on
//             arrival order (deterministic, no float atomics)
//
// Slot reuse is safe without a second barrier: a rank can only be writing slot[e & 1] again at epoch e+2 after passing the
// barrier of epoch e+1, which every peer signals from its kernel e+1, i.e. after its kernel e (the reader of slot[e & 1])
// has finished -- calls on one communicator are stream-ordered on every rank.  A peer that never arrives trips a
// 10 s timeout (globaltimer): the kernel raises a sticky status word in mapped host memory instead of hanging the GPU,
// and the next call on the communicator fails.
//
// The epoch lives in device memory and is advanced by the kernel itself, so the launch takes no per-call host
// arguments besides the pointers (it can sit inside a CUDA graph).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <vector>
#include "common.cuh"
#include "../../include/deepbayes_b200.h"

namespace dbx {

constexpr int kCommMaxWorld = 16;
constexpr int kCommBlocks = 64;
constexpr int kCommThreads = 512;

struct CommCtl {                       // lives at the start of every window
  unsigned long long flags[kCommMaxWorld];   // flags[r] = last epoch rank r has published to THIS rank
  unsigned long long epoch;                  // the epoch of the NEXT call (local)
  unsigned int arrived;                      // blocks of the running kernel that finished phase 0 (local)
  unsigned int departed;                     // blocks of the running kernel that finished phase 1 (local)
  unsigned int pad[2];
};
constexpr size_t kCtlBytes = 256;
static_assert(sizeof(CommCtl) <= kCtlBytes, "control block");

struct PeerTable { float* slots[kCommMaxWorld]; CommCtl* ctl[kCommMaxWorld]; };

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float4 ld_volatile_f4(const float* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// a | b (na + nb floats, both multiples of 4, 16-byte aligned) summed over the ranks, in place
__global__ void __launch_bounds__(kCommThreads, 1)
peer_allreduce_kernel(PeerTable pt, int rank, int world, long long cap, float* __restrict__ a, long long na, float* __restrict__ b,
                      long long nb, int* __restrict__ status) {
  CommCtl* me = pt.ctl[rank];
  const unsigned long long epoch = *((volatile unsigned long long*)&me->epoch);
  float* myslot = pt.slots[rank] + (epoch & 1ull) * cap;
  const long long n4 = (na + nb) >> 2, na4 = na >> 2;
  const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x, nth = (long long)gridDim.x * blockDim.x;
  // ---- phase 0: local sums -> own window
  for (long long i = tid; i < n4; i += nth) {
    const float4 v = i < na4 ? reinterpret_cast<const float4*>(a)[i] : reinterpret_cast<const float4*>(b)[i - na4];
    reinterpret_cast<float4*>(myslot)[i] = v;
  }
  __syncthreads();
  __shared__ int s_ok;
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int prev = atomicAdd(&me->arrived, 1u);
    if (prev == gridDim.x - 1) {          // the whole slot is written: publish the epoch to every peer (and to myself)
      __threadfence_system();
      for (int r = 0; r < world; ++r) st_release_sys(&pt.ctl[r]->flags[rank], epoch);
    }
    // ---- wait for every rank's slot of this epoch
    int ok = 1;
    const unsigned long long t0 = globaltimer_ns();
    for (int r = 0; r < world && ok; ++r) {
      unsigned int spins = 0;
      while (ld_acquire_sys(&me->flags[r]) < epoch) {
        if ((++spins & 1023u) == 0 && globaltimer_ns() - t0 > 10000000000ull) { ok = 0; break; }
        __nanosleep(64);
      }
    }
    if (!ok) *status = 1;
    s_ok = ok;
  }
  __syncthreads();
  // ---- phase 1: sum the slots of all ranks in rank order
  if (s_ok) {
    for (long long i = tid; i < n4; i += nth) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int r = 0; r < world; ++r) {
        const float4 v = ld_volatile_f4(pt.slots[r] + (epoch & 1ull) * cap + 4 * i);
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      if (i < na4) reinterpret_cast<float4*>(a)[i] = acc; else reinterpret_cast<float4*>(b)[i - na4] = acc;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(&me->departed, 1u);
    if (prev == gridDim.x - 1) {          // last block out: arm the next call
      me->arrived = 0; me->departed = 0;
      __threadfence();
      *((volatile unsigned long long*)&me->epoch) = epoch + 1;
    }
  }
}

}  // namespace dbx

using namespace dbx;

struct dbx_comm_s {
  int rank = 0, world = 1, device = 0;
  long long cap = 0;                 // floats per slot
  void* window = nullptr;            // own window: CommCtl | slot[2][cap]
  void* peer_base[kCommMaxWorld] = {};
  PeerTable pt = {};
  int* status_host = nullptr; int* status_dev = nullptr;
  bool connected = false;
  long long calls = 0;
};

extern "C" {

#define DBX_API __attribute__((visibility("default")))

DBX_API int dbx_comm_create(int32_t rank, int32_t world, int32_t device, int64_t max_floats, dbx_comm* out) {
  DBX_CHECK(out, "null argument");
  DBX_CHECK(world >= 1 && world <= kCommMaxWorld && rank >= 0 && rank < world, "dbx_comm_create: rank %d of %d (at most %d ranks)", rank, world, kCommMaxWorld);
  DBX_CHECK(max_floats > 0, "dbx_comm_create: max_floats must be positive");
  DBX_CUDA(cudaSetDevice(device));
  dbx_comm_s* c = new dbx_comm_s();
  c->rank = rank; c->world = world; c->device = device;
  c->cap = round_up(max_floats, 1024);
  const size_t bytes = kCtlBytes + 2 * (size_t)c->cap * 4;
  if (cudaMalloc(&c->window, bytes) != cudaSuccess) { delete c; return set_err("dbx_comm_create: cudaMalloc of the %zu-byte window failed", bytes); }
  cudaMemset(c->window, 0, bytes);
  CommCtl init = {}; init.epoch = 1;
  cudaMemcpy(c->window, &init, sizeof init, cudaMemcpyHostToDevice);
  if (cudaHostAlloc((void**)&c->status_host, sizeof(int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&c->status_dev, c->status_host, 0) != cudaSuccess) {
    cudaFree(c->window); delete c; return set_err("dbx_comm_create: mapped status word failed");
  }
  *c->status_host = 0;
  DBX_CUDA(cudaDeviceSynchronize());
  *out = c;
  return 0;
}

DBX_API int dbx_comm_handle(dbx_comm c, void* handle64) {
  DBX_CHECK(c && handle64, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  DBX_CUDA(cudaSetDevice(c->device));
  cudaIpcMemHandle_t h;
  DBX_CUDA(cudaIpcGetMemHandle(&h, c->window));
  memcpy(handle64, &h, 64);
  return 0;
}

DBX_API int dbx_comm_connect(dbx_comm c, const void* handles) {
  DBX_CHECK(c && handles, "null argument");
  DBX_CHECK(!c->connected, "dbx_comm_connect: already connected");
  DBX_CUDA(cudaSetDevice(c->device));
  for (int r = 0; r < c->world; ++r) {
    void* base = c->window;
    if (r != c->rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)handles + 64 * (size_t)r, 64);
      const cudaError_t e = cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        for (int q = 0; q < r; ++q) if (q != c->rank && c->peer_base[q]) { cudaIpcCloseMemHandle(c->peer_base[q]); c->peer_base[q] = nullptr; }
        return set_err("dbx_comm_connect: cudaIpcOpenMemHandle(rank %d) -> %s", r, cudaGetErrorString(e));
      }
      c->peer_base[r] = base;
    }
    c->pt.ctl[r] = (CommCtl*)base;
    c->pt.slots[r] = (float*)((char*)base + kCtlBytes);
  }
  c->connected = true;
  return 0;
}

DBX_API int dbx_comm_allreduce(dbx_comm c, float* a_dev, int64_t na, float* b_dev, int64_t nb, void* stream) {
  DBX_CHECK(c && a_dev && na > 0, "null argument");
  DBX_CHECK(c->connected, "dbx_comm_allreduce: communicator not connected (dbx_comm_connect)");
  DBX_CHECK(*c->status_host == 0, "dbx_comm_allreduce: a previous exchange timed out waiting for a peer rank");
  if (!b_dev) nb = 0;
  DBX_CHECK(na % 4 == 0 && nb % 4 == 0 && ((uintptr_t)a_dev & 15) == 0 && ((uintptr_t)b_dev & 15) == 0,
            "dbx_comm_allreduce: buffers must be 16-byte aligned with a multiple of 4 elements");
  DBX_CHECK(na + nb <= c->cap, "dbx_comm_allreduce: %lld floats exceed the window of %lld", (long long)(na + nb), c->cap);
  DBX_CUDA(cudaSetDevice(c->device));
  const long long n4 = (na + nb) / 4;
  const int grid = (int)std::max<long long>(1, std::min<long long>(kCommBlocks, cdiv(n4, kCommThreads)));
  DBX_LAUNCH(peer_allreduce_kernel, grid, kCommThreads, 0, (cudaStream_t)stream, c->pt, c->rank, c->world, c->cap, a_dev, (long long)na,
             b_dev, (long long)nb, c->status_dev);
  ++c->calls;
  return 0;
}

DBX_API int dbx_comm_status(dbx_comm c, int32_t* timed_out, int64_t* calls) {
  DBX_CHECK(c, "null argument");
  if (timed_out) *timed_out = *c->status_host;
  if (calls) *calls = c->calls;
  return 0;
}

DBX_API int dbx_comm_destroy(dbx_comm c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int r = 0; r < c->world; ++r)
    if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->window) cudaFree(c->window);
  if (c->status_host) cudaFreeHost(c->status_host);
  delete c;
  return 0;
}

}  // extern "C"
