// Tensor-level symmetric-memory collectives for the DTensor redistribute / TP / vocab-parallel paths (sm_100a).
//
// SURVEY §2F rows served here (the FSDP unit collectives are in symm_comm.cu, the GEMM-fused ones in gemm_fused_tp.cu):
//   C11/C17/C19  all_reduce      : NVLS two-shot (multimem.ld_reduce + multimem.st), P2P two-shot, or one-shot for small messages
//   C5/C10/C18   reduce_scatter  : NVLS (one multimem.ld_reduce per vector) or P2P pull of my slice, strided private output
//   C12          a2a_permute     : Shard(i) -> Shard(j) all-to-all with the two permutes folded into the strided put
//   C2/C3/C21    put_segments    : ragged->ragged interval exchange, scatter-from-source, gather-to-root (Muon) as one-sided puts
//   C20          vocab_ce        : vocab-parallel cross entropy, local max/sumexp + W-way stats exchange + grad in ONE launch
//
// All kernels are bracketed by epoch-flag barriers on the arena's signal pad (slot = "inputs ready", slot+1 = "everyone done"),
// so the symmetric buffers may be reused by the next call on the same stream without any host synchronisation.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "common.cuh"

using namespace vb;

namespace {

constexpr int kMaxPeers = 16;
struct Peers {
  void* p[kMaxPeers];
};
struct Flags {
  uint32_t* p[kMaxPeers];
};

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream(); }

VB_DEVICE void spin_flag(const uint32_t* f, uint32_t epoch, int slot, int peer) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
    if (clock64() - t0 > 20000000000LL) {
      printf("[vescale_b200] collective timeout: slot=%d peer=%d want epoch %u have %u\n", slot, peer, epoch, ld_relaxed_sys(f));
      __trap();
    }
  }
}

// "inputs ready": CTA 0 publishes, every CTA waits for all peers.
VB_DEVICE void start_barrier(const Flags& pads, const uint32_t* my_pad, int world, int rank, int slot, uint32_t epoch) {
  if (blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pads.p[threadIdx.x] + slot * world + rank, epoch);
  }
  if (threadIdx.x < world) spin_flag(my_pad + slot * world + threadIdx.x, epoch, slot, threadIdx.x);
  __syncthreads();
}

// "everyone done": the last CTA of this GPU to finish publishes and waits, so the kernel retires only when every peer
// has finished reading from / writing to this GPU's buffers.
VB_DEVICE void end_barrier(uint32_t* counter, const Flags& pads, const uint32_t* my_pad, int world, int rank, int slot, uint32_t epoch) {
  __shared__ int is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    is_last = (atomicAdd(counter, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (is_last) {
    if (threadIdx.x == 0) *counter = 0;
    if (threadIdx.x < world) {
      __threadfence_system();
      st_release_sys(pads.p[threadIdx.x] + slot * world + rank, epoch);
      spin_flag(my_pad + slot * world + threadIdx.x, epoch, slot, threadIdx.x);
    }
  }
}

// ------------------------------------------------------------------------------------------------- all-reduce
template <typename T>
struct Vec16;
template <>
struct Vec16<float> {
  static constexpr int N = 4;
  VB_DEVICE static void add(float* acc, const uint4& v) {
    const float* f = reinterpret_cast<const float*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) acc[k] += f[k];
  }
  VB_DEVICE static uint4 pack(const float* acc) { return *reinterpret_cast<const uint4*>(acc); }
  VB_DEVICE static uint4 mc_ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
  }
};
template <>
struct Vec16<__nv_bfloat16> {
  static constexpr int N = 8;
  VB_DEVICE static void add(float* acc, const uint4& v) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 f = __bfloat1622float2(h[k]);
      acc[2 * k] += f.x;
      acc[2 * k + 1] += f.y;
    }
  }
  VB_DEVICE static uint4 pack(const float* acc) {
    uint4 v;
    v.x = pack_bf16x2(acc[0], acc[1]);
    v.y = pack_bf16x2(acc[2], acc[3]);
    v.z = pack_bf16x2(acc[4], acc[5]);
    v.w = pack_bf16x2(acc[6], acc[7]);
    return v;
  }
  VB_DEVICE static uint4 mc_ld_reduce(const void* p) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p)
                 : "memory");
    return v;
  }
};

VB_DEVICE void mc_st(void* p, const uint4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// MODE 0: P2P two-shot (reduce my slice from all peers, write it to all peers).  MODE 1: NVLS two-shot.
// MODE 2: one-shot (every rank reduces everything into a private output; no peer writes).
template <typename T, int MODE>
__global__ void __launch_bounds__(512) all_reduce_kernel(const __grid_constant__ Peers bufs, void* mc, uint4* __restrict__ out, size_t nvec, int world, int rank, float scale,
                                                         const __grid_constant__ Flags pads, const uint32_t* my_pad, int slot, uint32_t epoch, uint32_t* counter) {
  start_barrier(pads, my_pad, world, rank, slot, epoch);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (MODE == 2) {
    for (size_t i = tid; i < nvec; i += stride) {
      float acc[Vec16<T>::N] = {};
      uint4 v[kMaxPeers / 2];  // all peer loads are issued before the first add: one NVLink round trip, not W of them
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int k = 0; k < kMaxPeers / 2; ++k) {
          const int pi = half * (kMaxPeers / 2) + k;
          if (pi < world) v[k] = ld_stream(reinterpret_cast<const uint4*>(bufs.p[(rank + pi) % world]) + i);
        }
#pragma unroll
        for (int k = 0; k < kMaxPeers / 2; ++k)
          if (half * (kMaxPeers / 2) + k < world) Vec16<T>::add(acc, v[k]);
      }
#pragma unroll
      for (int k = 0; k < Vec16<T>::N; ++k) acc[k] *= scale;
      out[i] = Vec16<T>::pack(acc);
    }
  } else {
    const size_t per = (nvec + world - 1) / world;
    const size_t lo = min(nvec, per * rank), hi = min(nvec, lo + per);
    for (size_t i = lo + tid; i < hi; i += stride) {
      float acc[Vec16<T>::N] = {};
      if (MODE == 1) {
        Vec16<T>::add(acc, Vec16<T>::mc_ld_reduce(reinterpret_cast<const uint4*>(mc) + i));
      } else {
        uint4 v[kMaxPeers / 2];
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
          for (int k = 0; k < kMaxPeers / 2; ++k) {
            const int pi = half * (kMaxPeers / 2) + k;
            if (pi < world) v[k] = ld_stream(reinterpret_cast<const uint4*>(bufs.p[(rank + pi) % world]) + i);
          }
#pragma unroll
          for (int k = 0; k < kMaxPeers / 2; ++k)
            if (half * (kMaxPeers / 2) + k < world) Vec16<T>::add(acc, v[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < Vec16<T>::N; ++k) acc[k] *= scale;
      const uint4 r = Vec16<T>::pack(acc);
      if (MODE == 1) {
        mc_st(reinterpret_cast<uint4*>(mc) + i, r);
      } else {
        for (int pi = 0; pi < world; ++pi) reinterpret_cast<uint4*>(bufs.p[(rank + pi) % world])[i] = r;
      }
    }
  }
  end_barrier(counter, pads, my_pad, world, rank, slot + 1, epoch);
}

// ------------------------------------------------------------------------------------------------- reduce-scatter
// Rank r receives the sum over ranks of slice r of a symmetric buffer (slice_vecs 16-byte vectors starting at r * slice_vecs)
// into the private tensor `out`.  The slice is a [rows, row_vecs] matrix and `out` may have a row stride, so column chunks of
// a wider result can be reduced separately.  MODE 0: P2P loads of every peer's copy, all in flight before the first add.
// MODE 1: one multimem.ld_reduce per vector -- the NVSwitch adds, this GPU receives 1/W of what the P2P form pulls.
template <typename T, int MODE>
__global__ void __launch_bounds__(512) reduce_scatter_kernel(const __grid_constant__ Peers bufs, const void* mc, uint4* __restrict__ out, size_t slice_vecs, size_t row_vecs,
                                                             size_t out_stride_vecs, int world, int rank, float scale, const __grid_constant__ Flags pads,
                                                             const uint32_t* my_pad, int slot, uint32_t epoch, uint32_t* counter) {
  start_barrier(pads, my_pad, world, rank, slot, epoch);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t tid = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t lo = slice_vecs * (size_t)rank;
  auto store = [&](size_t i, float* acc) {
#pragma unroll
    for (int k = 0; k < Vec16<T>::N; ++k) acc[k] *= scale;
    out[(i / row_vecs) * out_stride_vecs + i % row_vecs] = Vec16<T>::pack(acc);
  };
  if (MODE == 1) {
    constexpr int U = 4;  // switch round trips in flight per thread
    for (size_t base = tid; base < slice_vecs; base += stride * U) {
      uint4 r[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = base + u * stride;
        if (i < slice_vecs) r[u] = Vec16<T>::mc_ld_reduce(reinterpret_cast<const uint4*>(mc) + lo + i);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const size_t i = base + u * stride;
        if (i < slice_vecs) {
          float acc[Vec16<T>::N] = {};
          Vec16<T>::add(acc, r[u]);
          store(i, acc);
        }
      }
    }
  } else {
    for (size_t i = tid; i < slice_vecs; i += stride) {
      float acc[Vec16<T>::N] = {};
      uint4 v[kMaxPeers / 2];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
#pragma unroll
        for (int k = 0; k < kMaxPeers / 2; ++k) {
          const int pi = half * (kMaxPeers / 2) + k;
          if (pi < world) v[k] = ld_stream(reinterpret_cast<const uint4*>(bufs.p[(rank + pi) % world]) + lo + i);
        }
#pragma unroll
        for (int k = 0; k < kMaxPeers / 2; ++k)
          if (half * (kMaxPeers / 2) + k < world) Vec16<T>::add(acc, v[k]);
      }
      store(i, acc);
    }
  }
  end_barrier(counter, pads, my_pad, world, rank, slot + 1, epoch);
}

// ------------------------------------------------------------------------------------------------- all-to-all with folded permutes
// Logical loop nest (a, i, b, j, c): the source keeps dim I local and holds all of J; peer p receives J-chunk p and places
// my I-chunk at index `rank` of its full I.  Strides are in units of VEC bytes; c is the contiguous inner run.
struct A2AArgs {
  int64_t n[5];
  int64_t ss[5];
  int64_t ds[5];
  int64_t src_peer_stride;  // advance of the source base per destination peer (J chunk)
  int64_t dst_rank_stride;  // offset inside every destination for data coming from `rank` (I chunk)
};

template <typename V>
__global__ void __launch_bounds__(512) a2a_permute_kernel(const V* __restrict__ src, const __grid_constant__ Peers dst, const __grid_constant__ A2AArgs a, int world, int rank, const __grid_constant__ Flags pads,
                                                          const uint32_t* my_pad, int slot, uint32_t epoch, uint32_t* counter) {
  start_barrier(pads, my_pad, world, rank, slot, epoch);
  const int64_t total = a.n[0] * a.n[1] * a.n[2] * a.n[3] * a.n[4];
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int pi = 0; pi < world; ++pi) {
    const int p = (rank + pi) % world;
    const V* s = src + (int64_t)p * a.src_peer_stride;
    V* d = reinterpret_cast<V*>(dst.p[p]) + (int64_t)rank * a.dst_rank_stride;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += stride) {
      int64_t r = t;
      const int64_t c = r % a.n[4];
      r /= a.n[4];
      const int64_t j = r % a.n[3];
      r /= a.n[3];
      const int64_t b = r % a.n[2];
      r /= a.n[2];
      const int64_t i = r % a.n[1];
      const int64_t aa = r / a.n[1];
      d[aa * a.ds[0] + i * a.ds[1] + b * a.ds[2] + j * a.ds[3] + c * a.ds[4]] = s[aa * a.ss[0] + i * a.ss[1] + b * a.ss[2] + j * a.ss[3] + c * a.ss[4]];
    }
  }
  end_barrier(counter, pads, my_pad, world, rank, slot + 1, epoch);
}

// ------------------------------------------------------------------------------------------------- one-sided segment puts
// table[n][4] (int64, units of VEC bytes): src_off, dst_peer, dst_off, count
template <typename V>
__global__ void __launch_bounds__(512) put_segments_kernel(const V* __restrict__ src, const __grid_constant__ Peers dst, const int64_t* __restrict__ table, int nseg, int world,
                                                           int rank, const __grid_constant__ Flags pads, const uint32_t* my_pad, int slot, uint32_t epoch, uint32_t* counter) {
  start_barrier(pads, my_pad, world, rank, slot, epoch);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int s = 0; s < nseg; ++s) {
    const int64_t so = table[4 * s], peer = table[4 * s + 1], dof = table[4 * s + 2], n = table[4 * s + 3];
    const V* sp = src + so;
    V* dp = reinterpret_cast<V*>(dst.p[peer]) + dof;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < n; t += stride) dp[t] = sp[t];
  }
  end_barrier(counter, pads, my_pad, world, rank, slot + 1, epoch);
}

// ------------------------------------------------------------------------------------------------- vocab-parallel cross entropy
// logits [T, Vloc] bf16 hold this rank's vocabulary slice [v0, v0+Vloc).  Phase 1 computes (max, sumexp, target logit) per row
// and puts them into every peer's stats[parity][rank][row]; a per-(CTA, peer) epoch flag in the same symmetric block replaces
// the two NCCL all-reduces of the reference (legacy loss.py:138,141; vp_cross_entropy.py:47,79,84).  Phase 2 merges the W
// partial statistics, writes loss[row] and overwrites the logits with (softmax - onehot)/n_valid.
// The grid is persistent (<= resident capacity) so CTA i of every rank is co-scheduled with CTA i of its peers.
constexpr int kVCEThreads = 512;
__global__ void __launch_bounds__(kVCEThreads) vocab_ce_kernel(__nv_bfloat16* __restrict__ logits, const int64_t* __restrict__ target,
                                                               const float* __restrict__ n_valid, float* __restrict__ loss, int T, int Vloc, int64_t v0,
                                                               int64_t ignore_index, const __grid_constant__ Peers stats, int world, int rank, uint32_t epoch, int max_rows,
                                                               int max_ctas) {
  __shared__ float red[33];
  const int parity = epoch & 1;
  // symmetric block layout (floats): stats[2][W][max_rows][4]  then flags[W][max_ctas] (uint32)
  const size_t stats_floats = (size_t)2 * world * max_rows * 4;
  const int nvec = Vloc / 8;
  for (int row = blockIdx.x; row < T; row += gridDim.x) {
    const __nv_bfloat16* x = logits + (size_t)row * Vloc;
    float m = -INFINITY, s = 0.f;
    for (int v = threadIdx.x; v < nvec; v += kVCEThreads) {
      float f[8];
      unpack8(ld8(x + v * 8), f);
      float lm = f[0];
#pragma unroll
      for (int k = 1; k < 8; ++k) lm = fmaxf(lm, f[k]);
      const float nm = fmaxf(m, lm);
      float acc = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc += __expf(f[k] - nm);
      s = s * __expf(m - nm) + acc;
      m = nm;
    }
    const float gm = block_max<kVCEThreads>(m, red);
    s = (m == -INFINITY) ? 0.f : s * __expf(m - gm);
    const float gs = block_sum<kVCEThreads>(s, red);
    if (threadIdx.x < world) {
      const int64_t tg = target[row] - v0;
      const float tl = (tg >= 0 && tg < Vloc) ? __bfloat162float(x[tg]) : 0.f;
      float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(stats.p[threadIdx.x]) + (((size_t)parity * world + rank) * max_rows + row) * 4);
      *dst = make_float4(gm, gs, tl, 0.f);
    }
  }
  __syncthreads();
  if (threadIdx.x < world) {
    __threadfence_system();
    uint32_t* f = reinterpret_cast<uint32_t*>(reinterpret_cast<float*>(stats.p[threadIdx.x]) + stats_floats) + (size_t)rank * max_ctas + blockIdx.x;
    st_release_sys(f, epoch);
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(reinterpret_cast<const float*>(stats.p[rank]) + stats_floats) + (size_t)threadIdx.x * max_ctas + blockIdx.x;
    spin_flag(mine, epoch, -1, threadIdx.x);
  }
  __syncthreads();
  const float* my_stats = reinterpret_cast<const float*>(stats.p[rank]) + (size_t)parity * world * max_rows * 4;
  const float nv = n_valid[0];
  for (int row = blockIdx.x; row < T; row += gridDim.x) {
    __nv_bfloat16* x = logits + (size_t)row * Vloc;
    float gm = -INFINITY;
    for (int p = 0; p < world; ++p) gm = fmaxf(gm, my_stats[((size_t)p * max_rows + row) * 4]);
    float gs = 0.f, tl = 0.f;
    for (int p = 0; p < world; ++p) {
      const float4 st = *reinterpret_cast<const float4*>(my_stats + ((size_t)p * max_rows + row) * 4);
      gs += st.y * __expf(st.x - gm);
      tl += st.z;
    }
    const int64_t tgt = target[row];
    const bool valid = tgt != ignore_index;
    const int64_t tg = tgt - v0;
    const float inv = valid ? 1.f / (gs * nv) : 0.f;
    const float sub = valid ? 1.f / nv : 0.f;
    if (threadIdx.x == 0) loss[row] = valid ? (__logf(gs) + gm - tl) : 0.f;
    for (int v = threadIdx.x; v < nvec; v += kVCEThreads) {
      float f[8];
      unpack8(ld8(x + v * 8), f);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float g = __expf(f[k] - gm) * inv;
        if ((int64_t)(v * 8 + k) == tg) g -= sub;
        f[k] = g;
      }
      st8(x + v * 8, pack8(f));
    }
  }
}

Peers to_peers(const std::vector<int64_t>& v) {
  TORCH_CHECK((int)v.size() <= kMaxPeers, "at most ", kMaxPeers, " peers");
  Peers p{};
  for (size_t i = 0; i < v.size(); ++i) p.p[i] = reinterpret_cast<void*>(v[i]);
  return p;
}
Flags to_flags(const std::vector<int64_t>& v) {
  Flags p{};
  for (size_t i = 0; i < v.size(); ++i) p.p[i] = reinterpret_cast<uint32_t*>(v[i]);
  return p;
}

int grid_for_bytes(size_t nbytes, int cap) {
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const size_t want = (nbytes / 16 + 2047) / 2048;  // ~4 vectors per thread
  return (int)std::max<size_t>(1, std::min<size_t>(want, cap > 0 ? cap : sms));
}

}  // namespace

// in place on the symmetric buffer (two-shot) or into `out` (one-shot, `out` defined)
void symm_all_reduce(std::vector<int64_t> buf_ptrs, int64_t multicast_ptr, c10::optional<at::Tensor> out, int64_t numel, int64_t dtype_code, double scale,
                     int64_t rank, std::vector<int64_t> pad_ptrs, int64_t slot, int64_t epoch, at::Tensor counter, int64_t num_ctas) {
  const int world = buf_ptrs.size();
  const int esz = dtype_code == 0 ? 4 : 2;
  TORCH_CHECK((numel * esz) % 16 == 0, "symm_all_reduce: byte size must be a multiple of 16");
  c10::cuda::CUDAGuard guard(counter.device());
  const size_t nvec = (size_t)numel * esz / 16;
  Peers bp = to_peers(buf_ptrs);
  Flags pf = to_flags(pad_ptrs);
  uint32_t* ctr = reinterpret_cast<uint32_t*>(counter.data_ptr());
  const bool oneshot = out.has_value();
  const int grid = grid_for_bytes(oneshot ? nvec * 16 : nvec * 16 / world, (int)num_ctas);
  uint4* outp = oneshot ? reinterpret_cast<uint4*>(out->data_ptr()) : nullptr;
  void* mc = reinterpret_cast<void*>(multicast_ptr);
#define VB_AR(T, MODE)                                                                                                                          \
  all_reduce_kernel<T, MODE><<<grid, 512, 0, cur_stream()>>>(bp, mc, outp, nvec, world, (int)rank, (float)scale, pf, pf.p[rank], (int)slot, \
                                                             (uint32_t)epoch, ctr)
  if (dtype_code == 0) {
    if (oneshot) VB_AR(float, 2);
    else if (multicast_ptr != 0) VB_AR(float, 1);
    else VB_AR(float, 0);
  } else {
    if (oneshot) VB_AR(__nv_bfloat16, 2);
    else if (multicast_ptr != 0) VB_AR(__nv_bfloat16, 1);
    else VB_AR(__nv_bfloat16, 0);
  }
#undef VB_AR
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// out[rows, row] (row stride out_row_stride elements) = scale * sum over ranks of slice `rank` of the symmetric buffer
void symm_reduce_scatter_t(std::vector<int64_t> buf_ptrs, int64_t multicast_ptr, at::Tensor out, int64_t slice_numel, int64_t row_numel, int64_t out_row_stride,
                           int64_t dtype_code, double scale, int64_t rank, std::vector<int64_t> pad_ptrs, int64_t slot, int64_t epoch, at::Tensor counter,
                           int64_t num_ctas) {
  const int world = buf_ptrs.size();
  const int esz = dtype_code == 0 ? 4 : 2;
  TORCH_CHECK(out.is_cuda() && (int)out.element_size() == esz, "symm_reduce_scatter_t: out dtype does not match dtype_code");
  TORCH_CHECK(row_numel > 0 && slice_numel % row_numel == 0, "symm_reduce_scatter_t: the slice must be whole rows");
  TORCH_CHECK((row_numel * esz) % 16 == 0 && (out_row_stride * esz) % 16 == 0 && reinterpret_cast<uintptr_t>(out.data_ptr()) % 16 == 0,
              "symm_reduce_scatter_t: rows, the output row stride and the output base must be multiples of 16 bytes");
  TORCH_CHECK(out_row_stride >= row_numel);
  c10::cuda::CUDAGuard guard(counter.device());
  const size_t slice_vecs = (size_t)slice_numel * esz / 16, row_vecs = (size_t)row_numel * esz / 16, ostride = (size_t)out_row_stride * esz / 16;
  Peers bp = to_peers(buf_ptrs);
  Flags pf = to_flags(pad_ptrs);
  uint32_t* ctr = reinterpret_cast<uint32_t*>(counter.data_ptr());
  const int grid = grid_for_bytes(slice_vecs * 16, (int)num_ctas);
  const void* mc = reinterpret_cast<const void*>(multicast_ptr);
  uint4* outp = reinterpret_cast<uint4*>(out.data_ptr());
#define VB_RS(T, MODE)                                                                                                                                  \
  reduce_scatter_kernel<T, MODE><<<grid, 512, 0, cur_stream()>>>(bp, mc, outp, slice_vecs, row_vecs, ostride, world, (int)rank, (float)scale, pf, pf.p[rank], \
                                                                 (int)slot, (uint32_t)epoch, ctr)
  if (dtype_code == 0) {
    if (multicast_ptr != 0) VB_RS(float, 1);
    else VB_RS(float, 0);
  } else {
    if (multicast_ptr != 0) VB_RS(__nv_bfloat16, 1);
    else VB_RS(__nv_bfloat16, 0);
  }
#undef VB_RS
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void symm_a2a_permute(const at::Tensor& src, std::vector<int64_t> dst_ptrs, std::vector<int64_t> n, std::vector<int64_t> ss, std::vector<int64_t> ds,
                      int64_t src_peer_stride, int64_t dst_rank_stride, int64_t vec_bytes, int64_t rank, std::vector<int64_t> pad_ptrs, int64_t slot,
                      int64_t epoch, at::Tensor counter, int64_t num_ctas) {
  TORCH_CHECK(src.is_cuda() && n.size() == 5 && ss.size() == 5 && ds.size() == 5);
  c10::cuda::CUDAGuard guard(src.device());
  const int world = dst_ptrs.size();
  A2AArgs a{};
  size_t total = 1;
  for (int k = 0; k < 5; ++k) {
    a.n[k] = n[k];
    a.ss[k] = ss[k];
    a.ds[k] = ds[k];
    total *= n[k];
  }
  a.src_peer_stride = src_peer_stride;
  a.dst_rank_stride = dst_rank_stride;
  Peers dp = to_peers(dst_ptrs);
  Flags pf = to_flags(pad_ptrs);
  uint32_t* ctr = reinterpret_cast<uint32_t*>(counter.data_ptr());
  const int grid = grid_for_bytes(total * vec_bytes * world, (int)num_ctas);
#define VB_A2A(V) \
  a2a_permute_kernel<V><<<grid, 512, 0, cur_stream()>>>(reinterpret_cast<const V*>(src.data_ptr()), dp, a, world, (int)rank, pf, pf.p[rank], (int)slot, (uint32_t)epoch, ctr)
  if (vec_bytes == 16) VB_A2A(uint4);
  else if (vec_bytes == 8) VB_A2A(uint2);
  else if (vec_bytes == 4) VB_A2A(uint32_t);
  else if (vec_bytes == 2) VB_A2A(uint16_t);
  else TORCH_CHECK(false, "a2a_permute: vec_bytes must be 2, 4, 8 or 16");
#undef VB_A2A
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void symm_put_segments(const at::Tensor& src, std::vector<int64_t> dst_ptrs, const at::Tensor& table, int64_t vec_bytes, int64_t total_vecs, int64_t rank,
                       std::vector<int64_t> pad_ptrs, int64_t slot, int64_t epoch, at::Tensor counter, int64_t num_ctas) {
  TORCH_CHECK(src.is_cuda() && table.is_cuda() && table.scalar_type() == at::kLong && table.is_contiguous());
  c10::cuda::CUDAGuard guard(src.device());
  const int world = dst_ptrs.size();
  const int nseg = table.numel() / 4;
  Peers dp = to_peers(dst_ptrs);
  Flags pf = to_flags(pad_ptrs);
  uint32_t* ctr = reinterpret_cast<uint32_t*>(counter.data_ptr());
  const int grid = grid_for_bytes((size_t)total_vecs * vec_bytes, (int)num_ctas);
#define VB_PUT(V)                                                                                                                                   \
  put_segments_kernel<V><<<grid, 512, 0, cur_stream()>>>(reinterpret_cast<const V*>(src.data_ptr()), dp, table.data_ptr<int64_t>(), nseg, world, \
                                                         (int)rank, pf, pf.p[rank], (int)slot, (uint32_t)epoch, ctr)
  if (vec_bytes == 16) VB_PUT(uint4);
  else if (vec_bytes == 4) VB_PUT(uint32_t);
  else if (vec_bytes == 2) VB_PUT(uint16_t);
  else if (vec_bytes == 1) VB_PUT(uint8_t);
  else TORCH_CHECK(false, "put_segments: vec_bytes must be 1, 2, 4 or 16");
#undef VB_PUT
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

at::Tensor symm_vocab_ce_(at::Tensor logits, const at::Tensor& target, const at::Tensor& n_valid, int64_t vocab_start, int64_t ignore_index,
                          std::vector<int64_t> stats_ptrs, int64_t rank, int64_t epoch, int64_t max_rows, int64_t max_ctas) {
  TORCH_CHECK(logits.is_cuda() && logits.scalar_type() == at::kBFloat16 && logits.dim() == 2 && logits.is_contiguous());
  TORCH_CHECK(target.scalar_type() == at::kLong && target.is_contiguous() && n_valid.scalar_type() == at::kFloat);
  const int T = logits.size(0), V = logits.size(1);
  TORCH_CHECK(V % 8 == 0, "vocab_ce: local vocabulary must be a multiple of 8");
  TORCH_CHECK(T <= max_rows, "vocab_ce: more rows than the symmetric stats block was sized for");
  c10::cuda::CUDAGuard guard(logits.device());
  auto loss = at::empty({T}, logits.options().dtype(at::kFloat));
  if (T == 0) return loss;
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = (int)std::min<int64_t>(std::min<int64_t>(T, max_ctas), 2 * sms);
  vocab_ce_kernel<<<grid, kVCEThreads, 0, cur_stream()>>>((__nv_bfloat16*)logits.data_ptr(), target.data_ptr<int64_t>(), n_valid.data_ptr<float>(),
                                                         loss.data_ptr<float>(), T, V, vocab_start, ignore_index, to_peers(stats_ptrs),
                                                         (int)stats_ptrs.size(), (int)rank, (uint32_t)epoch, (int)max_rows, (int)max_ctas);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
  return loss;
}
