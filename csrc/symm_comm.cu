// Symmetric-memory collectives over NVLink 5 / NVSwitch for the FSDP unit buffers (sm_100a).
//
// Every rank maps every peer's arena (cuMem fabric/fd handles exchanged at bootstrap) so a kernel can load
// from / store to any peer with ordinary global instructions; NVSwitch gives each GPU 900 GB/s per direction
// to any mix of peers, so the kernels are written as *pull* operations that stripe requests over all peers
// at once.  Cross-GPU ordering uses 32-bit epoch flags in a per-rank signal pad with st.release.sys /
// ld.acquire.sys — no host synchronisation, no NCCL kernel, no separate cast/scale/norm passes.
//
//   signal_all / wait_all    : one-sided flag barrier pieces (epoch-valued, monotonic, never reset)
//   all_gather_pull          : full[p*S .. (p+1)*S) <- peer p's shard, all peers in flight together
//   reduce_scatter_fused     : my fp32 grad shard = scale * sum_p peer_p.full_grad[my slice]  (+ sum of squares
//                              for the global grad norm) — C4/C7 of SURVEY §2F fused with cast/scale/norm
//   rs_adamw_fused           : the same reduction feeding AdamW directly (no gradient ever written to HBM)
//   multimem variants        : NVLS in-switch reduction (multimem.ld_reduce) when a multicast mapping exists
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>

#include "common.cuh"

using namespace vb;

namespace {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  const void* p[kMaxPeers];
};
struct PeerFlags {
  uint32_t* p[kMaxPeers];
};

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream(); }

// ------------------------------------------------------------------------------------------------- signalling
// Pad layout (uint32): slot-major, [slot][src_rank].  Rank r signals slot s by writing `epoch` into
// peer_pad[s*W + r] on every peer; waiting means spinning on my own pad until all W entries reach `epoch`.
__global__ void signal_all_kernel(PeerFlags pads, int world, int rank, int slot, uint32_t epoch) {
  const int p = threadIdx.x;
  if (p < world) {
    __threadfence_system();
    st_release_sys(pads.p[p] + slot * world + rank, epoch);
  }
}

// Spin with a watchdog: a protocol bug must surface as a trapped kernel with a message, not as a hung box.
VB_DEVICE void spin_until(const uint32_t* f, uint32_t epoch, int slot, int peer) {
  const long long t0 = clock64();
  while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
    if (clock64() - t0 > 20000000000LL) {  // ~10 s at 2 GHz
      printf("[vescale_b200] signal timeout: slot=%d peer=%d want epoch %u have %u\n", slot, peer, epoch, ld_relaxed_sys(f));
      __trap();
    }
  }
}

__global__ void wait_all_kernel(const uint32_t* my_pad, int world, int slot, uint32_t epoch) {
  const int p = threadIdx.x;
  if (p < world) spin_until(my_pad + slot * world + p, epoch, slot, p);
}

VB_DEVICE void block_signal_then_wait(const PeerFlags& pads, const uint32_t* my_pad, int world, int rank, int slot, uint32_t epoch, bool do_signal) {
  // every CTA waits; only CTA 0 signals (after making this GPU's prior writes visible system-wide)
  if (do_signal && blockIdx.x == 0 && threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(pads.p[threadIdx.x] + slot * world + rank, epoch);
  }
  if (threadIdx.x < world) spin_until(my_pad + slot * world + threadIdx.x, epoch, slot, threadIdx.x);
  __syncthreads();
}

// ------------------------------------------------------------------------------------------------- all-gather (pull)
// Work item = 16-byte vector v of peer p's shard.  Items are interleaved peer-fastest so each CTA keeps
// requests to all peers in flight.
template <int UNROLL>
VB_DEVICE void pull_range(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t begin, size_t end) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = begin + blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < end; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) v[u] = ld_stream(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) dst[i + u * stride] = v[u];
  }
  for (; i < end; i += stride) dst[i] = ld_stream(src + i);
}

// range_mode 0: whole unit.  1: only global vectors [range_lo, range_hi) (small parameters needed before the fused first
// GEMM).  2: everything except that range (the weight the fused AG⊕GEMM kernel gathers itself).
template <int UNROLL>
__global__ void __launch_bounds__(512) all_gather_pull_kernel(PeerPtrs shards, uint4* __restrict__ full, size_t vec_per_shard, int world, int rank,
                                                              PeerFlags pads, const uint32_t* my_pad, int slot, uint32_t epoch, int use_flags,
                                                              int range_mode, size_t range_lo, size_t range_hi) {
  if (use_flags) block_signal_then_wait(pads, my_pad, world, rank, slot, epoch, true);
  for (int pi = 0; pi < world; ++pi) {
    const int p = (rank + pi) % world;  // own shard first, then neighbours: spreads simultaneous requests over distinct owners
    const uint4* src = reinterpret_cast<const uint4*>(shards.p[p]);
    uint4* dst = full + (size_t)p * vec_per_shard;
    if (range_mode == 0) {
      pull_range<UNROLL>(src, dst, 0, vec_per_shard);
    } else {
      const size_t base = (size_t)p * vec_per_shard;
      const size_t a = range_lo > base ? min(range_lo - base, vec_per_shard) : 0;
      const size_t b = range_hi > base ? min(range_hi - base, vec_per_shard) : 0;
      if (range_mode == 1) {
        pull_range<UNROLL>(src, dst, a, b);
      } else {
        pull_range<UNROLL>(src, dst, 0, a);
        pull_range<UNROLL>(src, dst, max(a, b), vec_per_shard);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------- reduce-scatter fused
VB_DEVICE void acc8(float* acc, const uint4& v) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float2 f = __bfloat1622float2(h[k]);
    acc[2 * k] += f.x;
    acc[2 * k + 1] += f.y;
  }
}

struct AdamArgs {
  float* master;
  float* m;
  float* v;
  __nv_bfloat16* p_out;
  const int64_t* wd_table;
  int nseg;
  const float* coef;
  float lr, b1, b2, eps, wd, bc1, bc2;
};

// MODE 0: write fp32 grad shard + sumsq.   MODE 1: feed AdamW directly (grad never stored).
template <int MODE, int WORLD>
__global__ void __launch_bounds__(512) reduce_scatter_fused_kernel(PeerPtrs grads, size_t shard_off_vec, float* __restrict__ out, float* __restrict__ sumsq,
                                                                   size_t nvec, int rank, float scale, PeerFlags pads, const uint32_t* my_pad,
                                                                   int slot, uint32_t epoch, AdamArgs ad) {
  __shared__ float red[33];
  __shared__ int64_t seg[64 * 3];
  // wait until every peer has finished producing this bucket's gradients (and tell them mine are done)
  block_signal_then_wait(pads, my_pad, WORLD, rank, slot, epoch, true);
  if (MODE == 1) {
    for (int i = threadIdx.x; i < ad.nseg * 3 && i < 64 * 3; i += blockDim.x) seg[i] = ad.wd_table[i];
    __syncthreads();
  }
  float ss = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // requests in flight per thread: WORLD peers x U iterations (>= 8 x 16 B), see reduce_scatter_multimem_kernel
  constexpr int U = WORLD >= 8 ? 1 : (WORLD >= 4 ? 2 : (WORLD >= 2 ? 4 : 8));
  auto process = [&](size_t i, const uint4* v) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int pi = 0; pi < WORLD; ++pi) acc8(acc, v[pi]);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k] *= scale;
      ss += acc[k] * acc[k];
    }
    if (MODE == 0) {
      float4* o = reinterpret_cast<float4*>(out + i * 8);
      o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
      o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    } else {
      const size_t e = i * 8;
      float decay = 0.f;
      for (int s = 0; s < ad.nseg; ++s)
        if ((int64_t)e >= seg[3 * s] && (int64_t)e < seg[3 * s + 1]) decay = seg[3 * s + 2] ? ad.wd : 0.f;
      const float gs = ad.coef[0];
      const float inv_bc1 = 1.f / ad.bc1, inv_sqrt_bc2 = rsqrtf(ad.bc2);
      float pm[8], mm[8], vv[8];
      *reinterpret_cast<float4*>(pm) = reinterpret_cast<float4*>(ad.master + e)[0];
      *reinterpret_cast<float4*>(pm + 4) = reinterpret_cast<float4*>(ad.master + e)[1];
      *reinterpret_cast<float4*>(mm) = reinterpret_cast<float4*>(ad.m + e)[0];
      *reinterpret_cast<float4*>(mm + 4) = reinterpret_cast<float4*>(ad.m + e)[1];
      *reinterpret_cast<float4*>(vv) = reinterpret_cast<float4*>(ad.v + e)[0];
      *reinterpret_cast<float4*>(vv + 4) = reinterpret_cast<float4*>(ad.v + e)[1];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float g = acc[k] * gs;
        mm[k] = ad.b1 * mm[k] + (1.f - ad.b1) * g;
        vv[k] = ad.b2 * vv[k] + (1.f - ad.b2) * g * g;
        pm[k] = pm[k] * (1.f - ad.lr * decay) - ad.lr * (mm[k] * inv_bc1) / (sqrtf(vv[k]) * inv_sqrt_bc2 + ad.eps);
      }
      reinterpret_cast<float4*>(ad.master + e)[0] = *reinterpret_cast<float4*>(pm);
      reinterpret_cast<float4*>(ad.master + e)[1] = *reinterpret_cast<float4*>(pm + 4);
      reinterpret_cast<float4*>(ad.m + e)[0] = *reinterpret_cast<float4*>(mm);
      reinterpret_cast<float4*>(ad.m + e)[1] = *reinterpret_cast<float4*>(mm + 4);
      reinterpret_cast<float4*>(ad.v + e)[0] = *reinterpret_cast<float4*>(vv);
      reinterpret_cast<float4*>(ad.v + e)[1] = *reinterpret_cast<float4*>(vv + 4);
      st8(ad.p_out + e, pack8(pm));
    }
  };
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i + (U - 1) * stride < nvec; i += U * stride) {
    uint4 v[U][WORLD];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int pi = 0; pi < WORLD; ++pi) {
        const int p = (rank + pi) % WORLD;
        v[u][pi] = ld_stream(reinterpret_cast<const uint4*>(grads.p[p]) + shard_off_vec + i + u * stride);
      }
#pragma unroll
    for (int u = 0; u < U; ++u) process(i + u * stride, v[u]);
  }
  for (; i < nvec; i += stride) {
    uint4 v[WORLD];
#pragma unroll
    for (int pi = 0; pi < WORLD; ++pi) {
      const int p = (rank + pi) % WORLD;
      v[pi] = ld_stream(reinterpret_cast<const uint4*>(grads.p[p]) + shard_off_vec + i);
    }
    process(i, v);
  }
  ss = block_sum<512>(ss, red);
  if (threadIdx.x == 0 && sumsq != nullptr) atomicAdd(sumsq, ss);
}

// NVLS: the switch performs the W-way reduction; each GPU receives only its reduced slice.
__global__ void __launch_bounds__(512) reduce_scatter_multimem_kernel(const void* mc_base, size_t shard_off_vec, float* __restrict__ out,
                                                                      float* __restrict__ sumsq, size_t nvec, int world, int rank, float scale,
                                                                      PeerFlags pads, const uint32_t* my_pad, int slot, uint32_t epoch) {
  __shared__ float red[33];
  block_signal_then_wait(pads, my_pad, world, rank, slot, epoch, true);
  float ss = 0.f;
  const uint4* src = reinterpret_cast<const uint4*>(mc_base) + shard_off_vec;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // A switch-reduced load has ~3 us of latency: with one 16 B request per thread the kernel moved ~180 GB/s on half the SMs
  // (profiles/step_profile_n2_ce_r2.txt).  UNROLL requests in flight per thread let a quarter of the CTAs saturate the links.
  constexpr int UNROLL = 8;
  auto consume = [&](size_t i, const uint4& v) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    acc8(acc, v);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      acc[k] *= scale;
      ss += acc[k] * acc[k];
    }
    float4* o = reinterpret_cast<float4*>(out + i * 8);
    o[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    o[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  };
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  for (; i + (UNROLL - 1) * stride < nvec; i += UNROLL * stride) {
    uint4 v[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u)
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                   : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                   : "l"(src + i + u * stride)
                   : "memory");
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) consume(i + u * stride, v[u]);
  }
  for (; i < nvec; i += stride) {
    uint4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(src + i)
                 : "memory");
    consume(i, v);
  }
  ss = block_sum<512>(ss, red);
  if (threadIdx.x == 0 && sumsq != nullptr) atomicAdd(sumsq, ss);
}

PeerPtrs to_ptrs(const std::vector<int64_t>& v) {
  TORCH_CHECK((int)v.size() <= kMaxPeers, "at most ", kMaxPeers, " peers");
  PeerPtrs p{};
  for (size_t i = 0; i < v.size(); ++i) p.p[i] = reinterpret_cast<const void*>(v[i]);
  return p;
}
PeerFlags to_flags(const std::vector<int64_t>& v) {
  PeerFlags p{};
  for (size_t i = 0; i < v.size(); ++i) p.p[i] = reinterpret_cast<uint32_t*>(v[i]);
  return p;
}

int comm_grid(int sms, int frac_num, int frac_den) { return std::max(1, sms * frac_num / frac_den); }

}  // namespace

// ------------------------------------------------------------------------------------------------- host API
void symm_signal(std::vector<int64_t> pad_ptrs, int64_t rank, int64_t slot, int64_t epoch) {
  const int world = pad_ptrs.size();
  signal_all_kernel<<<1, 32, 0, cur_stream()>>>(to_flags(pad_ptrs), world, (int)rank, (int)slot, (uint32_t)epoch);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void symm_wait(int64_t my_pad, int64_t world, int64_t slot, int64_t epoch) {
  wait_all_kernel<<<1, 32, 0, cur_stream()>>>(reinterpret_cast<const uint32_t*>(my_pad), (int)world, (int)slot, (uint32_t)epoch);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void symm_all_gather(std::vector<int64_t> shard_ptrs, at::Tensor full, int64_t shard_bytes, int64_t rank, std::vector<int64_t> pad_ptrs,
                     int64_t slot, int64_t epoch, int64_t num_ctas, int64_t range_mode, int64_t range_lo_bytes, int64_t range_hi_bytes) {
  TORCH_CHECK(full.is_cuda() && full.is_contiguous() && shard_bytes % 16 == 0);
  const int world = shard_ptrs.size();
  TORCH_CHECK((int64_t)full.numel() * full.element_size() == shard_bytes * world, "all_gather: full buffer size mismatch");
  c10::cuda::CUDAGuard guard(full.device());
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  int grid = num_ctas > 0 ? (int)num_ctas : comm_grid(sms, 1, 4);
  if (range_mode == 1) grid = std::max<int>(1, std::min<int64_t>(grid, (range_hi_bytes - range_lo_bytes) / (16 * 512) + 1));
  const bool use_flags = !pad_ptrs.empty();
  PeerFlags pf = use_flags ? to_flags(pad_ptrs) : PeerFlags{};
  all_gather_pull_kernel<4><<<grid, 512, 0, cur_stream()>>>(to_ptrs(shard_ptrs), reinterpret_cast<uint4*>(full.data_ptr()), (size_t)shard_bytes / 16, world,
                                                           (int)rank, pf, use_flags ? pf.p[rank] : nullptr, (int)slot, (uint32_t)epoch, use_flags ? 1 : 0,
                                                           (int)range_mode, (size_t)range_lo_bytes / 16, (size_t)(range_hi_bytes + 15) / 16);
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

// Copy-engine all-gather: the same pull, but issued as W peer-to-peer cudaMemcpyAsync's (DMA engines over NVLink) after a one-warp
// flag exchange, so the gather occupies NO SM while the unit in front of it computes.  The SM pull kernel above takes a quarter of
// the SMs for the whole transfer and slows every GEMM it overlaps; the copy engines reach the same ~700+ GB/s
// per direction for the >= 1 MB shards of an FSDP unit.  Peers are visited starting from rank+1 so that at any instant every GPU is
// read by exactly one peer.
__global__ void handshake_kernel(PeerFlags pads, const uint32_t* my_pad, int world, int rank, int slot, uint32_t epoch) {
  const int p = threadIdx.x;
  if (p < world) {
    __threadfence_system();
    st_release_sys(pads.p[p] + slot * world + rank, epoch);
    spin_until(my_pad + slot * world + p, epoch, slot, p);
  }
}

void symm_all_gather_ce(std::vector<int64_t> shard_ptrs, at::Tensor full, int64_t shard_bytes, int64_t rank, std::vector<int64_t> pad_ptrs, int64_t slot,
                        int64_t epoch, int64_t lo_bytes, int64_t hi_bytes) {
  TORCH_CHECK(full.is_cuda() && full.is_contiguous());
  const int world = shard_ptrs.size();
  TORCH_CHECK((int64_t)full.numel() * full.element_size() == shard_bytes * world, "all_gather_ce: full buffer size mismatch");
  c10::cuda::CUDAGuard guard(full.device());
  cudaStream_t st = cur_stream();
  if (!pad_ptrs.empty()) {
    PeerFlags pf = to_flags(pad_ptrs);
    handshake_kernel<<<1, 32, 0, st>>>(pf, pf.p[rank], world, (int)rank, (int)slot, (uint32_t)epoch);
    C10_CUDA_KERNEL_LAUNCH_CHECK();
  }
  char* dst = reinterpret_cast<char*>(full.data_ptr());
  // [lo_bytes, hi_bytes) of the flat gathered unit (0, 0 = everything)
  const int64_t total = shard_bytes * world;
  const int64_t lo = (lo_bytes == 0 && hi_bytes == 0) ? 0 : lo_bytes, hi = (lo_bytes == 0 && hi_bytes == 0) ? total : hi_bytes;
  for (int pi = 1; pi <= world; ++pi) {
    const int p = (int)((rank + pi) % world);  // own shard last (local copy)
    const int64_t a = std::max<int64_t>(lo, p * shard_bytes), b = std::min<int64_t>(hi, (p + 1) * shard_bytes);
    if (b <= a) continue;
    C10_CUDA_CHECK(cudaMemcpyAsync(dst + a, reinterpret_cast<const char*>(shard_ptrs[p]) + (a - p * shard_bytes), (size_t)(b - a), cudaMemcpyDeviceToDevice, st));
  }
}

#define VB_RS_DISPATCH(MODE)                                                                                                                   \
  switch (world) {                                                                                                                              \
    case 1: reduce_scatter_fused_kernel<MODE, 1><<<grid, 512, 0, cur_stream()>>>(gp, off_vec, outp, ssp, nvec, (int)rank, (float)scale, pf, pf.p[rank], (int)slot, (uint32_t)epoch, ad); break; \
    case 2: reduce_scatter_fused_kernel<MODE, 2><<<grid, 512, 0, cur_stream()>>>(gp, off_vec, outp, ssp, nvec, (int)rank, (float)scale, pf, pf.p[rank], (int)slot, (uint32_t)epoch, ad); break; \
    case 4: reduce_scatter_fused_kernel<MODE, 4><<<grid, 512, 0, cur_stream()>>>(gp, off_vec, outp, ssp, nvec, (int)rank, (float)scale, pf, pf.p[rank], (int)slot, (uint32_t)epoch, ad); break; \
    case 8: reduce_scatter_fused_kernel<MODE, 8><<<grid, 512, 0, cur_stream()>>>(gp, off_vec, outp, ssp, nvec, (int)rank, (float)scale, pf, pf.p[rank], (int)slot, (uint32_t)epoch, ad); break; \
    default: TORCH_CHECK(false, "reduce_scatter_fused: world size must be 1, 2, 4 or 8, got ", world);                                            \
  }

void symm_reduce_scatter(std::vector<int64_t> grad_ptrs, at::Tensor out, c10::optional<at::Tensor> sumsq, int64_t shard_elems, int64_t rank,
                         double scale, std::vector<int64_t> pad_ptrs, int64_t slot, int64_t epoch, int64_t multicast_ptr, int64_t num_ctas) {
  TORCH_CHECK(out.is_cuda() && out.scalar_type() == at::kFloat && out.numel() == shard_elems && shard_elems % 8 == 0);
  const int world = grad_ptrs.size();
  c10::cuda::CUDAGuard guard(out.device());
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  // 8 requests in flight per thread (see the kernels): 32 CTAs keep > 2 MB outstanding on the NVLS path, a third of the SMs on
  // the P2P path — the reduce-scatter overlaps the next unit's backward GEMMs and every SM it holds slows them down
  const int grid = num_ctas > 0 ? (int)num_ctas : (multicast_ptr != 0 ? 32 : comm_grid(sms, 1, 3));
  const size_t nvec = shard_elems / 8, off_vec = (size_t)rank * nvec;
  float* outp = out.data_ptr<float>();
  float* ssp = sumsq.has_value() ? sumsq->data_ptr<float>() : nullptr;
  PeerFlags pf = to_flags(pad_ptrs);
  if (multicast_ptr != 0) {
    reduce_scatter_multimem_kernel<<<grid, 512, 0, cur_stream()>>>(reinterpret_cast<const void*>(multicast_ptr), off_vec, outp, ssp, nvec, world, (int)rank,
                                                                  (float)scale, pf, pf.p[rank], (int)slot, (uint32_t)epoch);
  } else {
    PeerPtrs gp = to_ptrs(grad_ptrs);
    AdamArgs ad{};
    VB_RS_DISPATCH(0)
  }
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}

void symm_rs_adamw(std::vector<int64_t> grad_ptrs, at::Tensor master, at::Tensor m, at::Tensor v, at::Tensor p_out, const at::Tensor& wd_table,
                   const at::Tensor& coef, c10::optional<at::Tensor> sumsq, int64_t rank, double scale, std::vector<int64_t> pad_ptrs, int64_t slot,
                   int64_t epoch, double lr, double b1, double b2, double eps, double wd, double bc1, double bc2, int64_t num_ctas) {
  TORCH_CHECK(master.is_cuda() && master.scalar_type() == at::kFloat && p_out.scalar_type() == at::kBFloat16 && master.numel() % 8 == 0);
  const int world = grad_ptrs.size();
  c10::cuda::CUDAGuard guard(master.device());
  const int sms = at::cuda::getCurrentDeviceProperties()->multiProcessorCount;
  const int grid = num_ctas > 0 ? (int)num_ctas : comm_grid(sms, 1, 2);
  const size_t nvec = master.numel() / 8, off_vec = (size_t)rank * nvec;
  float* outp = nullptr;
  float* ssp = sumsq.has_value() ? sumsq->data_ptr<float>() : nullptr;
  PeerFlags pf = to_flags(pad_ptrs);
  PeerPtrs gp = to_ptrs(grad_ptrs);
  AdamArgs ad{master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), (__nv_bfloat16*)p_out.data_ptr(), wd_table.data_ptr<int64_t>(),
              (int)wd_table.size(0), coef.data_ptr<float>(), (float)lr, (float)b1, (float)b2, (float)eps, (float)wd, (float)bc1, (float)bc2};
  VB_RS_DISPATCH(1)
  C10_CUDA_KERNEL_LAUNCH_CHECK();
}
