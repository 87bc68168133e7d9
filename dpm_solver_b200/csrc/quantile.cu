// quantile.cu -- exact per-sample quantile of |x0| for dynamic thresholding
// (DPM_Solver.dynamic_thresholding_fn, dpm_solver_pytorch.py:416-423).
//
// torch.quantile(|x0|.reshape(B,-1), q, dim=1) sorts every sample; here one thread-block
// cluster owns one sample. Each CTA of the cluster streams its slice of (x, eps[, eps_u]) from
// HBM exactly once, recomputes x0 with the same device function the update kernel uses, and
// parks the fp32 bit pattern of |x0| (monotone as uint32 for non-negative floats) in shared
// memory. A 3-digit (11/11/10 bit) radix select then runs out of shared memory; the per-digit
// histograms of the CTAs are merged with distributed-shared-memory atomics into CTA 0 of the
// cluster. The two adjacent order statistics are combined with torch's CPU lerp
// (fma(w<0.5 ? w : w-1, hi-lo, w<0.5 ? lo : hi)) and floored with max_val (:423).
// HBM traffic: one read of the inputs, B floats written.
#include <cooperative_groups.h>

#include "common.cuh"
#include "launch.cuh"

namespace cg = cooperative_groups;

namespace dpm {

constexpr int kQThreads = 512;
constexpr int kBins = 2048;

struct QParams {
  uint64_t lo;        // floor(pos)
  uint32_t two;       // 1 if ceil(pos) != floor(pos)
  float w;            // pos - floor(pos)
  float max_val;
  uint32_t cap;       // key capacity per CTA (elements); 0 => keys are not cached
  uint32_t slice;     // elements of the sample owned by one CTA (multiple of 8 on the packet path)
  float* s_out;
};

struct Sel {
  uint32_t bin, cnt;
  uint64_t rank;
};

// Block-wide: find the bin of `tot[0..nbins)` that holds 0-based rank `k`; every thread returns
// the same answer. nbins == kQThreads * per_thread.
template <int PER>
__device__ __forceinline__ Sel select_bin(const uint32_t* tot, uint64_t k, uint32_t* warp_sums,
                                          Sel* out) {
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  uint32_t h[PER], local = 0;
#pragma unroll
  for (int i = 0; i < PER; ++i) { h[i] = tot[tid * PER + i]; local += h[i]; }
  uint32_t incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) warp_sums[wid] = incl;
  __syncthreads();
  if (wid == 0) {
    uint32_t v = lane < kQThreads / 32 ? warp_sums[lane] : 0, s = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += t;
    }
    if (lane < kQThreads / 32) warp_sums[lane] = s - v;  // exclusive
  }
  __syncthreads();
  uint64_t excl = (uint64_t)warp_sums[wid] + (incl - local);
  if (k >= excl && k < excl + local) {
    uint64_t c = excl;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      if (k >= c && k < c + h[i]) { out->bin = tid * PER + i; out->cnt = h[i]; out->rank = k - c; }
      c += h[i];
    }
  }
  __syncthreads();
  Sel r = *out;
  __syncthreads();
  return r;
}

template <typename TE, typename TS, int NE, bool VEC>
__global__ void __launch_bounds__(kQThreads)
    k_quantile(const __grid_constant__ KParams p, const __grid_constant__ QParams qp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);          // [kBins] local digit histogram
  uint32_t* total = hist + kBins;                                  // [3][kBins] cluster totals (rank 0)
  uint32_t* ctrl = total + 3 * kBins;                              // [64] warp sums, min slot, Sel
  uint32_t* keys = ctrl + 64;                                      // [cap]
  uint32_t* warp_sums = ctrl;                                      // 16 used
  uint32_t* min_slot = ctrl + 32;
  Sel* sel = reinterpret_cast<Sel*>(ctrl + 40);

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t crank = cluster.block_rank();
  const uint32_t csize = cluster.num_blocks();
  const uint64_t sample = blockIdx.x / csize;
  const int tid = threadIdx.x;

  const uint64_t s_begin = sample * p.per_sample;                  // first element of the sample
  uint64_t c_begin = (uint64_t)crank * qp.slice;                   // slice inside the sample
  uint64_t c_end = c_begin + qp.slice;
  if (c_begin > p.per_sample) c_begin = p.per_sample;
  if (c_end > p.per_sample) c_end = p.per_sample;
  const uint32_t cnt = (uint32_t)(c_end - c_begin);

  for (int i = tid; i < kBins; i += kQThreads) hist[i] = 0;
  for (int i = tid; i < 3 * kBins; i += kQThreads) total[i] = 0;
  if (tid == 0) *min_slot = 0xffffffffu;
  __syncthreads();

  const TS* __restrict__ gxe = static_cast<const TS*>(p.xe);
  const TE* __restrict__ gec = static_cast<const TE*>(p.ec);
  const TE* __restrict__ geu = static_cast<const TE*>(p.eu);

  // key(i) for element i of this CTA's slice, recomputed from global memory
  auto key_scalar = [&](uint32_t i) -> uint32_t {
    const size_t g = s_begin + c_begin + i;
    float xe = load_any(p.xe, p.state_dtype, g);
    float ec = load_any(p.ec, p.model_dtype, g);
    float eu = NE == 2 ? load_any(p.eu, p.model_dtype, g) : 0.f;
    float v = model_value<NE>(p, xe, ec, eu, 1.f, false);
    return __float_as_uint(fabsf(v));
  };

  // ---- pass A: stream the slice once, stage keys, histogram of bits 31..21 ----
  if (VEC) {
    const uint32_t npk = cnt / kPacket;
    const size_t e0 = s_begin + c_begin;
    for (uint32_t pk0 = 0; pk0 < npk; pk0 += 2 * kQThreads) {
      Raw<TS> rx[2];
      Raw<TE> rc[2], ru[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t pk = pk0 + u * kQThreads + tid;
        if (pk < npk) {
          const size_t e = e0 + (size_t)pk * kPacket;
          ldg_pk(rx[u], gxe + e);
          ldg_pk(rc[u], gec + e);
          if (NE == 2) ldg_pk(ru[u], geu + e);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const uint32_t pk = pk0 + u * kQThreads + tid;
        if (pk < npk) {
          float fx[8], fc[8], fu[8];
          unpack(rx[u], fx);
          unpack(rc[u], fc);
          if (NE == 2) unpack(ru[u], fu);
          uint32_t k8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float v = model_value<NE>(p, fx[i], fc[i], NE == 2 ? fu[i] : 0.f, 1.f, false);
            k8[i] = __float_as_uint(fabsf(v));
            atomicAdd(&hist[k8[i] >> 21], 1u);
          }
          if (qp.cap) {
            uint4* dst = reinterpret_cast<uint4*>(keys + (size_t)pk * kPacket);
            dst[0] = make_uint4(k8[0], k8[1], k8[2], k8[3]);
            dst[1] = make_uint4(k8[4], k8[5], k8[6], k8[7]);
          }
        }
      }
    }
  } else {
    for (uint32_t i = tid; i < cnt; i += kQThreads) {
      uint32_t k = key_scalar(i);
      atomicAdd(&hist[k >> 21], 1u);
      if (qp.cap) keys[i] = k;
    }
  }
  auto key_at = [&](uint32_t i) -> uint32_t { return qp.cap ? keys[i] : key_scalar(i); };

  uint32_t* total0 = csize > 1 ? cluster.map_shared_rank(total, 0) : total;
  uint32_t* min0 = csize > 1 ? cluster.map_shared_rank(min_slot, 0) : min_slot;

  auto merge = [&](int pass, int nb) {
    __syncthreads();
    if (pass == 0 && csize > 1) cluster.sync();  // every CTA has zeroed its arrays
    for (int i = tid; i < nb; i += kQThreads) {
      uint32_t v = hist[i];
      if (v) atomicAdd(&total0[pass * kBins + i], v);
    }
    if (csize > 1) cluster.sync(); else __syncthreads();
  };

  merge(0, kBins);
  Sel sa = select_bin<kBins / kQThreads>(total0, qp.lo, warp_sums, sel);

  // ---- pass B: bits 20..10 among keys whose top digit matches ----
  for (int i = tid; i < kBins; i += kQThreads) hist[i] = 0;
  __syncthreads();
  for (uint32_t i = tid; i < cnt; i += kQThreads) {
    uint32_t k = key_at(i);
    if ((k >> 21) == sa.bin) atomicAdd(&hist[(k >> 10) & 2047u], 1u);
  }
  merge(1, kBins);
  Sel sb = select_bin<kBins / kQThreads>(total0 + kBins, sa.rank, warp_sums, sel);
  const uint32_t prefix22 = (sa.bin << 11) | sb.bin;

  // ---- pass C: bits 9..0 ----
  for (int i = tid; i < kBins; i += kQThreads) hist[i] = 0;
  __syncthreads();
  for (uint32_t i = tid; i < cnt; i += kQThreads) {
    uint32_t k = key_at(i);
    if ((k >> 10) == prefix22) atomicAdd(&hist[k & 1023u], 1u);
  }
  merge(2, 1024);
  Sel sc = select_bin<kBins / kQThreads>(total0 + 2 * kBins, sb.rank, warp_sums, sel);
  const uint32_t key_lo = (prefix22 << 10) | sc.bin;

  // ---- upper neighbour ----
  uint32_t key_hi = key_lo;
  const bool need_next = qp.two && (sc.rank + 1 >= sc.cnt);  // uniform across the cluster
  if (need_next) {
    uint32_t m = 0xffffffffu;
    for (uint32_t i = tid; i < cnt; i += kQThreads) {
      uint32_t k = key_at(i);
      if (k > key_lo && k < m) m = k;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((tid & 31) == 0 && m != 0xffffffffu) atomicMin(min0, m);
    if (csize > 1) cluster.sync(); else __syncthreads();
    key_hi = *min0;
  }
  if (crank == 0 && tid == 0) {
    const float a = __uint_as_float(key_lo), b = __uint_as_float(key_hi);
    const float d = b - a;
    // at::native::lerp, CPU vectorised path: fmadd(coeff, end - start, base)
    float s = qp.w < 0.5f ? fmaf(qp.w, d, a) : fmaf(qp.w - 1.f, d, b);
    s = fmaxf(s, qp.max_val);  // torch.maximum(s, max_val) :423
    qp.s_out[sample] = s;
  }
  if (csize > 1) cluster.sync();  // keep CTA 0's shared memory alive until every peer has read it
}

typedef void (*QKernel)(const KParams, const QParams);

template <typename TE, typename TS>
static QKernel pick_q(int ne, bool vec) {
  if (ne == 2) return vec ? k_quantile<TE, TS, 2, true> : k_quantile<TE, TS, 2, false>;
  return vec ? k_quantile<TE, TS, 1, true> : k_quantile<TE, TS, 1, false>;
}

int launch_quantile(float* s_out, const KParams& p, uint64_t n_samples, float q, float max_val,
                    cudaStream_t stream) {
  const int md = p.model_dtype, sd = p.state_dtype;
  const uint64_t ps = p.per_sample;
  bool vec = (ps % kPacket == 0);
  // packet path needs aligned bases (sample starts stay aligned because ps % 8 == 0)
  auto al = [](const void* ptr, int dt) { return (reinterpret_cast<uintptr_t>(ptr) & (dt == DPM_F32 ? 31 : 15)) == 0; };
  if (!al(p.xe, sd) || !al(p.ec, md) || (p.n_model == 2 && !al(p.eu, md))) vec = false;

  QKernel k = nullptr;
  if (md == DPM_F32 && sd == DPM_F32) k = pick_q<float, float>(p.n_model, vec);
  else if (md == DPM_BF16 && sd == DPM_BF16) k = pick_q<__nv_bfloat16, __nv_bfloat16>(p.n_model, vec);
  else if (md == DPM_F16 && sd == DPM_F16) k = pick_q<__half, __half>(p.n_model, vec);
  else if (md == DPM_BF16 && sd == DPM_F32) k = pick_q<__nv_bfloat16, float>(p.n_model, vec);
  else if (md == DPM_F16 && sd == DPM_F32) k = pick_q<__half, float>(p.n_model, vec);
  else { set_error("dynamic threshold: unsupported dtype mix (model %d, state %d)", md, sd); return DPM_ERR_UNSUPPORTED; }

  // torch.quantile rank arithmetic, in fp32: pos = fl(q * (n-1))
  const float pos = q * (float)(ps - 1);
  const float fl = floorf(pos);
  QParams qp;
  qp.lo = (uint64_t)fl;
  qp.two = ceilf(pos) != fl ? 1u : 0u;
  qp.w = pos - fl;
  qp.max_val = max_val;
  qp.s_out = s_out;
  if (qp.lo >= ps) qp.lo = ps - 1;
  if (qp.lo + 1 >= ps) qp.two = 0;

  // cluster size: smallest power of two whose per-CTA slice fits the shared-memory key cache
  const size_t fixed = (size_t)(kBins + 3 * kBins + 64) * sizeof(uint32_t);
  const size_t budget = (size_t)max_smem_optin();
  const uint64_t cap_max = budget > fixed ? (budget - fixed) / sizeof(uint32_t) : 0;
  int cs = 1;
  uint64_t slice = ps;
  auto slice_for = [&](int c) {
    uint64_t s = (ps + c - 1) / c;
    return (s + kPacket - 1) / kPacket * kPacket;
  };
  while (cs < 16 && slice_for(cs) > cap_max) cs *= 2;
  slice = slice_for(cs);
  // prefer >= 2 CTAs' worth of parallelism per sample when samples are few and large
  while (cs < 8 && n_samples * cs < (uint64_t)sm_count() && slice_for(cs * 2) >= 4096) { cs *= 2; slice = slice_for(cs); }
  uint32_t cap = slice <= cap_max ? (uint32_t)slice : 0;  // 0: recompute keys from L2/HBM each pass
  if (cap == 0) { cs = 8; slice = slice_for(cs); }
  qp.cap = cap;
  qp.slice = (uint32_t)slice;
  if (slice > 0xffffffffull) { set_error("per_sample too large"); return DPM_ERR_UNSUPPORTED; }

  const size_t smem = fixed + (size_t)cap * sizeof(uint32_t);
  int rc = ensure_max_smem(reinterpret_cast<const void*>(k), /*nonportable_cluster=*/true);
  if (rc != 0) return rc;
  cudaError_t e;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(n_samples * cs), 1, 1);
  cfg.blockDim = dim3(kQThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (n_samples * cs > 0x7fffffffull) { set_error("too many samples"); return DPM_ERR_UNSUPPORTED; }
  e = cudaLaunchKernelEx(&cfg, k, p, qp);
  if (e != cudaSuccess) { set_error("quantile launch failed: %s", cudaGetErrorString(e)); return (int)e; }
  count_launch();
  return DPM_OK;
}

}  // namespace dpm
