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
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "launch.cuh"

namespace cg = cooperative_groups;

namespace dpm {

constexpr int kQThreads = 512;
constexpr int kBins = 2048;
constexpr int kSamples = 1024;      // sample keys per cluster (fast path)
constexpr int kLocalCand = 2048;    // bracket keys one CTA may collect
constexpr int kGlobalCand = 4096;   // bracket keys per cluster (kLocalCand + kGlobalCand = 3*kBins)

struct QParams {
  uint64_t lo;        // floor(pos)
  uint32_t two;       // 1 if ceil(pos) != floor(pos)
  float w;            // pos - floor(pos)
  float max_val;
  uint32_t cap;       // key capacity per CTA (elements); 0 => keys are not cached
  uint32_t slice;     // elements of the sample owned by one CTA (multiple of 8 on the packet path)
  uint32_t fast;      // try the sampled-pivot bracket first
  int32_t margin;     // half width of the bracket in sample ranks
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

// ---- block-level helpers for the fast path ---------------------------------------------------
// in-place bitonic sort of n (power of two) uint32 keys in shared memory, ascending
__device__ __forceinline__ void bitonic_sort(uint32_t* a, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int i = threadIdx.x; i < n; i += kQThreads) {
        const int l = i ^ j;
        if (l > i) {
          const uint32_t x = a[i], y = a[l];
          const bool asc = (i & k) == 0;
          if ((x > y) == asc) { a[i] = y; a[l] = x; }
        }
      }
      __syncthreads();
    }
  }
}

// ctrl word indices (all uint32 in shared memory)
enum { C_WARP = 0 /*16*/, C_MIN = 32, C_SEL = 40 /*Sel: 4 words*/, C_LO = 48, C_HI = 49, C_LT = 50, C_IN = 51,
       C_OVF = 52, C_NLOCAL = 53, C_GCOUNT = 54 };

template <typename TE, typename TS, int NE, bool VEC>
__global__ void __launch_bounds__(kQThreads)
    k_quantile(const __grid_constant__ KParams p, const __grid_constant__ QParams qp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t* hist = reinterpret_cast<uint32_t*>(smem_raw);          // [kBins] local digit histogram
  uint32_t* aux = hist + kBins;                                    // [3*kBins]: see below
  uint32_t* ctrl = aux + 3 * kBins;                                // [64]
  uint32_t* keys = ctrl + 64;                                      // [cap]
  // aux, exact path : total[3][kBins], cluster-wide digit totals (meaningful in CTA 0)
  // aux, fast path  : lcand[kLocalCand] | gcand[kGlobalCand] (CTA 0; the sample array aliases gcand)
  uint32_t* total = aux;
  uint32_t* lcand = aux;
  uint32_t* gcand = aux + kLocalCand;
  uint32_t* warp_sums = ctrl + C_WARP;
  uint32_t* min_slot = ctrl + C_MIN;
  Sel* sel = reinterpret_cast<Sel*>(ctrl + C_SEL);

  cg::cluster_group cluster = cg::this_cluster();
  const uint32_t crank = cluster.block_rank();
  const uint32_t csize = cluster.num_blocks();
  const uint64_t sample = blockIdx.x / csize;
  const int tid = threadIdx.x, lane = tid & 31;

  const uint64_t s_begin = sample * p.per_sample;                  // first element of the sample
  uint64_t c_begin = (uint64_t)crank * qp.slice;                   // slice inside the sample
  uint64_t c_end = c_begin + qp.slice;
  if (c_begin > p.per_sample) c_begin = p.per_sample;
  if (c_end > p.per_sample) c_end = p.per_sample;
  const uint32_t cnt = (uint32_t)(c_end - c_begin);

  for (int i = tid; i < 64; i += kQThreads) ctrl[i] = (i == C_MIN) ? 0xffffffffu : 0u;

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

  // ---- stage the slice: one pass over HBM, keys parked in shared memory ----
  if (qp.cap) {
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
            float fx[8], fc[8], fu[8], one[8], fT[8];
            unpack(rx[u], fx);
            unpack(rc[u], fc);
#pragma unroll
            for (int i = 0; i < 8; ++i) { fu[i] = 0.f; one[i] = 1.f; }
            if (NE == 2) unpack(ru[u], fu);
            model_values8<NE>(p, fx, fc, fu, one, false, true, fT);
            uint4* dst = reinterpret_cast<uint4*>(keys + (size_t)pk * kPacket);
            dst[0] = make_uint4(__float_as_uint(fabsf(fT[0])), __float_as_uint(fabsf(fT[1])),
                                __float_as_uint(fabsf(fT[2])), __float_as_uint(fabsf(fT[3])));
            dst[1] = make_uint4(__float_as_uint(fabsf(fT[4])), __float_as_uint(fabsf(fT[5])),
                                __float_as_uint(fabsf(fT[6])), __float_as_uint(fabsf(fT[7])));
          }
        }
      }
    } else {
      for (uint32_t i = tid; i < cnt; i += kQThreads) keys[i] = key_scalar(i);
    }
  }
  auto key_at = [&](uint32_t i) -> uint32_t { return qp.cap ? keys[i] : key_scalar(i); };
  __syncthreads();

  uint32_t key_lo = 0, key_hi = 0;
  bool done = false;

  // =============================== fast path: sampled-pivot bracket ===============================
  // 1024 evenly strided sample keys -> sort -> two pivots around the target rank -> one counting
  // pass (no atomics per element) that also compacts the few keys inside the bracket -> CTA 0
  // finishes with an exact radix select over those candidates. Exact whenever the target ranks fall
  // inside the bracket (checked with the exact counts); otherwise the full radix select below runs.
  if (qp.fast) {
    uint32_t* samp0 = csize > 1 ? cluster.map_shared_rank(gcand, 0) : gcand;
    const uint32_t per = kSamples / csize;                         // samples contributed by this CTA
    for (uint32_t j = tid; j < per; j += kQThreads) {
      const uint32_t idx = cnt ? (uint32_t)(((uint64_t)j * cnt) / per) : 0;
      samp0[crank * per + j] = cnt ? keys[idx] : 0xffffffffu;
    }
    if (csize > 1) cluster.sync(); else __syncthreads();           // (1) samples in CTA 0
    if (crank == 0) {
      bitonic_sort(gcand, kSamples);
      if (tid == 0) {
        const int64_t ps = (int64_t)(((unsigned __int128)qp.lo * kSamples) / p.per_sample);
        const int64_t lo_i = ps - qp.margin, hi_i = ps + qp.margin + 1;
        const uint32_t lo_k = lo_i < 0 ? 0u : gcand[lo_i];
        const uint32_t hi_k = hi_i >= kSamples ? 0xffffffffu : gcand[hi_i];
        for (uint32_t r = 0; r < csize; ++r) {
          uint32_t* c = csize > 1 ? cluster.map_shared_rank(ctrl, r) : ctrl;
          c[C_LO] = lo_k;
          c[C_HI] = hi_k;
        }
      }
    }
    if (csize > 1) cluster.sync(); else __syncthreads();           // (2) pivots everywhere
    const uint32_t lo_k = ctrl[C_LO], hi_k = ctrl[C_HI];
    uint32_t c_lt = 0;
    for (uint32_t i0 = 0; i0 < cnt; i0 += kQThreads) {
      const uint32_t i = i0 + tid;
      const uint32_t k = i < cnt ? keys[i] : 0xffffffffu;
      const bool in = i < cnt && k >= lo_k && k <= hi_k;
      c_lt += (i < cnt && k < lo_k) ? 1u : 0u;
      const uint32_t bal = __ballot_sync(0xffffffffu, in);
      if (bal) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&ctrl[C_NLOCAL], (uint32_t)__popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        const uint32_t pos = base + __popc(bal & ((1u << lane) - 1u));
        if (in && pos < kLocalCand) lcand[pos] = k;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c_lt += __shfl_xor_sync(0xffffffffu, c_lt, o);
    if (lane == 0 && c_lt) atomicAdd(&ctrl[C_LT + 8], c_lt);       // local scratch word (ctrl[58])
    __syncthreads();
    const uint32_t n_local = ctrl[C_NLOCAL];
    const uint32_t my_lt = ctrl[C_LT + 8];
    // publish counts to every CTA, reserve a range in CTA 0's candidate list, copy the local list
    uint32_t* ctrl0 = csize > 1 ? cluster.map_shared_rank(ctrl, 0) : ctrl;
    uint32_t* gc0 = csize > 1 ? cluster.map_shared_rank(gcand, 0) : gcand;
    if (tid == 0) {
      const uint32_t base = atomicAdd(&ctrl0[C_GCOUNT], n_local);
      ctrl[C_SEL + 4] = base;                                      // ctrl[44]
      const uint32_t ovf = (n_local > kLocalCand || base + n_local > kGlobalCand) ? 1u : 0u;
      for (uint32_t r = 0; r < csize; ++r) {
        uint32_t* c = csize > 1 ? cluster.map_shared_rank(ctrl, r) : ctrl;
        atomicAdd(&c[C_LT], my_lt);
        atomicAdd(&c[C_IN], n_local);
        if (ovf) atomicOr(&c[C_OVF], 1u);
      }
    }
    __syncthreads();
    {
      const uint32_t base = ctrl[C_SEL + 4];
      if (n_local <= kLocalCand && base + n_local <= kGlobalCand) {
        if (crank == 0 && csize == 1) {
          // single CTA: lcand and gcand are distinct regions of the same aux array
        }
        for (uint32_t i = tid; i < n_local; i += kQThreads) gc0[base + i] = lcand[i];
      }
    }
    if (csize > 1) cluster.sync(); else __syncthreads();           // (3) counts + candidates in CTA 0
    const uint64_t C_lt = ctrl[C_LT], C_in = ctrl[C_IN];
    const bool ok = !ctrl[C_OVF] && qp.lo >= C_lt && (qp.lo + qp.two) < C_lt + C_in;
    if (ok) {
      if (crank != 0) return;                                      // CTA 0 owns everything it still needs
      const uint32_t m = (uint32_t)C_in;
      uint64_t rank = qp.lo - C_lt;
      // exact 11/11/10-bit radix select over the candidates (block-local)
      uint32_t prefix = 0;
      Sel sc;
#pragma unroll 1
      for (int pass = 0; pass < 3; ++pass) {
        for (int i = tid; i < kBins; i += kQThreads) hist[i] = 0;
        __syncthreads();
        for (uint32_t i = tid; i < m; i += kQThreads) {
          const uint32_t k = gcand[i];
          if (pass == 0) atomicAdd(&hist[k >> 21], 1u);
          else if (pass == 1) { if ((k >> 21) == prefix) atomicAdd(&hist[(k >> 10) & 2047u], 1u); }
          else { if ((k >> 10) == prefix) atomicAdd(&hist[k & 1023u], 1u); }
        }
        __syncthreads();
        sc = select_bin<kBins / kQThreads>(hist, rank, warp_sums, sel);
        rank = sc.rank;
        prefix = pass == 0 ? sc.bin : (pass == 1 ? ((prefix << 11) | sc.bin) : ((prefix << 10) | sc.bin));
      }
      key_lo = prefix;
      key_hi = key_lo;
      if (qp.two && (sc.rank + 1 >= sc.cnt)) {
        uint32_t mn = 0xffffffffu;
        for (uint32_t i = tid; i < m; i += kQThreads) {
          const uint32_t k = gcand[i];
          if (k > key_lo && k < mn) mn = k;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        if (lane == 0 && mn != 0xffffffffu) atomicMin(min_slot, mn);
        __syncthreads();
        key_hi = *min_slot;
      }
      done = true;
    }
  }

  // =============================== exact path: full radix select ===================================
  if (!done) {
    for (int i = tid; i < kBins; i += kQThreads) hist[i] = 0;
    for (int i = tid; i < 3 * kBins; i += kQThreads) total[i] = 0;
    if (tid == 0) *min_slot = 0xffffffffu;
    __syncthreads();
    uint32_t* total0 = csize > 1 ? cluster.map_shared_rank(total, 0) : total;
    uint32_t* min0 = csize > 1 ? cluster.map_shared_rank(min_slot, 0) : min_slot;
    if (csize > 1) cluster.sync();                                 // every CTA has zeroed its arrays

    auto merge = [&](int pass, int nb) {
      __syncthreads();
      for (int i = tid; i < nb; i += kQThreads) {
        uint32_t v = hist[i];
        if (v) atomicAdd(&total0[pass * kBins + i], v);
      }
      if (csize > 1) cluster.sync(); else __syncthreads();
    };

    for (uint32_t i = tid; i < cnt; i += kQThreads) atomicAdd(&hist[key_at(i) >> 21], 1u);
    merge(0, kBins);
    Sel sa = select_bin<kBins / kQThreads>(total0, qp.lo, warp_sums, sel);

    for (int i = tid; i < kBins; i += kQThreads) hist[i] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < cnt; i += kQThreads) {
      uint32_t k = key_at(i);
      if ((k >> 21) == sa.bin) atomicAdd(&hist[(k >> 10) & 2047u], 1u);
    }
    merge(1, kBins);
    Sel sb = select_bin<kBins / kQThreads>(total0 + kBins, sa.rank, warp_sums, sel);
    const uint32_t prefix22 = (sa.bin << 11) | sb.bin;

    for (int i = tid; i < kBins; i += kQThreads) hist[i] = 0;
    __syncthreads();
    for (uint32_t i = tid; i < cnt; i += kQThreads) {
      uint32_t k = key_at(i);
      if ((k >> 10) == prefix22) atomicAdd(&hist[k & 1023u], 1u);
    }
    merge(2, 1024);
    Sel sc = select_bin<kBins / kQThreads>(total0 + 2 * kBins, sb.rank, warp_sums, sel);
    key_lo = (prefix22 << 10) | sc.bin;
    key_hi = key_lo;
    const bool need_next = qp.two && (sc.rank + 1 >= sc.cnt);      // uniform across the cluster
    if (need_next) {
      uint32_t m = 0xffffffffu;
      for (uint32_t i = tid; i < cnt; i += kQThreads) {
        uint32_t k = key_at(i);
        if (k > key_lo && k < m) m = k;
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) m = min(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (lane == 0 && m != 0xffffffffu) atomicMin(min0, m);
      if (csize > 1) cluster.sync(); else __syncthreads();
      key_hi = *min0;
    }
    if (csize > 1) cluster.sync();  // keep CTA 0's shared memory alive until every peer has read it
  }

  if (crank == 0 && tid == 0) {
    const float a = __uint_as_float(key_lo), b = __uint_as_float(key_hi);
    const float d = b - a;
    // at::native::lerp, CPU vectorised path: fmadd(coeff, end - start, base)
    float s = qp.w < 0.5f ? fmaf(qp.w, d, a) : fmaf(qp.w - 1.f, d, b);
    s = fmaxf(s, qp.max_val);  // torch.maximum(s, max_val) :423
    qp.s_out[sample] = s;
  }
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
  // sampled-pivot bracket: 4 sigma of the sample-rank of the target quantile, plus slack
  {
    const double f = ps > 1 ? (double)qp.lo / (double)(ps - 1) : 0.0;
    qp.margin = (int32_t)(4.0 * sqrt((double)kSamples * f * (1.0 - f)) + 3.0);
    qp.fast = (cap != 0 && ps >= 8192 && cs <= 16) ? 1u : 0u;
    if (const char* e = getenv("DPM_QUANTILE_EXACT_ONLY")) qp.fast = (e[0] == '1') ? 0u : qp.fast;
  }
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
