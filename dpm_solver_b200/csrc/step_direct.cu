// step_direct.cu -- variant 0: persistent grid, 128/256-bit global loads straight to registers.
//
// One thread owns kUnroll packets of 8 consecutive elements per tile; all loads of a tile are
// issued before the first use so that (streams x kUnroll) 16/32-byte requests are in flight per
// thread. HBM-bound: (k+2)*s algorithmic bytes per element for a pure order-k update.
#include "common.cuh"
#include "launch.cuh"

namespace dpm {

#ifndef DPM_UNROLL
#define DPM_UNROLL 2
#endif
constexpr int kUnroll = DPM_UNROLL;
constexpr int kMaxThreads = 512;

// FAST: the straight-line packet code of common.cuh (fast_model8 / fast_update8; launch-time test
// fast_path_ok). FAST = false keeps every parameterisation, per-element thresholds and IEEE divisions.
// RND (only with FAST = false): reference-rounding mode (dpm_step_desc.raw_round) on the vector path -- the 16-bit
// CFG combine (three rounded ops) and the rounded differences of raw 16-bit buffers, per element, between 128/256-bit
// loads and stores.
template <typename TE, typename TS, int NE, int FORM, bool FAST, bool RND = false>
__global__ void __launch_bounds__(kMaxThreads)
    k_step_direct(const __grid_constant__ KParams p) {
  using Needs = FormNeeds<FORM>;
  const TS* __restrict__ gx = static_cast<const TS*>(p.x);
  const TS* __restrict__ gxe = static_cast<const TS*>(p.xe);
  const TS* __restrict__ gm0 = static_cast<const TS*>(p.m0);
  const TS* __restrict__ gm1 = static_cast<const TS*>(p.m1);
  const TS* __restrict__ gm2 = static_cast<const TS*>(p.m2);
  const TE* __restrict__ gec = static_cast<const TE*>(p.ec);
  const TE* __restrict__ geu = static_cast<const TE*>(p.eu);
  TS* __restrict__ gmo = static_cast<TS*>(p.m_out);
  TS* __restrict__ go = static_cast<TS*>(p.out);
  TS* __restrict__ go2 = static_cast<TS*>(p.out2);

  const uint32_t npk = p.npk;
  const uint32_t tile_pk = blockDim.x * kUnroll;
  const bool sep_xe = (NE > 0) && p.use_xe && !(Needs::kX && p.xe_is_x);
  const bool clamp = (NE > 0) && (p.thr != nullptr);
  pdl_trigger();
  pdl_wait();

  for (uint64_t tile0 = (uint64_t)blockIdx.x * tile_pk; tile0 < npk;
       tile0 += (uint64_t)gridDim.x * tile_pk) {
    Raw<TS> rx[kUnroll], rxe[kUnroll], rm0[kUnroll], rm1[kUnroll], rm2[kUnroll];
    Raw<TE> rec[kUnroll], reu[kUnroll];
    // ---- issue every load of the tile ----
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t pk = tile0 + (uint64_t)u * blockDim.x + threadIdx.x;
      if (pk < npk) {
        const size_t e = (size_t)pk * kPacket;
        if (Needs::kX) ldg_pk(rx[u], gx + e);
        if (NE > 0) {
          ldg_pk(rec[u], gec + e);
          if (NE == 2) ldg_pk(reu[u], geu + e);
          if (sep_xe) ldg_pk(rxe[u], gxe + e);
        } else {
          ldg_pk(rm0[u], gm0 + e);
        }
        if (Needs::kM1) ldg_pk(rm1[u], gm1 + e);
        if (Needs::kM2) ldg_pk(rm2[u], gm2 + e);
      }
    }
    // ---- compute + store ----
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const uint64_t pk = tile0 + (uint64_t)u * blockDim.x + threadIdx.x;
      if (pk < npk) {
        const size_t e = (size_t)pk * kPacket;
        float fx[8], fxe[8], fT[8], fm1[8], fm2[8], fo[8];
        if (Needs::kX) unpack(rx[u], fx);
        if (Needs::kM1) unpack(rm1[u], fm1);
        if (Needs::kM2) unpack(rm2[u], fm2);
        if (NE > 0 && FAST) {
          float fec[8], feu[8];
          unpack(rec[u], fec);
          if (NE == 2) unpack(reu[u], feu);
          const float s_thr = clamp ? __ldg(p.thr + (uint32_t)pk / p.pk_per_sample) : 1.f;   // packet index < 2^32 (npk)
          if (sep_xe) {
            unpack(rxe[u], fxe);
            fast_model8<NE>(p, fxe, fec, feu, clamp, s_thr, fT);
          } else {
            fast_model8<NE>(p, fx, fec, feu, clamp, s_thr, fT);   // fx is only read when predict_x0 (then it is loaded)
          }
          Raw<TS> rmo;
          round_pack(rmo, fT);
          if (gmo != nullptr) stg_pk(gmo + e, rmo);
        } else if (NE > 0) {
          float fec[8], feu[8];
          unpack(rec[u], fec);
          if (NE == 2) unpack(reu[u], feu);
          if (sep_xe) {
            unpack(rxe[u], fxe);
          } else if (Needs::kX) {
#pragma unroll
            for (int i = 0; i < 8; ++i) fxe[i] = fx[i];
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) fxe[i] = 0.f;
          }
          float thr8[8];
          const bool thr_uniform = p.pk_per_sample != 0;
          if (clamp) {
            if (thr_uniform) {
              const float tpk = __ldg(p.thr + (uint32_t)(pk / p.pk_per_sample));
#pragma unroll
              for (int i = 0; i < 8; ++i) thr8[i] = tpk;
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) thr8[i] = __ldg(p.thr + (e + i) / p.per_sample);
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) thr8[i] = 1.f;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i)
            fT[i] = model_value<NE, RND>(p, fxe[i], fec[i], NE == 2 ? feu[i] : 0.f, thr8[i], clamp);
          Raw<TS> rmo;
          round_pack(rmo, fT);
          if (gmo != nullptr) stg_pk(gmo + e, rmo);
        } else {
          unpack(rm0[u], fT);
        }
        if (FORM != DPM_FORM_NONE) {
          if (FAST) {
            fast_update8<FORM>(p, fx, fT, fm1, fm2, fo);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              fo[i] = update_value<FORM, RND>(p, fx[i], fT[i], Needs::kM1 ? fm1[i] : 0.f,
                                              Needs::kM2 ? fm2[i] : 0.f);
          }
          Raw<TS> ro;
          pack(ro, fo);
          stg_pk(go + e, ro);
          if (go2 != nullptr) stg_pk(go2 + e, ro);
        }
      }
    }
  }
}

// ---- fully generic element-wise kernel: any dtype mix, any alignment, tails ------------------
// RND = true: reference-rounding mode (common.cuh); the only kernel that implements it so far.
template <bool RND>
__global__ void __launch_bounds__(256) k_step_scalar(const __grid_constant__ KParams pc) {
  KParams p = pc;
  if (pc.dev_coef != nullptr) {
    // scalars produced on the device by the adaptive controller (adaptive_ctl.cu: CO_* layout)
    const float* c = pc.dev_coef;
    p.a = c[0]; p.c0 = c[1]; p.c1 = c[2]; p.c2 = c[3];
    p.w0 = c[4]; p.w1 = c[5]; p.w2 = c[6]; p.w3 = c[7]; p.w4 = c[8];
    p.alpha_e = c[9]; p.sigma_e = c[10];
    p.fast_div = 0;
  }
  const int sd = p.state_dtype, md = p.model_dtype;
  const bool need_x = p.form != DPM_FORM_NONE;
  const bool need_m1 = p.form == DPM_FORM_LIN2 || p.form == DPM_FORM_LIN3 ||
                       p.form == DPM_FORM_DIFF2 || p.form == DPM_FORM_MS3 ||
                       p.form == DPM_FORM_SS3T;
  const bool need_m2 = p.form == DPM_FORM_LIN3 || p.form == DPM_FORM_MS3 || p.form == DPM_FORM_SS3T;
  const bool clamp = p.n_model > 0 && p.thr != nullptr;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < p.n;
       i += (size_t)gridDim.x * blockDim.x) {
    float x = need_x ? load_any(p.x, sd, i) : 0.f;
    float m1 = need_m1 ? load_any(p.m1, sd, i) : 0.f;
    float m2 = need_m2 ? load_any(p.m2, sd, i) : 0.f;
    float T0;
    if (p.n_model > 0) {
      float xe = p.use_xe ? ((need_x && p.xe_is_x) ? x : load_any(p.xe, sd, i)) : 0.f;
      float ec = load_any(p.ec, md, i);
      float eu = p.n_model == 2 ? load_any(p.eu, md, i) : 0.f;
      float thr = clamp ? p.thr[(i + p.elem_offset) / p.per_sample] : 1.f;
      float mv = p.n_model == 2 ? model_value<2, RND>(p, xe, ec, eu, thr, clamp)
                                : model_value<1, RND>(p, xe, ec, eu, thr, clamp);
      T0 = round_any(sd, mv);
      if (p.m_out) store_any(p.m_out, sd, i, mv);
    } else {
      T0 = load_any(p.m0, sd, i);
    }
    float o;
    switch (p.form) {
      case DPM_FORM_LIN1: o = update_value<DPM_FORM_LIN1, RND>(p, x, T0, m1, m2); break;
      case DPM_FORM_LIN2: o = update_value<DPM_FORM_LIN2, RND>(p, x, T0, m1, m2); break;
      case DPM_FORM_LIN3: o = update_value<DPM_FORM_LIN3, RND>(p, x, T0, m1, m2); break;
      case DPM_FORM_DIFF2: o = update_value<DPM_FORM_DIFF2, RND>(p, x, T0, m1, m2); break;
      case DPM_FORM_MS3: o = update_value<DPM_FORM_MS3, RND>(p, x, T0, m1, m2); break;
      case DPM_FORM_SS3T: o = update_value<DPM_FORM_SS3T, RND>(p, x, T0, m1, m2); break;
      default: continue;
    }
    store_any(p.out, sd, i, o);
    if (p.out2) store_any(p.out2, sd, i, o);
  }
}

// ---- dispatch ---------------------------------------------------------------------------------
typedef void (*StepKernel)(const KParams);

// NE == 0 has no model conversion: only SS3T (division by w4) has a generic twin there.
template <typename TE, typename TS, int NE, bool FAST>
static StepKernel pick_form(int form) {
  constexpr bool kTwin = NE > 0 || !FAST;   // is <FAST = false> a distinct kernel for this (NE, form)?
  switch (form) {
    case DPM_FORM_NONE: return NE > 0 ? k_step_direct<TE, TS, NE, DPM_FORM_NONE, FAST> : nullptr;
    case DPM_FORM_LIN1: return k_step_direct<TE, TS, NE, DPM_FORM_LIN1, FAST || NE == 0>;
    case DPM_FORM_LIN2: return k_step_direct<TE, TS, NE, DPM_FORM_LIN2, FAST || NE == 0>;
    case DPM_FORM_LIN3: return k_step_direct<TE, TS, NE, DPM_FORM_LIN3, FAST || NE == 0>;
    case DPM_FORM_DIFF2: return k_step_direct<TE, TS, NE, DPM_FORM_DIFF2, FAST || NE == 0>;
    case DPM_FORM_MS3: return k_step_direct<TE, TS, NE, DPM_FORM_MS3, FAST || NE == 0>;
    case DPM_FORM_SS3T: return k_step_direct<TE, TS, NE, DPM_FORM_SS3T, FAST>;
  }
  (void)kTwin;
  return nullptr;
}
template <typename TE, typename TS, int NE>
static StepKernel pick_fast(int form, bool fast) {
  return fast ? pick_form<TE, TS, NE, true>(form) : pick_form<TE, TS, NE, false>(form);
}
template <typename TE, typename TS>
static StepKernel pick_ne(int ne, int form, bool fast) {
  switch (ne) {
    case 1: return pick_fast<TE, TS, 1>(form, fast);
    case 2: return pick_fast<TE, TS, 2>(form, fast);
  }
  return nullptr;
}
static StepKernel pick_direct(int md, int sd, int ne, int form, bool fast) {
  if (ne == 0) {
    switch (sd) {
      case DPM_F32: return pick_fast<float, float, 0>(form, fast);
      case DPM_BF16: return pick_fast<__nv_bfloat16, __nv_bfloat16, 0>(form, fast);
      case DPM_F16: return pick_fast<__half, __half, 0>(form, fast);
    }
    return nullptr;
  }
  if (md == DPM_F32 && sd == DPM_F32) return pick_ne<float, float>(ne, form, fast);
  if (md == DPM_BF16 && sd == DPM_BF16) return pick_ne<__nv_bfloat16, __nv_bfloat16>(ne, form, fast);
  if (md == DPM_F16 && sd == DPM_F16) return pick_ne<__half, __half>(ne, form, fast);
  if (md == DPM_BF16 && sd == DPM_F32) return pick_ne<__nv_bfloat16, float>(ne, form, fast);
  if (md == DPM_F16 && sd == DPM_F32) return pick_ne<__half, float>(ne, form, fast);
  return nullptr;  // other mixes run on the generic kernel
}

// reference-rounding mode: fp32 state; raw network outputs in bf16 / f16 (NE >= 1), or fp32 buffers holding such raw
// outputs (NE == 0, differences rounded)
template <typename TE, int NE>
static StepKernel pick_rnd_form(int form) {
  switch (form) {
    case DPM_FORM_NONE: return NE > 0 ? k_step_direct<TE, float, NE, DPM_FORM_NONE, false, true> : nullptr;
    case DPM_FORM_LIN1: return k_step_direct<TE, float, NE, DPM_FORM_LIN1, false, true>;
    case DPM_FORM_LIN2: return k_step_direct<TE, float, NE, DPM_FORM_LIN2, false, true>;
    case DPM_FORM_LIN3: return k_step_direct<TE, float, NE, DPM_FORM_LIN3, false, true>;
    case DPM_FORM_DIFF2: return k_step_direct<TE, float, NE, DPM_FORM_DIFF2, false, true>;
    case DPM_FORM_MS3: return k_step_direct<TE, float, NE, DPM_FORM_MS3, false, true>;
    case DPM_FORM_SS3T: return k_step_direct<TE, float, NE, DPM_FORM_SS3T, false, true>;
  }
  return nullptr;
}
static StepKernel pick_rnd(int md, int sd, int ne, int form) {
  if (sd != DPM_F32) return nullptr;
  if (ne == 0) return pick_rnd_form<float, 0>(form);
  if (md == DPM_BF16) return ne == 1 ? pick_rnd_form<__nv_bfloat16, 1>(form) : pick_rnd_form<__nv_bfloat16, 2>(form);
  if (md == DPM_F16) return ne == 1 ? pick_rnd_form<__half, 1>(form) : pick_rnd_form<__half, 2>(form);
  return nullptr;
}

int launch_step_direct(const KParams& p, const Tuning& t, cudaStream_t stream) {
  StepKernel k = p.raw_round ? pick_rnd(p.model_dtype, p.state_dtype, p.n_model, p.form)
                             : pick_direct(p.model_dtype, p.state_dtype, p.n_model, p.form, fast_path_ok(p));
  if (k == nullptr) return 1;  // not served here
  const int threads = t.threads > 0 ? t.threads : 256;
  const uint32_t tile_pk = (uint32_t)threads * kUnroll;
  uint64_t tiles = ((uint64_t)p.npk + tile_pk - 1) / tile_pk;
  uint64_t cap = (uint64_t)sm_count() * (t.ctas_per_sm > 0 ? t.ctas_per_sm : 8);
  uint32_t grid = (uint32_t)(tiles < cap ? tiles : cap);
  if (grid == 0) return 0;
  cudaError_t le = launch_pdl(k, grid, (unsigned)threads, 0, stream, p);
  if (le != cudaSuccess) { set_error("step launch failed: %s", cudaGetErrorString(le)); cudaGetLastError(); return (int)le; }
  count_launch();
  return 0;
}

int launch_step_scalar(const KParams& p, cudaStream_t stream) {
  if (p.n == 0) return 0;
  const int threads = 256;
  uint64_t blocks = (p.n + threads - 1) / threads;
  uint64_t cap = (uint64_t)sm_count() * 8;
  uint32_t grid = (uint32_t)(blocks < cap ? blocks : cap);
  if (p.raw_round) k_step_scalar<true><<<grid, threads, 0, stream>>>(p);
  else k_step_scalar<false><<<grid, threads, 0, stream>>>(p);
  count_launch();
  return 0;
}

}  // namespace dpm
