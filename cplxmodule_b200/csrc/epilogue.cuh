// Shared epilogue of the (variational) linear / conv forward kernels:
//   y = mu + bias + eps * sqrt(max(s2, 1e-8))
// (cplxmodule/nn/relevance/complex/base.py:56, real/base.py:49; bias add
// cplx.py:644-646).  One thread owns a run of C consecutive output columns of
// one output row.
#pragma once
#include "common.cuh"
#include "noise.cuh"

namespace cplxk {

struct EpiParams {
  const void* b_re;
  const void* b_im;
  const void* eps_re;
  const void* eps_im;
  void* y_re;
  void* y_im;
  void* s2_out;        // nullable: the variance accumulator, saved for the backward pass
  int64_t M, N;        // logical output matrix [M, N] (rows = samples / output pixels)
  int64_t plane_elems; // elements of one noise plane in torch's randn call (= M*N)
  NoiseParams noise;
};

// optional by-product of the variational forward: the layer's KL sum, evaluated by the operand
// pre-pass on the weight rows it reads anyway (kind < 0: not requested)
struct KlFuse {
  int kind;
  float* sum;            // device float
  void* ws;              // KlWorkspace
  int64_t row_begin = 0; // weight rows [row_begin, row_end) enter the sum (a rank's KL shard);
  int64_t row_end = -1;  // row_end < 0: all rows
  void* event = nullptr; // cudaEvent_t recorded right after the pre-pass launch (nullable)
  void* fp = nullptr;    // device uint64: fingerprint of the parameters the sum was computed from (nullable)
};

template <typename T, int C>
__device__ __forceinline__ void store_s2_run(const EpiParams& p, int64_t row_off, int64_t n0,
                                             int nvalid, const float (&s2)[C]) {
  if (!p.s2_out) return;
  T* dst = static_cast<T*>(p.s2_out) + row_off + n0;
  constexpr int V = Elem<T>::kVec;
  if (nvalid == C && (reinterpret_cast<uintptr_t>(dst) & 15u) == 0 && C % V == 0) {
#pragma unroll
    for (int c = 0; c < C / V; ++c) {
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = s2[c * V + j];
      o.store(dst + c * V);
    }
  } else {
#pragma unroll
    for (int j = 0; j < C; ++j)
      if (j < nvalid) dst[j] = Elem<T>::from_f(s2[j]);
  }
}

// `row_off` = linear offset of (m, 0) inside a plane; consecutive n are contiguous.
template <typename T, bool kCplx, bool kVD, int C>
__device__ __forceinline__ void epilogue_run(const EpiParams& p, int64_t m, int64_t n0,
                                             float (&re)[C], float (&im)[C], float (&s2)[C]) {
  static_assert(C % 4 == 0, "runs are whole noise quads");
  if (m >= p.M || n0 >= p.N) return;
  const int64_t row_off = m * p.N;
  const int nvalid = (p.N - n0) < C ? static_cast<int>(p.N - n0) : C;

  if (p.b_re) {
    const T* br = static_cast<const T*>(p.b_re);
    const T* bi = static_cast<const T*>(p.b_im);
#pragma unroll
    for (int j = 0; j < C; ++j)
      if (j < nvalid) {
        re[j] += Elem<T>::to_f(__ldg(br + n0 + j));
        if constexpr (kCplx) im[j] += Elem<T>::to_f(__ldg(bi + n0 + j));
      }
  }

  if constexpr (kVD) {
    store_s2_run<T, C>(p, row_off, n0, nvalid, s2);
    float sd[C];
#pragma unroll
    for (int j = 0; j < C; ++j) sd[j] = sd_of(s2[j]);

    if (p.noise.mode == CPLXK_NOISE_INJECT) {
      const T* er = static_cast<const T*>(p.eps_re) + row_off + n0;
      const T* ei = kCplx ? static_cast<const T*>(p.eps_im) + row_off + n0 : nullptr;
#pragma unroll
      for (int j = 0; j < C; ++j)
        if (j < nvalid) {
          re[j] = fmaf(Elem<T>::to_f(__ldg(er + j)), sd[j], re[j]);
          if constexpr (kCplx) im[j] = fmaf(Elem<T>::to_f(__ldg(ei + j)), sd[j], im[j]);
        }
    } else if (p.noise.mode == CPLXK_NOISE_PHILOX_TORCH) {
      TorchNoiseCursor cur;
      cur.seek(static_cast<uint64_t>(row_off + n0), p.noise.threads);
#pragma unroll
      for (int j = 0; j < C; ++j)
        if (j < nvalid) re[j] = fmaf(cur.next(p.noise) * p.noise.scale, sd[j], re[j]);
      if constexpr (kCplx) {
        cur.seek(static_cast<uint64_t>(p.plane_elems + row_off + n0), p.noise.threads);
#pragma unroll
        for (int j = 0; j < C; ++j)
          if (j < nvalid) im[j] = fmaf(cur.next(p.noise) * p.noise.scale, sd[j], im[j]);
      }
    } else {  // CPLXK_NOISE_PHILOX_FAST: quads are per-row groups of four columns
      const uint64_t quads_per_row = static_cast<uint64_t>((p.N + 3) >> 2);
      const uint64_t q0 = static_cast<uint64_t>(m) * quads_per_row + static_cast<uint64_t>(n0 >> 2);
#pragma unroll
      for (int q = 0; q < C / 4; ++q) {
        float4 g = philox_fast_normal4(q0 + q, 0u, p.noise);
        re[4 * q + 0] = fmaf(g.x * p.noise.scale, sd[4 * q + 0], re[4 * q + 0]);
        re[4 * q + 1] = fmaf(g.y * p.noise.scale, sd[4 * q + 1], re[4 * q + 1]);
        re[4 * q + 2] = fmaf(g.z * p.noise.scale, sd[4 * q + 2], re[4 * q + 2]);
        re[4 * q + 3] = fmaf(g.w * p.noise.scale, sd[4 * q + 3], re[4 * q + 3]);
        if constexpr (kCplx) {
          float4 h = philox_fast_normal4(q0 + q, 1u, p.noise);
          im[4 * q + 0] = fmaf(h.x * p.noise.scale, sd[4 * q + 0], im[4 * q + 0]);
          im[4 * q + 1] = fmaf(h.y * p.noise.scale, sd[4 * q + 1], im[4 * q + 1]);
          im[4 * q + 2] = fmaf(h.z * p.noise.scale, sd[4 * q + 2], im[4 * q + 2]);
          im[4 * q + 3] = fmaf(h.w * p.noise.scale, sd[4 * q + 3], im[4 * q + 3]);
        }
      }
    }
  }

  // ---- store
  constexpr int V = Elem<T>::kVec;
  T* yr = static_cast<T*>(p.y_re) + row_off + n0;
  T* yi = kCplx ? static_cast<T*>(p.y_im) + row_off + n0 : nullptr;
  const bool vec = (nvalid == C) && ((reinterpret_cast<uintptr_t>(yr) & 15u) == 0) &&
                   (!kCplx || (reinterpret_cast<uintptr_t>(yi) & 15u) == 0) && (C % V == 0);
  if (vec) {
#pragma unroll
    for (int c = 0; c < C / V; ++c) {
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = re[c * V + j];
      o.store(yr + c * V);
      if constexpr (kCplx) {
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = im[c * V + j];
        o.store(yi + c * V);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < C; ++j)
      if (j < nvalid) {
        yr[j] = Elem<T>::from_f(re[j]);
        if constexpr (kCplx) yi[j] = Elem<T>::from_f(im[j]);
      }
  }
}


// ---- two-phase epilogue: the noise of a thread's run is produced (or fetched) BEFORE the
// accumulators are ready -- i.e. while the tensor pipe is still busy -- and kept in
// registers; the second phase only needs bias, sqrt, one FMA per plane and the stores.
template <typename T, bool kCplx, int R>
__device__ __forceinline__ void noise_prefetch(const EpiParams& p, int64_t m, int64_t n0,
                                               float (&nre)[R], float (&nim)[kCplx ? R : 1]) {
  static_assert(R % 4 == 0, "runs are whole noise quads");
#pragma unroll
  for (int j = 0; j < R; ++j) nre[j] = 0.f;
  if constexpr (kCplx) {
#pragma unroll
    for (int j = 0; j < R; ++j) nim[j] = 0.f;
  }
  if (m >= p.M || n0 >= p.N) return;
  const int64_t row_off = m * p.N;
  const int nvalid = (p.N - n0) < R ? static_cast<int>(p.N - n0) : R;
  if (p.noise.mode == CPLXK_NOISE_INJECT) {
    constexpr int V = Elem<T>::kVec;
    const T* er = static_cast<const T*>(p.eps_re) + row_off + n0;
    const T* ei = kCplx ? static_cast<const T*>(p.eps_im) + row_off + n0 : nullptr;
    const bool vec = (nvalid == R) && ((reinterpret_cast<uintptr_t>(er) & 15u) == 0) &&
                     (!kCplx || (reinterpret_cast<uintptr_t>(ei) & 15u) == 0) && (R % V == 0);
    if (vec) {
#pragma unroll
      for (int c = 0; c < R / V; ++c) {
        Vec16<T> a;
        a.load(er + c * V);
#pragma unroll
        for (int j = 0; j < V; ++j) nre[c * V + j] = a.v[j];
        if constexpr (kCplx) {
          a.load(ei + c * V);
#pragma unroll
          for (int j = 0; j < V; ++j) nim[c * V + j] = a.v[j];
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < R; ++j)
        if (j < nvalid) {
          nre[j] = Elem<T>::to_f(__ldg(er + j));
          if constexpr (kCplx) nim[j] = Elem<T>::to_f(__ldg(ei + j));
        }
    }
  } else if (p.noise.mode == CPLXK_NOISE_PHILOX_TORCH) {
    // A ROLLED loop (8 normals per plane per trip) that shifts the register-resident run
    // down by 8 and appends the new values: straight-line code for all R values would be
    // ~200 KB of SASS executed once per tile, i.e. instruction-fetch bound.  The 8 Philox
    // chains of a trip are independent (branch-free cursor arithmetic) so they interleave.
    static_assert(R % 8 == 0, "trip size");
    const uint32_t Tn = p.noise.threads;
    uint64_t slot_re = static_cast<uint64_t>(row_off + n0) / Tn;
    uint32_t idx_re = static_cast<uint32_t>(static_cast<uint64_t>(row_off + n0) - slot_re * Tn);
    uint64_t slot_im = 0;
    uint32_t idx_im = 0;
    if constexpr (kCplx) {
      slot_im = static_cast<uint64_t>(p.plane_elems + row_off + n0) / Tn;
      idx_im = static_cast<uint32_t>(static_cast<uint64_t>(p.plane_elems + row_off + n0) - slot_im * Tn);
    }
#pragma unroll 1
    for (int trip = 0; trip < R / 8; ++trip) {
#pragma unroll
      for (int j = 0; j < R - 8; ++j) nre[j] = nre[j + 8];
      if (idx_re + 7u < Tn) {   // the run does not wrap the thread index: one slot for all eight
        const uint64_t ctr = p.noise.ctr_base + (slot_re >> 2);
        const uint32_t comp = static_cast<uint32_t>(slot_re) & 3u;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          nre[R - 8 + j] = philox_torch_normal_at(ctr, idx_re + j, comp, p.noise) * p.noise.scale;
      } else {
#pragma unroll 1
        for (int j = 0; j < 8; ++j) {
          uint32_t i = idx_re + j;
          const bool wrap = i >= Tn;
          i = wrap ? i - Tn : i;
          const float v = philox_torch_normal(i, slot_re + (wrap ? 1u : 0u), p.noise) * p.noise.scale;
#pragma unroll
          for (int q = 0; q < 8; ++q)
            if (q == j) nre[R - 8 + q] = v;
        }
      }
      idx_re += 8;
      if (idx_re >= Tn) idx_re -= Tn, ++slot_re;
      if constexpr (kCplx) {
#pragma unroll
        for (int j = 0; j < R - 8; ++j) nim[j] = nim[j + 8];
        if (idx_im + 7u < Tn) {
          const uint64_t ctr = p.noise.ctr_base + (slot_im >> 2);
          const uint32_t comp = static_cast<uint32_t>(slot_im) & 3u;
#pragma unroll
          for (int j = 0; j < 8; ++j)
            nim[R - 8 + j] = philox_torch_normal_at(ctr, idx_im + j, comp, p.noise) * p.noise.scale;
        } else {
#pragma unroll 1
          for (int j = 0; j < 8; ++j) {
            uint32_t i = idx_im + j;
            const bool wrap = i >= Tn;
            i = wrap ? i - Tn : i;
            const float v = philox_torch_normal(i, slot_im + (wrap ? 1u : 0u), p.noise) * p.noise.scale;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              if (q == j) nim[R - 8 + q] = v;
          }
        }
        idx_im += 8;
        if (idx_im >= Tn) idx_im -= Tn, ++slot_im;
      }
    }
  } else {
    const uint64_t quads_per_row = static_cast<uint64_t>((p.N + 3) >> 2);
    const uint64_t q0 = static_cast<uint64_t>(m) * quads_per_row + static_cast<uint64_t>(n0 >> 2);
#pragma unroll 1
    for (int trip = 0; trip < R / 8; ++trip) {
#pragma unroll
      for (int j = 0; j < R - 8; ++j) nre[j] = nre[j + 8];
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float4 g = philox_fast_normal4(q0 + 2 * trip + q, 0u, p.noise);
        nre[R - 8 + 4 * q + 0] = g.x * p.noise.scale, nre[R - 8 + 4 * q + 1] = g.y * p.noise.scale;
        nre[R - 8 + 4 * q + 2] = g.z * p.noise.scale, nre[R - 8 + 4 * q + 3] = g.w * p.noise.scale;
      }
      if constexpr (kCplx) {
#pragma unroll
        for (int j = 0; j < R - 8; ++j) nim[j] = nim[j + 8];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          float4 h = philox_fast_normal4(q0 + 2 * trip + q, 1u, p.noise);
          nim[R - 8 + 4 * q + 0] = h.x * p.noise.scale, nim[R - 8 + 4 * q + 1] = h.y * p.noise.scale;
          nim[R - 8 + 4 * q + 2] = h.z * p.noise.scale, nim[R - 8 + 4 * q + 3] = h.w * p.noise.scale;
        }
      }
    }
  }
}

// second phase for C consecutive columns starting at n0 (noise pointers already offset)
template <typename T, bool kCplx, int C>
__device__ __forceinline__ void epilogue_finish(const EpiParams& p, int64_t m, int64_t n0,
                                                float (&re)[C], float (&im)[C], float (&s2)[C],
                                                const float* nre, const float* nim) {
  if (m >= p.M || n0 >= p.N) return;
  const int64_t row_off = m * p.N;
  const int nvalid = (p.N - n0) < C ? static_cast<int>(p.N - n0) : C;
  if (p.b_re) {
    const T* br = static_cast<const T*>(p.b_re);
    const T* bi = static_cast<const T*>(p.b_im);
#pragma unroll
    for (int j = 0; j < C; ++j)
      if (j < nvalid) {
        re[j] += Elem<T>::to_f(__ldg(br + n0 + j));
        if constexpr (kCplx) im[j] += Elem<T>::to_f(__ldg(bi + n0 + j));
      }
  }
  store_s2_run<T, C>(p, row_off, n0, nvalid, s2);
#pragma unroll
  for (int j = 0; j < C; ++j) {
    const float sd = sd_of(s2[j]);
    re[j] = fmaf(nre[j], sd, re[j]);
    if constexpr (kCplx) im[j] = fmaf(nim[j], sd, im[j]);
  }
  constexpr int V = Elem<T>::kVec;
  T* yr = static_cast<T*>(p.y_re) + row_off + n0;
  T* yi = kCplx ? static_cast<T*>(p.y_im) + row_off + n0 : nullptr;
  const bool vec = (nvalid == C) && ((reinterpret_cast<uintptr_t>(yr) & 15u) == 0) &&
                   (!kCplx || (reinterpret_cast<uintptr_t>(yi) & 15u) == 0) && (C % V == 0);
  if (vec) {
#pragma unroll
    for (int c = 0; c < C / V; ++c) {
      Vec16<T> o;
#pragma unroll
      for (int j = 0; j < V; ++j) o.v[j] = re[c * V + j];
      o.store(yr + c * V);
      if constexpr (kCplx) {
#pragma unroll
        for (int j = 0; j < V; ++j) o.v[j] = im[c * V + j];
        o.store(yi + c * V);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < C; ++j)
      if (j < nvalid) {
        yr[j] = Elem<T>::from_f(re[j]);
        if constexpr (kCplx) yi[j] = Elem<T>::from_f(im[j]);
      }
  }
}

}  // namespace cplxk
