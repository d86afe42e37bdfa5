// Warp-per-row vector helpers: a row of d floats (d % 4 == 0, d <= 1024) lives in NC float4 registers per
// lane, lane l owning columns (c*32 + l)*4 .. +3.  All global accesses are 128-bit and fully coalesced
// (a warp touches 512 contiguous bytes per chunk).
#pragma once
#include <cuda_bf16.h>

#include "common.cuh"

template <int NC>
struct RowVec {
  float4 v[NC];
};

template <int NC>
__device__ __forceinline__ void row_zero(RowVec<NC>& r) {
#pragma unroll
  for (int c = 0; c < NC; ++c) r.v[c] = make_float4(0.f, 0.f, 0.f, 0.f);
}

template <int NC>
__device__ __forceinline__ void row_load(RowVec<NC>& r, const float* __restrict__ p, int d, int lane) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    r.v[c] = col < d ? *reinterpret_cast<const float4*>(p + col) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

template <int NC>
__device__ __forceinline__ void row_store(const RowVec<NC>& r, float* __restrict__ p, int d, int lane) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    if (col < d) *reinterpret_cast<float4*>(p + col) = r.v[c];
  }
}

// the row's TF32 hi / lo pair (hi = upper 19 bits, lo = x - hi exactly): the operand format of the 3xTF32 tensor-core GEMMs,
// written by the producer so that no separate split pass is needed
template <int NC>
__device__ __forceinline__ void row_store_tf32_split(const RowVec<NC>& r, float* __restrict__ hi, float* __restrict__ lo, int d,
                                                     int lane) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    if (col < d) {
      float4 h;
      h.x = __uint_as_float(__float_as_uint(r.v[c].x) & 0xFFFFE000u); h.y = __uint_as_float(__float_as_uint(r.v[c].y) & 0xFFFFE000u);
      h.z = __uint_as_float(__float_as_uint(r.v[c].z) & 0xFFFFE000u); h.w = __uint_as_float(__float_as_uint(r.v[c].w) & 0xFFFFE000u);
      *reinterpret_cast<float4*>(hi + col) = h;
      *reinterpret_cast<float4*>(lo + col) = make_float4(r.v[c].x - h.x, r.v[c].y - h.y, r.v[c].z - h.z, r.v[c].w - h.w);
    }
  }
}

template <int NC>
__device__ __forceinline__ void row_add_store(const RowVec<NC>& r, float* __restrict__ p, int d, int lane) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    if (col < d) {
      float4 o = *reinterpret_cast<float4*>(p + col);
      o.x += r.v[c].x; o.y += r.v[c].y; o.z += r.v[c].z; o.w += r.v[c].w;
      *reinterpret_cast<float4*>(p + col) = o;
    }
  }
}

template <int NC>
__device__ __forceinline__ float row_dot(const RowVec<NC>& a, const RowVec<NC>& b) {
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NC; ++c)
    s += a.v[c].x * b.v[c].x + a.v[c].y * b.v[c].y + a.v[c].z * b.v[c].z + a.v[c].w * b.v[c].w;
  return warp_sum(s);
}

template <int NC>
__device__ __forceinline__ void row_scale(RowVec<NC>& r, float s) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    r.v[c].x *= s; r.v[c].y *= s; r.v[c].z *= s; r.v[c].w *= s;
  }
}

// r += s * a
template <int NC>
__device__ __forceinline__ void row_axpy(RowVec<NC>& r, float s, const RowVec<NC>& a) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    r.v[c].x = fmaf(s, a.v[c].x, r.v[c].x); r.v[c].y = fmaf(s, a.v[c].y, r.v[c].y);
    r.v[c].z = fmaf(s, a.v[c].z, r.v[c].z); r.v[c].w = fmaf(s, a.v[c].w, r.v[c].w);
  }
}

// r *= dropout multiplier of element (row_index * d + col)
template <int NC>
__device__ __forceinline__ void row_dropout(RowVec<NC>& r, const DropCfg& dc, long long row_index, int d, int lane) {
  if (dc.thresh == 0u) return;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    if (col < d) {
      uint64_t base = (uint64_t)row_index * (uint64_t)d + (uint64_t)col;
      r.v[c].x *= drop_mul(dc, base); r.v[c].y *= drop_mul(dc, base + 1);
      r.v[c].z *= drop_mul(dc, base + 2); r.v[c].w *= drop_mul(dc, base + 3);
    }
  }
}

// hi = r with the 13 low mantissa bits cleared (exactly TF32-representable), lo = r - hi (exact in fp32)
template <int NC>
__device__ __forceinline__ void row_split_tf32(const RowVec<NC>& r, RowVec<NC>& hi, RowVec<NC>& lo) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    hi.v[c].x = __uint_as_float(__float_as_uint(r.v[c].x) & 0xFFFFE000u); lo.v[c].x = r.v[c].x - hi.v[c].x;
    hi.v[c].y = __uint_as_float(__float_as_uint(r.v[c].y) & 0xFFFFE000u); lo.v[c].y = r.v[c].y - hi.v[c].y;
    hi.v[c].z = __uint_as_float(__float_as_uint(r.v[c].z) & 0xFFFFE000u); lo.v[c].z = r.v[c].z - hi.v[c].z;
    hi.v[c].w = __uint_as_float(__float_as_uint(r.v[c].w) & 0xFFFFE000u); lo.v[c].w = r.v[c].w - hi.v[c].w;
  }
}

// bf16 hi/lo split of a row (hi = bf16(x), lo = bf16(x - hi)) stored as 8-byte vectors (4 bf16 per lane per chunk)
template <int NC>
__device__ __forceinline__ void row_store_bf16_split(const RowVec<NC>& r, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                     int d, int lane) {
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    int col = (c * 32 + lane) * 4;
    if (col < d) {
      const __nv_bfloat162 h0 = __floats2bfloat162_rn(r.v[c].x, r.v[c].y), h1 = __floats2bfloat162_rn(r.v[c].z, r.v[c].w);
      const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
      const __nv_bfloat162 l0 = __floats2bfloat162_rn(r.v[c].x - f0.x, r.v[c].y - f0.y);
      const __nv_bfloat162 l1 = __floats2bfloat162_rn(r.v[c].z - f1.x, r.v[c].w - f1.y);
      *reinterpret_cast<uint2*>(hi + col) = make_uint2(*reinterpret_cast<const uint32_t*>(&h0), *reinterpret_cast<const uint32_t*>(&h1));
      *reinterpret_cast<uint2*>(lo + col) = make_uint2(*reinterpret_cast<const uint32_t*>(&l0), *reinterpret_cast<const uint32_t*>(&l1));
    }
  }
}

// y = norm_mode(x): returns the L2 norm of x.  y may alias x.
template <int NC>
__device__ __forceinline__ float row_normalize(const RowVec<NC>& x, RowVec<NC>& y, int mode) {
  float n = sqrtf(row_dot(x, x));
  y = x;
  if (mode == SRK_NORM_L2) {
    row_scale(y, 1.f / fmaxf(n, 1e-12f));
  } else if (mode == SRK_NORM_EPS) {
    row_scale(y, 1.f / (n + 1e-12f));
  } else if (mode == SRK_NORM_NISER) {
    row_scale(y, 1.f / (n + 1e-12f));
    float n1 = sqrtf(row_dot(y, y));
    row_scale(y, 1.f / n1);
  }
  return n;
}

// dx = d norm_mode(x) applied to dy; y = norm_mode(x) (final output), n = ||x||.
// For SRK_NORM_NISER `dy1` is an optional extra gradient arriving at the once-normalised rows.
template <int NC>
__device__ __forceinline__ void row_normalize_bwd(const RowVec<NC>& x, const RowVec<NC>& y, float n, int mode,
                                                  const RowVec<NC>& dy, const RowVec<NC>* dy1, RowVec<NC>& dx) {
  if (mode == SRK_NORM_NONE) {
    dx = dy;
  } else if (mode == SRK_NORM_L2) {
    dx = dy;
    if (n > 1e-12f) {
      float t = row_dot(y, dy);
      row_axpy(dx, -t, y);
      row_scale(dx, 1.f / n);
    } else {
      row_scale(dx, 1e12f);
    }
  } else {
    // y1 = x / (n + eps);  NISER additionally y2 = y1 / ||y1||
    const float ne = n + 1e-12f;
    RowVec<NC> g = dy;   // gradient wrt y1
    if (mode == SRK_NORM_NISER) {
      float n1 = n / ne;               // ||y1||
      float t = row_dot(y, dy);
      row_axpy(g, -t, y);
      row_scale(g, 1.f / n1);
      if (dy1) row_axpy(g, 1.f, *dy1);
    }
    float t = row_dot(x, g);
    dx = g;
    row_scale(dx, 1.f / ne);
    if (n > 0.f) row_axpy(dx, -t / (n * ne * ne), x);
  }
}

#define SRK_DISPATCH_NC(d, CALL)                                  \
  do {                                                            \
    int nc__ = ((d) + 127) / 128;                                 \
    if (nc__ <= 1) { constexpr int NC = 1; CALL; }                \
    else if (nc__ <= 2) { constexpr int NC = 2; CALL; }           \
    else if (nc__ <= 4) { constexpr int NC = 4; CALL; }           \
    else { constexpr int NC = 8; CALL; }                          \
  } while (0)

static inline int srk_check_dim(int d) {
  if (d <= 0 || d % 4 != 0 || d > 1024) {
    srk_set_error("embedding dim d=%d unsupported (need d %% 4 == 0 and d <= 1024)", d);
    return SRK_ERR_UNSUPPORTED;
  }
  return SRK_OK;
}
