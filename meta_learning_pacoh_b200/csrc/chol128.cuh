// potrf + trtri of one 128 x 128 tile in shared memory, shared by the large-n MLL path (gp_big.cu: diagonal tiles of the
// blocked Cholesky) and the predictive path (gp_post.cu: the context Gram matrix).
#pragma once
#include "common.cuh"

namespace pacoh {
namespace {

constexpr int LDT = 132;                // row stride (floats) of a shared-memory tile buffer
// named barrier of the 256 threads (warps 0-7) that own the tile
__device__ __forceinline__ void conv_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Cholesky factor of the 128 x 128 tile in shared memory (256 threads: r = tid & 127, h = tid >> 7).
// In: T = full symmetric tile.  Out: T lower triangle = L (upper part stale); *s_flag != 0 if a pivot was not positive.
__device__ __forceinline__ void potrf_128(float* T, int tid, float* s_flag) {
  const int r = tid & 127, h = tid >> 7, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) *s_flag = 0.0f;
  conv_sync();
#pragma unroll 1
  for (int p = 0; p < 4; ++p) {
    const int c0 = 32 * p;
    // (1) diagonal 32 x 32 block: one warp, lane = row, the row in registers, columns eliminated left to right
    if (warp == 0) {
      float a[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 v = lds4(T + (c0 + lane) * LDT + c0 + 4 * j4);
        a[4 * j4] = v.x; a[4 * j4 + 1] = v.y; a[4 * j4 + 2] = v.z; a[4 * j4 + 3] = v.w;
      }
      bool bad = false;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        float djj = __shfl_sync(0xffffffffu, a[j], j);
        if (!(djj > 1e-20f)) { bad = true; djj = 1.0f; }
        const float d = sqrtf(djj);
        const float lj = lane == j ? d : a[j] / d;      // column j of L (rows >= j are meaningful)
        a[j] = lj;
#pragma unroll
        for (int c = j + 1; c < 32; ++c) {
          const float lc = __shfl_sync(0xffffffffu, lj, c);
          a[c] = fmaf(-lj, lc, a[c]);                   // meaningful for rows >= c
        }
      }
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        sts4(T + (c0 + lane) * LDT + c0 + 4 * j4, make_float4(a[4 * j4], a[4 * j4 + 1], a[4 * j4 + 2], a[4 * j4 + 3]));
      if (bad && lane == 0) *s_flag = 1.0f;
    }
    conv_sync();
    if (p == 3) break;
    // (2) panel below the block: row r solves x L_pp^T = a_r (forward substitution, L_pp broadcast from shared memory)
    if (h == 0 && r >= c0 + 32) {
      float x[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 v = lds4(T + r * LDT + c0 + 4 * j4);
        x[4 * j4] = v.x; x[4 * j4 + 1] = v.y; x[4 * j4 + 2] = v.z; x[4 * j4 + 3] = v.w;
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float s = x[c];
#pragma unroll
        for (int m4 = 0; m4 < (c + 3) / 4; ++m4) {
          const float4 l = lds4(T + (c0 + c) * LDT + c0 + 4 * m4);
          if (4 * m4 + 0 < c) s = fmaf(-x[4 * m4 + 0], l.x, s);
          if (4 * m4 + 1 < c) s = fmaf(-x[4 * m4 + 1], l.y, s);
          if (4 * m4 + 2 < c) s = fmaf(-x[4 * m4 + 2], l.z, s);
          if (4 * m4 + 3 < c) s = fmaf(-x[4 * m4 + 3], l.w, s);
        }
        x[c] = s / T[(c0 + c) * LDT + c0 + c];
      }
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        sts4(T + r * LDT + c0 + 4 * j4, make_float4(x[4 * j4], x[4 * j4 + 1], x[4 * j4 + 2], x[4 * j4 + 3]));
    }
    conv_sync();
    // (3) trailing update of the rows below the block: T[r][c] -= L[r][c0..] . L[c][c0..] for c0 + 32 <= c < end of r's block
    if (r >= c0 + 32) {
      float own[32];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 v = lds4(T + r * LDT + c0 + 4 * j4);
        own[4 * j4] = v.x; own[4 * j4 + 1] = v.y; own[4 * j4 + 2] = v.z; own[4 * j4 + 3] = v.w;
      }
      const int cend = (r | 31) + 1;
#pragma unroll 2
      for (int c = c0 + 32 + h; c < cend; c += 2) {
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int m4 = 0; m4 < 8; ++m4) {
          const float4 l = lds4(T + c * LDT + c0 + 4 * m4);
          s0 = fmaf(own[4 * m4], l.x, s0); s1 = fmaf(own[4 * m4 + 1], l.y, s1);
          s0 = fmaf(own[4 * m4 + 2], l.z, s0); s1 = fmaf(own[4 * m4 + 3], l.w, s1);
        }
        T[r * LDT + c] -= s0 + s1;
      }
    }
    conv_sync();
  }
}

// potrf + trtri with a separate output tile.  In: T = full symmetric tile.  Out: T lower triangle = L (upper part stale),
// X = U = L^-T (upper triangular, zeros below the diagonal).  Returns false (for every thread) if a pivot was not positive.
__device__ bool potrf_trtri_128(float* T, float* X, int tid, float* s_flag) {
  const int r = tid & 127, h = tid >> 7;
  potrf_128(T, tid, s_flag);
  const bool ok = *s_flag == 0.0f;
  // ---- X = U = L^-T: thread r (h == 0) owns row r of U = column r of L^-1: x L^T = e_r, 32 columns at a time
  if (h == 0) {
    const int rb = r >> 5;
#pragma unroll 1
    for (int cb = 0; cb < 4; ++cb) {
      float acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = (32 * cb + c == r) ? 1.0f : 0.0f;
      if (cb >= rb) {
#pragma unroll 1
        for (int mb = rb; mb < cb; ++mb) {
          float xm[32];
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 v = lds4(X + r * LDT + 32 * mb + 4 * j4);
            xm[4 * j4] = v.x; xm[4 * j4 + 1] = v.y; xm[4 * j4 + 2] = v.z; xm[4 * j4 + 3] = v.w;
          }
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
            for (int m4 = 0; m4 < 8; ++m4) {
              const float4 l = lds4(T + (32 * cb + c) * LDT + 32 * mb + 4 * m4);
              s0 = fmaf(xm[4 * m4], l.x, s0); s1 = fmaf(xm[4 * m4 + 1], l.y, s1);
              s0 = fmaf(xm[4 * m4 + 2], l.z, s0); s1 = fmaf(xm[4 * m4 + 3], l.w, s1);
            }
            acc[c] -= s0 + s1;
          }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float s = acc[c];
#pragma unroll
          for (int m4 = 0; m4 < (c + 3) / 4; ++m4) {
            const float4 l = lds4(T + (32 * cb + c) * LDT + 32 * cb + 4 * m4);
            if (4 * m4 + 0 < c) s = fmaf(-acc[4 * m4 + 0], l.x, s);
            if (4 * m4 + 1 < c) s = fmaf(-acc[4 * m4 + 1], l.y, s);
            if (4 * m4 + 2 < c) s = fmaf(-acc[4 * m4 + 2], l.z, s);
            if (4 * m4 + 3 < c) s = fmaf(-acc[4 * m4 + 3], l.w, s);
          }
          acc[c] = s / T[(32 * cb + c) * LDT + 32 * cb + c];
        }
      } else {
#pragma unroll
        for (int c = 0; c < 32; ++c) acc[c] = 0.0f;
      }
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4)
        sts4(X + r * LDT + 32 * cb + 4 * j4, make_float4(acc[4 * j4], acc[4 * j4 + 1], acc[4 * j4 + 2], acc[4 * j4 + 3]));
    }
  }
  conv_sync();
  return ok;
}


// potrf + trtri IN PLACE: T lower triangle + diagonal = L, strict upper triangle = U = L^-T (U_rr = 1 / L_rr is implied).
// Thread r (h == 0) owns row r of U and only ever writes T[r][c > r]; everything it reads from other rows lies on or below
// the diagonal, so one tile is enough (three CTAs per SM in big_potrf_kernel).
__device__ bool potrf_trtri_128_inplace(float* T, int tid, float* s_flag) {
  const int r = tid & 127, h = tid >> 7;
  potrf_128(T, tid, s_flag);
  const bool ok = *s_flag == 0.0f;
  if (h == 0) {
    const int rb = r >> 5, rl = r & 31;
    const float urr = 1.0f / T[r * LDT + r];
#pragma unroll 1
    for (int cb = rb; cb < 4; ++cb) {
      float acc[32];
#pragma unroll
      for (int c = 0; c < 32; ++c) acc[c] = (32 * cb + c == r) ? 1.0f : 0.0f;
#pragma unroll 1
      for (int mb = rb; mb < cb; ++mb) {
        float xm[32];
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
          const float4 v = lds4(T + r * LDT + 32 * mb + 4 * j4);
          xm[4 * j4] = v.x; xm[4 * j4 + 1] = v.y; xm[4 * j4 + 2] = v.z; xm[4 * j4 + 3] = v.w;
        }
        if (mb == rb) {                  // own block: entries left of the diagonal are L, the diagonal is L_rr
#pragma unroll
          for (int m = 0; m < 32; ++m) xm[m] = m < rl ? 0.0f : (m == rl ? urr : xm[m]);
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
          for (int m4 = 0; m4 < 8; ++m4) {
            const float4 l = lds4(T + (32 * cb + c) * LDT + 32 * mb + 4 * m4);
            s0 = fmaf(xm[4 * m4], l.x, s0); s1 = fmaf(xm[4 * m4 + 1], l.y, s1);
            s0 = fmaf(xm[4 * m4 + 2], l.z, s0); s1 = fmaf(xm[4 * m4 + 3], l.w, s1);
          }
          acc[c] -= s0 + s1;
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        float s = acc[c];
#pragma unroll
        for (int m4 = 0; m4 < (c + 3) / 4; ++m4) {
          const float4 l = lds4(T + (32 * cb + c) * LDT + 32 * cb + 4 * m4);
          if (4 * m4 + 0 < c) s = fmaf(-acc[4 * m4 + 0], l.x, s);
          if (4 * m4 + 1 < c) s = fmaf(-acc[4 * m4 + 1], l.y, s);
          if (4 * m4 + 2 < c) s = fmaf(-acc[4 * m4 + 2], l.z, s);
          if (4 * m4 + 3 < c) s = fmaf(-acc[4 * m4 + 3], l.w, s);
        }
        acc[c] = s / T[(32 * cb + c) * LDT + 32 * cb + c];
      }
      // store the entries right of the diagonal only
#pragma unroll
      for (int c = 0; c < 32; ++c)
        if (32 * cb + c > r) T[r * LDT + 32 * cb + c] = acc[c];
    }
  }
  conv_sync();
  return ok;
}

}  // namespace
}  // namespace pacoh
