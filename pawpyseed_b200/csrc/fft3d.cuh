// Pruned, scatter-fused, band-interleaved inverse 3-D FFT for plane-wave -> real-space boxes (sm_100a).
//
// Replaces linalg.c:14-45 (zero fill + scatter + DftiComputeBackward) for the projection pipeline:
//   pass Z : only the (g1,g2) columns that contain plane waves (~pi/4 (2Gmax/N)^2 = 35 % of them) are
//            transformed; their input is read straight from the sorted coefficient rows (scatter fused),
//   pass Y : only the x-planes g1 inside the sphere (~2Gmax/N = 68 %) are transformed,
//   pass X : all lines.
// HBM traffic per band: 16 B * N * (0.35*2 + 0.68*2 + 1) = 49 N instead of 16 N (scatter) + ~96 N (cuFFT, three
// full passes).  Boxes are stored band-interleaved, X[group][x][y][z][FFT_B], FFT_B = 16 bands = 256 B per grid
// point, so every pass - whatever its direction - and the sphere gather of the projection kernel move whole
// 256-B segments.  Each CTA transforms NB = 16 bands x LPC adjacent lines held in shared memory as [n][NB]
// (batch fastest -> conflict-free), with a two-factor Cooley-Tukey n = R1*R2 (R <= 20, radices 2,3,5,7 and
// their products <= 20 evaluated in registers; n <= 400).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "fft_launch.h"

namespace pawb200 {

// exp(+2 pi i m / R) for R <= FFT_MAXR (filled by the host at start-up)
__constant__ double2 c_small_tw[FFT_MAXR + 1][FFT_MAXR];

__device__ __forceinline__ double2 cmulf(double2 a, double2 b) {
  return make_double2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ double2 cadd(double2 a, double2 b) { return make_double2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ double2 csub(double2 a, double2 b) { return make_double2(a.x - b.x, a.y - b.y); }

__host__ __device__ constexpr int smallest_factor(int r) {
  for (int p = 2; p * p <= r; p++)
    if (r % p == 0) return p;
  return r;
}

// V[k] = sum_j v[j] exp(+2 pi i j k / R), in registers, stride S between elements of v
template <int R, int S>
struct SmallDFT {
  static __device__ __forceinline__ void run(double2* v) {
    constexpr int P = smallest_factor(R);
    if constexpr (R == 1) {
      return;
    } else if constexpr (R == 2) {
      const double2 a = v[0], b = v[S];
      v[0] = cadd(a, b);
      v[S] = csub(a, b);
    } else if constexpr (R == 4) {
      const double2 a = cadd(v[0], v[2 * S]), b = csub(v[0], v[2 * S]);
      const double2 c = cadd(v[S], v[3 * S]), d = csub(v[S], v[3 * S]);
      const double2 id = make_double2(-d.y, d.x);   // +i * d
      v[0] = cadd(a, c);
      v[S] = cadd(b, id);
      v[2 * S] = csub(a, c);
      v[3 * S] = csub(b, id);
    } else if constexpr (P == R) {
      // odd prime: pair up j and R-j
      double2 out[R];
      double2 sp[(R - 1) / 2], sm[(R - 1) / 2];
#pragma unroll
      for (int j = 1; j <= (R - 1) / 2; j++) {
        sp[j - 1] = cadd(v[j * S], v[(R - j) * S]);
        sm[j - 1] = csub(v[j * S], v[(R - j) * S]);
      }
      out[0] = v[0];
#pragma unroll
      for (int j = 0; j < (R - 1) / 2; j++) out[0] = cadd(out[0], sp[j]);
#pragma unroll
      for (int k = 1; k <= (R - 1) / 2; k++) {
        double2 re = v[0], im = make_double2(0, 0);
#pragma unroll
        for (int j = 1; j <= (R - 1) / 2; j++) {
          const double2 w = c_small_tw[R][(j * k) % R];
          re.x += sp[j - 1].x * w.x;
          re.y += sp[j - 1].y * w.x;
          im.x += sm[j - 1].x * w.y;
          im.y += sm[j - 1].y * w.y;
        }
        // X[k] = re + i*im (im is real-weighted difference): i*(a+ib) = -b + i a
        out[k] = make_double2(re.x - im.y, re.y + im.x);
        out[R - k] = make_double2(re.x + im.y, re.y - im.x);
      }
#pragma unroll
      for (int k = 0; k < R; k++) v[k * S] = out[k];
    } else {
      // R = P * Q Cooley-Tukey in registers: j = j1*Q + j2, k = k1 + P*k2
      constexpr int Q = R / P;
      double2 t[R];
#pragma unroll
      for (int j2 = 0; j2 < Q; j2++) {
#pragma unroll
        for (int j1 = 0; j1 < P; j1++) t[j2 * P + j1] = v[(j1 * Q + j2) * S];
        SmallDFT<P, 1>::run(t + j2 * P);            // -> y[k1] at t[j2*P + k1]
#pragma unroll
        for (int k1 = 1; k1 < P; k1++)
          if (j2 > 0) t[j2 * P + k1] = cmulf(t[j2 * P + k1], c_small_tw[R][(j2 * k1) % R]);
      }
      // for each k1: Q-point DFT over j2 (stride P in t)
#pragma unroll
      for (int k1 = 0; k1 < P; k1++) SmallDFT<Q, P>::run(t + k1);   // t[k1 + P*k2] = X[k1 + P*k2]
#pragma unroll
      for (int k = 0; k < R; k++) v[k * S] = t[k];
    }
  }
};

// RMAX (template parameter of the pass kernels) prunes the large-radix bodies so that grids whose factors
// are all <= 10 (e.g. 90 = 9 x 10) compile to a low-register kernel with 6 resident CTAs per SM.
#define PAWB200_RADIX_SWITCH(R, CALL)                                                            \
  switch (R) {                                                                                    \
    case 2: CALL(2); break;   case 3: CALL(3); break;   case 4: CALL(4); break;                   \
    case 5: CALL(5); break;   case 6: CALL(6); break;   case 7: CALL(7); break;                   \
    case 8: CALL(8); break;   case 9: CALL(9); break;   case 10: CALL(10); break;                 \
    default:                                                                                      \
      if constexpr (RMAX > 10) {                                                                  \
        switch (R) {                                                                              \
          case 12: CALL(12); break; case 14: CALL(14); break; case 15: CALL(15); break;           \
          case 16: CALL(16); break;                                                               \
          default:                                                                                \
            if constexpr (RMAX > 16) {                                                            \
              switch (R) {                                                                        \
                case 18: CALL(18); break; case 20: CALL(20); break; default: break;               \
              }                                                                                   \
            }                                                                                     \
            break;                                                                                \
        }                                                                                         \
      }                                                                                           \
      break;                                                                                      \
  }

// Two-factor line transform n = R1*R2 of FFT_B interleaved bands, thread = (q, band):
//   phase 1 (q = j2 < R2): v[j1] = x[j1*R2 + j2]  (fetched by `load(row)`), R1-point DFT, twiddle
//                          w_n^(j2*k1), written to shared buf[(k1*R2 + j2)][band];
//   phase 2 (q = k1 < R1): R2-point DFT over buf[k1*R2 + j2], result X[k1 + R1*k2] handed to `store(row, v)`.
// Inputs come straight from global memory into registers and outputs go straight back, so a line costs one
// shared-memory round trip and one barrier (buf is double buffered by the caller).
template <int R, class Load>
__device__ __forceinline__ void line_phase1(double2* __restrict__ buf, const double2* __restrict__ tw, int R2,
                                            int j2, int b, Load load) {
  double2 v[R];
#pragma unroll
  for (int j1 = 0; j1 < R; j1++) v[j1] = load(j1 * R2 + j2);
  SmallDFT<R, 1>::run(v);
#pragma unroll
  for (int k1 = 0; k1 < R; k1++) {
    double2 x = v[k1];
    if (k1 > 0) x = cmulf(x, tw[j2 * k1]);          // tw[0] = (1, 0) exactly: no per-thread branch for j2 == 0
    buf[(k1 * R2 + j2) * FFT_B + b] = x;
  }
}
template <int R, class Store>
__device__ __forceinline__ void line_phase2(const double2* __restrict__ buf, int R1, int k1, int b, Store store) {
  double2 v[R];
#pragma unroll
  for (int j2 = 0; j2 < R; j2++) v[j2] = buf[(k1 * R + j2) * FFT_B + b];
  SmallDFT<R, 1>::run(v);
#pragma unroll
  for (int k2 = 0; k2 < R; k2++) store(k1 + R1 * k2, v[k2]);
}

// L2 prefetch of a contiguous byte range (multiple of 16 B, 16-B aligned): the passes below ask for the inputs of
// their NEXT work item while they transform the current one, so the demand loads of phase 1 find their lines in L2
// instead of waiting a DRAM round trip with only ~20 resident warps per SM to cover it.
__device__ __forceinline__ void l2_prefetch(const void* p, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ void l2_prefetch_line(const void* p) {      // the 128-byte line holding p
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// phase 2 with the common store pattern out[(k1 + R1 * k2) * stride]: one running pointer instead of R offsets
template <int R>
__device__ __forceinline__ void line_phase2_strided(const double2* __restrict__ buf, int R1, int k1, int b,
                                                    double2* __restrict__ out, int stride) {
  double2 v[R];
#pragma unroll
  for (int j2 = 0; j2 < R; j2++) v[j2] = buf[(k1 * R + j2) * FFT_B + b];
  SmallDFT<R, 1>::run(v);
  double2* po = out + k1 * stride;
  const long inc = (long)R1 * stride;
#pragma unroll
  for (int k2 = 0; k2 < R; k2++) {
    *po = v[k2];
    po += inc;
  }
}

template <int RMAX> struct FftLaunch {
  static constexpr int THREADS = RMAX * FFT_B;          // q-slots x 16 bands
  // resident CTAs per SM the register budget is tuned for (10: 160 thr x 5, 12: 192 x 3, 14: 224 x 2, 16: 256 x 2,
  // 20: 320 x 1 - lines up to 400 points fill the shared memory of an SM on their own)
#ifndef PAWB200_FFT_MINB10
#define PAWB200_FFT_MINB10 5
#endif
  static constexpr int MINB = RMAX <= 10 ? PAWB200_FFT_MINB10 : RMAX <= 12 ? 3 : RMAX <= 16 ? 2 : 1;
};

// ---- pass Z: coefficients -> T1[group][col][z][FFT_B] ---------------------------------------------
// Reads the 16-slot interleaved coefficient copy (interleave_coeff_kernel): thread (q, b) pulls the plane waves
// of slot b that phase 1 needs straight into registers, 16 threads per 128-byte row.  A column's plane waves
// cover one cyclic run of z (the cutoff sphere is convex), so "is z present, and where" is arithmetic on four
// integers per column instead of a staged, zero-filled shared-memory column.
template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_pass_z_kernel(FftGeom g, const float2* __restrict__ Cil, long ldil, int slot0, int nslot, double scale,
                  double2* __restrict__ T1, int ngroups) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* bufs = reinterpret_cast<double2*>(fft_smem);      // [2][n3][FFT_B] exchange buffers
  double2* tw = bufs + 2 * g.n3 * FFT_B;                     // [n3]
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  const int R1 = g.r1[2], R2 = g.r2[2], n3 = g.n3;
  for (int i = tid; i < n3; i += blockDim.x) tw[i] = g.tw[2][i];
  // (group, column) of this CTA's line, advanced incrementally: no 64-bit divisions in the line loop
  const int ncol = g.ncol, step = (int)gridDim.x;
  int grp = (int)blockIdx.x / ncol, col = (int)blockIdx.x % ncol;
  int4 run = grp < ngroups ? __ldg(g.col_run + col) : make_int4(0, 0, 0, 0);
  __syncthreads();
  for (int it = 0; grp < ngroups; it++) {
    const int4 cur = run;
    int ngrp = grp, ncl = col + step;
    while (ncl >= ncol) { ncl -= ncol; ngrp++; }
    const bool more = ngrp < ngroups;
    if (more) run = __ldg(g.col_run + ncl);                    // next column's run, off the critical path
    if (g.pf && tid == 0 && more && run.y > 0) {
      // the next line's plane waves are one contiguous run of 128-byte rows of the interleaved coefficients
      l2_prefetch(Cil + ((long)((slot0 >> 4) + ngrp) * ldil + run.x) * FFT_B, (unsigned)run.y * (FFT_B * sizeof(float2)));
    }
    double2* buf = bufs + (it & 1) * n3 * FFT_B;
    if (q < R2) {
      const int slot = slot0 + grp * FFT_B + b;
      const int cnt = slot < slot0 + nslot ? cur.y : 0;                  // pad slots of the last group stay zero
      const float2* base = Cil + ((long)(slot >> 4) * ldil + cur.x) * FFT_B + (slot & 15);
      auto load = [&](int row) {
        int d = row - cur.z;
        if (d < 0) d += n3;
        if (d >= cnt) return make_double2(0, 0);
        const int j = d < cur.w ? cur.y - cur.w + d : d - cur.w;
        const float2 c = __ldg(base + j * FFT_B);
        return make_double2(scale * (double)c.x, scale * (double)c.y);
      };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
      PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
    }
    __syncthreads();
    if (q < R1) {
      double2* out = T1 + (((long)grp * g.ncol + col) * n3) * FFT_B + b;
#define P2(R) line_phase2_strided<R>(buf, R1, q, b, out, FFT_B)
      PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
    }
    grp = ngrp;
    col = ncl;
  }
}

// ---- pass Z, staged variant (fallback when a column's plane waves are not one cyclic z-run) --------------
// slot -> coefficient row like scatter_pw_kernel: band = slot / halves, half = slot % halves.
// The sparse column (z-run of plane waves) is staged through shared memory with coalesced reads.
template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_pass_z_staged_kernel(FftGeom g, const float2* __restrict__ C, long ldc, int halves, int half_len, int slot0,
                  int nslot, double scale, double2* __restrict__ T1, int ngroups) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* in = reinterpret_cast<double2*>(fft_smem);        // [n3][FFT_B] sparse input column
  double2* buf = in + g.n3 * FFT_B;                          // [n3][FFT_B] exchange buffer
  double2* tw = buf + g.n3 * FFT_B;                          // [n3]
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  const int warp = tid >> 5, lane = tid & 31, nwarp = blockDim.x >> 5;
  const int R1 = g.r1[2], R2 = g.r2[2];
  for (int i = tid; i < g.n3; i += blockDim.x) tw[i] = g.tw[2][i];
  const long nlines = (long)ngroups * g.ncol;
  for (long line = blockIdx.x; line < nlines; line += gridDim.x) {
    const int grp = (int)(line / g.ncol), col = (int)(line % g.ncol);
    __syncthreads();                                         // previous line is done with `in` and `buf`
    for (int i = tid; i < g.n3 * FFT_B; i += blockDim.x) in[i] = make_double2(0, 0);
    __syncthreads();
    const int s = g.col_start[col], cnt = g.col_cnt[col];
    for (int bb = warp; bb < FFT_B; bb += nwarp) {           // warp <-> band, lanes <-> consecutive plane waves
      const int slot = slot0 + grp * FFT_B + bb;
      if (slot >= slot0 + nslot) continue;                   // pad bands of the last group stay zero
      const float2* row = C + (long)(slot / halves) * ldc + (long)(slot % halves) * half_len;
      for (int j = lane; j < cnt; j += 32) {
        const float2 c = __ldg(row + s + j);
        in[g.zpos[s + j] * FFT_B + bb] = make_double2(scale * (double)c.x, scale * (double)c.y);
      }
    }
    __syncthreads();
    if (q < R2) {
      auto load = [&](int row) { return in[row * FFT_B + b]; };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
      PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
    }
    __syncthreads();
    if (q < R1) {
      double2* out = T1 + (((long)grp * g.ncol + col) * g.n3) * FFT_B + b;
      auto store = [&](int row, double2 v) { out[(long)row * FFT_B] = v; };
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
      PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
    }
  }
}

// ---- pass Y: T1 -> T2[group][plane][y][z][FFT_B] ------------------------------------------------------
// Work unit = (group, plane, chunk of FFT_ZC z-lines): the plane's row->column table is fetched into shared
// memory once per unit, so the data loads do not wait behind a dependent global lookup.
template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_pass_y_kernel(FftGeom g, const double2* __restrict__ T1, double2* __restrict__ T2, int ngroups) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* bufs = reinterpret_cast<double2*>(fft_smem);      // [2][n2][FFT_B]
  double2* tw = bufs + 2 * g.n2 * FFT_B;
  int* ssrc = reinterpret_cast<int*>(tw + g.n2);             // [n2] column holding each y row of the plane
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  const int R1 = g.r1[1], R2 = g.r2[1];
  for (int i = tid; i < g.n2; i += blockDim.x) tw[i] = g.tw[1][i];
  const int nzc = (g.n3 + FFT_ZC - 1) / FFT_ZC;
  const int nunits = ngroups * g.nplane * nzc;                // 32-bit: <= 8 groups x 400 planes x 45 chunks
  int it = 0;
  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const int zc = unit % nzc;
    const int p = (unit / nzc) % g.nplane;
    const int grp = unit / (nzc * g.nplane);
    __syncthreads();                                         // previous unit no longer reads ssrc
    // element offset of the column holding each y row, premultiplied (32-bit: ncol * n3 * 16 < 2^31 for n <= 400)
    const int colstride = g.n3 * FFT_B;
    for (int i = tid; i < g.n2; i += blockDim.x) {
      const int c = g.ysrc[p * g.n2 + i];
      ssrc[i] = c >= 0 ? c * colstride : -1;
    }
    // first line of the next unit of this CTA (one line ahead only: a whole unit ahead is 80 MB in flight GPU-wide and
    // was evicted before use - ncu showed pass Y reading T1 twice)
    long nu_base = -1;
    int nu_plane = 0;
    if (g.pf && unit + (int)gridDim.x < nunits) {
      const int nu = unit + (int)gridDim.x;
      nu_plane = (nu / nzc) % g.nplane;
      nu_base = ((long)(nu / (nzc * g.nplane)) * g.ncol * g.n3 + (nu % nzc) * FFT_ZC) * FFT_B;
    }
    __syncthreads();
    const int z1 = min(g.n3, (zc + 1) * FFT_ZC);
    for (int z = zc * FFT_ZC; z < z1; z++, it++) {
      double2* buf = bufs + (it & 1) * g.n2 * FFT_B;
      if (g.pf) {               // inputs of the next line: z + 1 of this unit, else the first line of the next unit
        const bool last = z + 1 >= z1;
        if (!last || nu_base >= 0) {
          const double2* nin = last ? T1 + nu_base : T1 + ((long)grp * g.ncol * g.n3 + z + 1) * FFT_B;
          for (int i = tid; i < 2 * g.n2; i += blockDim.x) {
            const int o = last ? __ldg(g.ysrc + nu_plane * g.n2 + (i >> 1)) * colstride : ssrc[i >> 1];
            if (o >= 0) l2_prefetch_line(reinterpret_cast<const char*>(nin + o) + (i & 1) * 128);
          }
        }
      }
      if (q < R2) {
        const double2* in = T1 + ((long)grp * g.ncol * g.n3 + z) * FFT_B + b;
        auto load = [&](int row) {
          const int o = ssrc[row];
          return o >= 0 ? in[o] : make_double2(0, 0);
        };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
        PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
      }
      __syncthreads();
      if (q < R1) {
        double2* out = T2 + ((((long)grp * g.nplane + p) * g.n2) * g.n3 + z) * FFT_B + b;
#define P2(R) line_phase2_strided<R>(buf, R1, q, b, out, colstride)
        PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
      }
    }
  }
}

// ---- pass X: T2 -> X[group][x][y][z][FFT_B] --------------------------------------------------------------
template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_pass_x_kernel(FftGeom g, const double2* __restrict__ T2, double2* __restrict__ X, int ngroups) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* bufs = reinterpret_cast<double2*>(fft_smem);      // [2][n1][FFT_B]
  double2* tw = bufs + 2 * g.n1 * FFT_B;
  int* sxsrc = reinterpret_cast<int*>(tw + g.n1);            // [n1] plane holding each x row (or -1)
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  const int R1 = g.r1[0], R2 = g.r2[0];
  const long plane = (long)g.n2 * g.n3;
  for (int i = tid; i < g.n1; i += blockDim.x) {
    tw[i] = g.tw[0][i];
    const int p = g.xsrc[i];
    sxsrc[i] = p >= 0 ? p * (int)(plane * FFT_B) : -1;     // premultiplied element offset (nplane * plane * 16 < 2^31)
  }
  __syncthreads();
  const int iplane = (int)plane, step = (int)gridDim.x;
  int grp = (int)blockIdx.x / iplane, yz = (int)blockIdx.x % iplane;       // advanced incrementally: no divisions per line
  for (int it = 0; grp < ngroups; it++) {
    double2* buf = bufs + (it & 1) * g.n1 * FFT_B;
    if (q < R2) {
      const double2* in = T2 + ((long)grp * g.nplane * plane + yz) * FFT_B + b;
      auto load = [&](int row) {
        const int o = sxsrc[row];
        return o >= 0 ? in[o] : make_double2(0, 0);
      };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
      PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
    }
    __syncthreads();
    // next line of this CTA; its inputs are requested now (phase 1's registers are dead, phase 2 covers the latency)
    int ngrp = grp, nyz = yz + step;
    while (nyz >= iplane) { nyz -= iplane; ngrp++; }
    if (g.pf && ngrp < ngroups) {
      const double2* nin = T2 + ((long)ngrp * g.nplane * plane + nyz) * FFT_B;
      for (int i = tid; i < 2 * g.nplane; i += blockDim.x)
        l2_prefetch_line(reinterpret_cast<const char*>(nin + (i >> 1) * (int)(plane * FFT_B)) + (i & 1) * 128);
    }
    if (q < R1) {
      double2* out = X + ((long)grp * g.n1 * plane + yz) * FFT_B + b;
      const int xstride = (int)(plane * FFT_B);                 // n1 * plane * 16 < 2^31 for n <= 400
#define P2(R) line_phase2_strided<R>(buf, R1, q, b, out, xstride)
      PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
    }
    grp = ngrp;
    yz = nyz;
  }
}

// ---- three-factor lines: n = R1 * R2 * R3 (radices <= 10) -----------------------------------------------------
// Line lengths without a two-factor split into radices <= 20 (175 = 5*5*7, 243 = 3*9*9, 384 = 6*8*8, ...):
//   j = j1*M + j2*R3 + j3 (M = R2*R3),   k = k1 + R1*(k2 + R2*k3)
//   phase 1, task m < M:        R1-point DFT over j1, twiddle w_n^(m k1)        -> A[k1*M + m]
//   phase 2, task (k1, j3):     R2-point DFT over j2, twiddle w_n^(R1 j3 k2)    -> B[k1*M + k2*R3 + j3]
//   phase 3, task (k1, k2):     R3-point DFT over j3                            -> X[k1 + R1*(k2 + R2*k3)]
// Thread = (q, band) as in the two-factor passes; a thread loops over the tasks q, q + Q, ... of a phase.  Two
// exchange buffers A and B and two barriers per line: the next line's phase 1 only writes A, which nobody reads
// after the second barrier, and its phase 2 (the next writer of B) comes after a barrier every thread reaches only
// when it has finished reading B.
constexpr int FFT3_RMAX = 10;

template <class Load, class Store>
__device__ __forceinline__ void line3_transform(double2* __restrict__ A, double2* __restrict__ B,
                                                const double2* __restrict__ tw, int R1, int R2, int R3, int q, int Q,
                                                int b, Load load, Store store) {
  constexpr int RMAX = FFT3_RMAX;
  const int M = R2 * R3;
  for (int m = q; m < M; m += Q) {
#define T1(R)                                                        \
  {                                                                  \
    double2 v[R];                                                    \
    _Pragma("unroll") for (int j1 = 0; j1 < R; j1++) v[j1] = load(j1 * M + m); \
    SmallDFT<R, 1>::run(v);                                          \
    _Pragma("unroll") for (int k1 = 0; k1 < R; k1++) {               \
      double2 x = v[k1];                                             \
      if (k1 > 0) x = cmulf(x, tw[m * k1]);                          \
      A[(k1 * M + m) * FFT_B + b] = x;                               \
    }                                                                \
  }
    PAWB200_RADIX_SWITCH(R1, T1)
#undef T1
  }
  __syncthreads();
  for (int t = q; t < R1 * R3; t += Q) {
    const int k1 = t / R3, j3 = t - k1 * R3;
#define T2(R)                                                        \
  {                                                                  \
    double2 v[R];                                                    \
    _Pragma("unroll") for (int j2 = 0; j2 < R; j2++) v[j2] = A[(k1 * M + j2 * R3 + j3) * FFT_B + b]; \
    SmallDFT<R, 1>::run(v);                                          \
    _Pragma("unroll") for (int k2 = 0; k2 < R; k2++) {               \
      double2 x = v[k2];                                             \
      if (k2 > 0) x = cmulf(x, tw[R1 * j3 * k2]);                    \
      B[(k1 * M + k2 * R3 + j3) * FFT_B + b] = x;                    \
    }                                                                \
  }
    PAWB200_RADIX_SWITCH(R2, T2)
#undef T2
  }
  __syncthreads();
  for (int t = q; t < R1 * R2; t += Q) {
    const int k1 = t / R2, k2 = t - k1 * R2;
#define T3(R)                                                        \
  {                                                                  \
    double2 v[R];                                                    \
    _Pragma("unroll") for (int j3 = 0; j3 < R; j3++) v[j3] = B[(k1 * M + k2 * R + j3) * FFT_B + b]; \
    SmallDFT<R, 1>::run(v);                                          \
    _Pragma("unroll") for (int k3 = 0; k3 < R; k3++) store(k1 + R1 * (k2 + R2 * k3), v[k3]); \
  }
    PAWB200_RADIX_SWITCH(R3, T3)
#undef T3
  }
}

// PASS: 2 = Z (coefficients -> T1), 1 = Y (T1 -> T2), 0 = X (T2 -> X); same work decomposition, inputs and outputs as
// the two-factor kernels above, one line per iteration.
template <int PASS>
__global__ void __launch_bounds__(512, 1)
fft3_pass_kernel(FftGeom g, const float2* __restrict__ Cil, long ldil, int slot0, int nslot, double scale,
                 const double2* __restrict__ in_base, double2* __restrict__ out_base, int ngroups) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  const int n = PASS == 0 ? g.n1 : PASS == 1 ? g.n2 : g.n3;
  double2* A = reinterpret_cast<double2*>(fft_smem);         // [n][FFT_B]
  double2* B = A + n * FFT_B;                                // [n][FFT_B]
  double2* tw = B + n * FFT_B;                               // [n]
  int* ssrc = reinterpret_cast<int*>(tw + n);                // X: [n1] plane offsets; Y: [n2] column offsets
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B, Q = blockDim.x / FFT_B;
  const int R1 = g.r1[PASS], R2 = g.r2[PASS], R3 = g.r3[PASS];
  for (int i = tid; i < n; i += blockDim.x) tw[i] = g.tw[PASS][i];
  const long plane = (long)g.n2 * g.n3;
  if (PASS == 0)
    for (int i = tid; i < g.n1; i += blockDim.x) {
      const int p = g.xsrc[i];
      ssrc[i] = p >= 0 ? p * (int)(plane * FFT_B) : -1;
    }
  __syncthreads();
  if (PASS == 2) {
    const long nlines = (long)ngroups * g.ncol;
    for (long line = blockIdx.x; line < nlines; line += gridDim.x) {
      const int grp = (int)(line / g.ncol), col = (int)(line % g.ncol);
      const int4 cur = __ldg(g.col_run + col);
      const int slot = slot0 + grp * FFT_B + b;
      const int cnt = slot < slot0 + nslot ? cur.y : 0;
      const float2* base = Cil + ((long)(slot >> 4) * ldil + cur.x) * FFT_B + (slot & 15);
      const int n3 = g.n3;
      auto load = [&](int row) {
        int d = row - cur.z;
        if (d < 0) d += n3;
        if (d >= cnt) return make_double2(0, 0);
        const int j = d < cur.w ? cur.y - cur.w + d : d - cur.w;
        const float2 c = __ldg(base + j * FFT_B);
        return make_double2(scale * (double)c.x, scale * (double)c.y);
      };
      double2* out = out_base + (((long)grp * g.ncol + col) * n3) * FFT_B + b;
      auto store = [&](int row, double2 v) { out[row * FFT_B] = v; };
      line3_transform(A, B, tw, R1, R2, R3, q, Q, b, load, store);
    }
  } else if (PASS == 1) {
    const int colstride = g.n3 * FFT_B;
    const long nlines = (long)ngroups * g.nplane * g.n3;
    int cur_plane = -1;
    for (long line = blockIdx.x; line < nlines; line += gridDim.x) {
      const int z = (int)(line % g.n3);
      const int p = (int)((line / g.n3) % g.nplane);
      const int grp = (int)(line / ((long)g.n3 * g.nplane));
      if (p != cur_plane) {                                  // (uniform) refresh the plane's row -> column offsets
        __syncthreads();
        for (int i = tid; i < g.n2; i += blockDim.x) {
          const int c = g.ysrc[p * g.n2 + i];
          ssrc[i] = c >= 0 ? c * colstride : -1;
        }
        cur_plane = p;
        __syncthreads();
      }
      const double2* in = in_base + ((long)grp * g.ncol * g.n3 + z) * FFT_B + b;
      auto load = [&](int row) {
        const int o = ssrc[row];
        return o >= 0 ? in[o] : make_double2(0, 0);
      };
      double2* out = out_base + ((((long)grp * g.nplane + p) * g.n2) * g.n3 + z) * FFT_B + b;
      auto store = [&](int row, double2 v) { out[row * colstride] = v; };
      line3_transform(A, B, tw, R1, R2, R3, q, Q, b, load, store);
    }
  } else {
    const int xstride = (int)(plane * FFT_B);
    const long nlines = (long)ngroups * plane;
    for (long line = blockIdx.x; line < nlines; line += gridDim.x) {
      const long yz = line % plane;
      const int grp = (int)(line / plane);
      const double2* in = in_base + ((long)grp * g.nplane * plane + yz) * FFT_B + b;
      auto load = [&](int row) {
        const int o = ssrc[row];
        return o >= 0 ? in[o] : make_double2(0, 0);
      };
      double2* out = out_base + ((long)grp * g.n1 * plane + yz) * FFT_B + b;
      auto store = [&](int row, double2 v) { out[row * xstride] = v; };
      line3_transform(A, B, tw, R1, R2, R3, q, Q, b, load, store);
    }
  }
}

// ---- fused pass Y + X: T1 -> (L2-resident ring) -> X ----------------------------------------------------------
// The stand-alone passes write the y-transformed planes T2 (0.68 N points per band) to HBM and read them back:
// 45 % of the transform's traffic.  Here both passes run in ONE persistent kernel and T2 only ever exists for a few
// z values at a time: a "chunk" c = (band group, zch consecutive z) owns slot c % ring of a ring buffer of a few tens
// of MB that is written and re-read within microseconds, i.e. it stays in the 126 MB L2 and never reaches DRAM.
//
// Work items are single lines, handed out IN ORDER through a global ticket counter (atomicAdd), so every item an
// item depends on has a smaller ticket and is held by a running CTA - the waits below cannot deadlock, whatever the
// residency of the grid.  Ticket order, step s = 0, 1, ...:  the X lines of chunk s - lead, then the Y lines of chunk
// s.  A Y line of chunk c waits until the X lines of chunk c - ring have consumed the slot; an X line of chunk c waits
// for all Y lines of c.  `lead` steps (more items than there are CTAs) separate producer and consumer, so in steady
// state the waits are satisfied on the first poll.  Completion is signalled per warp (syncwarp, fence, one atomic);
// consumers poll with ld.acquire and read the ring with ld.global.cg (L2) - slots are recycled, L1 may be stale.
struct YxArgs {
  int zch, nzc, ring, lead, nchunks;
  unsigned* ticket;         // next item
  unsigned* ydone;          // [nchunks] warps that finished a Y line of the chunk
  unsigned* xdone;          // [nchunks] warps that finished an X line of the chunk
};

// Polls are relaxed loads: an acquire load invalidates the whole L1 of the SM (CCTL.IVALL) on every poll, and the
// only data ordered behind a poll - the ring - is read with ld.global.cg, which never looks at L1.
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_count(const unsigned* p, unsigned target) {
  while (ld_acquire_u32(p) < target) __nanosleep(64);
}

// decoded work item (thread 0 decodes the ticket once - the integer divisions would otherwise be repeated by every
// thread - and publishes it through shared memory together with the ticket prefetch)
struct YxItem {
  int kind;                 // 0: past the end, 1: no-op (void / padded), 2: Y line, 3: X line; bit 2 set: counts on completion
  int c;                    // chunk
  int row;                  // X: y;  Y: plane
  int zz, z, grp, slot;
  int dep;                  // index of the counter this item waits on (into ydone/xdone), -1: none
  unsigned target;          // value the counter must reach
  int done;                 // index of the counter it increments on completion, -1: none
  int pad[2];
};

template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_pass_yx_kernel(FftGeom g, YxArgs a, const double2* __restrict__ T1, double2* T2c, double2* __restrict__ X) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  const int nmax = max(g.n1, g.n2);
  double2* bufs = reinterpret_cast<double2*>(fft_smem);      // [2][nmax][FFT_B] exchange buffers
  double2* twy = bufs + 2 * nmax * FFT_B;                    // [n2]
  double2* twx = twy + g.n2;                                 // [n1]
  int4* srun = reinterpret_cast<int4*>(twx + g.n1);          // [nplane] y-run of every active x-plane
  YxItem* sitem = reinterpret_cast<YxItem*>(srun + g.nplane);   // [2]
  int* sxsrc = reinterpret_cast<int*>(sitem + 2);            // [n1] plane holding each x row (or -1)
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B, lane = tid & 31;
  const unsigned nwarps = (blockDim.x + 31) >> 5;             // the last warp may be half full (16-thread q slots)
  for (int i = tid; i < g.n2; i += blockDim.x) twy[i] = g.tw[1][i];
  for (int i = tid; i < g.n1; i += blockDim.x) {
    twx[i] = g.tw[0][i];
    sxsrc[i] = g.xsrc[i];
  }
  for (int i = tid; i < g.nplane; i += blockDim.x) srun[i] = g.plane_run[i];
  const unsigned items_x = (unsigned)g.n2 * a.zch, items_y = (unsigned)g.nplane * a.zch;
  const unsigned per_step = items_x + items_y;
  const unsigned total = (unsigned)(a.nchunks + a.lead) * per_step;
  unsigned* counters = a.ydone;                               // ydone[nchunks] followed by xdone[nchunks]
  auto fetch = [&](YxItem* it) {                              // thread 0 only
    const unsigned t = atomicAdd(a.ticket, 1u);
    YxItem d;
    d.kind = 0; d.c = 0; d.row = 0; d.zz = 0; d.z = 0; d.grp = 0; d.slot = 0; d.dep = -1; d.target = 0; d.done = -1;
    if (t < total) {
      const int step = (int)(t / per_step);
      unsigned r = t - (unsigned)step * per_step;
      const bool isx = r < items_x;
      const int c = isx ? step - a.lead : step;
      const bool valid = isx ? (c >= 0) : (c < a.nchunks);
      if (!isx) r -= items_x;
      d.kind = 1;
      if (valid) {
        d.c = c;
        d.row = (int)(r / a.zch);
        d.zz = (int)(r - (unsigned)d.row * a.zch);
        d.grp = c / a.nzc;
        d.z = (c - d.grp * a.nzc) * a.zch + d.zz;
        d.slot = c % a.ring;
        d.done = isx ? a.nchunks + c : c;
        if (d.z < g.n3) {
          d.kind = isx ? 3 : 2;
          if (isx) { d.dep = c; d.target = items_y * nwarps; }
          else if (c >= a.ring) { d.dep = a.nchunks + c - a.ring; d.target = items_x * nwarps; }
        }
      }
    }
    *it = d;
  };
  if (tid == 0) fetch(sitem);
  __syncthreads();
  const long plane = (long)g.n2 * g.n3;
  const long slot_elems = (long)g.nplane * g.n2 * a.zch * FFT_B;
  YxItem cur = sitem[0];
  // lane 0 of every warp: value of the current item's dependency counter, sampled during the previous item
  unsigned seen = (lane == 0 && cur.dep >= 0) ? ld_acquire_u32(counters + cur.dep) : 0u;
  int pending = -1;                 // completion of the previous item, not yet published
  auto publish = [&]() {
    if (lane == 0 && pending >= 0) red_release_add(counters + pending, 1u);
    pending = -1;
  };
  for (int it = 0; cur.kind != 0; it++) {
    if (tid == 0) fetch(sitem + ((it + 1) & 1));               // next item, off the critical path
    const bool isx = cur.kind == 3, work = cur.kind >= 2;
    double2* buf = bufs + (it & 1) * nmax * FFT_B;
    double2* ring = T2c + (long)cur.slot * slot_elems;
    if (work) {
      if (lane == 0 && cur.dep >= 0 && seen < cur.target) {
        publish();                 // never wait while holding an unpublished completion
        wait_count(counters + cur.dep, cur.target);
      }
      __syncwarp();
      if (isx) {
        if (q < g.r2[0]) {
          const double2* in = ring + ((long)cur.row * a.zch + cur.zz) * FFT_B + b;
          const long pstride = (long)g.n2 * a.zch * FFT_B;
          auto load = [&](int x) {
            const int p = sxsrc[x];
            return p >= 0 ? __ldcg(in + (long)p * pstride) : make_double2(0, 0);
          };
          const int R2 = g.r2[0];
#define P1(R) line_phase1<R>(buf, twx, R2, q, b, load)
          PAWB200_RADIX_SWITCH(g.r1[0], P1)
#undef P1
        }
      } else {
        if (q < g.r2[1]) {
          const int4 run = srun[cur.row];
          const double2* in = T1 + ((long)cur.grp * g.ncol * g.n3 + cur.z) * FFT_B + b;
          const long cstride = (long)g.n3 * FFT_B;
          const int n2 = g.n2;
          auto load = [&](int y) {
            int d = y - run.z;
            if (d < 0) d += n2;
            if (d >= run.y) return make_double2(0, 0);
            const int col = run.x + (d < run.w ? run.y - run.w + d : d - run.w);
            return __ldg(in + (long)col * cstride);
          };
          const int R2 = g.r2[1];
#define P1(R) line_phase1<R>(buf, twy, R2, q, b, load)
          PAWB200_RADIX_SWITCH(g.r1[1], P1)
#undef P1
        }
      }
    }
    __syncthreads();
    const YxItem nxt = sitem[(it + 1) & 1];
    // The previous item's stores were issued a whole phase ago: publishing them now costs a short fence.  The
    // next item's dependency counter is sampled here too, so that its value is in a register when the item starts.
    publish();
    seen = (lane == 0 && nxt.dep >= 0) ? ld_acquire_u32(counters + nxt.dep) : 0u;
    if (work) {
      if (isx) {
        if (q < g.r1[0]) {
          double2* out = X + ((long)cur.grp * g.n1 * plane + (long)cur.row * g.n3 + cur.z) * FFT_B + b;
          auto store = [&](int x, double2 v) { out[(long)x * plane * FFT_B] = v; };
          const int R1 = g.r1[0];
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
          PAWB200_RADIX_SWITCH(g.r2[0], P2)
#undef P2
        }
      } else {
        if (q < g.r1[1]) {
          double2* out = ring + ((long)cur.row * g.n2 * a.zch + cur.zz) * FFT_B + b;
          const long ystride = (long)a.zch * FFT_B;
          auto store = [&](int y, double2 v) { __stcg(out + (long)y * ystride, v); };
          const int R1 = g.r1[1];
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
          PAWB200_RADIX_SWITCH(g.r2[1], P2)
#undef P2
        }
      }
    }
    if (cur.done >= 0) {        // padded z (z >= n3) items only count
      __syncwarp();
      pending = cur.done;
    }
    cur = nxt;
  }
  publish();
}

// ---- fused pass Y + X, phased variant (PAWB200_FFT_FUSED=2) --------------------------------------------------------
// Same idea as fft_pass_yx_kernel - T2 only ever exists as a few chunks in an L2-resident ring - but without any
// per-line synchronisation.  The items of the whole launch form one static list, phase p = [Y lines of chunk p,
// X lines of chunk p - 1], and CTA i simply takes items i, i + G, i + 2G, ...  A CTA talks to the others only
// when its item moves on to another (chunk, kind) group: it publishes how many items of the old group it completed
// (one barrier, one fence, one atomic) and checks that the new group's dependency - all Y lines of the chunk for
// an X line, all X lines of the chunk that used the ring slot before for a Y line (ring = 3 slots) - is complete.
// Both dependencies were produced at least LX (or LY) items earlier in the list, so in steady state the check is a
// single load.  All CTAs must be co-resident (cooperative launch): a waiting CTA depends on CTAs with SMALLER
// and LARGER indices alike.
struct Yx2Args {
  int zch, nzc, ring, nchunks;
  unsigned* ydone;          // [nchunks] Y items of the chunk that are complete
  unsigned* xdone;          // [nchunks]
};

template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_pass_yx2_kernel(FftGeom g, Yx2Args a, const double2* __restrict__ T1, double2* T2c, double2* __restrict__ X) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  const int nmax = max(g.n1, g.n2);
  double2* bufs = reinterpret_cast<double2*>(fft_smem);      // [2][nmax][FFT_B] exchange buffers
  double2* twy = bufs + 2 * nmax * FFT_B;                    // [n2]
  double2* twx = twy + g.n2;                                 // [n1]
  int4* srun = reinterpret_cast<int4*>(twx + g.n1);          // [nplane] y-run of every active x-plane
  int* sxsrc = reinterpret_cast<int*>(srun + g.nplane);      // [n1] plane holding each x row (or -1)
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  for (int i = tid; i < g.n2; i += blockDim.x) twy[i] = g.tw[1][i];
  for (int i = tid; i < g.n1; i += blockDim.x) {
    twx[i] = g.tw[0][i];
    sxsrc[i] = g.xsrc[i];
  }
  for (int i = tid; i < g.nplane; i += blockDim.x) srun[i] = g.plane_run[i];
  __syncthreads();
  const unsigned LY = (unsigned)g.nplane * a.zch, LX = (unsigned)g.n2 * a.zch, per = LY + LX;
  const long plane = (long)g.n2 * g.n3;
  const long slot_elems = (long)g.nplane * g.n2 * a.zch * FFT_B;
  const long cstride = (long)g.n3 * FFT_B;                   // T1: column stride
  const long pstride = (long)g.n2 * a.zch * FFT_B;           // ring: plane stride
  const long ystride = (long)a.zch * FFT_B;                  // ring: y stride
  unsigned phase = 0, r = blockIdx.x;
  while (r >= per) { r -= per; phase++; }
  int key = -1;               // (chunk, kind) group of the items in progress: 2 * chunk + kind
  unsigned cnt = 0;           // items of that group this CTA has completed
  int it = 0;
  auto publish = [&]() {
    if (key < 0) return;
    __syncthreads();          // every thread's stores of the group are issued
    if (tid == 0) {
      __threadfence();
      atomicAdd(((key & 1) ? a.xdone : a.ydone) + (key >> 1), cnt);
    }
  };
  while (phase <= (unsigned)a.nchunks) {
    const bool isx = r >= LY;
    const int c = isx ? (int)phase - 1 : (int)phase;
    const bool valid = isx ? phase >= 1 : phase < (unsigned)a.nchunks;
    // the next item of this CTA (for the L2 prefetch of a Y line's inputs)
    unsigned nphase = phase, nr = r + gridDim.x;
    while (nr >= per) { nr -= per; nphase++; }
    if (valid) {
      const int k = 2 * c + (isx ? 1 : 0);
      if (k != key) {
        publish();
        key = k;
        cnt = 0;
        if (tid == 0) {
          const unsigned* dep = nullptr;
          unsigned target = 0;
          if (isx) { dep = a.ydone + c; target = LY; }
          else if (c >= a.ring) { dep = a.xdone + (c - a.ring); target = LX; }
          if (dep) {
            while (ld_acquire_u32(dep) < target) __nanosleep(40);
            __threadfence();
          }
        }
        __syncthreads();
      }
      cnt++;
      const unsigned rr = isx ? r - LY : r;
      const int row = (int)(rr / (unsigned)a.zch), zz = (int)(rr - (unsigned)row * a.zch);
      const int grp = c / a.nzc;
      const int z = (c - grp * a.nzc) * a.zch + zz;
      if (z < g.n3) {
        double2* buf = bufs + (it & 1) * nmax * FFT_B;
        it++;
        double2* ring = T2c + (long)(c % a.ring) * slot_elems;
        if (g.pf && nphase < (unsigned)a.nchunks && nr < LY) {
          // inputs of this CTA's next Y line: one 256-byte segment per active column of its plane
          const int nrow = (int)(nr / (unsigned)a.zch), nzz = (int)(nr - (unsigned)nrow * a.zch);
          const int ngrp = (int)nphase / a.nzc;
          const int nz = ((int)nphase - ngrp * a.nzc) * a.zch + nzz;
          const int4 nrun = srun[nrow];
          if (nz < g.n3 && tid < nrun.y) {
            const char* pp = reinterpret_cast<const char*>(T1 + (((long)ngrp * g.ncol + nrun.x + tid) * g.n3 + nz) * FFT_B);
            l2_prefetch_line(pp);
            l2_prefetch_line(pp + 128);
          }
        }
        if (isx) {
          if (q < g.r2[0]) {
            const double2* in = ring + ((long)row * a.zch + zz) * FFT_B + b;
            auto load = [&](int x) {
              const int p = sxsrc[x];
              return p >= 0 ? __ldcg(in + (long)p * pstride) : make_double2(0, 0);
            };
            const int R2 = g.r2[0];
#define P1(R) line_phase1<R>(buf, twx, R2, q, b, load)
            PAWB200_RADIX_SWITCH(g.r1[0], P1)
#undef P1
          }
        } else {
          if (q < g.r2[1]) {
            const int4 run = srun[row];
            const double2* in = T1 + ((long)grp * g.ncol * g.n3 + z) * FFT_B + b;
            const int n2 = g.n2;
            auto load = [&](int y) {
              int d = y - run.z;
              if (d < 0) d += n2;
              if (d >= run.y) return make_double2(0, 0);
              const int col = run.x + (d < run.w ? run.y - run.w + d : d - run.w);
              return __ldg(in + (long)col * cstride);
            };
            const int R2 = g.r2[1];
#define P1(R) line_phase1<R>(buf, twy, R2, q, b, load)
            PAWB200_RADIX_SWITCH(g.r1[1], P1)
#undef P1
          }
        }
        __syncthreads();
        if (isx) {
          if (q < g.r1[0]) {
            double2* out = X + ((long)grp * g.n1 * plane + (long)row * g.n3 + z) * FFT_B + b;
            auto store = [&](int x, double2 v) { out[(long)x * plane * FFT_B] = v; };
            const int R1 = g.r1[0];
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
            PAWB200_RADIX_SWITCH(g.r2[0], P2)
#undef P2
          }
        } else {
          if (q < g.r1[1]) {
            double2* out = ring + ((long)row * g.n2 * a.zch + zz) * FFT_B + b;
            auto store = [&](int y, double2 v) { __stcg(out + (long)y * ystride, v); };
            const int R1 = g.r1[1];
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
            PAWB200_RADIX_SWITCH(g.r2[1], P2)
#undef P2
          }
        }
      }
    }
    phase = nphase;
    r = nr;
  }
  publish();
}

// ---- TMA-fed passes Y and X -----------------------------------------------------------------------------------
// The register-direct passes above keep only as many bytes in flight as their ~20 resident warps per SM have loads
// outstanding (ncu: long_scoreboard-bound at 47-60 % of DRAM peak).  Here the inputs of a whole work unit - ZL
// lines - are fetched by the TMA engine (cp.async.bulk.tensor, one elected thread, completion on an mbarrier) into a
// ring of shared-memory stages, several units ahead of the arithmetic, so the bytes in flight no longer depend on
// how many warps are waiting:
//   pass X unit = ZL consecutive (y,z) lines of a group: ONE 3-D box {256 B, ZL, nplane} of T2[grp][plane][yz][16];
//   pass Y unit = (group, plane, ZL consecutive z): boxes {256 B, ZL, 16 columns} of T1[grp][col][z][16] covering the
//                 plane's column run (the tail box over-fetches the next plane's first columns; they are ignored).
// A stage is laid out like the box, [row][ZL][16] (row = plane / column).  The compute side is the unchanged
// two-phase line transform (thread = (radix slot, band)); phase 1 reads the stage instead of global memory, phase 2
// stores straight to global.  Stage reuse needs no second barrier array: a stage is refilled only after the
// __syncthreads that ends the phase-1 reads of its last line.
struct TmaPassArgs {
  int zl;                   // lines per unit
  int stages;               // ring depth
  int stage_rows;           // rows (planes / padded columns) a stage holds
  int dbl;                  // exchange buffer double-buffered (one barrier per line) or single (two)
};

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int c0, int c1, int c2, void* bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}

constexpr int FFT_TMA_COLBOX = 16;     // columns per box of pass Y

// PASS: 0 = X (T2 -> X), 1 = Y (T1 -> T2)
template <int RMAX, int PASS>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_pass_tma_kernel(const __grid_constant__ CUtensorMap tmap, FftGeom g, TmaPassArgs a, double2* __restrict__ out_base,
                    int ngroups) {
  extern __shared__ __align__(128) unsigned char fft_smem[];
  const int n = PASS == 0 ? g.n1 : g.n2;
  const int R1 = g.r1[PASS == 0 ? 0 : 1], R2 = g.r2[PASS == 0 ? 0 : 1];
  const int zl = a.zl, S = a.stages;
  const int stage_elems = a.stage_rows * zl * FFT_B;                     // double2 per stage
  double2* stage0 = reinterpret_cast<double2*>(fft_smem);                // [S][stage_rows][zl][16]
  double2* bufs = stage0 + (long)S * stage_elems;                        // [1 or 2][n][16] exchange
  double2* tw = bufs + (a.dbl ? 2 : 1) * n * FFT_B;                      // [n]
  int4* srun = reinterpret_cast<int4*>(tw + n);                          // Y: [nplane] runs
  int* ssrc = reinterpret_cast<int*>(srun + (PASS == 1 ? g.nplane : 0));  // X: [n1] plane of each x row
  unsigned long long* full = reinterpret_cast<unsigned long long*>(
      (reinterpret_cast<uintptr_t>(ssrc + (PASS == 0 ? g.n1 : 0)) + 7) & ~(uintptr_t)7);   // [S] mbarriers
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  for (int i = tid; i < n; i += blockDim.x) tw[i] = g.tw[PASS == 0 ? 0 : 1][i];
  if (PASS == 0) {
    for (int i = tid; i < g.n1; i += blockDim.x) ssrc[i] = g.xsrc[i];
  } else {
    for (int i = tid; i < g.nplane; i += blockDim.x) srun[i] = g.plane_run[i];
  }
  if (tid == 0) {
    for (int s = 0; s < S; s++) mbar_init(full + s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  // work units
  const long plane = (long)g.n2 * g.n3;
  const long upg = PASS == 0 ? (plane + zl - 1) / zl : (long)g.nplane * ((g.n3 + zl - 1) / zl);   // units per group
  const long nunits = (long)ngroups * upg;
  const int nzc = (g.n3 + zl - 1) / zl;

  auto issue = [&](long unit, int s) {          // elected thread: fetch the inputs of `unit` into stage s
    const int grp = (int)(unit / upg);
    const long u = unit % upg;
    double2* dst = stage0 + (long)s * stage_elems;
    if (PASS == 0) {
      const long yz0 = u * zl;
      const int nbox = (g.nplane + 255) / 256;
      mbar_expect_tx(full + s, (unsigned)(stage_elems * sizeof(double2)));
      for (int j = 0; j < nbox; j++)
        tma_load_3d(dst + (long)j * 256 * zl * FFT_B, &tmap, 0, (int)yz0, grp * g.nplane + j * 256, full + s);
    } else {
      const int p = (int)(u / nzc), z0 = (int)(u % nzc) * zl;
      const int4 run = srun[p];
      const int nbox = (run.y + FFT_TMA_COLBOX - 1) / FFT_TMA_COLBOX;
      mbar_expect_tx(full + s, (unsigned)(nbox * FFT_TMA_COLBOX * zl * FFT_B * sizeof(double2)));
      for (int j = 0; j < nbox; j++)
        tma_load_3d(dst + (long)j * FFT_TMA_COLBOX * zl * FFT_B, &tmap, 0, z0,
                    grp * g.ncol + run.x + j * FFT_TMA_COLBOX, full + s);
    }
  };

  long nmine = 0;
  if ((long)blockIdx.x < nunits) nmine = (nunits - blockIdx.x + gridDim.x - 1) / gridDim.x;
  if (tid == 0)
    for (int s = 0; s < S - 1 && s < nmine; s++) issue(blockIdx.x + (long)s * gridDim.x, s);
  int lineno = 0;
  for (long i = 0; i < nmine; i++) {
    const int s = (int)(i % S);
    const unsigned parity = (unsigned)((i / S) & 1);
    // the stage of unit i-1 was consumed before the barrier that ended its last phase 1: refill it
    if (tid == 0 && i + S - 1 < nmine) issue(blockIdx.x + (i + S - 1) * gridDim.x, (int)((i + S - 1) % S));
    mbar_wait(full + s, parity);
    const long unit = blockIdx.x + i * gridDim.x;
    const int grp = (int)(unit / upg);
    const long u = unit % upg;
    const double2* st = stage0 + (long)s * stage_elems;
    for (int l = 0; l < zl; l++, lineno++) {
      double2* buf = bufs + (a.dbl ? (lineno & 1) : 0) * n * FFT_B;
      bool live;
      long yz = 0;
      int p = 0, z = 0;
      if (PASS == 0) {
        yz = u * zl + l;
        live = yz < plane;
      } else {
        p = (int)(u / nzc);
        z = (int)(u % nzc) * zl + l;
        live = z < g.n3;
      }
      if (live && q < R2) {
        const double2* in = st + (long)l * FFT_B + b;
        const long rstride = (long)zl * FFT_B;
        if (PASS == 0) {
          auto load = [&](int x) {
            const int pl = ssrc[x];
            return pl >= 0 ? in[(long)pl * rstride] : make_double2(0, 0);
          };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
          PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
        } else {
          const int4 run = srun[p];
          const int n2 = g.n2;
          auto load = [&](int y) {
            int d = y - run.z;
            if (d < 0) d += n2;
            if (d >= run.y) return make_double2(0, 0);
            const int j = d < run.w ? run.y - run.w + d : d - run.w;
            return in[(long)j * rstride];
          };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
          PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
        }
      }
      __syncthreads();
      if (live && q < R1) {
        if (PASS == 0) {
          double2* out = out_base + ((long)grp * g.n1 * plane + yz) * FFT_B + b;
          auto store = [&](int x, double2 v) { out[(long)x * plane * FFT_B] = v; };
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
          PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
        } else {
          double2* out = out_base + ((((long)grp * g.nplane + p) * g.n2) * g.n3 + z) * FFT_B + b;
          const long ystride = (long)g.n3 * FFT_B;
          auto store = [&](int y, double2 v) { out[(long)y * ystride] = v; };
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
          PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
        }
      }
      if (!a.dbl) __syncthreads();
    }
  }
}

// ---- forward transform, pruned on the OUTPUT side: X -> T2 -> T1 -> interleaved plane-wave coefficients ------------
// fwd_fft3d + gather (linalg.c:47-79) for whole band groups: F[G] = scale * sum_r x[r] e^{-2 pi i G.r/N} is evaluated
// as conj(inverse transform of conj(x)), so the line transforms above are reused unchanged; the data stays conjugated
// between the passes and is conjugated back (and narrowed to complex64, like the reference's store) by pass Z.
// Pruning mirrors the inverse transform: pass X keeps only the x-planes that hold plane waves, pass Y only the
// active columns, pass Z only each column's z-run.
template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_fwd_pass_x_kernel(FftGeom g, const double2* __restrict__ X, double2* __restrict__ T2, int ngroups) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* bufs = reinterpret_cast<double2*>(fft_smem);      // [2][n1][FFT_B]
  double2* tw = bufs + 2 * g.n1 * FFT_B;
  int* sxsrc = reinterpret_cast<int*>(tw + g.n1);
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  const int R1 = g.r1[0], R2 = g.r2[0];
  const long plane = (long)g.n2 * g.n3;
  const int xstride = (int)(plane * FFT_B);                   // 32-bit element offsets: n1 * plane * 16 < 2^31 (n <= 400)
  for (int i = tid; i < g.n1; i += blockDim.x) {
    tw[i] = g.tw[0][i];
    const int p = g.xsrc[i];
    sxsrc[i] = p >= 0 ? p * xstride : -1;
  }
  __syncthreads();
  const int iplane = (int)plane, step = (int)gridDim.x;
  int grp = (int)blockIdx.x / iplane, yz = (int)blockIdx.x % iplane;       // advanced incrementally
  for (int it = 0; grp < ngroups; it++) {
    double2* buf = bufs + (it & 1) * g.n1 * FFT_B;
    if (q < R2) {
      const double2* in = X + ((long)grp * g.n1 * plane + yz) * FFT_B + b;
      auto load = [&](int row) {
        const double2 v = in[row * xstride];
        return make_double2(v.x, -v.y);
      };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
      PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
    }
    __syncthreads();
    if (q < R1) {
      double2* out = T2 + ((long)grp * g.nplane * plane + yz) * FFT_B + b;
      auto store = [&](int row, double2 v) {
        const int o = sxsrc[row];
        if (o >= 0) out[o] = v;
      };
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
      PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
    }
    yz += step;
    while (yz >= iplane) { yz -= iplane; grp++; }
  }
}

template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_fwd_pass_y_kernel(FftGeom g, const double2* __restrict__ T2, double2* __restrict__ T1, int ngroups) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* bufs = reinterpret_cast<double2*>(fft_smem);      // [2][n2][FFT_B]
  double2* tw = bufs + 2 * g.n2 * FFT_B;
  int* ssrc = reinterpret_cast<int*>(tw + g.n2);             // [n2] column of each y row of the current plane
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  const int R1 = g.r1[1], R2 = g.r2[1];
  for (int i = tid; i < g.n2; i += blockDim.x) tw[i] = g.tw[1][i];
  const int nzc = (g.n3 + FFT_ZC - 1) / FFT_ZC;
  const int nunits = ngroups * g.nplane * nzc;
  const int colstride = g.n3 * FFT_B;
  int it = 0;
  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const int zc = unit % nzc;
    const int p = (unit / nzc) % g.nplane;
    const int grp = unit / (nzc * g.nplane);
    __syncthreads();
    for (int i = tid; i < g.n2; i += blockDim.x) {
      const int c = g.ysrc[p * g.n2 + i];
      ssrc[i] = c >= 0 ? c * colstride : -1;
    }
    __syncthreads();
    const int z1 = min(g.n3, (zc + 1) * FFT_ZC);
    for (int z = zc * FFT_ZC; z < z1; z++, it++) {
      double2* buf = bufs + (it & 1) * g.n2 * FFT_B;
      if (q < R2) {
        const double2* in = T2 + ((((long)grp * g.nplane + p) * g.n2) * g.n3 + z) * FFT_B + b;
        auto load = [&](int row) { return in[row * colstride]; };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
        PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
      }
      __syncthreads();
      if (q < R1) {
        double2* out = T1 + ((long)grp * g.ncol * g.n3 + z) * FFT_B + b;
        auto store = [&](int row, double2 v) {
          const int o = ssrc[row];
          if (o >= 0) out[o] = v;
        };
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
        PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
      }
    }
  }
}

// pass Z: T1 -> Cil-layout coefficients out[group][ldil][FFT_B] (complex64, box order), scaled and conjugated back
template <int RMAX>
__global__ void __launch_bounds__(FftLaunch<RMAX>::THREADS, FftLaunch<RMAX>::MINB)
fft_fwd_pass_z_kernel(FftGeom g, const double2* __restrict__ T1, float2* __restrict__ out_il, long ldil, double scale,
                      int ngroups) {
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* bufs = reinterpret_cast<double2*>(fft_smem);      // [2][n3][FFT_B]
  double2* tw = bufs + 2 * g.n3 * FFT_B;
  const int tid = threadIdx.x, q = tid / FFT_B, b = tid % FFT_B;
  const int R1 = g.r1[2], R2 = g.r2[2], n3 = g.n3;
  for (int i = tid; i < n3; i += blockDim.x) tw[i] = g.tw[2][i];
  __syncthreads();
  const int ncol = g.ncol, step = (int)gridDim.x;
  int grp = (int)blockIdx.x / ncol, col = (int)blockIdx.x % ncol;
  for (int it = 0; grp < ngroups; it++) {
    const int4 run = __ldg(g.col_run + col);
    double2* buf = bufs + (it & 1) * n3 * FFT_B;
    if (q < R2) {
      const double2* in = T1 + (((long)grp * g.ncol + col) * n3) * FFT_B + b;
      auto load = [&](int row) { return in[row * FFT_B]; };
#define P1(R) line_phase1<R>(buf, tw, R2, q, b, load)
      PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
    }
    __syncthreads();
    if (q < R1) {
      float2* out = out_il + ((long)grp * ldil + run.x) * FFT_B + b;
      auto store = [&](int row, double2 v) {
        int d = row - run.z;
        if (d < 0) d += n3;
        if (d >= run.y) return;
        const int j = d < run.w ? run.y - run.w + d : d - run.w;
        out[j * FFT_B] = make_float2((float)(v.x * scale), (float)(-v.y * scale));
      };
#define P2(R) line_phase2<R>(buf, R1, q, b, store)
      PAWB200_RADIX_SWITCH(R2, P2)
#undef P2
    }
    col += step;
    while (col >= ncol) { col -= ncol; grp++; }
  }
}

}  // namespace pawb200
