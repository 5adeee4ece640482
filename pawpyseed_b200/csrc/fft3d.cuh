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
// (batch fastest -> conflict-free), with a two-factor Cooley-Tukey n = R1*R2 (R <= 16, radices 2,3,5,7 and
// their products <= 16 evaluated in registers).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pawb200 {

constexpr int FFT_B = 16;          // interleaved bands per group
constexpr int FFT_MAXR = 16;

// exp(+2 pi i m / R) for R <= 16 (filled by the host at start-up)
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

template <int R>
__device__ __forceinline__ void fft_phase1(double2* __restrict__ data, const double2* __restrict__ tw, int R2,
                                           int NB, int j2, int e) {
  double2 v[R];
#pragma unroll
  for (int j1 = 0; j1 < R; j1++) v[j1] = data[(j1 * R2 + j2) * NB + e];
  SmallDFT<R, 1>::run(v);
#pragma unroll
  for (int k1 = 0; k1 < R; k1++) {
    double2 x = v[k1];
    if (k1 > 0 && j2 > 0) x = cmulf(x, tw[j2 * k1]);
    data[(k1 * R2 + j2) * NB + e] = x;
  }
}

template <int R>
__device__ __forceinline__ void fft_phase2_load(const double2* __restrict__ data, int NB, int k1, int e,
                                                double2* v) {
#pragma unroll
  for (int j2 = 0; j2 < R; j2++) v[j2] = data[(k1 * R + j2) * NB + e];
  SmallDFT<R, 1>::run(v);
}
template <int R>
__device__ __forceinline__ void fft_phase2_store(double2* __restrict__ data, int R1, int NB, int k1, int e,
                                                 const double2* v) {
#pragma unroll
  for (int k2 = 0; k2 < R; k2++) data[(k1 + R1 * k2) * NB + e] = v[k2];
}

#define PAWB200_RADIX_SWITCH(R, CALL)                                                            \
  switch (R) {                                                                                    \
    case 2: CALL(2); break;   case 3: CALL(3); break;   case 4: CALL(4); break;                   \
    case 5: CALL(5); break;   case 6: CALL(6); break;   case 7: CALL(7); break;                   \
    case 8: CALL(8); break;   case 9: CALL(9); break;   case 10: CALL(10); break;                 \
    case 12: CALL(12); break; case 14: CALL(14); break; case 15: CALL(15); break;                 \
    case 16: CALL(16); break; default: break;                                                     \
  }

// In-place inverse DFT of NB independent length-n = R1*R2 sequences stored as data[n][NB].
// Every thread of the CTA must call this (it contains __syncthreads()); blockDim >= max(R1,R2)*NB.
__device__ __forceinline__ void fft_lines_smem(double2* data, const double2* tw, int R1, int R2, int NB) {
  const int tid = threadIdx.x;
  {
    const int j2 = tid / NB, e = tid % NB;
    if (j2 < R2) {
#define P1(R) fft_phase1<R>(data, tw, R2, NB, j2, e)
      PAWB200_RADIX_SWITCH(R1, P1)
#undef P1
    }
  }
  __syncthreads();
  {
    const int k1 = tid / NB, e = tid % NB;
    double2 v[FFT_MAXR];
    const bool act = k1 < R1;
    if (act) {
#define P2L(R) fft_phase2_load<R>(data, NB, k1, e, v)
      PAWB200_RADIX_SWITCH(R2, P2L)
#undef P2L
    }
    __syncthreads();
    if (act) {
#define P2S(R) fft_phase2_store<R>(data, R1, NB, k1, e, v)
      PAWB200_RADIX_SWITCH(R2, P2S)
#undef P2S
    }
  }
  __syncthreads();
}

struct FftGeom {            // device-side description of one (k-point, grid) pruned transform
  int n1, n2, n3;           // grid
  int r1[3], r2[3];         // n_d = r1[d] * r2[d] (index 0: x, 1: y, 2: z)
  int ncol, nplane;         // active (g1,g2) columns / active g1 planes
  const int* col_start;     // [ncol] first sorted plane-wave index of the column
  const int* col_cnt;       // [ncol]
  const int* col_ypos;      // [ncol] wrapped g2
  const int* zpos;          // [npw]  wrapped g3 of each sorted plane wave
  const int* plane_col0;    // [nplane] first column of the plane
  const int* plane_ncol;    // [nplane]
  const int* plane_xpos;    // [nplane] wrapped g1
  const double2* tw[3];     // exp(+2 pi i m / n_d), m < n_d
};

// ---- pass Z: coefficients -> T1[group][col][z][FFT_B] ---------------------------------------------
// slot -> coefficient row like scatter_pw_kernel: band = slot / halves, half = slot % halves
template <int LPC>
__global__ void __launch_bounds__(FFT_MAXR * FFT_B * LPC, LPC == 1 ? 3 : 1)
fft_pass_z_kernel(FftGeom g, const float2* __restrict__ C, long ldc, int halves, int half_len, int slot0,
                  int nslot, double scale, double2* __restrict__ T1) {
  constexpr int NB = FFT_B * LPC;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* data = reinterpret_cast<double2*>(fft_smem);      // [n3][NB]
  double2* tw = data + g.n3 * NB;                             // [n3]
  const int grp = blockIdx.y, col0 = blockIdx.x * LPC;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < g.n3 * NB; i += nthr) data[i] = make_double2(0, 0);
  for (int i = tid; i < g.n3; i += nthr) tw[i] = g.tw[2][i];
  __syncthreads();
  // sparse load: warp <-> (line, band), lanes <-> consecutive plane waves of the column (contiguous 8-B reads)
  const int warp = tid >> 5, lane = tid & 31, nwarp = nthr >> 5;
  for (int q = warp; q < NB; q += nwarp) {
    const int lc = q / FFT_B, b = q % FFT_B;
    const int col = col0 + lc;
    if (col >= g.ncol) continue;
    int slot = slot0 + grp * FFT_B + b;
    if (slot >= slot0 + nslot) continue;            // pad bands of the last group stay zero
    const int band = slot / halves, half = slot % halves;
    const float2* row = C + (long)band * ldc + (long)half * half_len;
    const int s = g.col_start[col], cnt = g.col_cnt[col];
    for (int j = lane; j < cnt; j += 32) {
      const float2 c = __ldg(row + s + j);
      data[g.zpos[s + j] * NB + q] = make_double2(scale * (double)c.x, scale * (double)c.y);
    }
  }
  __syncthreads();
  fft_lines_smem(data, tw, g.r1[2], g.r2[2], NB);
  // store: T1[((grp*ncol + col)*n3 + z)*FFT_B + b]
  for (int i = tid; i < g.n3 * NB; i += nthr) {
    const int b = i % FFT_B, z = (i / FFT_B) % g.n3, lc = i / (FFT_B * g.n3);
    const int col = col0 + lc;
    if (col < g.ncol) T1[(((long)grp * g.ncol + col) * g.n3 + z) * FFT_B + b] = data[z * NB + lc * FFT_B + b];
  }
}

// ---- pass Y: T1 -> T2[group][plane][y][z][FFT_B] ------------------------------------------------------
template <int LPC>
__global__ void __launch_bounds__(FFT_MAXR * FFT_B * LPC, LPC == 1 ? 3 : 1)
fft_pass_y_kernel(FftGeom g, const double2* __restrict__ T1, double2* __restrict__ T2) {
  constexpr int NB = FFT_B * LPC;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* data = reinterpret_cast<double2*>(fft_smem);      // [n2][NB]
  double2* tw = data + g.n2 * NB;
  const int nzc = (g.n3 + LPC - 1) / LPC;
  const int p = blockIdx.x / nzc, z0 = (blockIdx.x % nzc) * LPC, grp = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < g.n2 * NB; i += nthr) data[i] = make_double2(0, 0);
  for (int i = tid; i < g.n2; i += nthr) tw[i] = g.tw[1][i];
  __syncthreads();
  const int c0 = g.plane_col0[p], nc = g.plane_ncol[p];
  for (int i = tid; i < nc * NB; i += nthr) {
    const int e = i % NB, c = c0 + i / NB;
    const int lz = e / FFT_B, b = e % FFT_B;
    if (z0 + lz < g.n3)
      data[g.col_ypos[c] * NB + e] = T1[(((long)grp * g.ncol + c) * g.n3 + z0 + lz) * FFT_B + b];
  }
  __syncthreads();
  fft_lines_smem(data, tw, g.r1[1], g.r2[1], NB);
  for (int i = tid; i < g.n2 * NB; i += nthr) {
    const int e = i % NB, y = i / NB;
    const int lz = e / FFT_B, b = e % FFT_B;
    if (z0 + lz < g.n3)
      T2[((((long)grp * g.nplane + p) * g.n2 + y) * g.n3 + z0 + lz) * FFT_B + b] = data[y * NB + e];
  }
}

// ---- pass X: T2 -> X[group][x][y][z][FFT_B] --------------------------------------------------------------
template <int LPC>
__global__ void __launch_bounds__(FFT_MAXR * FFT_B * LPC, LPC == 1 ? 3 : 1)
fft_pass_x_kernel(FftGeom g, const double2* __restrict__ T2, double2* __restrict__ X) {
  constexpr int NB = FFT_B * LPC;
  extern __shared__ __align__(16) unsigned char fft_smem[];
  double2* data = reinterpret_cast<double2*>(fft_smem);      // [n1][NB]
  double2* tw = data + g.n1 * NB;
  const int nzc = (g.n3 + LPC - 1) / LPC;
  const int y = blockIdx.x / nzc, z0 = (blockIdx.x % nzc) * LPC, grp = blockIdx.y;
  const int tid = threadIdx.x, nthr = blockDim.x;
  for (int i = tid; i < g.n1 * NB; i += nthr) data[i] = make_double2(0, 0);
  for (int i = tid; i < g.n1; i += nthr) tw[i] = g.tw[0][i];
  __syncthreads();
  for (int i = tid; i < g.nplane * NB; i += nthr) {
    const int e = i % NB, p = i / NB;
    const int lz = e / FFT_B, b = e % FFT_B;
    if (z0 + lz < g.n3)
      data[g.plane_xpos[p] * NB + e] = T2[((((long)grp * g.nplane + p) * g.n2 + y) * g.n3 + z0 + lz) * FFT_B + b];
  }
  __syncthreads();
  fft_lines_smem(data, tw, g.r1[0], g.r2[0], NB);
  const long ngrid = (long)g.n1 * g.n2 * g.n3;
  for (int i = tid; i < g.n1 * NB; i += nthr) {
    const int e = i % NB, x = i / NB;
    const int lz = e / FFT_B, b = e % FFT_B;
    if (z0 + lz < g.n3)
      X[((long)grp * ngrid + ((long)x * g.n2 + y) * g.n3 + z0 + lz) * FFT_B + b] = data[x * NB + e];
  }
}

}  // namespace pawb200
