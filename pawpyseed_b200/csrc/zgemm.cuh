// Complex "A * B^H" GEMM on the FP64 tensor cores (DMMA.8x8x4), sm_100a.
//
//   O[m][n] (+)= sum_k A[m][k] * conj(B[n][k])       A: M x K, B: N x K, both K-contiguous
//
// This is the band-band pseudo overlap  <psi~_R,n | psi~_S,m> = sum_G conj(C_R[n][G]) C_S[m][G]
// (pseudoprojector.c:63-90, one cblas_cdotc_sub per pair in the reference) with T = float2 -
// complex64 coefficients widened exactly to FP64 in registers - and the one-centre augmentation
// contraction sum_p conj(P_R[n][p]) (dO P_S)[m][p] (projector.c:890-959) with T = double2.
//
// tcgen05 has no f64 kind, so FP64 tensor work on Blackwell is warp-level mma.sync; the
// 4-real-product form is used (real: ArBr + AiBi, imag: AiBr - ArBi).
//
// Scheduling: persistent CTAs with a static, deterministic stream-K partition of the
// (tile, k-iteration) space, so 600x600 (100 tiles) and 2000x2000 (1024 tiles) outputs both
// keep all 148 SMs busy; tiles cut across CTAs go through a workspace and are summed in a
// fixed order by zgemm_fixup_kernel (bit-reproducible, no atomics).
#pragma once
#include <cuda_runtime.h>

#include "kernels.cuh"

namespace pawb200 {

constexpr int ZG_BM = 64;                 // CTA tile rows
constexpr int ZG_STAGES = 4;
constexpr int ZG_THREADS = 128;           // 2 x 2 warps

// Two arithmetic variants of the complex product:
//   4M: real += ArBr + AiBi, imag += AiBr - ArBi           (4 DMMA per 8x8x4 tile step, 32x32 warp tile)
//   3M: P1 += ArBr, P2 += AiBi, P3 += (Ar+Ai)(Br-Bi);      (3 DMMA, Karatsuba; 32x24 warp tile because of the
//       real = P1+P2, imag = P3-P1+P2                       third accumulator set) - 25 % less tensor work
template <bool K3M> struct ZgShape {
  static constexpr int BN = K3M ? 48 : 64;      // CTA tile columns
  static constexpr int NT = K3M ? 3 : 4;        // n-tiles (8 wide) per warp
};

template <typename T> struct ZgTraits;
template <> struct ZgTraits<float2> {
  static constexpr int KT = 16;           // k elements per stage (128 B per row)
  static constexpr int LD = KT + 4;       // element row stride in smem: conflict-free 8-B fragments
  static constexpr int CHUNK = 2;         // elements per 16-B cp.async
};
template <> struct ZgTraits<double2> {
  static constexpr int KT = 8;
  static constexpr int LD = KT + 4;
  static constexpr int CHUNK = 1;
};

template <typename T, bool K3M>
constexpr size_t zgemm_smem_bytes() {
  return sizeof(T) * ZG_STAGES * (ZG_BM + ZgShape<K3M>::BN) * ZgTraits<T>::LD;
}

__device__ __forceinline__ void widen(const float2 v, double& re, double& im) {
  re = (double)v.x;
  im = (double)v.y;
}
__device__ __forceinline__ void widen(const double2 v, double& re, double& im) {
  re = v.x;
  im = v.y;
}

struct ZgPlan {
  int M, N;
  int tiles_m, tiles_n;
  int bn;               // CTA tile columns of the variant in use
  long kiters;          // K_padded / KT
  long total;           // tiles * kiters
  int G;                // persistent CTAs
};

__host__ __device__ inline long zg_unit_begin(const ZgPlan& p, int g) {
  return (long)(((__int128)p.total * g) / p.G);
}

template <typename T, bool K3M>
__global__ void __launch_bounds__(ZG_THREADS, 2)
zgemm_abh_kernel(const T* __restrict__ A, long lda, const T* __restrict__ B, long ldb, ZgPlan plan,
                 double2* __restrict__ out, long ldo, int accumulate, double2* __restrict__ ws) {
  constexpr int KT = ZgTraits<T>::KT, LD = ZgTraits<T>::LD, CH = ZgTraits<T>::CHUNK;
  constexpr int BN = ZgShape<K3M>::BN, NT = ZgShape<K3M>::NT, NACC = K3M ? 3 : 2;
  constexpr int CPR = KT / CH;                       // 16-B chunks per tile row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sA = reinterpret_cast<T*>(smem_raw);            // [ST][BM][LD]
  T* sB = sA + ZG_STAGES * ZG_BM * LD;               // [ST][BN][LD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int wm = warp >> 1, wn = warp & 1;
  const int g = blockIdx.x;
  const long u0 = zg_unit_begin(plan, g), u1 = zg_unit_begin(plan, g + 1);

  long u = u0;
  while (u < u1) {
    const int tile = (int)(u / plan.kiters);
    const long kb = u % plan.kiters;
    const long ke = (plan.kiters - kb < u1 - u) ? plan.kiters : kb + (u1 - u);
    const int tm = tile / plan.tiles_n, tn = tile % plan.tiles_n;
    const int row0 = tm * ZG_BM, col0 = tn * BN;

    double acc[NACC][4][NT][2];
#pragma unroll
    for (int q = 0; q < NACC; q++)
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < NT; j++) acc[q][i][j][0] = acc[q][i][j][1] = 0;

    auto issue = [&](long kit, int st) {
      if (kit < ke) {
        const long k0 = kit * KT;
#pragma unroll
        for (int c = tid; c < ZG_BM * CPR; c += ZG_THREADS) {
          const int r = c / CPR, q = c % CPR;
          int gr = row0 + r;
          if (gr >= plan.M) gr = plan.M - 1;
          cp_async16(sA + (st * ZG_BM + r) * LD + q * CH, A + (long)gr * lda + k0 + q * CH);
        }
#pragma unroll
        for (int c = tid; c < BN * CPR; c += ZG_THREADS) {
          const int r = c / CPR, q = c % CPR;
          int gr = col0 + r;
          if (gr >= plan.N) gr = plan.N - 1;
          cp_async16(sB + (st * BN + r) * LD + q * CH, B + (long)gr * ldb + k0 + q * CH);
        }
      }
      cp_async_commit();
    };

    __syncthreads();   // previous segment's readers are done with the stage ring
#pragma unroll
    for (int s = 0; s < ZG_STAGES - 1; s++) issue(kb + s, s);

    for (long kit = kb; kit < ke; kit++) {
      cp_async_wait<ZG_STAGES - 2>();
      __syncthreads();
      issue(kit + ZG_STAGES - 1, (int)((kit - kb + ZG_STAGES - 1) % ZG_STAGES));
      const int st = (int)((kit - kb) % ZG_STAGES);
      const T* tA = sA + (st * ZG_BM + wm * 32 + (lane >> 2)) * LD + (lane & 3);
      const T* tB = sB + (st * BN + wn * (8 * NT) + (lane >> 2)) * LD + (lane & 3);
#pragma unroll
      for (int kk = 0; kk < KT / 4; kk++) {
        double ar[4], ai[4], br[NT], bi[NT];
#pragma unroll
        for (int i = 0; i < 4; i++) widen(tA[(8 * i) * LD + 4 * kk], ar[i], ai[i]);
#pragma unroll
        for (int j = 0; j < NT; j++) widen(tB[(8 * j) * LD + 4 * kk], br[j], bi[j]);
        if constexpr (K3M) {
          double as[4], bd[NT];
#pragma unroll
          for (int i = 0; i < 4; i++) as[i] = ar[i] + ai[i];
#pragma unroll
          for (int j = 0; j < NT; j++) bd[j] = br[j] - bi[j];
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
              dmma884(acc[0][i][j][0], acc[0][i][j][1], ar[i], br[j]);
              dmma884(acc[1][i][j][0], acc[1][i][j][1], ai[i], bi[j]);
              dmma884(acc[2][i][j][0], acc[2][i][j][1], as[i], bd[j]);
            }
        } else {
          double nbi[NT];
#pragma unroll
          for (int j = 0; j < NT; j++) nbi[j] = -bi[j];
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
              dmma884(acc[0][i][j][0], acc[0][i][j][1], ar[i], br[j]);
              dmma884(acc[1][i][j][0], acc[1][i][j][1], ai[i], br[j]);
            }
#pragma unroll
          for (int i = 0; i < 4; i++)
#pragma unroll
            for (int j = 0; j < NT; j++) {
              dmma884(acc[0][i][j][0], acc[0][i][j][1], ai[i], bi[j]);
              dmma884(acc[1][i][j][0], acc[1][i][j][1], ar[i], nbi[j]);
            }
        }
      }
    }
    cp_async_wait<0>();

    // ---- store: full tiles go to `out`, partial k-ranges to the workspace -------------
    const bool full = (kb == 0 && ke == plan.kiters);
    const int slot = (u == u0) ? 0 : 1;
    double2* w = ws + ((long)g * 2 + slot) * (ZG_BM * BN);
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
      for (int j = 0; j < NT; j++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
          double2 v;
          if constexpr (K3M)
            v = make_double2(acc[0][i][j][e] + acc[1][i][j][e],
                             acc[2][i][j][e] - acc[0][i][j][e] + acc[1][i][j][e]);
          else
            v = make_double2(acc[0][i][j][e], acc[1][i][j][e]);
          const int lr = wm * 32 + 8 * i + (lane >> 2);
          const int lc = wn * (8 * NT) + 8 * j + 2 * (lane & 3) + e;
          if (full) {
            const int r = row0 + lr, c = col0 + lc;
            if (r < plan.M && c < plan.N) {
              double2* o = out + (long)r * ldo + c;
              if (accumulate) {
                const double2 old = *o;
                v.x += old.x;
                v.y += old.y;
              }
              *o = v;
            }
          } else {
            w[lr * BN + lc] = v;
          }
        }
    u += ke - kb;
  }
}

// Sums the workspace partials of every tile that was cut across CTAs, in ascending CTA order.
__global__ void __launch_bounds__(256)
zgemm_fixup_kernel(ZgPlan plan, const double2* __restrict__ ws, double2* __restrict__ out, long ldo,
                   int accumulate) {
  const int tile = blockIdx.x;
  const int BN = plan.bn;
  const long t0 = (long)tile * plan.kiters, t1 = t0 + plan.kiters;
  // first CTA whose range ends after t0
  int g = (int)(((__int128)t0 * plan.G) / plan.total);
  while (g > 0 && zg_unit_begin(plan, g) > t0) g--;
  while (zg_unit_begin(plan, g + 1) <= t0) g++;
  if (zg_unit_begin(plan, g) <= t0 && zg_unit_begin(plan, g + 1) >= t1) return;  // whole tile, one CTA
  const int tm = tile / plan.tiles_n, tn = tile % plan.tiles_n;
  for (int e = threadIdx.x; e < ZG_BM * BN; e += blockDim.x) {
    const int r = tm * ZG_BM + e / BN, c = tn * BN + e % BN;
    if (r >= plan.M || c >= plan.N) continue;
    double2 acc = accumulate ? out[(long)r * ldo + c] : make_double2(0, 0);
    for (int gg = g; gg < plan.G && zg_unit_begin(plan, gg) < t1; gg++) {
      const long b0 = zg_unit_begin(plan, gg), b1 = zg_unit_begin(plan, gg + 1);
      if (b1 <= t0 || b0 == b1) continue;
      const int slot = (b0 >= t0) ? 0 : 1;
      const double2 v = ws[((long)gg * 2 + slot) * (ZG_BM * BN) + e];
      acc.x += v.x;
      acc.y += v.y;
    }
    out[(long)r * ldo + c] = acc;
  }
}

}  // namespace pawb200
