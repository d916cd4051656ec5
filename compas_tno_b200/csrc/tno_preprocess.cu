// One-time pre-processing of a fixed-plan problem on the GPU, batched over plan GEOMETRIES.
//
// The reference computes, once per form diagram on the host (src/compas_tno/problems/problems.py:516-528),
//     Edinv = -pinv(E[:, dep]),   B[dep] = Edinv E[:, ind],   B[ind] = I,   d[dep] = -Edinv ph
// with E = [Cit U ; Cit V] the (2 ni x m) horizontal equilibrium matrix, U = diag(C x), V = diag(C y)
// (problems.py:314-326).  E[:, dep] has full column rank (the independents are the columns a pivoted QR pushes past the
// rank, algorithms/independents.py:126-154), so pinv(Ed) Y is the least-squares solution of Ed X = Y.  A Monte-Carlo over
// the plan geometry (configuration 4 of BASELINE.json) needs this for every sampled geometry: a dense SVD per sample on the
// host (3354 x 3335 at discretisation 40) dwarfs the evaluations themselves.
//
// Here G geometries advance together: for every geometry the augmented matrix
//     A = [ E[:, dep] | -E[:, ind] | ph ]        (M = 2 ni rows, N = m - k columns to factorise, NR = k + 1 right-hand sides)
// is assembled on the device from the vertex coordinates (k_pre_assemble), reduced by a BLOCKED Householder QR in compact
// WY form (panels of 16 columns):
//     k_pre_panel   one CTA per geometry, one WARP per panel column: every column step is one reduction pass
//                   (s_c = x . a_c for all 16 columns at once; c = own column gives the norm, c < j the entries of T) and
//                   one update pass; the panel lives in shared memory when it fits (M <= ~1700), else it is worked on in
//                   place (L2 resident)
//     k_pre_update  A2 <- (I - V T' V') A2 for the trailing columns and the right-hand sides, one CTA per 32-column tile:
//                   W = V' A2 and A2 -= V (T' W) are genuine dense contractions and run on the FP64 tensor pipe
//                   (mma.sync.aligned.m8n8k4.f64) -- or on the DFMA pipe with use_mma = 0 (same results to rounding; the A/B
//                   timing is in profiles/)
// and finished by a blocked back substitution R X = (Q' Y)[0:N] (k_pre_backsolve, 8 right-hand sides per CTA) and a
// scatter into B (G x m x k) and d (G x m).  The rows [N, M) of Q' Y are the least-squares residual: its largest entry and
// the smallest / largest |r_jj| are returned per geometry so that a caller can tell when a frozen set of independents no
// longer fits a perturbed geometry (the reference would re-run find_independents there).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <string>
#include <vector>

#include "tno_plan.hpp"

namespace tno {

constexpr int PRE_NB = 16;    // panel width = warps of the panel kernel
constexpr int PRE_TC = 32;    // trailing columns per CTA
constexpr int PRE_RC = 64;    // rows per staged chunk
constexpr int PRE_CB = 8;     // right-hand sides per back-substitution CTA
constexpr int PRE_BS = 32;    // rows per back-substitution block

struct PreDims {
  int n, m, ni, k;
  int M, N, NR, NC, lda;
  int64_t strideA;   // doubles per geometry
};

// ---- assembly -----------------------------------------------------------------------------------------------------
// thread per (geometry, edge): the up to four entries of column colmap[e] (rows: the free end points, x block then y
// block); the ph column by the same grid afterwards.  A is zeroed beforehand.
__global__ void k_pre_assemble(PreDims D, const int* __restrict__ edges, const int* __restrict__ vfree,
                               const int* __restrict__ colmap, const double* __restrict__ xy, const double* __restrict__ ph,
                               int64_t ph_stride, double* __restrict__ Aall, int64_t G) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per = (int64_t)D.m + D.M;
  if (t >= G * per) return;
  const int64_t g = t / per;
  const int it = (int)(t - g * per);
  double* A = Aall + g * D.strideA;
  if (it < D.m) {
    const int e = it;
    const int u = edges[2 * e], v = edges[2 * e + 1];
    const double* X = xy + g * 2 * (int64_t)D.n;
    const double du = X[2 * v] - X[2 * u], dv = X[2 * v + 1] - X[2 * u + 1];   // (C x)_e, (C y)_e
    const int col = colmap[e];
    const double sg = col >= D.N ? -1.0 : 1.0;                                  // right-hand side block holds -E[:, ind]
    double* a = A + (int64_t)col * D.lda;
    const int fu = vfree[u], fv = vfree[v];
    if (fu >= 0) {   // C[e, u] = -1
      a[fu] = -sg * du;
      a[D.ni + fu] = -sg * dv;
    }
    if (fv >= 0) {   // C[e, v] = +1
      a[fv] = sg * du;
      a[D.ni + fv] = sg * dv;
    }
  } else {
    const int r = it - D.m;
    A[(int64_t)(D.NC - 1) * D.lda + r] = ph ? ph[g * ph_stride + r] : 0.0;
  }
}

// ---- panel factorisation --------------------------------------------------------------------------------------------
// Columns [j0, j0 + nbc) of every geometry, rows [j0, M).  Warp c owns panel column c.  Output: Householder vectors below
// the diagonal (unit diagonal implied), R on and above it, T (PRE_NB x PRE_NB, upper triangular) per geometry.
__global__ void __launch_bounds__(PRE_NB * 32) k_pre_panel(PreDims D, double* __restrict__ Aall, int j0, int nbc,
                                                           double* __restrict__ Tall, double* __restrict__ rdiag,
                                                           int use_smem) {
  extern __shared__ double psm[];
  double* Ts = psm;                       // PRE_NB x PRE_NB
  double* ssm = Ts + PRE_NB * PRE_NB;     // PRE_NB reductions
  double* zsm = ssm + PRE_NB;             // PRE_NB
  double* pan = zsm + PRE_NB;             // panel copy (use_smem)
  const int g = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* A = Aall + (int64_t)g * D.strideA;
  const int Mj = D.M - j0;
  double* Pn = use_smem ? pan : A + (int64_t)j0 * D.lda + j0;
  const int pitch = use_smem ? Mj : D.lda;
  if (use_smem) {
    for (int c = warp; c < nbc; c += PRE_NB) {
      const double* src = A + (int64_t)(j0 + c) * D.lda + j0;
      for (int r = lane; r < Mj; r += 32) pan[(size_t)c * pitch + r] = src[r];
    }
  }
  for (int i = tid; i < PRE_NB * PRE_NB; i += blockDim.x) Ts[i] = 0.0;
  __syncthreads();
  double dmin = rdiag[2 * g], dmax = rdiag[2 * g + 1];
  for (int jj = 0; jj < nbc && jj < Mj; ++jj) {
    double* xcol = Pn + (size_t)jj * pitch;
    // reduction pass: s_c = sum_{r > jj} x_r a_rc for every panel column c
    if (warp < nbc) {
      const double* col = Pn + (size_t)warp * pitch;
      double s0 = 0.0, s1 = 0.0;
      int r = jj + 1 + lane;
      for (; r + 32 < Mj; r += 64) {
        s0 = fma(xcol[r], col[r], s0);
        s1 = fma(xcol[r + 32], col[r + 32], s1);
      }
      if (r < Mj) s0 = fma(xcol[r], col[r], s0);
      s0 += s1;
      for (int off = 16; off > 0; off >>= 1) s0 += __shfl_xor_sync(0xffffffffu, s0, off);
      if (lane == 0) ssm[warp] = s0;
    }
    __syncthreads();
    const double alpha = xcol[jj], nrm2 = ssm[jj];
    double beta = alpha, tau = 0.0, scale = 0.0;
    if (nrm2 > 0.0) {
      beta = -copysign(sqrt(fma(alpha, alpha, nrm2)), alpha);
      tau = (beta - alpha) / beta;
      scale = 1.0 / (alpha - beta);
    }
    // update pass: a_c -= tau (v . a_c) v  for the columns to the right; v = [1 ; scale x]
    if (warp > jj && warp < nbc) {
      double* col = Pn + (size_t)warp * pitch;
      const double f = tau * fma(scale, ssm[warp], col[jj]);
      const double fs = f * scale;
      for (int r = jj + 1 + lane; r < Mj; r += 32) col[r] = fma(-fs, xcol[r], col[r]);
      if (lane == 0) col[jj] -= f;
    }
    if (warp == 0 && lane < jj) zsm[lane] = fma(scale, ssm[lane], Pn[(size_t)lane * pitch + jj]);   // v_i . v_jj, i < jj
    __syncthreads();
    if (warp == jj) {
      for (int r = jj + 1 + lane; r < Mj; r += 32) xcol[r] *= scale;
      if (lane == 0) xcol[jj] = beta;
    } else if (warp == ((jj + 1) & (PRE_NB - 1))) {
      // T[0:jj, jj] = -tau T[0:jj, 0:jj] z ;  T[jj, jj] = tau
      if (lane < jj) {
        double acc = 0.0;
        for (int p = lane; p < jj; ++p) acc = fma(Ts[lane * PRE_NB + p], zsm[p], acc);
        Ts[lane * PRE_NB + jj] = -tau * acc;
      }
      if (lane == 0) Ts[jj * PRE_NB + jj] = tau;
    }
    dmin = fmin(dmin, fabs(beta));
    dmax = fmax(dmax, fabs(beta));
    __syncthreads();
  }
  if (use_smem) {
    for (int c = warp; c < nbc; c += PRE_NB) {
      double* dst = A + (int64_t)(j0 + c) * D.lda + j0;
      for (int r = lane; r < Mj; r += 32) dst[r] = pan[(size_t)c * pitch + r];
    }
  }
  for (int i = tid; i < PRE_NB * PRE_NB; i += blockDim.x) Tall[(int64_t)g * PRE_NB * PRE_NB + i] = Ts[i];
  if (tid == 0) {
    rdiag[2 * g] = dmin;
    rdiag[2 * g + 1] = dmax;
  }
}

// ---- trailing update ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void pre_dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

// A2 <- (I - V T' V') A2 for the columns [cfirst + 32 blockIdx.x, + 32) of geometry blockIdx.y, rows [j0, M).
// 256 threads.  Shared memory: Vs (64 x 17), As (64 x 33), W (16 x 33), W2 (16 x 33), Ts (16 x 16).
template <bool MMA>
__global__ void __launch_bounds__(256) k_pre_update(PreDims D, double* __restrict__ Aall, int j0, int nbc, int cfirst,
                                                    const double* __restrict__ Tall) {
  constexpr int PV = PRE_NB + 1, PA = PRE_TC + 1;
  __shared__ double Vs[PRE_RC * PV];
  __shared__ double As[PRE_RC * PA];
  __shared__ double Ws[PRE_NB * PA];
  __shared__ double W2[PRE_NB * PA];
  __shared__ double Ts[PRE_NB * PRE_NB];
  const int g = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double* A = Aall + (int64_t)g * D.strideA;
  const int c0 = cfirst + PRE_TC * blockIdx.x;
  const int ncol = min(PRE_TC, D.NC - c0);
  const int Mj = D.M - j0;
  for (int i = tid; i < PRE_NB * PRE_NB; i += 256) Ts[i] = Tall[(int64_t)g * PRE_NB * PRE_NB + i];
  auto load_chunk = [&](int r0) {
    // V chunk: unit lower trapezoid, zero above the diagonal and for missing panel columns
    for (int i = tid; i < PRE_RC * PRE_NB; i += 256) {
      const int r = i & (PRE_RC - 1), q = i / PRE_RC;
      const int rl = r0 + r;
      double v = 0.0;
      if (q < nbc && rl < Mj) v = rl > q ? A[(int64_t)(j0 + q) * D.lda + j0 + rl] : (rl == q ? 1.0 : 0.0);
      Vs[r * PV + q] = v;
    }
    for (int i = tid; i < PRE_RC * PRE_TC; i += 256) {
      const int r = i & (PRE_RC - 1), c = i / PRE_RC;
      const int rl = r0 + r;
      As[r * PA + c] = (c < ncol && rl < Mj) ? A[(int64_t)(c0 + c) * D.lda + j0 + rl] : 0.0;
    }
  };
  // ---- pass 1: W = V' A2 (16 x 32) ----
  double w0 = 0.0, w1 = 0.0;
  const int mt = warp >> 2, nt = warp & 3;       // MMA: this warp's 8 x 8 tile of W
  const int cc = tid & 31, q0 = 2 * (tid >> 5);  // FMA: column and the two rows of W of this thread
  for (int r0 = 0; r0 < Mj; r0 += PRE_RC) {
    __syncthreads();
    load_chunk(r0);
    __syncthreads();
    if constexpr (MMA) {
      // A_frag(i, kk) = V[kk, 8 mt + i] ; B_frag(kk, j) = A2[kk, 8 nt + j] ; lane holds A_frag(l / 4, l % 4), B_frag(l % 4, l / 4)
#pragma unroll 4
      for (int k4 = 0; k4 < PRE_RC; k4 += 4) {
        const double a = Vs[(k4 + (lane & 3)) * PV + 8 * mt + (lane >> 2)];
        const double b = As[(k4 + (lane & 3)) * PA + 8 * nt + (lane >> 2)];
        pre_dmma(w0, w1, a, b);
      }
    } else {
#pragma unroll 8
      for (int r = 0; r < PRE_RC; ++r) {
        const double a = As[r * PA + cc];
        w0 = fma(Vs[r * PV + q0], a, w0);
        w1 = fma(Vs[r * PV + q0 + 1], a, w1);
      }
    }
  }
  if constexpr (MMA) {
    Ws[(8 * mt + (lane >> 2)) * PA + 8 * nt + 2 * (lane & 3)] = w0;
    Ws[(8 * mt + (lane >> 2)) * PA + 8 * nt + 2 * (lane & 3) + 1] = w1;
  } else {
    Ws[q0 * PA + cc] = w0;
    Ws[(q0 + 1) * PA + cc] = w1;
  }
  __syncthreads();
  // ---- W2 = T' W : W2[q, c] = sum_{p <= q} T[p, q] W[p, c] ----
  for (int i = tid; i < PRE_NB * PRE_TC; i += 256) {
    const int c = i & (PRE_TC - 1), q = i / PRE_TC;
    double acc = 0.0;
    for (int p = 0; p <= q; ++p) acc = fma(Ts[p * PRE_NB + q], Ws[p * PA + c], acc);
    W2[q * PA + c] = acc;
  }
  // ---- pass 2: A2 -= V W2 ----
  for (int r0 = 0; r0 < Mj; r0 += PRE_RC) {
    __syncthreads();
    load_chunk(r0);
    __syncthreads();
    if constexpr (MMA) {
      // 8 x 4 output tiles of 8 x 8; warp w takes tile row w, all four tile columns; K = 16 in four steps
      const int row = 8 * warp + (lane >> 2);
#pragma unroll
      for (int tn = 0; tn < 4; ++tn) {
        const int col = 8 * tn + 2 * (lane & 3);
        double c0r = As[row * PA + col], c1r = As[row * PA + col + 1];
#pragma unroll
        for (int k4 = 0; k4 < PRE_NB; k4 += 4) {
          const double a = -Vs[row * PV + k4 + (lane & 3)];                       // A_frag(i, kk) = -V[8 w + i, k4 + kk]
          const double b = W2[(k4 + (lane & 3)) * PA + 8 * tn + (lane >> 2)];     // B_frag(kk, j) = W2[k4 + kk, 8 tn + j]
          pre_dmma(c0r, c1r, a, b);
        }
        As[row * PA + col] = c0r;
        As[row * PA + col + 1] = c1r;
      }
    } else {
      double wreg[PRE_NB];
#pragma unroll
      for (int q = 0; q < PRE_NB; ++q) wreg[q] = W2[q * PA + cc];
      for (int r = tid >> 5; r < PRE_RC; r += 8) {
        double acc = As[r * PA + cc];
#pragma unroll
        for (int q = 0; q < PRE_NB; ++q) acc = fma(-Vs[r * PV + q], wreg[q], acc);
        As[r * PA + cc] = acc;
      }
    }
    __syncthreads();
    for (int i = tid; i < PRE_RC * PRE_TC; i += 256) {
      const int r = i & (PRE_RC - 1), c = i / PRE_RC;
      const int rl = r0 + r;
      if (c < ncol && rl < Mj) A[(int64_t)(c0 + c) * D.lda + j0 + rl] = As[r * PA + c];
    }
  }
}

// ---- back substitution and scatter ------------------------------------------------------------------------------------
// R X = Y for the right-hand sides [8 blockIdx.x, + 8) of geometry blockIdx.y; X overwrites Y (rows [0, N) of the
// right-hand-side columns).  Blocks of 32 rows from the bottom: the triangular block is solved by one thread per right-hand
// side from shared memory, the rows above are updated by the whole CTA.
__global__ void __launch_bounds__(256) k_pre_backsolve(PreDims D, double* __restrict__ Aall) {
  __shared__ double Rs[PRE_BS * (PRE_BS + 1)];
  __shared__ double Xs[PRE_BS * PRE_CB];
  const int g = blockIdx.y, tid = threadIdx.x;
  double* A = Aall + (int64_t)g * D.strideA;
  const int cb0 = D.N + PRE_CB * blockIdx.x;
  const int ncb = min(PRE_CB, D.NC - cb0);
  const int N = D.N;
  for (int b1 = N; b1 > 0; b1 -= PRE_BS) {
    const int b0 = max(0, b1 - PRE_BS), nb = b1 - b0;
    for (int i = tid; i < nb * nb; i += 256) {
      const int r = i % nb, c = i / nb;
      Rs[r * (PRE_BS + 1) + c] = A[(int64_t)(b0 + c) * D.lda + b0 + r];
    }
    for (int i = tid; i < nb * ncb; i += 256) {
      const int r = i % nb, c = i / nb;
      Xs[r * PRE_CB + c] = A[(int64_t)(cb0 + c) * D.lda + b0 + r];
    }
    __syncthreads();
    if (tid < ncb) {
      for (int r = nb - 1; r >= 0; --r) {
        double acc = Xs[r * PRE_CB + tid];
        for (int c = r + 1; c < nb; ++c) acc = fma(-Rs[r * (PRE_BS + 1) + c], Xs[c * PRE_CB + tid], acc);
        Xs[r * PRE_CB + tid] = acc / Rs[r * (PRE_BS + 1) + r];
      }
    }
    __syncthreads();
    for (int i = tid; i < nb * ncb; i += 256) {
      const int r = i % nb, c = i / nb;
      A[(int64_t)(cb0 + c) * D.lda + b0 + r] = Xs[r * PRE_CB + c];
    }
    // rows above: Y[0:b0, :] -= R[0:b0, b0:b1] X
    for (int r = tid; r < b0; r += 256) {
      double acc[PRE_CB];
#pragma unroll
      for (int c = 0; c < PRE_CB; ++c) acc[c] = 0.0;
      for (int j = 0; j < nb; ++j) {
        const double rv = A[(int64_t)(b0 + j) * D.lda + r];
#pragma unroll
        for (int c = 0; c < PRE_CB; ++c) acc[c] = fma(rv, Xs[j * PRE_CB + c], acc[c]);
      }
#pragma unroll
      for (int c = 0; c < PRE_CB; ++c)
        if (c < ncb) A[(int64_t)(cb0 + c) * D.lda + r] -= acc[c];
    }
    __syncthreads();
  }
}

// B (G x m x k), d (G x m) from X; info[g] = { max |residual rows|, min |r_jj|, max |r_jj|, max |X| }
__global__ void k_pre_scatter(PreDims D, const double* __restrict__ Aall, const int* __restrict__ dep, const int* __restrict__ ind,
                              const double* __restrict__ rdiag, double* __restrict__ B, double* __restrict__ dvec,
                              double* __restrict__ info) {
  __shared__ double red[2][8];
  const int g = blockIdx.x, tid = threadIdx.x;
  const double* A = Aall + (int64_t)g * D.strideA;
  double* Bg = B + (int64_t)g * D.m * D.k;
  double* dg = dvec + (int64_t)g * D.m;
  double xmax = 0.0, rmax = 0.0;
  for (int64_t i = tid; i < (int64_t)D.N * D.NR; i += blockDim.x) {
    const int r = (int)(i % D.N), c = (int)(i / D.N);
    const double v = A[(int64_t)(D.N + c) * D.lda + r];
    xmax = fmax(xmax, fabs(v));
    if (c < D.k) Bg[(int64_t)dep[r] * D.k + c] = v;
    else dg[dep[r]] = v;
  }
  for (int i = tid; i < D.k * D.k; i += blockDim.x) {
    const int j = i / D.k, c = i - j * D.k;
    Bg[(int64_t)ind[j] * D.k + c] = j == c ? 1.0 : 0.0;
  }
  for (int j = tid; j < D.k; j += blockDim.x) dg[ind[j]] = 0.0;
  for (int64_t i = tid; i < (int64_t)(D.M - D.N) * D.NR; i += blockDim.x) {
    const int r = D.N + (int)(i % (D.M - D.N)), c = (int)(i / (D.M - D.N));
    rmax = fmax(rmax, fabs(A[(int64_t)(D.N + c) * D.lda + r]));
  }
  for (int off = 16; off > 0; off >>= 1) {
    xmax = fmax(xmax, __shfl_xor_sync(0xffffffffu, xmax, off));
    rmax = fmax(rmax, __shfl_xor_sync(0xffffffffu, rmax, off));
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = xmax;
    red[1][tid >> 5] = rmax;
  }
  __syncthreads();
  if (tid == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
      xmax = fmax(xmax, red[0][w]);
      rmax = fmax(rmax, red[1][w]);
    }
    info[4 * g + 0] = rmax;
    info[4 * g + 1] = rdiag[2 * g];
    info[4 * g + 2] = rdiag[2 * g + 1];
    info[4 * g + 3] = xmax;
  }
}

__global__ void k_pre_init_rdiag(double* rdiag, int64_t G) {
  const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g < G) {
    rdiag[2 * g] = 1e300;
    rdiag[2 * g + 1] = 0.0;
  }
}

static PreDims pre_dims(int n, int m, int ni, int k) {
  PreDims D;
  D.n = n;
  D.m = m;
  D.ni = ni;
  D.k = k;
  D.M = 2 * ni;
  D.N = m - k;
  D.NR = k + 1;
  D.NC = D.N + D.NR;
  D.lda = (D.M + 1) / 2 * 2;
  D.strideA = (int64_t)D.lda * D.NC;
  return D;
}

}  // namespace tno

using namespace tno;

extern "C" {

int64_t tno_preprocess_work_bytes(int32_t n, int32_t m, int32_t ni, int32_t k, int64_t n_geometries) {
  if (n < 1 || m < 1 || ni < 1 || k < 0 || k >= m || n_geometries < 0) return -1;
  const PreDims D = pre_dims(n, m, ni, k);
  // A | T | rdiag | index arrays (edges, vfree, colmap, dep, ind)
  return (D.strideA + PRE_NB * PRE_NB + 2) * 8 * n_geometries + ((int64_t)2 * m + n + m + m + k + 64) * 4 + 256;
}

int tno_preprocess_batch(int32_t n, int32_t m, int32_t ni, int32_t k, const int32_t* edges, const int32_t* free_index,
                         const int32_t* ind, const double* xy_dev, const double* ph_dev, int64_t ph_stride,
                         int64_t n_geometries, double* B_dev, double* d_dev, double* info_dev, void* work_dev,
                         int64_t work_bytes, int32_t use_mma, void* cuda_stream) {
  if (!edges || !free_index || (k > 0 && !ind) || !xy_dev || !B_dev || !d_dev || !info_dev || !work_dev) {
    set_error("tno_preprocess_batch: null argument");
    return TNO_EINVAL;
  }
  if (n < 1 || m < 1 || ni < 1 || k < 0 || k >= m || n_geometries < 0 || 2 * ni < m - k) {
    set_error("tno_preprocess_batch: bad sizes (need 0 <= k < m and 2 ni >= m - k)");
    return TNO_EINVAL;
  }
  const int64_t G = n_geometries;
  if (G == 0) return TNO_OK;
  if (work_bytes < tno_preprocess_work_bytes(n, m, ni, k, G)) {
    set_error("tno_preprocess_batch: work buffer smaller than tno_preprocess_work_bytes()");
    return TNO_EINVAL;
  }
  if (G > 65535) {
    set_error("tno_preprocess_batch: at most 65535 geometries per call");
    return TNO_EINVAL;
  }
  const PreDims D = pre_dims(n, m, ni, k);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  // host index tables
  std::vector<int32_t> colmap(m, -1), dep;
  for (int j = 0; j < k; ++j) {
    if (ind[j] < 0 || ind[j] >= m || colmap[ind[j]] >= 0) {
      set_error("tno_preprocess_batch: ind must hold k distinct edge indices");
      return TNO_EINVAL;
    }
    colmap[ind[j]] = D.N + j;
  }
  for (int e = 0; e < m; ++e)
    if (colmap[e] < 0) {
      colmap[e] = (int32_t)dep.size();
      dep.push_back(e);
    }
  for (int e = 0; e < m; ++e)
    if (edges[2 * e] < 0 || edges[2 * e] >= n || edges[2 * e + 1] < 0 || edges[2 * e + 1] >= n) {
      set_error("tno_preprocess_batch: edge end point out of range");
      return TNO_EINVAL;
    }
  for (int v = 0; v < n; ++v)
    if (free_index[v] >= ni) {
      set_error("tno_preprocess_batch: free_index out of range");
      return TNO_EINVAL;
    }
  // carve the work buffer
  char* w = (char*)work_dev;
  double* A = (double*)w;
  w += D.strideA * 8 * G;
  double* T = (double*)w;
  w += (int64_t)PRE_NB * PRE_NB * 8 * G;
  double* rdiag = (double*)w;
  w += 16 * G;
  int32_t* d_edges = (int32_t*)w;
  w += (int64_t)2 * m * 4;
  int32_t* d_vfree = (int32_t*)w;
  w += (int64_t)n * 4;
  int32_t* d_colmap = (int32_t*)w;
  w += (int64_t)m * 4;
  int32_t* d_dep = (int32_t*)w;
  w += (int64_t)m * 4;
  int32_t* d_ind = (int32_t*)w;
  TNO_CUDA_OK(cudaMemcpyAsync(d_edges, edges, (size_t)2 * m * 4, cudaMemcpyHostToDevice, st));
  TNO_CUDA_OK(cudaMemcpyAsync(d_vfree, free_index, (size_t)n * 4, cudaMemcpyHostToDevice, st));
  TNO_CUDA_OK(cudaMemcpyAsync(d_colmap, colmap.data(), (size_t)m * 4, cudaMemcpyHostToDevice, st));
  TNO_CUDA_OK(cudaMemcpyAsync(d_dep, dep.data(), dep.size() * 4, cudaMemcpyHostToDevice, st));
  if (k > 0) TNO_CUDA_OK(cudaMemcpyAsync(d_ind, ind, (size_t)k * 4, cudaMemcpyHostToDevice, st));
  TNO_CUDA_OK(cudaStreamSynchronize(st));   // the host tables above are stack / vector memory
  TNO_CUDA_OK(cudaMemsetAsync(A, 0, (size_t)D.strideA * 8 * G, st));
  k_pre_init_rdiag<<<(unsigned)((G + 255) / 256), 256, 0, st>>>(rdiag, G);
  {
    const int64_t total = G * ((int64_t)m + D.M);
    k_pre_assemble<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(D, d_edges, d_vfree, d_colmap, xy_dev, ph_dev, ph_stride, A, G);
  }
  int dev = 0, max_smem = 0;
  TNO_CUDA_OK(cudaGetDevice(&dev));
  TNO_CUDA_OK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  const size_t base_smem = (size_t)(PRE_NB * PRE_NB + 2 * PRE_NB) * 8;
  TNO_CUDA_OK(cudaFuncSetAttribute(k_pre_panel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
  for (int j0 = 0; j0 < D.N; j0 += PRE_NB) {
    const int nbc = std::min(PRE_NB, D.N - j0);
    const size_t pan_bytes = (size_t)(D.M - j0) * nbc * 8;
    const int use_smem = base_smem + pan_bytes + 1024 <= (size_t)max_smem ? 1 : 0;
    k_pre_panel<<<(unsigned)G, PRE_NB * 32, base_smem + (use_smem ? pan_bytes : 0), st>>>(D, A, j0, nbc, T, rdiag, use_smem);
    const int cfirst = j0 + nbc;
    const int ntile = (D.NC - cfirst + PRE_TC - 1) / PRE_TC;
    if (ntile > 0) {
      dim3 grid((unsigned)ntile, (unsigned)G);
      if (use_mma) k_pre_update<true><<<grid, 256, 0, st>>>(D, A, j0, nbc, cfirst, T);
      else k_pre_update<false><<<grid, 256, 0, st>>>(D, A, j0, nbc, cfirst, T);
    }
  }
  {
    dim3 grid((unsigned)((D.NR + PRE_CB - 1) / PRE_CB), (unsigned)G);
    k_pre_backsolve<<<grid, 256, 0, st>>>(D, A);
  }
  k_pre_scatter<<<(unsigned)G, 256, 0, st>>>(D, A, d_dep, d_ind, rdiag, B_dev, d_dev, info_dev);
  TNO_CUDA_OK(cudaGetLastError());
  return TNO_OK;
}

}  // extern "C"
