// Kernel family 1: "interleaved" -- one thread per candidate, values batch-interleaved in HBM.
//
// Every per-candidate number lives in a scratch "slot"; slot s of candidate c is stored at
// scratch[s * NC + c], so the 32 lanes of a warp (32 consecutive candidates, all executing the same
// symbolic program) always touch 256 contiguous bytes.  Pattern indices are warp-uniform loads.
// This family has no size limit (the factor does not have to fit on chip) and is the general path;
// the on-chip family (tno_onchip.cu) is selected when factor + right-hand sides fit in shared memory.
//
// Math restated from the reference (paths relative to /root/reference/src/compas_tno/):
//   q = B q_ind + d                     algorithms/equilibrium.py:190
//   X_free = (Ci^T Q Ci)^-1 (P_i - Ci^T Q Cb X_b)     algorithms/equilibrium.py:230-234
//   dz/dq_ind, dz/dzb, dz/dlambda       problems/jacobian.py:157-160,190-192,297-308,348-353
//   reactions, f, grad f, reac_bounds   problems/objectives.py:144-146, derivatives.py:303-331,
//                                       constraints.py:146-149, jacobian.py:204-255
// The matrix is negative definite on this branch (compression = negative q), so we factor
// -(Ci^T Q Ci) = L D L^T with a fixed pattern and no pivoting; D > 0 is checked per candidate.
#include <cstdio>
#include <cstdlib>

#include "tno_plan.hpp"

namespace tno {

#define WS(slot) ws[(int64_t)(slot) * NC + c]

// Horizontal load basis of a candidate: (c, s) = its load direction (tno_batch_inputs.load_dir), (1, 0) without one;
// the second basis is only dereferenced when the plan has it, and c = 1, s = 0 reproduces the single-basis values bit
// for bit.
#define DIR_C WS(P.slot.s_dir)
#define DIR_S WS(P.slot.s_dir + 1)
__device__ __forceinline__ double mix2(const double* __restrict__ a, const double* __restrict__ b, int i, double c, double s) {
  return b ? fma(s, b[i], c * a[i]) : a[i];
}

__global__ void k_pack(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, const double* __restrict__ x,
                       int64_t ldx, const double* __restrict__ pz_scale, const double* __restrict__ load_dir) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int nvar = P.var.nvar;
  for (int j = 0; j < nvar; ++j) WS(P.slot.s_x + j) = x[c * ldx + j];
  WS(P.slot.s_pzs) = pz_scale ? pz_scale[c] : 1.0;
  DIR_C = load_dir ? load_dir[2 * c] : 1.0;
  DIR_S = load_dir ? load_dir[2 * c + 1] : 0.0;
}

// q_e = lambdh * d_e + sum_j B[e, j] q_ind[j]
__global__ void k_q(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, int etile) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int e0 = blockIdx.y * etile;
  const int e1 = min(P.m, e0 + etile);
  const int k = P.k;
  const double lam = P.var.o_lambdh >= 0 ? WS(P.slot.s_x + P.var.o_lambdh) : 1.0;
  const double dc = DIR_C, ds = DIR_S;
  for (int e = e0; e < e1; ++e) {
    double acc = lam * mix2(P.d, P.d_b, e, dc, ds);
    const double* Brow = P.B + (int64_t)e * k;
    for (int j = 0; j < k; ++j) acc = fma(Brow[j], WS(P.slot.s_x + j), acc);
    WS(P.slot.s_q + e) = acc;
  }
}

// values of -(Ci^T Q Ci) scattered into the pattern of L (fill positions get 0)
__global__ void k_assemble(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, int ptile) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int p0 = blockIdx.y * ptile;
  const int p1 = min(P.nnz_l, p0 + ptile);
  for (int p = p0; p < p1; ++p) {
    double acc = 0.0;
    for (int t = P.asm_ptr[p]; t < P.asm_ptr[p + 1]; ++t) acc += (double)P.asm_coef[t] * WS(P.slot.s_q + P.asm_edge[t]);
    WS(P.slot.s_l + p) = acc;
  }
}

// right-looking L D L^T; the diagonal slot ends up holding 1 / d_j
__global__ void k_factor(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, int32_t* __restrict__ status) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  int st = 0;
  const int sl = P.slot.s_l;
  for (int j = 0; j < P.ni; ++j) {
    const int p0 = P.colptr[j], p1 = P.colptr[j + 1];
    const double dj = WS(sl + p0);
    if (!(dj > 0.0)) st |= TNO_STATUS_NOT_PD;
    const double inv = 1.0 / dj;
    WS(sl + p0) = inv;
    int64_t t = P.upd_ptr[j];
    for (int pa = p0 + 1; pa < p1; ++pa) {
      const double va = WS(sl + pa);
      const double la = va * inv;
      {
        const int tg = P.upd_pos[t++];
        WS(sl + tg) = fma(-va, la, WS(sl + tg));
      }
      for (int pb = pa + 1; pb < p1; ++pb) {
        const int tg = P.upd_pos[t++];
        WS(sl + tg) = fma(-WS(sl + pb), la, WS(sl + tg));
      }
      WS(sl + pa) = la;
    }
  }
  status[c] = st;
}

// ---- level-scheduled factorisation and solves ----------------------------------------------------------------------
// A thread per candidate leaves a large problem latency bound: at discretisation 40 a chunk of ~9k candidates is two
// warps per SM walking 0.65 M dependent updates.  These kernels take the parallelism of the elimination tree as well:
// one launch per level, a thread per (candidate, item of the level); a warp still holds 32 candidates of ONE item, so
// every access to the batch-interleaved values stays coalesced and the index loads are warp-uniform.
// Storage while they run: off-diagonals hold U = L D (unscaled), diagonals hold 1 / d.
//   factor (left-looking):  U(i, j) = A(i, j) - sum_c U(i, c) U(j, c) / d_c ,  d_j likewise
//   forward (pull):         z_i = (b_i - sum_c U(i, c) z_c) / d_i            (L y = b and y = D z in one sweep)
//   backward:               x_j = z_j - (sum_i U(i, j) x_i) / d_j
__global__ void __launch_bounds__(128) k_factor_level(DevPlan P, DevLevels L, double* __restrict__ ws, int64_t NC, int64_t ncand,
                                                      int item0, int item1, int tile, int32_t* __restrict__ status) {
  const int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int sl = P.slot.s_l;
  const int i0 = item0 + blockIdx.x * tile;
  const int i1 = min(item1, i0 + tile);
  for (int it = i0; it < i1; ++it) {
    const int pe = L.items[it];
    const int p = pe & 0x7fffffff;
    const int d0 = L.dotptr[it], d1 = L.dotptr[it + 1];
    double v = WS(sl + p);
#pragma unroll 8
    for (int t = d0; t < d1; ++t) {
      const int a = L.dot[3 * t], b = L.dot[3 * t + 1], dg = L.dot[3 * t + 2];
      v = fma(-(WS(sl + a) * WS(sl + dg)), WS(sl + b), v);
    }
    if (pe < 0) {
      if (!(v > 0.0)) atomicOr(status + c, TNO_STATUS_NOT_PD);
      v = 1.0 / v;
    }
    WS(sl + p) = v;
  }
}

template <int RB>
__global__ void __launch_bounds__(128) k_forward_level(DevPlan P, DevLevels L, double* __restrict__ ws, int64_t NC, int64_t ncand,
                                                       int row0, int nrows, int col0, int ncols, int ngroups) {
  const int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int item = blockIdx.x / ngroups, grp = blockIdx.x - item * ngroups;
  if (item >= nrows) return;
  const int i = L.cols[row0 + item];
  const int cb = col0 + grp * RB;
  const int nr = min(RB, col0 + ncols - cb);
  const int nrhs = P.rhs.nrhs;
  const int sl = P.slot.s_l, sr = P.slot.s_r + cb;
  double acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = (r < nr) ? WS(sr + i * nrhs + r) : 0.0;
  const int t1 = L.rowptr[i + 1];
#pragma unroll(RB > 4 ? 4 : 8)
  for (int t = L.rowptr[i]; t < t1; ++t) {
    const double u = WS(sl + L.rowpos[t]);
    const int kc = L.rowcol[t];
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (r < nr) acc[r] = fma(-u, WS(sr + kc * nrhs + r), acc[r]);
  }
  const double inv = WS(sl + P.colptr[i]);
#pragma unroll
  for (int r = 0; r < RB; ++r)
    if (r < nr) WS(sr + i * nrhs + r) = acc[r] * inv;
}

template <int RB>
__global__ void __launch_bounds__(128) k_backward_level(DevPlan P, DevLevels L, double* __restrict__ ws, int64_t NC, int64_t ncand,
                                                        int row0, int nrows, int col0, int ncols, int ngroups) {
  const int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int item = blockIdx.x / ngroups, grp = blockIdx.x - item * ngroups;
  if (item >= nrows) return;
  const int j = L.cols[row0 + item];
  const int p0 = P.colptr[j], p1 = P.colptr[j + 1];
  if (p1 - p0 == 1) return;
  const int cb = col0 + grp * RB;
  const int nr = min(RB, col0 + ncols - cb);
  const int nrhs = P.rhs.nrhs;
  const int sl = P.slot.s_l, sr = P.slot.s_r + cb;
  double acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = 0.0;
#pragma unroll(RB > 4 ? 4 : 8)
  for (int p = p0 + 1; p < p1; ++p) {
    const double u = WS(sl + p);
    const int i = P.rowidx[p];
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (r < nr) acc[r] = fma(u, WS(sr + i * nrhs + r), acc[r]);
  }
  const double inv = WS(sl + p0);
#pragma unroll
  for (int r = 0; r < RB; ++r)
    if (r < nr) WS(sr + j * nrhs + r) = fma(-inv, acc[r], WS(sr + j * nrhs + r));
}

// Narrow levels (the top of the tree: one or two long rows per level) leave the kernels above latency bound: few
// threads, each walking a long dependent list.  The *_seg variants split the list of one item over SEG threads per
// candidate (a block = 32 candidates x SEG segments) and reduce through shared memory in a fixed order.
template <int SEG>
__global__ void __launch_bounds__(32 * SEG) k_factor_level_seg(DevPlan P, DevLevels L, double* __restrict__ ws, int64_t NC,
                                                               int64_t ncand, int item0, int32_t* __restrict__ status) {
  __shared__ double red[SEG][32];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.y * 32 + lane;
  const bool ok = c < ncand;
  const int sl = P.slot.s_l;
  const int it = item0 + blockIdx.x;
  const int pe = L.items[it];
  const int p = pe & 0x7fffffff;
  const int d0 = L.dotptr[it], d1 = L.dotptr[it + 1];
  double v = 0.0;
  if (ok) {
#pragma unroll 4
    for (int t = d0 + seg; t < d1; t += SEG) {
      const int a = L.dot[3 * t], b = L.dot[3 * t + 1], dg = L.dot[3 * t + 2];
      v = fma(-(WS(sl + a) * WS(sl + dg)), WS(sl + b), v);
    }
  }
  red[seg][lane] = v;
  __syncthreads();
  if (seg == 0 && ok) {
    double tot = WS(sl + p);
#pragma unroll
    for (int g = 0; g < SEG; ++g) tot += red[g][lane];
    if (pe < 0) {
      if (!(tot > 0.0)) atomicOr(status + c, TNO_STATUS_NOT_PD);
      tot = 1.0 / tot;
    }
    WS(sl + p) = tot;
  }
}

template <int RB, int SEG, bool BACKWARD>
__global__ void __launch_bounds__(32 * SEG) k_solve_level_seg(DevPlan P, DevLevels L, double* __restrict__ ws, int64_t NC,
                                                              int64_t ncand, int row0, int nrows, int col0, int ncols, int ngroups) {
  __shared__ double red[SEG][RB][32];
  const int lane = threadIdx.x & 31, seg = threadIdx.x >> 5;
  const int64_t c = (int64_t)blockIdx.y * 32 + lane;
  const bool ok = c < ncand;
  const int item = blockIdx.x / ngroups, grp = blockIdx.x - item * ngroups;
  const int i = L.cols[row0 + item];
  const int cb = col0 + grp * RB;
  const int nr = min(RB, col0 + ncols - cb);
  const int nrhs = P.rhs.nrhs;
  const int sl = P.slot.s_l, sr = P.slot.s_r + cb;
  const int pd = P.colptr[i];
  double acc[RB];
#pragma unroll
  for (int r = 0; r < RB; ++r) acc[r] = 0.0;
  if (ok) {
    if (BACKWARD) {
      const int p1 = P.colptr[i + 1];
#pragma unroll 4
      for (int p = pd + 1 + seg; p < p1; p += SEG) {
        const double u = WS(sl + p);
        const int ii = P.rowidx[p];
#pragma unroll
        for (int r = 0; r < RB; ++r)
          if (r < nr) acc[r] = fma(u, WS(sr + ii * nrhs + r), acc[r]);
      }
    } else {
      const int t1 = L.rowptr[i + 1];
#pragma unroll 4
      for (int t = L.rowptr[i] + seg; t < t1; t += SEG) {
        const double u = WS(sl + L.rowpos[t]);
        const int kc = L.rowcol[t];
#pragma unroll
        for (int r = 0; r < RB; ++r)
          if (r < nr) acc[r] = fma(u, WS(sr + kc * nrhs + r), acc[r]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < RB; ++r) red[seg][r][lane] = acc[r];
  __syncthreads();
  if (seg == 0 && ok) {
    const double inv = WS(sl + pd);
#pragma unroll
    for (int r = 0; r < RB; ++r) {
      if (r >= nr) continue;
      double tot = 0.0;
#pragma unroll
      for (int g = 0; g < SEG; ++g) tot += red[g][r][lane];
      const double rhs = WS(sr + i * nrhs + r);
      WS(sr + i * nrhs + r) = BACKWARD ? fma(-inv, tot, rhs) : (rhs - tot) * inv;
    }
  }
}

__device__ __forceinline__ double zfixed(const DevPlan& P, const double* ws, int64_t NC, int64_t c, int b) {
  return P.var.o_zb >= 0 ? WS(P.slot.s_x + P.var.o_zb + b) : P.X[3 * P.fixed[b] + 2];
}
// x (axis 0) or y (axis 1) of support b: a variable with xyb (x block first, then y: constraints.py:52-55)
__device__ __forceinline__ double xyfixed(const DevPlan& P, const double* ws, int64_t NC, int64_t c, int b, int axis) {
  return P.var.o_xyb >= 0 ? WS(P.slot.s_x + P.var.o_xyb + axis * P.nb + b) : P.X[3 * P.fixed[b] + axis];
}

// phase-1 right-hand sides of  -(Ci^T Q Ci) X = -(P_i - Ci^T Q Cb X_b)  and the zb / lambdv blocks
__global__ void k_rhs1(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, int rtile) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int r0 = blockIdx.y * rtile;
  const int r1 = min(P.ni, r0 + rtile);
  const int nrhs = P.rhs.nrhs;
  const double lamh = P.var.o_lambdh >= 0 ? WS(P.slot.s_x + P.var.o_lambdh) : 1.0;
  const double lamv = P.var.o_lambdv >= 0 ? WS(P.slot.s_x + P.var.o_lambdv) : 1.0;
  const double pzs = WS(P.slot.s_pzs);
  const double dc = DIR_C, ds = DIR_S;
  for (int r = r0; r < r1; ++r) {
    const int v = P.row_vertex[r];
    double px, py, pz;
    if (P.var.o_lambdh >= 0) {
      px = lamh * mix2(P.px0, P.px0_b, v, dc, ds);
      py = lamh * mix2(P.py0, P.py0_b, v, dc, ds);
    } else {
      px = P.P[3 * v + 0];
      py = P.P[3 * v + 1];
    }
    pz = (P.var.o_lambdv >= 0) ? (lamv * P.pzv[v] + P.pz0[v]) : P.P[3 * v + 2];
    pz *= pzs;
    double bx = -px, by = -py, bz = -pz;
    const int base = P.slot.s_r + r * nrhs;
    if (P.rhs.c_zb >= 0)
      for (int b = 0; b < P.nb; ++b) WS(base + P.rhs.c_zb + b) = 0.0;
    for (int t = P.vadj_ptr[v]; t < P.vadj_ptr[v + 1]; ++t) {
      const int o = P.vadj_other[t];
      const int ro = P.vrow[o];
      if (ro >= 0) continue;
      const int b = -1 - ro;
      const double qe = WS(P.slot.s_q + P.vadj_edge[t]);
      bx = fma(-qe, xyfixed(P, ws, NC, c, b, 0), bx);
      by = fma(-qe, xyfixed(P, ws, NC, c, b, 1), by);
      bz = fma(-qe, zfixed(P, ws, NC, c, b), bz);
      if (P.rhs.c_zb >= 0) WS(base + P.rhs.c_zb + b) -= qe;
    }
    WS(base + P.rhs.cx) = bx;
    WS(base + P.rhs.cy) = by;
    WS(base + P.rhs.cz) = bz;
    if (P.rhs.c_lambdv >= 0) WS(base + P.rhs.c_lambdv) = -P.pzv[v] * pzs;   // d(pz)/d(lambdv) = pz_scale * pzv
  }
}

// L D L^T solve of RB right-hand-side columns [col0, col0+RB) (guarded by ncols) for one candidate
template <int RB>
__global__ void k_solve(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, int col0, int ncols) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int cb = col0 + blockIdx.y * RB;
  const int nr = min(RB, col0 + ncols - cb);
  const int nrhs = P.rhs.nrhs;
  const int sl = P.slot.s_l, sr = P.slot.s_r + cb;
  double y[RB];
  // forward (unit lower) fused with the diagonal scaling
  for (int j = 0; j < P.ni; ++j) {
    const int p0 = P.colptr[j], p1 = P.colptr[j + 1];
#pragma unroll
    for (int r = 0; r < RB; ++r) y[r] = (r < nr) ? WS(sr + j * nrhs + r) : 0.0;
    for (int p = p0 + 1; p < p1; ++p) {
      const double l = WS(sl + p);
      const int i = P.rowidx[p];
#pragma unroll
      for (int r = 0; r < RB; ++r)
        if (r < nr) WS(sr + i * nrhs + r) = fma(-l, y[r], WS(sr + i * nrhs + r));
    }
    const double inv = WS(sl + p0);
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (r < nr) WS(sr + j * nrhs + r) = y[r] * inv;
  }
  // backward (unit upper = L^T)
  for (int j = P.ni - 1; j >= 0; --j) {
    const int p0 = P.colptr[j], p1 = P.colptr[j + 1];
    if (p1 - p0 == 1) continue;
#pragma unroll
    for (int r = 0; r < RB; ++r) y[r] = (r < nr) ? WS(sr + j * nrhs + r) : 0.0;
    for (int p = p0 + 1; p < p1; ++p) {
      const double l = WS(sl + p);
      const int i = P.rowidx[p];
#pragma unroll
      for (int r = 0; r < RB; ++r)
        if (r < nr) y[r] = fma(-l, WS(sr + i * nrhs + r), y[r]);
    }
#pragma unroll
    for (int r = 0; r < RB; ++r)
      if (r < nr) WS(sr + j * nrhs + r) = y[r];
  }
}

__device__ __forceinline__ double zval(const DevPlan& P, const double* ws, int64_t NC, int64_t c, int v) {
  const int r = P.vrow[v];
  return r >= 0 ? WS(P.slot.s_r + r * P.rhs.nrhs + P.rhs.cz) : zfixed(P, ws, NC, c, -1 - r);
}
__device__ __forceinline__ double xval(const DevPlan& P, const double* ws, int64_t NC, int64_t c, int v, int axis) {
  const int r = P.vrow[v];
  return r >= 0 ? WS(P.slot.s_r + r * P.rhs.nrhs + axis) : xyfixed(P, ws, NC, c, -1 - r, axis);
}

// phase-2 right-hand sides:  Ci^T W B  (k columns) and  Ci^T W d0  (lambdh column), W = diag(C z)
constexpr int RHS2_JT = 32;   // columns per thread of k_rhs2 (= the launch's jtile)
__global__ void k_rhs2(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, int rtile, int jtile) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int r0 = blockIdx.y * rtile;
  const int r1 = min(P.ni, r0 + rtile);
  const int nrhs = P.rhs.nrhs, k = P.k;
  const int j0 = blockIdx.z * jtile, j1 = min(k, j0 + jtile);   // column tile: problems with hundreds of independents
  const double dc = DIR_C, ds = DIR_S;
  for (int r = r0; r < r1; ++r) {
    const int v = P.row_vertex[r];
    const int base = P.slot.s_r + r * nrhs;
    const double zv = WS(base + P.rhs.cz);
    const int a0 = P.vadj_ptr[v], a1 = P.vadj_ptr[v + 1];
    // edges outside, the (at most RHS2_JT) columns of the tile in registers: the end-point heights and the row of B of
    // an edge are read once per edge instead of once per (edge, column)
    {
      double acc[RHS2_JT];
#pragma unroll
      for (int jj = 0; jj < RHS2_JT; ++jj) acc[jj] = 0.0;
      for (int t = a0; t < a1; ++t) {
        const double zo = zval(P, ws, NC, c, P.vadj_other[t]);
        const double sg = (double)P.vadj_sign[t];
        const double w = sg > 0 ? (zv - zo) : (zo - zv);  // (C z)_e
        const double sw = sg * w;
        const double* Brow = P.B + (int64_t)P.vadj_edge[t] * k + j0;
#pragma unroll
        for (int jj = 0; jj < RHS2_JT; ++jj)
          if (j0 + jj < j1) acc[jj] = fma(sw, Brow[jj], acc[jj]);
      }
#pragma unroll
      for (int jj = 0; jj < RHS2_JT; ++jj)
        if (j0 + jj < j1) WS(base + P.rhs.c_q + j0 + jj) = acc[jj];
    }
    if (P.rhs.c_lambdh >= 0 && blockIdx.z == 0) {
      double acc = 0.0;
      for (int t = a0; t < a1; ++t) {
        const double zo = zval(P, ws, NC, c, P.vadj_other[t]);
        const double sg = (double)P.vadj_sign[t];
        const double w = sg > 0 ? (zv - zo) : (zo - zv);
        acc = fma(sg * w, mix2(P.d, P.d_b, P.vadj_edge[t], dc, ds), acc);
      }
      WS(base + P.rhs.c_lambdh) = acc;
    }
    if (P.rhs.c_qx >= 0) {  // Ci^T U B and Ci^T V B, U = diag(C x), V = diag(C y)  (jacobian.py:173-186)
      const double xv = WS(base + P.rhs.cx), yv = WS(base + P.rhs.cy);
#pragma unroll 1
      for (int axis = 0; axis < 2; ++axis) {
        const double cv = axis == 0 ? xv : yv;
        double acc[RHS2_JT];
#pragma unroll
        for (int jj = 0; jj < RHS2_JT; ++jj) acc[jj] = 0.0;
        for (int t = a0; t < a1; ++t) {
          // C[e, v] (C x)_e = x_v - x_o whichever end v is
          const double du = cv - xval(P, ws, NC, c, P.vadj_other[t], axis);
          const double* Brow = P.B + (int64_t)P.vadj_edge[t] * k + j0;
#pragma unroll
          for (int jj = 0; jj < RHS2_JT; ++jj)
            if (j0 + jj < j1) acc[jj] = fma(du, Brow[jj], acc[jj]);
        }
        const int cq = axis == 0 ? P.rhs.c_qx : P.rhs.c_qy;
#pragma unroll
        for (int jj = 0; jj < RHS2_JT; ++jj)
          if (j0 + jj < j1) WS(base + cq + j0 + jj) = acc[jj];
      }
    }
  }
}

__device__ __forceinline__ void envelope_eval(const DevPlan& P, double x, double y, double t, double* ub, double* lb,
                                               double* dub, double* dlb) {
  // closed forms of the parametric envelopes (compas_tna is not vendored)
  const double dx = x - P.env[0], dy = y - P.env[1];
  double d2;
  if (P.envelope_kind == TNO_ENV_DOME) {
    d2 = dx * dx + dy * dy;
  } else {
    const double dd = (fabs(dy) >= fabs(dx)) ? dx : dy;
    d2 = dd * dd;
  }
  const double re = P.env[2] + 0.5 * t, ri = P.env[2] - 0.5 * t;
  const double u = sqrt(re * re - d2);
  *ub = u;
  *dub = re / (2.0 * u);
  const double in = ri * ri - d2;
  if (in > 0.0) {
    const double l = sqrt(in);
    *lb = l;
    *dlb = -ri / (2.0 * l);
  } else {
    *lb = P.env[3];
    *dlb = 0.0;
  }
}

// reactions, objective, gradient, reac_bounds rows, envelope update, feasibility margin, status
__global__ void k_epilogue(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, double* __restrict__ f_out,
                           double* __restrict__ gmin_out, int32_t* __restrict__ status, const double* __restrict__ bounds) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int n = P.n, m = P.m, nb = P.nb, k = P.k, nvar = P.var.nvar, nrhs = P.rhs.nrhs;
  const bool has_t = P.var.o_t >= 0;
  const double lamh = P.var.o_lambdh >= 0 ? WS(P.slot.s_x + P.var.o_lambdh) : 1.0;
  const double lamv = P.var.o_lambdv >= 0 ? WS(P.slot.s_x + P.var.o_lambdv) : 1.0;
  const double pzs = WS(P.slot.s_pzs);
  const double dc = DIR_C, ds = DIR_S;
  double gmin = 1e300;
  bool bad = false;

  if (P.constraints & TNO_CON_FUNICULAR) {
    for (int e = 0; e < m; ++e) {
      const double q = WS(P.slot.s_q + e);
      gmin = fmin(gmin, fmin(q - P.qmin[e], P.qmax[e] - q));
      bad |= !isfinite(q);
    }
  }
  if (has_t) {
    const double t = WS(P.slot.s_x + P.var.o_t);
    for (int v = 0; v < n; ++v) {
      double ub, lb, dub, dlb;
      envelope_eval(P, xval(P, ws, NC, c, v, 0), xval(P, ws, NC, c, v, 1), t, &ub, &lb, &dub, &dlb);
      WS(P.slot.s_env + v) = ub;
      WS(P.slot.s_env + n + v) = lb;
      WS(P.slot.s_env + 2 * n + v) = dub;
      WS(P.slot.s_env + 3 * n + v) = dlb;
    }
  }
  if (P.constraints & TNO_CON_ENVELOPE) {
    for (int v = 0; v < n; ++v) {
      const double z = zval(P, ws, NC, c, v);
      const double ub = has_t ? WS(P.slot.s_env + v) : (bounds ? bounds[c * 2 * n + v] : P.ub[v]);
      const double lb = has_t ? WS(P.slot.s_env + n + v) : (bounds ? bounds[c * 2 * n + n + v] : P.lb[v]);
      gmin = fmin(gmin, fmin(z - lb, ub - z));
      bad |= !isfinite(z);
    }
  }

  if (P.con.r_envxy >= 0) {
    for (int v = 0; v < n; ++v) {
      const double x = xval(P, ws, NC, c, v, 0), y = xval(P, ws, NC, c, v, 1);
      gmin = fmin(gmin, fmin(fmin(x - P.xlimits[2 * v], P.xlimits[2 * v + 1] - x), fmin(y - P.ylimits[2 * v], P.ylimits[2 * v + 1] - y)));
      bad |= !isfinite(x) || !isfinite(y);
    }
  }

  const bool want_thrust = (P.objective == TNO_OBJ_MIN || P.objective == TNO_OBJ_MAX);
  const bool want_reac = (P.constraints & TNO_CON_REAC_BOUNDS) != 0;
  double f = 0.0;
  if (want_thrust)
    for (int j = 0; j < nvar; ++j) WS(P.slot.s_grad + j) = 0.0;

  if (want_thrust || want_reac) {
    const double osign = (P.objective == TNO_OBJ_MAX) ? -1.0 : 1.0;
    for (int b = 0; b < nb; ++b) {
      const int vb = P.fixed[b];
      const int a0 = P.vadj_ptr[vb], a1 = P.vadj_ptr[vb + 1];
      const double xb = xyfixed(P, ws, NC, c, b, 0), yb = xyfixed(P, ws, NC, c, b, 1), zb = zfixed(P, ws, NC, c, b);
      double px, py, pz;
      if (P.var.o_lambdh >= 0) {
        px = lamh * mix2(P.px0, P.px0_b, vb, dc, ds);
        py = lamh * mix2(P.py0, P.py0_b, vb, dc, ds);
      } else {
        px = P.P[3 * vb + 0];
        py = P.P[3 * vb + 1];
      }
      pz = (P.var.o_lambdv >= 0) ? (lamv * P.pzv[vb] + P.pz0[vb]) : P.P[3 * vb + 2];
      pz *= pzs;
      // R_b = Cb^T Q C X - P_b
      double Rx = 0.0, Ry = 0.0, Rz = 0.0;
      for (int t = a0; t < a1; ++t) {
        const int o = P.vadj_other[t];
        const double sg = (double)P.vadj_sign[t];
        const double qe = WS(P.slot.s_q + P.vadj_edge[t]);
        const double xo = xval(P, ws, NC, c, o, 0), yo = xval(P, ws, NC, c, o, 1), zo = zval(P, ws, NC, c, o);
        const double u = sg > 0 ? (xb - xo) : (xo - xb);
        const double vv = sg > 0 ? (yb - yo) : (yo - yb);
        const double w = sg > 0 ? (zb - zo) : (zo - zb);
        Rx = fma(sg * qe, u, Rx);
        Ry = fma(sg * qe, vv, Ry);
        Rz = fma(sg * qe, w, Rz);
      }
      Rx -= px;
      Ry -= py;
      Rz -= pz;
      if (want_thrust) {
        const double Rn = sqrt(Rx * Rx + Ry * Ry);
        f += Rn;
        const double cx = osign * Rx / Rn, cy = osign * Ry / Rn;
        // the q block of the gradient is formed by k_thrust_grad (a thread per candidate and column tile)
        WS(P.slot.s_cxy + b) = cx;
        WS(P.slot.s_cxy + nb + b) = cy;
        if (P.var.o_xyb >= 0) {
          // xb / yb blocks: (Rx / R)^T Cb^T Q C A and (Ry / R)^T Cb^T Q C A, A = [A^-1 (-Ci^T Q Cb); I]  (derivatives.py:310-326)
          for (int cb = 0; cb < nb; ++cb) {
            double acc = 0.0;
            for (int t = a0; t < a1; ++t) {
              const int ro = P.vrow[P.vadj_other[t]];
              const double Ao = ro >= 0 ? WS(P.slot.s_r + ro * nrhs + P.rhs.c_zb + cb) : ((-1 - ro) == cb ? 1.0 : 0.0);
              acc = fma(WS(P.slot.s_q + P.vadj_edge[t]), (cb == b ? 1.0 : 0.0) - Ao, acc);
            }
            WS(P.slot.s_grad + P.var.o_xyb + cb) += cx * acc;
            WS(P.slot.s_grad + P.var.o_xyb + nb + cb) += cy * acc;
          }
        }
      }
      if (want_reac) {
        // constraints.py:146-149 and jacobian.py:204-255, 310-346, 355-373 (restated literally)
        const int vbv = vb;
        const double sx = Rx < 0 ? -1.0 : 1.0, sy = Ry < 0 ? -1.0 : 1.0, sz = Rz < 0 ? -1.0 : 1.0;
        const double rz2 = Rz * Rz;
        const double rise = zb - (P.s ? P.s[vbv] : 0.0);
        const double gx_ = fabs(P.b[2 * b + 0]) - rise * fabs(Rx / Rz);
        const double gy_ = fabs(P.b[2 * b + 1]) - rise * fabs(Ry / Rz);
        WS(P.slot.s_rg + b) = gx_;
        WS(P.slot.s_rg + nb + b) = gy_;
        gmin = fmin(gmin, fmin(gx_, gy_));
        bad |= !isfinite(gx_) || !isfinite(gy_);
        const int rowx = P.slot.s_rj + b * nvar, rowy = P.slot.s_rj + (nb + b) * nvar;
        for (int j = 0; j < nvar; ++j) {
          WS(rowx + j) = 0.0;
          WS(rowy + j) = 0.0;
        }
        const bool has_env = (P.constraints & TNO_CON_ENVELOPE) != 0;  // dz/dq is only formed with 'envelope'
        // q_ind columns
        for (int j = 0; j < k; ++j) {
          double dRx = 0.0, dRy = 0.0, dRz = 0.0;
          for (int t = a0; t < a1; ++t) {
            const int o = P.vadj_other[t];
            const int e = P.vadj_edge[t];
            const double sg = (double)P.vadj_sign[t];
            const double qe = WS(P.slot.s_q + e);
            const double xo = xval(P, ws, NC, c, o, 0), yo = xval(P, ws, NC, c, o, 1), zo = zval(P, ws, NC, c, o);
            const double u = sg > 0 ? (xb - xo) : (xo - xb);
            const double vv = sg > 0 ? (yb - yo) : (yo - yb);
            const double w = sg > 0 ? (zb - zo) : (zo - zb);
            const double Bej = P.B[(int64_t)e * k + j];
            dRx = fma(sg * u, Bej, dRx);
            dRy = fma(sg * vv, Bej, dRy);
            dRz = fma(sg * w, Bej, dRz);
            const int ro = P.vrow[o];
            if (has_env && ro >= 0) {
              const double dzo = WS(P.slot.s_r + ro * nrhs + P.rhs.c_q + j);
              // (C dz)_e with dz_b = 0:  head - tail
              dRz = fma(sg * qe, sg > 0 ? -dzo : dzo, dRz);
            }
            if (P.rhs.c_qx >= 0 && ro >= 0) {  // + Cb^T Q C dx/dq_ind, dy/dq_ind (formed with 'envelopexy', jacobian.py:206-207)
              dRx = fma(-qe, WS(P.slot.s_r + ro * nrhs + P.rhs.c_qx + j), dRx);
              dRy = fma(-qe, WS(P.slot.s_r + ro * nrhs + P.rhs.c_qy + j), dRy);
            }
          }
          WS(rowx + j) = zb * sx * (-Rz * dRx + Rx * dRz) / rz2 / sz;
          WS(rowy + j) = zb * sy * (-Rz * dRy + Ry * dRz) / rz2 / sz;
        }
        // zb columns:  dRz/dzb = Cb^T Q C A ;  the reference adds COLUMN b of it to row b
        if (P.var.o_zb >= 0) {
          for (int cb = 0; cb < nb; ++cb) {
            // entry (cb, b) of Cb^T Q C A : row cb = support cb
            const int vc = P.fixed[cb];
            double acc = 0.0;
            for (int t = P.vadj_ptr[vc]; t < P.vadj_ptr[vc + 1]; ++t) {
              const int o = P.vadj_other[t];
              const double sg = (double)P.vadj_sign[t];
              const double qe = WS(P.slot.s_q + P.vadj_edge[t]);
              const int ro = P.vrow[o];
              const double Ao = ro >= 0 ? WS(P.slot.s_r + ro * nrhs + P.rhs.c_zb + b) : ((-1 - ro) == b ? 1.0 : 0.0);
              const double Ac = (cb == b) ? 1.0 : 0.0;
              const double diff = sg > 0 ? (Ac - Ao) : (Ao - Ac);
              acc = fma(sg * qe, diff, acc);
            }
            double vx = sz * zb * fabs(Rx) / rz2 * acc;
            double vy = sz * zb * fabs(Ry) / rz2 * acc;
            if (cb == b) {
              vx += -fabs(Rx / Rz);
              vy += -fabs(Ry / Rz);
            }
            WS(rowx + P.var.o_zb + cb) = vx;
            WS(rowy + P.var.o_zb + cb) = vy;
          }
        }
        if (P.var.o_lambdh >= 0) {
          double dRx = 0.0, dRy = 0.0, dRz = 0.0;
          for (int t = a0; t < a1; ++t) {
            const int o = P.vadj_other[t];
            const int e = P.vadj_edge[t];
            const double sg = (double)P.vadj_sign[t];
            const double qe = WS(P.slot.s_q + e);
            const double xo = xval(P, ws, NC, c, o, 0), yo = xval(P, ws, NC, c, o, 1), zo = zval(P, ws, NC, c, o);
            const double u = sg > 0 ? (xb - xo) : (xo - xb);
            const double vv = sg > 0 ? (yb - yo) : (yo - yb);
            const double w = sg > 0 ? (zb - zo) : (zo - zb);
            const double d0 = mix2(P.d, P.d_b, e, dc, ds);
            dRx = fma(sg * u, d0, dRx);
            dRy = fma(sg * vv, d0, dRy);
            dRz = fma(sg * w, d0, dRz);
            const int ro = P.vrow[o];
            if (ro >= 0) {
              const double dzo = WS(P.slot.s_r + ro * nrhs + P.rhs.c_lambdh);
              dRz = fma(sg * qe, sg > 0 ? -dzo : dzo, dRz);
            }
          }
          dRx -= mix2(P.px0, P.px0_b, vb, dc, ds);
          dRy -= mix2(P.py0, P.py0_b, vb, dc, ds);
          WS(rowx + P.var.o_lambdh) = zb * sx * (-Rz * dRx + Rx * dRz) / rz2 / sz;
          WS(rowy + P.var.o_lambdh) = zb * sy * (-Rz * dRy + Ry * dRz) / rz2 / sz;
        }
        if (P.var.o_lambdv >= 0) {
          double dRz = 0.0;
          for (int t = a0; t < a1; ++t) {
            const int o = P.vadj_other[t];
            const double sg = (double)P.vadj_sign[t];
            const double qe = WS(P.slot.s_q + P.vadj_edge[t]);
            const int ro = P.vrow[o];
            if (ro >= 0) {
              const double dzo = WS(P.slot.s_r + ro * nrhs + P.rhs.c_lambdv);
              dRz = fma(sg * qe, sg > 0 ? -dzo : dzo, dRz);
            }
          }
          dRz -= P.pzv[vb] * pzs;
          WS(rowx + P.var.o_lambdv) = sz * zb * fabs(Rx) / rz2 * dRz;
          WS(rowy + P.var.o_lambdv) = sz * zb * fabs(Ry) / rz2 * dRz;
        }
      }
    }
  }

  if (P.objective == TNO_OBJ_BESTFIT) {
    // f = sum_v (z_v - s_v)^2 ;  grad = (2 (z - s))^T [dz/dq_ind | dz/dzb]   (objectives.py:170-200, derivatives.py:355-413)
    f = 0.0;
    for (int v = 0; v < n; ++v) {
      const double dz = zval(P, ws, NC, c, v) - P.s[v];
      f = fma(dz, dz, f);
    }
    const int ncol = P.k + (P.var.o_zb >= 0 ? nb : 0);
    for (int j = 0; j < ncol; ++j) {
      const bool isq = j < P.k;
      const int pc = isq ? P.rhs.c_q + j : P.rhs.c_zb + (j - P.k);
      double acc = 0.0;
      if (!isq) {
        const int b = j - P.k;
        acc = 2.0 * (zfixed(P, ws, NC, c, b) - P.s[P.fixed[b]]);
      }
      for (int r = 0; r < P.ni; ++r) {
        const double w = 2.0 * (WS(P.slot.s_r + r * nrhs + P.rhs.cz) - P.s[P.row_vertex[r]]);
        acc = fma(w, WS(P.slot.s_r + r * nrhs + pc), acc);
      }
      WS(P.slot.s_grad + (isq ? j : P.var.o_zb + (j - P.k))) = acc;
    }
  }
  if (P.objective == TNO_OBJ_LOADPATH) {
    // f = sum_e |q_e| l_e^2 ;  grad_q = B^T (sign(q) l^2) + omega^T dz/dq_ind,  grad_zb = omega^T dz/dzb  with
    // omega = 2 C^T (|q| o w)  (objectives.py:240-275, derivatives.py:640-718; the dx/dq, dy/dq terms of the
    // reference vanish on a fixed plan because B spans the null space of the horizontal equilibrium)
    f = 0.0;
    for (int j = 0; j < nvar; ++j) WS(P.slot.s_grad + j) = 0.0;
    for (int e = 0; e < m; ++e) {
      const int a = P.edge_u[e], b = P.edge_v[e];
      const double u = xval(P, ws, NC, c, b, 0) - xval(P, ws, NC, c, a, 0);
      const double v = xval(P, ws, NC, c, b, 1) - xval(P, ws, NC, c, a, 1);
      const double w = zval(P, ws, NC, c, b) - zval(P, ws, NC, c, a);
      const double l2 = u * u + v * v + w * w;
      const double qe = WS(P.slot.s_q + e);
      f = fma(fabs(qe), l2, f);
      const double sl = qe > 0.0 ? l2 : (qe < 0.0 ? -l2 : 0.0);
      for (int j = 0; j < k; ++j) WS(P.slot.s_grad + j) = fma(sl, P.B[(int64_t)e * k + j], WS(P.slot.s_grad + j));
    }
    for (int v = 0; v < n; ++v) {
      const double zv = zval(P, ws, NC, c, v);
      double om = 0.0;
      for (int t = P.vadj_ptr[v]; t < P.vadj_ptr[v + 1]; ++t)
        om = fma(fabs(WS(P.slot.s_q + P.vadj_edge[t])), zv - zval(P, ws, NC, c, P.vadj_other[t]), om);
      om *= 2.0;
      const int r = P.vrow[v];
      if (r >= 0) {
        for (int j = 0; j < k; ++j)
          WS(P.slot.s_grad + j) = fma(om, WS(P.slot.s_r + r * nrhs + P.rhs.c_q + j), WS(P.slot.s_grad + j));
        if (P.var.o_zb >= 0)
          for (int b = 0; b < nb; ++b)
            WS(P.slot.s_grad + P.var.o_zb + b) = fma(om, WS(P.slot.s_r + r * nrhs + P.rhs.c_zb + b), WS(P.slot.s_grad + P.var.o_zb + b));
      } else if (P.var.o_zb >= 0) {
        WS(P.slot.s_grad + P.var.o_zb + (-1 - r)) += om;
      }
    }
  }
  if (P.objective == TNO_OBJ_ECOMP_LINEAR || P.objective == TNO_OBJ_ECOMP_NONLINEAR) {
    // f = -sum(R o dXb), R = Cb^T Q C X - P_b ;  grad = -sum_b dXb[b] . dR_b/dx   (objectives.py:278-340,
    // derivatives.py:503-620; dx/dq_ind, dy/dq_ind vanish on a fixed plan).  'simplified' quadratic term: sum stiff q^2.
    f = 0.0;
    for (int j = 0; j < nvar; ++j) WS(P.slot.s_grad + j) = 0.0;
    for (int b = 0; b < nb; ++b) {
      const int vb = P.fixed[b];
      const int a0 = P.vadj_ptr[vb], a1 = P.vadj_ptr[vb + 1];
      const double xb = xyfixed(P, ws, NC, c, b, 0), yb = xyfixed(P, ws, NC, c, b, 1), zb = zfixed(P, ws, NC, c, b);
      const double dx = P.dXb[3 * b + 0], dy = P.dXb[3 * b + 1], dz = P.dXb[3 * b + 2];
      double Rx = -P.P[3 * vb + 0], Ry = -P.P[3 * vb + 1], Rz = -pzs * P.P[3 * vb + 2];
      for (int t = a0; t < a1; ++t) {
        const int o = P.vadj_other[t];
        const double qe = WS(P.slot.s_q + P.vadj_edge[t]);  // C[e, vb] (C X)_e = X_b - X_o whichever end the support is
        Rx = fma(qe, xb - xval(P, ws, NC, c, o, 0), Rx);
        Ry = fma(qe, yb - xval(P, ws, NC, c, o, 1), Ry);
        Rz = fma(qe, zb - zval(P, ws, NC, c, o), Rz);
      }
      f -= Rx * dx + Ry * dy + Rz * dz;
      for (int j = 0; j < k; ++j) {
        double dRx = 0.0, dRy = 0.0, dRz = 0.0;
        for (int t = a0; t < a1; ++t) {
          const int o = P.vadj_other[t];
          const int e = P.vadj_edge[t];
          const double Bej = P.B[(int64_t)e * k + j];
          dRx = fma(xb - xval(P, ws, NC, c, o, 0), Bej, dRx);
          dRy = fma(yb - xval(P, ws, NC, c, o, 1), Bej, dRy);
          dRz = fma(zb - zval(P, ws, NC, c, o), Bej, dRz);
          const int ro = P.vrow[o];
          if (ro >= 0) dRz = fma(-WS(P.slot.s_q + e), WS(P.slot.s_r + ro * nrhs + P.rhs.c_q + j), dRz);
        }
        WS(P.slot.s_grad + j) -= dx * dRx + dy * dRy + dz * dRz;
      }
      if (P.var.o_zb >= 0)
        for (int cb = 0; cb < nb; ++cb) {
          double acc = 0.0;
          for (int t = a0; t < a1; ++t) {
            const int ro = P.vrow[P.vadj_other[t]];
            const double Ao = ro >= 0 ? WS(P.slot.s_r + ro * nrhs + P.rhs.c_zb + cb) : ((-1 - ro) == cb ? 1.0 : 0.0);
            acc = fma(WS(P.slot.s_q + P.vadj_edge[t]), (cb == b ? 1.0 : 0.0) - Ao, acc);
          }
          WS(P.slot.s_grad + P.var.o_zb + cb) -= dz * acc;
        }
    }
    if (P.objective == TNO_OBJ_ECOMP_NONLINEAR)
      for (int e = 0; e < m; ++e) {
        const double qe = WS(P.slot.s_q + e), se = P.stiff[e];
        f = fma(se * qe, qe, f);
        for (int j = 0; j < k; ++j) WS(P.slot.s_grad + j) = fma(2.0 * se * qe, P.B[(int64_t)e * k + j], WS(P.slot.s_grad + j));
      }
  }
  if (P.objective == TNO_OBJ_MAX) f = -f;
  if (P.objective == TNO_OBJ_T) f = WS(P.slot.s_x + nvar - 1);
  if (P.objective == TNO_OBJ_NEG_LAST) f = -WS(P.slot.s_x + nvar - 1);
  if (P.objective == TNO_OBJ_FEASIBILITY) f = 1.0;
  bad |= !isfinite(f);
  if (f_out) f_out[c] = f;
  if (gmin_out) gmin_out[c] = gmin;
  if (bad) status[c] |= TNO_STATUS_NONFINITE;
}

// q block of the thrust gradient (derivatives.py:303-309):
//   grad_j = sum_b  cx_b (Cb^T U B + Cb^T Q C dx/dq_ind)[b, j] + cy_b (Cb^T V B + Cb^T Q C dy/dq_ind)[b, j]
// with cx_b = +-Rx_b / |R_h,b|, cy_b likewise (left in the scratch by k_epilogue).  One thread per candidate and tile of
// columns, so that problems with hundreds of independents do not serialise k x nb x degree terms in one thread.
constexpr int TG_JT = 16;   // columns per thread of k_thrust_grad (= the launch's jtile)
__global__ void k_thrust_grad(DevPlan P, double* __restrict__ ws, int64_t NC, int64_t ncand, int jtile) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ncand) return;
  const int k = P.k, nb = P.nb, nrhs = P.rhs.nrhs;
  const int j0 = blockIdx.y * jtile, j1 = min(k, j0 + jtile);
  // edge terms outside, the columns of the tile in registers: the geometry of a support edge is read once
  double acc[TG_JT];
#pragma unroll
  for (int jj = 0; jj < TG_JT; ++jj) acc[jj] = 0.0;
  for (int b = 0; b < nb; ++b) {
    const int vb = P.fixed[b];
    const double xb = xyfixed(P, ws, NC, c, b, 0), yb = xyfixed(P, ws, NC, c, b, 1);
    const double cx = WS(P.slot.s_cxy + b), cy = WS(P.slot.s_cxy + nb + b);
    for (int t = P.vadj_ptr[vb]; t < P.vadj_ptr[vb + 1]; ++t) {
      const int o = P.vadj_other[t];
      const int e = P.vadj_edge[t];
      // C[e, vb] (C x)_e = x_b - x_o whichever end the support is
      const double coef = cx * (xb - xval(P, ws, NC, c, o, 0)) + cy * (yb - xval(P, ws, NC, c, o, 1));
      const double* Brow = P.B + (int64_t)e * k + j0;
#pragma unroll
      for (int jj = 0; jj < TG_JT; ++jj)
        if (j0 + jj < j1) acc[jj] = fma(coef, Brow[jj], acc[jj]);
      const int ro = P.vrow[o];
      // + Cb^T Q C dx/dq_ind: the free neighbours move with q once the supports do.  Only with the xyb variable: on a
      // fixed plan the reference keeps dx/dq_ind = 0 in this gradient (derivatives.py:257-301), and the Dx / Dy panels of an
      // envelopexy-only problem are not solved unless J is requested.
      if (P.rhs.c_qx >= 0 && P.var.o_xyb >= 0 && ro >= 0) {
        const double qe = WS(P.slot.s_q + e);
        const int bx = P.slot.s_r + ro * nrhs + P.rhs.c_qx + j0, by = P.slot.s_r + ro * nrhs + P.rhs.c_qy + j0;
#pragma unroll
        for (int jj = 0; jj < TG_JT; ++jj)
          if (j0 + jj < j1) acc[jj] = fma(-qe, cx * WS(bx + jj) + cy * WS(by + jj), acc[jj]);
      }
    }
  }
#pragma unroll
  for (int jj = 0; jj < TG_JT; ++jj)
    if (j0 + jj < j1) WS(P.slot.s_grad + j0 + jj) = acc[jj];
}

// Per-candidate envelope: g[r_env + v] = z_v - lb_v, g[r_env + n + v] = ub_v - z_v with the candidate's own bounds.
__global__ void __launch_bounds__(256) k_bounds_rows(DevPlan P, const double* __restrict__ ws, int64_t NC, const double* __restrict__ bounds,
                                                     double* __restrict__ g) {
  const int v = blockIdx.y * blockDim.x + threadIdx.x;
  if (v >= P.n) return;
  const int64_t c = blockIdx.x;
  const int r = P.vrow[v];
  const double z = r >= 0 ? WS(P.slot.s_r + r * P.rhs.nrhs + P.rhs.cz)
                          : (P.var.o_zb >= 0 ? WS(P.slot.s_x + P.var.o_zb + (-1 - r)) : P.X[3 * v + 2]);
  const double* b = bounds + c * 2 * P.n;
  double* gc = g + c * P.con.ng + P.con.r_env;
  gc[v] = z - b[P.n + v];
  gc[P.n + v] = b[v] - z;
}

// Per-candidate load direction: rows e and m + e of the funicular block get +-(c d[e] + s d_b[e]) in the lambdh column.
__global__ void __launch_bounds__(256) k_dir_column(DevPlan P, const double* __restrict__ load_dir, double* __restrict__ J,
                                                    int64_t cand_stride) {
  const int e = blockIdx.y * blockDim.x + threadIdx.x;
  if (e >= P.m) return;
  const int64_t cand = blockIdx.x;
  const double v = fma(load_dir[2 * cand + 1], P.d_b[e], load_dir[2 * cand] * P.d[e]);
  double* Jc = J + cand * cand_stride + (int64_t)P.con.r_fun * P.var.nvar + P.var.o_lambdh;
  Jc[(int64_t)e * P.var.nvar] = v;
  Jc[(int64_t)(P.m + e) * P.var.nvar] = -v;
}

// Table-driven transposing emitter: out[c][e] = sa * ws[a][c] + sb * ws[b][c] + const.
// A CTA stages TW elements x 32 candidates in shared memory: reads are coalesced along the candidate
// axis of the scratch, writes are coalesced along the element axis of the user-facing row-major output.
constexpr int EMIT_TW = 128;
__global__ void __launch_bounds__(256) k_emit(const EmitDesc2* __restrict__ table, int64_t count, double* __restrict__ out,
                                              int64_t out_stride, const double* __restrict__ ws, int64_t NC, int64_t ncand) {
  __shared__ double tile[EMIT_TW][33];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t e0 = (int64_t)blockIdx.x * EMIT_TW;
  const int64_t cand0 = (int64_t)blockIdx.y * 32;
  const int ne = (int)min((int64_t)EMIT_TW, count - e0);
  const int64_t c = cand0 + lane;
  // four elements per pass: descriptors first, then all gathers, so that the loads of a pass overlap
  const bool cok = c < ncand;
  for (int eb = warp; eb < ne; eb += 32) {
    EmitDesc2 d[4];
    double va[4], vb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = eb + 8 * u;
      d[u] = e < ne ? table[e0 + e] : EmitDesc2{-1, -1, 0.f, 0.f, 0.0};
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      va[u] = (cok && d[u].slot_a >= 0) ? ws[(int64_t)d[u].slot_a * NC + c] : 0.0;
      vb[u] = (cok && d[u].slot_b >= 0) ? ws[(int64_t)d[u].slot_b * NC + c] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int e = eb + 8 * u;
      if (e < ne) tile[e][lane] = fma((double)d[u].sign_b, vb[u], fma((double)d[u].sign_a, va[u], d[u].c));
    }
  }
  __syncthreads();
  for (int cc = warp; cc < 32; cc += 8) {
    const int64_t cg = cand0 + cc;
    if (cg >= ncand) continue;
    double* o = out + cg * out_stride + e0;
    for (int e = lane; e < ne; e += 32) o[e] = tile[e][cc];
  }
}

__global__ void k_copy_status(const int32_t* __restrict__ src, int32_t* __restrict__ dst, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[i];
}

static inline dim3 grid1(int64_t ncand, int block, int gy = 1) {
  return dim3((unsigned)((ncand + block - 1) / block), (unsigned)gy, 1);
}

const char* const kKernelNames[KK_COUNT] = {"k_pack", "k_q", "k_assemble", "k_factor", "k_rhs1", "k_solve(phase1)", "k_rhs2",
                                            "k_solve(phase2)", "k_epilogue", "k_emit(g,grad,z,q)", "k_emit(J)", "k_copy_status",
                                            "k_onchip"};

// Level lists of the level-scheduled kernels (built on first use; the plan keeps them).
static int build_levels(Plan* PL) {
  const Symbolic& S = PL->sym;
  std::vector<int> items, dotptr(1, 0), dot;
  PL->lv_item_ptr.assign(S.nlevels + 1, 0);
  PL->lv_seg.clear();
  for (int l = 0; l < S.nlevels; ++l) {
    for (int q = S.level_ptr[l]; q < S.level_ptr[l + 1]; ++q) {
      const int j = S.level_cols[q];
      for (int p = S.colptr[j]; p < S.colptr[j + 1]; ++p) {
        items.push_back(p == S.colptr[j] ? (int)(0x80000000u | (unsigned)p) : p);
        for (int64_t t = S.dot_ptr[p]; t < S.dot_ptr[p + 1]; ++t) {
          dot.push_back(S.dot_a[t]);
          dot.push_back(S.dot_b[t]);
          dot.push_back(S.colptr[S.dot_c[t]]);
        }
        dotptr.push_back((int)(dot.size() / 3));
      }
    }
    PL->lv_item_ptr[l + 1] = (int)items.size();
    // long pair lists on few items: split every list over the threads of a block (k_factor_level_seg)
    const int nit = PL->lv_item_ptr[l + 1] - PL->lv_item_ptr[l];
    const int64_t npairs = (int64_t)dotptr.back() - dotptr[PL->lv_item_ptr[l]];
    PL->lv_seg.push_back(nit > 0 && npairs >= 16 * (int64_t)nit && nit < 4096);
  }
  if (dot.empty()) dot.push_back(0);
  int rc;
  if ((rc = upload_vec(PL, items, &PL->lv.items))) return rc;
  if ((rc = upload_vec(PL, dotptr, &PL->lv.dotptr))) return rc;
  if ((rc = upload_vec(PL, dot, &PL->lv.dot))) return rc;
  if ((rc = upload_vec(PL, S.level_cols, &PL->lv.cols))) return rc;
  if ((rc = upload_vec(PL, S.rowptr, &PL->lv.rowptr))) return rc;
  std::vector<int> rowcol = S.rowcol, rowpos = S.rowpos;
  if (rowcol.empty()) { rowcol.push_back(0); rowpos.push_back(0); }
  if ((rc = upload_vec(PL, rowcol, &PL->lv.rowcol))) return rc;
  if ((rc = upload_vec(PL, rowpos, &PL->lv.rowpos))) return rc;
  PL->lv_ready = true;
  return TNO_OK;
}

int run_interleaved(Plan* PL, const double* x_dev, int64_t N, int64_t ldx, const tno_batch_inputs* opt,
                    const tno_outputs* out, cudaStream_t stream) {
  const DevPlan& P = PL->dp;
  const Symbolic& S = PL->sym;
  const int64_t NC = PL->chunk;
  double* ws = PL->scratch;
  int32_t* st = PL->status_scratch;
  const int TB = 128;
  // level-scheduled factor / solves unless the plan (TNO_KERNEL_INTERLEAVED_TPC) or the developer knob asks for the
  // thread-per-candidate originals
  static const bool v1 = getenv("TNO_INTERLEAVED_V1") != nullptr;
  const bool levelled = !(v1 || PL->thread_per_candidate);
  if (levelled && !PL->lv_ready) {
    const int rc = build_levels(PL);
    if (rc) return rc;
  }
#define LAUNCH(kind, kernel, grid, block, ...)              \
  do {                                                      \
    PL->prof_begin(kind, stream);                           \
    kernel<<<grid, block, 0, stream>>>(__VA_ARGS__);        \
    PL->prof_end(stream);                                   \
    PL->launches += 1;                                      \
  } while (0)
  for (int64_t c0 = 0; c0 < N; c0 += NC) {
    const int64_t nc = std::min(NC, N - c0);
    const double* xs = x_dev + c0 * ldx;
    const double* pzs = (opt && opt->pz_scale) ? opt->pz_scale + c0 : nullptr;
    const double* bnd = (opt && opt->bounds) ? opt->bounds + 2 * c0 * P.n : nullptr;
    if (bnd && (P.var.o_t >= 0 || P.con.r_env < 0)) {
      set_error("per-candidate bounds need the envelope constraint and no thickness variable");
      return TNO_EUNSUPPORTED;
    }
    const double* ldir = (opt && opt->load_dir) ? opt->load_dir + 2 * c0 : nullptr;
    if (ldir && !P.d_b) {
      set_error("load_dir needs a plan created with the second load basis (d_b, px0_b, py0_b)");
      return TNO_EUNSUPPORTED;
    }
    LAUNCH(KK_PACK, k_pack, grid1(nc, TB), TB, P, ws, NC, nc, xs, ldx, pzs, ldir);
    const int etile = P.k > 64 ? 4 : 32;   // many independents: shorter edge tiles keep the machine busy
    LAUNCH(KK_Q, k_q, grid1(nc, TB, (P.m + etile - 1) / etile), TB, P, ws, NC, nc, etile);
    const int ptile = 128;
    LAUNCH(KK_ASSEMBLE, k_assemble, grid1(nc, TB, (P.nnz_l + ptile - 1) / ptile), TB, P, ws, NC, nc, ptile);
    const unsigned cblocks = (unsigned)((nc + TB - 1) / TB), cblocks32 = (unsigned)((nc + 31) / 32);
    constexpr int SEG = 8;
    if (levelled) {
      TNO_CUDA_OK(cudaMemsetAsync(st, 0, (size_t)nc * 4, stream));
      for (int l = 0; l < S.nlevels; ++l) {
        const int i0 = PL->lv_item_ptr[l], i1 = PL->lv_item_ptr[l + 1];
        // a few thousand blocks per launch are plenty; longer tiles amortise the block start-up on the wide leaf levels
        const int tile = std::max(1, (int)(((int64_t)(i1 - i0) * cblocks) / 16384));
        if (PL->lv_seg[l])
          LAUNCH(KK_FACTOR, k_factor_level_seg<SEG>, dim3((unsigned)(i1 - i0), cblocks32, 1), 32 * SEG, P, PL->lv, ws, NC, nc, i0, st);
        else
          LAUNCH(KK_FACTOR, k_factor_level, dim3((unsigned)((i1 - i0 + tile - 1) / tile), cblocks, 1), TB, P, PL->lv, ws, NC, nc, i0, i1, tile, st);
      }
    } else {
      LAUNCH(KK_FACTOR, k_factor, grid1(nc, TB), TB, P, ws, NC, nc, st);
    }
    const int rtile = P.k > 64 ? 8 : 32;
    const int gyr = (P.ni + rtile - 1) / rtile;
    LAUNCH(KK_RHS1, k_rhs1, grid1(nc, TB, gyr), TB, P, ws, NC, nc, rtile);
    constexpr int RB = 4;
    auto solve_levelled = [&](int kind, int col0, int ncols) {
      const int ngroups = (ncols + RB - 1) / RB;
      const auto seg_fwd = k_solve_level_seg<RB, SEG, false>;
      const auto seg_bwd = k_solve_level_seg<RB, SEG, true>;
      for (int pass = 0; pass < 2; ++pass) {
        // forward over ascending levels, backward over descending ones (the roots have no off-diagonals)
        for (int q = 0; q < S.nlevels - pass; ++q) {
          const int l = pass ? S.nlevels - 2 - q : q;
          const int r0 = S.level_ptr[l], nr = S.level_ptr[l + 1] - r0;
          const bool seg = (int64_t)nr * ngroups * cblocks < 4096;  // less than ~two waves of 128-thread blocks
          if (seg) {
            const dim3 g((unsigned)(nr * ngroups), cblocks32, 1);
            if (pass)
              LAUNCH(kind, seg_bwd, g, 32 * SEG, P, PL->lv, ws, NC, nc, r0, nr, col0, ncols, ngroups);
            else
              LAUNCH(kind, seg_fwd, g, 32 * SEG, P, PL->lv, ws, NC, nc, r0, nr, col0, ncols, ngroups);
          } else {
            const dim3 g((unsigned)(nr * ngroups), cblocks, 1);
            if (pass)
              LAUNCH(kind, k_backward_level<RB>, g, TB, P, PL->lv, ws, NC, nc, r0, nr, col0, ncols, ngroups);
            else
              LAUNCH(kind, k_forward_level<RB>, g, TB, P, PL->lv, ws, NC, nc, r0, nr, col0, ncols, ngroups);
          }
        }
      }
    };
    if (levelled)
      solve_levelled(KK_SOLVE1, 0, P.rhs.nrhs1);
    else
      LAUNCH(KK_SOLVE1, k_solve<RB>, grid1(nc, TB, (P.rhs.nrhs1 + RB - 1) / RB), TB, P, ws, NC, nc, 0, P.rhs.nrhs1);
    // the dz/dq_ind columns only feed the Jacobian and the bestfit gradient
    const bool need2 = P.rhs.nrhs2 > 0 && (out->J != nullptr || (P.objective >= TNO_OBJ_BESTFIT && P.objective != TNO_OBJ_FEASIBILITY && out->grad != nullptr) ||
                                            (P.rhs.c_qx >= 0 && P.var.o_xyb >= 0 && out->grad != nullptr));
    if (need2) {
      const int jtile = RHS2_JT;
      LAUNCH(KK_RHS2, k_rhs2, dim3((unsigned)((nc + TB - 1) / TB), (unsigned)gyr, (unsigned)((P.k + jtile - 1) / jtile)), TB, P, ws, NC, nc, rtile, jtile);
      if (levelled)
        solve_levelled(KK_SOLVE2, P.rhs.nrhs1, P.rhs.nrhs2);
      else
        LAUNCH(KK_SOLVE2, k_solve<RB>, grid1(nc, TB, (P.rhs.nrhs2 + RB - 1) / RB), TB, P, ws, NC, nc, P.rhs.nrhs1, P.rhs.nrhs2);
    }
    LAUNCH(KK_EPILOGUE, k_epilogue, grid1(nc, TB), TB, P, ws, NC, nc, out->f ? out->f + c0 : nullptr,
           out->gmin ? out->gmin + c0 : nullptr, st, bnd);
    if (out->grad && (P.objective == TNO_OBJ_MIN || P.objective == TNO_OBJ_MAX)) {
      const int jtile = TG_JT;
      LAUNCH(KK_EPILOGUE, k_thrust_grad, grid1(nc, TB, (P.k + jtile - 1) / jtile), TB, P, ws, NC, nc, jtile);
    }
    // stride = elements per candidate of the user-facing array, offset = first element the table describes
    auto emit = [&](int kind, const EmitTable& T, double* dst, int64_t stride = -1, int64_t offset = 0) {
      if (!dst || T.count == 0) return;
      if (stride < 0) stride = T.count;
      dim3 g((unsigned)((T.count + EMIT_TW - 1) / EMIT_TW), (unsigned)((nc + 31) / 32), 1);
      LAUNCH(kind, k_emit, g, 256, T.dev, T.count, dst + c0 * stride + offset, stride, ws, NC, nc);
    };
    emit(KK_EMIT_SMALL, PL->t_g, out->g);
    if (bnd && out->g)  // the table carries the plan's bounds: the envelope rows are rewritten with the candidate's own
      LAUNCH(KK_EMIT_SMALL, k_bounds_rows, dim3((unsigned)nc, (unsigned)((P.n + 255) / 256), 1), 256, P, ws, NC, bnd, out->g + c0 * P.con.ng);
    emit(KK_EMIT_SMALL, PL->t_grad, out->grad);
    emit(KK_EMIT_SMALL, PL->t_z, out->z);
    emit(KK_EMIT_SMALL, PL->t_q, out->q);
    if (out->J) {
      const int64_t jstride = (int64_t)P.con.ng * P.var.nvar;
      if (PL->jfun) launch_fill_funicular(PL, PL->jfun, PL->jfun_count, out->J + c0 * jstride, nc, stream);
      if (ldir && PL->jfun)  // the lambdh column of the funicular rows is +-d0 of the candidate's direction
        LAUNCH(KK_EMIT_J, k_dir_column, dim3((unsigned)nc, (unsigned)((P.m + 255) / 256), 1), 256, P, ldir, out->J + c0 * jstride, jstride);
      emit(KK_EMIT_J, PL->t_J, out->J, jstride, PL->jfun_count);
    }
    if (out->status) LAUNCH(KK_STATUS, k_copy_status, grid1(nc, 256), 256, st, out->status + c0, nc);
  }
#undef LAUNCH
  TNO_CUDA_OK(cudaGetLastError());
  return TNO_OK;
}

}  // namespace tno
