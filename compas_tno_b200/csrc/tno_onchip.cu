// Kernel family 2: "on-chip" -- one CTA per candidate, factor and right-hand-side panel resident in
// shared memory, one fused launch for the whole evaluation (persistent grid-stride over candidates).
//
// HBM traffic of a candidate is its algorithmic traffic only: x in; z, f, grad, g and the dense J out.
// The numeric factor (nnz_l doubles) and the panel R (ni x W doubles) never leave the SM.  Pattern
// indices are 16-bit streams shared by all candidates (L2 / L1 resident).
//
// Parallelism inside a candidate comes from the elimination tree: columns of one level set are
// independent, so the factorisation (pull / "dot product" form of L D L^T) and both triangular solves
// are level-synchronous sweeps; lanes of a sweep item run over right-hand-side columns (coalesced,
// conflict-free panel rows), narrow levels at the top of the tree split one dot product over several
// lanes and reduce with shuffles.
//
// Math restated from the reference, same citations as tno_interleaved.cu.
#include <algorithm>
#include <cstdio>

#include "tno_plan.hpp"

namespace tno {

struct OcArgs {
  const double* x;
  int64_t ldx;
  const double* pzs;
  double* f;
  double* grad;
  double* g;
  double* J;
  double* z;
  double* q;
  double* gmin;
  int32_t* status;
  int64_t N;
};

__device__ __forceinline__ void oc_envelope(const DevPlan& P, double x, double y, double t, double* ub, double* lb,
                                            double* dub, double* dlb) {
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

// Level-synchronous L D L^T solve of `ncols` panel columns.  PHASE 1 columns: [0, n1) then x (W-2), y (W-1);
// PHASE 2 columns: [c_q, c_q + ncols).  G lanes (power of two) cover the columns of one row.
template <int PHASE>
__device__ __forceinline__ void oc_solve(const DevPlan& P, const DevOnchip& O, const double* __restrict__ Lv,
                                         double* __restrict__ R, int ncols, int G, int tid, int nthreads) {
  const int W = O.W;
  const int r = tid & (G - 1);
  const int slot = tid / G;
  const int nslots = nthreads / G;
  const int cgroups = (ncols + G - 1) / G;
  auto colmap = [&](int cidx) -> int {
    if (PHASE == 1) return cidx < O.n1 ? cidx : (W - 2 + (cidx - O.n1));
    return O.c_q + cidx;
  };
  // forward: y_i = b_i - sum_j L_ij y_j   (unit lower, rows of one level are independent)
  for (int lev = 1; lev < O.nlevels; ++lev) {
    const int i0 = O.s_lev_ptr[lev], i1 = O.s_lev_ptr[lev + 1];
    const int nitems = (i1 - i0) * cgroups;
    for (int it = slot; it < nitems; it += nslots) {
      const int ri = it / cgroups;
      const int cidx = (it - ri * cgroups) * G + r;
      if (cidx < ncols) {
        const int row = O.s_row[i0 + ri];
        const int col = colmap(cidx);
        double acc = R[row * W + col];
        const int t0 = O.rowptr[row], t1 = O.rowptr[row + 1];
#pragma unroll 4
        for (int t = t0; t < t1; ++t) {
          const uint32_t pc = __ldg(O.row_pc + t);
          acc = fma(-Lv[pc & 0xffffu], R[(pc >> 16) * W + col], acc);
        }
        R[row * W + col] = acc;
      }
    }
    __syncthreads();
  }
  // backward: x_j = y_j / d_j - sum_{i > j} L_ij x_i
  for (int lev = O.nlevels - 1; lev >= 0; --lev) {
    const int i0 = O.s_lev_ptr[lev], i1 = O.s_lev_ptr[lev + 1];
    const int nitems = (i1 - i0) * cgroups;
    for (int it = slot; it < nitems; it += nslots) {
      const int ri = it / cgroups;
      const int cidx = (it - ri * cgroups) * G + r;
      if (cidx < ncols) {
        const int j = O.s_row[i0 + ri];
        const int col = colmap(cidx);
        const int p0 = P.colptr[j], p1 = P.colptr[j + 1];
        double acc = R[j * W + col] * Lv[p0];
#pragma unroll 4
        for (int p = p0 + 1; p < p1; ++p) acc = fma(-Lv[p], R[(int)__ldg(O.rowidx16 + p) * W + col], acc);
        R[j * W + col] = acc;
      }
    }
    __syncthreads();
  }
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_onchip(const __grid_constant__ DevPlan P, const __grid_constant__ DevOnchip O,
                                                            const __grid_constant__ OcArgs A) {
  extern __shared__ double sm[];
  double* xs = sm + O.o_x;
  double* Lv = sm + O.o_l;
  double* R = sm + O.o_r;
  double* qs = R;  // q lives in the panel region until the matrix is assembled
  double* qsup = sm + O.o_small + O.o_qsup;
  double* react = sm + O.o_small + O.o_react;  // nb x 3
  double* gxy = sm + O.o_small + O.o_gxy;      // 2 x nb x k : Cb^T U B, Cb^T V B
  double* red = sm + O.o_small + O.o_red;      // THREADS / 32 + 2
  double* envd = O.o_env >= 0 ? sm + O.o_env : nullptr;  // dub[n] | dlb[n]
  __shared__ int s_flags;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int n = P.n, m = P.m, ni = P.ni, nb = P.nb, k = P.k, nvar = P.var.nvar, ng = P.con.ng, W = O.W;
  const bool has_t = P.var.o_t >= 0;
  const bool want_thrust = (P.objective == TNO_OBJ_MIN || P.objective == TNO_OBJ_MAX);
  const bool want_reac = (P.constraints & TNO_CON_REAC_BOUNDS) != 0;

  for (int64_t cand = blockIdx.x; cand < A.N; cand += gridDim.x) {
    // ---- 0. variables ------------------------------------------------------------------------------
    if (tid < nvar) xs[tid] = A.x[cand * A.ldx + tid];
    if (tid == 0) s_flags = 0;
    __syncthreads();
    const double lamh = P.var.o_lambdh >= 0 ? xs[P.var.o_lambdh] : 1.0;
    const double lamv = P.var.o_lambdv >= 0 ? xs[P.var.o_lambdv] : 1.0;
    const double pzs = A.pzs ? A.pzs[cand] : 1.0;
    double gmin = 1e300;
    bool bad = false;
    double* gout = A.g ? A.g + cand * ng : nullptr;
    double* Jout = A.J ? A.J + cand * (int64_t)ng * nvar : nullptr;

    // ---- 1. q = B q_ind + lambdh d ; funicular rows of g -------------------------------------------
    for (int e = tid; e < m; e += THREADS) {
      double acc = lamh * P.d[e];
      for (int j = 0; j < k; ++j) acc = fma(__ldg(O.Bt + (int64_t)j * m + e), xs[j], acc);
      qs[e] = acc;
      if (A.q) A.q[cand * m + e] = acc;
      if (P.con.r_fun >= 0) {
        const double g0 = acc - P.qmin[e], g1 = P.qmax[e] - acc;
        if (gout) {
          gout[P.con.r_fun + e] = g0;
          gout[P.con.r_fun + m + e] = g1;
        }
        gmin = fmin(gmin, fmin(g0, g1));
      }
      bad |= !isfinite(acc);
    }
    __syncthreads();

    // ---- 2. assemble -(Ci^T Q Ci) into the pattern of L; keep q of the support edges ----------------
    for (int p = tid; p < P.nnz_l; p += THREADS) {
      double acc = 0.0;
      for (int t = P.asm_ptr[p]; t < P.asm_ptr[p + 1]; ++t) acc += (double)P.asm_coef[t] * qs[P.asm_edge[t]];
      Lv[p] = acc;
    }
    for (int t = tid; t < O.nsup; t += THREADS) qsup[t] = qs[O.sup_edge[t]];
    __syncthreads();

    // ---- 3. numeric L D L^T, level by level (off-diagonals stay unscaled until step 4) --------------
    for (int lev = 0; lev < O.nlevels; ++lev) {
      const int i0 = O.f_lev_ptr[lev], i1 = O.f_lev_ptr[lev + 1];
      const int lpi = O.f_lev_lpi[lev];
      const int sub = tid & (lpi - 1);
      const int per_pass = THREADS / lpi;
      const int npass = (i1 - i0 + per_pass - 1) / per_pass;
      for (int ps = 0; ps < npass; ++ps) {
        const int it = i0 + ps * per_pass + tid / lpi;
        const bool valid = it < i1;
        double acc = 0.0;
        int pos = 0, j = 0;
        if (valid) {
          pos = O.f_item_pos[it];
          j = O.f_item_col[it];
          const int t0 = O.f_item_pptr[it], t1 = O.f_item_pptr[it + 1];
          for (int t = t0 + sub; t < t1; t += lpi) {
            const unsigned long long pr = __ldg(O.f_pairs + t);
            const int a = (int)(pr & 0xffffu), b = (int)((pr >> 16) & 0xffffu), dp = (int)((pr >> 32) & 0xffffu);
            acc = fma(Lv[a] * Lv[dp], Lv[b], acc);  // U_ic * (1/d_c) * U_jc
          }
        }
        for (int off = lpi >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (valid && sub == 0) {
          const double v = Lv[pos] - acc;
          if (pos == P.colptr[j]) {
            if (!(v > 0.0)) atomicOr(&s_flags, TNO_STATUS_NOT_PD);
            Lv[pos] = 1.0 / v;
          } else {
            Lv[pos] = v;
          }
        }
      }
      __syncthreads();
    }
    // ---- 4. L_ij = U_ij / d_j ----------------------------------------------------------------------
    for (int p = tid; p < P.nnz_l; p += THREADS) {
      const int dp = O.ent_dp[p];
      if (dp != p) Lv[p] *= Lv[dp];
    }
    // (the panel region is free from here on: q has been consumed)
    // ---- 5. phase-1 right-hand sides ----------------------------------------------------------------
    for (int r = tid; r < ni; r += THREADS) {
      const int v = P.row_vertex[r];
      double px, py, pz;
      if (P.var.o_lambdh >= 0) {
        px = lamh * P.px0[v];
        py = lamh * P.py0[v];
      } else {
        px = P.P[3 * v + 0];
        py = P.P[3 * v + 1];
      }
      pz = (P.var.o_lambdv >= 0) ? (lamv * P.pzv[v] + P.pz0[v]) : P.P[3 * v + 2];
      pz *= pzs;
      double bx = -px, by = -py, bz = -pz;
      double* row = R + r * W;
      if (O.c_zb >= 0)
        for (int b = 0; b < nb; ++b) row[O.c_zb + b] = 0.0;
      for (int t = P.vadj_ptr[v]; t < P.vadj_ptr[v + 1]; ++t) {
        const int o = P.vadj_other[t];
        const int ro = P.vrow[o];
        if (ro >= 0) continue;
        const int b = -1 - ro;
        const double qe = qsup[O.edge_sup[P.vadj_edge[t]]];
        const double zb = P.var.o_zb >= 0 ? xs[P.var.o_zb + b] : P.X[3 * o + 2];
        bx = fma(-qe, P.X[3 * o + 0], bx);
        by = fma(-qe, P.X[3 * o + 1], by);
        bz = fma(-qe, zb, bz);
        if (O.c_zb >= 0) row[O.c_zb + b] -= qe;
      }
      row[W - 2] = bx;
      row[W - 1] = by;
      row[O.c_z] = bz;
      if (O.c_lambdv >= 0) row[O.c_lambdv] = -P.pzv[v];
    }
    __syncthreads();

    // ---- 6. solve phase 1: z | Az | z_lambdv | x | y ------------------------------------------------
    {
      const int nc1 = O.n1 + 2;
      int G = 1;
      while (G < nc1 && G < 32) G <<= 1;
      oc_solve<1>(P, O, Lv, R, nc1, G, tid, THREADS);
    }

    // ---- 7. everything that needs x, y (transient columns) or only z --------------------------------
    auto zfix = [&](int b) -> double { return P.var.o_zb >= 0 ? xs[P.var.o_zb + b] : P.X[3 * P.fixed[b] + 2]; };
    auto zval = [&](int v) -> double {
      const int r = P.vrow[v];
      return r >= 0 ? R[r * W + O.c_z] : zfix(-1 - r);
    };
    auto xyval = [&](int v, int axis) -> double {
      const int r = P.vrow[v];
      return r >= 0 ? R[r * W + W - 2 + axis] : P.X[3 * v + axis];
    };
    for (int v = tid; v < n; v += THREADS) {
      const double z = zval(v);
      if (A.z) A.z[cand * n + v] = z;
      bad |= !isfinite(z);
      if (P.con.r_env >= 0) {
        double ub, lb;
        if (has_t) {
          double dub, dlb;
          oc_envelope(P, xyval(v, 0), xyval(v, 1), xs[P.var.o_t], &ub, &lb, &dub, &dlb);
          envd[v] = dub;
          envd[n + v] = dlb;
        } else {
          ub = P.ub[v];
          lb = P.lb[v];
        }
        const double g0 = z - lb, g1 = ub - z;
        if (gout) {
          gout[P.con.r_env + v] = g0;
          gout[P.con.r_env + n + v] = g1;
        }
        gmin = fmin(gmin, fmin(g0, g1));
      }
    }
    if (want_thrust || want_reac) {
      // R_b = Cb^T Q C X - P_b
      for (int b = tid; b < nb; b += THREADS) {
        const int vb = P.fixed[b];
        const double xb = P.X[3 * vb + 0], yb = P.X[3 * vb + 1], zb = zfix(b);
        double Rx = 0.0, Ry = 0.0, Rz = 0.0, ud0 = 0.0, vd0 = 0.0;
        for (int t = O.sup_ptr[b]; t < O.sup_ptr[b + 1]; ++t) {
          const int o = O.sup_other[t];
          const double sg = (double)O.sup_sign[t];
          const double qe = qsup[t];
          const double xo = xyval(o, 0), yo = xyval(o, 1), zo = zval(o);
          const double u = sg > 0 ? (xb - xo) : (xo - xb);
          const double vv = sg > 0 ? (yb - yo) : (yo - yb);
          const double w = sg > 0 ? (zb - zo) : (zo - zb);
          Rx = fma(sg * qe, u, Rx);
          Ry = fma(sg * qe, vv, Ry);
          Rz = fma(sg * qe, w, Rz);
          if (P.var.o_lambdh >= 0) {
            const double d0 = P.d[O.sup_edge[t]];
            ud0 = fma(sg * u, d0, ud0);
            vd0 = fma(sg * vv, d0, vd0);
          }
        }
        if (P.var.o_lambdh >= 0) {
          gxy[2 * nb * k + b] = ud0;
          gxy[2 * nb * k + nb + b] = vd0;
        }
        double px, py, pz;
        if (P.var.o_lambdh >= 0) {
          px = lamh * P.px0[vb];
          py = lamh * P.py0[vb];
        } else {
          px = P.P[3 * vb + 0];
          py = P.P[3 * vb + 1];
        }
        pz = (P.var.o_lambdv >= 0) ? (lamv * P.pzv[vb] + P.pz0[vb]) : P.P[3 * vb + 2];
        pz *= pzs;
        react[3 * b + 0] = Rx - px;
        react[3 * b + 1] = Ry - py;
        react[3 * b + 2] = Rz - pz;
      }
      // (Cb^T U B)[b, j], (Cb^T V B)[b, j]
      for (int it = tid; it < nb * k; it += THREADS) {
        const int b = it / k, j = it - b * k;
        const int vb = P.fixed[b];
        const double xb = P.X[3 * vb + 0], yb = P.X[3 * vb + 1];
        double gx = 0.0, gy = 0.0;
        for (int t = O.sup_ptr[b]; t < O.sup_ptr[b + 1]; ++t) {
          const int o = O.sup_other[t];
          const double sg = (double)O.sup_sign[t];
          const double xo = xyval(o, 0), yo = xyval(o, 1);
          const double u = sg > 0 ? (xb - xo) : (xo - xb);
          const double vv = sg > 0 ? (yb - yo) : (yo - yb);
          const double Bej = P.B[(int64_t)O.sup_edge[t] * k + j];
          gx = fma(sg * u, Bej, gx);
          gy = fma(sg * vv, Bej, gy);
        }
        gxy[it] = gx;
        gxy[nb * k + it] = gy;
      }
    }
    __syncthreads();

    // ---- 8. phase-2 right-hand sides: Ci^T W B | Ci^T W d0 ------------------------------------------
    const int nc2 = (O.c_q >= 0) ? (k + (O.c_lambdh >= 0 ? 1 : 0)) : 0;
    if (nc2 > 0) {
      int G = 1;
      while (G < nc2 && G < 32) G <<= 1;
      const int cgroups = (nc2 + G - 1) / G;
      const int r_ = tid & (G - 1), slot = tid / G, nslots = THREADS / G;
      for (int it = slot; it < ni * cgroups; it += nslots) {
        const int r = it / cgroups;
        const int cidx = (it - r * cgroups) * G + r_;
        if (cidx >= nc2) continue;
        const int v = P.row_vertex[r];
        const double zv = R[r * W + O.c_z];
        double acc = 0.0;
        for (int t = P.vadj_ptr[v]; t < P.vadj_ptr[v + 1]; ++t) {
          const double zo = zval(P.vadj_other[t]);
          const double sg = (double)P.vadj_sign[t];
          const double w = sg > 0 ? (zv - zo) : (zo - zv);
          const int e = P.vadj_edge[t];
          const double coef = cidx < k ? P.B[(int64_t)e * k + cidx] : P.d[e];
          acc = fma(sg * w, coef, acc);
        }
        R[r * W + O.c_q + cidx] = acc;
      }
      __syncthreads();
      // ---- 9. solve phase 2 ----
      oc_solve<2>(P, O, Lv, R, nc2, G, tid, THREADS);
    }

    // ---- 10. reac_bounds rows, objective, gradient ---------------------------------------------------
    if (want_reac) {
      const bool has_env = (P.constraints & TNO_CON_ENVELOPE) != 0;
      // one item per (support b, Jacobian column)
      for (int it = tid; it < nb * nvar; it += THREADS) {
        const int b = it / nvar, col = it - b * nvar;
        const int vb = P.fixed[b];
        const double Rx = react[3 * b], Ry = react[3 * b + 1], Rz = react[3 * b + 2];
        const double sx = Rx < 0 ? -1.0 : 1.0, sy = Ry < 0 ? -1.0 : 1.0, sz = Rz < 0 ? -1.0 : 1.0;
        const double rz2 = Rz * Rz;
        const double zb = zfix(b);
        double vx = 0.0, vy = 0.0;
        const int t0 = O.sup_ptr[b], t1 = O.sup_ptr[b + 1];
        if (col < k || col == P.var.o_lambdh) {
          const bool isq = col < k;
          const int pcol = isq ? (O.c_q + col) : O.c_lambdh;
          double dRx = isq ? gxy[b * k + col] : 0.0, dRy = isq ? gxy[nb * k + b * k + col] : 0.0, dRz = 0.0;
          for (int t = t0; t < t1; ++t) {
            const int o = O.sup_other[t];
            const int e = O.sup_edge[t];
            const double sg = (double)O.sup_sign[t];
            const double qe = qsup[t];
            const double zo = zval(o);
            const double w = sg > 0 ? (zb - zo) : (zo - zb);
            const double coef = isq ? P.B[(int64_t)e * k + col] : P.d[e];
            dRz = fma(sg * w, coef, dRz);
            const int ro = P.vrow[o];
            if ((has_env || !isq) && ro >= 0) {
              const double dzo = R[ro * W + pcol];
              dRz = fma(sg * qe, sg > 0 ? -dzo : dzo, dRz);
            }
          }
          if (!isq) {  // lambdh column: Cb^T U d0 - px0, Cb^T V d0 - py0 (formed in step 7 while x, y were live)
            dRx = gxy[2 * nb * k + b] - P.px0[vb];
            dRy = gxy[2 * nb * k + nb + b] - P.py0[vb];
          }
          vx = zb * sx * (-Rz * dRx + Rx * dRz) / rz2 / sz;
          vy = zb * sy * (-Rz * dRy + Ry * dRz) / rz2 / sz;
        } else if (P.var.o_zb >= 0 && col >= P.var.o_zb && col < P.var.o_zb + nb) {
          const int cb = col - P.var.o_zb;
          double acc = 0.0;
          for (int t = O.sup_ptr[cb]; t < O.sup_ptr[cb + 1]; ++t) {
            const int o = O.sup_other[t];
            const double sg = (double)O.sup_sign[t];
            const double qe = qsup[t];
            const int ro = P.vrow[o];
            const double Ao = ro >= 0 ? R[ro * W + O.c_zb + b] : ((-1 - ro) == b ? 1.0 : 0.0);
            const double Ac = (cb == b) ? 1.0 : 0.0;
            acc = fma(sg * qe, sg > 0 ? (Ac - Ao) : (Ao - Ac), acc);
          }
          vx = sz * zb * fabs(Rx) / rz2 * acc;
          vy = sz * zb * fabs(Ry) / rz2 * acc;
          if (cb == b) {
            vx += -fabs(Rx / Rz);
            vy += -fabs(Ry / Rz);
          }
        } else if (col == P.var.o_lambdv) {
          double dRz = 0.0;
          for (int t = t0; t < t1; ++t) {
            const int o = O.sup_other[t];
            const double sg = (double)O.sup_sign[t];
            const int ro = P.vrow[o];
            if (ro >= 0) {
              const double dzo = R[ro * W + O.c_lambdv];
              dRz = fma(sg * qsup[t], sg > 0 ? -dzo : dzo, dRz);
            }
          }
          dRz -= P.pzv[vb];
          vx = sz * zb * fabs(Rx) / rz2 * dRz;
          vy = sz * zb * fabs(Ry) / rz2 * dRz;
        }
        if (Jout) {
          Jout[(int64_t)(P.con.r_reac + b) * nvar + col] = vx;
          Jout[(int64_t)(P.con.r_reac + nb + b) * nvar + col] = vy;
        }
        if (col == 0) {
          const double rise = zb - (P.s ? P.s[vb] : 0.0);
          const double gx_ = fabs(P.b[2 * b + 0]) - rise * fabs(Rx / Rz);
          const double gy_ = fabs(P.b[2 * b + 1]) - rise * fabs(Ry / Rz);
          if (gout) {
            gout[P.con.r_reac + b] = gx_;
            gout[P.con.r_reac + nb + b] = gy_;
          }
          gmin = fmin(gmin, fmin(gx_, gy_));
          bad |= !isfinite(gx_) || !isfinite(gy_);
        }
      }
    }
    double fval = 0.0;
    if (want_thrust) {
      const double osign = (P.objective == TNO_OBJ_MAX) ? -1.0 : 1.0;
      if (tid < nvar && A.grad) {
        double gsum = 0.0;
        if (tid < k)
          for (int b = 0; b < nb; ++b) {
            const double Rx = react[3 * b], Ry = react[3 * b + 1];
            const double Rn = sqrt(Rx * Rx + Ry * Ry);
            gsum += (osign * Rx / Rn) * gxy[b * k + tid] + (osign * Ry / Rn) * gxy[nb * k + b * k + tid];
          }
        A.grad[cand * nvar + tid] = gsum;
      }
      if (tid == 0) {
        for (int b = 0; b < nb; ++b) fval += sqrt(react[3 * b] * react[3 * b] + react[3 * b + 1] * react[3 * b + 1]);
        fval *= osign;
      }
    } else {
      if (tid < nvar && A.grad)
        A.grad[cand * nvar + tid] = (tid == nvar - 1) ? (P.objective == TNO_OBJ_T ? 1.0 : (P.objective == TNO_OBJ_NEG_LAST ? -1.0 : 0.0)) : 0.0;
      if (tid == 0) fval = P.objective == TNO_OBJ_T ? xs[nvar - 1] : (P.objective == TNO_OBJ_NEG_LAST ? -xs[nvar - 1] : 0.0);
    }
    if (tid == 0) {
      bad |= !isfinite(fval);
      if (A.f) A.f[cand] = fval;
    }

    // ---- 11. dense Jacobian: funicular block from B, envelope block from the panel -------------------
    if (Jout) {
      if (P.con.r_fun >= 0) {
        double* o1 = Jout + (int64_t)P.con.r_fun * nvar;
        double* o2 = o1 + (int64_t)m * nvar;
        for (int e = tid; e < m * nvar; e += THREADS) {
          const int row = __umulhi((unsigned)e, O.nvar_magic);
          const int col = e - row * nvar;
          double v = 0.0;
          if (col < k)
            v = __ldg(P.B + (int64_t)row * k + col);
          else if (col == P.var.o_lambdh)
            v = P.d[row];
          o1[e] = v;
          o2[e] = -v;
        }
      }
      if (P.con.r_env >= 0) {
        double* o1 = Jout + (int64_t)P.con.r_env * nvar;
        double* o2 = o1 + (int64_t)n * nvar;
        for (int e = tid; e < n * nvar; e += THREADS) {
          const int v = __umulhi((unsigned)e, O.nvar_magic);
          const int col = e - v * nvar;
          const int r = P.vrow[v];
          const int src = O.colsrc[col];
          double a, bneg;
          if (src >= 0) {
            a = r >= 0 ? R[r * W + src] : 0.0;
            if (r < 0 && P.var.o_zb >= 0 && col == P.var.o_zb + (-1 - r)) a = 1.0;
            bneg = -a;
          } else {  // thickness column: [-dlb/dt ; +dub/dt]
            a = -envd[n + v];
            bneg = envd[v];
          }
          o1[e] = a;
          o2[e] = bneg;
        }
      }
    }

    // ---- 12. feasibility margin and status -----------------------------------------------------------
    for (int off = 16; off > 0; off >>= 1) gmin = fmin(gmin, __shfl_xor_sync(0xffffffffu, gmin, off));
    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) {
      red[warp] = gmin;
      if (anybad) atomicOr(&s_flags, TNO_STATUS_NONFINITE);
    }
    __syncthreads();
    if (tid == 0) {
      double g = red[0];
      for (int w = 1; w < THREADS / 32; ++w) g = fmin(g, red[w]);
      if (A.gmin) A.gmin[cand] = g;
      if (A.status) A.status[cand] = s_flags;
    }
    __syncthreads();  // panel, flags and small arrays are reused by the next candidate
  }
}

// ------------------------------------------------------------------------------------------------------
// host: schedule construction
int build_onchip(Plan* PL, const tno_problem_desc* D) {
  const Symbolic& S = PL->sym;
  DevPlan& P = PL->dp;
  DevOnchip& O = PL->oc;
  PL->onchip_ok = false;
  const int ni = P.ni, n = P.n, m = P.m, nb = P.nb, k = P.k, nvar = P.var.nvar;
  const int nnz = P.nnz_l;
  if (nnz >= 65536 || ni >= 65536) return TNO_OK;

  // panel columns: [z | Az | z_lambdv | Dq | z_lambdh], x and y transient in the last two columns
  int pos = 0;
  O.c_z = pos++;
  O.c_zb = -1;
  O.c_lambdv = -1;
  O.c_q = -1;
  O.c_lambdh = -1;
  if (P.var.o_zb >= 0) { O.c_zb = pos; pos += nb; }
  if (P.var.o_lambdv >= 0) { O.c_lambdv = pos; pos += 1; }
  O.n1 = pos;
  if (P.rhs.c_q >= 0) {
    O.c_q = pos; pos += k;
    if (P.var.o_lambdh >= 0) { O.c_lambdh = pos; pos += 1; }
  }
  O.W = std::max(pos, O.n1 + 2);
  O.nlevels = S.nlevels;
  O.nvar_magic = (uint32_t)((((uint64_t)1 << 32) + nvar - 1) / nvar);

  // support incidence
  std::vector<int> sup_ptr(nb + 1, 0), sup_edge, sup_other, edge_sup(m, -1);
  std::vector<float> sup_sign;
  for (int b = 0; b < nb; ++b) {
    const int vb = PL->h_fixed[b];
    for (int t = PL->h_vadj_ptr[vb]; t < PL->h_vadj_ptr[vb + 1]; ++t) {
      edge_sup[PL->h_vadj_edge[t]] = (int)sup_edge.size();  // an edge between two supports keeps the last index; q is the same
      sup_edge.push_back(PL->h_vadj_edge[t]);
      sup_other.push_back(PL->h_vadj_other[t]);
      sup_sign.push_back(PL->h_vadj_sign[t]);
    }
    sup_ptr[b + 1] = (int)sup_edge.size();
  }
  O.nsup = (int)sup_edge.size();

  // shared memory layout (doubles)
  int off = 0;
  O.o_x = off; off += (nvar + 3) / 4 * 4;
  O.o_l = off; off += (nnz + 3) / 4 * 4;
  O.o_r = off; off += std::max(ni * O.W, m);
  off = (off + 3) / 4 * 4;
  O.o_env = -1;
  if (P.var.o_t >= 0) { O.o_env = off; off += 2 * n; }
  O.o_small = off;
  int so = 0;
  O.o_qsup = so; so += O.nsup;
  O.o_react = so; so += 3 * nb;
  O.o_gxy = so; so += 2 * nb * k + 2 * nb;  // + (Cb^T U d0, Cb^T V d0) for the lambdh column
  O.o_red = so; so += 40;
  off += so;
  O.smem_doubles = off;
  const size_t smem_bytes = (size_t)off * 8;
  int max_optin = 0;
  TNO_CUDA_OK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, PL->device));
  if (smem_bytes > (size_t)max_optin) return TNO_OK;
  int smem_sm = 0;
  TNO_CUDA_OK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, PL->device));
  PL->onchip_threads = 256;
  PL->onchip_ctas_per_sm = std::max(1, std::min(8, (int)((size_t)smem_sm / (smem_bytes + 1024))));

  // ---- factorisation items, level ordered ----
  std::vector<int> f_lev_ptr(S.nlevels + 1, 0), f_lev_lpi(S.nlevels, 1), f_item_pptr(1, 0);
  std::vector<unsigned long long> f_pairs;
  std::vector<unsigned short> f_item_pos, f_item_col, ent_dp(nnz), s_row(ni), rowidx16(nnz);
  std::vector<uint32_t> row_bd(S.rowpos.size()), row_pc(S.rowpos.size());
  for (size_t t = 0; t < S.rowpos.size(); ++t) {
    const int c = S.rowcol[t];
    row_bd[t] = (uint32_t)S.rowpos[t] | ((uint32_t)S.colptr[c] << 16);
    row_pc[t] = (uint32_t)S.rowpos[t] | ((uint32_t)c << 16);
  }
  for (int j = 0; j < ni; ++j)
    for (int p = S.colptr[j]; p < S.colptr[j + 1]; ++p) {
      ent_dp[p] = (unsigned short)S.colptr[j];
      rowidx16[p] = (unsigned short)S.rowidx[p];
    }
  for (int lev = 0; lev < S.nlevels; ++lev) {
    int64_t pairs = 0;
    for (int ci = S.level_ptr[lev]; ci < S.level_ptr[lev + 1]; ++ci) {
      const int j = S.level_cols[ci];
      for (int p = S.colptr[j]; p < S.colptr[j + 1]; ++p) {
        f_item_pos.push_back((unsigned short)p);
        f_item_col.push_back((unsigned short)j);
        for (int64_t t = S.dot_ptr[p]; t < S.dot_ptr[p + 1]; ++t)
          f_pairs.push_back((unsigned long long)S.dot_a[t] | ((unsigned long long)S.dot_b[t] << 16) |
                            ((unsigned long long)S.colptr[S.dot_c[t]] << 32));
        f_item_pptr.push_back((int)f_pairs.size());
        pairs += S.dot_ptr[p + 1] - S.dot_ptr[p];
      }
    }
    f_lev_ptr[lev + 1] = (int)f_item_pos.size();
    const int items = f_lev_ptr[lev + 1] - f_lev_ptr[lev];
    int lpi = 1;
    if (items > 0 && pairs / items >= 4)
      while (lpi < 32 && items * lpi * 2 <= PL->onchip_threads) lpi <<= 1;
    f_lev_lpi[lev] = lpi;
  }
  for (int i = 0; i < ni; ++i) s_row[i] = (unsigned short)S.level_cols[i];

  std::vector<double> Bt((size_t)k * m);
  for (int e = 0; e < m; ++e)
    for (int j = 0; j < k; ++j) Bt[(size_t)j * m + e] = D->B[(size_t)e * k + j];
  std::vector<int> colsrc(nvar, -1);
  for (int j = 0; j < k; ++j) colsrc[j] = O.c_q >= 0 ? O.c_q + j : -2;
  if (P.var.o_zb >= 0)
    for (int b = 0; b < nb; ++b) colsrc[P.var.o_zb + b] = O.c_zb + b;
  if (P.var.o_lambdh >= 0) colsrc[P.var.o_lambdh] = O.c_lambdh;
  if (P.var.o_lambdv >= 0) colsrc[P.var.o_lambdv] = O.c_lambdv;
  if (P.var.o_t >= 0) colsrc[P.var.o_t] = -1;

  int rc;
#define UPO(field, hostvec) \
  if ((rc = upload_vec(PL, hostvec, &O.field))) return rc;
  UPO(Bt, Bt);
  UPO(f_lev_ptr, f_lev_ptr);
  UPO(f_lev_lpi, f_lev_lpi);
  UPO(f_item_pos, f_item_pos);
  UPO(f_item_col, f_item_col);
  UPO(f_item_pptr, f_item_pptr);
  UPO(f_pairs, f_pairs);
  UPO(rowptr, S.rowptr);
  UPO(row_bd, row_bd);
  UPO(row_pc, row_pc);
  UPO(ent_dp, ent_dp);
  UPO(s_lev_ptr, S.level_ptr);
  UPO(s_row, s_row);
  UPO(rowidx16, rowidx16);
  UPO(edge_sup, edge_sup);
  UPO(sup_ptr, sup_ptr);
  UPO(sup_edge, sup_edge);
  UPO(sup_other, sup_other);
  UPO(sup_sign, sup_sign);
  UPO(colsrc, colsrc);
#undef UPO
  PL->onchip_ok = true;
  return TNO_OK;
}

int run_onchip(Plan* PL, const double* x_dev, int64_t N, int64_t ldx, const tno_batch_inputs* opt,
               const tno_outputs* out, cudaStream_t stream) {
  OcArgs A;
  A.x = x_dev;
  A.ldx = ldx;
  A.pzs = (opt && opt->pz_scale) ? opt->pz_scale : nullptr;
  A.f = out->f;
  A.grad = out->grad;
  A.g = out->g;
  A.J = out->J;
  A.z = out->z;
  A.q = out->q;
  A.gmin = out->gmin;
  A.status = out->status;
  A.N = N;
  const size_t smem = (size_t)PL->oc.smem_doubles * 8;
  const int grid = (int)std::min<int64_t>(N, (int64_t)PL->sm_count * PL->onchip_ctas_per_sm);
  auto kern = PL->onchip_ctas_per_sm >= 2 ? k_onchip<256, 2> : k_onchip<256, 1>;
  TNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PL->prof_begin(KK_ONCHIP, stream);
  kern<<<grid, 256, smem, stream>>>(PL->dp, PL->oc, A);
  PL->prof_end(stream);
  PL->launches += 1;
  TNO_CUDA_OK(cudaGetLastError());
  return TNO_OK;
}

}  // namespace tno
