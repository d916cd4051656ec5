// Kernel family 2: "on-chip" -- one CTA per candidate, factor and right-hand-side panel resident in
// shared memory, ONE fused launch for the whole evaluation (persistent grid-stride over candidates).
//
// HBM traffic of a candidate is its algorithmic traffic only: x in; z, f, grad, g and the dense J out.
// The numeric factor (nnz_l doubles) and the panel R (ni x W doubles) never leave the SM.
//
// Parallelism inside a candidate comes from the elimination tree.  The plan orders the matrix by level
// (height above the leaves), so the columns -- and rows -- of one level are contiguous and independent:
//   * factorisation: pull ("dot product") form of L D L^T, one barrier per level; the (a, b, d) index
//     triples of a level are streamed from L2 into the idle panel with cp.async two levels ahead, so no
//     global-memory latency sits on the level-to-level critical path;
//   * forward solve: pull form over the rows of a level, lanes = right-hand-side columns (conflict-free,
//     contiguous panel rows), long rows split over sub-slots and reduced with shuffles;
//   * backward solve: PUSH form over the same row-major storage -- a vertex has at most one ancestor per
//     level, so the updates of one level never collide and a dense top row is spread over the whole CTA;
//   * all solve indices (row pointers, 16-bit columns, level table) live in shared memory for the CTA's
//     lifetime, shared by every candidate it processes.
//
// Math restated from the reference, same citations as tno_interleaved.cu.
#include <algorithm>
#include <cstdio>
#include <cstring>

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

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

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

// ---- level kernels of the triangular solves -----------------------------------------------------------
// A "team" of NT threads (the whole CTA, or one warp for the narrow levels near the root) handles the rows
// of one level.  G = 2^gshift lanes cover the columns of a row, TWO adjacent columns per lane (one 16-byte
// shared-memory access), S = 2^sshift sub-slots split the entries of a row.

// forward, pull form: y_i -= sum_j L_ij y_j ; sub-slot partial sums are combined with shuffles (S * G <= 32)
template <int NT>
__device__ __forceinline__ void oc_pull_level(const OcLevel& L, int sshift, int gshift, int tl, int c0, int ncols, int W,
                                              const unsigned short* __restrict__ rowptr16,
                                              const unsigned short* __restrict__ rowcol16, const double* __restrict__ Lv,
                                              double* __restrict__ R) {
  const int lshift = sshift + gshift;
  const int rows = L.row1 - L.row0;
  const int r = tl & ((1 << gshift) - 1);
  const int sub = (tl >> gshift) & ((1 << sshift) - 1);
  const int S = 1 << sshift;
  const bool colok = 2 * r < ncols;
  const int col = c0 + 2 * r;
  for (int base = 0; base < rows; base += NT >> lshift) {
    const int it = base + (tl >> lshift);
    const bool valid = it < rows && colok;
    const int row = L.row0 + it;
    double ax = 0.0, ay = 0.0;
    if (valid) {
      const int t1 = rowptr16[row + 1];
#pragma unroll 4
      for (int t = rowptr16[row] + sub; t < t1; t += S) {
        const double l = Lv[t];
        const double2 y = *reinterpret_cast<const double2*>(R + (int)rowcol16[t] * W + col);
        ax = fma(l, y.x, ax);
        ay = fma(l, y.y, ay);
      }
    }
    for (int off = (1 << lshift) >> 1; off >= (1 << gshift); off >>= 1) {
      ax += __shfl_xor_sync(0xffffffffu, ax, off);
      ay += __shfl_xor_sync(0xffffffffu, ay, off);
    }
    if (valid && sub == 0) {
      double2* d = reinterpret_cast<double2*>(R + row * W + col);
      double2 v = *d;
      v.x -= ax;
      v.y -= ay;
      *d = v;
    }
  }
}

// backward, push form: x_i is final -> x_j -= L_ij x_i for every entry j of row i.  A vertex has at most one
// ancestor per level, so the targets of one level are disjoint: no atomics, any number of sub-slots.
template <int NT>
__device__ __forceinline__ void oc_push_level(const OcLevel& L, int sshift, int gshift, int tl, int c0, int ncols, int W,
                                              const unsigned short* __restrict__ rowptr16,
                                              const unsigned short* __restrict__ rowcol16, const double* __restrict__ Lv,
                                              double* __restrict__ R) {
  const int lshift = sshift + gshift;
  const int rows = L.row1 - L.row0;
  const int r = tl & ((1 << gshift) - 1);
  const int sub = (tl >> gshift) & ((1 << sshift) - 1);
  const int S = 1 << sshift;
  const bool colok = 2 * r < ncols;
  const int col = c0 + 2 * r;
  for (int base = 0; base < rows; base += NT >> lshift) {
    const int it = base + (tl >> lshift);
    if (it < rows && colok) {
      const int row = L.row0 + it;
      const double2 xi = *reinterpret_cast<const double2*>(R + row * W + col);
      const int t1 = rowptr16[row + 1];
#pragma unroll 4
      for (int t = rowptr16[row] + sub; t < t1; t += S) {
        const double l = Lv[t];
        double2* d = reinterpret_cast<double2*>(R + (int)rowcol16[t] * W + col);
        double2 v = *d;
        v.x = fma(-l, xi.x, v.x);
        v.y = fma(-l, xi.y, v.y);
        *d = v;
      }
    }
  }
}

// L D L^T solve of the panel columns [c0, c0 + ncols) (c0 even) for all rows.  Levels [1, lsplit) are wide and use the
// whole CTA with one block barrier each; levels [lsplit, nlevels) fit one warp, which walks them with warp barriers
// only while the other warps wait at a single block barrier.
template <int THREADS>
__device__ __forceinline__ void oc_solve(const DevOnchip& O, int phase, const OcLevel* __restrict__ levels,
                                         const unsigned short* __restrict__ rowptr16, const unsigned short* __restrict__ rowcol16,
                                         const double* __restrict__ Lv, double* __restrict__ R, int ni, int c0, int ncols, int tid) {
  const int W = O.W;
  const int gshift = O.gshift[phase];
  const int lsplit = O.lsplit[phase];
  const int H = O.nlevels;
  for (int lev = 1; lev < lsplit; ++lev) {
    const OcLevel L = levels[lev];
    oc_pull_level<THREADS>(L, L.pull_s[phase][0], gshift, tid, c0, ncols, W, rowptr16, rowcol16, Lv, R);
    __syncthreads();
  }
  if (tid < 32) {
    for (int lev = lsplit; lev < H; ++lev) {
      const OcLevel L = levels[lev];
      oc_pull_level<32>(L, L.pull_s[phase][1], gshift, tid, c0, ncols, W, rowptr16, rowcol16, Lv, R);
      __syncwarp();
    }
  }
  __syncthreads();
  // diagonal: y_i / d_i
  {
    const int r = tid & ((1 << gshift) - 1);
    if (2 * r < ncols) {
      for (int row = tid >> gshift; row < ni; row += THREADS >> gshift) {
        double2* d = reinterpret_cast<double2*>(R + row * W + c0 + 2 * r);
        const double s = Lv[O.nlow + row];
        double2 v = *d;
        v.x *= s;
        v.y *= s;
        *d = v;
      }
    }
  }
  __syncthreads();
  if (tid < 32) {
    for (int lev = H - 1; lev >= lsplit && lev >= 1; --lev) {
      const OcLevel L = levels[lev];
      oc_push_level<32>(L, L.push_s[phase][1], gshift, tid, c0, ncols, W, rowptr16, rowcol16, Lv, R);
      __syncwarp();
    }
  }
  __syncthreads();
  for (int lev = lsplit - 1; lev >= 1; --lev) {
    const OcLevel L = levels[lev];
    oc_push_level<THREADS>(L, L.push_s[phase][0], gshift, tid, c0, ncols, W, rowptr16, rowcol16, Lv, R);
    __syncthreads();
  }
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) k_onchip(const __grid_constant__ DevPlan P, const __grid_constant__ DevOnchip O,
                                                            const __grid_constant__ OcArgs A) {
  extern __shared__ __align__(16) unsigned char smraw[];
  double* sm = reinterpret_cast<double*>(smraw);
  double* xs = sm + O.o_x;
  double* Lv = sm + O.o_l;
  double* R = sm + O.o_r;
  double* qs = R;  // q lives in the panel region until the matrix is assembled
  double* qsup = sm + O.o_qsup;
  double* react = sm + O.o_react;  // nb x 3
  double* gxy = sm + O.o_gxy;      // Cb^T U B | Cb^T V B (nb x k each) | Cb^T U d0 | Cb^T V d0 (nb each)
  double* azsup = sm + O.o_azsup;  // nsup x nb : dz/dzb at the neighbours of the supports
  double* red = sm + O.o_red;
  const unsigned short* rowptr16 = reinterpret_cast<const unsigned short*>(smraw + O.o_idx_bytes + O.idx_rowptr);
  const unsigned short* rowcol16 = reinterpret_cast<const unsigned short*>(smraw + O.o_idx_bytes + O.idx_rowcol);
  const OcLevel* levels = reinterpret_cast<const OcLevel*>(smraw + O.o_idx_bytes + O.idx_levels);
  __shared__ int s_flags;

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int n = P.n, m = P.m, ni = P.ni, nb = P.nb, k = P.k, nvar = P.var.nvar, ng = P.con.ng, W = O.W;
  const int nlow = O.nlow;
  const bool has_t = P.var.o_t >= 0;
  const bool want_thrust = (P.objective == TNO_OBJ_MIN || P.objective == TNO_OBJ_MAX);
  const bool want_reac = (P.constraints & TNO_CON_REAC_BOUNDS) != 0;

  // index region: loaded once per CTA, reused by every candidate
  for (int i = tid; i < O.idx_bytes / 16; i += THREADS)
    reinterpret_cast<uint4*>(smraw + O.o_idx_bytes)[i] = reinterpret_cast<const uint4*>(O.idx_init)[i];
  __syncthreads();

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
      for (int t = O.asm_ptr[p]; t < O.asm_ptr[p + 1]; ++t) acc += (double)O.asm_coef[t] * qs[O.asm_edge[t]];
      Lv[p] = acc;
    }
    for (int t = tid; t < O.nsup; t += THREADS) qsup[t] = qs[O.sup_edge[t]];
    __syncthreads();  // q consumed: the panel region is free for the factor stream

    // ---- 3. numeric L D L^T, level by level (off-diagonals stay unscaled until step 4) --------------
    {
      unsigned char* stage_base = reinterpret_cast<unsigned char*>(R);
      const int seg_bytes = O.fseg_units * 16;
      auto issue = [&](int lev) {
        if (O.stage && lev < O.nlevels) {
          const OcLevel L = levels[lev];
          unsigned char* dst = stage_base + (lev % 3) * seg_bytes;
          const uint4* src = O.f_stream + L.seg_off;
          for (unsigned u = tid; u < L.seg_units; u += THREADS) cp_async16(dst + 16 * u, src + u);
        }
        cp_async_commit();
      };
      issue(0);
      issue(1);
      for (int lev = 0; lev < O.nlevels; ++lev) {
        cp_async_wait<1>();
        __syncthreads();
        issue(lev + 2);
        const OcLevel L = levels[lev];
        const uint2* rec = O.stage ? reinterpret_cast<const uint2*>(stage_base + (lev % 3) * seg_bytes)
                                   : reinterpret_cast<const uint2*>(O.f_stream + L.seg_off);
        const unsigned long long* pairs = reinterpret_cast<const unsigned long long*>(rec + L.nitems + 1);
        const int lpi = 1 << L.lpi_shift;
        const int sub = tid & (lpi - 1);
        const int per_pass = THREADS >> L.lpi_shift;
        for (int base = 0; base < L.nitems; base += per_pass) {
          const int it = base + (tid >> L.lpi_shift);
          const bool valid = it < L.nitems;
          double acc = 0.0;
          unsigned pf = 0;
          if (valid) {
            const uint2 r0 = rec[it];
            pf = r0.x;
            const int t1 = rec[it + 1].y;
            for (int t = r0.y + sub; t < t1; t += lpi) {
              const unsigned long long pr = pairs[t];
              const int a = (int)(pr & 0xffffu), b = (int)((pr >> 16) & 0xffffu), dp = (int)((pr >> 32) & 0xffffu);
              acc = fma(Lv[a] * Lv[dp], Lv[b], acc);  // U_ic * (1/d_c) * U_jc
            }
          }
          for (int off = lpi >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
          if (valid && sub == 0) {
            const int pos = (int)(pf & 0xffffu);
            const double v = Lv[pos] - acc;
            if (pf >> 31) {
              if (!(v > 0.0)) atomicOr(&s_flags, TNO_STATUS_NOT_PD);
              Lv[pos] = 1.0 / v;
            } else {
              Lv[pos] = v;
            }
          }
        }
      }
      cp_async_wait<0>();
      __syncthreads();
    }
    // ---- 4. L_ij = U_ij / d_j ----------------------------------------------------------------------
    for (int p = tid; p < nlow; p += THREADS) Lv[p] *= Lv[nlow + (int)rowcol16[p]];

    // ---- 5. phase-1 right-hand sides: z | z_lambdv | Az | x | y -------------------------------------
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
      for (int cc = 0; cc < ((O.ncols1 + 1) & ~1); ++cc) row[cc] = 0.0;  // incl. padding columns touched by 2-column lanes
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
        if (O.c_az >= 0) row[O.c_az + b] -= qe;
      }
      row[O.c_x] = bx;
      row[O.c_y] = by;
      row[O.c_z] = bz;
      if (O.c_lambdv >= 0) row[O.c_lambdv] = -P.pzv[v];
    }
    __syncthreads();

    // ---- 6. solve phase 1 ---------------------------------------------------------------------------
    oc_solve<THREADS>(O, 0, levels, rowptr16, rowcol16, Lv, R, ni, 0, O.ncols1, tid);

    // ---- 7. everything that needs the transient columns (x, y, Az) or only z ------------------------
    auto zfix = [&](int b) -> double { return P.var.o_zb >= 0 ? xs[P.var.o_zb + b] : P.X[3 * P.fixed[b] + 2]; };
    auto zval = [&](int v) -> double {
      const int r = P.vrow[v];
      return r >= 0 ? R[r * W + O.c_z] : zfix(-1 - r);
    };
    auto xyval = [&](int v, int axis) -> double {
      const int r = P.vrow[v];
      return r >= 0 ? R[r * W + O.c_x + axis] : P.X[3 * v + axis];
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
          if (Jout) {  // thickness column: [-dlb/dt ; +dub/dt]
            Jout[(int64_t)(P.con.r_env + v) * nvar + P.var.o_t] = -dlb;
            Jout[(int64_t)(P.con.r_env + n + v) * nvar + P.var.o_t] = dub;
          }
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
    if (Jout && P.con.r_env >= 0 && P.var.o_zb >= 0) {
      // dz/dzb block straight to the Jacobian (its panel columns are about to be reused)
      for (int idx = tid; idx < n * nb; idx += THREADS) {
        const int v = idx / nb, b = idx - v * nb;
        const int r = P.vrow[v];
        const double a = r >= 0 ? R[r * W + O.c_az + b] : ((-1 - r) == b ? 1.0 : 0.0);
        Jout[(int64_t)(P.con.r_env + v) * nvar + P.var.o_zb + b] = a;
        Jout[(int64_t)(P.con.r_env + n + v) * nvar + P.var.o_zb + b] = -a;
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
      if (want_reac && O.c_az >= 0) {
        for (int it = tid; it < O.nsup * nb; it += THREADS) {
          const int t = it / nb, b = it - t * nb;
          const int ro = P.vrow[O.sup_other[t]];
          azsup[it] = ro >= 0 ? R[ro * W + O.c_az + b] : ((-1 - ro) == b ? 1.0 : 0.0);
        }
      }
    }
    __syncthreads();

    // ---- 8. phase-2 right-hand sides: Ci^T W B | Ci^T W d0 ------------------------------------------
    if (O.ncols2 > 0) {
      const int nc2 = O.ncols2;
      const int g2 = O.rhs2_shift;  // 2^g2 lanes walk the columns of one row
      const int r_ = tid & ((1 << g2) - 1);
      for (int r = tid >> g2; r < ni; r += THREADS >> g2)
      for (int cidx = r_; cidx < nc2; cidx += 1 << g2) {
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
      if (nc2 & 1)  // the spare column next to an odd column count takes part in the two-column accesses: keep it finite
        for (int r = tid; r < ni; r += THREADS) R[r * W + O.c_q + nc2] = 0.0;
      __syncthreads();
      // ---- 9. solve phase 2 ----
      oc_solve<THREADS>(O, 1, levels, rowptr16, rowcol16, Lv, R, ni, O.c_q, nc2, tid);
    }

    // ---- 10. reac_bounds rows, objective, gradient ---------------------------------------------------
    if (want_reac) {
      const bool has_env = (P.constraints & TNO_CON_ENVELOPE) != 0;
      for (int it = tid; it < nb * nvar; it += THREADS) {  // one item per (support b, Jacobian column)
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
          // dRz/dzb = Cb^T Q C A ; the reference adds COLUMN b of it to row b (jacobian.py:222)
          const int cb = col - P.var.o_zb;
          double acc = 0.0;
          for (int t = O.sup_ptr[cb]; t < O.sup_ptr[cb + 1]; ++t) {
            const double sg = (double)O.sup_sign[t];
            const double Ao = azsup[t * nb + b];
            const double Ac = (cb == b) ? 1.0 : 0.0;
            acc = fma(sg * qsup[t], sg > 0 ? (Ac - Ao) : (Ao - Ac), acc);
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
            const double sg = (double)O.sup_sign[t];
            const int ro = P.vrow[O.sup_other[t]];
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
          const int src = O.colsrc[col];
          if (src == -1) continue;  // zb and thickness columns were written in step 7
          const int r = P.vrow[v];
          const double a = (src >= 0 && r >= 0) ? R[r * W + src] : 0.0;
          o1[e] = a;
          o2[e] = -a;
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
  const int ni = P.ni, m = P.m, nb = P.nb, k = P.k, nvar = P.var.nvar;
  const int nnz = P.nnz_l;
  const int nlow = nnz - ni;
  if (nnz >= 65536 || ni >= 65535 || S.nlevels < 1) return TNO_OK;
  const int THREADS = 256;

  // panel columns
  int pos = 0;
  O.c_z = pos++;
  O.c_lambdv = -1;
  O.c_q = -1;
  O.c_lambdh = -1;
  if (P.var.o_lambdv >= 0) O.c_lambdv = pos++;
  if (pos & 1) ++pos;  // phase-2 columns start on a 16-byte boundary (two columns per lane)
  const int n_persist = pos;
  O.ncols2 = 0;
  if (P.rhs.c_q >= 0) {
    O.c_q = pos;
    O.ncols2 = k + (P.var.o_lambdh >= 0 ? 1 : 0);
    if (P.var.o_lambdh >= 0) O.c_lambdh = O.c_q + k;
  }
  O.c_az = (P.var.o_zb >= 0) ? n_persist : -1;
  O.c_x = n_persist + (P.var.o_zb >= 0 ? nb : 0);
  O.c_y = O.c_x + 1;
  O.ncols1 = O.c_y + 1;
  O.W = std::max(O.ncols1, n_persist + O.ncols2);
  O.W = (O.W + 2) / 2 * 2;  // even, and one spare column for the odd lane of a two-column access
  O.nlevels = S.nlevels;
  for (int ph = 0; ph < 2; ++ph) {
    const int ncols = ph == 0 ? O.ncols1 : O.ncols2;
    const int lanes = std::max(1, (ncols + 1) / 2);
    int g = 0;
    while ((1 << g) < lanes) ++g;
    if (g > 5) return TNO_OK;  // more than 64 columns per phase: not handled on chip
    O.gshift[ph] = g;
    if (ph == 1) {
      int g2 = 0;
      while ((1 << g2) < std::min(std::max(ncols, 1), 32)) ++g2;
      O.rhs2_shift = g2;
    }
    int ls = S.nlevels;
    while (ls > 1 && ((S.level_ptr[ls] - S.level_ptr[ls - 1]) << g) <= 32) --ls;
    O.lsplit[ph] = ls;
  }
  O.nlow = nlow;
  O.nvar_magic = (uint32_t)((((uint64_t)1 << 32) + nvar - 1) / nvar);

  // CSC position -> on-chip position (CSR order of the strictly lower part, diagonals last)
  std::vector<int> newpos(nnz, -1);
  for (int j = 0; j < ni; ++j) newpos[S.colptr[j]] = nlow + j;
  for (int t = 0; t < nlow; ++t) newpos[S.rowpos[t]] = t;

  // support incidence
  std::vector<int> sup_ptr(nb + 1, 0), sup_edge, sup_other, edge_sup(m, -1);
  std::vector<float> sup_sign;
  for (int b = 0; b < nb; ++b) {
    const int vb = PL->h_fixed[b];
    for (int t = PL->h_vadj_ptr[vb]; t < PL->h_vadj_ptr[vb + 1]; ++t) {
      edge_sup[PL->h_vadj_edge[t]] = (int)sup_edge.size();
      sup_edge.push_back(PL->h_vadj_edge[t]);
      sup_other.push_back(PL->h_vadj_other[t]);
      sup_sign.push_back(PL->h_vadj_sign[t]);
    }
    sup_ptr[b + 1] = (int)sup_edge.size();
  }
  O.nsup = (int)sup_edge.size();

  // ---- factor stream, one segment per level ----
  std::vector<OcLevel> levels(S.nlevels);
  std::vector<uint32_t> stream;  // 32-bit words; segments padded to 16 bytes
  uint32_t max_units = 0;
  for (int lev = 0; lev < S.nlevels; ++lev) {
    OcLevel& L = levels[lev];
    std::memset(&L, 0, sizeof(L));
    L.row0 = (unsigned short)S.level_ptr[lev];
    L.row1 = (unsigned short)S.level_ptr[lev + 1];
    std::vector<uint32_t> recs;
    std::vector<unsigned long long> prs;
    int nitems = 0;
    for (int j = S.level_ptr[lev]; j < S.level_ptr[lev + 1]; ++j) {
      if (S.level_cols[j] != j) {
        set_error("internal: level ordering is not contiguous");
        return TNO_EINVAL;
      }
      for (int p = S.colptr[j]; p < S.colptr[j + 1]; ++p) {
        recs.push_back((uint32_t)newpos[p] | (p == S.colptr[j] ? 0x80000000u : 0u));
        recs.push_back((uint32_t)prs.size());
        for (int64_t t = S.dot_ptr[p]; t < S.dot_ptr[p + 1]; ++t)
          prs.push_back((unsigned long long)newpos[S.dot_a[t]] | ((unsigned long long)newpos[S.dot_b[t]] << 16) |
                        ((unsigned long long)(nlow + S.dot_c[t]) << 32));
        ++nitems;
      }
    }
    recs.push_back(0);
    recs.push_back((uint32_t)prs.size());
    if (nitems >= 65536) return TNO_OK;
    L.nitems = (unsigned short)nitems;
    int lpi_shift = 0;
    if (nitems > 0 && (int64_t)prs.size() / nitems >= 4)
      while (lpi_shift < 5 && ((nitems << (lpi_shift + 1)) * 2) <= THREADS * 2 && (nitems << (lpi_shift + 1)) <= THREADS) ++lpi_shift;
    L.lpi_shift = (unsigned char)lpi_shift;
    {
      const int rows = S.level_ptr[lev + 1] - S.level_ptr[lev];
      for (int ph = 0; ph < 2; ++ph) {
        const int g = O.gshift[ph];
        for (int team = 0; team < 2; ++team) {
          const int NT = team == 0 ? THREADS : 32;
          int sp = 0;  // pull: S * G <= 32 (shuffle reduction) and rows * S * G <= NT
          while (sp + 1 + g <= 5 && (rows << (sp + 1 + g)) <= NT) ++sp;
          int sq = 0;  // push: rows * S * G <= NT
          while (sq < 5 && (rows << (sq + 1 + g)) <= NT) ++sq;
          L.pull_s[ph][team] = (unsigned char)sp;
          L.push_s[ph][team] = (unsigned char)sq;
        }
      }
    }
    L.seg_off = (uint32_t)(stream.size() / 4);
    stream.insert(stream.end(), recs.begin(), recs.end());
    for (unsigned long long pr : prs) {
      stream.push_back((uint32_t)(pr & 0xffffffffu));
      stream.push_back((uint32_t)(pr >> 32));
    }
    while (stream.size() % 4) stream.push_back(0);
    L.seg_units = (uint32_t)(stream.size() / 4) - L.seg_off;
    max_units = std::max(max_units, L.seg_units);
  }
  O.fseg_units = (int)max_units;

  // ---- assembly gather lists in on-chip position order ----
  std::vector<int> asm_ptr(nnz + 1, 0), asm_edge;
  std::vector<float> asm_coef;
  {
    std::vector<std::vector<std::pair<int, float>>> contrib(nnz);
    for (int e = 0; e < m; ++e) {
      const int u = D->edges[2 * e], v = D->edges[2 * e + 1];
      const int ru = PL->h_vrow[u], rv = PL->h_vrow[v];
      if (ru >= 0) contrib[nlow + ru].push_back({e, -1.f});
      if (rv >= 0) contrib[nlow + rv].push_back({e, -1.f});
      if (ru >= 0 && rv >= 0) contrib[newpos[S.pos(std::max(ru, rv), std::min(ru, rv))]].push_back({e, +1.f});
    }
    for (int p = 0; p < nnz; ++p) {
      for (auto& c : contrib[p]) {
        asm_edge.push_back(c.first);
        asm_coef.push_back(c.second);
      }
      asm_ptr[p + 1] = (int)asm_edge.size();
    }
  }

  // ---- shared-memory index region image: rowptr16 | rowcol16 | levels ----
  std::vector<unsigned char> idx;
  auto align16 = [&]() {
    while (idx.size() % 16) idx.push_back(0);
  };
  O.idx_rowptr = 0;
  for (int i = 0; i <= ni; ++i) {
    const unsigned short v = (unsigned short)S.rowptr[i];
    idx.push_back((unsigned char)(v & 0xff));
    idx.push_back((unsigned char)(v >> 8));
  }
  align16();
  O.idx_rowcol = (int)idx.size();
  for (int t = 0; t < nlow; ++t) {
    const unsigned short v = (unsigned short)S.rowcol[t];
    idx.push_back((unsigned char)(v & 0xff));
    idx.push_back((unsigned char)(v >> 8));
  }
  align16();
  O.idx_levels = (int)idx.size();
  {
    const unsigned char* lp = reinterpret_cast<const unsigned char*>(levels.data());
    idx.insert(idx.end(), lp, lp + levels.size() * sizeof(OcLevel));
  }
  align16();
  O.idx_bytes = (int)idx.size();

  // ---- shared memory layout ----
  int off = 0;  // doubles
  auto take = [&](int count) {
    const int o = off;
    off += (count + 1) / 2 * 2;  // keep 16-byte alignment
    return o;
  };
  O.o_x = take(nvar);
  O.o_l = take(nnz);
  const int panel = std::max(ni * O.W, m);
  const int stage_doubles = 3 * O.fseg_units * 2;
  O.stage = 1;
  O.o_r = take(std::max(panel, stage_doubles));
  O.o_env = -1;
  O.o_qsup = take(O.nsup);
  O.o_react = take(3 * nb);
  O.o_gxy = take(2 * nb * k + 2 * nb);
  O.o_azsup = take(((P.constraints & TNO_CON_REAC_BOUNDS) && P.var.o_zb >= 0) ? O.nsup * nb : 0);
  O.o_red = take(40);
  O.o_idx_bytes = off * 8;
  O.smem_bytes = O.o_idx_bytes + O.idx_bytes;
  int max_optin = 0, smem_sm = 0;
  TNO_CUDA_OK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, PL->device));
  TNO_CUDA_OK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, PL->device));
  if (O.smem_bytes > max_optin) return TNO_OK;
  PL->onchip_threads = THREADS;
  PL->onchip_ctas_per_sm = std::max(1, std::min(8, smem_sm / (O.smem_bytes + 1024)));

  std::vector<double> Bt((size_t)k * m);
  for (int e = 0; e < m; ++e)
    for (int j = 0; j < k; ++j) Bt[(size_t)j * m + e] = D->B[(size_t)e * k + j];
  std::vector<int> colsrc(nvar, -1);
  for (int j = 0; j < k; ++j) colsrc[j] = O.c_q >= 0 ? O.c_q + j : -2;
  if (P.var.o_lambdh >= 0) colsrc[P.var.o_lambdh] = O.c_lambdh >= 0 ? O.c_lambdh : -2;
  if (P.var.o_lambdv >= 0) colsrc[P.var.o_lambdv] = O.c_lambdv;

  int rc;
  const void* dptr = nullptr;
#define UPO(field, hostvec) \
  if ((rc = upload_vec(PL, hostvec, &O.field))) return rc;
  UPO(Bt, Bt);
  UPO(asm_ptr, asm_ptr);
  UPO(asm_edge, asm_edge);
  UPO(asm_coef, asm_coef);
  if ((rc = upload_bytes(PL, stream.data(), stream.size() * 4, &dptr))) return rc;
  O.f_stream = reinterpret_cast<const uint4*>(dptr);
  if ((rc = upload_bytes(PL, idx.data(), idx.size(), &dptr))) return rc;
  O.idx_init = reinterpret_cast<const unsigned char*>(dptr);
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
  const size_t smem = (size_t)PL->oc.smem_bytes;
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
