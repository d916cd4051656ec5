// Device side of the batched multi-start driver: S independent SQP iterations advance together, iterates and
// every callback output (f, grad f, g, dense J) stay in HBM between steps.  This file holds the two per-start kernels
// that are not evaluations of the thrust network:
//
//   k_sqp_qp    the quadratic subproblem of one SQP step,
//                   min 1/2 d'H d + c'd   s.t.  A d + b >= 0,  lb <= x + d <= ub
//               (what scipy's SLSQP solves with its least-squares QP, reference solvers/scipy.py:172 -> fmin_slsqp),
//               by a primal-dual interior point method (Mehrotra predictor-corrector) on the DENSE Jacobian A = J exactly
//               as tno_eval_batch left it in device memory.  nvar is small (tens), ng large (thousands): every iteration
//               forms the nvar x nvar normal matrix M = H + A' diag(lam / s) A -- a genuine dense contraction, optionally on
//               the FP64 tensor pipe (mma.sync m8n8k4) -- and reads A three times.  Rows shared by all starts (the
//               candidate-independent funicular rows [B 0; -B 0]) are read from one L2-resident copy.
//   k_sqp_bfgs  damped BFGS update of H with the Lagrangian gradients at the old and the new iterate.
//
// One CTA per start, grid-stride over the starts.  The host loop (compas_tno_b200/multistart.py) does the line search on
// the L1 merit function with batched evaluations.
#include <cstdio>
#include <cstdlib>

#include "tno_plan.hpp"

namespace tno {

constexpr int QP_NT = 256;     // threads per start
constexpr int QP_TR = 64;      // rows of A staged per tile
constexpr int QP_NVMAX = 64;   // variables per start
constexpr int QP_SLOTS = (QP_NVMAX * (QP_NVMAX + 1) / 2 + QP_NT - 1) / QP_NT;   // (i, j) pairs of M per thread
constexpr int QP_MT = 5;       // 8 x 8 tiles of M per warp on the tensor pipe: 36 lower tiles at nvar = 64 over 8 warps
constexpr int QP_ACC = QP_SLOTS > 2 * QP_MT ? QP_SLOTS : 2 * QP_MT;

struct QpArgs {
  int64_t S;
  int nv, ng, nshared;
  const double* H;      // S x nv x nv
  const double* c;      // S x nv        gradient of the objective
  const double* b;      // S x ng        constraint values g(x)
  const double* A;      // S x ng x nv   Jacobian; rows [0, nshared) are taken from Ash instead
  const double* Ash;    // nshared x nv  rows shared by all starts (may be null when nshared = 0)
  const double* x;      // S x nv
  const double* lb;     // nv
  const double* ub;     // nv
  const unsigned char* active;   // S or null: 0 = leave this start alone (d = 0)
  double* d;            // S x nv
  double* lam;          // S x (ng + 2 nv)   multipliers: constraints, lower bounds, upper bounds
  double* info;         // S x 8: c'd | sum lam |b| | sum max(0, -g) | max lam | IPM iterations | converged | mu | max |r_p|
  double* vold;         // S x nv: c - A_g' lam_g  (Lagrangian gradient at x with the new multipliers)
  double* work;         // S x 4 x (ng + 2 nv): s, r_p, ds, dl
  int max_iter;
  double tol;
  uint32_t nv_magic;    // ceil(2^32 / nv)
};

__device__ __forceinline__ double blk_sum(double v, double* red, int tid) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < QP_NT / 32; ++w) s += red[w];
  return s;
}
__device__ __forceinline__ double blk_min(double v, double* red, int tid) {
  for (int off = 16; off > 0; off >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, off));
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = v;
  __syncthreads();
  double s = red[0];
  for (int w = 1; w < QP_NT / 32; ++w) s = fmin(s, red[w]);
  return s;
}
__device__ __forceinline__ double blk_max(double v, double* red, int tid) { return -blk_min(-v, red, tid); }

// MMA: M += A_t' diag(w) A_t of one staged tile with mma.sync.aligned.m8n8k4.f64 instead of one FMA per (pair, row).
template <bool MMA>
__global__ void __launch_bounds__(QP_NT) k_sqp_qp(const QpArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nv = a.nv, ng = a.ng, mt = ng + 2 * nv;
  const int ldt = nv | 1;   // odd row stride of the tile: conflict-free column walks
  const int npairs = nv * (nv + 1) / 2;
  double* M = sm;                       // nv x nv
  double* tile = M + nv * nv;           // QP_TR x ldt
  double* wt = tile + QP_TR * ldt;      // per tile row: lam / s
  double* lt = wt + QP_TR;              //               lam
  double* ct = lt + QP_TR;              //               coefficient of the A' (...) accumulation of the pass
  double* it = ct + QP_TR;              //               1 / s
  double* dv = it + QP_TR;              // nv: d
  double* dd = dv + nv;                 // nv: Newton direction
  double* rd = dd + nv;                 // nv: dual residual
  double* ra = rd + nv;                 // nv: right-hand side of the affine step
  double* vi = ra + nv;                 // nv: A' (1 / s)
  double* rh = vi + nv;                 // nv: right-hand side being solved
  double* red = rh + nv;                // 16
  unsigned char* pi = reinterpret_cast<unsigned char*>(red + 16);
  unsigned char* pj = pi + npairs;
  for (int p = tid; p < npairs; p += QP_NT) {   // pair p = (i, j), j <= i, row-major over the lower triangle
    int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while (i * (i + 1) / 2 > p) --i;
    while ((i + 1) * (i + 2) / 2 <= p) ++i;
    pi[p] = (unsigned char)i;
    pj[p] = (unsigned char)(p - i * (i + 1) / 2);
  }
  for (int i = tid; i < QP_TR * ldt; i += QP_NT) tile[i] = 0.0;
  __syncthreads();
  const int vt = QP_NT - 1 - tid;   // the vector accumulations run on the LAST threads (the first carry the pairs)

  for (int64_t st = blockIdx.x; st < a.S; st += gridDim.x) {
    double* dout = a.d + st * nv;
    double* info = a.info + st * 8;
    if (a.active && !a.active[st]) {
      if (tid < nv) dout[tid] = 0.0;
      continue;
    }
    const double* H = a.H + st * nv * nv;
    const double* c = a.c + st * nv;
    const double* b = a.b + st * ng;
    const double* Aown = a.A + st * (int64_t)ng * nv;
    const double* x = a.x + st * nv;
    double* lam = a.lam + st * mt;
    double* s = a.work + st * 4 * mt;
    double* rp = s + mt;
    double* ds = rp + mt;
    double* dl = ds + mt;
    auto bound_b = [&](int r) -> double {   // r >= ng: value of the bound row
      const int i = r - ng;
      return i < nv ? x[i] - a.lb[i] : a.ub[i - nv] - x[i - nv];
    };
    auto load_tile = [&](int r0, int nrow) {
      const int64_t base = (int64_t)r0 * nv;
      for (int idx = tid; idx < QP_TR * nv; idx += QP_NT) {   // rows past the end are zero filled
        const int r = __umulhi((unsigned)idx, a.nv_magic);
        const double v = r < nrow ? (r0 + r < a.nshared ? a.Ash : Aown)[base + idx] : 0.0;
        tile[r * ldt + (idx - r * nv)] = v;
      }
    };
    // ---- start: d = 0, s = max(b, 1), lam = 1 ----
    for (int r = tid; r < mt; r += QP_NT) {
      const double br = r < ng ? b[r] : bound_b(r);
      const bool en = br < 1e15;   // an infinite bound never binds: kept out of the complementarity products
      s[r] = en ? fmax(br, 1.0) : 1.0;
      lam[r] = en ? 1.0 : 0.0;
    }
    if (tid < nv) dv[tid] = 0.0;
    double cmax = tid < nv ? fabs(c[tid]) : 0.0;
    cmax = 1.0 + blk_max(cmax, red, tid);
    double mcount = 0.0;
    for (int r = tid; r < mt; r += QP_NT) mcount += lam[r] > 0.0 ? 1.0 : 0.0;
    mcount = fmax(blk_sum(mcount, red, tid), 1.0);
    int iters = 0;
    bool converged = false;
    double mu = 0.0, rpmax = 0.0;
    for (; iters < a.max_iter; ++iters) {
      // ---- pass A: r_p, M = H + A' W A, A' lam, A' (-lam - w r_p), A' (1 / s) ----
      double acc[QP_ACC];
#pragma unroll
      for (int q = 0; q < QP_ACC; ++q) acc[q] = 0.0;
      double vA = 0.0, vB = 0.0, vD = 0.0, mu_p = 0.0, rp_p = 0.0;
      for (int r0 = 0; r0 < ng; r0 += QP_TR) {
        const int nrow = min(QP_TR, ng - r0);
        load_tile(r0, nrow);
        __syncthreads();
        if (tid < QP_TR) {
          double w = 0.0, l = 0.0, cf = 0.0, is = 0.0;
          if (tid < nrow) {
            const int r = r0 + tid;
            double ad = 0.0;
            for (int cc = 0; cc < nv; ++cc) ad = fma(tile[tid * ldt + cc], dv[cc], ad);
            const double sr = s[r];
            l = lam[r];
            const double res = ad + b[r] - sr;
            rp[r] = res;
            is = 1.0 / sr;
            w = fmin(l * is, 1e24);
            cf = -l - w * res;
            mu_p = fma(sr, l, mu_p);
            rp_p = fmax(rp_p, fabs(res));
          }
          wt[tid] = w;
          lt[tid] = l;
          ct[tid] = cf;
          it[tid] = is;
        }
        __syncthreads();
        if constexpr (MMA) {
          // M (nt x nt tiles of 8 x 8) += sum over the tile rows in steps of four: D = A_frag (8 x 4) B_frag (4 x 8) + C.
          // A_frag(i, kk) = w_kk A[kk, 8 ti + i], B_frag(kk, j) = A[kk, 8 tj + j].  Lane l holds A_frag(l / 4, l % 4),
          // B_frag(l % 4, l / 4) and C(l / 4, 2 (l % 4) + {0, 1}).  Tiles (ti >= tj) are dealt round-robin to the warps;
          // a warp's (at most QP_SLOTS / 2) tiles accumulate in acc[2 q], acc[2 q + 1].
          const int nt = (nv + 7) >> 3, ntiles = nt * (nt + 1) / 2;
#pragma unroll
          for (int q = 0; q < QP_MT; ++q) {
            const int tl = warp + q * (QP_NT / 32);
            if (tl < ntiles) {
              int ti = 0;
              while ((ti + 1) * (ti + 2) / 2 <= tl) ++ti;
              const int tj = tl - ti * (ti + 1) / 2;
              const int ia = 8 * ti + (lane >> 2), jb = 8 * tj + (lane >> 2), kk = lane & 3;
              const bool oka = ia < nv, okb = jb < nv;
              double c0 = acc[2 * q], c1 = acc[2 * q + 1];
              for (int t0 = 0; t0 < QP_TR; t0 += 4) {
                const double av = oka ? wt[t0 + kk] * tile[(t0 + kk) * ldt + ia] : 0.0;
                const double bv = okb ? tile[(t0 + kk) * ldt + jb] : 0.0;
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                             : "+d"(c0), "+d"(c1)
                             : "d"(av), "d"(bv));
              }
              acc[2 * q] = c0;
              acc[2 * q + 1] = c1;
            }
          }
        } else {
#pragma unroll
          for (int q = 0; q < QP_SLOTS; ++q) {
            const int p = tid + q * QP_NT;
            if (p < npairs) {
              const int i = pi[p], j = pj[p];
              double sum = 0.0;
#pragma unroll 8
              for (int t = 0; t < QP_TR; ++t) sum = fma(wt[t] * tile[t * ldt + i], tile[t * ldt + j], sum);
              acc[q] += sum;
            }
          }
        }
        if (vt < nv) {
#pragma unroll 8
          for (int t = 0; t < QP_TR; ++t) {
            const double av = tile[t * ldt + vt];
            vA = fma(lt[t], av, vA);
            vB = fma(ct[t], av, vB);
            vD = fma(it[t], av, vD);
          }
        }
        __syncthreads();
      }
      // M from the accumulators (+ H)
      if constexpr (MMA) {
        const int nt = (nv + 7) >> 3, ntiles = nt * (nt + 1) / 2;
#pragma unroll
        for (int q = 0; q < QP_MT; ++q) {
          const int tl = warp + q * (QP_NT / 32);
          if (tl < ntiles) {
            int ti = 0;
            while ((ti + 1) * (ti + 2) / 2 <= tl) ++ti;
            const int tj = tl - ti * (ti + 1) / 2;
            const int i = 8 * ti + (lane >> 2), j0 = 8 * tj + 2 * (lane & 3);
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const int j = j0 + u;
              if (i < nv && j < nv && j <= i) {
                const double v = acc[2 * q + u] + H[i * nv + j];
                M[i * nv + j] = v;
                M[j * nv + i] = v;
              }
            }
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < QP_SLOTS; ++q) {
          const int p = tid + q * QP_NT;
          if (p < npairs) {
            const int i = pi[p], j = pj[p];
            const double v = acc[q] + H[i * nv + j];
            M[i * nv + j] = v;
            M[j * nv + i] = v;
          }
        }
      }
      __syncthreads();
      if (vt < nv) {
        // bound rows of variable vt: +e (lower), -e (upper)
        const int i = vt, rl = ng + i, ru = ng + nv + i;
        const double dvi = dv[i];
        double mdiag = 0.0;
        {
          const double sr = s[rl], l = lam[rl], res = dvi + bound_b(rl) - sr;
          rp[rl] = res;
          if (l > 0.0) {
            const double w = l / sr;
            mdiag += w;
            vA += l;
            vB += -l - w * res;
            vD += 1.0 / sr;
            mu_p = fma(sr, l, mu_p);
            rp_p = fmax(rp_p, fabs(res));
          }
        }
        {
          const double sr = s[ru], l = lam[ru], res = -dvi + bound_b(ru) - sr;
          rp[ru] = res;
          if (l > 0.0) {
            const double w = l / sr;
            mdiag += w;
            vA -= l;
            vB -= -l - w * res;
            vD -= 1.0 / sr;
            mu_p = fma(sr, l, mu_p);
            rp_p = fmax(rp_p, fabs(res));
          }
        }
        M[i * nv + i] += mdiag;
        double hd = c[i];
        for (int j = 0; j < nv; ++j) hd = fma(H[i * nv + j], dv[j], hd);
        const double r = hd - vA;
        rd[i] = r;
        ra[i] = -r + vB;
        vi[i] = vD;
        rh[i] = -r + vB;
      }
      mu = blk_sum(mu_p, red, tid) / mcount;
      rpmax = blk_max(rp_p, red, tid);
      const double rdmax = blk_max(vt < nv ? fabs(rd[vt]) : 0.0, red, tid);
      if (!(isfinite(mu) && isfinite(rpmax) && isfinite(rdmax))) break;   // data of this start is not finite: give up, d as it is
      if (rdmax / cmax < a.tol && rpmax < a.tol && mu < a.tol) {
        converged = true;
        break;
      }
      // ---- Cholesky of M (whole CTA; trailing update over the pair table) ----
      const double mscale = blk_max(tid < nv ? fabs(M[tid * nv + tid]) : 0.0, red, tid);
      for (int jj = 0; jj < nv; ++jj) {
        if (tid == 0) {
          double pv = M[jj * nv + jj];
          if (!(pv > 1e-14 * mscale)) pv = 1e-14 * mscale;   // keeps the factor finite when M is numerically singular
          M[jj * nv + jj] = sqrt(pv);
        }
        __syncthreads();
        const double ljj = M[jj * nv + jj];
        for (int i = jj + 1 + tid; i < nv; i += QP_NT) M[i * nv + jj] /= ljj;
        __syncthreads();
        for (int p = tid; p < npairs; p += QP_NT) {
          const int i = pi[p], j = pj[p];
          if (j > jj) M[i * nv + j] = fma(-M[i * nv + jj], M[j * nv + jj], M[i * nv + j]);
        }
        __syncthreads();
      }
      auto solve = [&]() {   // rh <- M^-1 rh (warp 0), result copied to dd
        if (warp == 0) {
          for (int j = 0; j < nv; ++j) {
            const double y = rh[j] / M[j * nv + j];
            __syncwarp();
            if (lane == 0) rh[j] = y;
            for (int i = j + 1 + lane; i < nv; i += 32) rh[i] = fma(-M[i * nv + j], y, rh[i]);
            __syncwarp();
          }
          for (int j = nv - 1; j >= 0; --j) {
            const double y = rh[j] / M[j * nv + j];
            __syncwarp();
            if (lane == 0) rh[j] = y;
            for (int i = lane; i < j; i += 32) rh[i] = fma(-M[j * nv + i], y, rh[i]);
            __syncwarp();
          }
          for (int i = lane; i < nv; i += 32) dd[i] = rh[i];
        }
        __syncthreads();
      };
      solve();
      {
        // an inconsistent linearisation (or a numerically singular normal matrix) shows up as a non-finite direction:
        // stop with the last finite d and report "not converged"
        const double fin = blk_min(tid < nv ? (isfinite(dd[tid]) ? 1.0 : 0.0) : 1.0, red, tid);
        if (fin == 0.0) break;
      }
      // ---- pass B: affine step on the rows, A' (ds dl / s), step lengths ----
      double ap = 1.0, ad_ = 1.0, vC = 0.0;
      for (int r0 = 0; r0 < ng; r0 += QP_TR) {
        const int nrow = min(QP_TR, ng - r0);
        load_tile(r0, nrow);
        __syncthreads();
        if (tid < QP_TR) {
          double cf = 0.0;
          if (tid < nrow) {
            const int r = r0 + tid;
            double adot = 0.0;
            for (int cc = 0; cc < nv; ++cc) adot = fma(tile[tid * ldt + cc], dd[cc], adot);
            const double sr = s[r], l = lam[r];
            const double dsr = adot + rp[r];
            const double dlr = -l - l / sr * dsr;
            ds[r] = dsr;
            dl[r] = dlr;
            if (dsr < 0.0) ap = fmin(ap, -sr / dsr);
            if (dlr < 0.0) ad_ = fmin(ad_, -l / dlr);
            cf = dsr * dlr / sr;
          }
          ct[tid] = cf;
        }
        __syncthreads();
        if (vt < nv) {
#pragma unroll 8
          for (int t = 0; t < QP_TR; ++t) vC = fma(ct[t], tile[t * ldt + vt], vC);
        }
        __syncthreads();
      }
      if (vt < nv) {
        const int i = vt;
        for (int side = 0; side < 2; ++side) {
          const int r = ng + side * nv + i;
          const double sg = side ? -1.0 : 1.0;
          const double sr = s[r], l = lam[r];
          if (l > 0.0) {
            const double dsr = sg * dd[i] + rp[r];
            const double dlr = -l - l / sr * dsr;
            ds[r] = dsr;
            dl[r] = dlr;
            if (dsr < 0.0) ap = fmin(ap, -sr / dsr);
            if (dlr < 0.0) ad_ = fmin(ad_, -l / dlr);
            vC += sg * dsr * dlr / sr;
          } else {
            ds[r] = 0.0;
            dl[r] = 0.0;
          }
        }
      }
      ap = blk_min(ap, red, tid);
      ad_ = blk_min(ad_, red, tid);
      double ma = 0.0;
      for (int r = tid; r < mt; r += QP_NT)
        if (lam[r] > 0.0) ma = fma(s[r] + ap * ds[r], lam[r] + ad_ * dl[r], ma);
      ma = blk_sum(ma, red, tid) / mcount;
      double sigma = ma / mu;
      sigma = sigma * sigma * sigma;
      const double smu = sigma * mu;
      if (vt < nv) rh[vt] = ra[vt] + smu * vi[vt] - vC;
      __syncthreads();
      solve();
      {
        const double fin = blk_min(tid < nv ? (isfinite(dd[tid]) ? 1.0 : 0.0) : 1.0, red, tid);
        if (fin == 0.0) break;
      }
      // ---- pass D: corrected step on the rows, step lengths ----
      ap = 1.0;
      ad_ = 1.0;
      for (int r0 = 0; r0 < ng; r0 += QP_TR) {
        const int nrow = min(QP_TR, ng - r0);
        load_tile(r0, nrow);
        __syncthreads();
        if (tid < nrow) {
          const int r = r0 + tid;
          double adot = 0.0;
          for (int cc = 0; cc < nv; ++cc) adot = fma(tile[tid * ldt + cc], dd[cc], adot);
          const double sr = s[r], l = lam[r];
          const double rc = smu - sr * l - ds[r] * dl[r];
          const double dsr = adot + rp[r];
          const double dlr = (rc - l * dsr) / sr;
          ds[r] = dsr;
          dl[r] = dlr;
          if (dsr < 0.0) ap = fmin(ap, -sr / dsr);
          if (dlr < 0.0) ad_ = fmin(ad_, -l / dlr);
        }
        __syncthreads();
      }
      if (vt < nv) {
        const int i = vt;
        for (int side = 0; side < 2; ++side) {
          const int r = ng + side * nv + i;
          const double sg = side ? -1.0 : 1.0;
          const double sr = s[r], l = lam[r];
          if (l > 0.0) {
            const double rc = smu - sr * l - ds[r] * dl[r];
            const double dsr = sg * dd[i] + rp[r];
            const double dlr = (rc - l * dsr) / sr;
            ds[r] = dsr;
            dl[r] = dlr;
            if (dsr < 0.0) ap = fmin(ap, -sr / dsr);
            if (dlr < 0.0) ad_ = fmin(ad_, -l / dlr);
          }
        }
      }
      ap = fmin(1.0, 0.995 * blk_min(ap, red, tid));
      ad_ = fmin(1.0, 0.995 * blk_min(ad_, red, tid));
      for (int r = tid; r < mt; r += QP_NT)
        if (lam[r] > 0.0) {
          s[r] += ap * ds[r];
          lam[r] += ad_ * dl[r];
        }
      if (tid < nv) dv[tid] = fma(ap, dd[tid], dv[tid]);
      __syncthreads();
    }
    // ---- outputs: d, c'd, sum lam |b|, violation, Lagrangian gradient with the new multipliers ----
    double vL = 0.0;
    for (int r0 = 0; r0 < ng; r0 += QP_TR) {
      const int nrow = min(QP_TR, ng - r0);
      load_tile(r0, nrow);
      if (tid < QP_TR) lt[tid] = tid < nrow ? lam[r0 + tid] : 0.0;
      __syncthreads();
      if (vt < nv) {
#pragma unroll 8
        for (int t = 0; t < QP_TR; ++t) vL = fma(lt[t], tile[t * ldt + vt], vL);
      }
      __syncthreads();
    }
    if (vt < nv) a.vold[st * nv + vt] = c[vt] - vL;
    double lamb = 0.0, viol = 0.0, lmax = 0.0;
    for (int r = tid; r < mt; r += QP_NT) {
      const double br = r < ng ? b[r] : bound_b(r);
      const double l = lam[r];
      if (l > 0.0) lamb = fma(l, fabs(br), lamb);
      lmax = fmax(lmax, l);
      if (r < ng) viol += fmax(0.0, -br);
    }
    lamb = blk_sum(lamb, red, tid);
    viol = blk_sum(viol, red, tid);
    lmax = blk_max(lmax, red, tid);
    const double gd = blk_sum(tid < nv ? c[tid] * dv[tid] : 0.0, red, tid);
    if (tid < nv) dout[tid] = dv[tid];
    if (tid == 0) {
      info[0] = gd;
      info[1] = lamb;
      info[2] = viol;
      info[3] = lmax;
      info[4] = (double)iters;
      info[5] = converged ? 1.0 : 0.0;
      info[6] = mu;
      info[7] = rpmax;
    }
    __syncthreads();
  }
}

struct BfgsArgs {
  int64_t S;
  int nv, ng, nshared;
  double* H;             // S x nv x nv, updated in place
  const double* c_new;   // S x nv      gradient of the objective at the new iterate
  const double* A_new;   // S x ng x nv Jacobian at the new iterate
  const double* Ash;
  const double* lam;     // S x (ng + 2 nv)
  const double* vold;    // S x nv
  const double* step;    // S x nv      x_new - x_old
  const unsigned char* active;
  uint32_t nv_magic;
};

// y = (c_new - A_new' lam_g) - vold ;  Powell-damped BFGS:  H <- H - (H s)(H s)' / s'Hs + y y' / s'y
__global__ void __launch_bounds__(QP_NT) k_sqp_bfgs(const BfgsArgs a) {
  extern __shared__ __align__(16) double sm[];
  const int tid = threadIdx.x;
  const int nv = a.nv, ng = a.ng, mt = ng + 2 * nv;
  const int ldt = nv | 1;
  double* tile = sm;                 // QP_TR x ldt
  double* lt = tile + QP_TR * ldt;   // QP_TR
  double* y = lt + QP_TR;            // nv
  double* hs = y + nv;               // nv
  double* sv = hs + nv;              // nv
  double* red = sv + nv;             // 16
  for (int64_t st = blockIdx.x; st < a.S; st += gridDim.x) {
    if (a.active && !a.active[st]) continue;
    double* H = a.H + st * nv * nv;
    const double* Aown = a.A_new + st * (int64_t)ng * nv;
    const double* lam = a.lam + st * mt;
    double vL = 0.0;
    for (int r0 = 0; r0 < ng; r0 += QP_TR) {
      const int nrow = min(QP_TR, ng - r0);
      const int64_t base = (int64_t)r0 * nv;
      for (int idx = tid; idx < nrow * nv; idx += QP_NT) {
        const int r = __umulhi((unsigned)idx, a.nv_magic);
        tile[r * ldt + (idx - r * nv)] = (r0 + r < a.nshared ? a.Ash : Aown)[base + idx];
      }
      if (tid < QP_TR) lt[tid] = tid < nrow ? lam[r0 + tid] : 0.0;
      __syncthreads();
      if (tid < nv) {
        for (int t = 0; t < nrow; ++t) vL = fma(lt[t], tile[t * ldt + tid], vL);
      }
      __syncthreads();
    }
    if (tid < nv) {
      y[tid] = (a.c_new[st * nv + tid] - vL) - a.vold[st * nv + tid];
      sv[tid] = a.step[st * nv + tid];
    }
    __syncthreads();
    if (tid < nv) {
      double h = 0.0;
      for (int j = 0; j < nv; ++j) h = fma(H[tid * nv + j], sv[j], h);
      hs[tid] = h;
    }
    __syncthreads();
    const double shs = blk_sum(tid < nv ? sv[tid] * hs[tid] : 0.0, red, tid);
    double sy = blk_sum(tid < nv ? sv[tid] * y[tid] : 0.0, red, tid);
    if (sy < 0.2 * shs) {
      const double th = 0.8 * shs / (shs - sy);
      if (tid < nv) y[tid] = th * y[tid] + (1.0 - th) * hs[tid];
      __syncthreads();
      sy = blk_sum(tid < nv ? sv[tid] * y[tid] : 0.0, red, tid);
    }
    if (shs > 0.0 && sy > 0.0 && isfinite(shs) && isfinite(sy)) {
      for (int e = tid; e < nv * nv; e += QP_NT) {
        const int i = e / nv, j = e - i * nv;
        H[e] += -hs[i] * hs[j] / shs + y[i] * y[j] / sy;
      }
    }
    __syncthreads();
  }
}

static size_t qp_smem_bytes(int nv) {
  const int ldt = nv | 1;
  const size_t doubles = (size_t)nv * nv + (size_t)QP_TR * ldt + 4 * QP_TR + 6 * (size_t)nv + 16;
  return doubles * 8 + (size_t)nv * (nv + 1) + 16;
}

}  // namespace tno

using namespace tno;

extern "C" {

int tno_sqp_qp_solve(int64_t n_starts, int32_t nvar, int32_t ng, int32_t nshared, const double* H, const double* grad,
                     const double* g, const double* J, const double* J_shared, const double* x, const double* lb,
                     const double* ub, const unsigned char* active, double* d, double* lam, double* info, double* lagr_grad,
                     double* work, int32_t max_iter, double tol, int32_t use_mma, void* cuda_stream) {
  if (!H || !grad || !g || !J || !x || !lb || !ub || !d || !lam || !info || !lagr_grad || !work) {
    set_error("tno_sqp_qp_solve: null argument");
    return TNO_EINVAL;
  }
  if (n_starts < 0 || nvar < 1 || nvar > QP_NVMAX || ng < 1 || nshared < 0 || nshared > ng || (nshared > 0 && !J_shared)) {
    set_error("tno_sqp_qp_solve: bad sizes (1 <= nvar <= 64, 0 <= nshared <= ng, J_shared with nshared)");
    return TNO_EINVAL;
  }
  if (n_starts == 0) return TNO_OK;
  QpArgs a;
  a.S = n_starts;
  a.nv = nvar;
  a.ng = ng;
  a.nshared = nshared;
  a.H = H;
  a.c = grad;
  a.b = g;
  a.A = J;
  a.Ash = J_shared;
  a.x = x;
  a.lb = lb;
  a.ub = ub;
  a.active = active;
  a.d = d;
  a.lam = lam;
  a.info = info;
  a.vold = lagr_grad;
  a.work = work;
  a.max_iter = max_iter > 0 ? max_iter : 40;
  a.tol = tol > 0.0 ? tol : 1e-10;
  a.nv_magic = (uint32_t)((((uint64_t)1 << 32) + nvar - 1) / nvar);
  int dev = 0, sms = 148;
  TNO_CUDA_OK(cudaGetDevice(&dev));
  TNO_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const size_t smem = qp_smem_bytes(nvar);
  auto kern = use_mma ? k_sqp_qp<true> : k_sqp_qp<false>;
  TNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 1;
  TNO_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, QP_NT, smem));
  const int grid = (int)std::min<int64_t>(n_starts, (int64_t)sms * std::max(per_sm, 1));
  kern<<<grid, QP_NT, smem, (cudaStream_t)cuda_stream>>>(a);
  TNO_CUDA_OK(cudaGetLastError());
  return TNO_OK;
}

int tno_sqp_bfgs_update(int64_t n_starts, int32_t nvar, int32_t ng, int32_t nshared, double* H, const double* grad_new,
                        const double* J_new, const double* J_shared, const double* lam, const double* lagr_grad_old,
                        const double* step, const unsigned char* active, void* cuda_stream) {
  if (!H || !grad_new || !J_new || !lam || !lagr_grad_old || !step) {
    set_error("tno_sqp_bfgs_update: null argument");
    return TNO_EINVAL;
  }
  if (n_starts < 0 || nvar < 1 || nvar > QP_NVMAX || ng < 1 || nshared < 0 || nshared > ng || (nshared > 0 && !J_shared)) {
    set_error("tno_sqp_bfgs_update: bad sizes");
    return TNO_EINVAL;
  }
  if (n_starts == 0) return TNO_OK;
  BfgsArgs a;
  a.S = n_starts;
  a.nv = nvar;
  a.ng = ng;
  a.nshared = nshared;
  a.H = H;
  a.c_new = grad_new;
  a.A_new = J_new;
  a.Ash = J_shared;
  a.lam = lam;
  a.vold = lagr_grad_old;
  a.step = step;
  a.active = active;
  a.nv_magic = (uint32_t)((((uint64_t)1 << 32) + nvar - 1) / nvar);
  int dev = 0, sms = 148;
  TNO_CUDA_OK(cudaGetDevice(&dev));
  TNO_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int ldt = nvar | 1;
  const size_t smem = ((size_t)QP_TR * ldt + QP_TR + 3 * (size_t)nvar + 16) * 8;
  const int grid = (int)std::min<int64_t>(n_starts, (int64_t)sms * 4);
  k_sqp_bfgs<<<grid, QP_NT, smem, (cudaStream_t)cuda_stream>>>(a);
  TNO_CUDA_OK(cudaGetLastError());
  return TNO_OK;
}

}  // extern "C"
