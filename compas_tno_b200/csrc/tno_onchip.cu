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
#include <cstdlib>
#include <cstring>
#include <map>
#include <numeric>

#include "tno_plan.hpp"
#ifndef TNO_AB
#define TNO_AB 8
#endif
#ifndef TNO_QU
#define TNO_QU 4
#endif

namespace tno {

struct OcArgs {
  const double* x;
  int64_t ldx;
  const double* pzs;
  const double* dir;   // N x 2 load directions or null
  const double* bounds;  // N x 2n per-candidate [ub | lb] or null
  double* f;
  double* grad;
  double* g;
  double* J;
  double* z;
  double* q;
  double* gmin;
  int32_t* status;
  int64_t N;
  long long* phase_clocks;  // optional (TNO_PHASE_CLOCKS=1): clock64() of CTA 0 at the phase boundaries of its first candidate
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

// Reciprocal of a positive double without the IEEE division's slow path: hardware seed (about 20 bits) and two Newton
// steps on the FP64 pipe; the result is within an ulp or two of 1 / d.
__device__ __forceinline__ double oc_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  e = fma(-d, x, 1.0);
  x = fma(x, e, x);
  return x;
}

// Barrier of one candidate group.  A CTA holds G groups of NT threads, one candidate each, which share the index region
// but nothing else: every group synchronises on its own named barrier (id = group + 1); a single-group CTA keeps
// __syncthreads().
template <int NT>
__device__ __forceinline__ void oc_sync(int bar) {
  if (bar == 0) __syncthreads();
  else asm volatile("bar.sync %0, %1;" ::"r"(bar), "n"(NT) : "memory");
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

// ---- kernels of the triangular solves ----------------------------------------------------------------
// G = 2^gshift lanes cover the columns of a row, TWO adjacent columns per lane (one 16-byte shared-memory
// access), S = 2^sshift sub-slots split the entries of a row.  Rows [row0, row1) are mutually independent.

// forward, pull form: y_i -= sum_j L_ij y_j over the CSR entries of row i; sub-slot partial sums are combined
// with shuffles (S * G <= 32)
template <int NT>
__device__ __forceinline__ void oc_pull_rows(int row0, int row1, int sshift, int gshift, int tl, int c0, int ncols, int W,
                                             const unsigned short* __restrict__ rowptr16,
                                             const unsigned short* __restrict__ rowcol16, const double* __restrict__ V,
                                             double* __restrict__ R, const unsigned short* __restrict__ rowlist = nullptr) {
  const int lshift = sshift + gshift;
  const int rows = row1 - row0;
  const int r = tl & ((1 << gshift) - 1);
  const int sub = (tl >> gshift) & ((1 << sshift) - 1);
  const int S = 1 << sshift;
  const bool colok = 2 * r < ncols;
  const int col = c0 + 2 * r;
  for (int base = 0; base < rows; base += NT >> lshift) {
    const int it = base + (tl >> lshift);
    const bool valid = it < rows && colok;
    const int row = rowlist ? (it < rows ? (int)rowlist[it] : row0) : row0 + it;   // rowlist: processing order of the level's rows
    double ax = 0.0, ay = 0.0;
    if (valid) {
      const int t1 = rowptr16[row + 1];
#pragma unroll 4
      for (int t = rowptr16[row] + sub; t < t1; t += S) {
        const double l = V[t];
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

// backward, push form: x_i is final -> x_j -= L_ij x_i for every CSR entry j of row i.  A vertex has at most one
// ancestor per level (and per chain level), so the targets of independent rows are disjoint: no atomics.
template <int NT>
__device__ __forceinline__ void oc_push_rows(int row0, int row1, int sshift, int gshift, int tl, int c0, int ncols, int W,
                                             const unsigned short* __restrict__ rowptr16,
                                             const unsigned short* __restrict__ rowcol16, const double* __restrict__ V,
                                             double* __restrict__ R, const unsigned short* __restrict__ rowlist = nullptr) {
  const int lshift = sshift + gshift;
  const int rows = row1 - row0;
  const int r = tl & ((1 << gshift) - 1);
  const int sub = (tl >> gshift) & ((1 << sshift) - 1);
  const int S = 1 << sshift;
  const bool colok = 2 * r < ncols;
  const int col = c0 + 2 * r;
  for (int base = 0; base < rows; base += NT >> lshift) {
    const int it = base + (tl >> lshift);
    if (it < rows && colok) {
      const int row = rowlist ? (int)rowlist[it] : row0 + it;
      const double2 xi = *reinterpret_cast<const double2*>(R + row * W + col);
      const int t1 = rowptr16[row + 1];
#pragma unroll 4
      for (int t = rowptr16[row] + sub; t < t1; t += S) {
        const double l = V[t];
        double2* d = reinterpret_cast<double2*>(R + (int)rowcol16[t] * W + col);
        double2 v = *d;
        v.x = fma(-l, xi.x, v.x);
        v.y = fma(-l, xi.y, v.y);
        *d = v;
      }
    }
  }
}

// Dense part of a chain level: y_C = Z_C r_C (forward, TRANSPOSED = false) or x_C = Z_C^T y_C (backward), Z unit lower
// triangular stored packed in V.  All rows of the level are computed into registers first, then written back.
constexpr int OC_ZPASS = 4;
template <int NT, bool TRANSPOSED>
__device__ __forceinline__ void oc_chain_apply(const OcRange& L, const OcChain* __restrict__ chains,
                                               const unsigned char* __restrict__ rowchain, int wide_rows, int gshift, int tl, int c0,
                                               int ncols, int W, const double* __restrict__ V, double* __restrict__ R, int bar) {
  const int rows = L.row1 - L.row0;
  const int r = tl & ((1 << gshift) - 1);
  const bool colok = 2 * r < ncols;
  const int col = c0 + 2 * r;
  const int B = OC_ZPASS * (NT >> gshift);  // rows per batch
  // forward results only depend on LOWER rows of the chain, so batches run from the top rows down (and the other way
  // round for the transposed product): a batch never reads what an earlier batch has overwritten.
  for (int done = 0; done < rows; done += B) {
    const int lo = TRANSPOSED ? done : max(0, rows - done - B);
    const int hi = TRANSPOSED ? min(rows, done + B) : rows - done;
    double2 res[OC_ZPASS];
#pragma unroll
    for (int ps = 0; ps < OC_ZPASS; ++ps) {
      const int it = lo + ps * (NT >> gshift) + (tl >> gshift);
      res[ps] = make_double2(0.0, 0.0);
      if (it < hi && colok) {
        const int row = L.row0 + it;
        const OcChain C = chains[L.chain0 + rowchain[row - wide_rows]];
        const int il = row - C.row0;
        double2 acc = *reinterpret_cast<const double2*>(R + row * W + col);
        if (!TRANSPOSED) {
          const double* z = V + C.zoff + (il * (il - 1)) / 2;
#pragma unroll 4
          for (int j = 0; j < il; ++j) {
            const double2 y = *reinterpret_cast<const double2*>(R + (C.row0 + j) * W + col);
            acc.x = fma(z[j], y.x, acc.x);
            acc.y = fma(z[j], y.y, acc.y);
          }
        } else {
#pragma unroll 4
          for (int i = il + 1; i < C.T; ++i) {
            const double zz = V[C.zoff + (i * (i - 1)) / 2 + il];
            const double2 y = *reinterpret_cast<const double2*>(R + (C.row0 + i) * W + col);
            acc.x = fma(zz, y.x, acc.x);
            acc.y = fma(zz, y.y, acc.y);
          }
        }
        res[ps] = acc;
      }
    }
    oc_sync<NT>(bar);
#pragma unroll
    for (int ps = 0; ps < OC_ZPASS; ++ps) {
      const int it = lo + ps * (NT >> gshift) + (tl >> gshift);
      if (it < hi && colok) *reinterpret_cast<double2*>(R + (L.row0 + it) * W + col) = res[ps];
    }
  }
}

// The same dense products on the FP64 tensor pipe: Y_C = Z_C R_C as 8 x 8 output tiles, one warp per tile,
// mma.sync.m8n8k4: D (8 rows x 8 panel columns) = sum over 4-row blocks of  Z[tile rows, block] x R_C[block, columns],
// the unit diagonal enters through the accumulator (C = R_C tile).  A lane loads ONE element of Z and ONE of R per 256
// multiply-adds (the FMA variant above: a 16-byte panel load and a broadcast Z load per 2), which is what the kernel is
// short of: shared-memory bandwidth and issue slots.  Tiles are ordered so that a batch never reads rows an earlier
// batch has overwritten (forward: high rows first; transposed: low rows first).
constexpr int OC_ZMMA_ITEMS = 3;   // work items (8 rows x 16 columns) per warp and batch
template <int NT, bool TRANSPOSED>
__device__ __forceinline__ void oc_chain_apply_mma(const OcRange& L, const uint2* __restrict__ zt, int tl, int c0, int ncols, int W,
                                                   const double* __restrict__ V, double* __restrict__ R, int bar) {
  constexpr int NW = NT / 32;
  const int lane = tl & 31, warp = tl >> 5;
  const int g = lane >> 2, kq = lane & 3;
  const int ncols_e = (ncols + 1) & ~1;
  const int npair = (ncols + 15) >> 4;          // pairs of 8-column tiles: one work item = an 8-row tile x one pair
  const int nmt = (int)L.zt1 - (int)L.zt0;
  const int nitems = nmt * npair;
  for (int base = 0; base < nitems; base += NW * OC_ZMMA_ITEMS) {
    double2 res[OC_ZMMA_ITEMS][2];
#pragma unroll
    for (int ps = 0; ps < OC_ZMMA_ITEMS; ++ps) {
      const int t = base + ps * NW + warp;
      res[ps][0] = res[ps][1] = make_double2(0.0, 0.0);
      if (t < nitems) {   // warp-uniform
        const int di = t / npair, np = t - di * npair;
        const uint2 d = zt[(int)L.zt0 + (TRANSPOSED ? nmt - 1 - di : di)];
        const int zoff = (int)(d.x & 0xffffu), row0 = (int)(d.x >> 16), T = (int)(d.y & 0xffu), mt = (int)((d.y >> 8) & 0xffu);
        const int i = 8 * mt + g;
        const bool rok = i < T;
        const int colc = 16 * np + 2 * kq;      // first output column of this lane in the first tile (second: + 8)
        double2 ca = make_double2(0.0, 0.0), cb2 = make_double2(0.0, 0.0);
        if (rok && colc < ncols_e) ca = *reinterpret_cast<const double2*>(R + (row0 + i) * W + c0 + colc);
        if (rok && colc + 8 < ncols_e) cb2 = *reinterpret_cast<const double2*>(R + (row0 + i) * W + c0 + colc + 8);
        const bool bok0 = 16 * np + g < ncols_e, bok1 = 16 * np + 8 + g < ncols_e;
        const double* rb = R + row0 * W + c0 + 16 * np + g;
        const int nkb = (T + 3) >> 2;
        const int kb0 = TRANSPOSED ? 2 * mt : 0;
        const int kb1 = TRANSPOSED ? nkb : min(nkb, 2 * mt + 2);
        const int ri = zoff + (i * (i - 1)) / 2;
        int kk = 4 * kb0 + kq;
        int zo = zoff + (kk * (kk - 1)) / 2 + i;   // transposed: position of Z[kk][i]
#pragma unroll 2
        for (int kb = kb0; kb < kb1; ++kb) {
          double a = 0.0;
          if (!TRANSPOSED) {
            if (rok && kk < i) a = V[ri + kk];
          } else {
            if (rok && kk > i && kk < T) a = V[zo];
          }
          const double b0 = (bok0 && kk < T) ? rb[kk * W] : 0.0;
          const double b1 = (bok1 && kk < T) ? rb[kk * W + 8] : 0.0;
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(ca.x), "+d"(ca.y)
                       : "d"(a), "d"(b0));
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                       : "+d"(cb2.x), "+d"(cb2.y)
                       : "d"(a), "d"(b1));
          zo += 4 * kk + 6;
          kk += 4;
        }
        res[ps][0] = ca;
        res[ps][1] = cb2;
      }
    }
    oc_sync<NT>(bar);
#pragma unroll
    for (int ps = 0; ps < OC_ZMMA_ITEMS; ++ps) {
      const int t = base + ps * NW + warp;
      if (t < nitems) {
        const int di = t / npair, np = t - di * npair;
        const uint2 d = zt[(int)L.zt0 + (TRANSPOSED ? nmt - 1 - di : di)];
        const int row0 = (int)(d.x >> 16), T = (int)(d.y & 0xffu), mt = (int)((d.y >> 8) & 0xffu);
        const int i = 8 * mt + g, colc = 16 * np + 2 * kq;
        if (i < T && colc < ncols_e) *reinterpret_cast<double2*>(R + (row0 + i) * W + c0 + colc) = res[ps][0];
        if (i < T && colc + 8 < ncols_e) *reinterpret_cast<double2*>(R + (row0 + i) * W + c0 + colc + 8) = res[ps][1];
      }
    }
    if (base + NW * OC_ZMMA_ITEMS < nitems) oc_sync<NT>(bar);
  }
}

struct OcWarpFactor {
  const uint32_t* off;    // [warp]: start of the warp's record stream in wf_stream, 16-byte units (NW + 1 entries)
  const uint32_t* desc;   // [warp][level]: offset inside the warp's buffer (16-byte units) | items << 16 | log2 lanes per item << 28
};
struct OcWarpLocal {
  const unsigned short* ptr;
  const unsigned short* rows;
  const unsigned char* s;
};

// L D L^T solve of the panel columns [c0, c0 + ncols) (c0 even) for all rows.
template <int THREADS>
__device__ __forceinline__ void oc_solve(const DevOnchip& O, int phase, const OcRange* __restrict__ wide,
                                         const OcRange* __restrict__ clev, const OcChain* __restrict__ chains,
                                         const unsigned char* __restrict__ rowchain, const unsigned short* __restrict__ rowptr16,
                                         const unsigned short* __restrict__ rowcol16, const double* __restrict__ V,
                                         double* __restrict__ R, int ni, int c0, int ncols, int tid,
                                         const unsigned short* ct_ptr, const unsigned short* ct_col, const unsigned short* ct_ent,
                                         const unsigned char* ct_row, int bar, const uint2* __restrict__ zt,
                                         const unsigned short* __restrict__ rowlist, const OcWarpLocal wl, long long* tc = nullptr) {
  const int W = O.W;
  const int gshift = O.gshift[phase];
  long long tq = tc ? clock64() : 0;
  auto lap = [&](int slot) {
    if (tc) {
      const long long now = clock64();
      tc[slot] += now - tq;
      tq = now;
    }
  };
  // ---- forward ----
  const int wlw = tid >> 5, wll = tid & 31;
  if (O.wl_on) {
    // every warp runs the rows of its own subtrees level by level: no block-wide barrier inside the wide part
    for (int lev = 1; lev < O.nwide; ++lev) {
      const int a = wl.ptr[wlw * (O.nwide + 1) + lev], b = wl.ptr[wlw * (O.nwide + 1) + lev + 1];
      if (b > a)
        oc_pull_rows<32>(0, b - a, wl.s[((wlw * O.nwide + lev) * 2 + phase) * 2], gshift, wll, c0, ncols, W, rowptr16, rowcol16, V, R,
                         wl.rows + a);
      __syncwarp();
    }
    oc_sync<THREADS>(bar);
  } else {
    for (int lev = 1; lev < O.nwide; ++lev) {
      const OcRange L = wide[lev];
      oc_pull_rows<THREADS>(L.row0, L.row1, L.pull_s[phase], gshift, tid, c0, ncols, W, rowptr16, rowcol16, V, R);
      oc_sync<THREADS>(bar);
    }
  }
  lap(0);
  for (int cl = 0; cl < O.nclev; ++cl) {
    const OcRange L = clev[cl];
    oc_pull_rows<THREADS>(L.row0, L.row1, L.pull_s[phase], gshift, tid, c0, ncols, W, rowptr16, rowcol16, V, R,
                          L.rl_on ? rowlist + L.rl0 : nullptr);
    oc_sync<THREADS>(bar);
    lap(1);
    if (O.zmma) oc_chain_apply_mma<THREADS, false>(L, zt, tid, c0, ncols, W, V, R, bar);
    else oc_chain_apply<THREADS, false>(L, chains, rowchain, O.wide_rows, gshift, tid, c0, ncols, W, V, R, bar);
    oc_sync<THREADS>(bar);
    lap(2);
  }
  // ---- diagonal: y_i / d_i ----
  {
    const int r = tid & ((1 << gshift) - 1);
    if (2 * r < ncols) {
      for (int row = tid >> gshift; row < ni; row += THREADS >> gshift) {
        double2* d = reinterpret_cast<double2*>(R + row * W + c0 + 2 * r);
        const double s = V[O.nlo + row];
        double2 v = *d;
        v.x *= s;
        v.y *= s;
        *d = v;
      }
    }
  }
  oc_sync<THREADS>(bar);
  lap(3);
  // ---- backward ----
  for (int cl = O.nclev - 1; cl >= 0; --cl) {
    const OcRange L = clev[cl];
    if (O.zmma) oc_chain_apply_mma<THREADS, true>(L, zt, tid, c0, ncols, W, V, R, bar);
    else oc_chain_apply<THREADS, true>(L, chains, rowchain, O.wide_rows, gshift, tid, c0, ncols, W, V, R, bar);
    oc_sync<THREADS>(bar);
    lap(4);
    // the rows of one chain share their descendants, so their updates are gathered per outside column ("target")
    // instead of pushed per row: x_j -= sum_{i in chain} L_ij x_i.  Targets of one chain level are disjoint.
    {
      const int r = tid & ((1 << gshift) - 1);
      const int col = c0 + 2 * r;
      if (2 * r < ncols) {
        for (int t = L.tgt0 + (tid >> gshift); t < L.tgt1; t += THREADS >> gshift) {
          const int e1 = ct_ptr[t + 1];
          double ax = 0.0, ay = 0.0;
#pragma unroll 4
          for (int e = ct_ptr[t]; e < e1; ++e) {
            const double l = V[ct_ent[e]];
            const double2 x = *reinterpret_cast<const double2*>(R + (L.row0 + (int)ct_row[e]) * W + col);
            ax = fma(l, x.x, ax);
            ay = fma(l, x.y, ay);
          }
          double2* d = reinterpret_cast<double2*>(R + (int)ct_col[t] * W + col);
          double2 v = *d;
          v.x -= ax;
          v.y -= ay;
          *d = v;
        }
      }
    }
    oc_sync<THREADS>(bar);
    lap(5);
  }
  if (O.wl_on) {
    for (int lev = O.nwide - 1; lev >= 1; --lev) {
      const int a = wl.ptr[wlw * (O.nwide + 1) + lev], b = wl.ptr[wlw * (O.nwide + 1) + lev + 1];
      if (b > a)
        oc_push_rows<32>(0, b - a, wl.s[((wlw * O.nwide + lev) * 2 + phase) * 2 + 1], gshift, wll, c0, ncols, W, rowptr16, rowcol16, V, R,
                         wl.rows + a);
      __syncwarp();
    }
    oc_sync<THREADS>(bar);
  } else {
    for (int lev = O.nwide - 1; lev >= 1; --lev) {
      const OcRange L = wide[lev];
      oc_push_rows<THREADS>(L.row0, L.row1, L.push_s[phase], gshift, tid, c0, ncols, W, rowptr16, rowcol16, V, R);
      oc_sync<THREADS>(bar);
    }
  }
  lap(6);
}

// Dense work of one chain level, whole CTA, all chains of the level at once:
//   (B) right-looking L D L^T of every dense triangle, ONE barrier per column step: element (i, k) subtracts
//       u_ij u_kj / d_j for every j < k; the owner of a diagonal inverts it as soon as it is final, so the
//       division overlaps the other threads' updates.  Columns are scaled to L = U D^-1 afterwards.
//   (C) Z = L^-1: one thread per column solves L z = e_c by forward substitution into a second buffer
//       (columns are independent, no barriers), then the blocks are copied back over L.
template <int THREADS>
__device__ __forceinline__ void oc_chain_dense(const DevOnchip& O, const OcDense& DN, double* __restrict__ V,
                                               double* __restrict__ tmpz, int* s_flags, int tid, int bar, long long* t_b) {
  long long tb0 = t_b ? clock64() : 0;
  auto lapd = [&](int slot) {
    if (t_b) {
      const long long now = clock64();
      t_b[slot] += now - tb0;
      tb0 = now;
    }
  };
  const int zbase = O.nlo + (O.nv - O.nlo - O.nz);  // first dense position (= nlo + ni)
  // ---- (B) ----
  // every thread keeps its (up to OC_DE) element descriptors in registers for the whole factorisation of the level
  constexpr int OC_DE = 4;
  uint4 dsc[OC_DE];
  const int nde = (int)(DN.de1 - DN.de0);
#pragma unroll
  for (int q = 0; q < OC_DE; ++q) {
    const int e = q * THREADS + tid;
    dsc[q] = e < nde ? __ldg(O.de + DN.de0 + e) : make_uint4(0u, 0u, 0x00ffu, 0u);  // k = 255 > i = 0: never active
  }
  const bool in_regs = nde <= OC_DE * THREADS;
  auto step_elem = [&](const uint4 d, int j) {
    const int kk = (int)(d.z & 0xffu), ii = (int)((d.z >> 8) & 0xffu);
    if (kk > j && kk <= ii) {
      const int sp = (int)(d.y & 0xffffu);
      const double inv = V[(int)(d.y >> 16) + j];
      const double uij = V[(int)(d.x & 0xffffu) + j];
      const double ukj = V[(int)(d.x >> 16) + j];
      double sv = fma(-uij * inv, ukj, V[sp]);
      if (kk == ii && j == ii - 1) {  // diagonal just became final
        if (!(sv > 0.0)) atomicOr(s_flags, TNO_STATUS_NOT_PD);
        sv = oc_rcp(sv);
      }
      V[sp] = sv;
    }
  };
  for (uint32_t e = DN.de0 + tid; e < DN.de1; e += THREADS) {
    const uint4 d = __ldg(O.de + e);
    if ((d.z & 0xffffu) == 0u) {  // element (0, 0): final after the S-step
      const int sp = (int)(d.y & 0xffffu);
      const double v = V[sp];
      if (!(v > 0.0)) atomicOr(s_flags, TNO_STATUS_NOT_PD);
      V[sp] = oc_rcp(v);
    }
  }
  oc_sync<THREADS>(bar);
  lapd(0);
  for (int j = 0; j + 1 < (int)DN.tmax; ++j) {
    if (in_regs) {
      // elements are sorted by descending k: once a thread's first element is inactive, all of them are
      if ((int)(dsc[0].z & 0xffu) > j && (int)(dsc[0].z & 0xffu) != 0xff) {
#pragma unroll
        for (int q = 0; q < OC_DE; ++q) step_elem(dsc[q], j);
      }
    } else {
      for (uint32_t e = DN.de0 + tid; e < DN.de1; e += THREADS) step_elem(__ldg(O.de + e), j);
    }
    oc_sync<THREADS>(bar);
  }
  for (uint32_t e = DN.de0 + tid; e < DN.de1; e += THREADS) {
    const uint4 d = __ldg(O.de + e);
    const int kk = (int)(d.z & 0xffu), ii = (int)((d.z >> 8) & 0xffu);
    if (kk < ii) V[(int)(d.y & 0xffffu)] *= V[(int)(d.y >> 16) + kk];  // l_ik = u_ik / d_k
  }
  lapd(1);
  oc_sync<THREADS>(bar);
  lapd(2);
  // ---- (C) ----
  // Z = L^-1 by 8 x 8 blocks.  Stage 0: the diagonal blocks, one thread per column (at most 7 dependent rows).  Stage s =
  // 1 .. nb - 1: the blocks (ib, ib - s) of the s-th block sub-diagonal, one warp per block, two elements per lane
  // (lane = column c + 8 (row mod 4)):
  //     M = sum_{kb = jb .. ib - 1} L(ib, kb) Z(kb, jb)      (everything it reads comes from earlier stages)
  //     Z(ib, jb) = - Z(ib, ib) M                             (M passes between the lanes with shuffles)
  // i.e. nb - 1 barrier-separated steps for a chain of 8 nb rows instead of a dependent sweep over all of its rows.
  for (uint32_t cidx = DN.dc0 + tid; cidx < DN.dc1; cidx += THREADS) {
    const uint32_t w = __ldg(O.dc + cidx);
    const int zoff = (int)(w & 0xffffu), T = (int)((w >> 16) & 0xffu), c = (int)(w >> 24);
    const double* Lz = V + zoff;
    double* Zt = tmpz + (zoff - zbase);
    const int iend = min(T, (c | 7) + 1);
    for (int i = c + 1; i < iend; ++i) {
      const int rb = (i * (i - 1)) / 2;
      double a = Lz[rb + c];
      for (int kk = c + 1; kk < i; ++kk) a = fma(Lz[rb + kk], Zt[(kk * (kk - 1)) / 2 + c], a);
      Zt[rb + c] = -a;
    }
  }
  oc_sync<THREADS>(bar);
  lapd(3);
  {
    const int lane = tid & 31, warp = tid >> 5;
    const int c = lane & 7, r0 = lane >> 3;
    for (int s = 1; s < (int)DN.nbmax; ++s) {
      for (uint32_t bi = DN.db0 + DN.sb[s] + warp; bi < DN.db0 + DN.sb[s + 1]; bi += THREADS / 32) {
        const uint32_t w = __ldg(O.db + bi);
        const int zoff = (int)(w & 0xffffu), T = (int)((w >> 16) & 0xffu), ib = (int)((w >> 24) & 0xfu);
        const double* Lz = V + zoff;
        double* Zt = tmpz + (zoff - zbase);
        const int j = 8 * (ib - s) + c, kend = 8 * ib;
        double mm[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 8 * ib + r0 + 4 * h;
          double a = 0.0;
          if (i < T) {
            const int rb = (i * (i - 1)) / 2;
            a = Lz[rb + j];
#pragma unroll 4
            for (int k = j + 1; k < kend; ++k) a = fma(Lz[rb + k], Zt[(k * (k - 1)) / 2 + j], a);
          }
          mm[h] = a;
        }
        double acc[2] = {0.0, 0.0};
#pragma unroll
        for (int kr = 0; kr < 8; ++kr) {
          const double mk = __shfl_sync(0xffffffffu, kr < 4 ? mm[0] : mm[1], ((kr & 3) << 3) | c);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int r = r0 + 4 * h, i = 8 * ib + r;
            if (kr == r) acc[h] += mk;
            else if (kr < r && i < T) acc[h] = fma(Zt[(i * (i - 1)) / 2 + 8 * ib + kr], mk, acc[h]);
          }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int i = 8 * ib + r0 + 4 * h;
          if (i < T) Zt[(i * (i - 1)) / 2 + j] = -acc[h];
        }
      }
      oc_sync<THREADS>(bar);
    }
  }
  lapd(4);
  oc_sync<THREADS>(bar);
  for (uint32_t p = DN.zlo + tid; p < DN.zhi; p += THREADS) V[p] = tmpz[p - zbase];
  oc_sync<THREADS>(bar);
  lapd(5);
}

// Pull-form items of one staged piece of the factorisation: item = one entry of L (position, diagonal flag) with its
// list of (a, b, d) triples; 2^lpi_shift lanes share an item's dot product.  NT threads (the CTA, or one warp for the
// warp-local wide levels) walk the items.
template <int NT>
__device__ __forceinline__ void oc_factor_items(const uint2* __restrict__ rec, const unsigned long long* __restrict__ pairs, int nitems,
                                                int lpi_shift, int kind, int tl, double* __restrict__ V, double* __restrict__ tmpbuf,
                                                int* s_flags) {
  const int lpi = 1 << lpi_shift;
  const int sub = tl & (lpi - 1);
  const int per_pass = NT >> lpi_shift;
  for (int base = 0; base < nitems; base += per_pass) {
    const int it = base + (tl >> lpi_shift);
    const bool valid = it < nitems;
    double acc = 0.0;
    unsigned pf = 0;
    if (valid) {
      const uint2 r0 = rec[it];
      pf = r0.x;
      const int t1 = rec[it + 1].y;
#pragma unroll 2
      for (int t = r0.y + sub; t < t1; t += lpi) {
        const unsigned long long pr = pairs[t];
        const int a = (int)(pr & 0xffffu), b = (int)((pr >> 16) & 0xffffu), dp = (int)((pr >> 32) & 0xffffu);
        acc = fma(V[a] * V[dp], V[b], acc);  // U_ic * (1 / d_c) * U_jc   (rectangular rows: S_ic * (-1) * Z_jc)
      }
    }
    for (int off = lpi >> 1; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (valid && sub == 0) {
      const int pos = (int)(pf & 0xffffu);
      const double v = V[pos] - acc;
      if (kind == 1) {
        tmpbuf[it] = v;
      } else if (pf >> 31) {
        if (!(v > 0.0)) atomicOr(s_flags, TNO_STATUS_NOT_PD);
        V[pos] = oc_rcp(v);
      } else {
        V[pos] = v;
      }
    }
  }
}

// Register-tile variant of the dense work of one chain level.  Every thread owns one TS x TS tile (ti, tj), ti >= tj, of a
// chain triangle (TS = 2, or 4 when the level has more 2 x 2 tiles than threads) and keeps two things in registers for the
// whole level: the tile of S (becoming U = L D column by column) and the same tile of Z = L^-1.  Step j = 0 .. T - 2 is
// ONE barrier-separated step for both:
//     l_ij = u_ij / d_j                       (column j of U and 1 / d_j come from a small exchange buffer)
//     S_ik -= l_ij u_kj      for k > j        (right-looking L D L^T, the same arithmetic as the element-wise variant)
//     Z_i: -= l_ij Z_j:      for i > j        (the row operations that reduce L to the identity, applied to the identity)
// after which the owners of column j + 1 / row j + 1 publish them (double buffered) and the owner of the diagonal
// inverts d_{j+1}.  L itself is never stored: the solves only use Z and 1 / d.  Small tiles keep the FP64 work of a step
// short per warp (a warp issues one DFMA every ~4 cycles: tools/latbench.cu), which is what the step latency is made of.
// Only the warps that hold tiles take part (named barrier).
template <int THREADS, int TS>
__device__ __forceinline__ void oc_chain_tiles(const DevOnchip& O, const OcDense& DN, double* __restrict__ V,
                                               double* __restrict__ buf, int* s_flags, int tid, int grp) {
  static_assert(TS == 2 || TS == 4, "tile size");
  constexpr int LTS = TS == 2 ? 1 : 2;
  const int nthr = (int)DN.tile_threads;
  if (tid >= nthr) return;
  const int ntl = (int)(DN.dt1 - DN.dt0);
  const bool has = tid < ntl;
  const uint4 d = has ? __ldg(O.dt + DN.dt0 + tid) : make_uint4(0u, 0u, 0u, 0u);
  const int zoff = (int)(d.x & 0xffffu), diag0 = (int)(d.x >> 16);
  const int T = (int)(d.y & 0xffu), ti = (int)((d.y >> 8) & 0xffu), tj = (int)((d.y >> 16) & 0xffu);
  const int cb = (int)d.z;
  const int cbtot = (int)DN.cbtot;
  double* colb[2] = {buf + cb, buf + cbtot + cb};                 // column j of U (rows > j), buffer j & 1
  double* zrb[2] = {buf + 2 * cbtot + cb, buf + 3 * cbtot + cb};  // row j of Z, buffer j & 1
  double* dinv = buf + 4 * cbtot + cb;
  const int barid = 8 + grp;
  const int i0 = TS * ti, k0 = TS * tj;
  auto ld = [](const double* p, double* v) {   // TS doubles from a 16-byte aligned address
#pragma unroll
    for (int q = 0; q < TS / 2; ++q) {
      const double2 t = *reinterpret_cast<const double2*>(p + 2 * q);
      v[2 * q] = t.x;
      v[2 * q + 1] = t.y;
    }
  };
  auto st = [](double* p, const double* v) {
#pragma unroll
    for (int q = 0; q < TS / 2; ++q) *reinterpret_cast<double2*>(p + 2 * q) = make_double2(v[2 * q], v[2 * q + 1]);
  };
  double s[TS][TS], z[TS][TS];
#pragma unroll
  for (int r = 0; r < TS; ++r)
#pragma unroll
    for (int c = 0; c < TS; ++c) {
      const int i = i0 + r, kk = k0 + c;
      double v = 0.0;
      if (has && i < T && kk <= i) v = V[i == kk ? diag0 + i : zoff + (i * (i - 1)) / 2 + kk];
      s[r][c] = v;
      z[r][c] = (i == kk) ? 1.0 : 0.0;
    }
  if (has && tj == 0) {
    double v[TS];
#pragma unroll
    for (int r = 0; r < TS; ++r) v[r] = s[r][0];
    st(colb[0] + i0, v);
    if (ti == 0) {
#pragma unroll
      for (int c = 0; c < TS; ++c) v[c] = c == 0 ? 1.0 : 0.0;
      st(zrb[0], v);
      const double dd = s[0][0];
      if (!(dd > 0.0)) atomicOr(s_flags, TNO_STATUS_NOT_PD);
      const double rc = oc_rcp(dd);
      dinv[0] = rc;
      V[diag0] = rc;
    }
  }
  asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nthr) : "memory");
  const int nsteps = (int)DN.tmax - 1;
  for (int jb = 0; 4 * jb < nsteps; ++jb) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * jb + jj;
      if (j >= nsteps) break;
      if (has && j < T - 1) {
        if (i0 + TS - 1 > j) {
          const double inv = dinv[j];
          double lr[TS];
          ld(colb[jj & 1] + i0, lr);
#pragma unroll
          for (int r = 0; r < TS; ++r) lr[r] = (i0 + r > j) ? lr[r] * inv : 0.0;
          if (k0 + TS - 1 > j) {
            double uc[TS];
            ld(colb[jj & 1] + k0, uc);
#pragma unroll
            for (int c = 0; c < TS; ++c) uc[c] = (k0 + c > j) ? uc[c] : 0.0;
#pragma unroll
            for (int r = 0; r < TS; ++r)
#pragma unroll
              for (int c = 0; c < TS; ++c) s[r][c] = fma(-lr[r], uc[c], s[r][c]);
          }
          if (k0 <= j) {
            double zc[TS];
            ld(zrb[jj & 1] + k0, zc);
#pragma unroll
            for (int r = 0; r < TS; ++r)
#pragma unroll
              for (int c = 0; c < TS; ++c) z[r][c] = fma(-lr[r], zc[c], z[r][c]);
          }
        }
        // column j + 1 of U and row j + 1 of Z are final now
        const int c1i = (jj + 1) & (TS - 1);       // compile-time after unrolling
        const int tb = (4 * jb + jj + 1) >> LTS;   // tile row / column holding index j + 1
        if (tj == tb) {
          double v[TS];
#pragma unroll
          for (int r = 0; r < TS; ++r) v[r] = s[r][c1i];
          if (ti == tb) {   // the pivot first: its reciprocal is what the other threads wait for
            const double dd = s[c1i][c1i];
            if (!(dd > 0.0)) atomicOr(s_flags, TNO_STATUS_NOT_PD);
            const double rc = oc_rcp(dd);
            dinv[j + 1] = rc;
            V[diag0 + j + 1] = rc;
          }
          st(colb[(jj + 1) & 1] + i0, v);
        }
        if (ti == tb) {
          double v[TS];
#pragma unroll
          for (int c = 0; c < TS; ++c) v[c] = z[c1i][c];
          st(zrb[(jj + 1) & 1] + k0, v);
        }
      }
      asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nthr) : "memory");
    }
  }
  if (has) {
#pragma unroll
    for (int r = 0; r < TS; ++r)
#pragma unroll
      for (int c = 0; c < TS; ++c) {
        const int i = i0 + r, kk = k0 + c;
        if (i < T && kk < i) V[zoff + (i * (i - 1)) / 2 + kk] = z[r][c];
      }
  }
}

// Two columns per exchange step (2 x 2 tiles only).  Block step b eliminates columns j = 2b and j + 1 at once: the owners
// publish BOTH columns of U and BOTH rows of Z as they are before the block is touched, plus the pivot record
//     inv0 = 1 / u_jj ,  m = u_{j+1,j} inv0 (= l_{j+1,j}) ,  inv1 = 1 / (u_{j+1,j+1} - u_{j+1,j} m)
// and every consumer applies the elimination of column j to column j + 1 / row j + 1 itself (two extra DFMA per row and
// column of its tile) before the rank-2 update.  Half as many exchange steps (31 instead of 61 at disc 20) for ~60 %
// more arithmetic per step; the exchange step is what the dense part costs (see the note above oc_chain_tiles).
template <int THREADS>
__device__ __forceinline__ void oc_chain_tiles2(const DevOnchip& O, const OcDense& DN, double* __restrict__ V,
                                                double* __restrict__ buf, int* s_flags, int tid, int grp) {
  const int nthr = (int)DN.tile_threads;
  if (tid >= nthr) return;
  const int ntl = (int)(DN.dt1 - DN.dt0);
  const bool has = tid < ntl;
  const uint4 d = has ? __ldg(O.dt + DN.dt0 + tid) : make_uint4(0u, 0u, 0u, 0u);
  const int zoff = (int)(d.x & 0xffffu), diag0 = (int)(d.x >> 16);
  const int T = (int)(d.y & 0xffu), ti = (int)((d.y >> 8) & 0xffu), tj = (int)((d.y >> 16) & 0xffu);
  const int cb = (int)d.z;
  const int cbtot = (int)DN.cbtot;
  // exchange buffers (double2 per row / column), two generations each, and the pivot records (4 doubles per block)
  double2* const xb = reinterpret_cast<double2*>(buf) + cb;
  auto colb = [&](int gen) { return xb + gen * cbtot; };         // column pair of block b, generation b & 1
  auto zrb = [&](int gen) { return xb + (2 + gen) * cbtot; };    // row pair
  double* piv = buf + 8 * cbtot + 2 * cb;   // block b of this chain at piv + 4 b
  const int barid = 8 + grp;
  const int i0 = 2 * ti, k0 = 2 * tj;
  const int nbT = (T + 1) >> 1;             // blocks of this chain
  double s[2][2], z[2][2];
#pragma unroll
  for (int r = 0; r < 2; ++r)
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const int i = i0 + r, kk = k0 + c;
      double v = (i == kk && i >= T) ? 1.0 : 0.0;   // padding row of an odd chain: identity
      if (has && i < T && kk <= i) v = V[i == kk ? diag0 + i : zoff + (i * (i - 1)) / 2 + kk];
      s[r][c] = v;
      z[r][c] = (i == kk) ? 1.0 : 0.0;
    }
  auto publish = [&](int gen, int b) {   // this thread's share of block b: column pair, row pair, pivot record
    if (tj == b) {
      colb(gen)[i0] = make_double2(s[0][0], s[0][1]);
      colb(gen)[i0 + 1] = make_double2(s[1][0], s[1][1]);
      if (ti == b) {
        // two INDEPENDENT reciprocals: 1 / d0 and 1 / det of the 2 x 2 pivot block; the second pivot of the scalar
        // factorisation is d1 = det / d0, so 1 / d1 = d0 / det without waiting for 1 / d0
        const double d0 = s[0][0];
        const double det = fma(d0, s[1][1], -s[1][0] * s[1][0]);
        if (!(d0 > 0.0) || !(det > 0.0)) atomicOr(s_flags, TNO_STATUS_NOT_PD);
        const double inv0 = oc_rcp(d0);
        const double inv1 = d0 * oc_rcp(det);
        const double m = s[1][0] * inv0;
        *reinterpret_cast<double2*>(piv + 4 * b) = make_double2(inv0, m);
        piv[4 * b + 2] = inv1;
        V[diag0 + 2 * b] = inv0;
        if (2 * b + 1 < T) V[diag0 + 2 * b + 1] = inv1;
      }
    }
    if (ti == b) {
      zrb(gen)[k0] = make_double2(z[0][0], z[1][0]);
      zrb(gen)[k0 + 1] = make_double2(z[0][1], z[1][1]);
    }
  };
  if (has) publish(0, 0);
  asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nthr) : "memory");
  const int nblocks = ((int)DN.tmax + 1) >> 1;
  for (int b = 0; b < nblocks; ++b) {
    const int gen = b & 1;
    if (has && b < nbT) {
      const double2 p01 = *reinterpret_cast<const double2*>(piv + 4 * b);
      const double inv0 = p01.x, m = p01.y;
      if (ti > b) {
        const double inv1 = piv[4 * b + 2];
        double l0[2], l1[2];
#pragma unroll
        for (int r = 0; r < 2; ++r) {
          const double2 a = colb(gen)[i0 + r];
          l0[r] = a.x * inv0;
          l1[r] = fma(-a.x, m, a.y) * inv1;
        }
        if (tj > b) {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const double2 u = colb(gen)[k0 + c];
            const double u1 = fma(-u.x, m, u.y);
#pragma unroll
            for (int r = 0; r < 2; ++r) s[r][c] = fma(-l1[r], u1, fma(-l0[r], u.x, s[r][c]));
          }
        } else {
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const double2 zz = zrb(gen)[k0 + c];
            const double z1 = fma(-m, zz.x, zz.y);
#pragma unroll
            for (int r = 0; r < 2; ++r) z[r][c] = fma(-l1[r], z1, fma(-l0[r], zz.x, z[r][c]));
          }
        }
      } else if (ti == b) {
        // the block's own rows: row 2b of Z is final, row 2b + 1 takes the elimination of column 2b
#pragma unroll
        for (int c = 0; c < 2; ++c) z[1][c] = fma(-m, z[0][c], z[1][c]);
      }
      if (b + 1 < nbT) publish(gen ^ 1, b + 1);
    }
    asm volatile("bar.sync %0, %1;" ::"r"(barid), "r"(nthr) : "memory");
  }
  if (has) {
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int i = i0 + r, kk = k0 + c;
        if (i < T && kk < i) V[zoff + (i * (i - 1)) / 2 + kk] = z[r][c];
      }
  }
}

// DIR: per-candidate horizontal load directions (tno_batch_inputs.load_dir); a separate instantiation so that the
// common single-basis path carries none of its registers or branches.
// G: candidate groups per CTA.  Every group of THREADS threads works on its own candidate with its own value array and
// panel (o_idx_bytes bytes each) and its own named barrier; the index region behind them is shared by all groups, which is
// what lets a third candidate fit on an SM.
// CLK: in-kernel phase clocks (TNO_PHASE_CLOCKS=1); the production instantiations carry none of that code.
template <int THREADS, int MINB, int G, bool DIR, bool CLK>
__global__ void __launch_bounds__(THREADS * G, MINB) k_onchip(const __grid_constant__ DevPlan P, const __grid_constant__ DevOnchip O,
                                                                const __grid_constant__ OcArgs A) {
  extern __shared__ __align__(16) unsigned char smraw[];
  const int grp = G > 1 ? (int)threadIdx.x / THREADS : 0;
  const int bar = G > 1 ? grp + 1 : 0;
  double* sm = reinterpret_cast<double*>(smraw + (size_t)grp * O.o_idx_bytes);
  const unsigned char* idxb = smraw + (size_t)G * O.o_idx_bytes;
  double* xs = sm + O.o_x;
  double* V = sm + O.o_v;
  double* R = sm + O.o_r;
  double* qs = R;  // q lives in the panel region until the matrix is assembled
  double* qsup = sm + O.o_qsup;
  double* react = sm + O.o_react;  // nb x 3
  double* gxy = sm + O.o_gxy;      // Cb^T U B | Cb^T V B (nb x k each) | Cb^T U d0 | Cb^T V d0 (nb each)
  double* azsup = sm + O.o_azsup;  // nsup x nb : dz/dzb at the neighbours of the supports
  double* red = sm + O.o_red;
  double* pers = sm + O.o_pers;    // z_lambdv of the free rows after phase 1 (lambdv variable only)
  double* bf = sm + O.o_bf;        // bestfit / loadpath objectives: per-warp partial column sums (THREADS / 32 x 16)
  double* lpe = sm + O.o_lpe;      // loadpath objective: sign(q_e) l_e^2 per edge (m)
  double* lpa = sm + O.o_lpa;      //                     |q_e| w_e per edge (m), then the direct gradient term (k)
  double* lpw = sm + O.o_lpw;      //                     omega = 2 C^T (|q| o w) per free row, then per support (ni + nb)
  const unsigned short* rowptr16 = reinterpret_cast<const unsigned short*>(idxb + O.idx_rowptr);
  const unsigned short* rowcol16 = reinterpret_cast<const unsigned short*>(idxb + O.idx_rowcol);
  const OcRange* wide = reinterpret_cast<const OcRange*>(idxb + O.idx_wide);
  const OcRange* clev = reinterpret_cast<const OcRange*>(idxb + O.idx_clev);
  const OcChain* chains = reinterpret_cast<const OcChain*>(idxb + O.idx_chains);
  const unsigned char* rowchain = idxb + O.idx_rowchain;
  const OcChunk* chunks = reinterpret_cast<const OcChunk*>(idxb + O.idx_chunks);
  const OcOp* ops = reinterpret_cast<const OcOp*>(idxb + O.idx_ops);
  const OcDense* dense = reinterpret_cast<const OcDense*>(idxb + O.idx_dense);
  const short* vrow_s = reinterpret_cast<const short*>(idxb + O.idx_vrow);
  const short* rowv_s = reinterpret_cast<const short*>(idxb + O.idx_rowv);
  const int* colsrc_s = reinterpret_cast<const int*>(idxb + O.idx_colsrc);
  const int* supptr_s = reinterpret_cast<const int*>(idxb + O.idx_supptr);
  const int* supother_s = reinterpret_cast<const int*>(idxb + O.idx_supother);
  const int* supedge_s = reinterpret_cast<const int*>(idxb + O.idx_supedge);
  const float* supsign_s = reinterpret_cast<const float*>(idxb + O.idx_supsign);
  const int* fixed_s = reinterpret_cast<const int*>(idxb + O.idx_fixed);
  const unsigned short* radjp_s = reinterpret_cast<const unsigned short*>(idxb + O.idx_radjp);
  const double* supxyz_s = reinterpret_cast<const double*>(idxb + O.idx_supxyz);   // x, y, z of the supports
  const uint2* zt_s = reinterpret_cast<const uint2*>(idxb + O.idx_zt);
  const OcWarpFactor wf_s = {reinterpret_cast<const uint32_t*>(idxb + O.idx_wf_off), reinterpret_cast<const uint32_t*>(idxb + O.idx_wf_desc)};
  const unsigned short* rowlist_s = reinterpret_cast<const unsigned short*>(idxb + O.idx_rowlist);
  const OcWarpLocal wl_s = {reinterpret_cast<const unsigned short*>(idxb + O.idx_wl_ptr),
                            reinterpret_cast<const unsigned short*>(idxb + O.idx_wl_rows), idxb + O.idx_wl_s};
  double* zs = sm + O.o_zs;        // z of the free rows (after phase 1) followed by zb of the supports
  const unsigned short* ct_ptr = O.ct_in_smem ? reinterpret_cast<const unsigned short*>(idxb + O.idx_ct_ptr) : O.ct_ptr;
  const unsigned short* ct_col = O.ct_in_smem ? reinterpret_cast<const unsigned short*>(idxb + O.idx_ct_col) : O.ct_col;
  const unsigned short* ct_ent = O.ct_in_smem ? reinterpret_cast<const unsigned short*>(idxb + O.idx_ct_ent) : O.ct_ent;
  const unsigned char* ct_row = O.ct_in_smem ? idxb + O.idx_ct_row : O.ct_row;
  __shared__ int s_flags_all[4];
  int& s_flags = s_flags_all[grp];

  const int tid = G > 1 ? (int)threadIdx.x - grp * THREADS : (int)threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;
  const int n = P.n, m = P.m, ni = P.ni, nb = P.nb, k = P.k, nvar = P.var.nvar, ng = P.con.ng, W = O.W;
  const int nlo = O.nlo;
  const bool has_t = P.var.o_t >= 0;
  const bool want_thrust = (P.objective == TNO_OBJ_MIN || P.objective == TNO_OBJ_MAX);
  const bool want_reac = (P.constraints & TNO_CON_REAC_BOUNDS) != 0;
  const bool want_bestfit = P.objective == TNO_OBJ_BESTFIT;
  const bool want_loadpath = P.objective == TNO_OBJ_LOADPATH;
  const bool want_ecomp = P.objective == TNO_OBJ_ECOMP_LINEAR || P.objective == TNO_OBJ_ECOMP_NONLINEAR;
  const bool want_ecomp_nl = P.objective == TNO_OBJ_ECOMP_NONLINEAR;
  const bool want_wsum = want_bestfit || want_loadpath;  // gradient = weighted column sums of the sensitivity panels
  // the dz/dq_ind panel is only needed by the Jacobian and by the bestfit / loadpath gradients
  const bool need2 = O.ncols2 > 0 && (A.J != nullptr || ((want_wsum || want_ecomp) && A.grad != nullptr));

  // index region: loaded once per CTA, reused by every candidate
  for (int i = threadIdx.x; i < O.idx_bytes / 16; i += THREADS * G)
    reinterpret_cast<uint4*>(smraw + (size_t)G * O.o_idx_bytes)[i] = reinterpret_cast<const uint4*>(O.idx_init)[i];
  __syncthreads();

  const int64_t clk_cand = A.N > 4 * (int64_t)gridDim.x * G ? 4 * (int64_t)gridDim.x * G : 0;   // CTA 0, group 0, its 5th candidate
  for (int64_t cand = (int64_t)blockIdx.x * G + grp; cand < A.N; cand += (int64_t)gridDim.x * G) {
    const bool clk_here = CLK && A.phase_clocks && tid == 0 && cand == clk_cand;
    if (clk_here) A.phase_clocks[0] = clock64();
    // ---- 0. variables ------------------------------------------------------------------------------
    if (tid < nvar) xs[tid] = A.x[cand * A.ldx + tid];
    if (tid == 0) s_flags = 0;
    oc_sync<THREADS>(bar);
    const double lamh = P.var.o_lambdh >= 0 ? xs[P.var.o_lambdh] : 1.0;
    const double lamv = P.var.o_lambdv >= 0 ? xs[P.var.o_lambdv] : 1.0;
    const double pzs = A.pzs ? A.pzs[cand] : 1.0;
    // horizontal load basis of the candidate: (c, s) = load direction; (1, 0) reproduces the single-basis values bit for bit
    const bool has_dir = DIR && A.dir != nullptr;
    const double dirc = has_dir ? A.dir[2 * cand] : 1.0, dirs = has_dir ? A.dir[2 * cand + 1] : 0.0;
    auto mix2 = [&](const double* a, const double* b, int i) -> double {
      if constexpr (DIR) return has_dir ? fma(dirs, __ldg(b + i), dirc * __ldg(a + i)) : __ldg(a + i);
      else return __ldg(a + i);
    };
    double gmin = 1e300;
    double fpart = 0.0;
    bool bad = false;
    double* gout = A.g ? A.g + cand * ng : nullptr;
    double* Jout = A.J ? A.J + cand * (int64_t)ng * nvar : nullptr;

    if (O.dbg & 256) {
      // touch the candidate's output pages long before they are written (address translation in the background)
      const double* pf = nullptr;
      if (tid == 0 && Jout) pf = Jout + (int64_t)(P.con.r_env >= 0 ? P.con.r_env : 0) * nvar;
      if (tid == 32 && Jout) pf = Jout + (int64_t)ng * nvar - 1;
      if (tid == 64 && gout) pf = gout;
      if (tid == 96 && A.z) pf = A.z + cand * n;
      if (tid == 128 && A.grad) pf = A.grad + cand * nvar;
      if (pf) asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
    }
    if (clk_here) A.phase_clocks[1] = clock64();
    // ---- 1. q = B q_ind + lambdh d ; funicular rows of g -------------------------------------------
    for (int e0 = tid; e0 < m; e0 += 4 * THREADS) {
      // four edges per thread at once: 4 x unroll independent loads of B in flight
      constexpr int QU = TNO_QU;
      double acc4[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc4[i] = (e0 + i * THREADS < m) ? lamh * mix2(P.d, P.d_b, e0 + i * THREADS) : 0.0;
#pragma unroll QU
      for (int j = 0; j < k; ++j) {
        const double xj = xs[j];
        const double* bj = O.Bt + (int64_t)j * m + e0;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (e0 + i * THREADS < m) acc4[i] = fma(__ldg(bj + i * THREADS), xj, acc4[i]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
      const int e = e0 + i * THREADS;
      if (e >= m) continue;
      const double acc = acc4[i];
      qs[e] = acc;
      if (has_dir && Jout && P.con.r_fun >= 0) {  // the lambdh column of the funicular rows follows the candidate's direction
        const double de = mix2(P.d, P.d_b, e);
        Jout[(int64_t)(P.con.r_fun + e) * nvar + P.var.o_lambdh] = de;
        Jout[(int64_t)(P.con.r_fun + m + e) * nvar + P.var.o_lambdh] = -de;
      }
      if (want_ecomp_nl) {  // 'simplified' quadratic energy sum stiff q^2 and the weight of its gradient B^T (2 stiff q)
        const double se = __ldg(P.stiff + e);
        fpart = fma(se * acc, acc, fpart);
        lpe[e] = 2.0 * se * acc;
      }
      if (A.q) A.q[cand * m + e] = acc;
      if (P.con.r_fun >= 0) {
        const double g0 = acc - __ldg(P.qmin + e), g1 = __ldg(P.qmax + e) - acc;
        if (gout) {
          gout[P.con.r_fun + e] = g0;
          gout[P.con.r_fun + m + e] = g1;
        }
        gmin = fmin(gmin, fmin(g0, g1));
      }
      bad |= !isfinite(acc);
      }
    }
    oc_sync<THREADS>(bar);

    if (clk_here) A.phase_clocks[2] = clock64();
    // ---- 2. assemble -(Ci^T Q Ci) into the pattern of L; keep q of the support edges ----------------
    if (O.simple_asm) {
      // off-diagonal positions: one edge or fill; diagonals: minus the sum over the incident edges
      // all index loads of a thread (L2) are requested before the first of them is used
      {
        constexpr int AB = TNO_AB;
        for (int p0 = tid; p0 < nlo; p0 += AB * THREADS) {
          unsigned e[AB];
#pragma unroll
          for (int u = 0; u < AB; ++u) e[u] = p0 + u * THREADS < nlo ? __ldg(O.asm_e16 + p0 + u * THREADS) : 0xffffu;
#pragma unroll
          for (int u = 0; u < AB; ++u)
            if (p0 + u * THREADS < nlo) V[p0 + u * THREADS] = e[u] == 0xffffu ? 0.0 : qs[e[u]];
        }
        for (int p0 = tid; p0 < O.nz; p0 += AB * THREADS) {
          unsigned e[AB];
#pragma unroll
          for (int u = 0; u < AB; ++u) e[u] = p0 + u * THREADS < O.nz ? __ldg(O.asm_z16 + p0 + u * THREADS) : 0xffffu;
#pragma unroll
          for (int u = 0; u < AB; ++u)
            if (p0 + u * THREADS < O.nz) V[nlo + ni + p0 + u * THREADS] = e[u] == 0xffffu ? 0.0 : qs[e[u]];
        }
      }
      for (int r = tid; r < ni; r += THREADS) {
        double acc = 0.0;
        const int t0 = O.radjp_in_smem ? (int)radjp_s[r] : __ldg(O.radj_ptr + r);
        const int t1 = O.radjp_in_smem ? (int)radjp_s[r + 1] : __ldg(O.radj_ptr + r + 1);
        unsigned w[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) w[u] = t0 + u < t1 ? __ldg(O.radj + t0 + u) : 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (t0 + u < t1) acc -= qs[w[u] >> 16];
        for (int t = t0 + 8; t < t1; ++t) acc -= qs[__ldg(O.radj + t) >> 16];
        if (r < O.lev0_rows) {   // a leaf of the elimination tree: its pivot is final as assembled
          if (!(acc > 0.0)) atomicOr(&s_flags, TNO_STATUS_NOT_PD);
          acc = oc_rcp(acc);
        }
        V[nlo + r] = acc;
      }
    } else {
      for (int p = tid; p < O.nv; p += THREADS) {
        double acc = 0.0;
        for (int t = O.asm_ptr[p]; t < O.asm_ptr[p + 1]; ++t) acc += (double)O.asm_coef[t] * qs[O.asm_edge[t]];
        if (p >= nlo && p < nlo + O.lev0_rows) {
          if (!(acc > 0.0)) atomicOr(&s_flags, TNO_STATUS_NOT_PD);
          acc = oc_rcp(acc);
        }
        V[p] = acc;
      }
    }
    if (tid == 0) V[O.pos_m1] = -1.0;
    for (int t = tid; t < O.nsup; t += THREADS) qsup[t] = qs[supedge_s[t]];
    oc_sync<THREADS>(bar);  // q consumed: the panel region is free for the factor stream

    if (clk_here) A.phase_clocks[3] = clock64();
    // ---- 3. numeric L D L^T: a short program of staged pull chunks (wide levels, chain S-steps, rectangular
    //         chain rows) and dense chain ops.  Off-diagonal CSR entries stay unscaled (U = L D) until step 4.
    {
      unsigned char* stage_base = reinterpret_cast<unsigned char*>(R);
      double* tmpbuf = R + O.o_tmp_doubles;
      const int seg_bytes = O.fseg_units * 16;
      auto issue = [&](int c) {
        if (c < O.nchunks) {
          const OcChunk K = chunks[c];
          unsigned char* dst = stage_base + ((c + 2) % 3) * seg_bytes;   // chunk 0 lands in the third buffer (see the warp-local part)
          const uint4* src = O.f_stream + K.seg_off;
          for (unsigned u = tid; u < K.seg_units; u += THREADS) cp_async16(dst + 16 * u, src + u);
        }
        cp_async_commit();
      };
      issue(0);
      if (O.wf_on) {
        // Wide levels, warp-local: every warp stages the records of its own subtrees (a few KB, all levels at once) into its
        // slice of the first two ring buffers and factorises them level by level with __syncwarp() only -- the operands of an
        // item are entries of columns of the same subtree, produced by the same warp one level earlier.
        const int w = tid >> 5, ln = tid & 31;
        unsigned char* wbuf = stage_base + (size_t)w * O.wf_buf_bytes;
        const uint4* src = O.wf_stream + wf_s.off[w];
        for (int u = ln; u < (int)wf_s.off[w + 1] - (int)wf_s.off[w]; u += 32) cp_async16(wbuf + 16 * u, src + u);
        cp_async_commit();
        cp_async_wait<0>();
        __syncwarp();
        for (int lev = 1; lev < O.nwide; ++lev) {
          const uint32_t dsc = wf_s.desc[w * O.nwide + lev];
          const int ni_l = (int)((dsc >> 16) & 0xfffu);
          if (ni_l > 0) {
            const uint2* rec = reinterpret_cast<const uint2*>(wbuf + 16 * (dsc & 0xffffu));
            oc_factor_items<32>(rec, reinterpret_cast<const unsigned long long*>(rec + ni_l + 1), ni_l, (int)(dsc >> 28), 0, ln, V, nullptr,
                                &s_flags);
          }
          __syncwarp();
        }
        oc_sync<THREADS>(bar);
      }
      issue(1);
      const bool clk = clk_here;
      long long t_dense = 0, t_wait = 0, t_ldl[6] = {0, 0, 0, 0, 0, 0};
      for (int op = 0; op < O.nops; ++op) {
        const OcOp OP = ops[op];
        if (clk && op < 96) A.phase_clocks[32 + op] = clock64();
        if (OP.type == OC_OP_DENSE) {
          oc_sync<THREADS>(bar);
          const long long t0 = clk ? clock64() : 0;
          if (dense[OP.a].tile_threads > 0 && dense[OP.a].tile_size == 2 && !(O.dbg & 512)) oc_chain_tiles2<THREADS>(O, dense[OP.a], V, R + O.o_tmpz_doubles, &s_flags, tid, grp);
          else if (dense[OP.a].tile_threads > 0 && dense[OP.a].tile_size == 2) oc_chain_tiles<THREADS, 2>(O, dense[OP.a], V, R + O.o_tmpz_doubles, &s_flags, tid, grp);
          else if (dense[OP.a].tile_threads > 0) oc_chain_tiles<THREADS, 4>(O, dense[OP.a], V, R + O.o_tmpz_doubles, &s_flags, tid, grp);
          else oc_chain_dense<THREADS>(O, dense[OP.a], V, R + O.o_tmpz_doubles, &s_flags, tid, bar, clk ? t_ldl : nullptr);
          oc_sync<THREADS>(bar);
          if (clk) t_dense += clock64() - t0;
          continue;
        }
        const int c = OP.a;
        const long long tw0 = clk ? clock64() : 0;
        cp_async_wait<1>();
        oc_sync<THREADS>(bar);
        if (clk) t_wait += clock64() - tw0;
        if (!(O.dbg & 32)) issue(c + 2);
        const OcChunk K = chunks[c];
        const uint2* rec = reinterpret_cast<const uint2*>(stage_base + ((c + 2) % 3) * seg_bytes);
        const unsigned long long* pairs = reinterpret_cast<const unsigned long long*>(rec + K.nitems + 1);
        if (!(O.dbg & 16)) oc_factor_items<THREADS>(rec, pairs, (int)K.nitems, (int)K.lpi_shift, (int)K.kind, tid, V, tmpbuf, &s_flags);
        if (K.kind == 1) {  // rectangular chain rows read each other's S values: write back after everyone has read
          oc_sync<THREADS>(bar);
          for (int it = tid; it < K.nitems; it += THREADS) V[(int)(rec[it].x & 0xffffu)] = tmpbuf[it];
        }
      }
      cp_async_wait<0>();
      oc_sync<THREADS>(bar);
      if (clk) {
        A.phase_clocks[14] = t_dense;
        A.phase_clocks[15] = t_ldl[0] + t_ldl[1] + t_ldl[2];
        for (int q = 0; q < 6; ++q) A.phase_clocks[24 + q] = t_ldl[q];
        A.phase_clocks[30] = t_wait;
      }
    }
    if (clk_here) A.phase_clocks[4] = clock64();
    // ---- 4. L_ij = U_ij / d_j for the CSR entries (the dense chain blocks are already final) ---------
    for (int p = tid; p < nlo; p += THREADS) V[p] *= V[nlo + (int)rowcol16[p]];

    if (clk_here) A.phase_clocks[5] = clock64();
    // ---- 5. phase-1 right-hand sides: z | z_lambdv | Az | x | y -------------------------------------
    // (the scale pass above only touches V; the panel is cleared with coalesced 16-byte stores first)
    for (int it = tid; it < ni * (W >> 1); it += THREADS) reinterpret_cast<double2*>(R)[it] = make_double2(0.0, 0.0);
    oc_sync<THREADS>(bar);
    for (int r = tid; r < ni; r += THREADS) {
      const int v = rowv_s[r];
      // loads (L2) first: the vertex loads and the row's adjacency words travel together
      const int t0 = O.radjp_in_smem ? (int)radjp_s[r] : __ldg(O.radj_ptr + r);
      const int t1 = O.radjp_in_smem ? (int)radjp_s[r + 1] : __ldg(O.radj_ptr + r + 1);
      unsigned w[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) w[u] = t0 + u < t1 ? __ldg(O.radj + t0 + u) : 0u;
      double px, py, pz;
      if (P.var.o_lambdh >= 0) {
        px = lamh * mix2(P.px0, P.px0_b, v);
        py = lamh * mix2(P.py0, P.py0_b, v);
      } else {
        px = __ldg(P.P + 3 * v + 0);
        py = __ldg(P.P + 3 * v + 1);
      }
      const double pzv_v = (P.var.o_lambdv >= 0) ? __ldg(P.pzv + v) : 0.0;
      pz = (P.var.o_lambdv >= 0) ? (lamv * pzv_v + __ldg(P.pz0 + v)) : __ldg(P.P + 3 * v + 2);
      pz *= pzs;
      double bx = -px, by = -py, bz = -pz;
      double* row = R + r * W;
      auto sup_edge_term = [&](unsigned ww) {   // an edge to support b: its q from the saved support-edge values
        const int b = (int)(ww & 0x7fffu), e = (int)(ww >> 16);
        double qe = 0.0;
        for (int tt = supptr_s[b]; tt < supptr_s[b + 1]; ++tt)
          if (supedge_s[tt] == e) qe = qsup[tt];
        const double zb = P.var.o_zb >= 0 ? xs[P.var.o_zb + b] : supxyz_s[3 * b + 2];
        bx = fma(-qe, supxyz_s[3 * b + 0], bx);
        by = fma(-qe, supxyz_s[3 * b + 1], by);
        bz = fma(-qe, zb, bz);
        if (O.c_az >= 0) row[O.c_az + b] -= qe;
      };
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (t0 + u < t1 && (w[u] & 0x8000u)) sup_edge_term(w[u]);
      for (int t = t0 + 8; t < t1; ++t) {
        const unsigned ww = __ldg(O.radj + t);
        if (ww & 0x8000u) sup_edge_term(ww);
      }
      row[O.c_x] = bx;
      row[O.c_y] = by;
      row[O.c_z] = bz;
      if (O.c_lambdv >= 0) row[O.c_lambdv] = -pzv_v * pzs;   // d(pz)/d(lambdv) = pz_scale * pzv
    }
    oc_sync<THREADS>(bar);

    if (clk_here) A.phase_clocks[6] = clock64();
    // ---- 6. solve phase 1 ---------------------------------------------------------------------------
    oc_solve<THREADS>(O, 0, wide, clev, chains, rowchain, rowptr16, rowcol16, V, R, ni, 0, O.ncols1, tid, ct_ptr, ct_col, ct_ent, ct_row, bar, zt_s, rowlist_s, wl_s);

    if (clk_here) A.phase_clocks[7] = clock64();
    // ---- 7. everything that needs the transient columns (x, y, Az) or only z ------------------------
    auto zfix = [&](int b) -> double { return P.var.o_zb >= 0 ? xs[P.var.o_zb + b] : supxyz_s[3 * b + 2]; };
    // persistent phase-1 results leave the panel: phase 2 reuses all of its columns
    for (int r = tid; r < ni; r += THREADS) {
      zs[r] = R[r * W + O.c_z];
      if (O.npers > 1) pers[r] = R[r * W + O.c_lambdv];
    }
    if (tid < nb) zs[ni + tid] = zfix(tid);
    oc_sync<THREADS>(bar);
    auto zval = [&](int v) -> double {  // branch-free: free rows first, supports behind them
      const int r = vrow_s[v];
      return zs[r >= 0 ? r : ni - 1 - r];
    };
    auto xyval = [&](int v, int axis) -> double {
      const int r = vrow_s[v];
      return r >= 0 ? R[r * W + O.c_x + axis] : supxyz_s[3 * (-1 - r) + axis];
    };
    for (int v = tid; v < n; v += THREADS) {
      const double z = zval(v);
      if (A.z) A.z[cand * n + v] = z;
      bad |= !isfinite(z);
      if (want_bestfit) {
        const double dz = z - __ldg(P.s + v);
        fpart = fma(dz, dz, fpart);
      }
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
          ub = __ldg(P.ub + v);
          lb = __ldg(P.lb + v);
          if (DIR && A.bounds) {   // per-candidate envelope (same instantiation as the load directions: off the common path)
            ub = A.bounds[cand * 2 * n + v];
            lb = A.bounds[cand * 2 * n + n + v];
          }
        }
        const double g0 = z - lb, g1 = ub - z;
        if (gout) {
          gout[P.con.r_env + v] = g0;
          gout[P.con.r_env + n + v] = g1;
        }
        gmin = fmin(gmin, fmin(g0, g1));
      }
    }
    if (want_loadpath) {
      // f = sum_e |q_e| l_e^2 ;  grad_q = B^T (sign(q) l^2) + omega^T dz/dq_ind ,  grad_zb = omega^T dz/dzb  with
      // omega = 2 C^T (|q| o w)   (objectives.py:240-275, derivatives.py:640-718; the reference's dx/dq, dy/dq terms
      // vanish on a fixed plan because B spans the null space of the horizontal equilibrium).
      // q left the panel when the matrix was assembled: it is recomputed here, bit for bit as in step 1.
      for (int e = tid; e < m; e += THREADS) {
        double qe = lamh * P.d[e];
#pragma unroll 4
        for (int j = 0; j < k; ++j) qe = fma(__ldg(O.Bt + (int64_t)j * m + e), xs[j], qe);
        const int a = __ldg(P.edge_u + e), b = __ldg(P.edge_v + e);
        const double u = xyval(b, 0) - xyval(a, 0), v = xyval(b, 1) - xyval(a, 1), w = zval(b) - zval(a);
        const double l2 = u * u + v * v + w * w;
        fpart = fma(fabs(qe), l2, fpart);
        lpe[e] = qe > 0.0 ? l2 : (qe < 0.0 ? -l2 : 0.0);
        lpa[e] = fabs(qe) * w;
      }
      oc_sync<THREADS>(bar);
      for (int v = tid; v < n; v += THREADS) {
        double om = 0.0;
        for (int t = __ldg(P.vadj_ptr + v); t < __ldg(P.vadj_ptr + v + 1); ++t)
          om = fma((double)__ldg(P.vadj_sign + t), lpa[__ldg(P.vadj_edge + t)], om);
        const int r = vrow_s[v];
        lpw[r >= 0 ? r : ni - 1 - r] = 2.0 * om;
      }
      oc_sync<THREADS>(bar);
    }
    if ((want_loadpath || want_ecomp_nl) && A.grad) {
      {
        // direct term: lpa[j] = sum_e lpe[e] B[e, j]; half-warps read 16 consecutive doubles of one row of B
        constexpr int NW = THREADS / 32;
        const int csub = lane & 15, rsub = lane >> 4;
        for (int cb = 0; cb < k; cb += 16) {
          const bool cok = cb + csub < k;
          double acc = 0.0;
          if (cok)
            for (int e = warp * 2 + rsub; e < m; e += 2 * NW) acc = fma(lpe[e], __ldg(P.B + (int64_t)e * k + cb + csub), acc);
          acc += __shfl_xor_sync(0xffffffffu, acc, 16);
          if (lane < 16) bf[warp * 16 + lane] = acc;
          oc_sync<THREADS>(bar);
          if (tid < 16 && cb + tid < k) {
            double sum = 0.0;
            for (int w = 0; w < NW; ++w) sum += bf[w * 16 + tid];
            lpa[cb + tid] = sum;
          }
          oc_sync<THREADS>(bar);
        }
      }
    }
    // bestfit / loadpath gradients: dst[c] = addend[c] + sum_r wgt(r) R[r, col0 + c]  (+ the identity rows of the
    // supports).  Half-warps walk 16 consecutive panel columns of one row (conflict-free), rows are spread over the
    // 2 * THREADS / 32 half-warps; fixed reduction order, so results are reproducible run to run.
    auto wgt = [&](int i) -> double {  // i < ni: free row, else support i - ni
      if (want_loadpath) return lpw[i];
      return 2.0 * (zs[i] - __ldg(P.s + (i < ni ? (int)rowv_s[i] : fixed_s[i - ni])));
    };
    auto wcolsum = [&](int col0, int ncols, double* dst, const double* addend, bool add_fixed) {
      constexpr int NW = THREADS / 32;
      const int csub = lane & 15, rsub = lane >> 4;
      for (int cb = 0; cb < ncols; cb += 16) {
        const bool cok = cb + csub < ncols;
        double acc = 0.0;
        for (int r = warp * 2 + rsub; r < ni; r += 2 * NW) {
          if (cok) acc = fma(wgt(r), R[r * W + col0 + cb + csub], acc);
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 16);
        if (lane < 16) bf[warp * 16 + lane] = acc;
        oc_sync<THREADS>(bar);
        if (tid < 16 && cb + tid < ncols) {
          double sum = 0.0;
          for (int w = 0; w < NW; ++w) sum += bf[w * 16 + tid];
          if (add_fixed) sum += wgt(ni + cb + tid);
          if (addend) sum += addend[cb + tid];
          dst[cb + tid] = sum;
        }
        oc_sync<THREADS>(bar);
      }
    };
    if (want_wsum && A.grad && P.var.o_zb >= 0) wcolsum(O.c_az, nb, A.grad + cand * nvar + P.var.o_zb, nullptr, true);
    if (Jout && P.con.r_env >= 0 && P.var.o_zb >= 0) {
      // dz/dzb block straight to the Jacobian (its panel columns are about to be reused)
      for (int idx = tid; idx < n * nb; idx += THREADS) {
        const int v = idx / nb, b = idx - v * nb;
        const int r = vrow_s[v];
        const double a = r >= 0 ? R[r * W + O.c_az + b] : ((-1 - r) == b ? 1.0 : 0.0);
        if (O.dbg & 128) continue;
        Jout[(int64_t)(P.con.r_env + v) * nvar + P.var.o_zb + b] = a;
        Jout[(int64_t)(P.con.r_env + n + v) * nvar + P.var.o_zb + b] = -a;
      }
    }
    if (want_thrust || want_reac || want_ecomp) {
      // R_b = Cb^T Q C X - P_b
      for (int b = tid; b < nb; b += THREADS) {
        const int vb = fixed_s[b];
        const double xb = supxyz_s[3 * b + 0], yb = supxyz_s[3 * b + 1], zb = zfix(b);
        double Rx = 0.0, Ry = 0.0, Rz = 0.0, ud0 = 0.0, vd0 = 0.0;
        for (int t = supptr_s[b]; t < supptr_s[b + 1]; ++t) {
          const int o = supother_s[t];
          const double sg = (double)supsign_s[t];
          const double qe = qsup[t];
          const double xo = xyval(o, 0), yo = xyval(o, 1), zo = zval(o);
          const double u = sg > 0 ? (xb - xo) : (xo - xb);
          const double vv = sg > 0 ? (yb - yo) : (yo - yb);
          const double w = sg > 0 ? (zb - zo) : (zo - zb);
          Rx = fma(sg * qe, u, Rx);
          Ry = fma(sg * qe, vv, Ry);
          Rz = fma(sg * qe, w, Rz);
          if (P.var.o_lambdh >= 0) {
            const double d0 = mix2(P.d, P.d_b, supedge_s[t]);
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
          px = lamh * mix2(P.px0, P.px0_b, vb);
          py = lamh * mix2(P.py0, P.py0_b, vb);
        } else {
          px = __ldg(P.P + 3 * vb + 0);
          py = __ldg(P.P + 3 * vb + 1);
        }
        pz = (P.var.o_lambdv >= 0) ? (lamv * __ldg(P.pzv + vb) + __ldg(P.pz0 + vb)) : __ldg(P.P + 3 * vb + 2);
        pz *= pzs;
        react[3 * b + 0] = Rx - px;
        react[3 * b + 1] = Ry - py;
        react[3 * b + 2] = Rz - pz;
        if (want_thrust) {   // |R_h| and the direction cosines of the thrust objective, once per support
          const double Rhx = Rx - px, Rhy = Ry - py;
          const double Rn = sqrt(Rhx * Rhx + Rhy * Rhy);
          const double osign = (P.objective == TNO_OBJ_MAX) ? -1.0 : 1.0;
          react[3 * nb + 3 * b + 0] = Rn;
          react[3 * nb + 3 * b + 1] = osign * Rhx / Rn;
          react[3 * nb + 3 * b + 2] = osign * Rhy / Rn;
        }
      }
      // (Cb^T U B)[b, j], (Cb^T V B)[b, j]
      for (int it = tid; it < nb * k; it += THREADS) {
        const int b = it / k, j = it - b * k;
        const double xb = supxyz_s[3 * b + 0], yb = supxyz_s[3 * b + 1];
        double gx = 0.0, gy = 0.0;
        const int t0 = supptr_s[b], t1 = supptr_s[b + 1];
        double Bv[4];   // the coefficient loads (L2) of up to four support edges travel together
#pragma unroll
        for (int u = 0; u < 4; ++u) Bv[u] = t0 + u < t1 ? __ldg(P.B + (int64_t)supedge_s[t0 + u] * k + j) : 0.0;
        auto term = [&](int t, double Bej) {
          const int o = supother_s[t];
          const double sg = (double)supsign_s[t];
          const double xo = xyval(o, 0), yo = xyval(o, 1);
          const double u = sg > 0 ? (xb - xo) : (xo - xb);
          const double vv = sg > 0 ? (yb - yo) : (yo - yb);
          gx = fma(sg * u, Bej, gx);
          gy = fma(sg * vv, Bej, gy);
        };
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (t0 + u < t1) term(t0 + u, Bv[u]);
        for (int t = t0 + 4; t < t1; ++t) term(t, __ldg(P.B + (int64_t)supedge_s[t] * k + j));
        gxy[it] = gx;
        gxy[nb * k + it] = gy;
      }
      if ((want_reac || want_ecomp) && O.c_az >= 0) {
        for (int it = tid; it < O.nsup * nb; it += THREADS) {
          const int t = it / nb, b = it - t * nb;
          const int ro = vrow_s[supother_s[t]];
          azsup[it] = ro >= 0 ? R[ro * W + O.c_az + b] : ((-1 - ro) == b ? 1.0 : 0.0);
        }
      }
    }
    oc_sync<THREADS>(bar);

    if (clk_here) A.phase_clocks[8] = clock64();
    // ---- 8-11 run once per PASS over the phase-2 panel columns: when the panel is narrower than the dz/dq_ind block
    //      (W < ncols2, chosen by the host when that lets another candidate fit on the SM), the columns are solved
    //      pw2 at a time and every pass writes its own columns of J and of the gradient.
    const int npass = need2 ? O.npass2 : 1;
    double fval = 0.0;
    for (int pass = 0; pass < npass; ++pass) {
    const int cb = pass * O.pw2;                                    // first phase-2 column of this pass
    const int ncp = need2 ? min(O.pw2, O.ncols2 - cb) : 0;          // columns of this pass
    const bool last_pass = pass == npass - 1;
    // ---- 8. phase-2 right-hand sides: Ci^T W B | Ci^T W d0 ------------------------------------------
    if (need2) {
      const int nc2 = O.ncols2;
      // Row-wise ("pull") accumulation: row r of Ci^T W [B | d0] is  sum over the edges e at vertex r of
      // (z_r - z_other(e)) [B | d0][e, :]  whichever way the edge points (C[e, r] (C z)_e = z_r - z_other).  G lanes
      // cover the column pairs of a row, rows are independent: no colouring, no read-modify-write, ONE barrier.  The
      // coefficient rows are stored in row-adjacency order (Br: one 16-byte-aligned record per (row, edge) pair), so a
      // row's loads are contiguous, independent of one another and issued back to back.
      {
        // (Requesting a row's records one row ahead does not pay here: the loads of both rows end up on one scoreboard,
        // so the wait for the current row also waits for the row just requested, and the extra registers spill.)
        const int gs = O.gshift[1];
        const int c = 2 * (tid & ((1 << gs) - 1));
        const int gc = cb + c;                // phase-2 column of the lane's first panel column
        const bool cok = c < W;
        for (int r = tid >> gs; r < ni; r += THREADS >> gs) {
          const int t0 = O.radjp_in_smem ? (int)radjp_s[r] : __ldg(O.radj_ptr + r);
          const int t1 = O.radjp_in_smem ? (int)radjp_s[r + 1] : __ldg(O.radj_ptr + r + 1);
          const double zr = zs[r];
          double ax = 0.0, ay = 0.0;
          if (cok) {
#pragma unroll 4
            for (int t = t0; t < t1; ++t) {
              const unsigned o = __ldg(O.radj + t) & 0xffffu;
              const double dz = zr - zs[(o & 0x8000u) ? ni + (o & 0x7fffu) : o];
              double2 cf = __ldg(reinterpret_cast<const double2*>(O.Br + (int64_t)t * O.br_stride + gc));
              if constexpr (DIR) {
                // lambdh column with a second load basis: d0 of the candidate's direction (the basis sits one column on)
                if (has_dir && (gc == k || gc + 1 == k)) {
                  const double d0b = __ldg(O.Br + (int64_t)t * O.br_stride + k + 1);
                  if (gc == k) cf.x = fma(dirs, d0b, dirc * cf.x);
                  else cf.y = fma(dirs, d0b, dirc * cf.y);
                }
              }
              ax = fma(dz, cf.x, ax);
              ay = fma(dz, cf.y, ay);
            }
            if (gc >= nc2) ax = 0.0;        // padding columns (and the slot of the second load basis) stay zero
            if (gc + 1 >= nc2) ay = 0.0;
            *reinterpret_cast<double2*>(R + r * W + c) = make_double2(ax, ay);
          }
        }
        oc_sync<THREADS>(bar);
      }
      if (clk_here && pass == 0) A.phase_clocks[9] = clock64();
      // ---- 9. solve phase 2 ----
      {
        const bool clk2 = clk_here && pass == 0;
        long long tcs[7];
        if (clk2)
          for (int q = 0; q < 7; ++q) tcs[q] = 0;
        oc_solve<THREADS>(O, 1, wide, clev, chains, rowchain, rowptr16, rowcol16, V, R, ni, 0, ncp, tid, ct_ptr, ct_col, ct_ent, ct_row,
                          bar, zt_s, rowlist_s, wl_s, clk2 ? tcs : nullptr);
        if (clk2)
          for (int q = 0; q < 7; ++q) A.phase_clocks[16 + q] = tcs[q];
      }
    }

    if (clk_here && pass == 0) A.phase_clocks[10] = clock64();
    // ---- 10. reac_bounds rows, objective, gradient ---------------------------------------------------
    if (want_reac) {
      const bool has_env = (P.constraints & TNO_CON_ENVELOPE) != 0;
      for (int it = tid; it < nb * nvar; it += THREADS) {  // one item per (support b, Jacobian column)
        const int b = it / nvar, col = it - b * nvar;
        const bool ispanel = col < k || col == P.var.o_lambdh;
        const int gcol = col < k ? col : k;                 // phase-2 column of a q_ind / lambdh variable
        // a panel column belongs to the pass that solves it, everything else to the first pass
        if (need2 && ispanel ? (gcol < cb || gcol >= cb + ncp) : pass != 0) continue;
        const int vb = fixed_s[b];
        const double Rx = react[3 * b], Ry = react[3 * b + 1], Rz = react[3 * b + 2];
        const double sx = Rx < 0 ? -1.0 : 1.0, sy = Ry < 0 ? -1.0 : 1.0, sz = Rz < 0 ? -1.0 : 1.0;
        const double rz2 = Rz * Rz;
        const double zb = zfix(b);
        double vx = 0.0, vy = 0.0;
        const int t0 = supptr_s[b], t1 = supptr_s[b + 1];
        if (ispanel) {
          const bool isq = col < k;
          const int pcol = gcol - cb;
          double dRx = isq ? gxy[b * k + col] : 0.0, dRy = isq ? gxy[nb * k + b * k + col] : 0.0, dRz = 0.0;
          for (int t = t0; t < t1; ++t) {
            const int o = supother_s[t];
            const int e = supedge_s[t];
            const double sg = (double)supsign_s[t];
            const double qe = qsup[t];
            const double zo = zval(o);
            const double w = sg > 0 ? (zb - zo) : (zo - zb);
            const double coef = isq ? P.B[(int64_t)e * k + col] : mix2(P.d, P.d_b, e);
            dRz = fma(sg * w, coef, dRz);
            const int ro = vrow_s[o];
            if (need2 && (has_env || !isq) && ro >= 0) {
              const double dzo = R[ro * W + pcol];
              dRz = fma(sg * qe, sg > 0 ? -dzo : dzo, dRz);
            }
          }
          if (!isq) {  // lambdh column: Cb^T U d0 - px0, Cb^T V d0 - py0 (formed in step 7 while x, y were live)
            dRx = gxy[2 * nb * k + b] - mix2(P.px0, P.px0_b, vb);
            dRy = gxy[2 * nb * k + nb + b] - mix2(P.py0, P.py0_b, vb);
          }
          vx = zb * sx * (-Rz * dRx + Rx * dRz) / rz2 / sz;
          vy = zb * sy * (-Rz * dRy + Ry * dRz) / rz2 / sz;
        } else if (P.var.o_zb >= 0 && col >= P.var.o_zb && col < P.var.o_zb + nb) {
          // dRz/dzb = Cb^T Q C A ; the reference adds COLUMN b of it to row b (jacobian.py:222)
          const int cbz = col - P.var.o_zb;
          double acc = 0.0;
          for (int t = supptr_s[cbz]; t < supptr_s[cbz + 1]; ++t) {
            const double sg = (double)supsign_s[t];
            const double Ao = azsup[t * nb + b];
            const double Ac = (cbz == b) ? 1.0 : 0.0;
            acc = fma(sg * qsup[t], sg > 0 ? (Ac - Ao) : (Ao - Ac), acc);
          }
          vx = sz * zb * fabs(Rx) / rz2 * acc;
          vy = sz * zb * fabs(Ry) / rz2 * acc;
          if (cbz == b) {
            vx += -fabs(Rx / Rz);
            vy += -fabs(Ry / Rz);
          }
        } else if (col == P.var.o_lambdv) {
          double dRz = 0.0;
          for (int t = t0; t < t1; ++t) {
            const double sg = (double)supsign_s[t];
            const int ro = vrow_s[supother_s[t]];
            if (ro >= 0) {
              const double dzo = pers[ro];
              dRz = fma(sg * qsup[t], sg > 0 ? -dzo : dzo, dRz);
            }
          }
          dRz -= P.pzv[vb] * pzs;
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
    if (want_thrust) {
      if (pass == 0) {
        const double osign = (P.objective == TNO_OBJ_MAX) ? -1.0 : 1.0;
        if (tid < nvar && A.grad) {
          double gsum = 0.0;
          if (tid < k)
            for (int b = 0; b < nb; ++b)
              gsum += react[3 * nb + 3 * b + 1] * gxy[b * k + tid] + react[3 * nb + 3 * b + 2] * gxy[nb * k + b * k + tid];
          A.grad[cand * nvar + tid] = gsum;
        }
        if (tid == 0) {
          for (int b = 0; b < nb; ++b) fval += react[3 * nb + 3 * b];
          fval *= osign;
        }
      }
    } else if (want_ecomp) {
      // f = -sum(R o dXb) [+ sum stiff q^2] ;  grad = -sum_b dXb[b] . dR_b/dx [+ B^T (2 stiff q)]
      // (objectives.py:278-340, derivatives.py:503-620; dx/dq_ind and dy/dq_ind vanish on a fixed plan)
      if (A.grad)
        for (int col = tid; col < nvar; col += THREADS) {
          if (col < k ? (col < cb || col >= cb + ncp) : pass != 0) continue;
          double gsum = 0.0;
          if (col < k) {
            for (int b = 0; b < nb; ++b) {
              const double zb = zfix(b);
              double dRz = 0.0;
              for (int t = supptr_s[b]; t < supptr_s[b + 1]; ++t) {
                const int o = supother_s[t];
                const int e = supedge_s[t];
                dRz = fma(zb - zval(o), __ldg(P.B + (int64_t)e * k + col), dRz);  // C[e, b] (C z)_e = z_b - z_o
                const int ro = vrow_s[o];
                if (ro >= 0) dRz = fma(-qsup[t], R[ro * W + col - cb], dRz);
              }
              gsum -= __ldg(P.dXb + 3 * b) * gxy[b * k + col] + __ldg(P.dXb + 3 * b + 1) * gxy[nb * k + b * k + col] +
                      __ldg(P.dXb + 3 * b + 2) * dRz;
            }
            if (want_ecomp_nl) gsum += lpa[col];
          } else {
            const int cbz = col - P.var.o_zb;
            for (int b = 0; b < nb; ++b) {
              double acc = 0.0;
              for (int t = supptr_s[b]; t < supptr_s[b + 1]; ++t) acc = fma(qsup[t], (cbz == b ? 1.0 : 0.0) - azsup[t * nb + cbz], acc);
              gsum -= __ldg(P.dXb + 3 * b + 2) * acc;
            }
          }
          A.grad[cand * nvar + col] = gsum;
        }
      if (last_pass) {
        for (int off = 16; off > 0; off >>= 1) fpart += __shfl_xor_sync(0xffffffffu, fpart, off);
        if (lane == 0) red[warp] = fpart;
        oc_sync<THREADS>(bar);
        if (tid == 0) {
          for (int w = 0; w < THREADS / 32; ++w) fval += red[w];
          for (int b = 0; b < nb; ++b)
            fval -= react[3 * b] * P.dXb[3 * b] + react[3 * b + 1] * P.dXb[3 * b + 1] + react[3 * b + 2] * P.dXb[3 * b + 2];
        }
        oc_sync<THREADS>(bar);  // red is reused by the feasibility reduction below
      }
    } else if (want_wsum) {
      {
        const int nq = min(cb + ncp, k) - cb;   // q_ind columns of this pass (the lambdh column is not part of these gradients)
        if (A.grad && nq > 0) wcolsum(0, nq, A.grad + cand * nvar + cb, want_loadpath ? lpa + cb : nullptr, false);
      }
      if (last_pass) {
        for (int off = 16; off > 0; off >>= 1) fpart += __shfl_xor_sync(0xffffffffu, fpart, off);
        if (lane == 0) red[warp] = fpart;
        oc_sync<THREADS>(bar);
        if (tid == 0)
          for (int w = 0; w < THREADS / 32; ++w) fval += red[w];
        oc_sync<THREADS>(bar);  // red is reused by the feasibility reduction below
      }
    } else if (pass == 0) {
      if (tid < nvar && A.grad)
        A.grad[cand * nvar + tid] = (tid == nvar - 1) ? (P.objective == TNO_OBJ_T ? 1.0 : (P.objective == TNO_OBJ_NEG_LAST ? -1.0 : 0.0)) : 0.0;
      if (tid == 0)
        fval = P.objective == TNO_OBJ_T ? xs[nvar - 1]
               : (P.objective == TNO_OBJ_NEG_LAST ? -xs[nvar - 1] : (P.objective == TNO_OBJ_FEASIBILITY ? 1.0 : 0.0));
    }
    if (last_pass && tid == 0) {
      bad |= !isfinite(fval);
      if (A.f) A.f[cand] = fval;
    }

    if (clk_here && pass == 0) A.phase_clocks[11] = clock64();
    // ---- 11. dense Jacobian: envelope block from the panel (the funicular block is k_fill_funicular's) ------
    if (Jout && P.con.r_env >= 0) {
      double* o1 = Jout + (int64_t)P.con.r_env * nvar;
      double* o2 = o1 + (int64_t)n * nvar;
      if (ncp > 0) {
        // the lanes of a row walk the pass's panel columns (contiguous in the panel and in J), rows over the rest
        const int es = O.emit_shift;
        const int c = tid & ((1 << es) - 1);
        if (c < ncp) {
          const int gcol = cb + c;
          const int col = gcol < k ? gcol : P.var.o_lambdh;
          for (int v = tid >> es; v < n; v += THREADS >> es) {
            const int r = vrow_s[v];
            const double a = r >= 0 ? R[r * W + c] : 0.0;
            if (O.dbg & 64) continue;
            o1[(int64_t)v * nvar + col] = a;
            o2[(int64_t)v * nvar + col] = -a;
          }
        }
      }
      if (pass == 0) {
        // columns that do not come from the phase-2 panel: lambdv (persistent copy), and zeros when there is no phase 2
        // (zb and thickness columns were written in step 7)
        for (int col = 0; col < nvar; ++col) {
          const int src = colsrc_s[col];
          if (src >= 0 || src == -1) continue;
          for (int v = tid; v < n; v += THREADS) {
            const int r = vrow_s[v];
            const double a = (r >= 0 && src == -3) ? pers[r] : 0.0;
            o1[(int64_t)v * nvar + col] = a;
            o2[(int64_t)v * nvar + col] = -a;
          }
        }
      }
    }
    if (!last_pass) oc_sync<THREADS>(bar);   // the panel is rewritten by the next pass
    }  // pass

    if (clk_here) A.phase_clocks[12] = clock64();
    // ---- 12. feasibility margin and status -----------------------------------------------------------
    for (int off = 16; off > 0; off >>= 1) gmin = fmin(gmin, __shfl_xor_sync(0xffffffffu, gmin, off));
    const unsigned anybad = __ballot_sync(0xffffffffu, bad);
    if (lane == 0) {
      red[warp] = gmin;
      if (anybad) atomicOr(&s_flags, TNO_STATUS_NONFINITE);
    }
    oc_sync<THREADS>(bar);
    if (tid == 0) {
      double g = red[0];
      for (int w = 1; w < THREADS / 32; ++w) g = fmin(g, red[w]);
      if (A.gmin) A.gmin[cand] = g;
      if (A.status) A.status[cand] = s_flags;
    }
    oc_sync<THREADS>(bar);  // panel, flags and small arrays are reused by the next candidate
    if (clk_here) A.phase_clocks[13] = clock64();
  }
}

// The funicular rows of J ([B 0; -B 0] plus the d0 column) do not depend on the candidate: a pure streaming
// kernel replicates the precomputed block.  Each thread keeps one 16-byte chunk in registers and writes it to
// CPB candidates, so the block is read once per CPB candidates (L2) and the stores run at HBM speed.
constexpr int FILL_CPB = 16;
__global__ void __launch_bounds__(256) k_fill_funicular(const double2* __restrict__ src, int64_t chunks, double* __restrict__ J,
                                                        int64_t cand_stride, int64_t row_off, int64_t N) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= chunks) return;
  const double2 v = src[c];
  const int64_t c0 = (int64_t)blockIdx.y * FILL_CPB;
  const int64_t c1 = c0 + FILL_CPB < N ? c0 + FILL_CPB : N;
  for (int64_t cand = c0; cand < c1; ++cand)
    reinterpret_cast<double2*>(J + cand * cand_stride + row_off)[c] = v;
}

// ------------------------------------------------------------------------------------------------------
// host: schedule construction
struct OcLayout {
  int threads;     // threads per candidate group
  int groups;      // candidate groups per CTA (they share the index region)
  int multipass;   // 1: panel as wide as phase 1 needs, the phase-2 columns solved in several passes
};
static int build_onchip_impl(Plan* PL, const tno_problem_desc* D, const OcLayout LY, bool dry, int* cands_per_sm, int* npass_out);

// Chooses the shared-memory layout that keeps the most candidates resident per SM (the kernel is bound by the latency of
// a candidate's dependent steps, so throughput follows the number of candidates in flight), then builds that schedule.
// Layouts tried: the full-width panel or a panel as wide as phase 1 needs with the phase-2 columns solved in passes;
// one or two candidate groups per CTA sharing one index region (a third group would have to live with 80 registers per
// thread and gained nothing when measured); 512-thread groups when only one candidate fits anyway.
int build_onchip(Plan* PL, const tno_problem_desc* D) {
  std::vector<OcLayout> tries;
  if (const char* env = getenv("TNO_ONCHIP_LAYOUT")) {   // developer knob: "threads,groups,multipass"
    OcLayout LY{256, 1, 0};
    if (sscanf(env, "%d,%d,%d", &LY.threads, &LY.groups, &LY.multipass) != 3 || (LY.threads != 256 && LY.threads != 512) ||
        LY.groups < 1 || LY.groups > 2 || (LY.threads == 512 && LY.groups != 1)) {
      set_error("TNO_ONCHIP_LAYOUT must be threads(256|512),groups(1-2),multipass(0|1)");
      return TNO_EINVAL;
    }
    tries.push_back(LY);
  } else {
    for (int mp = 0; mp < 2; ++mp)
      tries.push_back(OcLayout{256, 1, mp});   // two groups per CTA (TNO_ONCHIP_LAYOUT=256,2,x) measured 4 % slower than two CTAs
  }
  int best = -1, best_cands = 0, best_pass = 0;
  for (size_t i = 0; i < tries.size(); ++i) {
    int cands = 0, npass = 1;
    const int rc = build_onchip_impl(PL, D, tries[i], true, &cands, &npass);
    if (rc != TNO_OK) return rc;
    if (!PL->onchip_ok) continue;
    if (tries[i].multipass && npass == 1) continue;   // same as the single-pass layout
    // more candidates in flight first; then fewer passes; then fewer groups (more registers per thread)
    const bool better = best < 0 || cands > best_cands || (cands == best_cands && npass < best_pass);
    if (better) {
      best = (int)i;
      best_cands = cands;
      best_pass = npass;
    }
  }
  PL->onchip_ok = false;
  if (best < 0) return TNO_OK;
  OcLayout LY = tries[best];
  if (best_cands == 1 && LY.threads == 256 && !getenv("TNO_ONCHIP_256") && !getenv("TNO_ONCHIP_LAYOUT")) LY.threads = 512;
  int cands = 0, npass = 1;
  int rc = build_onchip_impl(PL, D, LY, false, &cands, &npass);
  if (rc == TNO_OK && !PL->onchip_ok && LY.threads == 512) {   // should not happen (same memory footprint)
    LY.threads = 256;
    rc = build_onchip_impl(PL, D, LY, false, &cands, &npass);
  }
  return rc;
}

void launch_fill_funicular(Plan* PL, const double* block, int64_t count, double* J, int64_t N, cudaStream_t stream) {
  const int64_t chunks = count / 2;  // count = 2 m nvar is even
  const int64_t cand_stride = (int64_t)PL->dp.con.ng * PL->dp.var.nvar;
  const int64_t per_launch = (int64_t)65535 * FILL_CPB;   // grid.y limit: about one million candidates per launch
  PL->prof_begin(KK_EMIT_J, stream);
  for (int64_t c0 = 0; c0 < N; c0 += per_launch) {
    const int64_t nc = std::min(per_launch, N - c0);
    dim3 g((unsigned)((chunks + 255) / 256), (unsigned)((nc + FILL_CPB - 1) / FILL_CPB), 1);
    k_fill_funicular<<<g, 256, 0, stream>>>(reinterpret_cast<const double2*>(block), chunks, J + c0 * cand_stride, cand_stride,
                                            (int64_t)PL->dp.con.r_fun * PL->dp.var.nvar, nc);
    PL->launches += 1;
  }
  PL->prof_end(stream);
}

static int build_onchip_impl(Plan* PL, const tno_problem_desc* D, const OcLayout LY, bool dry, int* cands_per_sm, int* npass_out) {
  const int THREADS = LY.threads;
  const Symbolic& S = PL->sym;
  DevPlan& P = PL->dp;
  DevOnchip& O = PL->oc;
  PL->onchip_ok = false;
  const int ni = P.ni, m = P.m, nb = P.nb, k = P.k, nvar = P.var.nvar;
  const int nnz = P.nnz_l;
  if (P.var.o_xyb >= 0 || P.con.r_envxy >= 0) { set_error("on-chip family not applicable: xyb / envelopexy run on the interleaved family"); return TNO_OK; }
  if (nnz >= 60000 || ni >= 32768 || S.nlevels < 1 || m >= 65535) { set_error("on-chip family not applicable: nnz >= 60000 || ni >= 32768 || S.nlevels < 1 || m >= 65535"); return TNO_OK; }

  // ---- panel columns.  phase 1: [z | z_lambdv | Az (nb) | x | y]   phase 2: [Dq (k) | z_lambdh] -------------
  int pos = 0;
  O.c_z = pos++;
  O.c_lambdv = -1;
  O.c_q = -1;
  O.c_lambdh = -1;
  if (P.var.o_lambdv >= 0) O.c_lambdv = pos++;
  O.npers = pos;
  O.c_az = (P.var.o_zb >= 0) ? pos : -1;
  O.c_x = pos + (P.var.o_zb >= 0 ? nb : 0);
  O.c_y = O.c_x + 1;
  O.ncols1 = O.c_y + 1;
  O.ncols2 = 0;
  if (P.rhs.c_q >= 0) {
    O.c_q = 0;
    O.ncols2 = k + (P.var.o_lambdh >= 0 ? 1 : 0);
    if (P.var.o_lambdh >= 0) O.c_lambdh = k;
  }
  O.W = (std::max(O.ncols1, O.ncols2) + 1) / 2 * 2;
  O.npass2 = 1;
  if (LY.multipass && O.ncols2 > 0) {
    // panel as wide as phase 1 needs (at least 8 columns); the phase-2 columns are solved W at a time
    O.W = std::max(8, (O.ncols1 + 1) / 2 * 2);
    O.npass2 = (O.ncols2 + O.W - 1) / O.W;
  }
  O.pw2 = O.W;
  {
    int es = 0;
    while ((1 << es) < std::max(1, std::min(O.pw2, O.ncols2))) ++es;
    if (es > 8) { set_error("on-chip family not applicable: more than 256 panel columns"); return TNO_OK; }
    O.emit_shift = es;
  }
  *npass_out = O.npass2;
  O.nvar_magic = (uint32_t)((((uint64_t)1 << 32) + nvar - 1) / nvar);
  O.rhs2_shift = 0;
  for (int ph = 0; ph < 2; ++ph) {
    const int ncols = ph == 0 ? O.ncols1 : std::min(O.ncols2, O.pw2);
    const int lanes = std::max(1, (ncols + 1) / 2);
    int g = 0;
    while ((1 << g) < lanes) ++g;
    if (g > 5) { set_error("on-chip family not applicable: g > 5"); return TNO_OK; }  // more than 64 columns per phase: not handled on chip
    O.gshift[ph] = g;
    if (ph == 1) {
      int g2 = 0;
      while ((1 << g2) < std::min(std::max(ncols, 1), 32)) ++g2;
      O.rhs2_shift = g2;
    }
  }

  // ---- chains -------------------------------------------------------------------------------------------
  const int nch = (int)S.chain_clevel.size();
  const int ncl = (int)S.clevel_ptr.size() - 1;
  const int wide_rows = S.wide_rows;
  const int nwide = nch > 0 ? S.wide_levels : S.nlevels;
  std::vector<int> chain_of(ni, -1), zoff_rel(std::max(nch, 1), 0);
  int nz = 0;
  for (int c = 0; c < nch; ++c) {
    const int T = S.chain_ptr[c + 1] - S.chain_ptr[c];
    if (T > 64) { set_error("on-chip family not applicable: T > 64"); return TNO_OK; }
    for (int r = S.chain_ptr[c]; r < S.chain_ptr[c + 1]; ++r) chain_of[r] = c;
    zoff_rel[c] = nz;
    nz += T * (T - 1) / 2;
  }
  // outside CSR
  std::vector<int> rowptr_o(ni + 1, 0), rowcol_o, csr_of_t(S.rowcol.size(), -1);
  for (int i = 0; i < ni; ++i) {
    for (int t = S.rowptr[i]; t < S.rowptr[i + 1]; ++t) {
      const int j = S.rowcol[t];
      if (chain_of[i] >= 0 && chain_of[i] == chain_of[j]) continue;
      csr_of_t[t] = (int)rowcol_o.size();
      rowcol_o.push_back(j);
    }
    rowptr_o[i + 1] = (int)rowcol_o.size();
  }
  const int nlo = (int)rowcol_o.size();
  const int nv = nlo + ni + nz + 1;
  if (nv >= 65535) { set_error("on-chip family not applicable: nv >= 65535"); return TNO_OK; }
  O.nlo = nlo;
  O.nz = nz;
  O.nv = nv - 1;
  O.pos_m1 = nv - 1;
  O.wide_rows = wide_rows;
  O.nwide = nwide;
  O.nclev = nch > 0 ? ncl : 0;
  O.nchains = nch;
  auto zpos = [&](int i, int j) {  // i > j, same chain
    const int c = chain_of[i];
    const int il = i - S.chain_ptr[c], jl = j - S.chain_ptr[c];
    return nlo + ni + zoff_rel[c] + il * (il - 1) / 2 + jl;
  };
  // CSC position -> V position
  std::vector<int> newpos(nnz, -1);
  for (int j = 0; j < ni; ++j) newpos[S.colptr[j]] = nlo + j;
  for (size_t t = 0; t < S.rowpos.size(); ++t) {
    int i = 0;  // row of CSR entry t
    (void)i;
  }
  for (int i = 0; i < ni; ++i)
    for (int t = S.rowptr[i]; t < S.rowptr[i + 1]; ++t)
      newpos[S.rowpos[t]] = csr_of_t[t] >= 0 ? csr_of_t[t] : zpos(i, S.rowcol[t]);

  // ---- partition of the wide forest: whole subtrees dealt to the warps, heaviest first to the lightest warp.  Used by the
  //      warp-local wide part of the solves and of the factorisation.
  const int NW = THREADS / 32;
  std::vector<int> wl_root(ni), wl_warp_of(ni, 0);
  bool wl_balanced = false;
  {
    std::iota(wl_root.begin(), wl_root.end(), 0);
    for (int j = wide_rows - 1; j >= 0; --j)
      if (S.parent[j] >= 0 && S.parent[j] < wide_rows) wl_root[j] = wl_root[S.parent[j]];
    std::vector<long long> weight(ni, 0);
    for (int j = 0; j < wide_rows; ++j) weight[wl_root[j]] += 1 + (rowptr_o[j + 1] - rowptr_o[j]);
    std::vector<int> roots;
    for (int j = 0; j < wide_rows; ++j)
      if (wl_root[j] == j) roots.push_back(j);
    std::stable_sort(roots.begin(), roots.end(), [&](int a, int b) { return weight[a] > weight[b]; });
    std::vector<long long> load(NW, 0);
    long long total = 0;
    for (int r : roots) {
      const int w = (int)(std::min_element(load.begin(), load.end()) - load.begin());
      load[w] += weight[r];
      wl_warp_of[r] = w;
      total += weight[r];
    }
    const long long maxload = *std::max_element(load.begin(), load.end());
    // worth it when the heaviest warp carries at most twice its share (one subtree = the whole tree would not be)
    wl_balanced = nch > 0 && nwide > 1 && total > 0 && maxload * NW <= 2 * total;
  }

  // ---- support incidence ----
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
  O.dbg = getenv("TNO_DBG") ? atoi(getenv("TNO_DBG")) : 0;

  // ---- factorisation program ----------------------------------------------------------------------------
  struct Item {
    uint32_t pf;                          // V position | diagonal flag
    std::vector<unsigned long long> prs;  // a | b << 16 | dp << 32
  };
  std::vector<OcChunk> chunks;
  std::vector<OcOp> ops;
  std::vector<OcDense> dense_r;
  std::vector<uint4> de;
  std::vector<uint32_t> dc, db;
  std::vector<uint4> dt;
  int max_tile_buf = 0;
  std::vector<uint32_t> stream;
  uint32_t max_units = 0;
  int max_tmp_items = 0;
  // three staging buffers, the temp buffer of the rectangular chain rows and the inverse's nz doubles share the panel
  // region while the matrix is factorised: the chunk size follows the panel so that they do not enlarge it
  size_t CHUNK_BYTES = 12 * 1024;
  // dense scratch behind the staging buffers: the element-wise variant inverts into nz doubles, the register-tile variant
  // only exchanges columns / rows (5 x the padded chain rows of a level)
  int dense_scratch = 0;
  {
    static const bool elementwise = getenv("TNO_ONCHIP_DENSE_ELEMENTS") != nullptr;
    for (int cl = 0; cl < O.nclev; ++cl) {
      const int c0 = S.clevel_ptr[cl], c1 = S.clevel_ptr[cl + 1];
      int need = nz;
      for (int TS = getenv("TNO_ONCHIP_TILE4") ? 4 : 2; TS <= 4 && !elementwise; TS += 2) {
        int tiles = 0, cbo = 0;
        for (int c = c0; c < c1; ++c) {
          const int nt = (S.chain_ptr[c + 1] - S.chain_ptr[c] + TS - 1) / TS;
          tiles += nt * (nt + 1) / 2;
          cbo += TS * nt;
        }
        if (tiles <= THREADS) {
          need = 10 * cbo;   // two-column steps of the 2 x 2 tiles: (column pair, row pair) x 2 buffers + pivot records
          break;
        }
      }
      dense_scratch = std::max(dense_scratch, need);
    }
    static const char* cap_env = getenv("TNO_ONCHIP_CHUNK_KB");   // developer knob
    const long long cap = (cap_env ? atoi(cap_env) : 20) * 1024LL;
    const long long room = ((long long)std::max(ni * O.W, m) - dense_scratch - 2) * 8 - 4096;
    CHUNK_BYTES = (size_t)std::min<long long>(cap, std::max<long long>(2048, room / 3 / 16 * 16));
  }
  static const bool sort_items = getenv("TNO_ONCHIP_NO_ITEM_SORT") == nullptr;   // developer knob (A/B measurements)
  auto emit_step = [&](std::vector<Item> items, int kind) {
    // The items of a kind-0 step are independent of one another: longest dot products first, so that the items of one chunk
    // have similar lengths (a chunk takes as long as its longest item; the lanes-per-item split is chosen per chunk).
    // Kind-1 items keep their order (a later chunk must not overwrite what an earlier one still reads).
    if (kind == 0 && sort_items)
      std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.prs.size() > b.prs.size(); });
    size_t i0 = 0;
    while (i0 < items.size()) {
      size_t i1 = i0, bytes = 8;
      while (i1 < items.size()) {
        const size_t add = 8 + 8 * items[i1].prs.size();
        if (i1 > i0 && (bytes + add > CHUNK_BYTES || (kind == 1 && i1 - i0 >= 448))) break;   // kind 1: the temp buffer holds 4 KB
        bytes += add;
        ++i1;
      }
      OcChunk K;
      std::memset(&K, 0, sizeof(K));
      K.seg_off = (uint32_t)(stream.size() / 4);
      K.nitems = (unsigned short)(i1 - i0);
      K.kind = (unsigned char)kind;
      size_t npairs = 0;
      for (size_t i = i0; i < i1; ++i) {
        stream.push_back(items[i].pf);
        stream.push_back((uint32_t)npairs);
        npairs += items[i].prs.size();
      }
      stream.push_back(0);
      stream.push_back((uint32_t)npairs);
      for (size_t i = i0; i < i1; ++i)
        for (unsigned long long pr : items[i].prs) {
          stream.push_back((uint32_t)(pr & 0xffffffffu));
          stream.push_back((uint32_t)(pr >> 32));
        }
      while (stream.size() % 4) stream.push_back(0);
      K.seg_units = (uint32_t)(stream.size() / 4) - K.seg_off;
      const int nitems = (int)(i1 - i0);
      // lanes per item: minimise  passes x (pairs of the longest item per lane + shuffle steps + a per-pass overhead),
      // all in units of one dependent pair iteration
      int lpi_shift = 0;
      {
        size_t maxp = 0;
        for (size_t i = i0; i < i1; ++i) maxp = std::max(maxp, items[i].prs.size());
        double best = 1e300;
        for (int sft = 0; sft <= 5; ++sft) {
          const int passes = ((nitems << sft) + THREADS - 1) / THREADS;
          const double cost = passes * ((double)((maxp + (1u << sft) - 1) >> sft) + 0.8 * sft + 2.0);
          if (cost < best - 1e-9) {
            best = cost;
            lpi_shift = sft;
          }
        }
        static const bool old_rule = getenv("TNO_ONCHIP_OLD_LPI") != nullptr;   // developer knob (A/B measurements)
        if (old_rule) {
          lpi_shift = 0;
          if (nitems > 0 && npairs / nitems >= 4)
            while (lpi_shift < 5 && (nitems << (lpi_shift + 1)) <= THREADS) ++lpi_shift;
        }
      }
      K.lpi_shift = (unsigned char)lpi_shift;
      max_units = std::max(max_units, K.seg_units);
      if (kind == 1) max_tmp_items = std::max(max_tmp_items, nitems);
      OcOp OP;
      std::memset(&OP, 0, sizeof(OP));
      OP.type = OC_OP_CHUNK;
      OP.a = (unsigned short)chunks.size();
      ops.push_back(OP);
      chunks.push_back(K);
      i0 = i1;
    }
  };
  auto make_item = [&](int p, bool diag_flag, int col_limit) {
    // pairs over common columns c < col_limit (all common columns when col_limit < 0)
    Item it;
    it.pf = (uint32_t)newpos[p] | (diag_flag ? 0x80000000u : 0u);
    for (int64_t t = S.dot_ptr[p]; t < S.dot_ptr[p + 1]; ++t) {
      if (col_limit >= 0 && S.dot_c[t] >= col_limit) continue;
      it.prs.push_back((unsigned long long)newpos[S.dot_a[t]] | ((unsigned long long)newpos[S.dot_b[t]] << 16) |
                       ((unsigned long long)(nlo + S.dot_c[t]) << 32));
    }
    return it;
  };
  // The leaves (level 0) have no updates at all: their off-diagonal entries are final as assembled and the assembly
  // step inverts their pivots, so the factor program starts at level 1.
  O.lev0_rows = S.level_ptr[1];
  for (int lev = 1; lev < nwide; ++lev)
    for (int j = S.level_ptr[lev]; j < S.level_ptr[lev + 1]; ++j)
      if (S.level_cols[j] != j) {
        set_error("internal: level ordering is not contiguous");
        return TNO_EINVAL;
      }
  // Wide levels 1 ..: warp-local when the forest partition is balanced (every warp gets the items of the columns of its own
  // subtrees as one small record stream: [records | pairs] per level, the layout of a staged chunk), else one staged step per
  // level for the whole CTA.
  std::vector<uint32_t> wf_words, wf_off(NW + 1, 0), wf_desc((size_t)NW * std::max(nwide, 1), 0);
  O.wf_on = 0;
  O.wf_buf_bytes = 0;
  {
    // Built and measured (round 2): neutral at disc 20 (1.992 vs 1.996 M evals/s) -- a warp's own dependent chain through
    // the levels takes as long as the block-wide steps it replaces -- so it stays behind a knob.
    static const bool off = getenv("TNO_ONCHIP_WARP_FACTOR") == nullptr;
    bool wf = wl_balanced && !off;
    size_t maxbytes = 0;
    for (int w = 0; w < NW && wf; ++w) {
      wf_off[w] = (uint32_t)(wf_words.size() / 4);
      const size_t w0 = wf_words.size();
      for (int lev = 1; lev < nwide && wf; ++lev) {
        std::vector<Item> items;
        for (int j = S.level_ptr[lev]; j < S.level_ptr[lev + 1]; ++j)
          if (wl_warp_of[wl_root[j]] == w)
            for (int p = S.colptr[j]; p < S.colptr[j + 1]; ++p) items.push_back(make_item(p, p == S.colptr[j], -1));
        if (items.empty()) continue;
        std::stable_sort(items.begin(), items.end(), [](const Item& a, const Item& b) { return a.prs.size() > b.prs.size(); });
        const size_t off_units = (wf_words.size() - w0) / 4;
        const int nitems = (int)items.size();
        size_t npairs = 0;
        for (const Item& it : items) {
          wf_words.push_back(it.pf);
          wf_words.push_back((uint32_t)npairs);
          npairs += it.prs.size();
        }
        wf_words.push_back(0);
        wf_words.push_back((uint32_t)npairs);
        for (const Item& it : items)
          for (unsigned long long pr : it.prs) {
            wf_words.push_back((uint32_t)(pr & 0xffffffffu));
            wf_words.push_back((uint32_t)(pr >> 32));
          }
        while (wf_words.size() % 4) wf_words.push_back(0);
        int lpi_shift = 0;
        double best = 1e300;
        const size_t maxp = items[0].prs.size();
        for (int sft = 0; sft <= 5; ++sft) {
          const int passes = ((nitems << sft) + 31) / 32;
          const double cost = passes * ((double)((maxp + (1u << sft) - 1) >> sft) + 0.8 * sft + 2.0);
          if (cost < best - 1e-9) {
            best = cost;
            lpi_shift = sft;
          }
        }
        if (off_units >= 65536 || nitems >= 4096) wf = false;
        wf_desc[(size_t)w * nwide + lev] = (uint32_t)off_units | ((uint32_t)nitems << 16) | ((uint32_t)lpi_shift << 28);
      }
      maxbytes = std::max(maxbytes, (wf_words.size() - w0) * 4);
    }
    wf_off[NW] = (uint32_t)(wf_words.size() / 4);
    // the warps' buffers share the first two ring buffers while the first chain chunk is already on its way into the third
    if (wf && (size_t)NW * maxbytes > 2 * (CHUNK_BYTES / 16 * 16)) wf = false;
    if (wf) {
      O.wf_on = 1;
      O.wf_buf_bytes = (int)maxbytes;
      max_units = std::max<uint32_t>(max_units, (uint32_t)(((size_t)NW * maxbytes + 31) / 32));   // 2 buffers >= NW x maxbytes
    } else {
      wf_words.clear();
      std::fill(wf_desc.begin(), wf_desc.end(), 0u);
      for (int lev = 1; lev < nwide; ++lev) {
        std::vector<Item> items;
        for (int j = S.level_ptr[lev]; j < S.level_ptr[lev + 1]; ++j)
          for (int p = S.colptr[j]; p < S.colptr[j + 1]; ++p) items.push_back(make_item(p, p == S.colptr[j], -1));
        emit_step(items, 0);
      }
    }
    if (wf_words.empty()) wf_words.assign(4, 0);
  }
  for (int cl = 0; cl < O.nclev; ++cl) {
    const int c0 = S.clevel_ptr[cl], c1 = S.clevel_ptr[cl + 1];
    std::vector<Item> sitems, ritems;
    for (int c = c0; c < c1; ++c) {
      const int r0 = S.chain_ptr[c], r1 = S.chain_ptr[c + 1];
      for (int j = r0; j < r1; ++j)
        for (int p = S.colptr[j]; p < S.colptr[j + 1]; ++p) sitems.push_back(make_item(p, false, r0));  // S = A - outside pairs
      // rectangular rows, columns in descending order:  U_ij = S_ij + sum_{c < j in chain, (i, c) present} S_ic Z_jc
      for (int j = r1 - 1; j >= r0; --j)
        for (int p = S.colptr[j] + 1; p < S.colptr[j + 1]; ++p) {
          const int i = S.rowidx[p];
          if (i < r1) continue;  // triangle entry
          Item it;
          it.pf = (uint32_t)newpos[p];
          for (int cc = r0; cc < j; ++cc) {
            const int pic = S.pos(i, cc);
            if (pic < 0) continue;
            it.prs.push_back((unsigned long long)newpos[pic] | ((unsigned long long)zpos(j, cc) << 16) |
                             ((unsigned long long)O.pos_m1 << 32));
          }
          ritems.push_back(it);
        }
    }
    emit_step(sitems, 0);
    OcOp OP;
    std::memset(&OP, 0, sizeof(OP));
    OP.type = OC_OP_DENSE;
    OP.a = (unsigned short)cl;
    ops.push_back(OP);
    {
      OcDense DN;
      std::memset(&DN, 0, sizeof(DN));
      DN.de0 = (uint32_t)de.size();
      DN.dc0 = (uint32_t)dc.size();
      DN.zlo = (uint32_t)(nlo + ni + zoff_rel[c0]);
      for (int c = c0; c < c1; ++c) {
        const int r0 = S.chain_ptr[c], T = S.chain_ptr[c + 1] - r0;
        const int zoff = nlo + ni + zoff_rel[c];
        DN.tmax = std::max<uint32_t>(DN.tmax, (uint32_t)T);
        DN.zhi = (uint32_t)(zoff + T * (T - 1) / 2);
        for (int i = 0; i < T; ++i)
          for (int kk = 0; kk <= i; ++kk) {
            const uint32_t rb_i = (uint32_t)(zoff + i * (i - 1) / 2), rb_k = (uint32_t)(zoff + kk * (kk - 1) / 2);
            const uint32_t self = kk == i ? (uint32_t)(nlo + r0 + i) : rb_i + (uint32_t)kk;
            uint4 d;
            d.x = rb_i | (rb_k << 16);
            d.y = self | ((uint32_t)(nlo + r0) << 16);
            d.z = (uint32_t)kk | ((uint32_t)i << 8);
            d.w = 0;
            de.push_back(d);
          }
        for (int cc = 0; cc + 1 < T; ++cc) dc.push_back((uint32_t)zoff | ((uint32_t)T << 16) | ((uint32_t)cc << 24));
      }
      std::stable_sort(de.begin() + DN.de0, de.end(),
                       [](const uint4& a, const uint4& b) { return (a.z & 0xffu) > (b.z & 0xffu); });
      DN.de1 = (uint32_t)de.size();
      DN.dc1 = (uint32_t)dc.size();
      // blocked inversion: the 8 x 8 blocks below the block diagonal, grouped by their distance from it
      DN.db0 = (uint32_t)db.size();
      for (int c = c0; c < c1; ++c) DN.nbmax = std::max<unsigned short>(DN.nbmax, (unsigned short)((S.chain_ptr[c + 1] - S.chain_ptr[c] + 7) / 8));
      for (int st = 1; st <= 8; ++st) {
        DN.sb[st] = (unsigned short)(db.size() - DN.db0);
        if (st == 8) break;
        for (int c = c0; c < c1; ++c) {
          const int T = S.chain_ptr[c + 1] - S.chain_ptr[c], nbk = (T + 7) / 8;
          for (int ib = st; ib < nbk; ++ib)
            db.push_back((uint32_t)(nlo + ni + zoff_rel[c]) | ((uint32_t)T << 16) | ((uint32_t)ib << 24) | ((uint32_t)st << 28));
        }
      }
      // register tiles: one 4 x 4 tile of a triangle per thread, when the level has no more tiles than threads
      {
        static const bool elementwise = getenv("TNO_ONCHIP_DENSE_ELEMENTS") != nullptr;   // developer knob (A/B measurements)
        DN.dt0 = DN.dt1 = (uint32_t)dt.size();
        DN.tile_threads = 0;
        for (int TS = getenv("TNO_ONCHIP_TILE4") ? 4 : 2; TS <= 4 && !elementwise && DN.tile_threads == 0; TS += 2) {   // TNO_ONCHIP_TILE4: developer knob (tests)
          std::vector<uint4> tiles;
          uint32_t cbo = 0;
          for (int c = c0; c < c1; ++c) {
            const int r0 = S.chain_ptr[c], T = S.chain_ptr[c + 1] - r0, nt = (T + TS - 1) / TS;
            for (int ti = 0; ti < nt; ++ti)
              for (int tj = 0; tj <= ti; ++tj) {
                uint4 d;
                d.x = (uint32_t)(nlo + ni + zoff_rel[c]) | ((uint32_t)(nlo + r0) << 16);
                d.y = (uint32_t)T | ((uint32_t)ti << 8) | ((uint32_t)tj << 16);
                d.z = cbo;
                d.w = 0;
                tiles.push_back(d);
              }
            cbo += (uint32_t)(TS * nt);
          }
          if ((int)tiles.size() > THREADS) continue;
          dt.insert(dt.end(), tiles.begin(), tiles.end());
          DN.dt1 = (uint32_t)dt.size();
          DN.cbtot = cbo;
          DN.tile_size = (uint32_t)TS;
          DN.tile_threads = (uint32_t)((tiles.size() + 31) / 32 * 32);
          max_tile_buf = std::max(max_tile_buf, (int)(10 * cbo));
        }
      }
      dense_r.push_back(DN);
    }
    if (!ritems.empty()) emit_step(ritems, 1);
  }
  O.nchunks = (int)chunks.size();
  O.nops = (int)ops.size();
  if (getenv("TNO_PHASE_CLOCKS") && !dry) {
    fprintf(stderr, "[tno factor program]");
    for (const OcOp& op : ops) {
      if (op.type == OC_OP_DENSE) fprintf(stderr, " DENSE(level %d)", (int)op.a);
      else fprintf(stderr, " chunk(kind %d, %d items, %u bytes, %d lanes/item)", (int)chunks[op.a].kind, (int)chunks[op.a].nitems,
                   chunks[op.a].seg_units * 16, 1 << chunks[op.a].lpi_shift);
    }
    fprintf(stderr, "\n");
  }
  O.fseg_units = (int)max_units;
  if (chunks.size() >= 65535) { set_error("on-chip family not applicable: chunks.size() >= 65535"); return TNO_OK; }

  // ---- solve ranges ----
  auto make_range = [&](int row0, int row1, int chain0, int chain1) {
    OcRange L;
    std::memset(&L, 0, sizeof(L));
    L.row0 = (unsigned short)row0;
    L.row1 = (unsigned short)row1;
    L.chain0 = (unsigned short)chain0;
    L.chain1 = (unsigned short)chain1;
    const int rows = row1 - row0;
    for (int ph = 0; ph < 2; ++ph) {
      const int g = O.gshift[ph];
      int sp = 0;
      while (sp + 1 + g <= 5 && (rows << (sp + 1 + g)) <= THREADS) ++sp;
      int sq = 0;
      while (sq < 5 && (rows << (sq + 1 + g)) <= THREADS) ++sq;
      L.pull_s[ph] = (unsigned char)sp;
      L.push_s[ph] = (unsigned char)sq;
    }
    return L;
  };
  std::vector<OcRange> wide_r, clev_r;
  for (int lev = 0; lev < nwide; ++lev) wide_r.push_back(make_range(S.level_ptr[lev], S.level_ptr[lev + 1], 0, 0));
  std::vector<OcChain> chain_r(std::max(nch, 1));
  std::memset(chain_r.data(), 0, chain_r.size() * sizeof(OcChain));
  std::vector<unsigned char> rowchain(std::max(ni - wide_rows, 1), 0);
  for (int cl = 0; cl < O.nclev; ++cl) {
    const int c0 = S.clevel_ptr[cl], c1 = S.clevel_ptr[cl + 1];
    if (c1 - c0 > 255) { set_error("on-chip family not applicable: c1 - c0 > 255"); return TNO_OK; }
    clev_r.push_back(make_range(S.chain_ptr[c0], S.chain_ptr[c1], c0, c1));
    for (int c = c0; c < c1; ++c) {
      chain_r[c].row0 = (unsigned short)S.chain_ptr[c];
      chain_r[c].T = (unsigned short)(S.chain_ptr[c + 1] - S.chain_ptr[c]);
      chain_r[c].zoff = (uint32_t)(nlo + ni + zoff_rel[c]);
      for (int r = S.chain_ptr[c]; r < S.chain_ptr[c + 1]; ++r) rowchain[r - wide_rows] = (unsigned char)(c - c0);
    }
  }
  // 8-row tiles of the chains' dense blocks (tensor-pipe variant of the dense chain products), high tile rows first
  std::vector<uint2> zt_tab;
  for (int cl = 0; cl < O.nclev; ++cl) {
    const int c0 = S.clevel_ptr[cl], c1 = S.clevel_ptr[cl + 1];
    clev_r[cl].zt0 = (unsigned short)zt_tab.size();
    int mtmax = 0;
    for (int c = c0; c < c1; ++c) mtmax = std::max(mtmax, (S.chain_ptr[c + 1] - S.chain_ptr[c] + 7) / 8);
    for (int mt = mtmax - 1; mt >= 0; --mt)
      for (int c = c0; c < c1; ++c) {
        const int T = S.chain_ptr[c + 1] - S.chain_ptr[c];
        if (8 * mt >= T) continue;
        uint2 d;
        d.x = (uint32_t)(nlo + ni + zoff_rel[c]) | ((uint32_t)S.chain_ptr[c] << 16);
        d.y = (uint32_t)T | ((uint32_t)mt << 8);
        zt_tab.push_back(d);
      }
    clev_r[cl].zt1 = (unsigned short)zt_tab.size();
  }
  if (zt_tab.empty()) zt_tab.push_back(uint2{0, 0});
  // ---- warp-local wide part of the solves: row lists per warp and level ----
  std::vector<unsigned short> wl_ptr, wl_rows;
  std::vector<unsigned char> wl_sub;
  O.wl_on = 0;
  {
    // Built and measured (round 2): +1.2 % before the other changes of the round, -0.7 % after them (1.993 vs 2.008 M evals/s
    // at disc 20): the barriers of the wide levels are not what their steps cost.  Behind a knob.
    static const bool off = getenv("TNO_ONCHIP_WARP_LOCAL") == nullptr;
    const std::vector<int>& root = wl_root;
    const std::vector<int>& warp_of = wl_warp_of;
    if (!off && wl_balanced) {
      O.wl_on = 1;
      wl_ptr.assign((size_t)NW * (nwide + 1), 0);
      wl_sub.assign((size_t)NW * nwide * 4, 0);
      for (int w = 0; w < NW; ++w) {
        for (int l = 0; l < nwide; ++l) {
          wl_ptr[(size_t)w * (nwide + 1) + l] = (unsigned short)wl_rows.size();   // start of level l = end of level l - 1
          if (l == 0) continue;   // the leaves have no entries: nothing to pull or push
          int count = 0;
          for (int j = S.level_ptr[l]; j < S.level_ptr[l + 1]; ++j)
            if (warp_of[root[j]] == w) {
              wl_rows.push_back((unsigned short)j);
              ++count;
            }
          for (int ph = 0; ph < 2; ++ph) {
            const int g = O.gshift[ph];
            int sp = 0;
            while (count > 0 && sp + 1 + g <= 5 && (count << (sp + 1 + g)) <= 32) ++sp;   // S * G <= 32 (shuffle reduction)
            wl_sub[(((size_t)w * nwide + l) * 2 + ph) * 2 + 0] = (unsigned char)sp;   // pull
            wl_sub[(((size_t)w * nwide + l) * 2 + ph) * 2 + 1] = (unsigned char)sp;   // push
          }
        }
        wl_ptr[(size_t)w * (nwide + 1) + nwide] = (unsigned short)wl_rows.size();
      }
    }
    if (wl_ptr.empty()) wl_ptr.push_back(0);
    if (wl_rows.empty()) wl_rows.push_back(0);
    if (wl_sub.empty()) wl_sub.push_back(0);
  }
  // processing order of the rows of a chain level in the forward (pull) step: longest rows first, and as many sub-slots per
  // row as minimise  sum over passes of (longest row of the pass / sub-slots + a per-pass overhead)
  std::vector<unsigned short> rowlist;
  {
    static const bool off = getenv("TNO_ONCHIP_ROWLIST") == nullptr;   // developer knob: measured 1 % slower than row order at disc 20, off
    for (int cl = 0; cl < O.nclev && !off; ++cl) {
      OcRange& L = clev_r[cl];
      std::vector<int> rows;
      for (int r = L.row0; r < L.row1; ++r) rows.push_back(r);
      std::stable_sort(rows.begin(), rows.end(), [&](int a, int b) {
        return rowptr_o[a + 1] - rowptr_o[a] > rowptr_o[b + 1] - rowptr_o[b];
      });
      L.rl0 = (unsigned short)rowlist.size();
      L.rl_on = 1;
      for (int r : rows) rowlist.push_back((unsigned short)r);
      for (int ph = 0; ph < 2; ++ph) {
        const int g = O.gshift[ph];
        int best_s = 0;
        double best = 1e300;
        for (int sft = 0; sft + g <= 5; ++sft) {
          const int per_pass = THREADS >> (sft + g);
          if (per_pass < 1) break;
          double cost = 0.0;
          for (size_t i0 = 0; i0 < rows.size(); i0 += per_pass) {
            const int len = rowptr_o[rows[i0] + 1] - rowptr_o[rows[i0]];   // longest row of the pass
            cost += (double)((len + (1 << sft) - 1) >> sft) + 4.0 + 1.5 * sft;
          }
          if (cost < best) {
            best = cost;
            best_s = sft;
          }
        }
        L.pull_s[ph] = (unsigned char)best_s;
      }
    }
    if (rowlist.empty()) rowlist.push_back(0);
  }
  O.zmma = getenv("TNO_ONCHIP_NO_DMMA") ? 0 : 1;   // developer knob (A/B measurements): dense chain products with DFMA
  std::vector<unsigned short> ct_ptr(1, 0);
  std::vector<unsigned short> ct_col;
  std::vector<unsigned short> ct_ent;
  std::vector<unsigned char> ct_row;
  bool ct_rows_fit = true;
  for (int cl = 0; cl < O.nclev; ++cl) {
    const int c0 = S.clevel_ptr[cl], c1 = S.clevel_ptr[cl + 1];
    clev_r[cl].tgt0 = (unsigned short)ct_col.size();
    // outside column -> entries of the level's chain rows in it (the chains of one level reach disjoint columns)
    std::map<int, std::vector<uint32_t>> by_col;
    for (int c = c0; c < c1; ++c)
      for (int i = S.chain_ptr[c]; i < S.chain_ptr[c + 1]; ++i)
        for (int t = rowptr_o[i]; t < rowptr_o[i + 1]; ++t) by_col[rowcol_o[t]].push_back((uint32_t)t | ((uint32_t)i << 16));
    // longest targets first: the targets of one pass (THREADS >> gshift of them) then have similar lengths, and a pass
    // takes as long as its longest target
    std::vector<std::pair<int, std::vector<uint32_t>>> tl(by_col.begin(), by_col.end());
    std::stable_sort(tl.begin(), tl.end(), [](const std::pair<int, std::vector<uint32_t>>& a, const std::pair<int, std::vector<uint32_t>>& b) {
      return a.second.size() > b.second.size();
    });
    for (auto& kv : tl) {
      ct_col.push_back((unsigned short)kv.first);
      for (uint32_t w : kv.second) {
        const int rel = (int)(w >> 16) - S.chain_ptr[c0];
        if (rel > 255) ct_rows_fit = false;
        ct_ent.push_back((unsigned short)(w & 0xffffu));
        ct_row.push_back((unsigned char)(rel & 0xff));
      }
      ct_ptr.push_back((unsigned short)ct_ent.size());
    }
    clev_r[cl].tgt1 = (unsigned short)ct_col.size();
  }
  if (!ct_rows_fit) { set_error("on-chip family not applicable: a chain level has more than 256 rows"); return TNO_OK; }
  if (ct_col.size() >= 65535 || ct_ent.size() >= 65535) { set_error("on-chip family not applicable: ct_col.size() >= 65535"); return TNO_OK; }
  if (ct_col.empty()) ct_col.push_back(0);
  if (ct_ent.empty()) ct_ent.push_back(0);
  if (ct_row.empty()) ct_row.push_back(0);
  if (clev_r.empty()) clev_r.push_back(make_range(0, 0, 0, 0));

  // ---- assembly maps in V position order ----
  std::vector<int> asm_ptr(nv, 0), asm_edge;
  std::vector<float> asm_coef;
  {
    std::vector<std::vector<std::pair<int, float>>> contrib(nv - 1);
    for (int e = 0; e < m; ++e) {
      const int u = D->edges[2 * e], v = D->edges[2 * e + 1];
      const int ru = PL->h_vrow[u], rv = PL->h_vrow[v];
      if (ru >= 0) contrib[nlo + ru].push_back({e, -1.f});
      if (rv >= 0) contrib[nlo + rv].push_back({e, -1.f});
      if (ru >= 0 && rv >= 0) contrib[newpos[S.pos(std::max(ru, rv), std::min(ru, rv))]].push_back({e, +1.f});
    }
    for (int p = 0; p < nv - 1; ++p) {
      for (auto& c : contrib[p]) {
        asm_edge.push_back(c.first);
        asm_coef.push_back(c.second);
      }
      asm_ptr[p + 1] = (int)asm_edge.size();
    }
  }
  std::vector<unsigned short> asm_e16(std::max(nlo, 1), 0xffff), asm_z16(std::max(nz, 1), 0xffff);
  O.simple_asm = 1;
  for (int p = 0; p < nv - 1 && O.simple_asm; ++p) {
    if (p >= nlo && p < nlo + ni) continue;
    const int cnt = asm_ptr[p + 1] - asm_ptr[p];
    if (cnt > 1 || (cnt == 1 && asm_coef[asm_ptr[p]] != 1.f)) O.simple_asm = 0;
    if (cnt == 1) (p < nlo ? asm_e16[p] : asm_z16[p - nlo - ni]) = (unsigned short)asm_edge[asm_ptr[p]];
  }

  // ---- row adjacency records ----
  std::vector<int> radj_ptr(ni + 1, 0);
  std::vector<uint32_t> radj, radj_pad;
  int maxdeg = 0;
  {
    std::vector<int> row_vertex(ni, -1);
    for (int v = 0; v < P.n; ++v)
      if (PL->h_vrow[v] >= 0) row_vertex[PL->h_vrow[v]] = v;
    for (int r = 0; r < ni; ++r) {
      const int v = row_vertex[r];
      for (int t = PL->h_vadj_ptr[v]; t < PL->h_vadj_ptr[v + 1]; ++t) {
        const int ro = PL->h_vrow[PL->h_vadj_other[t]];
        const uint32_t other = ro >= 0 ? (uint32_t)ro : (0x8000u | (uint32_t)(-1 - ro));
        radj.push_back(other | ((uint32_t)PL->h_vadj_edge[t] << 16));
      }
      radj_ptr[r + 1] = (int)radj.size();
      maxdeg = std::max(maxdeg, radj_ptr[r + 1] - radj_ptr[r]);
    }
    O.adj_deg = maxdeg <= 8 ? maxdeg : 0;
    if (O.adj_deg > 0) {
      radj_pad.assign((size_t)ni * O.adj_deg, 0);
      for (int r = 0; r < ni; ++r)
        for (int t = 0; t < O.adj_deg; ++t) {
          const int src = radj_ptr[r] + t;
          // padding: "other end" = the row itself (z_v - z_v = 0), any valid edge
          radj_pad[(size_t)r * O.adj_deg + t] = src < radj_ptr[r + 1] ? radj[src] : ((uint32_t)r | (radj[radj_ptr[r]] & 0xffff0000u));
        }
    } else {
      radj_pad.assign(1, 0);
    }
  }

  // ---- coefficient rows of the phase-2 right-hand sides in row-adjacency order ----
  // Br[t] = [B[e, 0..k-1] | d0[e] (lambdh) | d0_b[e] (second load basis)] of the edge e of adjacency record t, zero padded
  // to br_stride (even, >= the panel width): the kernel reads it with 16-byte loads, two columns per lane.
  std::vector<double> Br;
  {
    const int cols = k + (P.var.o_lambdh >= 0 ? 1 : 0) + (D->d_b ? 1 : 0);
    O.br_stride = std::max(O.W * O.npass2, (cols + 1) / 2 * 2);
    Br.assign(std::max<size_t>(radj.size(), 1) * O.br_stride, 0.0);
    for (size_t t = 0; t < radj.size(); ++t) {
      const int e = (int)(radj[t] >> 16);
      double* dst = Br.data() + t * O.br_stride;
      for (int j = 0; j < k; ++j) dst[j] = D->B[(size_t)e * k + j];
      if (P.var.o_lambdh >= 0) dst[k] = D->d[e];
      if (D->d_b) dst[k + 1] = D->d_b[e];
    }
  }

  std::vector<int> colsrc(nvar, -1);
  for (int j = 0; j < k; ++j) colsrc[j] = O.c_q >= 0 ? O.c_q + j : -2;
  if (P.var.o_lambdh >= 0) colsrc[P.var.o_lambdh] = O.c_lambdh >= 0 ? O.c_lambdh : -2;
  if (P.var.o_lambdv >= 0) colsrc[P.var.o_lambdv] = -3;  // from the persistent copy
  // ---- shared-memory index region image ----
  std::vector<unsigned char> idx;
  auto align16 = [&]() {
    while (idx.size() % 16) idx.push_back(0);
  };
  auto put16 = [&](int v) {
    idx.push_back((unsigned char)(v & 0xff));
    idx.push_back((unsigned char)((v >> 8) & 0xff));
  };
  auto put_bytes = [&](const void* ptr, size_t bytes) {
    const unsigned char* bp = reinterpret_cast<const unsigned char*>(ptr);
    idx.insert(idx.end(), bp, bp + bytes);
  };
  O.idx_rowptr = 0;
  for (int i = 0; i <= ni; ++i) put16(rowptr_o[i]);
  align16();
  O.idx_rowcol = (int)idx.size();
  for (int t = 0; t < nlo; ++t) put16(rowcol_o[t]);
  align16();
  O.idx_wide = (int)idx.size();
  put_bytes(wide_r.data(), wide_r.size() * sizeof(OcRange));
  align16();
  O.idx_clev = (int)idx.size();
  put_bytes(clev_r.data(), clev_r.size() * sizeof(OcRange));
  align16();
  O.idx_chains = (int)idx.size();
  put_bytes(chain_r.data(), chain_r.size() * sizeof(OcChain));
  align16();
  O.idx_rowchain = (int)idx.size();
  put_bytes(rowchain.data(), rowchain.size());
  align16();
  O.idx_chunks = (int)idx.size();
  put_bytes(chunks.data(), chunks.size() * sizeof(OcChunk));
  align16();
  O.idx_ops = (int)idx.size();
  put_bytes(ops.data(), ops.size() * sizeof(OcOp));
  align16();
  O.idx_dense = (int)idx.size();
  if (dense_r.empty()) dense_r.resize(1);
  put_bytes(dense_r.data(), dense_r.size() * sizeof(OcDense));
  align16();
  {
    std::vector<int> row_vertex(ni, -1);
    for (int v = 0; v < P.n; ++v)
      if (PL->h_vrow[v] >= 0) row_vertex[PL->h_vrow[v]] = v;
    auto put_table = [&](int* off, const void* ptr, size_t bytes) {
      *off = (int)idx.size();
      put_bytes(ptr, bytes);
      align16();
    };
    std::vector<short> vrow16(PL->h_vrow.begin(), PL->h_vrow.end()), rowv16(row_vertex.begin(), row_vertex.end());
    put_table(&O.idx_vrow, vrow16.data(), vrow16.size() * 2);
    put_table(&O.idx_rowv, rowv16.data(), rowv16.size() * 2);
    put_table(&O.idx_colsrc, colsrc.data(), colsrc.size() * 4);
    put_table(&O.idx_supptr, sup_ptr.data(), sup_ptr.size() * 4);
    put_table(&O.idx_supother, sup_other.data(), std::max<size_t>(sup_other.size(), 1) * 4);
    put_table(&O.idx_supedge, sup_edge.data(), std::max<size_t>(sup_edge.size(), 1) * 4);
    put_table(&O.idx_supsign, sup_sign.data(), std::max<size_t>(sup_sign.size(), 1) * 4);
    put_table(&O.idx_fixed, PL->h_fixed.data(), PL->h_fixed.size() * 4);
    // row pointers of the row adjacency (16 bit) when they fit
    O.radjp_in_smem = radj.size() < 65535 ? 1 : 0;
    std::vector<unsigned short> radjp16(O.radjp_in_smem ? ni + 1 : 1, 0);
    if (O.radjp_in_smem)
      for (int r = 0; r <= ni; ++r) radjp16[r] = (unsigned short)radj_ptr[r];
    put_table(&O.idx_radjp, radjp16.data(), radjp16.size() * 2);
    std::vector<double> supxyz((size_t)3 * std::max(nb, 1), 0.0);
    for (int b = 0; b < nb; ++b)
      for (int a = 0; a < 3; ++a) supxyz[3 * b + a] = D->X[(size_t)3 * PL->h_fixed[b] + a];
    put_table(&O.idx_supxyz, supxyz.data(), supxyz.size() * 8);
    put_table(&O.idx_zt, zt_tab.data(), zt_tab.size() * sizeof(uint2));
    put_table(&O.idx_rowlist, rowlist.data(), rowlist.size() * 2);
    put_table(&O.idx_wf_off, wf_off.data(), wf_off.size() * 4);
    put_table(&O.idx_wf_desc, wf_desc.data(), wf_desc.size() * 4);
    put_table(&O.idx_wl_ptr, wl_ptr.data(), wl_ptr.size() * 2);
    put_table(&O.idx_wl_rows, wl_rows.data(), wl_rows.size() * 2);
    put_table(&O.idx_wl_s, wl_sub.data(), wl_sub.size());
  }
  const int idx_bytes_core = (int)idx.size();
  O.idx_ct_ptr = (int)idx.size();
  put_bytes(ct_ptr.data(), ct_ptr.size() * 2);
  align16();
  O.idx_ct_col = (int)idx.size();
  put_bytes(ct_col.data(), ct_col.size() * 2);
  align16();
  O.idx_ct_ent = (int)idx.size();
  put_bytes(ct_ent.data(), ct_ent.size() * 2);
  align16();
  O.idx_ct_row = (int)idx.size();
  put_bytes(ct_row.data(), ct_row.size());
  align16();
  O.idx_bytes = (int)idx.size();
  O.ct_in_smem = 1;

  // ---- shared memory layout ----
  int off = 0;  // doubles
  auto take = [&](int count) {
    const int o = off;
    off += (count + 1) / 2 * 2;  // keep 16-byte alignment
    return o;
  };
  O.o_x = take(nvar);
  O.o_v = take(nv);
  const int panel = std::max(ni * O.W, m);
  const int stage_doubles = 3 * O.fseg_units * 2;
  O.o_tmp_doubles = stage_doubles;
  O.o_tmpz_doubles = (stage_doubles + max_tmp_items + 2) / 2 * 2;
  O.o_r = take(std::max(panel, O.o_tmpz_doubles + std::max(dense_scratch, max_tile_buf) + 2));
  O.o_zs = take(ni + nb);
  O.o_pers = take(O.npers > 1 ? ni : 0);
  O.o_qsup = take(O.nsup);
  O.o_react = take(6 * nb);   // reactions, then |R_h| and the direction cosines of the thrust objectives
  O.o_gxy = take(2 * nb * k + 2 * nb);
  O.o_azsup = take((((P.constraints & TNO_CON_REAC_BOUNDS) || P.objective == TNO_OBJ_ECOMP_LINEAR || P.objective == TNO_OBJ_ECOMP_NONLINEAR) &&
                     P.var.o_zb >= 0) ? O.nsup * nb : 0);
  O.o_red = take(40);
  const bool lp = P.objective == TNO_OBJ_LOADPATH, enl = P.objective == TNO_OBJ_ECOMP_NONLINEAR;
  O.o_bf = take((P.objective == TNO_OBJ_BESTFIT || lp || enl) ? (THREADS / 32) * 16 : 0);
  O.o_lpe = take((lp || enl) ? m : 0);
  O.o_lpa = take(lp ? std::max(m, k) : (enl ? k : 0));
  O.o_lpw = take(lp ? ni + P.nb : 0);
  O.o_idx_bytes = off * 8;   // bytes of one candidate group; the shared index region follows the last group
  int max_optin = 0, smem_sm = 0;
  TNO_CUDA_OK(cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, PL->device));
  TNO_CUDA_OK(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, PL->device));
  const int G = LY.groups;
  // resident CTAs per SM by shared memory and by registers (256-thread single-group CTAs are compiled for two per SM)
  auto ctas = [&](int bytes) {
    if (bytes > max_optin) return 0;
    const int by_smem = smem_sm / (bytes + 1024);
    return G > 1 || THREADS == 512 ? std::min(by_smem, 1) : std::min(by_smem, 2);
  };
  {
    // keep the column-wise chain index in shared memory only if that does not cost a resident CTA
    const int with_ct = G * O.o_idx_bytes + O.idx_bytes, without_ct = G * O.o_idx_bytes + idx_bytes_core;
    if (ctas(with_ct) < ctas(without_ct) || with_ct > max_optin) {
      O.ct_in_smem = 0;
      O.idx_bytes = idx_bytes_core;
    }
  }
  O.smem_bytes = G * O.o_idx_bytes + O.idx_bytes;
  if (O.smem_bytes > max_optin) { set_error("on-chip family not applicable: O.smem_bytes > max_optin"); return TNO_OK; }
  PL->onchip_threads = THREADS;
  PL->onchip_groups = G;
  PL->onchip_ctas = std::max(1, ctas(O.smem_bytes));
  PL->onchip_ctas_per_sm = PL->onchip_ctas * G;   // candidates in flight per SM
  *cands_per_sm = PL->onchip_ctas_per_sm;
  if (dry) {
    PL->onchip_ok = true;
    return TNO_OK;
  }

  std::vector<double> Bt((size_t)k * m);
  for (int e = 0; e < m; ++e)
    for (int j = 0; j < k; ++j) Bt[(size_t)j * m + e] = D->B[(size_t)e * k + j];
  std::vector<double> Jfun;
  if (P.con.r_fun >= 0) {
    Jfun.assign((size_t)2 * m * nvar, 0.0);
    for (int e = 0; e < m; ++e) {
      for (int j = 0; j < k; ++j) {
        Jfun[(size_t)e * nvar + j] = D->B[(size_t)e * k + j];
        Jfun[(size_t)(m + e) * nvar + j] = -D->B[(size_t)e * k + j];
      }
      if (P.var.o_lambdh >= 0) {
        Jfun[(size_t)e * nvar + P.var.o_lambdh] = D->d[e];
        Jfun[(size_t)(m + e) * nvar + P.var.o_lambdh] = -D->d[e];
      }
    }
  }

  int rc;
  const void* dptr = nullptr;
#define UPO(field, hostvec) \
  if ((rc = upload_vec(PL, hostvec, &O.field))) return rc;
  UPO(Bt, Bt);
  UPO(asm_ptr, asm_ptr);
  UPO(asm_edge, asm_edge);
  UPO(asm_coef, asm_coef);
  UPO(asm_e16, asm_e16);
  UPO(asm_z16, asm_z16);
  if ((rc = upload_bytes(PL, stream.data(), stream.size() * 4, &dptr))) return rc;
  O.f_stream = reinterpret_cast<const uint4*>(dptr);
  if ((rc = upload_bytes(PL, idx.data(), idx.size(), &dptr))) return rc;
  O.idx_init = reinterpret_cast<const unsigned char*>(dptr);
  if ((rc = upload_bytes(PL, wf_words.data(), wf_words.size() * 4, &dptr))) return rc;
  O.wf_stream = reinterpret_cast<const uint4*>(dptr);
  UPO(edge_sup, edge_sup);
  UPO(sup_ptr, sup_ptr);
  UPO(sup_edge, sup_edge);
  UPO(sup_other, sup_other);
  UPO(sup_sign, sup_sign);
  UPO(colsrc, colsrc);
  UPO(radj_ptr, radj_ptr);
  UPO(radj, radj);
  UPO(radj_pad, radj_pad);
  UPO(Br, Br);
  UPO(ct_ptr, ct_ptr);
  UPO(ct_col, ct_col);
  UPO(ct_ent, ct_ent);
  UPO(ct_row, ct_row);
  if (de.empty()) de.push_back(uint4{0, 0, 0, 0});
  if (dc.empty()) dc.push_back(0);
  UPO(de, de);
  UPO(dc, dc);
  if (db.empty()) db.push_back(0);
  UPO(db, db);
  if (dt.empty()) dt.push_back(uint4{0, 0, 0, 0});
  UPO(dt, dt);
  UPO(Jfun, Jfun);
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
  A.dir = (opt && opt->load_dir) ? opt->load_dir : nullptr;
  A.bounds = (opt && opt->bounds) ? opt->bounds : nullptr;
  if (A.bounds && (PL->dp.var.o_t >= 0 || PL->dp.con.r_env < 0)) {
    set_error("per-candidate bounds need the envelope constraint and no thickness variable");
    return TNO_EUNSUPPORTED;
  }
  if (A.dir && !PL->dp.d_b) {
    set_error("load_dir needs a plan created with the second load basis (d_b, px0_b, py0_b)");
    return TNO_EUNSUPPORTED;
  }
  A.f = out->f;
  A.grad = out->grad;
  A.g = out->g;
  A.J = out->J;
  A.z = out->z;
  A.q = out->q;
  A.gmin = out->gmin;
  A.status = out->status;
  A.N = N;
  A.phase_clocks = nullptr;
  static const bool want_clocks = getenv("TNO_PHASE_CLOCKS") != nullptr;
  if (want_clocks) {
    TNO_CUDA_OK(cudaMalloc((void**)&A.phase_clocks, 128 * sizeof(long long)));
    TNO_CUDA_OK(cudaMemsetAsync(A.phase_clocks, 0, 128 * sizeof(long long), stream));
  }
  const size_t smem = (size_t)PL->oc.smem_bytes;
  static const char* force_ctas = getenv("TNO_CTAS_PER_SM");  // developer knob
  const int G = PL->onchip_groups;
  const int ctas = force_ctas ? std::max(1, atoi(force_ctas)) : PL->onchip_ctas;
  const int grid = (int)std::min<int64_t>((N + G - 1) / G, (int64_t)PL->sm_count * ctas);
  const int threads = PL->onchip_threads * G;
  const bool dir = A.dir || A.bounds;
  // <256, 2, 1>: up to two single-candidate CTAs per SM; <256, 1, G>: one CTA of G candidate groups sharing the index
  // region; <512, 1, 1>: one 512-thread candidate per SM (large panels)
  void (*kern)(const DevPlan, const DevOnchip, const OcArgs) = nullptr;
  if (PL->onchip_threads == 512) kern = dir ? k_onchip<512, 1, 1, true, false> : (want_clocks ? k_onchip<512, 1, 1, false, true> : k_onchip<512, 1, 1, false, false>);
  else if (G == 1) kern = dir ? k_onchip<256, 2, 1, true, false> : (want_clocks ? k_onchip<256, 2, 1, false, true> : k_onchip<256, 2, 1, false, false>);
  else kern = dir ? k_onchip<256, 1, 2, true, false> : k_onchip<256, 1, 2, false, false>;
  TNO_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (out->J && PL->dp.con.r_fun >= 0)
    launch_fill_funicular(PL, PL->oc.Jfun, (int64_t)2 * PL->dp.m * PL->dp.var.nvar, out->J, N, stream);
  PL->prof_begin(KK_ONCHIP, stream);
  kern<<<grid, threads, smem, stream>>>(PL->dp, PL->oc, A);
  PL->prof_end(stream);
  PL->launches += 1;
  TNO_CUDA_OK(cudaGetLastError());
  if (want_clocks) {
    long long h[128];
    TNO_CUDA_OK(cudaStreamSynchronize(stream));
    TNO_CUDA_OK(cudaMemcpy(h, A.phase_clocks, sizeof(h), cudaMemcpyDeviceToHost));
    cudaFree(A.phase_clocks);
    static const char* names[13] = {"x", "q", "assemble", "factor", "scale", "rhs1", "solve1", "post1", "rhs2", "solve2",
                                    "reac/obj", "emitJ", "final"};
    fprintf(stderr, "[tno smem] total=%d idx=%d ct_in_smem=%d V=%d panel=%d doubles; factor program: %d ops, %d staged chunks of <= %d bytes\n",
            PL->oc.smem_bytes, PL->oc.idx_bytes, PL->oc.ct_in_smem, PL->oc.nv, PL->oc.o_zs - PL->oc.o_r, PL->oc.nops, PL->oc.nchunks,
            PL->oc.fseg_units * 16);
    fprintf(stderr, "[tno phase clocks] N=%lld grid=%d total=%lld :", (long long)N, grid, h[13] - h[0]);
    for (int i = 0; i < 13; ++i) fprintf(stderr, " %s=%lld", names[i], h[i + 1] - h[i]);
    fprintf(stderr, " | factor: dense=%lld (LDL^T part %lld: setup=%lld columns / tile steps=%lld barrier=%lld; inverse: diagonal blocks / tile publish=%lld stages=%lld copy=%lld) chunk waits=%lld",
            h[14], h[15], h[24], h[25], h[26], h[27], h[28], h[29], h[30]);
    fprintf(stderr, " | solve2: pull-wide=%lld pull-chain=%lld Z=%lld diag=%lld Zt=%lld gather=%lld push-wide=%lld", h[16], h[17], h[18],
            h[19], h[20], h[21], h[22]);
    fprintf(stderr, "\n[tno factor ops] start clocks relative to the factor phase:");
    for (int op = 0; op < PL->oc.nops && op < 96; ++op) fprintf(stderr, " %lld", h[32 + op] - h[3]);
    fprintf(stderr, " end=%lld\n", h[4] - h[3]);
  }
  return TNO_OK;
}

}  // namespace tno
