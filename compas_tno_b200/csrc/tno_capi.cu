// C ABI (include/tno_b200.h): plan construction (symbolic analysis + constant upload + emit tables)
// and the batched evaluation entry points.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <chrono>
#include <thread>
#include <vector>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include "tno_plan.hpp"

namespace tno {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

int upload_bytes(Plan* P, const void* host, size_t bytes, const void** dev) {
  *dev = nullptr;
  if (bytes == 0) return TNO_OK;
  void* p = nullptr;
  TNO_CUDA_OK(cudaMalloc(&p, bytes));
  P->dev_allocs.push_back(p);
  TNO_CUDA_OK(cudaMemcpy(p, host, bytes, cudaMemcpyHostToDevice));
  *dev = p;
  return TNO_OK;
}

namespace {

template <class T>
int upload(Plan* P, const std::vector<T>& host, const T** dev) {
  *dev = nullptr;
  if (host.empty()) return TNO_OK;
  void* p = nullptr;
  TNO_CUDA_OK(cudaMalloc(&p, host.size() * sizeof(T)));
  P->dev_allocs.push_back(p);
  TNO_CUDA_OK(cudaMemcpy(p, host.data(), host.size() * sizeof(T), cudaMemcpyHostToDevice));
  *dev = (const T*)p;
  return TNO_OK;
}

template <class T>
std::vector<T> vec(const T* p, size_t n) {
  return p ? std::vector<T>(p, p + n) : std::vector<T>();
}

int upload_table(Plan* P, const std::vector<EmitDesc2>& t, EmitTable* T) {
  const EmitDesc2* d = nullptr;
  int rc = upload(P, t, &d);
  if (rc) return rc;
  T->dev = const_cast<EmitDesc2*>(d);
  T->count = (int64_t)t.size();
  return TNO_OK;
}

EmitDesc2 cst(double c) { return EmitDesc2{-1, -1, 0.f, 0.f, c}; }
EmitDesc2 one(int slot, float sign, double c = 0.0) { return EmitDesc2{slot, -1, sign, 0.f, c}; }
EmitDesc2 two(int a, float sa, int b, float sb) { return EmitDesc2{a, b, sa, sb, 0.0}; }

// dst <- src, count doubles, with non-temporal stores where the ISA has them: the destination (a user's Jacobian
// buffer, hundreds of MB) is written once and not read here, so it should not be pulled through the caches first.
void copy_streaming(double* dst, const double* src, size_t count) {
#if defined(__SSE2__)
  size_t i = 0;
  if ((reinterpret_cast<uintptr_t>(dst) & 15u) && count) { dst[0] = src[0]; i = 1; }
  for (; i + 2 <= count; i += 2)
    _mm_stream_pd(dst + i, _mm_loadu_pd(src + i));
  if (i < count) dst[i] = src[i];
#else
  std::memcpy(dst, src, count * sizeof(double));
#endif
}

int fail(int code, const std::string& msg) {
  set_error(msg);
  return code;
}

int ensure_scratch(Plan* P, int64_t N) {
  // candidates per launch sequence: enough to fill the machine, bounded by a scratch budget
  const int64_t per_cand = (int64_t)P->dp.slot.nslots * 8 + 4;
  const int64_t budget = (int64_t)6 << 30;
  int64_t cmax = std::max<int64_t>(128, (budget / per_cand) / 128 * 128);
  cmax = std::min<int64_t>(cmax, 65536);
  if (const char* cap = getenv("TNO_CHUNK_MAX"))   // test knob: small chunks make a modest batch cross chunk boundaries
    cmax = std::max<int64_t>(128, std::min<int64_t>(cmax, atoll(cap) / 128 * 128));
  int64_t want = std::min<int64_t>(cmax, (N + 127) / 128 * 128);
  if (P->scratch && P->chunk >= want) return TNO_OK;
  if (P->scratch) {
    cudaFree(P->scratch);
    cudaFree(P->status_scratch);
    P->scratch = nullptr;
    P->status_scratch = nullptr;
  }
  P->chunk = want;
  P->scratch_bytes = P->chunk * (int64_t)P->dp.slot.nslots * 8;
  TNO_CUDA_OK(cudaMalloc((void**)&P->scratch, (size_t)P->scratch_bytes));
  TNO_CUDA_OK(cudaMalloc((void**)&P->status_scratch, (size_t)P->chunk * 4));
  P->info.chunk = (int32_t)P->chunk;
  P->info.scratch_bytes = P->scratch_bytes;
  return TNO_OK;
}

}  // namespace
}  // namespace tno

using namespace tno;

extern "C" {

int tno_abi_version(void) { return TNO_ABI_VERSION; }
const char* tno_last_error(void) { return g_error.c_str(); }

void tno_plan_destroy(tno_plan* plan) {
  Plan* P = reinterpret_cast<Plan*>(plan);
  if (!P) return;
  cudaSetDevice(P->device);
  for (void* p : P->dev_allocs) cudaFree(p);
  if (P->scratch) cudaFree(P->scratch);
  if (P->status_scratch) cudaFree(P->status_scratch);
  if (P->d_stage) cudaFree(P->d_stage);
  for (auto& r : P->prof_recs) {
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (auto& e : P->prof_pool) cudaEventDestroy(e);
  if (P->fetch_ev0) cudaEventDestroy(P->fetch_ev0);
  if (P->fetch_ev1) cudaEventDestroy(P->fetch_ev1);
  if (P->stream) cudaStreamDestroy(P->stream);
  delete P;
}

int tno_plan_create(const tno_problem_desc* D, tno_plan** out_plan) {
  if (!D || !out_plan) return fail(TNO_EINVAL, "null argument");
  *out_plan = nullptr;
  if (D->abi_version != TNO_ABI_VERSION) return fail(TNO_EINVAL, "abi_version mismatch");
  const int n = D->n, m = D->m, nb = D->nb, k = D->k;
  if (n <= 0 || m <= 0 || nb <= 0 || nb >= n || k <= 0 || k > m) return fail(TNO_EINVAL, "bad sizes");
  if (!D->edges || !D->fixed || !D->B || !D->d || !D->P || !D->X || !D->qmin || !D->qmax)
    return fail(TNO_EINVAL, "missing mandatory array");
  if (!(D->variables & TNO_VAR_Q)) return fail(TNO_EINVAL, "variables must contain q");
  if (D->variables & ~(TNO_VAR_Q | TNO_VAR_ZB | TNO_VAR_T | TNO_VAR_LAMBDH | TNO_VAR_LAMBDV | TNO_VAR_XYB))
    return fail(TNO_EUNSUPPORTED, "unsupported variable (supported: q, xyb, zb, t/n, lambdh, lambdv)");
  if (D->constraints & ~(TNO_CON_FUNICULAR | TNO_CON_ENVELOPE | TNO_CON_REAC_BOUNDS | TNO_CON_ENVELOPEXY))
    return fail(TNO_EUNSUPPORTED, "unsupported constraint (supported: funicular, envelopexy, envelope, reac_bounds)");
  if ((D->constraints & TNO_CON_ENVELOPEXY) && (!D->xlimits || !D->ylimits)) return fail(TNO_EINVAL, "envelopexy needs xlimits / ylimits");
  if ((D->variables & TNO_VAR_XYB) || (D->constraints & TNO_CON_ENVELOPEXY)) {
    // the reference stacks the xyb columns against the envelopexy AND envelope row blocks (jacobian.py:269-273)
    if ((D->variables & TNO_VAR_XYB) && (D->constraints & (TNO_CON_ENVELOPEXY | TNO_CON_ENVELOPE)) != (TNO_CON_ENVELOPEXY | TNO_CON_ENVELOPE))
      return fail(TNO_EUNSUPPORTED, "xyb variables need the envelopexy and envelope constraints");
    if (D->variables & (TNO_VAR_LAMBDH | TNO_VAR_LAMBDV)) return fail(TNO_EUNSUPPORTED, "load multipliers with xyb / envelopexy are not available");
    if (D->objective > TNO_OBJ_NEG_LAST && D->objective != TNO_OBJ_FEASIBILITY)
      return fail(TNO_EUNSUPPORTED, "xyb / envelopexy support the objectives min, max, t, n, feasibility");
    if (D->kernel == TNO_KERNEL_ONCHIP) return fail(TNO_EUNSUPPORTED, "xyb / envelopexy run on the interleaved kernel family");
  }
  if ((D->constraints & TNO_CON_ENVELOPE) && (!D->lb || !D->ub)) return fail(TNO_EINVAL, "envelope needs lb / ub");
  if ((D->variables & TNO_VAR_T) && D->envelope_kind != TNO_ENV_DOME && D->envelope_kind != TNO_ENV_CROSSVAULT)
    return fail(TNO_EINVAL, "thickness variable needs a parametric envelope");
  if ((D->variables & TNO_VAR_T) && (D->constraints & TNO_CON_REAC_BOUNDS))
    return fail(TNO_EUNSUPPORTED, "reac_bounds with variable thickness is not available (broken on the reference branch too)");
  if ((D->constraints & TNO_CON_REAC_BOUNDS) && !D->b) return fail(TNO_EINVAL, "reac_bounds needs b");
  if ((D->variables & TNO_VAR_LAMBDH) && (!D->px0 || !D->py0)) return fail(TNO_EINVAL, "lambdh needs px0 / py0");
  if ((D->variables & TNO_VAR_LAMBDV) && (!D->pzv || !D->pz0)) return fail(TNO_EINVAL, "lambdv needs pzv / pz0");
  if (D->d_b || D->px0_b || D->py0_b) {
    if (!(D->d_b && D->px0_b && D->py0_b)) return fail(TNO_EINVAL, "second load basis needs d_b, px0_b and py0_b");
    if (!(D->variables & TNO_VAR_LAMBDH)) return fail(TNO_EINVAL, "second load basis needs the lambdh variable");
  }
  if (D->objective < TNO_OBJ_NONE || D->objective > TNO_OBJ_ECOMP_NONLINEAR) return fail(TNO_EINVAL, "bad objective");
  const bool ecomp = D->objective == TNO_OBJ_ECOMP_LINEAR || D->objective == TNO_OBJ_ECOMP_NONLINEAR;
  if (D->objective == TNO_OBJ_BESTFIT || D->objective == TNO_OBJ_LOADPATH || ecomp) {
    // gradient_bestfit / gradient_loadpath / gradient_complementary_energy (derivatives.py:355-413, 640-718, 503-620):
    // of the variables this library supports they only know the q and zb blocks
    if (D->variables & ~(TNO_VAR_Q | TNO_VAR_ZB))
      return fail(TNO_EUNSUPPORTED, "bestfit / loadpath / Ecomp objectives support the variables q and zb only");
    if (D->objective == TNO_OBJ_BESTFIT && !D->s) return fail(TNO_EINVAL, "bestfit objective needs the target surface s");
    if (ecomp && !D->dXb) return fail(TNO_EINVAL, "Ecomp objectives need the support displacements dXb");
    if (D->objective == TNO_OBJ_ECOMP_NONLINEAR && !D->stiff) return fail(TNO_EINVAL, "Ecomp-nonlinear needs stiff");
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(TNO_ENODEVICE, "no CUDA device: this library has no CPU fallback");
  if (D->device < 0 || D->device >= ndev) return fail(TNO_ENODEVICE, "device ordinal out of range");

  Plan* P = new (std::nothrow) Plan();
  if (!P) return fail(TNO_ENOMEM, "out of host memory");
  struct Guard {
    Plan* p;
    ~Guard() {
      if (p) tno_plan_destroy(reinterpret_cast<tno_plan*>(p));
    }
  } guard{P};
  P->device = D->device;
  TNO_CUDA_OK(cudaSetDevice(D->device));
  {
    cudaDeviceProp prop;
    TNO_CUDA_OK(cudaGetDeviceProperties(&prop, D->device));
    P->sm_count = prop.multiProcessorCount;
    if (prop.major < 10) return fail(TNO_ENODEVICE, "device is not sm_100 class");
  }
  TNO_CUDA_OK(cudaStreamCreateWithFlags(&P->stream, cudaStreamNonBlocking));

  // ---- topology ---------------------------------------------------------------------------
  std::vector<int> vrow(n, 0), freepos(n, -1), fixed(D->fixed, D->fixed + nb);
  std::vector<char> is_fixed(n, 0);
  for (int b = 0; b < nb; ++b) {
    if (fixed[b] < 0 || fixed[b] >= n || is_fixed[fixed[b]]) return fail(TNO_EINVAL, "bad fixed list");
    is_fixed[fixed[b]] = 1;
  }
  const int ni = n - nb;
  std::vector<int> free_vertex;
  for (int v = 0; v < n; ++v)
    if (!is_fixed[v]) {
      freepos[v] = (int)free_vertex.size();
      free_vertex.push_back(v);
    }
  std::vector<int> eu(m), ev(m);
  std::vector<std::pair<int, int>> aedges;
  for (int e = 0; e < m; ++e) {
    eu[e] = D->edges[2 * e];
    ev[e] = D->edges[2 * e + 1];
    if (eu[e] < 0 || eu[e] >= n || ev[e] < 0 || ev[e] >= n || eu[e] == ev[e]) return fail(TNO_EINVAL, "bad edge");
    if (!is_fixed[eu[e]] && !is_fixed[ev[e]]) aedges.push_back({freepos[eu[e]], freepos[ev[e]]});
  }
  std::vector<double> coords(2 * (size_t)ni);
  for (int i = 0; i < ni; ++i) {
    coords[2 * i] = D->X[3 * free_vertex[i]];
    coords[2 * i + 1] = D->X[3 * free_vertex[i] + 1];
  }
  Symbolic& S = P->sym;
  try {
    const char* cw = getenv("TNO_CHAIN_WIDTH");  // developer knob: 0 disables the chain partition
    const int chain_width = cw ? atoi(cw) : 4;
    // The analysis depends on the topology only (the coordinates enter the nested-dissection ordering alone): plans of the
    // same form diagram -- one per sampled plan geometry, per objective, per rank-local replica -- reuse it from a small
    // process-wide cache (21 ms at disc 20, 240 ms at disc 40 otherwise).
    struct SymKey {
      int ni, ordering, chain_width;
      std::vector<std::pair<int, int>> edges;
      bool operator==(const SymKey& o) const { return ni == o.ni && ordering == o.ordering && chain_width == o.chain_width && edges == o.edges; }
    };
    static std::mutex sym_mutex;
    static std::vector<std::pair<SymKey, std::shared_ptr<const Symbolic>>> sym_cache;
    static const bool no_cache = getenv("TNO_NO_SYMBOLIC_CACHE") != nullptr || getenv("TNO_CHAIN_MAX") || getenv("TNO_CHAIN_MERGE");
    const bool cacheable = !no_cache && D->ordering != ORD_NESTED;
    std::shared_ptr<const Symbolic> hit;
    SymKey key{ni, D->ordering, chain_width, aedges};
    if (cacheable) {
      std::lock_guard<std::mutex> lock(sym_mutex);
      for (auto& kv : sym_cache)
        if (kv.first == key) hit = kv.second;
    }
    if (hit) {
      S = *hit;
    } else {
      analyse(ni, aedges, D->ordering, coords.data(), true, true, &S, chain_width);
      if (cacheable) {
        std::lock_guard<std::mutex> lock(sym_mutex);
        if (sym_cache.size() >= 8) sym_cache.erase(sym_cache.begin());
        sym_cache.emplace_back(std::move(key), std::make_shared<const Symbolic>(S));
      }
    }
  } catch (const std::exception& ex) {
    return fail(TNO_EINVAL, std::string("symbolic analysis failed: ") + ex.what());
  }
  const int nnz_l = (int)S.rowidx.size();
  std::vector<int> row_vertex(ni);
  for (int r = 0; r < ni; ++r) row_vertex[r] = free_vertex[S.perm[r]];
  for (int v = 0; v < n; ++v) vrow[v] = is_fixed[v] ? 0 : S.iperm[freepos[v]];
  for (int b = 0; b < nb; ++b) vrow[fixed[b]] = -1 - b;

  // vertex -> incident edges
  std::vector<int> vadj_ptr(n + 1, 0), vadj_edge(2 * (size_t)m), vadj_other(2 * (size_t)m);
  std::vector<float> vadj_sign(2 * (size_t)m);
  for (int e = 0; e < m; ++e) {
    vadj_ptr[eu[e] + 1]++;
    vadj_ptr[ev[e] + 1]++;
  }
  for (int v = 0; v < n; ++v) vadj_ptr[v + 1] += vadj_ptr[v];
  {
    std::vector<int> fill(vadj_ptr.begin(), vadj_ptr.end() - 1);
    for (int e = 0; e < m; ++e) {
      int t = fill[eu[e]]++;
      vadj_edge[t] = e;
      vadj_other[t] = ev[e];
      vadj_sign[t] = -1.f;
      t = fill[ev[e]]++;
      vadj_edge[t] = e;
      vadj_other[t] = eu[e];
      vadj_sign[t] = +1.f;
    }
  }
  // assembly gather lists: -(Ci^T Q Ci):  diag_i = -sum q_e,  offdiag_ij = +q_e
  std::vector<std::vector<std::pair<int, float>>> contrib(nnz_l);
  for (int e = 0; e < m; ++e) {
    const int ru = vrow[eu[e]], rv = vrow[ev[e]];
    if (ru >= 0) contrib[S.colptr[ru]].push_back({e, -1.f});
    if (rv >= 0) contrib[S.colptr[rv]].push_back({e, -1.f});
    if (ru >= 0 && rv >= 0) {
      const int p = S.pos(std::max(ru, rv), std::min(ru, rv));
      if (p < 0) return fail(TNO_EINVAL, "internal: edge missing from pattern");
      contrib[p].push_back({e, +1.f});
    }
  }
  std::vector<int> asm_ptr(nnz_l + 1, 0), asm_edge;
  std::vector<float> asm_coef;
  for (int p = 0; p < nnz_l; ++p) {
    for (auto& c : contrib[p]) {
      asm_edge.push_back(c.first);
      asm_coef.push_back(c.second);
    }
    asm_ptr[p + 1] = (int)asm_edge.size();
  }

  P->h_vrow = vrow;
  P->h_vadj_ptr = vadj_ptr;
  P->h_vadj_edge = vadj_edge;
  P->h_vadj_other = vadj_other;
  P->h_vadj_sign = vadj_sign;
  P->h_fixed = fixed;

  // ---- layouts ----------------------------------------------------------------------------
  DevPlan& dp = P->dp;
  dp.n = n; dp.m = m; dp.ni = ni; dp.nb = nb; dp.k = k;
  dp.variables = D->variables; dp.constraints = D->constraints;
  dp.objective = D->objective; dp.envelope_kind = D->envelope_kind;
  for (int i = 0; i < 8; ++i) dp.env[i] = D->env_params[i];
  dp.nnz_l = nnz_l;
  VarLayout& V = dp.var;
  int pos = k;
  if (D->variables & TNO_VAR_XYB) { V.o_xyb = pos; pos += 2 * nb; }
  if (D->variables & TNO_VAR_ZB) { V.o_zb = pos; pos += nb; }
  if (D->variables & TNO_VAR_T) { V.o_t = pos; pos += 1; }
  if (D->variables & TNO_VAR_LAMBDH) { V.o_lambdh = pos; pos += 1; }
  if (D->variables & TNO_VAR_LAMBDV) { V.o_lambdv = pos; pos += 1; }
  V.nvar = pos;
  ConLayout& C = dp.con;
  pos = 0;
  if (D->constraints & TNO_CON_FUNICULAR) { C.r_fun = pos; pos += 2 * m; }
  if (D->constraints & TNO_CON_ENVELOPEXY) { C.r_envxy = pos; pos += 4 * n; }
  if (D->constraints & TNO_CON_ENVELOPE) { C.r_env = pos; pos += 2 * n; }
  if (D->constraints & TNO_CON_REAC_BOUNDS) { C.r_reac = pos; pos += 2 * nb; }
  C.ng = pos;
  RhsLayout& R = dp.rhs;
  pos = 3;
  if (V.o_zb >= 0 || V.o_xyb >= 0) { R.c_zb = pos; pos += nb; }   // A = A^-1 (-Ci^T Q Cb): dz/dzb = dx/dxb = dy/dyb
  if (V.o_lambdv >= 0) { R.c_lambdv = pos; pos += 1; }
  R.nrhs1 = pos;
  const bool need_dq = (D->constraints & (TNO_CON_ENVELOPE | TNO_CON_REAC_BOUNDS)) != 0 || D->objective == TNO_OBJ_BESTFIT ||
                       D->objective == TNO_OBJ_LOADPATH || ecomp;
  if (need_dq) {
    R.c_q = pos; pos += k;
    if (V.o_lambdh >= 0) { R.c_lambdh = pos; pos += 1; }
  }
  const bool thrust = D->objective == TNO_OBJ_MIN || D->objective == TNO_OBJ_MAX;
  if ((D->constraints & TNO_CON_ENVELOPEXY) || (V.o_xyb >= 0 && thrust)) {
    R.c_qx = pos; pos += k;
    R.c_qy = pos; pos += k;
  }
  R.nrhs = pos;
  R.nrhs2 = R.nrhs - R.nrhs1;
  SlotLayout& SL = dp.slot;
  pos = 0;
  SL.s_x = pos; pos += V.nvar;
  SL.s_pzs = pos; pos += 1;
  SL.s_dir = pos; pos += 2;
  SL.s_q = pos; pos += m;
  SL.s_l = pos; pos += nnz_l;
  SL.s_r = pos; pos += ni * R.nrhs;
  SL.s_env = pos; pos += (V.o_t >= 0) ? 4 * n : 0;
  SL.s_grad = pos; pos += V.nvar;
  SL.s_rg = pos; pos += 2 * nb;
  SL.s_rj = pos; pos += (D->constraints & TNO_CON_REAC_BOUNDS) ? 2 * nb * V.nvar : 0;
  SL.s_cxy = pos; pos += 2 * nb;
  SL.nslots = pos;

  // ---- uploads ----------------------------------------------------------------------------
  int rc;
#define UP(field, hostvec)                    \
  if ((rc = upload(P, hostvec, &dp.field))) return rc;
  UP(colptr, S.colptr);
  UP(rowidx, S.rowidx);
  UP(upd_ptr, S.upd_ptr);
  UP(upd_pos, S.upd_pos);
  UP(asm_ptr, asm_ptr);
  UP(asm_edge, asm_edge);
  UP(asm_coef, asm_coef);
  UP(edge_u, eu);
  UP(edge_v, ev);
  UP(vrow, vrow);
  UP(row_vertex, row_vertex);
  UP(fixed, fixed);
  UP(vadj_ptr, vadj_ptr);
  UP(vadj_edge, vadj_edge);
  UP(vadj_other, vadj_other);
  UP(vadj_sign, vadj_sign);
  UP(B, vec(D->B, (size_t)m * k));
  UP(d, vec(D->d, (size_t)m));
  UP(P, vec(D->P, (size_t)n * 3));
  UP(X, vec(D->X, (size_t)n * 3));
  UP(lb, vec(D->lb, (size_t)n));
  UP(ub, vec(D->ub, (size_t)n));
  UP(s, vec(D->s, (size_t)n));
  UP(qmin, vec(D->qmin, (size_t)m));
  UP(qmax, vec(D->qmax, (size_t)m));
  UP(b, vec(D->b, (size_t)nb * 2));
  UP(px0, vec(D->px0, (size_t)n));
  UP(py0, vec(D->py0, (size_t)n));
  UP(pzv, vec(D->pzv, (size_t)n));
  UP(pz0, vec(D->pz0, (size_t)n));
  UP(dXb, vec(D->dXb, (size_t)nb * 3));
  UP(stiff, vec(D->stiff, (size_t)m));
  UP(d_b, vec(D->d_b, (size_t)m));
  UP(px0_b, vec(D->px0_b, (size_t)n));
  UP(py0_b, vec(D->py0_b, (size_t)n));
  UP(xlimits, vec(D->xlimits, (size_t)n * 2));
  UP(ylimits, vec(D->ylimits, (size_t)n * 2));
#undef UP

  // ---- emit tables (user-facing row-major layouts as affine gathers of the scratch slots) --
  const int nvar = V.nvar, ng = C.ng;
  auto zdesc = [&](int v, float sign) -> EmitDesc2 {
    if (vrow[v] >= 0) return one(SL.s_r + vrow[v] * R.nrhs + R.cz, sign);
    const int b = -1 - vrow[v];
    if (V.o_zb >= 0) return one(SL.s_x + V.o_zb + b, sign);
    return cst(sign * D->X[3 * v + 2]);
  };
  auto xydesc = [&](int v, int axis, float sign) -> EmitDesc2 {
    if (vrow[v] >= 0) return one(SL.s_r + vrow[v] * R.nrhs + (axis == 0 ? R.cx : R.cy), sign);
    const int b = -1 - vrow[v];
    if (V.o_xyb >= 0) return one(SL.s_x + V.o_xyb + axis * nb + b, sign);
    return cst(sign * D->X[3 * v + axis]);
  };
  {
    std::vector<EmitDesc2> t((size_t)ng);
    if (C.r_fun >= 0)
      for (int e = 0; e < m; ++e) {
        t[C.r_fun + e] = one(SL.s_q + e, +1.f, -D->qmin[e]);
        t[C.r_fun + m + e] = one(SL.s_q + e, -1.f, +D->qmax[e]);
      }
    if (C.r_envxy >= 0)
      for (int v = 0; v < n; ++v)
        for (int axis = 0; axis < 2; ++axis) {
          const double* lim = axis == 0 ? D->xlimits : D->ylimits;
          EmitDesc2 lo = xydesc(v, axis, +1.f), hi = xydesc(v, axis, -1.f);
          lo.c += -lim[2 * v + 0];
          hi.c += +lim[2 * v + 1];
          t[C.r_envxy + 2 * axis * n + v] = lo;
          t[C.r_envxy + (2 * axis + 1) * n + v] = hi;
        }
    if (C.r_env >= 0)
      for (int v = 0; v < n; ++v) {
        EmitDesc2 lo = zdesc(v, +1.f), hi = zdesc(v, -1.f);
        if (V.o_t >= 0) {
          lo.slot_b = SL.s_env + n + v; lo.sign_b = -1.f;   // - lb(t)
          hi.slot_b = SL.s_env + v;     hi.sign_b = +1.f;   // + ub(t)
        } else {
          lo.c += -D->lb[v];
          hi.c += +D->ub[v];
        }
        if (lo.slot_a < 0 && lo.slot_b >= 0) { lo.slot_a = lo.slot_b; lo.sign_a = lo.sign_b; lo.slot_b = -1; }
        if (hi.slot_a < 0 && hi.slot_b >= 0) { hi.slot_a = hi.slot_b; hi.sign_a = hi.sign_b; hi.slot_b = -1; }
        t[C.r_env + v] = lo;
        t[C.r_env + n + v] = hi;
      }
    if (C.r_reac >= 0)
      for (int i = 0; i < 2 * nb; ++i) t[C.r_reac + i] = one(SL.s_rg + i, +1.f);
    if ((rc = upload_table(P, t, &P->t_g))) return rc;
  }
  {
    std::vector<EmitDesc2> t((size_t)nvar, cst(0.0));
    if (D->objective == TNO_OBJ_MIN || D->objective == TNO_OBJ_MAX) {
      for (int j = 0; j < k; ++j) t[j] = one(SL.s_grad + j, +1.f);
      if (V.o_xyb >= 0)
        for (int j = 0; j < 2 * nb; ++j) t[V.o_xyb + j] = one(SL.s_grad + V.o_xyb + j, +1.f);
    }
    else if (D->objective == TNO_OBJ_BESTFIT || D->objective == TNO_OBJ_LOADPATH || ecomp)
      for (int j = 0; j < nvar; ++j) t[j] = one(SL.s_grad + j, +1.f);
    else if (D->objective == TNO_OBJ_T)
      t[nvar - 1] = cst(1.0);
    else if (D->objective == TNO_OBJ_NEG_LAST)
      t[nvar - 1] = cst(-1.0);
    if ((rc = upload_table(P, t, &P->t_grad))) return rc;
  }
  {
    std::vector<EmitDesc2> t((size_t)n);
    for (int v = 0; v < n; ++v) t[v] = zdesc(v, +1.f);
    if ((rc = upload_table(P, t, &P->t_z))) return rc;
    std::vector<EmitDesc2> tq((size_t)m);
    for (int e = 0; e < m; ++e) tq[e] = one(SL.s_q + e, +1.f);
    if ((rc = upload_table(P, tq, &P->t_q))) return rc;
  }
  {
    std::vector<EmitDesc2> t((size_t)ng * nvar, cst(0.0));
    auto at = [&](int row, int col) -> EmitDesc2& { return t[(size_t)row * nvar + col]; };
    if (C.r_fun >= 0)
      for (int e = 0; e < m; ++e) {
        for (int j = 0; j < k; ++j) {
          at(C.r_fun + e, j) = cst(D->B[(size_t)e * k + j]);
          at(C.r_fun + m + e, j) = cst(-D->B[(size_t)e * k + j]);
        }
        if (V.o_lambdh >= 0) {
          at(C.r_fun + e, V.o_lambdh) = cst(D->d[e]);
          at(C.r_fun + m + e, V.o_lambdh) = cst(-D->d[e]);
        }
      }
    if (C.r_env >= 0)
      for (int v = 0; v < n; ++v) {
        const int r1 = C.r_env + v, r2 = C.r_env + n + v;
        if (vrow[v] >= 0) {
          const int base = SL.s_r + vrow[v] * R.nrhs;
          for (int j = 0; j < k; ++j) {
            at(r1, j) = one(base + R.c_q + j, +1.f);
            at(r2, j) = one(base + R.c_q + j, -1.f);
          }
          if (V.o_zb >= 0)
            for (int b = 0; b < nb; ++b) {
              at(r1, V.o_zb + b) = one(base + R.c_zb + b, +1.f);
              at(r2, V.o_zb + b) = one(base + R.c_zb + b, -1.f);
            }
          if (V.o_lambdh >= 0) {
            at(r1, V.o_lambdh) = one(base + R.c_lambdh, +1.f);
            at(r2, V.o_lambdh) = one(base + R.c_lambdh, -1.f);
          }
          if (V.o_lambdv >= 0) {
            at(r1, V.o_lambdv) = one(base + R.c_lambdv, +1.f);
            at(r2, V.o_lambdv) = one(base + R.c_lambdv, -1.f);
          }
        } else if (V.o_zb >= 0) {
          const int b = -1 - vrow[v];
          at(r1, V.o_zb + b) = cst(1.0);
          at(r2, V.o_zb + b) = cst(-1.0);
        }
        if (V.o_t >= 0) {
          at(r1, V.o_t) = one(SL.s_env + 3 * n + v, -1.f);  // -dlb/dt
          at(r2, V.o_t) = one(SL.s_env + 2 * n + v, +1.f);  // +dub/dt
        }
      }
    if (C.r_envxy >= 0)   // rows [dx; -dx; dy; -dy]: q columns from the Dx / Dy panels, xb (yb) columns = A (jacobian.py:173-186,269-273)
      for (int v = 0; v < n; ++v)
        for (int axis = 0; axis < 2; ++axis) {
          const int r1 = C.r_envxy + 2 * axis * n + v, r2 = r1 + n;
          const int ocol = V.o_xyb >= 0 ? V.o_xyb + axis * nb : -1;
          if (vrow[v] >= 0) {
            const int base = SL.s_r + vrow[v] * R.nrhs;
            const int cq = axis == 0 ? R.c_qx : R.c_qy;
            for (int j = 0; j < k; ++j) {
              at(r1, j) = one(base + cq + j, +1.f);
              at(r2, j) = one(base + cq + j, -1.f);
            }
            if (ocol >= 0)
              for (int b = 0; b < nb; ++b) {
                at(r1, ocol + b) = one(base + R.c_zb + b, +1.f);
                at(r2, ocol + b) = one(base + R.c_zb + b, -1.f);
              }
          } else if (ocol >= 0) {
            const int b = -1 - vrow[v];
            at(r1, ocol + b) = cst(1.0);
            at(r2, ocol + b) = cst(-1.0);
          }
        }
    if (C.r_reac >= 0)
      for (int i = 0; i < 2 * nb; ++i)
        for (int j = 0; j < nvar; ++j) at(C.r_reac + i, j) = one(SL.s_rj + i * nvar + j, +1.f);
    if (C.r_fun == 0 && m > 0) {
      // the funicular rows are constants: they leave the table and are replicated by a streaming kernel
      const size_t cnt = (size_t)2 * m * nvar;
      std::vector<double> blockv(cnt);
      for (size_t i = 0; i < cnt; ++i) blockv[i] = t[i].c;
      if ((rc = upload(P, blockv, &P->jfun))) return rc;
      P->jfun_count = (int64_t)cnt;
      P->h_jfun = blockv;
      t.erase(t.begin(), t.begin() + cnt);
    }
    if ((rc = upload_table(P, t, &P->t_J))) return rc;
  }

  // ---- info -------------------------------------------------------------------------------
  tno_plan_info& I = P->info;
  I.n = n; I.m = m; I.ni = ni; I.nb = nb; I.k = k;
  I.nvar = nvar; I.ng = ng; I.nrhs = R.nrhs;
  I.nnz_a = S.nnz_a; I.nnz_l = nnz_l;
  I.etree_height = S.nlevels; I.max_colcount = S.max_colcount;
  I.factor_flops = S.factor_flops;
  I.solve_flops = (int64_t)4 * nnz_l * R.nrhs;
  P->kernel = TNO_KERNEL_INTERLEAVED;
  I.smem_bytes = 0;
  P->thread_per_candidate = D->kernel == TNO_KERNEL_INTERLEAVED_TPC;
  if (D->kernel != TNO_KERNEL_INTERLEAVED && D->kernel != TNO_KERNEL_INTERLEAVED_TPC) {
    if ((rc = build_onchip(P, D))) return rc;
    if (P->onchip_ok) {
      P->kernel = TNO_KERNEL_ONCHIP;
      P->auto_family = D->kernel == TNO_KERNEL_AUTO;
      I.smem_bytes = (int64_t)P->oc.smem_bytes;
      I.onchip_ctas_per_sm = P->onchip_ctas_per_sm;
    } else if (D->kernel == TNO_KERNEL_ONCHIP) {
      return fail(TNO_EUNSUPPORTED, std::string("problem does not fit the on-chip kernel family: ") + tno_last_error());
    }
  }
  I.kernel = P->kernel;

  guard.p = nullptr;
  *out_plan = reinterpret_cast<tno_plan*>(P);
  return TNO_OK;
}

int tno_plan_query(const tno_plan* plan, tno_plan_info* info) {
  if (!plan || !info) return fail(TNO_EINVAL, "null argument");
  *info = reinterpret_cast<const Plan*>(plan)->info;
  return TNO_OK;
}

int tno_plan_symbolic(const tno_plan* plan, int32_t* perm, int32_t* colptr, int32_t* rowidx, int32_t* parent,
                      int32_t* level) {
  if (!plan) return fail(TNO_EINVAL, "null plan");
  const Symbolic& S = reinterpret_cast<const Plan*>(plan)->sym;
  if (perm) std::copy(S.perm.begin(), S.perm.end(), perm);
  if (colptr) std::copy(S.colptr.begin(), S.colptr.end(), colptr);
  if (rowidx) std::copy(S.rowidx.begin(), S.rowidx.end(), rowidx);
  if (parent) std::copy(S.parent.begin(), S.parent.end(), parent);
  if (level) std::copy(S.level.begin(), S.level.end(), level);
  return TNO_OK;
}

int tno_symbolic_analyse(int32_t n_free, int32_t n_edges, const int32_t* edges, int32_t ordering, int32_t chain_width,
                         int32_t* perm, int32_t* colptr, int32_t* rowidx, int32_t rowidx_capacity, int32_t* parent,
                         int32_t* chain_ptr, int32_t chain_capacity, int32_t* info) {
  if (n_free <= 0 || n_edges < 0 || (n_edges > 0 && !edges) || !info) return fail(TNO_EINVAL, "bad argument");
  std::vector<std::pair<int, int>> e;
  for (int i = 0; i < n_edges; ++i) {
    const int a = edges[2 * i], b = edges[2 * i + 1];
    if (a < 0 || a >= n_free || b < 0 || b >= n_free) return fail(TNO_EINVAL, "edge endpoint out of range");
    e.push_back({a, b});
  }
  Symbolic S;
  try {
    analyse(n_free, e, ordering, nullptr, true, true, &S, chain_width);
  } catch (const std::exception& ex) {
    return fail(TNO_EINVAL, std::string("symbolic analysis failed: ") + ex.what());
  }
  const int nnz = (int)S.rowidx.size();
  const int nch = (int)S.chain_clevel.size();
  info[0] = nnz;
  info[1] = S.nlevels;
  info[2] = S.max_colcount;
  info[3] = S.wide_levels;
  info[4] = S.wide_rows;
  info[5] = nch;
  info[6] = (int)S.clevel_ptr.size() - 1;
  info[7] = (int)S.factor_flops;
  if (perm) std::copy(S.perm.begin(), S.perm.end(), perm);
  if (colptr) std::copy(S.colptr.begin(), S.colptr.end(), colptr);
  if (parent) std::copy(S.parent.begin(), S.parent.end(), parent);
  if (rowidx) {
    if (rowidx_capacity < nnz) return fail(TNO_EINVAL, "rowidx capacity too small");
    std::copy(S.rowidx.begin(), S.rowidx.end(), rowidx);
  }
  if (chain_ptr) {
    if (chain_capacity < nch + 1) return fail(TNO_EINVAL, "chain capacity too small");
    std::copy(S.chain_ptr.begin(), S.chain_ptr.end(), chain_ptr);
  }
  return TNO_OK;
}

int64_t tno_plan_launch_count(const tno_plan* plan) { return plan ? reinterpret_cast<const Plan*>(plan)->launches : 0; }

int tno_plan_profile(tno_plan* plan, int enable) {
  Plan* P = reinterpret_cast<Plan*>(plan);
  if (!P) return fail(TNO_EINVAL, "null plan");
  for (auto& r : P->prof_recs) {
    P->prof_pool.push_back(r.a);
    P->prof_pool.push_back(r.b);
  }
  P->prof_recs.clear();
  for (int i = 0; i < KK_COUNT; ++i) {
    P->prof_ms[i] = 0.0;
    P->prof_launches[i] = 0;
  }
  P->profiling = enable != 0;
  return TNO_OK;
}

int tno_plan_profile_read(tno_plan* plan, int32_t max_kinds, char* names, double* total_ms, int64_t* launches) {
  Plan* P = reinterpret_cast<Plan*>(plan);
  if (!P || !names || !total_ms || !launches) return fail(TNO_EINVAL, "null argument");
  TNO_CUDA_OK(cudaSetDevice(P->device));
  for (auto& r : P->prof_recs) {
    TNO_CUDA_OK(cudaEventSynchronize(r.b));
    float ms = 0.f;
    TNO_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
    P->prof_ms[r.kind] += ms;
    P->prof_launches[r.kind] += 1;
    P->prof_pool.push_back(r.a);
    P->prof_pool.push_back(r.b);
  }
  P->prof_recs.clear();
  const int nk = std::min<int>(max_kinds, KK_COUNT);
  for (int i = 0; i < nk; ++i) {
    std::strncpy(names + 32 * i, kKernelNames[i], 31);
    names[32 * i + 31] = 0;
    total_ms[i] = P->prof_ms[i];
    launches[i] = P->prof_launches[i];
  }
  return nk;
}

int tno_eval_batch(tno_plan* plan, const double* x_dev, int64_t N, int64_t ldx, const tno_batch_inputs* opt,
                   const tno_outputs* out_dev, void* cuda_stream) {
  Plan* P = reinterpret_cast<Plan*>(plan);
  if (!P || !x_dev || !out_dev) return fail(TNO_EINVAL, "null argument");
  if (N < 0 || ldx < P->dp.var.nvar) return fail(TNO_EINVAL, "bad N / ldx");
  if (N == 0) return TNO_OK;
  // the Jacobian is written with 16-byte stores; ng is even, so every candidate's block is aligned when the base is
  if (out_dev->J && (reinterpret_cast<uintptr_t>(out_dev->J) & 15u)) return fail(TNO_EINVAL, "out->J must be 16-byte aligned");
  TNO_CUDA_OK(cudaSetDevice(P->device));
  bool onchip = P->kernel == TNO_KERNEL_ONCHIP;
  if (onchip && P->auto_family && P->onchip_ctas_per_sm == 1) {
    // one resident candidate per SM: beyond ~6k candidates the level-scheduled interleaved family is faster
    // (dome 16x20, B200: 12.4 vs 10.5 ms at 8192, 24.5 vs 17.8 ms at 16384; 6.2 vs 7.2 ms at 4096)
    static const char* env = getenv("TNO_AUTO_SWITCH_N");
    const int64_t thr = env ? atoll(env) : 6144;
    if (N >= thr) onchip = false;
  }
  P->info.last_kernel = onchip ? TNO_KERNEL_ONCHIP : TNO_KERNEL_INTERLEAVED;
  if (onchip) return run_onchip(P, x_dev, N, ldx, opt, out_dev, (cudaStream_t)cuda_stream);
  int rc = ensure_scratch(P, N);
  if (rc) return rc;
  return run_interleaved(P, x_dev, N, ldx, opt, out_dev, (cudaStream_t)cuda_stream);
}

int tno_eval_batch_host(tno_plan* plan, const double* x_host, int64_t N, int64_t ldx, const tno_batch_inputs* in_host,
                        const tno_outputs* oh) {
  Plan* P = reinterpret_cast<Plan*>(plan);
  if (!P || !x_host || !oh) return fail(TNO_EINVAL, "null argument");
  const int nvar = P->dp.var.nvar, ng = P->dp.con.ng, n = P->dp.n, m = P->dp.m;
  if (N < 0 || ldx < nvar) return fail(TNO_EINVAL, "bad N / ldx");
  if (N == 0) return TNO_OK;
  TNO_CUDA_OK(cudaSetDevice(P->device));
  // device staging: x | pzs | f | grad | g | J | z | q | gmin | status
  size_t off = 0;
  auto take = [&](bool want, size_t count) {
    size_t o = off;
    if (want) off += (count + 31) / 32 * 32;
    return o;
  };
  const size_t o_x = take(true, (size_t)N * ldx);
  const double* pz_scale_host = in_host ? in_host->pz_scale : nullptr;
  const double* load_dir_host = in_host ? in_host->load_dir : nullptr;
  const size_t o_pzs = take(pz_scale_host != nullptr, (size_t)N);
  const size_t o_dir = take(load_dir_host != nullptr, (size_t)N * 2);
  const double* bounds_host = in_host ? in_host->bounds : nullptr;
  const size_t o_bnd = take(bounds_host != nullptr, (size_t)N * 2 * n);
  const size_t o_f = take(oh->f != nullptr, (size_t)N);
  const size_t o_grad = take(oh->grad != nullptr, (size_t)N * nvar);
  const size_t o_g = take(oh->g != nullptr, (size_t)N * ng);
  const size_t o_J = take(oh->J != nullptr, (size_t)N * ng * nvar);
  const size_t o_z = take(oh->z != nullptr, (size_t)N * n);
  const size_t o_q = take(oh->q != nullptr, (size_t)N * m);
  const size_t o_gmin = take(oh->gmin != nullptr, (size_t)N);
  const size_t o_st = take(oh->status != nullptr, ((size_t)N + 1) / 2);
  const size_t bytes = off * 8;
  if (bytes > P->d_stage_bytes) {
    if (P->d_stage) cudaFree(P->d_stage);
    P->d_stage = nullptr;
    P->d_stage_bytes = 0;
    TNO_CUDA_OK(cudaMalloc((void**)&P->d_stage, bytes));
    P->d_stage_bytes = bytes;
  }
  double* base = P->d_stage;
  cudaStream_t s = P->stream;
  TNO_CUDA_OK(cudaMemcpyAsync(base + o_x, x_host, (size_t)N * ldx * 8, cudaMemcpyHostToDevice, s));
  tno_batch_inputs opt{};
  if (pz_scale_host) {
    TNO_CUDA_OK(cudaMemcpyAsync(base + o_pzs, pz_scale_host, (size_t)N * 8, cudaMemcpyHostToDevice, s));
    opt.pz_scale = base + o_pzs;
  }
  if (load_dir_host) {
    TNO_CUDA_OK(cudaMemcpyAsync(base + o_dir, load_dir_host, (size_t)N * 16, cudaMemcpyHostToDevice, s));
    opt.load_dir = base + o_dir;
  }
  if (bounds_host) {
    TNO_CUDA_OK(cudaMemcpyAsync(base + o_bnd, bounds_host, (size_t)N * 2 * n * 8, cudaMemcpyHostToDevice, s));
    opt.bounds = base + o_bnd;
  }
  tno_outputs od{};
  od.f = oh->f ? base + o_f : nullptr;
  od.grad = oh->grad ? base + o_grad : nullptr;
  od.g = oh->g ? base + o_g : nullptr;
  od.J = oh->J ? base + o_J : nullptr;
  od.z = oh->z ? base + o_z : nullptr;
  od.q = oh->q ? base + o_q : nullptr;
  od.gmin = oh->gmin ? base + o_gmin : nullptr;
  od.status = oh->status ? reinterpret_cast<int32_t*>(base + o_st) : nullptr;
  int rc = tno_eval_batch(plan, base + o_x, N, ldx, &opt, &od, (void*)s);
  if (rc) return rc;
#define BACK(field, count, elt)                                                                              \
  if (oh->field) TNO_CUDA_OK(cudaMemcpyAsync(oh->field, od.field, (size_t)(count) * (elt), cudaMemcpyDeviceToHost, s));
  BACK(f, N, 8);
  BACK(grad, N * nvar, 8);
  BACK(g, N * ng, 8);
  if (oh->J) {
    // constant funicular rows are written by host threads, the rest crosses PCIe (not with per-candidate directions:
    // the lambdh column of those rows then differs per candidate)
    if (load_dir_host) {
      TNO_CUDA_OK(cudaMemcpyAsync(oh->J, od.J, (size_t)N * ng * nvar * 8, cudaMemcpyDeviceToHost, s));
    } else {
      rc = tno_fetch_jacobian_host(plan, od.J, oh->J, N, (void*)s);
      if (rc) return rc;
    }
  }
  BACK(z, N * n, 8);
  BACK(q, N * m, 8);
  BACK(gmin, N, 8);
  BACK(status, N, 4);
#undef BACK
  TNO_CUDA_OK(cudaStreamSynchronize(s));
  return TNO_OK;
}

int tno_fetch_jacobian_host(tno_plan* plan, const double* J_dev, double* J_host, int64_t N, void* cuda_stream) {
  Plan* P = reinterpret_cast<Plan*>(plan);
  if (!P || !J_dev || !J_host || N < 0) return fail(TNO_EINVAL, "null argument");
  if (N == 0) return TNO_OK;
  TNO_CUDA_OK(cudaSetDevice(P->device));
  cudaStream_t s = (cudaStream_t)cuda_stream;
  const size_t per = (size_t)P->dp.con.ng * P->dp.var.nvar;           // doubles per candidate
  const size_t cst = P->h_jfun.size();                                 // leading constant doubles per candidate
  if (cst == 0 || cst >= per) {
    TNO_CUDA_OK(cudaMemcpyAsync(J_host, J_dev, (size_t)N * per * 8, cudaMemcpyDeviceToHost, s));
    return TNO_OK;
  }
  // up to 16 host threads, or TNO_HOST_THREADS (one process per GPU on a shared host: cores / ranks)
  static const char* env_threads = getenv("TNO_HOST_THREADS");
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int64_t cap = env_threads ? std::max(1, atoi(env_threads)) : 16;
  const int64_t nthreads = std::max<int64_t>(1, std::min<int64_t>({(int64_t)hw, cap, N / 64 + 1}));
  // Split of the constant rows between the host threads and the DMA engine.  The device copy of J holds them too
  // (k_fill_funicular), so any suffix of the constant block can ride along with the candidate-dependent rows.  With
  // D = PCIe rate and H = host fill rate (~4.5 GB/s per streaming thread), both finish together when the host takes
  //   phi = (v + c) / (c (1 + D / H))     of the c constant bytes (v = candidate-dependent bytes per candidate).
  // Few threads per rank (8 ranks sharing one host) => small phi, the DMA engines of the 8 GPUs carry the load.
  // The model only starts the split: every fetch measures the DMA rate (events around the copy, read at the next call) and
  // the host fill rate (wall clock of the threads), both while the other side is running, and the next call uses them.
  static const char* env_phi = getenv("TNO_HOST_FILL_FRACTION");
  static const bool adaptive = getenv("TNO_HOST_FILL_STATIC") == nullptr;
  if (P->fetch_pending && cudaEventQuery(P->fetch_ev1) == cudaSuccess) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, P->fetch_ev0, P->fetch_ev1) == cudaSuccess && ms > 0.f) {
      const double rate = P->fetch_pending_bytes / (1e-3 * (double)ms);
      P->fetch_rate_d = P->fetch_rate_d > 0.0 ? 0.5 * (P->fetch_rate_d + rate) : rate;
    }
    P->fetch_pending = false;
  }
  double phi = (double)per / ((double)cst * (1.0 + 52.0 / (4.5 * (double)nthreads)));
  if (adaptive && P->fetch_rate_h > 0.0 && P->fetch_rate_d > 0.0 && N >= 256)
    phi = (double)per / ((double)cst * (1.0 + P->fetch_rate_d / P->fetch_rate_h));
  if (env_phi) phi = atof(env_phi);
  phi = std::min(1.0, std::max(0.0, phi));
  P->fetch_phi = phi;
  P->fetch_threads = (int)nthreads;
  const size_t row = (size_t)P->dp.var.nvar;
  size_t rows_h = std::min(cst / row, (size_t)(phi * (double)(cst / row) + 0.5));   // whole rows of J
  if ((rows_h * row) & 1) rows_h -= 1;                                              // the DMA part starts 16-byte aligned
  const size_t cst_h = rows_h * row;
  const bool measure = adaptive && !P->fetch_pending && N >= 256;
  if (measure) {
    if (!P->fetch_ev0) {
      TNO_CUDA_OK(cudaEventCreate(&P->fetch_ev0));
      TNO_CUDA_OK(cudaEventCreate(&P->fetch_ev1));
    }
    TNO_CUDA_OK(cudaEventRecord(P->fetch_ev0, s));
  }
  TNO_CUDA_OK(cudaMemcpy2DAsync(J_host + cst_h, per * 8, J_dev + cst_h, per * 8, (per - cst_h) * 8, (size_t)N, cudaMemcpyDeviceToHost, s));
  if (measure) {
    TNO_CUDA_OK(cudaEventRecord(P->fetch_ev1, s));
    P->fetch_pending = true;
    P->fetch_pending_bytes = (double)(per - cst_h) * 8.0 * (double)N;
  }
  const size_t cst_fill = cst_h;
  if (cst_fill == 0) return TNO_OK;
  const auto fill_t0 = std::chrono::steady_clock::now();
  const double* src = P->h_jfun.data();
  auto work = [=](int64_t a, int64_t b) {
    for (int64_t c = a; c < b; ++c) copy_streaming(J_host + (size_t)c * per, src, cst_fill);
#if defined(__SSE2__)
    _mm_sfence();
#endif
  };
  if (nthreads == 1) {
    work(0, N);
  } else {
    std::vector<std::thread> pool;
    const int64_t step = (N + nthreads - 1) / nthreads;
    for (int64_t a = step; a < N; a += step) pool.emplace_back(work, a, std::min(N, a + step));
    work(0, std::min(N, step));
    for (auto& t : pool) t.join();
  }
  if (N >= 256) {
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - fill_t0).count();
    if (dt > 0.0) {
      const double rate = (double)cst_fill * 8.0 * (double)N / dt;
      P->fetch_rate_h = P->fetch_rate_h > 0.0 ? 0.5 * (P->fetch_rate_h + rate) : rate;
    }
  }
  return TNO_OK;
}

int tno_plan_fetch_stats(const tno_plan* plan, double* stats) {
  const Plan* P = reinterpret_cast<const Plan*>(plan);
  if (!P || !stats) return fail(TNO_EINVAL, "null argument");
  stats[0] = P->fetch_phi;
  stats[1] = P->fetch_rate_h * 1e-9;
  stats[2] = P->fetch_rate_d * 1e-9;
  stats[3] = (double)P->fetch_threads;
  return TNO_OK;
}

}  // extern "C"
