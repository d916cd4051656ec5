// Internal plan layout shared by the C ABI and the kernel families.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/tno_b200.h"
#include "tno_symbolic.hpp"

namespace tno {

// Column layout of the right-hand-side panel R (ni rows x nrhs columns, permuted row order).
//   phase 1 (independent of z):  [ x | y | z | Az_0..Az_{nb-1} (zb variable) | z_lambdv (lambdv variable) ]
//   phase 2 (needs W = diag(C z)): [ Dq_0..Dq_{k-1} | z_lambdh (lambdh variable) ]
struct RhsLayout {
  int cx = 0, cy = 1, cz = 2;
  int c_zb = -1;      // first Az column
  int c_lambdv = -1;
  int c_q = -1;       // first Dq column
  int c_lambdh = -1;
  int nrhs1 = 0, nrhs2 = 0, nrhs = 0;
};

// Offsets of x = [q_ind | zb | t | lambdh | lambdv]  (-1 = absent)
struct VarLayout {
  int o_q = 0, o_zb = -1, o_t = -1, o_lambdh = -1, o_lambdv = -1;
  int nvar = 0;
};

// Offsets of g / rows of J
struct ConLayout {
  int r_fun = -1, r_env = -1, r_reac = -1;
  int ng = 0;
};

// Per-candidate scratch "slots" (one double each) of the interleaved kernel family.
struct SlotLayout {
  int s_x = 0;      // nvar
  int s_pzs = 0;    // 1
  int s_q = 0;      // m
  int s_l = 0;      // nnz_l
  int s_r = 0;      // ni * nrhs, row-major by permuted row
  int s_env = 0;    // 4 * n : ub, lb, dub/dt, dlb/dt  (thickness variable only)
  int s_grad = 0;   // nvar
  int s_rg = 0;     // 2 nb   reac_bounds constraint values
  int s_rj = 0;     // 2 nb * nvar   reac_bounds Jacobian rows
  int nslots = 0;
};

// Element descriptor of the table-driven emitter:
//   out[e] = sign_a * scratch[slot_a] + sign_b * scratch[slot_b] + c      (slot < 0: term absent)
struct EmitDesc2 {
  int32_t slot_a;
  int32_t slot_b;
  float sign_a;
  float sign_b;
  double c;
};

// Device-resident constant data of a plan (all pointers are device pointers).
struct DevPlan {
  int n, m, ni, nb, k;
  uint32_t variables, constraints;
  int objective, envelope_kind;
  double env[8];
  VarLayout var;
  ConLayout con;
  RhsLayout rhs;
  SlotLayout slot;
  int nnz_l;
  // symbolic
  const int* colptr;
  const int* rowidx;
  const int64_t* upd_ptr;
  const int* upd_pos;
  const int* asm_ptr;       // nnz_l + 1
  const int* asm_edge;
  const float* asm_coef;
  // topology
  const int* edge_u;
  const int* edge_v;
  const int* vrow;          // n: permuted row of a free vertex, or -1 - b for support b
  const int* row_vertex;    // ni
  const int* fixed;         // nb vertex ids
  const int* vadj_ptr;      // n + 1
  const int* vadj_edge;
  const int* vadj_other;
  const float* vadj_sign;   // C[e, v]
  // values
  const double* B;          // m x k
  const double* d;          // m
  const double* P;          // n x 3
  const double* X;          // n x 3
  const double* lb;
  const double* ub;
  const double* s;
  const double* qmin;
  const double* qmax;
  const double* b;          // nb x 2
  const double* px0;
  const double* py0;
  const double* pzv;
  const double* pz0;
};

// Device data of the on-chip kernel family (tno_onchip.cu): level-ordered index streams, all 16-bit.
// One level set of the elimination tree (rows == columns are contiguous per level after the level ordering).
struct OcLevel {
  unsigned short row0, row1;   // rows / columns [row0, row1)
  unsigned short nitems;       // entries of L in those columns (factorisation work items)
  unsigned char lpi_shift;     // log2 of the lanes sharing one item's dot product in the factorisation
  unsigned char pad0;
  uint32_t seg_off;            // factor stream segment: offset and length in 16-byte units
  uint32_t seg_units;
  // log2 of the sub-slots splitting one row, [phase 0/1][team: 0 = whole CTA, 1 = one warp]
  unsigned char pull_s[2][2];  // forward solve (pull): S * G <= 32
  unsigned char push_s[2][2];  // backward solve (push)
};

// Device data of the on-chip kernel family (tno_onchip.cu).
// Value array Lv (shared memory): strictly-lower entries of L in CSR (row-major) order [0, nlow), then the ni
// diagonal slots [nlow, nlow + ni) which hold 1 / d_j after the factorisation.
struct DevOnchip {
  int W;                 // panel columns: [z | z_lambdv | Dq (k) | z_lambdh]; Az, x, y transient in the Dq columns
  int c_z, c_lambdv, c_q, c_lambdh;   // persistent panel columns (-1 = absent)
  int c_az, c_x, c_y;    // transient phase-1 columns (overwritten by phase 2)
  int ncols1, ncols2;    // columns solved in phase 1: [0, ncols1); in phase 2: [c_q, c_q + ncols2)
  int rhs2_shift;        // log2 of the lanes walking the phase-2 right-hand-side columns of one row
  int gshift[2];         // log2 of the lanes covering the columns of one row (two columns per lane), per phase
  int lsplit[2];         // first level handled by a single warp (rows * G <= 32 from there to the root), per phase
  int nlevels, nlow, nsup;
  uint32_t nvar_magic;   // ceil(2^32 / nvar): flat index -> (row, col)
  // shared memory layout: offsets in doubles, o_idx in bytes
  int o_x, o_l, o_r, o_env, o_qsup, o_react, o_gxy, o_azsup, o_red, o_idx_bytes;
  int idx_rowptr, idx_rowcol, idx_levels;   // byte offsets inside the index region
  int idx_bytes;
  int smem_bytes;
  int stage;             // 1: factor stream staged through the panel with cp.async (3 buffers of fseg_units)
  int fseg_units;        // largest level segment, 16-byte units
  const double* Bt;      // k x m (transposed B)
  const int* asm_ptr;    // nnz_l + 1, assembly gather lists in Lv position order
  const int* asm_edge;
  const float* asm_coef;
  const uint4* f_stream; // per level: (nitems + 1) records {pos | diag << 31, first pair} then pairs {a | b << 16 | dp << 32}
  const unsigned char* idx_init;   // image of the shared-memory index region (rowptr16, rowcol16, levels)
  const int* edge_sup;   // m: index of an edge in the support-incidence list or -1
  const int* sup_ptr;    // nb + 1
  const int* sup_edge;   // nsup
  const int* sup_other;  // nsup: vertex at the other end
  const float* sup_sign; // nsup: C[e, support]
  const int* colsrc;     // nvar: panel column feeding Jacobian column c (-1: written in the phase-1 epilogue)
};

struct EmitTable {
  EmitDesc2* dev = nullptr;
  int64_t count = 0;
};

// optional per-kernel timing (CUDA events on the launching stream)
enum KernelKind { KK_PACK = 0, KK_Q, KK_ASSEMBLE, KK_FACTOR, KK_RHS1, KK_SOLVE1, KK_RHS2, KK_SOLVE2, KK_EPILOGUE,
                  KK_EMIT_SMALL, KK_EMIT_J, KK_STATUS, KK_ONCHIP, KK_COUNT };
extern const char* const kKernelNames[KK_COUNT];
struct ProfRec {
  int kind;
  cudaEvent_t a, b;
};

struct Plan {
  tno_plan_info info{};
  bool profiling = false;
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
  double prof_ms[KK_COUNT] = {0};
  int64_t prof_launches[KK_COUNT] = {0};
  void prof_begin(int kind, cudaStream_t s) {
    if (!profiling) return;
    ProfRec r;
    r.kind = kind;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
      if (prof_pool.empty()) {
        cudaEventCreate(e);
      } else {
        *e = prof_pool.back();
        prof_pool.pop_back();
      }
    }
    cudaEventRecord(r.a, s);
    prof_recs.push_back(r);
  }
  void prof_end(cudaStream_t s) {
    if (!profiling) return;
    cudaEventRecord(prof_recs.back().b, s);
  }
  DevPlan dp{};
  DevOnchip oc{};
  bool onchip_ok = false;
  int onchip_threads = 256;
  int onchip_ctas_per_sm = 1;
  Symbolic sym;
  std::vector<int> h_vrow, h_vadj_ptr, h_vadj_edge, h_vadj_other, h_fixed;
  std::vector<float> h_vadj_sign;
  int device = 0;
  int kernel = TNO_KERNEL_INTERLEAVED;
  std::vector<void*> dev_allocs;       // everything to cudaFree at destroy
  // emit tables
  EmitTable t_g, t_grad, t_z, t_q, t_J;
  // scratch (interleaved family): nslots x chunk doubles
  double* scratch = nullptr;
  int32_t* status_scratch = nullptr;
  int64_t chunk = 0;
  int64_t scratch_bytes = 0;
  // host-call staging
  cudaStream_t stream = nullptr;
  double* h_pinned = nullptr;
  size_t h_pinned_bytes = 0;
  double* d_stage = nullptr;
  size_t d_stage_bytes = 0;
  int64_t launches = 0;
  int sm_count = 148;
};

// kernel family entry points (defined in the .cu files); all asynchronous on `stream`
int run_interleaved(Plan* P, const double* x_dev, int64_t N, int64_t ldx, const tno_batch_inputs* opt,
                    const tno_outputs* out, cudaStream_t stream);

int build_onchip(Plan* P, const tno_problem_desc* D);   // host schedule + upload; sets onchip_ok
int run_onchip(Plan* P, const double* x_dev, int64_t N, int64_t ldx, const tno_batch_inputs* opt,
               const tno_outputs* out, cudaStream_t stream);

void set_error(const std::string& msg);
int upload_bytes(Plan* P, const void* host, size_t bytes, const void** dev);
template <class T>
inline int upload_vec(Plan* P, const std::vector<T>& h, const T** dev) {
  return upload_bytes(P, h.data(), h.size() * sizeof(T), (const void**)dev);
}

#define TNO_CUDA_OK(call)                                                                     \
  do {                                                                                        \
    cudaError_t e__ = (call);                                                                 \
    if (e__ != cudaSuccess) {                                                                 \
      ::tno::set_error(std::string(#call) + ": " + cudaGetErrorString(e__));                  \
      return TNO_ECUDA;                                                                       \
    }                                                                                         \
  } while (0)

}  // namespace tno
