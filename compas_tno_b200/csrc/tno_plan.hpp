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
//   phase 2 (needs W = diag(C z)): [ Dq_0..Dq_{k-1} | z_lambdh (lambdh variable) | Dx_0..Dx_{k-1} | Dy_0..Dy_{k-1} ]
//   (Dx = dx/dq_ind, Dy = dy/dq_ind: envelopexy constraint or xyb variable with a thrust objective)
struct RhsLayout {
  int cx = 0, cy = 1, cz = 2;
  int c_zb = -1;      // first Az column
  int c_lambdv = -1;
  int c_q = -1;       // first Dq column
  int c_lambdh = -1;
  int c_qx = -1, c_qy = -1;
  int nrhs1 = 0, nrhs2 = 0, nrhs = 0;
};

// Offsets of x = [q_ind | zb | t | lambdh | lambdv]  (-1 = absent)
struct VarLayout {
  int o_q = 0, o_xyb = -1, o_zb = -1, o_t = -1, o_lambdh = -1, o_lambdv = -1;   // o_xyb: xb block, yb = o_xyb + nb
  int nvar = 0;
};

// Offsets of g / rows of J
struct ConLayout {
  int r_fun = -1, r_envxy = -1, r_env = -1, r_reac = -1;
  int ng = 0;
};

// Per-candidate scratch "slots" (one double each) of the interleaved kernel family.
struct SlotLayout {
  int s_x = 0;      // nvar
  int s_pzs = 0;    // 1
  int s_dir = 0;    // 2 : (c, s) of the candidate's load direction, (1, 0) without one
  int s_q = 0;      // m
  int s_l = 0;      // nnz_l
  int s_r = 0;      // ni * nrhs, row-major by permuted row
  int s_env = 0;    // 4 * n : ub, lb, dub/dt, dlb/dt  (thickness variable only)
  int s_grad = 0;   // nvar
  int s_rg = 0;     // 2 nb   reac_bounds constraint values
  int s_rj = 0;     // 2 nb * nvar   reac_bounds Jacobian rows (reac_bounds only)
  int s_cxy = 0;    // 2 nb : +-Rx / |R_h|, +-Ry / |R_h| of the supports (thrust objectives), read by k_thrust_grad
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
  const double* dXb;        // nb x 3 (complementary-energy objectives)
  const double* stiff;      // m (Ecomp-nonlinear)
  const double* d_b;        // second horizontal load basis (per-candidate load directions), null = absent
  const double* px0_b;
  const double* py0_b;
  const double* xlimits;    // n x 2 (envelopexy)
  const double* ylimits;
};

// ---- on-chip kernel family (tno_onchip.cu) -----------------------------------------------------------
// Row ranges of one solve step: a wide level (rows are independent) or a chain level (the rows of all chains
// of one depth of the chain tree).  Sub-slot counts are log2, [phase 0/1].
struct OcRange {
  unsigned short row0, row1;
  unsigned char pull_s[2];   // forward (pull) sub-slots per row: S * G <= 32 (shuffle reduction)
  unsigned char push_s[2];   // backward (push) sub-slots per row: any power of two
  unsigned short chain0, chain1;   // chain levels only: chains [chain0, chain1)
  unsigned short tgt0, tgt1;       // chain levels only: outside columns reached by those chains, [tgt0, tgt1) in ct_*
  unsigned short zt0, zt1;         // chain levels only: 8-row tiles of the chains' dense blocks, [zt0, zt1) in the zt table
  unsigned short rl0, rl_on;       // chain levels only: the rows in processing order (longest first) start at rowlist[rl0]
};
struct OcChain {
  unsigned short row0, T;    // rows [row0, row0 + T)
  uint32_t zoff;             // position (in the value array) of its dense strictly-lower block, (i, j) at i(i-1)/2 + j
};
// One staged piece of the factorisation stream.
struct OcChunk {
  uint32_t seg_off, seg_units;   // 16-byte units inside f_stream
  unsigned short nitems;
  unsigned char lpi_shift;       // log2 lanes sharing one item's dot product
  unsigned char kind;            // 0: result written in place; 1: via the temp buffer (rectangular chain rows)
};
// Dense work of one chain level (all its chains at once, whole CTA).
struct OcDense {
  uint32_t de0, de1;     // triangle elements (diagonal included) in de[]
  uint32_t dc0, dc1;     // chain columns in dc[]
  uint32_t zlo, zhi;     // value positions of the dense blocks of this chain level
  uint32_t tmax;         // longest chain
  uint32_t db0;          // first block descriptor of the blocked inversion in db[]
  unsigned short sb[9];  // blocks of stage s (block sub-diagonal s = 1..7) are db[db0 + sb[s]] .. db[db0 + sb[s + 1])
  unsigned short nbmax;  // 8 x 8 block rows of the longest chain
  // register-tile variant (oc_chain_tiles): one 4 x 4 tile of a chain triangle per thread
  uint32_t dt0, dt1;     // tile descriptors in dt[]
  uint32_t cbtot;        // sum of the chains' padded sizes: length of one column / row exchange buffer
  uint32_t tile_threads; // threads taking part (multiple of 32); 0 = more tiles than threads, element-wise variant
  uint32_t tile_size;    // 2 or 4
};
// Factorisation program: a short list of ops executed in order by the whole CTA.
enum { OC_OP_CHUNK = 0, OC_OP_DENSE = 1 };
struct OcOp {
  unsigned short type;
  unsigned short a, b;   // CHUNK: a = chunk id.  DENSE: a = chain level
  unsigned short pad;
};

// Device data of the on-chip kernel family.
// Value array V (shared memory, 16-bit positions):
//   [0, nlo)            strictly-lower entries of L OUTSIDE the chain blocks, CSR (row-major) order
//   [nlo, nlo + ni)     diagonal slots; hold 1 / d_j after the factorisation
//   [nlo + ni, .. + nz) dense strictly-lower chain blocks: S, then L_C, finally Z_C = L_C^-1
//   [pos_m1]            the constant -1
struct DevOnchip {
  int W;                 // panel columns (even).  phase 1: [z | z_lambdv | Az (nb) | x | y], phase 2: [Dq (k) | z_lambdh]
  int c_z, c_lambdv, c_q, c_lambdh;
  int c_az, c_x, c_y;
  int npers;             // persistent phase-1 results kept outside the panel: z (, z_lambdv)
  int ncols1, ncols2;
  int npass2;            // passes over the phase-2 columns (1 unless the panel is narrower than ncols2)
  int pw2;               // phase-2 columns per pass
  int emit_shift;        // log2 lanes walking the panel columns of one vertex when the envelope rows of J are written
  int rhs2_shift;        // log2 lanes walking the phase-2 right-hand-side columns of one row
  int gshift[2];         // log2 lanes covering the columns of one row (two columns per lane), per phase
  int nlo, nz, nv, pos_m1;
  int wide_rows, nwide, nclev, nchains, nchunks, nops, nsup;
  int dbg;               // developer experiments (TNO_DBG)
  int adj_deg;           // padded adjacency width (0: use the pointer lists)
  uint32_t nvar_magic;   // ceil(2^32 / nvar)
  // shared memory layout: offsets in doubles, index region in bytes
  int o_x, o_v, o_r, o_pers, o_qsup, o_react, o_gxy, o_azsup, o_red, o_bf, o_lpe, o_lpa, o_lpw, o_idx_bytes;
  int idx_rowptr, idx_rowcol, idx_wide, idx_clev, idx_chains, idx_rowchain, idx_chunks, idx_ops;
  int idx_bytes, smem_bytes;
  int fseg_units;        // staging buffer size, 16-byte units (3 buffers)
  int o_tmp_doubles;     // offset (doubles, inside the panel region) of the temp buffer of kind-1 chunks
  int o_tmpz_doubles;    // offset (doubles, inside the panel region) of the nz-double buffer used by the block inversion
  int idx_dense;         // byte offset of the OcDense table inside the index region
  const uint4* de;       // element descriptors: x = rowbase_i | rowbase_k << 16, y = selfpos | dgbase << 16, z = k | i << 8
  const uint32_t* dc;    // column descriptors: zoff | T << 16 | c << 24
  const uint32_t* db;    // block descriptors of the blocked inversion: zoff | T << 16 | block row << 24 | stage << 28
  const uint4* dt;       // tile descriptors: x = zoff | first diagonal slot << 16, y = T | tile row << 8 | tile column << 16, z = offset in the exchange buffers
  int o_zs;              // smem offset (doubles) of zs[ni + nb]: z of the free rows, then zb of the supports
  int simple_asm;
  int lev0_rows;         // rows of level 0 (leaves): pivots inverted by the assembly step
  const double* Bt;      // k x m (transposed B)
  const int* asm_ptr;    // generic assembly gather lists in V position order (nlo + ni entries)
  const int* asm_edge;
  const float* asm_coef;
  const unsigned short* asm_e16;  // nlo: the edge feeding each CSR position, 0xFFFF for fill / chain-side entries
  const unsigned short* asm_z16;  // nz : the edge feeding each dense chain position, 0xFFFF for none
  const uint4* f_stream;
  const unsigned char* idx_init;  // image of the shared-memory index region
  const int* edge_sup;   // m
  const int* sup_ptr;    // nb + 1
  const int* sup_edge;   // nsup
  const int* sup_other;  // nsup
  const float* sup_sign; // nsup
  const int* colsrc;     // nvar
  const int* radj_ptr;   // ni + 1
  const uint32_t* radj;  // other end (row, or 0x8000 | support index) | edge << 16
  const uint32_t* radj_pad;  // ni x adj_deg, padded with self references (zero contribution)
  // phase-2 right-hand sides, row-wise: coefficient rows [B | d0 | d0_b] in row-adjacency order (one record per radj entry)
  const double* Br;
  int br_stride;             // doubles per record (even, >= W)
  // column-wise view of the chain rows' outside entries (backward solve): for every outside column ("target")
  // reached by the chains of one chain level, the entries (value position | row << 16) of the chain rows in it
  const unsigned short* ct_ptr;
  const unsigned short* ct_col;
  const unsigned short* ct_ent;   // value position of each entry
  const unsigned char* ct_row;    // its row, relative to the first row of the chain level
  int idx_ct_row;
  // small lookup tables copied into the shared-memory index region (byte offsets): no L2 round trip in front of
  // the accesses that depend on them
  int idx_vrow, idx_rowv, idx_colsrc, idx_supptr, idx_supother, idx_supedge, idx_supsign, idx_fixed;
  int idx_supxyz;        // 3 nb doubles: coordinates of the supports
  // warp-local wide part of the solves: the wide rows above level 0 form a forest of subtrees below the chains; whole
  // subtrees are dealt to the warps (balanced), and a warp runs its rows level by level with __syncwarp() only
  int wl_on;             // 1: in use (the partition is balanced enough)
  // warp-local factorisation of the wide levels (same partition): per-warp record streams staged into the ring buffers
  int wf_on, wf_buf_bytes;
  int idx_wf_off, idx_wf_desc;
  const uint4* wf_stream;
  int idx_wl_ptr;        // u16 [warp][level 0 .. nwide]: start of the warp's rows of each level in wl_rows
  int idx_wl_rows;       // u16 rows, by warp, then level
  int idx_wl_s;          // u8 [warp][level][phase][pull, push]: log2 sub-slots
  int idx_rowlist;       // u16: rows of every chain level sorted by their number of outside entries, descending
  int idx_zt;            // uint2 per 8-row tile of a chain block: x = zoff | first row << 16, y = T | tile row << 8 (sorted by tile row, descending, per chain level)
  int zmma;              // 1: the dense chain products Z R, Z^T R run on the FP64 tensor pipe (mma.sync m8n8k4)
  int idx_radjp, radjp_in_smem;   // 16-bit copy of radj_ptr in the index region (when the adjacency has < 65535 records)
  int ct_in_smem;        // 1: the three arrays live in the shared-memory index region at idx_ct_*
  int idx_ct_ptr, idx_ct_col, idx_ct_ent;
  const double* Jfun;    // 2m x nvar
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

// Level-scheduled variant of the interleaved family (tno_interleaved.cu): one launch per elimination-tree level,
// a thread per (candidate, item of the level).  Item = entry of L for the factorisation, row / column for the solves.
struct DevLevels {
  const int* items;      // entry positions in level order (bit 31 set: diagonal entry)
  const int* dotptr;     // nitems + 1 : pairs of every item
  const int* dot;        // 3 per pair: position of U(i, c), position of U(j, c), position of the diagonal of column c
  const int* cols;       // columns (= rows) in level order
  const int* rowptr;     // CSR view of the strictly lower part of L
  const int* rowcol;
  const int* rowpos;
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
  int onchip_threads = 256;     // threads per candidate group
  int onchip_groups = 1;        // candidate groups per CTA
  int onchip_ctas = 1;          // resident CTAs per SM
  int onchip_ctas_per_sm = 1;   // candidates in flight per SM = onchip_ctas * onchip_groups
  Symbolic sym;
  std::vector<int> h_vrow, h_vadj_ptr, h_vadj_edge, h_vadj_other, h_fixed;
  std::vector<float> h_vadj_sign;
  int device = 0;
  int kernel = TNO_KERNEL_INTERLEAVED;
  bool auto_family = false;            // created with TNO_KERNEL_AUTO: large batches may go to the other family
  bool thread_per_candidate = false;   // TNO_KERNEL_INTERLEAVED_TPC: the unscheduled factor / solve kernels
  std::vector<void*> dev_allocs;       // everything to cudaFree at destroy
  // emit tables
  EmitTable t_g, t_grad, t_z, t_q, t_J;   // t_J covers the rows below the funicular block when jfun is set
  const double* jfun = nullptr;           // 2m x nvar candidate-independent rows [B 0; -B 0] (+ d0 column) of J
  int64_t jfun_count = 0;
  std::vector<double> h_jfun;             // host copy of the same block (tno_fetch_jacobian_host)
  // adaptive split of the constant rows between host threads and the DMA engine (tno_fetch_jacobian_host)
  double fetch_rate_h = 0.0, fetch_rate_d = 0.0;   // measured bytes / s (smoothed), 0 = not measured yet
  double fetch_phi = 0.0;                          // fraction used by the last fetch
  int fetch_threads = 0;
  cudaEvent_t fetch_ev0 = nullptr, fetch_ev1 = nullptr;
  bool fetch_pending = false;
  double fetch_pending_bytes = 0.0;
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
  // level-scheduled interleaved kernels
  bool lv_ready = false;
  DevLevels lv{};
  std::vector<int> lv_item_ptr;   // nlevels + 1 : items of each level
  std::vector<char> lv_seg;       // per level: use the segmented factor kernel
};

// kernel family entry points (defined in the .cu files); all asynchronous on `stream`
int run_interleaved(Plan* P, const double* x_dev, int64_t N, int64_t ldx, const tno_batch_inputs* opt,
                    const tno_outputs* out, cudaStream_t stream);

int build_onchip(Plan* P, const tno_problem_desc* D);   // host schedule + upload; sets onchip_ok
int run_onchip(Plan* P, const double* x_dev, int64_t N, int64_t ldx, const tno_batch_inputs* opt,
               const tno_outputs* out, cudaStream_t stream);

// replicate the candidate-independent funicular rows of J for N candidates (streaming kernel, tno_onchip.cu)
void launch_fill_funicular(Plan* P, const double* block, int64_t count, double* J, int64_t N, cudaStream_t stream);

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
