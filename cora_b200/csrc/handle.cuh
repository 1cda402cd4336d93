// handle.cuh -- the device-side problem handle and its launch helpers.
#pragma once
#include <chrono>
#include <cstring>

#include "internal.cuh"
#include "stream_layout.hpp"

namespace cora_b200 {

template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  void alloc(size_t count) {
    free();
    n = count;
    if (count) CUDA_CHECK(cudaMalloc((void **)&p, count * sizeof(T)));
  }
  void reserve(size_t count) {  // grow only: keeps the allocation when it is already large enough
    if (n < count || p == nullptr) alloc(count);
  }
  void upload(const std::vector<T> &v, cudaStream_t s) {
    alloc(v.size() ? v.size() : 1);
    if (!v.empty()) CUDA_CHECK(cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s));
  }
  void free() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
  ~DevBuf() { free(); }
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
};

// names of the N x r work vectors
enum Vec {
  V_X = 0,   // current iterate
  V_G,       // Euclidean gradient Q X
  V_GRAD,    // Riemannian gradient
  V_PG,      // preconditioned gradient / STPCG v
  V_S,       // STPCG step
  V_R,       // STPCG residual
  V_P,       // STPCG direction
  V_HP,      // Hess p
  V_XP,      // proposed iterate
  V_GP,      // Q X+
  V_GRADP,   // grad at X+
  V_Z,       // external preconditioner output / scratch
  V_V,       // STPCG preconditioned residual v (V_PG must survive a rejected step)
  V_T0,      // scratch (tier-1 staging)
  V_T1,
  V_COUNT
};

struct ChainChol;  // chain_chol.cuh
struct ChainSym;   // chain_factor_dev.cuh

}  // namespace cora_b200

struct cora_b200_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int sm_count = 148;
  cora_b200::HostLayout HL;  // structure (values freed after upload)
  cora_b200::DevLayout DL{};
  // device copies of the layout
  cora_b200::DevBuf<int> d_tile_slots, d_bcol, d_grp_ptr, d_tile_long_ptr, d_long_grp, d_long_ptr, d_int2ref;
  cora_b200::DevBuf<int> d_chunk_beg, d_chunk_end, d_long_chunk_ptr, d_tile_sp_cnt, d_sp_gptr;
  cora_b200::DevBuf<long long> d_tile_sp_off;
  cora_b200::DevBuf<unsigned> d_sp_pk;
  cora_b200::DevBuf<double> d_sp_val;
  cora_b200::DevBuf<long long> d_tile_boff, d_tile_coff;
  cora_b200::DevBuf<unsigned> d_rem_pk, d_long_pk;
  cora_b200::DevBuf<double> d_bval, d_sdiag, d_rem_val, d_long_val, d_dinv, d_diag;
  // certificate values S + eta I on the same structure
  cora_b200::DevBuf<double> d_bvalS, d_sdiagS, d_lam_st, d_lam_ob;
  // workspace
  int ws_r = 0;
  cora_b200::DevBuf<double> ws[cora_b200::V_COUNT];
  cora_b200::DevBuf<double> d_stage;    // reference-layout staging, N x ws_r
  cora_b200::DevBuf<double> d_longbuf;  // numLong x D1 x ws_r
  cora_b200::DevBuf<double> d_partials;
  cora_b200::DevBuf<double> d_scal;
  cora_b200::DevBuf<unsigned> d_counter;
  cora_b200::DevBuf<cora_b200::CgCtrl> d_ctrl;
  double *h_scal = nullptr;             // pinned
  cora_b200::CgCtrl *h_ctrl = nullptr;  // pinned
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_chunk[2] = {nullptr, nullptr};
  // preconditioner
  int precond = CORA_B200_PRECON_JACOBI;
  double reg_max_cond = 1e6;
  double lambda_reg = -1.0;
  bool lambda_user = false;
  cora_b200::ChainChol *chol = nullptr;  // RegularizedCholesky factor of (Q + lambda I)[:-1,:-1]
  cora_b200::ChainSym *chain_sym = nullptr;  // structure of the chain factorisation (built on first use)
  cora_b200::ChainChol *chol_spare = nullptr;  // buffers of the last released factor, recycled by the next one
  int chain_sym_state = 0;                   // 0: not built, 1: built, 2: not a chain graph
  std::string chain_sym_error;
  int precond_requested = CORA_B200_PRECON_JACOBI;  // what the caller asked for (precond: what is applied)
  // Formulation::Implicit (translations marginalised, src/CORA_problem.cpp:714-757): iterates carry zero
  // translation rows; every data-matrix product runs on the translation-completed copy (implicit.cuh)
  int formulation = CORA_B200_FORMULATION_EXPLICIT;
  cora_b200::ChainChol *ltrans = nullptr;    // factor of Q33 with the last translation pinned (LtransCholRed_)
  cora_b200::DevBuf<double> d_imp[3];        // completed copy, T^T Y, L^-1 T^T Y
  int io_rows() const { return formulation == CORA_B200_FORMULATION_IMPLICIT ? HL.d * HL.n + HL.m : HL.N; }
  int last_cert_branch = CORA_B200_CERT_NONE;
  // resident iterate rank
  int resident_r = 0;
  int64_t launches = 0;
  int cg_chunk = 8;
  // optional per-launch timing of the dominant kernel (k_qprod<HESS> inside STPCG)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_ev;
  size_t prof_n = 0;
  cora_b200::DevBuf<double> d_snap;  // snapshot of the resident iterate
  // persistent TNT kernel (persistent.cuh)
  cora_b200::DevBuf<double> d_longpart, d_ppartials, d_trace, d_lamT, d_lamS, d_lmpart;
  cora_b200::DevBuf<unsigned long long> d_bar;
  cora_b200::DevBuf<int> d_cta_t0;
  cora_b200::DevBuf<unsigned char> d_tntdev;
  cora_b200::DevBuf<unsigned long long> d_prof_all;   // [grid][PH_COUNT] per-CTA phase times of the last launch
  std::vector<unsigned long long> h_prof_all;
  int prof_all_grid = 0;
  void *h_tntdev = nullptr;  // pinned TntDev
  std::vector<double> h_trace;
  int trace_cap = 0;
  int persistent_grid = 0, persistent_grid_r = -1, persistent_nbuf = 0, persistent_threads = 256;
  size_t persistent_smem = 0;
  // strip layout of the data matrix (stream_layout.hpp) for the rank-specialised streaming kernels
  cora_b200::StreamHost SH;
  cora_b200::DevBuf<unsigned> d_rec_off;
  cora_b200::DevBuf<unsigned char> d_rec;
  cora_b200::DevBuf<double> d_diagQ, d_sdiagP, d_diagL, d_sdiagL;
  cora_b200::DevBuf<int> d_warp_strip;
  bool persistent_stream = false;   // the configured rank runs on a streaming kernel
  int stream_stages = 2;            // ring depth per warp
  int stream_stage_doubles = 0, stream_xw = 0, stream_yw = 0;
  double stream_scalar_weight = 1.0;
  bool allow_stream = true;
  bool stream_interleave = false;   // warps of a CTA take the CTA's strips round-robin
  void *persistent_kfn = nullptr, *persistent_spmm_kfn = nullptr;
  bool use_persistent = true;
  int snap_r = 0;
};
