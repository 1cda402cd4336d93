// internal.cuh -- shared declarations of the cora_b200 CUDA library (not installed).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/cora_b200.h"
#include "layout.hpp"

namespace cora_b200 {

// ------------------------------------------------------------------ errors ----
struct Error : public std::runtime_error {
  int code;
  Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
void set_last_error(const std::string &m);

#define CUDA_CHECK(expr)                                                                 \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      throw ::cora_b200::Error(CORA_B200_ECUDA, std::string("CUDA error: ") +            \
                                                    cudaGetErrorString(_e) + " at " +    \
                                                    __FILE__ + ":" + std::to_string(__LINE__)); \
  } while (0)

// ------------------------------------------------------- device-side layout ----
struct DevLayout {
  int d, D1, n, m, l;
  int N;          // rows
  int TR, TP, numTiles;
  int nPoseRows;  // D1*n
  int G;
  int maxSlots;
  int numLong;
  const int *tile_slots;
  const long long *tile_boff;
  const long long *tile_coff;
  const double *bval;  // block-ELL values (swappable: Q or S = Q - Lambda + eta I)
  const int *bcol;
  const double *sdiag;  // diagonal of scalar rows (swappable)
  const int *grp_ptr;
  const unsigned *rem_pk;
  const double *rem_val;
  const int *tile_long_ptr;
  const int *long_grp;
  const int *long_ptr;
  const unsigned *long_pk;
  const double *long_val;
  int numChunks;              // hub rows split in chunks of kHubChunk entries (persistent kernel)
  const int *chunk_beg;
  const int *chunk_end;
  const int *long_chunk_ptr;  // per hub group: first chunk
  int TRP, maxTileSpill;      // per-tile spill slices (persistent kernel)
  const long long *tile_sp_off;
  const int *tile_sp_cnt;
  const int *sp_gptr;
  const unsigned *sp_pk;
  const double *sp_val;
  const double *dinv;  // 1/diag(Q), internal order
  const int *int2ref;
};

// ---------------------------------------------------------- STPCG control ------
// Device-resident state of one Steihaug-Toint solve
// (libs/Optimization/.../LinearAlgebra/IterativeSolvers.h:207-426).
enum CgMode { CG_MODE_STEP = 0, CG_MODE_BOUNDARY = 1 };
enum CgExit { CG_EXIT_NONE = 0, CG_EXIT_TARGET = 1, CG_EXIT_MAXIT = 2, CG_EXIT_BOUNDARY = 3, CG_EXIT_KERNEL = 4 };
struct CgCtrl {
  int state;  // 0 running, 1 done
  int mode;
  int it;
  int max_it;
  int exit_reason;
  int pad;
  double rv;  // <r, v>
  double target;
  double Delta, Delta2;
  double sMp, sM2, pM2, sM2_next;
  double alpha, beta, kappa, sigma;
  double hM;  // ||s||_M on exit
  double eps;
  double kappa_fgr, theta;
  double pp, HpHp;
};

// scalar slots in the device `scal` array
enum ScalSlot {
  SC_XG = 0,    // <x, Qx>, <grad, grad>            (k_qprod<GRAD> writes 3 slots)
  SC_GG = 1,
  SC_RV = 4,    // <g, P g>, <P g, P g>
  SC_HH = 8,    // <h, h>, <g, h>
  SC_GH = 9,
  SC_XG2 = 12,  // proposed point: <x+, Q x+>, <grad+, grad+>
  SC_GG2 = 13,
  SC_HHH = 16,  // <h, Hess h>, <Hh, Hh>, <h, h>
  SC_RV2 = 20,  // second <g, P g> slot (current / proposed alternate)
  SC_TMP = 24,  // scratch (8)
  SC_COUNT = 32
};

enum PostOp { POST_STORE = 0, POST_CG_HESS = 1, POST_CG_UPDATE = 2 };
enum QMode { QM_SPMM = 0, QM_GRAD = 1, QM_HESS = 2 };

struct QArgs {
  const double *X;  // multiplied vector (N x r internal)
  const double *Y;  // base point (GRAD: == X; HESS)
  const double *G;  // Euclidean gradient at Y (HESS)
  double *out;      // SPMM: Q X ; GRAD: Riemannian gradient ; HESS: Hess[X]
  double *out2;     // GRAD: Q X (Euclidean gradient)
  const double *longbuf;
  double *partials;   // numTiles x 4
  unsigned *counter;
  double *scal;
  CgCtrl *ctrl;  // nullptr: not gated
  int r;
  int mode;
  int post;
  int slot;
};

struct UArgs {  // k_cg_update / precondition+projection
  const double *Y;
  const double *P;   // search direction p
  const double *HP;  // Hess p
  double *S;         // step s
  double *R;         // residual r (axpy: updated in place; else read-only)
  const double *Z;   // externally preconditioned residual (zsrc == 2)
  double *V;         // output v = proj_Y(M^-1 r)
  double *partials;
  unsigned *counter;
  double *scal;
  CgCtrl *ctrl;
  int r;
  int do_axpy;  // 1: s += alpha p, r += alpha Hp (or boundary step)
  int zsrc;     // 0: r * dinv (Jacobi), 1: r (identity), 2: Z
  int do_proj;  // 0: stop after the axpy (external preconditioner follows)
  int post;
  int slot;
  int gated;
};

}  // namespace cora_b200
