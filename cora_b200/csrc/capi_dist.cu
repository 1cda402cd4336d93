// capi_dist.cu -- multi-GPU helper: independent restarts / staircase ranks run one per
// GPU with NO collective on the data path; the only exchange is the final gather of the best
// certified solution: ncclAllGather of a {f, certified} record, arg-min on every rank, ncclBroadcast
// of the winner's iterate over NVLink/NVSwitch.  NCCL is resolved with dlopen at first use (the
// library has no link-time dependency on it: CPU-only hosts load libcora_b200.so without NCCL).
#include <dlfcn.h>
#include <nccl.h>

#include "ops.cuh"

using namespace cora_b200;

namespace {
struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi &nccl() {
  static NcclApi api;
  if (api.lib) return api;
  // CORA_B200_NCCL_LIB: explicit path.  Otherwise the SONAME: a copy already loaded in the process (e.g. the one
  // bundled with PyTorch, which cora_b200/capi.py loads first) is reused by the dynamic loader.
  const char *names[] = {getenv("CORA_B200_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    if (!nm || !*nm) continue;
    api.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) throw Error(CORA_B200_ERUNTIME, std::string("cannot load NCCL (libnccl.so.2): ") + dlerror());
  auto sym = [&](const char *s) {
    void *p = dlsym(api.lib, s);
    if (!p) throw Error(CORA_B200_ERUNTIME, std::string("NCCL symbol missing: ") + s);
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  return api;
}
void nccl_check(ncclResult_t r, const char *what) {
  if (r != ncclSuccess) throw Error(CORA_B200_ERUNTIME, std::string(what) + ": " + nccl().GetErrorString(r));
}
}  // namespace

// arg-min of f over the certified ranks, over all ranks when none is certified (ties: lowest rank)
extern "C" int cora_b200_select_best(int world_size, const double *f, const int *certified, int *winner) {
  API_BEGIN
  require(world_size > 0 && f && certified && winner, "bad argument");
  int best = -1;
  bool any = false;
  for (int i = 0; i < world_size; ++i) any = any || certified[i] != 0;
  for (int i = 0; i < world_size; ++i) {
    if (any && !certified[i]) continue;
    if (f[i] != f[i]) continue;  // NaN never wins
    if (best < 0 || f[i] < f[best]) best = i;
  }
  *winner = best < 0 ? 0 : best;
  API_END
}

extern "C" int cora_b200_nccl_unique_id(void *id128) {
  API_BEGIN
  require(id128 != nullptr, "NULL id buffer");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  nccl_check(nccl().GetUniqueId((ncclUniqueId *)id128), "ncclGetUniqueId");
  API_END
}

extern "C" int cora_b200_nccl_init(void **comm, int device, int world_size, int rank, const void *id128) {
  API_BEGIN
  require(comm && id128 && world_size > 0 && rank >= 0 && rank < world_size, "bad argument");
  CUDA_CHECK(cudaSetDevice(device));
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  ncclComm_t c = nullptr;
  nccl_check(nccl().CommInitRank(&c, world_size, id, rank), "ncclCommInitRank");
  {  // warm the communicator: the first collective on a cold communicator costs ~1 s (channel set-up); pay it here,
     // not inside the gather that follows a solve
    cudaStream_t s = nullptr;
    CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    DevBuf<double> w;
    w.alloc(2 * (size_t)world_size + 2);
    CUDA_CHECK(cudaMemsetAsync(w.p, 0, w.n * sizeof(double), s));
    nccl_check(nccl().AllGather(w.p + 2 * world_size, w.p, 2, ncclDouble, c, s), "ncclAllGather (warm-up)");
    nccl_check(nccl().Broadcast(w.p, w.p, 2, ncclDouble, 0, c, s), "ncclBroadcast (warm-up)");
    CUDA_CHECK(cudaStreamSynchronize(s));
    cudaStreamDestroy(s);
  }
  *comm = (void *)c;
  API_END
}

extern "C" int cora_b200_nccl_destroy(void *comm) {
  API_BEGIN
  if (comm) nccl_check(nccl().CommDestroy((ncclComm_t)comm), "ncclCommDestroy");
  API_END
}

extern "C" int cora_b200_gather_best(void *nccl_comm, cora_b200_t *h, int world_size, int my_rank, int r_max,
                                     double f, int certified, double *X_inout, int *winner_rank, double *winner_f) {
  API_BEGIN
  require(nccl_comm && h && X_inout && winner_rank, "NULL argument");
  require(world_size > 0 && my_rank >= 0 && my_rank < world_size && r_max > 0, "bad argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  ensure_workspace(h, r_max);
  ncclComm_t comm = (ncclComm_t)nccl_comm;
  const size_t nE = (size_t)h->DL.N * r_max;
  DevBuf<double> rec;
  rec.alloc(2 * (size_t)world_size + 2);
  double mine[2] = {f, certified ? 1.0 : 0.0};
  CUDA_CHECK(cudaMemcpyAsync(rec.p + 2 * world_size, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  nccl_check(nccl().AllGather(rec.p + 2 * world_size, rec.p, 2, ncclDouble, comm, h->stream), "ncclAllGather");
  std::vector<double> all(2 * (size_t)world_size);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), rec.p, all.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  std::vector<double> fs(world_size);
  std::vector<int> cs(world_size);
  for (int i = 0; i < world_size; ++i) { fs[i] = all[2 * i]; cs[i] = all[2 * i + 1] != 0.0; }
  int win = 0;
  if (cora_b200_select_best(world_size, fs.data(), cs.data(), &win) != CORA_B200_OK)
    throw Error(CORA_B200_ERUNTIME, cora_b200_last_error());
  // the winner's iterate (reference layout, N x r_max column-major, zero padded by the caller)
  if (my_rank == win)
    CUDA_CHECK(cudaMemcpyAsync(h->d_stage.p, X_inout, nE * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  nccl_check(nccl().Broadcast(h->d_stage.p, h->d_stage.p, nE, ncclDouble, win, comm, h->stream), "ncclBroadcast");
  if (my_rank != win)
    CUDA_CHECK(cudaMemcpyAsync(X_inout, h->d_stage.p, nE * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  *winner_rank = win;
  if (winner_f) *winner_f = fs[win];
  API_END
}

// Device-resident variant: every rank holds its solution as the resident iterate of its handle (rank r, as left by
// cora_b200_solve / cora_b200_tnt_resident); the winner's iterate is broadcast handle to handle over NVLink with no
// host staging and becomes the resident iterate of every rank.
extern "C" int cora_b200_gather_best_resident(void *nccl_comm, cora_b200_t *h, int world_size, int my_rank, double f,
                                              int certified, int *winner_rank, double *winner_f) {
  API_BEGIN
  require(nccl_comm && h && winner_rank, "NULL argument");
  require(world_size > 0 && my_rank >= 0 && my_rank < world_size, "bad argument");
  require(h->resident_r > 0, "no resident iterate on this handle");
  CUDA_CHECK(cudaSetDevice(h->device));
  ncclComm_t comm = (ncclComm_t)nccl_comm;
  const int r = h->resident_r;
  DevBuf<double> rec;
  rec.alloc(3 * (size_t)world_size + 3);
  double mine[3] = {f, certified ? 1.0 : 0.0, (double)r};
  CUDA_CHECK(cudaMemcpyAsync(rec.p + 3 * world_size, mine, sizeof(mine), cudaMemcpyHostToDevice, h->stream));
  nccl_check(nccl().AllGather(rec.p + 3 * world_size, rec.p, 3, ncclDouble, comm, h->stream), "ncclAllGather");
  std::vector<double> all(3 * (size_t)world_size);
  CUDA_CHECK(cudaMemcpyAsync(all.data(), rec.p, all.size() * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  std::vector<double> fs(world_size);
  std::vector<int> cs(world_size);
  for (int i = 0; i < world_size; ++i) {
    fs[i] = all[3 * i]; cs[i] = all[3 * i + 1] != 0.0;
    if ((int)all[3 * i + 2] != r) throw Error(CORA_B200_EINVAL, "gather_best_resident: the ranks hold iterates of different rank");
  }
  int win = 0;
  if (cora_b200_select_best(world_size, fs.data(), cs.data(), &win) != CORA_B200_OK)
    throw Error(CORA_B200_ERUNTIME, cora_b200_last_error());
  double *X = h->ws[V_X].p;  // same internal layout on every rank (same problem)
  nccl_check(nccl().Broadcast(X, X, (size_t)h->DL.N * r, ncclDouble, win, comm, h->stream), "ncclBroadcast");
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  *winner_rank = win;
  if (winner_f) *winner_f = fs[win];
  API_END
}

// ------------------------------------------------------------ row-partitioned product over peer memory ----
#include "peer_product.cuh"

struct cora_b200_peer {
  PeerCtx c;
};

namespace {
constexpr size_t kIpc = sizeof(cudaIpcMemHandle_t);  // 64 bytes
}

// Create the peer context of this rank and export its three allocations: [operand X | landmark staging | flags],
// 3 x 64 bytes into `handles`.  n_landmark_rows: landmark rows of the local problem (staging size).
extern "C" int cora_b200_peer_create(cora_b200_t *h, int r, int n_landmark_rows, cora_b200_peer_t **out, void *handles) {
  API_BEGIN
  require(h && out && handles && r > 0 && n_landmark_rows >= 0, "bad argument");
  CUDA_CHECK(cudaSetDevice(h->device));
  ensure_workspace(h, r);
  h->resident_r = r;
  auto *p = new cora_b200_peer();
  PeerCtx &c = p->c;
  c.h = h; c.r = r; c.n_lm = n_landmark_rows;
  c.stage.alloc((size_t)std::max(n_landmark_rows, 1) * r);
  c.flags.alloc(2 * kPeerMaxWorld);
  c.err.alloc(1);
  CUDA_CHECK(cudaMemset(c.flags.p, 0, 2 * kPeerMaxWorld * sizeof(unsigned long long)));
  CUDA_CHECK(cudaMemset(c.err.p, 0, sizeof(int)));
  cudaIpcMemHandle_t hx, hs, hf;
  CUDA_CHECK(cudaIpcGetMemHandle(&hx, h->ws[V_X].p));
  CUDA_CHECK(cudaIpcGetMemHandle(&hs, c.stage.p));
  CUDA_CHECK(cudaIpcGetMemHandle(&hf, c.flags.p));
  std::memcpy((char *)handles, &hx, kIpc);
  std::memcpy((char *)handles + kIpc, &hs, kIpc);
  std::memcpy((char *)handles + 2 * kIpc, &hf, kIpc);
  *out = p;
  API_END
}

// all_handles: world x 3 x 64 bytes (every rank's export, in rank order).  Plan: ghost row g of this rank's operand
// (internal row ghost_dst[g]) is row ghost_src[g] of rank ghost_peer[g]'s operand; lm_rows: this rank's landmark rows.
extern "C" int cora_b200_peer_connect(cora_b200_peer_t *p, int world, int rank, const void *all_handles, int n_ghost,
                                      const int32_t *ghost_peer, const int32_t *ghost_src, const int32_t *ghost_dst,
                                      const int32_t *lm_rows) {
  API_BEGIN
  require(p && all_handles && world > 0 && world <= kPeerMaxWorld && rank >= 0 && rank < world, "bad argument");
  PeerCtx &c = p->c;
  CUDA_CHECK(cudaSetDevice(c.h->device));
  c.world = world; c.rank = rank;
  for (int q = 0; q < world; ++q) {
    if (q == rank) {
      c.peer_x[q] = c.h->ws[V_X].p; c.peer_stage[q] = c.stage.p; c.peer_flags[q] = c.flags.p;
      continue;
    }
    cudaIpcMemHandle_t hx, hs, hf;
    const char *base = (const char *)all_handles + (size_t)q * 3 * kIpc;
    std::memcpy(&hx, base, kIpc); std::memcpy(&hs, base + kIpc, kIpc); std::memcpy(&hf, base + 2 * kIpc, kIpc);
    void *px = nullptr, *ps = nullptr, *pf = nullptr;
    CUDA_CHECK(cudaIpcOpenMemHandle(&px, hx, cudaIpcMemLazyEnablePeerAccess));
    CUDA_CHECK(cudaIpcOpenMemHandle(&ps, hs, cudaIpcMemLazyEnablePeerAccess));
    CUDA_CHECK(cudaIpcOpenMemHandle(&pf, hf, cudaIpcMemLazyEnablePeerAccess));
    c.opened[3 * q] = px; c.opened[3 * q + 1] = ps; c.opened[3 * q + 2] = pf;
    c.peer_x[q] = (const double *)px; c.peer_stage[q] = (const double *)ps; c.peer_flags[q] = (unsigned long long *)pf;
  }
  c.n_ghost = n_ghost;
  auto up = [&](DevBuf<int> &b, const int32_t *src, int n) {
    std::vector<int> t(src, src + std::max(n, 0));
    b.upload(t, c.h->stream);
  };
  up(c.ghost_peer, ghost_peer, n_ghost); up(c.ghost_src, ghost_src, n_ghost); up(c.ghost_dst, ghost_dst, n_ghost);
  up(c.lm_rows, lm_rows, c.n_lm);
  CUDA_CHECK(cudaStreamSynchronize(c.h->stream));
  for (int g = 0; g < n_ghost; ++g) require(ghost_peer[g] >= 0 && ghost_peer[g] < world, "ghost owner out of range");
  API_END
}

// `reps` products Q X -> the handle's Q*X buffer, each = barrier + halo pull | local rows | barrier + landmark sum,
// all enqueued back to back on the handle's stream; returns the CUDA-event milliseconds of the whole batch.
extern "C" int cora_b200_peer_product(cora_b200_peer_t *p, int reps, float *ms_total) {
  API_BEGIN
  require(p && reps > 0, "bad argument");
  PeerCtx &c = p->c;
  H *h = c.h;
  require(c.world > 0, "cora_b200_peer_connect has not been called");
  CUDA_CHECK(cudaSetDevice(h->device));
  PeerPtrs P{};
  for (int q = 0; q < c.world; ++q) { P.x[q] = c.peer_x[q]; P.stage[q] = c.peer_stage[q]; P.flags[q] = c.peer_flags[q]; }
  cudaStream_t s = h->stream;
  cudaEvent_t e0, e1;
  CUDA_CHECK(cudaEventCreate(&e0)); CUDA_CHECK(cudaEventCreate(&e1));
  persistent_configure(h, c.r);
  CUDA_CHECK(cudaEventRecord(e0, s));
  for (int i = 0; i < reps; ++i) {
    ++c.epoch;
    k_peer_halo<<<1, 256, 0, s>>>(P, c.world, c.rank, c.epoch, c.r, c.n_ghost, c.ghost_peer.p, c.ghost_src.p, c.ghost_dst.p,
                                  h->ws[V_X].p, c.err.p);
    check_launch(h);
    spmm_persistent(h, c.r, h->ws[V_X].p, h->ws[V_G].p, 1, /*sync=*/false);
    k_peer_reduce<<<1, 256, 0, s>>>(P, c.world, c.rank, c.epoch, c.r, c.n_lm, c.lm_rows.p, c.stage.p, h->ws[V_G].p, c.err.p);
    check_launch(h);
  }
  CUDA_CHECK(cudaEventRecord(e1, s));
  CUDA_CHECK(cudaEventSynchronize(e1));
  float ms = 0.f;
  CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  int err = 0;
  CUDA_CHECK(cudaMemcpy(&err, c.err.p, sizeof(int), cudaMemcpyDeviceToHost));
  if (err) throw Error(CORA_B200_ERUNTIME, "row-partitioned product: a peer did not reach the barrier within 5 s");
  if (ms_total) *ms_total = ms;
  API_END
}

extern "C" int cora_b200_peer_destroy(cora_b200_peer_t *p) {
  if (!p) return CORA_B200_OK;
  cudaSetDevice(p->c.h->device);
  cudaStreamSynchronize(p->c.h->stream);
  for (void *q : p->c.opened)
    if (q) cudaIpcCloseMemHandle(q);
  delete p;
  return CORA_B200_OK;
}
