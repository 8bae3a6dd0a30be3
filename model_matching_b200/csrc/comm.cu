// comm.cu -- hypothesis-sharded scoring across GPUs.
//
// The reference scores every hypothesis in one sequential loop and keeps the first strict maximum
// (src/stocs.cpp:990-998).  Hypotheses are independent, so each GPU scores a contiguous block of
// the list against its own replica of the scene index; the per-GPU K best are packed as 64-byte
// records {lcp, inliers, global index, 3x4 transform} (reduce.cu: topk_kernel), ONE
// ncclAllGather moves nranks*K records (16 KB at 8 GPUs, K = 32) over NVLink, and a single-CTA
// kernel ranks them by (lcp descending, global index ascending): record 0 is the reference's
// winner over the whole list.  The collective is latency-bound; nothing here is worth fusing into
// the scoring kernel (it moves 2 KB per rank once per object).
//
// NCCL is bound at run time (dlopen) so that single-GPU users of libstocs_b200.so do not need it
// and so that a host process that already carries an NCCL (PyTorch bundles its own libnccl.so.2)
// shares that copy instead of loading a second one.
#include <dlfcn.h>
#include <nccl.h>

#include <cstring>
#include <mutex>
#include <new>

#include "stocs_ctx.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string err;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

// returns NULL on success, else a message
const char* load_nccl() {
  std::lock_guard<std::mutex> lk(g_nccl_mu);
  if (g_nccl.handle) return nullptr;
  const char* names[] = {getenv("STOCS_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* h = nullptr;
  for (const char* n : names) {
    if (!n || !*n) continue;
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    g_nccl.err = std::string("cannot load NCCL (libnccl.so.2; set STOCS_NCCL_LIB): ") + (dlerror() ? dlerror() : "");
    return g_nccl.err.c_str();
  }
  bool ok = true;
  auto sym = [&](const char* s) { void* p = dlsym(h, s); if (!p) { ok = false; g_nccl.err = std::string("NCCL symbol missing: ") + s; } return p; };
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
  g_nccl.CommInitAll = (decltype(g_nccl.CommInitAll))sym("ncclCommInitAll");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
  g_nccl.AllGather = (decltype(g_nccl.AllGather))sym("ncclAllGather");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
  if (!ok) { dlclose(h); return g_nccl.err.c_str(); }
  g_nccl.handle = h;
  return nullptr;
}

#define STOCS_NCCL(ctx, call)                                                              \
  do {                                                                                     \
    ncclResult_t _r = (call);                                                              \
    if (_r != ncclSuccess) {                                                               \
      (ctx)->err = std::string(#call) + ": " + g_nccl.GetErrorString(_r);                  \
      return STOCS_E_NCCL;                                                                 \
    }                                                                                      \
  } while (0)

// Rank n records (n <= 2048) by (lcp descending, index ascending), keep the K best.  Every record
// counts the records that beat it; that count is its output slot.  Empty records (index < 0 or
// lcp <= 0) never win; unused output slots are written as empty.  One CTA.
__global__ void __launch_bounds__(1024) merge_records_kernel(const stocs_b200_record* __restrict__ in, int n, int K,
                                                             stocs_b200_record* __restrict__ out) {
  __shared__ float s_lcp[2048];
  __shared__ long long s_idx[2048];
  __shared__ int s_valid;
  if (threadIdx.x == 0) s_valid = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = in[i].lcp;
    const long long ix = in[i].index;
    const bool ok = ix >= 0 && v > 0.f;
    s_lcp[i] = ok ? v : 0.f;
    s_idx[i] = ok ? ix : -1;
    if (ok) atomicAdd(&s_valid, 1);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long ix = s_idx[i];
    if (ix < 0) continue;
    const float v = s_lcp[i];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const long long jx = s_idx[j];
      if (jx < 0) continue;
      const float w = s_lcp[j];
      rank += (w > v || (w == v && jx < ix)) ? 1 : 0;
    }
    if (rank < K) out[rank] = in[i];
  }
  const int nv = s_valid;
  for (int k = nv + threadIdx.x; k < K; k += blockDim.x) {
    stocs_b200_record r;
    r.lcp = 0.f; r.inliers = 0; r.index = -1;
    for (int t = 0; t < 12; ++t) r.T[t] = 0.f;
    out[k] = r;
  }
}

// local block -> K packed records in d_send (device).  Enqueues on st.
int enqueue_local_topk(stocs_b200_ctx* ctx, const float* d_T, int64_t H_local, int64_t index_offset, int K,
                       float* d_lcp, int32_t* d_inl, stocs_b200_record* d_send, cudaStream_t st, bool time_it) {
  if (H_local > 0) {
    int rc = stocs_launch_score(ctx, d_T, H_local, d_lcp, d_inl, st, time_it, 0, nullptr, stocs_is_host_memory(d_T));
    if (rc) return rc;
  }
  // H_local == 0: the top-K kernels run over an empty array and emit K empty records
  return stocs_launch_topk(ctx, d_lcp, H_local, K, index_offset, nullptr, nullptr, st, d_T, d_inl, d_send);
}

}  // namespace

extern "C" {

void stocs_b200_shard_range(int64_t H, int rank, int nranks, int64_t* lo, int64_t* hi) {
  if (nranks < 1) nranks = 1;
  const int64_t per = (H + nranks - 1) / nranks;
  int64_t a = (int64_t)rank * per;
  if (a > H) a = H;
  int64_t b = a + per;
  if (b > H) b = H;
  if (lo) *lo = a;
  if (hi) *hi = b;
}

int stocs_b200_comm_unique_id(void* id128) {
  if (!id128) return STOCS_E_ARG;
  static_assert(sizeof(ncclUniqueId) == STOCS_B200_UNIQUE_ID_BYTES, "ncclUniqueId is 128 bytes");
  if (load_nccl()) return STOCS_E_NCCL;
  ncclUniqueId id;
  if (g_nccl.GetUniqueId(&id) != ncclSuccess) return STOCS_E_NCCL;
  memcpy(id128, &id, sizeof(id));
  return STOCS_OK;
}

int stocs_b200_comm_init(stocs_b200_ctx* ctx, const void* id128, int rank, int nranks) {
  if (!ctx) return STOCS_E_ARG;
  if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) STOCS_FAIL(ctx, STOCS_E_ARG, "comm_init: bad argument");
  if (nranks * 32 > 2048) STOCS_FAIL(ctx, STOCS_E_ARG, "comm_init: at most 64 ranks");
  if (ctx->comm) STOCS_FAIL(ctx, STOCS_E_STATE, "comm_init: communicator already initialised");
  if (const char* e = load_nccl()) STOCS_FAIL(ctx, STOCS_E_NCCL, e);
  cudaSetDevice(ctx->device);
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  ncclComm_t comm = nullptr;
  STOCS_NCCL(ctx, g_nccl.CommInitRank(&comm, nranks, id, rank));
  ctx->comm = comm;
  ctx->comm_rank = rank;
  ctx->comm_nranks = nranks;
  return STOCS_OK;
}

int stocs_b200_comm_destroy(stocs_b200_ctx* ctx) {
  if (!ctx) return STOCS_E_ARG;
  if (ctx->comm) {
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->comm_rank = 0;
  ctx->comm_nranks = 1;
  return STOCS_OK;
}

int stocs_b200_score_sharded_device(stocs_b200_ctx* ctx, const float* d_T16_local, int64_t H_local,
                                    int64_t index_offset, int K, stocs_b200_record* d_topk_out, void* stream) {
  if (!ctx) return STOCS_E_ARG;
  if (ctx->S <= 0 || ctx->M <= 0) STOCS_FAIL(ctx, STOCS_E_STATE, "score_sharded: upload_model and upload_scene first");
  if (H_local < 0 || K < 1 || K > 32 || !d_topk_out || (H_local > 0 && !d_T16_local) || index_offset < 0)
    STOCS_FAIL(ctx, STOCS_E_ARG, "score_sharded: bad argument");
  cudaSetDevice(ctx->device);
  cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
  DevBuf &b_lcp = ctx->pool[POOL_SHARD_LCP], &b_inl = ctx->pool[POOL_SHARD_INL];
  DevBuf &b_send = ctx->pool[POOL_COMM_SEND], &b_recv = ctx->pool[POOL_COMM_RECV];
  STOCS_CUDA(ctx, b_lcp.ensure((size_t)(H_local ? H_local : 1) * 4));
  STOCS_CUDA(ctx, b_inl.ensure((size_t)(H_local ? H_local : 1) * 4));
  const int nranks = ctx->comm ? ctx->comm_nranks : 1;
  if (nranks == 1)  // no collective: the local top-K is the global one
    return enqueue_local_topk(ctx, d_T16_local, H_local, index_offset, K, b_lcp.as<float>(), b_inl.as<int32_t>(),
                              d_topk_out, st, true);
  STOCS_CUDA(ctx, b_send.ensure((size_t)K * sizeof(stocs_b200_record)));
  STOCS_CUDA(ctx, b_recv.ensure((size_t)K * nranks * sizeof(stocs_b200_record)));
  int rc = enqueue_local_topk(ctx, d_T16_local, H_local, index_offset, K, b_lcp.as<float>(), b_inl.as<int32_t>(),
                              b_send.as<stocs_b200_record>(), st, true);
  if (rc) return rc;
  // the path's one collective
  STOCS_NCCL(ctx, g_nccl.AllGather(b_send.p, b_recv.p, (size_t)K * sizeof(stocs_b200_record), ncclInt8,
                                   (ncclComm_t)ctx->comm, st));
  merge_records_kernel<<<1, 1024, 0, st>>>(b_recv.as<stocs_b200_record>(), K * nranks, K, d_topk_out);
  STOCS_CUDA(ctx, cudaGetLastError());
  return STOCS_OK;
}

int stocs_b200_score_sharded(stocs_b200_ctx* ctx, const float* T16_local, int64_t H_local, int64_t index_offset, int K,
                             stocs_b200_record* topk_out, float* lcp_local, int32_t* inliers_local) {
  if (!ctx) return STOCS_E_ARG;
  if (H_local < 0 || K < 1 || K > 32 || !topk_out || (H_local > 0 && !T16_local) || index_offset < 0)
    STOCS_FAIL(ctx, STOCS_E_ARG, "score_sharded: bad argument");
  if (ctx->S <= 0 || ctx->M <= 0) STOCS_FAIL(ctx, STOCS_E_STATE, "score_sharded: upload_model and upload_scene first");
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  // 1. the local block through the host-buffer scoring call (page-locked transforms read in place,
  //    pageable ones staged in overlapped chunks); per-hypothesis results come back as in the
  //    single-GPU drop-in, into the caller's arrays or into a context-owned page-locked scratch
  if (H_local > 0 && !lcp_local) {
    if (ctx->h_pinned_bytes < (size_t)H_local * 4) {
      if (ctx->h_pinned) cudaFreeHost(ctx->h_pinned);
      ctx->h_pinned = nullptr; ctx->h_pinned_bytes = 0;
      STOCS_CUDA(ctx, cudaHostAlloc(&ctx->h_pinned, (size_t)H_local * 4, cudaHostAllocDefault));
      ctx->h_pinned_bytes = (size_t)H_local * 4;
    }
    lcp_local = (float*)ctx->h_pinned;
  }
  const int nranks = ctx->comm ? ctx->comm_nranks : 1;
  DevBuf &b_send = ctx->pool[POOL_COMM_SEND], &b_recv = ctx->pool[POOL_COMM_RECV], &b_out = ctx->pool[POOL_COMM_OUT];
  STOCS_CUDA(ctx, b_send.ensure((size_t)K * sizeof(stocs_b200_record)));
  STOCS_CUDA(ctx, b_recv.ensure((size_t)K * nranks * sizeof(stocs_b200_record)));
  STOCS_CUDA(ctx, b_out.ensure((size_t)K * sizeof(stocs_b200_record)));
  stocs_b200_record* d_local = nranks > 1 ? b_send.as<stocs_b200_record>() : b_out.as<stocs_b200_record>();
  // Page-locked, device-mapped transforms: ONE pass with ONE synchronisation.  The kernel reads the
  // transforms in place over PCIe; then the per-hypothesis results go back on the context stream
  // while, on the second stream, the K best are packed, all-gathered and merged.  (Going through
  // stocs_b200_score_lcp first costs its own top-32 reduction and a second synchronisation:
  // 0.58 -> 0.50 ms per step at 125 000 hypotheses per rank on 8 GPUs.)
  if (H_local > 0 && !getenv("STOCS_NO_ZERO_COPY")) {
    cudaPointerAttributes pa{};
    const bool mapped = cudaPointerGetAttributes(&pa, T16_local) == cudaSuccess && pa.type == cudaMemoryTypeHost &&
                        pa.devicePointer != nullptr;
    if (!mapped) cudaGetLastError();
    if (mapped) {
      const float* d_T = (const float*)pa.devicePointer;
      STOCS_CUDA(ctx, ctx->d_lcp.ensure((size_t)H_local * 4));
      STOCS_CUDA(ctx, ctx->d_inl.ensure((size_t)H_local * 4));
      ctx->top_valid = false;
      int rc = stocs_launch_score(ctx, d_T, H_local, ctx->d_lcp.as<float>(), ctx->d_inl.as<int32_t>(), st, true, 0, nullptr, true);
      if (rc) return rc;
      ctx->last_H = H_local;
      ctx->last_T_dev = d_T;
      cudaStream_t aux = ctx->aux_stream;
      STOCS_CUDA(ctx, cudaEventRecord(ctx->join_ev[0], st));
      STOCS_CUDA(ctx, cudaStreamWaitEvent(aux, ctx->join_ev[0], 0));
      STOCS_CUDA(ctx, cudaMemcpyAsync(lcp_local, ctx->d_lcp.p, (size_t)H_local * 4, cudaMemcpyDeviceToHost, st));
      if (inliers_local)
        STOCS_CUDA(ctx, cudaMemcpyAsync(inliers_local, ctx->d_inl.p, (size_t)H_local * 4, cudaMemcpyDeviceToHost, st));
      rc = stocs_launch_topk(ctx, ctx->d_lcp.as<float>(), H_local, K, index_offset, nullptr, nullptr, aux, d_T,
                             ctx->d_inl.as<int32_t>(), d_local);
      if (rc) return rc;
      if (nranks > 1) {
        STOCS_NCCL(ctx, g_nccl.AllGather(b_send.p, b_recv.p, (size_t)K * sizeof(stocs_b200_record), ncclInt8,
                                         (ncclComm_t)ctx->comm, aux));
        merge_records_kernel<<<1, 1024, 0, aux>>>(b_recv.as<stocs_b200_record>(), K * nranks, K, b_out.as<stocs_b200_record>());
        STOCS_CUDA(ctx, cudaGetLastError());
      }
      STOCS_CUDA(ctx, cudaMemcpyAsync(topk_out, b_out.p, (size_t)K * sizeof(stocs_b200_record), cudaMemcpyDeviceToHost, aux));
      STOCS_CUDA(ctx, cudaEventRecord(ctx->join_ev[1], aux));
      STOCS_CUDA(ctx, cudaStreamWaitEvent(st, ctx->join_ev[1], 0));
      STOCS_CUDA(ctx, cudaStreamSynchronize(st));
      return STOCS_OK;
    }
  }
  if (H_local > 0) {
    int rc = stocs_b200_score_lcp(ctx, T16_local, H_local, lcp_local, inliers_local);
    if (rc) return rc;
  }
  // 2. K best of the resident results as records, 3. ONE all-gather, 4. merge
  STOCS_CUDA(ctx, ctx->d_lcp.ensure(4));
  int rc = stocs_launch_topk(ctx, ctx->d_lcp.as<float>(), H_local, K, index_offset, nullptr, nullptr, st,
                             H_local > 0 ? ctx->last_T_dev : nullptr, H_local > 0 ? ctx->d_inl.as<int32_t>() : nullptr, d_local);
  if (rc) return rc;
  if (nranks > 1) {
    STOCS_NCCL(ctx, g_nccl.AllGather(b_send.p, b_recv.p, (size_t)K * sizeof(stocs_b200_record), ncclInt8,
                                     (ncclComm_t)ctx->comm, st));
    merge_records_kernel<<<1, 1024, 0, st>>>(b_recv.as<stocs_b200_record>(), K * nranks, K, b_out.as<stocs_b200_record>());
    STOCS_CUDA(ctx, cudaGetLastError());
  }
  STOCS_CUDA(ctx, cudaMemcpyAsync(topk_out, b_out.p, (size_t)K * sizeof(stocs_b200_record), cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaStreamSynchronize(st));
  return STOCS_OK;
}

}  // extern "C"

// ---- single-process group -----------------------------------------------------------------------
struct stocs_b200_group {
  std::vector<stocs_b200_ctx*> ctx;
  std::string err;
};

static int group_fail(stocs_b200_group* g, int i, int rc) {
  g->err = "device " + std::to_string(g->ctx[i]->device) + ": " + stocs_b200_last_error(g->ctx[i]);
  return rc;
}

extern "C" {

int stocs_b200_group_create(stocs_b200_group** out, const int* device_ids, int n_dev) {
  if (!out || !device_ids || n_dev < 1 || n_dev > 64) return STOCS_E_ARG;
  *out = nullptr;
  stocs_b200_group* g = new (std::nothrow) stocs_b200_group();
  if (!g) return STOCS_E_ARG;
  for (int i = 0; i < n_dev; ++i) {
    stocs_b200_ctx* c = nullptr;
    int rc = stocs_b200_create(&c, device_ids[i]);
    if (rc != STOCS_OK) { stocs_b200_group_destroy(g); return rc; }
    g->ctx.push_back(c);
  }
  if (n_dev > 1) {
    if (load_nccl()) { stocs_b200_group_destroy(g); return STOCS_E_NCCL; }
    std::vector<ncclComm_t> comms((size_t)n_dev);
    if (g_nccl.CommInitAll(comms.data(), n_dev, device_ids) != ncclSuccess) { stocs_b200_group_destroy(g); return STOCS_E_NCCL; }
    for (int i = 0; i < n_dev; ++i) {
      g->ctx[i]->comm = comms[i];
      g->ctx[i]->comm_rank = i;
      g->ctx[i]->comm_nranks = n_dev;
    }
  }
  *out = g;
  return STOCS_OK;
}

void stocs_b200_group_destroy(stocs_b200_group* g) {
  if (!g) return;
  for (stocs_b200_ctx* c : g->ctx) { stocs_b200_comm_destroy(c); stocs_b200_destroy(c); }
  delete g;
}

int stocs_b200_group_size(stocs_b200_group* g) { return g ? (int)g->ctx.size() : 0; }
stocs_b200_ctx* stocs_b200_group_ctx(stocs_b200_group* g, int i) {
  return (g && i >= 0 && i < (int)g->ctx.size()) ? g->ctx[i] : nullptr;
}
const char* stocs_b200_group_last_error(stocs_b200_group* g) { return g ? g->err.c_str() : "no group"; }

int stocs_b200_group_set_params(stocs_b200_group* g, float distance_threshold, int tr, int rot) {
  if (!g) return STOCS_E_ARG;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    int rc = stocs_b200_set_params(g->ctx[i], distance_threshold, tr, rot);
    if (rc) return group_fail(g, (int)i, rc);
  }
  return STOCS_OK;
}

int stocs_b200_group_upload_model(stocs_b200_group* g, const float* pos3, const float* nrm3, int M) {
  if (!g) return STOCS_E_ARG;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    int rc = stocs_b200_upload_model(g->ctx[i], pos3, nrm3, M);
    if (rc) return group_fail(g, (int)i, rc);
  }
  return STOCS_OK;
}

int stocs_b200_group_upload_scene(stocs_b200_group* g, const float* pos3, const float* nrm3, const float* cls,
                                  const int32_t* pixel_rc, int S) {
  if (!g) return STOCS_E_ARG;
  for (size_t i = 0; i < g->ctx.size(); ++i) {
    int rc = stocs_b200_upload_scene(g->ctx[i], pos3, nrm3, cls, pixel_rc, S);
    if (rc) return group_fail(g, (int)i, rc);
  }
  return STOCS_OK;
}

int stocs_b200_group_score_best(stocs_b200_group* g, const float* T16, int64_t H, int K, stocs_b200_record* topk_out,
                                float* lcp, int32_t* inliers) {
  if (!g || g->ctx.empty()) return STOCS_E_ARG;
  if (H < 0 || K < 1 || K > 32 || !topk_out || (H > 0 && !T16)) { g->err = "group_score_best: bad argument"; return STOCS_E_ARG; }
  const int n = (int)g->ctx.size();
  const size_t rec_bytes = (size_t)K * sizeof(stocs_b200_record);
  std::vector<int64_t> lo((size_t)n), hi((size_t)n);
  // 1. every device: stage its block, score, local top-K records (all asynchronous)
  for (int i = 0; i < n; ++i) {
    stocs_b200_ctx* c = g->ctx[i];
    if (c->S <= 0 || c->M <= 0) { g->err = "group_score_best: upload_model and upload_scene first"; return STOCS_E_STATE; }
    stocs_b200_shard_range(H, i, n, &lo[i], &hi[i]);
    const int64_t hl = hi[i] - lo[i];
    cudaSetDevice(c->device);
    cudaStream_t st = c->stream;
    DevBuf &b_T = c->pool[POOL_SHARD_T], &b_lcp = c->pool[POOL_SHARD_LCP], &b_inl = c->pool[POOL_SHARD_INL];
    DevBuf &b_send = c->pool[POOL_COMM_SEND], &b_recv = c->pool[POOL_COMM_RECV], &b_out = c->pool[POOL_COMM_OUT];
    cudaError_t e = b_T.ensure((size_t)(hl ? hl : 1) * 64);
    if (e == cudaSuccess) e = b_lcp.ensure((size_t)(hl ? hl : 1) * 4);
    if (e == cudaSuccess) e = b_inl.ensure((size_t)(hl ? hl : 1) * 4);
    if (e == cudaSuccess) e = b_send.ensure(rec_bytes);
    if (e == cudaSuccess) e = b_recv.ensure(rec_bytes * n);
    if (e == cudaSuccess) e = b_out.ensure(rec_bytes);
    if (e == cudaSuccess && hl > 0)
      e = cudaMemcpyAsync(b_T.p, T16 + 16 * lo[i], (size_t)hl * 64, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { c->err = std::string("group_score_best: ") + cudaGetErrorString(e); return group_fail(g, i, STOCS_E_CUDA); }
    int rc = enqueue_local_topk(c, b_T.as<float>(), hl, lo[i], K, b_lcp.as<float>(), b_inl.as<int32_t>(),
                                n > 1 ? b_send.as<stocs_b200_record>() : b_out.as<stocs_b200_record>(), st, false);
    if (rc) return group_fail(g, i, rc);
  }
  // 2. ONE all-gather (grouped: one call per device of this process)
  if (n > 1) {
    if (g_nccl.GroupStart() != ncclSuccess) { g->err = "ncclGroupStart failed"; return STOCS_E_NCCL; }
    for (int i = 0; i < n; ++i) {
      stocs_b200_ctx* c = g->ctx[i];
      ncclResult_t r = g_nccl.AllGather(c->pool[POOL_COMM_SEND].p, c->pool[POOL_COMM_RECV].p, rec_bytes, ncclInt8,
                                        (ncclComm_t)c->comm, c->stream);
      if (r != ncclSuccess) { g_nccl.GroupEnd(); g->err = std::string("ncclAllGather: ") + g_nccl.GetErrorString(r); return STOCS_E_NCCL; }
    }
    if (g_nccl.GroupEnd() != ncclSuccess) { g->err = "ncclGroupEnd failed"; return STOCS_E_NCCL; }
    // 3. merge (device 0 is the one whose answer is returned; every device holds the same records)
    stocs_b200_ctx* c0 = g->ctx[0];
    cudaSetDevice(c0->device);
    merge_records_kernel<<<1, 1024, 0, c0->stream>>>(c0->pool[POOL_COMM_RECV].as<stocs_b200_record>(), K * n, K,
                                                     c0->pool[POOL_COMM_OUT].as<stocs_b200_record>());
  }
  // 4. results
  for (int i = 0; i < n; ++i) {
    stocs_b200_ctx* c = g->ctx[i];
    const int64_t hl = hi[i] - lo[i];
    cudaSetDevice(c->device);
    cudaError_t e = cudaSuccess;
    if (i == 0) e = cudaMemcpyAsync(topk_out, c->pool[POOL_COMM_OUT].p, rec_bytes, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && lcp && hl > 0)
      e = cudaMemcpyAsync(lcp + lo[i], c->pool[POOL_SHARD_LCP].p, (size_t)hl * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess && inliers && hl > 0)
      e = cudaMemcpyAsync(inliers + lo[i], c->pool[POOL_SHARD_INL].p, (size_t)hl * 4, cudaMemcpyDeviceToHost, c->stream);
    if (e != cudaSuccess) { c->err = std::string("group_score_best: ") + cudaGetErrorString(e); return group_fail(g, i, STOCS_E_CUDA); }
  }
  for (int i = 0; i < n; ++i) {
    stocs_b200_ctx* c = g->ctx[i];
    cudaSetDevice(c->device);
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) { c->err = std::string("group_score_best: ") + cudaGetErrorString(e); return group_fail(g, i, STOCS_E_CUDA); }
  }
  return STOCS_OK;
}

}  // extern "C"
