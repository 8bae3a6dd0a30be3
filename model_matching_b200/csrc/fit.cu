// fit.cu -- rigid transform per (base, congruent quad) and the fused online pipeline.
//
// fit_kernel replaces ComputeRigidTransformation (reference src/stocs.cpp:270-361) and
// stocs_estimator::get_rigid_transform_from_congruent_pair (:871-941): Gram-Schmidt frames on
// the first three correspondences, R = Fp^T Fq, the reference's (R*R).diagonal() orthogonality
// test, T = [R | c1 - R c2], and the un-centred pose with translation
// c1 + c_scene - R (c2 + c_model).  One thread per item; deviations D1 (degenerate input is
// rejected) and D4 (no SVD) as documented in oracle/stocs_oracle.cpp and DESIGN.md.
//
// stocs_b200_run_pipeline chains sample -> congruent -> select -> fit -> score -> reduce on the
// device (run_stocs_estimation, src/stocs_match_one_object.cpp:79-165).
#include <vector>

#include "stocs_ctx.h"

using namespace stocsm;

int stocs_launch_sample(stocs_b200_ctx* ctx, uint64_t seed, uint32_t first_base_no, int n_bases, int* d_ids,
                        float* d_inv, uint8_t* d_valid, cudaStream_t st);
int stocs_congruent_device(stocs_b200_ctx* ctx, int n_bases, const int* d_base_idx4, const float* d_inv2,
                           DevBuf& quads_buf, std::vector<long long>& h_quad_off, cudaStream_t st);

namespace {

__device__ __forceinline__ V3 ld3(const float4* p, int i) { const float4 v = p[i]; return v3(v.x, v.y, v.z); }
__device__ __forceinline__ float comp(const V3& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

__device__ __forceinline__ bool make_frame(V3 a0, V3 a1, V3 a2, V3 f[3]) {
  V3 v1 = sub(a1, a0);
  if (sqnorm(v1) == 0) return false;
  v1 = normalized(v1);
  const V3 d = sub(a2, a0);
  V3 v2 = sub(d, scale(v1, dot(d, v1)));
  if (sqnorm(v2) == 0) return false;
  v2 = normalized(v2);
  f[0] = v1; f[1] = v2; f[2] = cross(v1, v2);
  return true;
}

// item t: base ids base_idx4[4*bmap(t)..], quad quads4[4*qmap(t)..]
struct FitArgs {
  const float4* __restrict__ spos4;
  const float4* __restrict__ mpos4;
  const int* __restrict__ base_idx4;
  const int* __restrict__ quads4;
  const int* __restrict__ item_base;   // optional: item -> base row (NULL: identity)
  const long long* __restrict__ item_quad;  // optional: item -> quad row (NULL: identity)
  float* __restrict__ Tc;
  float* __restrict__ Tw;
  uint8_t* __restrict__ ok;
  long long n;
  float cs[3], cm[3];
  int S, M;
};

__global__ void fit_kernel(FitArgs a) {
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.n) return;
  const int* bid = a.base_idx4 + 4 * (size_t)(a.item_base ? a.item_base[t] : t);
  const int* qd = a.quads4 + 4 * (size_t)(a.item_quad ? a.item_quad[t] : t);
  float* Tc = a.Tc + 16 * (size_t)t;
  float* Tw = a.Tw ? a.Tw + 16 * (size_t)t : nullptr;
  bool good = true;
  for (int k = 0; k < 3; ++k) good = good && bid[k] >= 0 && bid[k] < a.S && qd[k] >= 0 && qd[k] < a.M;
  float R[3][3];
  V3 c1 = v3(0, 0, 0), c2 = v3(0, 0, 0);
  if (good) {
    const V3 p0 = ld3(a.spos4, bid[0]), p1 = ld3(a.spos4, bid[1]), p2 = ld3(a.spos4, bid[2]);
    const V3 q0 = ld3(a.mpos4, qd[0]), q1 = ld3(a.mpos4, qd[1]), q2 = ld3(a.mpos4, qd[2]);
    c1 = divs(add(add(p0, p1), p2), 3.0f);
    c2 = divs(add(add(q0, q1), q2), 3.0f);
    V3 fp[3], fq[3];
    good = make_frame(p0, p1, p2, fp) && make_frame(q0, q1, q2, fq);
    if (good) {
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
          R[i][j] = sum3(comp(fp[0], i) * comp(fq[0], j), comp(fp[1], i) * comp(fq[1], j), comp(fp[2], i) * comp(fq[2], j));
      for (int i = 0; i < 3; ++i) {
        const float rr = sum3(R[i][0] * R[0][i], R[i][1] * R[1][i], R[i][2] * R[2][i]);
        if (rr - 1.0f > 1e-6f) good = false;
        if (rr != rr) good = false;
      }
    }
  }
  if (!good) {
    // rejected items carry NaN so that, if they are scored anyway, they can never win
    const float qnan = __int_as_float(0x7fc00000);
    for (int k = 0; k < 16; ++k) { Tc[k] = qnan; if (Tw) Tw[k] = qnan; }
    a.ok[t] = 0;
    return;
  }
  const V3 nc2 = v3(-c2.x, -c2.y, -c2.z);
  const float c1a[3] = {c1.x, c1.y, c1.z};
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) { Tc[j * 4 + i] = R[i][j]; if (Tw) Tw[j * 4 + i] = R[i][j]; }
    Tc[12 + i] = c1a[i] + sum3(R[i][0] * nc2.x, R[i][1] * nc2.y, R[i][2] * nc2.z);
  }
  Tc[3] = Tc[7] = Tc[11] = 0.f; Tc[15] = 1.f;
  if (Tw) {
    const V3 sa = add(c1, v3(a.cs[0], a.cs[1], a.cs[2]));
    const V3 sb = add(c2, v3(a.cm[0], a.cm[1], a.cm[2]));
    const float saa[3] = {sa.x, sa.y, sa.z};
    for (int i = 0; i < 3; ++i) Tw[12 + i] = saa[i] - sum3(R[i][0] * sb.x, R[i][1] * sb.y, R[i][2] * sb.z);
    Tw[3] = Tw[7] = Tw[11] = 0.f; Tw[15] = 1.f;
  }
  a.ok[t] = 1;
}

// item list of the pipeline: for base b with cnt quads, take all when cnt < max_sets, else
// max_sets quads spread evenly over the (sorted) list: index floor(k * cnt / max_sets).
__global__ void select_items_kernel(const long long* __restrict__ quad_off, const long long* __restrict__ item_off,
                                    int n_bases, int max_sets, int* __restrict__ item_base,
                                    long long* __restrict__ item_quad) {
  const int b = blockIdx.x;
  const long long q0 = quad_off[b], cnt = quad_off[b + 1] - q0;
  const long long o = item_off[b];
  const long long take = cnt < max_sets ? cnt : max_sets;
  for (long long k = threadIdx.x; k < take; k += blockDim.x) {
    const long long src = (cnt < max_sets) ? k : (k * cnt) / max_sets;
    item_base[o + k] = b;
    item_quad[o + k] = q0 + src;
  }
}

}  // namespace

static void fill_fit_args(stocs_b200_ctx* ctx, FitArgs& a) {
  a.spos4 = ctx->d_spos4.as<float4>();
  a.mpos4 = ctx->d_mpos4.as<float4>();
  for (int k = 0; k < 3; ++k) { a.cs[k] = ctx->cs[k]; a.cm[k] = ctx->cm[k]; }
  a.S = ctx->S; a.M = ctx->M;
  a.item_base = nullptr; a.item_quad = nullptr;
}

extern "C" int stocs_b200_fit_transforms(stocs_b200_ctx* ctx, int64_t n, const int32_t* base_idx4, const int32_t* quads4,
                                         float* T_centred16, float* T_world16, uint8_t* ok) {
  if (!ctx) return STOCS_E_ARG;
  if (ctx->S <= 0 || ctx->M <= 0) STOCS_FAIL(ctx, STOCS_E_STATE, "fit_transforms: upload_model and upload_scene first");
  if (n < 0 || (n > 0 && (!base_idx4 || !quads4 || !T_centred16 || !ok))) STOCS_FAIL(ctx, STOCS_E_ARG, "fit_transforms: bad argument");
  if (n == 0) return STOCS_OK;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  STOCS_CUDA(ctx, ctx->d_tmp.ensure((size_t)n * 32));
  STOCS_CUDA(ctx, ctx->d_tmp2.ensure((size_t)n * (128 + 1) + 64));
  int* d_b = ctx->d_tmp.as<int>();
  int* d_q = d_b + 4 * (size_t)n;
  float* d_Tc = ctx->d_tmp2.as<float>();
  float* d_Tw = d_Tc + 16 * (size_t)n;
  uint8_t* d_ok = (uint8_t*)(d_Tw + 16 * (size_t)n);
  STOCS_CUDA(ctx, cudaMemcpyAsync(d_b, base_idx4, (size_t)n * 16, cudaMemcpyHostToDevice, st));
  STOCS_CUDA(ctx, cudaMemcpyAsync(d_q, quads4, (size_t)n * 16, cudaMemcpyHostToDevice, st));
  FitArgs a;
  fill_fit_args(ctx, a);
  a.base_idx4 = d_b; a.quads4 = d_q; a.Tc = d_Tc; a.Tw = d_Tw; a.ok = d_ok; a.n = n;
  fit_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a);
  STOCS_CUDA(ctx, cudaGetLastError());
  STOCS_CUDA(ctx, cudaMemcpyAsync(T_centred16, d_Tc, (size_t)n * 64, cudaMemcpyDeviceToHost, st));
  if (T_world16) STOCS_CUDA(ctx, cudaMemcpyAsync(T_world16, d_Tw, (size_t)n * 64, cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaMemcpyAsync(ok, d_ok, (size_t)n, cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaStreamSynchronize(st));
  return STOCS_OK;
}

int stocs_launch_sample_instance(stocs_b200_ctx* ctx, uint64_t seed, int base_num, float dispersion, int* d_ids,
                                 float* d_inv, uint8_t* d_valid, cudaStream_t st);  // sample_instance.cu

// mode 0: class-mode bases (independent, one launch for all); mode 1: instance-mode bases (sequentially
// coupled: one launch per base, base numbers 1..n_bases, enqueued back to back without host round trips)
static int run_pipeline_impl(stocs_b200_ctx* ctx, uint64_t seed, int n_bases, int max_sets, int mode, float dispersion,
                             stocs_b200_pipeline_result* result) {
  if (!ctx) return STOCS_E_ARG;
  if (ctx->S <= 0 || ctx->M <= 0) STOCS_FAIL(ctx, STOCS_E_STATE, "run_pipeline: upload_model and upload_scene first");
  if (n_bases <= 0 || max_sets <= 0 || !result) STOCS_FAIL(ctx, STOCS_E_ARG, "run_pipeline: bad argument");
  if (mode == 1 && n_bases > 255) STOCS_FAIL(ctx, STOCS_E_ARG, "run_pipeline: instance mode numbers its bases 1..255");
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  StageTrace tr(st);
  memset(result, 0, sizeof(*result));
  result->best_index = -1;
  result->best_base = -1;
  // 1. bases
  DevBuf& d_bases = ctx->pool[POOL_PIPE_BASES];
  STOCS_CUDA(ctx, d_bases.ensure((size_t)n_bases * 25 + 64));
  int* d_ids = d_bases.as<int>();
  float* d_inv = (float*)(d_ids + 4 * (size_t)n_bases);
  uint8_t* d_valid = (uint8_t*)(d_inv + 2 * (size_t)n_bases);
  int rc = STOCS_OK;
  if (mode == 0) {
    rc = stocs_launch_sample(ctx, seed, 0, n_bases, d_ids, d_inv, d_valid, st);
  } else {
    STOCS_CUDA(ctx, cudaMemsetAsync(d_valid, 0, (size_t)n_bases, st));
    for (int b = 0; b < n_bases && rc == STOCS_OK; ++b)
      rc = stocs_launch_sample_instance(ctx, seed, b + 1, dispersion, d_ids + 4 * (size_t)b, d_inv + 2 * (size_t)b, d_valid + b, st);
  }
  if (rc) return rc;
  std::vector<int> h_ids((size_t)4 * n_bases);
  std::vector<float> h_inv((size_t)2 * n_bases);
  std::vector<uint8_t> h_valid((size_t)n_bases);
  STOCS_CUDA(ctx, cudaMemcpyAsync(h_ids.data(), d_ids, (size_t)n_bases * 16, cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaMemcpyAsync(h_inv.data(), d_inv, (size_t)n_bases * 8, cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaMemcpyAsync(h_valid.data(), d_valid, (size_t)n_bases, cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaStreamSynchronize(st));
  tr.mark("pipeline: sample bases");
  // keep the valid bases, in order (base_set of the reference driver)
  std::vector<int> v_ids; std::vector<float> v_inv;
  for (int b = 0; b < n_bases; ++b)
    if (h_valid[b]) {
      for (int k = 0; k < 4; ++k) v_ids.push_back(h_ids[4 * (size_t)b + k]);
      v_inv.push_back(h_inv[2 * (size_t)b]); v_inv.push_back(h_inv[2 * (size_t)b + 1]);
    }
  const int nv = (int)(v_ids.size() / 4);
  result->n_valid_bases = nv;
  if (nv == 0) return STOCS_OK;
  STOCS_CUDA(ctx, cudaMemcpyAsync(d_ids, v_ids.data(), (size_t)nv * 16, cudaMemcpyHostToDevice, st));
  STOCS_CUDA(ctx, cudaMemcpyAsync(d_inv, v_inv.data(), (size_t)nv * 8, cudaMemcpyHostToDevice, st));
  // 2. congruent sets
  DevBuf& d_quads = ctx->pool[POOL_CONG_QUADS];
  std::vector<long long> quad_off;
  rc = stocs_congruent_device(ctx, nv, d_ids, d_inv, d_quads, quad_off, st);
  if (rc) return rc;
  result->n_congruent_sets = quad_off[nv];
  tr.mark("pipeline: congruent sets");
  // 3. at most max_sets transforms per base
  std::vector<long long> item_off((size_t)nv + 1, 0);
  for (int b = 0; b < nv; ++b) {
    const long long cnt = quad_off[b + 1] - quad_off[b];
    item_off[b + 1] = item_off[b] + (cnt < max_sets ? cnt : max_sets);
  }
  const long long n_items = item_off[nv];
  if (n_items == 0) return STOCS_OK;
  DevBuf &d_off = ctx->pool[POOL_PIPE_OFF], &d_items = ctx->pool[POOL_PIPE_ITEMS], &d_fit = ctx->pool[POOL_PIPE_FIT];
  auto cleanup = [&]() {};  // pool slots persist
#define PL(call) do { cudaError_t _e = (call); if (_e != cudaSuccess) { ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e); cleanup(); return STOCS_E_CUDA; } } while (0)
  PL(d_off.ensure((size_t)(nv + 1) * 16));
  long long* d_qoff = d_off.as<long long>();
  long long* d_ioff = d_qoff + (nv + 1);
  PL(cudaMemcpyAsync(d_qoff, quad_off.data(), (size_t)(nv + 1) * 8, cudaMemcpyHostToDevice, st));
  PL(cudaMemcpyAsync(d_ioff, item_off.data(), (size_t)(nv + 1) * 8, cudaMemcpyHostToDevice, st));
  PL(d_items.ensure((size_t)n_items * 12));
  long long* d_item_quad = d_items.as<long long>();
  int* d_item_base = (int*)(d_item_quad + n_items);
  select_items_kernel<<<nv, 128, 0, st>>>(d_qoff, d_ioff, nv, max_sets, d_item_base, d_item_quad);
  PL(d_fit.ensure((size_t)n_items * (64 + 64 + 1 + 4 + 4) + 1024));
  float* d_Tc = d_fit.as<float>();
  float* d_Tw = d_Tc + 16 * (size_t)n_items;
  float* d_lcp = d_Tw + 16 * (size_t)n_items;
  int* d_inl = (int*)(d_lcp + n_items);
  uint8_t* d_ok = (uint8_t*)(d_inl + n_items);
  FitArgs a;
  fill_fit_args(ctx, a);
  a.base_idx4 = d_ids; a.quads4 = d_quads.as<int>(); a.item_base = d_item_base; a.item_quad = d_item_quad;
  a.Tc = d_Tc; a.Tw = d_Tw; a.ok = d_ok; a.n = n_items;
  fit_kernel<<<(unsigned)((n_items + 127) / 128), 128, 0, st>>>(a);
  tr.mark("pipeline: select + fit");
  // 4. score + 5. best
  rc = stocs_launch_score(ctx, d_Tc, n_items, d_lcp, d_inl, st, true);
  if (rc) { cleanup(); return rc; }
  long long* d_bi = (long long*)(ctx->d_small.as<char>() + 512);
  float* d_bv = (float*)(ctx->d_small.as<char>() + 512 + 256);
  rc = stocs_launch_topk(ctx, d_lcp, n_items, 1, 0, (int64_t*)d_bi, d_bv, st);
  if (rc) { cleanup(); return rc; }
  long long bi = -1; float bv = 0.f;
  std::vector<uint8_t> h_ok((size_t)n_items);
  PL(cudaMemcpyAsync(&bi, d_bi, 8, cudaMemcpyDeviceToHost, st));
  PL(cudaMemcpyAsync(&bv, d_bv, 4, cudaMemcpyDeviceToHost, st));
  PL(cudaMemcpyAsync(h_ok.data(), d_ok, (size_t)n_items, cudaMemcpyDeviceToHost, st));
  PL(cudaStreamSynchronize(st));
  tr.mark("pipeline: score + best");
  long long n_ok = 0, rank_of_best = -1;
  for (long long i = 0; i < n_items; ++i) { if (i == bi) rank_of_best = n_ok; n_ok += h_ok[i] ? 1 : 0; }
  result->n_transforms = n_ok;
  result->best_lcp = bv;
  if (bi >= 0) {
    result->best_index = rank_of_best;  // index into the list of pushed transforms, as in the reference
    int bb = 0;
    while (bb + 1 < nv && item_off[bb + 1] <= bi) ++bb;
    result->best_base = bb;
    PL(cudaMemcpyAsync(result->best_T_centred, d_Tc + 16 * bi, 64, cudaMemcpyDeviceToHost, st));
    PL(cudaMemcpyAsync(result->best_T_world, d_Tw + 16 * bi, 64, cudaMemcpyDeviceToHost, st));
    PL(cudaStreamSynchronize(st));
  }
#undef PL
  cleanup();
  return STOCS_OK;
}

extern "C" int stocs_b200_run_pipeline(stocs_b200_ctx* ctx, uint64_t seed, int n_bases, int max_sets,
                                       stocs_b200_pipeline_result* result) {
  return run_pipeline_impl(ctx, seed, n_bases, max_sets, 0, 0.f, result);
}

extern "C" int stocs_b200_run_pipeline_instance(stocs_b200_ctx* ctx, uint64_t seed, int n_bases, int max_sets,
                                                float dispersion, stocs_b200_pipeline_result* result) {
  return run_pipeline_impl(ctx, seed, n_bases, max_sets, 1, dispersion, result);
}
