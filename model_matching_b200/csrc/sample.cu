// sample.cu -- probability-weighted 4-point base sampling, class mode.
//
// Replaces stocs_estimator::sample_class_base (reference src/stocs.cpp:363-519),
// sample_point_from_distribution (:133-148), try_sampled_base (:224-268) and
// segment_distance_and_invariants (:155-222).  One CTA (32 warps) per base, bases are
// independent in class mode.  Each of the four draws is an exact categorical draw over
// fixed-point weights floor(p * 2^40): integer sums are associative, so the block-parallel
// reduction + the selected warp's scan pick exactly the index a sequential CDF walk picks
// (documented deviation D3: counter-based Philox stream instead of a wall-clock seed).
// Between draws one pass over the scene zeroes candidates by the reference's predicates
// (PPF key present in the model map, internal angle >= 30 deg, coplanarity <= 0.015,
// >= 0.01 m from the chosen points); survivors are kept as one bit per point.
#include "sample_common.cuh"
#include "stocs_ctx.h"

using namespace stocsm;
using namespace stocs_sample;

namespace {

struct SampleArgs {
  const float4* __restrict__ spos4;
  const float4* __restrict__ sattr;
  int S;
  PpfView ppf;
  unsigned long long seed;
  uint32_t first_base;
  uint32_t* alive;  // n_bases * words
  int words;
  int* out_ids;
  float* out_inv;
  uint8_t* out_valid;
};

constexpr int kChunk = 16384;   // scene points per candidate-list round (uint16 offsets, 32 KB of shared memory)

__global__ void __launch_bounds__(1024) sample_bases_kernel(SampleArgs a) {
  const int base = blockIdx.x;
  const uint32_t base_no = a.first_base + (uint32_t)base;
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  uint32_t* alive = a.alive + (size_t)base * a.words;
  const int tiles = a.words;
  const int tpw = (tiles + 31) / 32;
  const int t0 = w * tpw, t1 = min(tiles, t0 + tpw);
  __shared__ unsigned long long s_wsum[32];
  __shared__ unsigned long long s_rem;
  __shared__ int s_b[4];
  __shared__ int s_pw;
  __shared__ uint16_t s_list[kChunk];
  __shared__ uint32_t s_alive[kChunk / 32];
  __shared__ int s_n;

  V3 pb[3], nb[3];
  V3 v_1 = v3(0, 0, 0);
  float pA = 0, pB = 0, pC = 0, denom = 0;
  const float plane_threshold = 0.015f, min_distance_base = 0.01f;

  for (int stage = 0; stage < 4; ++stage) {
    if (stage >= 1) {
      const float4 p = a.spos4[s_b[stage - 1]], n = a.sattr[s_b[stage - 1]];
      pb[stage - 1] = v3(p.x, p.y, p.z);
      nb[stage - 1] = v3(n.x, n.y, n.z);
    }
    if (stage == 2) v_1 = normalized(sub(pb[1], pb[0]));
    if (stage == 3) {
      const double x1 = pb[0].x, y1 = pb[0].y, z1 = pb[0].z;
      const double x2 = pb[1].x, y2 = pb[1].y, z2 = pb[1].z;
      const double x3 = pb[2].x, y3 = pb[2].y, z3 = pb[2].z;
      denom = (float)(-x3 * y2 * z1 + x2 * y3 * z1 + x3 * y1 * z2 - x1 * y3 * z2 - x2 * y1 * z3 + x1 * y2 * z3);
      if (denom != 0) {
        pA = (float)((-y2 * z1 + y3 * z1 + y1 * z2 - y3 * z2 - y1 * z3 + y2 * z3) / denom);
        pB = (float)((x2 * z1 - x3 * z1 - x1 * z2 + x3 * z2 + x1 * z3 - x2 * z3) / denom);
        pC = (float)((-x2 * y1 + x3 * y1 + x1 * y2 - x3 * y2 - x1 * y3 + x2 * y3) / denom);
      }
    }
    // ---- update pass (stages 1..3): survivors of the previous stage that pass this stage's predicates.
    // The predicates cost three atan2 each, and only points within the model's diameter of the last
    // chosen point can pass (their PPF distance bin must exist in the model map): a cheap distance test
    // over all points builds a dense candidate list in shared memory, and the 1024 threads then evaluate
    // the predicates over that list -- with one warp tile per 32 consecutive points the few candidates
    // left most lanes idle (0.26 ms per 100 bases on the 13 419-point YCB frame).  The resulting bitmap
    // is the same set of points.
    if (stage >= 1) {
      const int bsel = s_b[stage - 1];
      if (tid < 32) s_wsum[tid] = 0ull;   // (ordered before the adds below by the first barrier of the chunk loop)
      for (int c0 = 0; c0 < a.S; c0 += kChunk) {
        const int cn = min(kChunk, a.S - c0);           // points of this chunk
        const int cw = (cn + 31) >> 5;                  // bitmap words of this chunk
        if (tid < cw) s_alive[tid] = 0u;
        if (tid == 0) s_n = 0;
        __syncthreads();
        for (int o = tid; o < cw * 32; o += 1024) {     // whole warps stay in the loop: ballot below
          const int i = c0 + o;
          bool cand = false;
          if (o < cn && ((alive[i >> 5] >> (i & 31)) & 1u) && i != bsel) {
            const float4 p4 = a.spos4[i];
            cand = ppf_distance_may_exist(a.ppf, pb[stage - 1], v3(p4.x, p4.y, p4.z));
          }
          const unsigned bal = __ballot_sync(0xffffffffu, cand);
          int at = 0;
          if (lane == 0 && bal) at = atomicAdd(&s_n, __popc(bal));
          at = __shfl_sync(0xffffffffu, at, 0);
          if (cand) s_list[at + __popc(bal & ((1u << lane) - 1u))] = (uint16_t)o;
        }
        __syncthreads();
        const int n_cand = s_n;
        for (int e = tid; e < n_cand; e += 1024) {
          const int o = (int)s_list[e];
          const int i = c0 + o;
          const float4 p4 = a.spos4[i], n4 = a.sattr[i];
          const V3 p = v3(p4.x, p4.y, p4.z), n = v3(n4.x, n4.y, n4.z);
          const Ppf4 f = ppf_compute_dev(pb[stage - 1], nb[stage - 1], p, n, a.ppf.tr, a.ppf.rot);
          bool zero = !ppf_key_exists(a.ppf, f);
          if (stage == 2) {
            const V3 v_2 = normalized(sub(p, pb[0]));
            zero = zero || internal_angle_below_30(dot(v_1, v_2));
          } else if (stage == 3) {
            float planar = 10000.0f;
            if (denom != 0) planar = (float)fabs((double)((pA * p.x + pB * p.y) + pC * p.z) - 1.0);
            zero = zero || (planar > plane_threshold) || (norm(sub(p, pb[0])) < min_distance_base) ||
                   (norm(sub(p, pb[1])) < min_distance_base) || (norm(sub(p, pb[2])) < min_distance_base);
          }
          if (!zero) {
            atomicOr(&s_alive[o >> 5], 1u << (o & 31));
            // the survivor's weight goes straight to the sum of the warp range that owns its tile
            // (integer adds: any order gives the sums the per-range pass of stage 0 would)
            atomicAdd(&s_wsum[(i >> 5) / tpw], prob_weight(n4.w));
          }
        }
        __syncthreads();
        if (tid < cw) alive[(c0 >> 5) + tid] = s_alive[tid];
        __syncthreads();
      }
    }
    // ---- stage 0: every point is alive; weights summed per warp range
    if (stage == 0) {
      unsigned long long lsum = 0;
      for (int t = t0; t < t1; ++t) {
        const int i = t * 32 + lane;
        const bool al = i < a.S;
        const unsigned word = __ballot_sync(0xffffffffu, al);
        if (lane == 0) alive[t] = word;
        if (al) lsum += prob_weight(a.sattr[i].w);
      }
      lsum = warp_sum_u64(lsum);
      if (lane == 0) s_wsum[w] = lsum;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long total = 0;
      for (int k = 0; k < 32; ++k) total += s_wsum[k];
      if (total == 0) {
        s_pw = -1;
      } else {
        unsigned long long r = mulhi_u64(draw_u64(a.seed, base_no, (uint32_t)stage), total);
        int k = 0;
        while (k < 31 && r >= s_wsum[k]) { r -= s_wsum[k]; ++k; }
        s_pw = k;
        s_rem = r;
      }
    }
    __syncthreads();
    if (s_pw < 0) {  // the reference's "FAILED SAMPLING:: Zero probability returned" => return false
      if (tid == 0) {
        a.out_valid[base] = 0;
        for (int k = 0; k < 4; ++k) a.out_ids[4 * base + k] = -1;
        a.out_inv[2 * base] = 0.f; a.out_inv[2 * base + 1] = 0.f;
      }
      return;
    }
    if (w == s_pw) {
      unsigned long long rem = s_rem;
      for (int t = t0; t < t1; ++t) {
        const int i = t * 32 + lane;
        const uint32_t word = alive[t];
        if (word == 0u) continue;          // (warp-uniform) nothing to walk over in this tile
        unsigned long long wt = 0;
        if (i < a.S && ((word >> lane) & 1u)) wt = prob_weight(a.sattr[i].w);
        unsigned long long inc = wt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const unsigned long long up = __shfl_up_sync(0xffffffffu, inc, o);
          if (lane >= o) inc += up;
        }
        const unsigned long long tile_total = __shfl_sync(0xffffffffu, inc, 31);
        if (rem < tile_total) {
          const unsigned hit = __ballot_sync(0xffffffffu, inc > rem);
          if (lane == 0) s_b[stage] = t * 32 + (__ffs(hit) - 1);
          break;
        }
        rem -= tile_total;
      }
    }
    __syncthreads();
  }
  // try_sampled_base: the 12 ordered segment pairings, one thread each (binary64 arithmetic); thread 0
  // then takes the first strict minimum in the reference's (i, j) loop order
  __shared__ float s_sd[12], s_i1[12], s_i2[12];
  __shared__ int s_perm[12][4];
  if (tid < 12) {
    const int i = tid / 3;
    int j = tid - 3 * i; if (j >= i) ++j;        // the three j != i, ascending
    int k = 0; while (k == i || k == j) k++;
    int l = 0; while (l == i || l == j || l == k) l++;
    V3 b[4];
    for (int q = 0; q < 4; ++q) { const float4 p = a.spos4[s_b[q]]; b[q] = v3(p.x, p.y, p.z); }
    double li1, li2;
    s_sd[tid] = (float)seg_dist_inv(b[i], b[j], b[k], b[l], li1, li2);
    s_i1[tid] = (float)li1; s_i2[tid] = (float)li2;
    s_perm[tid][0] = i; s_perm[tid][1] = j; s_perm[tid][2] = k; s_perm[tid][3] = l;
  }
  __syncthreads();
  if (tid == 0) {
    const int ids[4] = {s_b[0], s_b[1], s_b[2], s_b[3]};
    float min_distance = 3.402823466e+38f, inv1 = 0.f, inv2 = 0.f;
    int best = -1;
    for (int c = 0; c < 12; ++c)
      if (s_sd[c] < min_distance) { min_distance = s_sd[c]; best = c; inv1 = s_i1[c]; inv2 = s_i2[c]; }
    const bool ok = best >= 0;
    for (int k = 0; k < 4; ++k) a.out_ids[4 * base + k] = ok ? ids[s_perm[best][k]] : ids[k];
    a.out_inv[2 * base] = inv1; a.out_inv[2 * base + 1] = inv2;
    a.out_valid[base] = ok ? 1 : 0;
  }
}

}  // namespace

PpfView stocs_ppf_view(const stocs_b200_ctx* ctx);

// device outputs: d_ids (n*4 int), d_inv (n*2 float), d_valid (n bytes)
int stocs_launch_sample(stocs_b200_ctx* ctx, uint64_t seed, uint32_t first_base_no, int n_bases, int* d_ids,
                        float* d_inv, uint8_t* d_valid, cudaStream_t st) {
  SampleArgs a;
  a.spos4 = ctx->d_spos4.as<float4>();
  a.sattr = ctx->d_sattr.as<float4>();
  a.S = ctx->S;
  a.ppf = stocs_ppf_view(ctx);
  a.seed = seed;
  a.words = (ctx->S + 31) / 32;
  a.out_ids = d_ids; a.out_inv = d_inv; a.out_valid = d_valid;
  // survivors bitmap: batches of at most 4 * num_sms bases
  const int batch = ctx->num_sms * 4;
  STOCS_CUDA(ctx, ctx->d_work.ensure((size_t)batch * a.words * 4));
  a.alive = ctx->d_work.as<uint32_t>();
  for (int off = 0; off < n_bases; off += batch) {
    const int n = (n_bases - off < batch) ? (n_bases - off) : batch;
    SampleArgs b = a;
    b.first_base = first_base_no + (uint32_t)off;
    b.out_ids = d_ids + 4 * (size_t)off;
    b.out_inv = d_inv + 2 * (size_t)off;
    b.out_valid = d_valid + off;
    sample_bases_kernel<<<n, 1024, 0, st>>>(b);
  }
  STOCS_CUDA(ctx, cudaGetLastError());
  return STOCS_OK;
}

extern "C" int stocs_b200_sample_bases(stocs_b200_ctx* ctx, uint64_t seed, uint32_t first_base_no, int n_bases,
                                       int32_t* base_idx4, float* inv2, uint8_t* valid) {
  if (!ctx) return STOCS_E_ARG;
  if (ctx->S <= 0 || ctx->M <= 0) STOCS_FAIL(ctx, STOCS_E_STATE, "sample_bases: upload_model and upload_scene first");
  if (n_bases < 0 || (n_bases > 0 && (!base_idx4 || !inv2 || !valid))) STOCS_FAIL(ctx, STOCS_E_ARG, "sample_bases: bad argument");
  if (n_bases == 0) return STOCS_OK;
  cudaSetDevice(ctx->device);
  cudaStream_t st = ctx->stream;
  STOCS_CUDA(ctx, ctx->d_tmp2.ensure((size_t)n_bases * (16 + 8 + 1) + 64));
  int* d_ids = ctx->d_tmp2.as<int>();
  float* d_inv = (float*)(d_ids + 4 * (size_t)n_bases);
  uint8_t* d_valid = (uint8_t*)(d_inv + 2 * (size_t)n_bases);
  int rc = stocs_launch_sample(ctx, seed, first_base_no, n_bases, d_ids, d_inv, d_valid, st);
  if (rc) return rc;
  STOCS_CUDA(ctx, cudaMemcpyAsync(base_idx4, d_ids, (size_t)n_bases * 16, cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaMemcpyAsync(inv2, d_inv, (size_t)n_bases * 8, cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaMemcpyAsync(valid, d_valid, (size_t)n_bases, cudaMemcpyDeviceToHost, st));
  STOCS_CUDA(ctx, cudaStreamSynchronize(st));
  return STOCS_OK;
}
