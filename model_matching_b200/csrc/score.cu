// score.cu -- LCP scoring of rigid-transform hypotheses (the dominant kernel).
//
// Replaces stocs_estimator::compute_alignment_score_for_rigid_transform (reference
// src/stocs.cpp:1006-1041) and the kd-tree query it calls
// (include/super4pcs/accelerators/kdtree.h:394-459).
//
// One warp per hypothesis, persistent CTAs, dynamic work counter.  Model points live in shared
// memory as float4.
//   Phase A (per block of 256 model points), two loops:
//     1. the test every point takes: fused affine map "transform o world->grid" (9 FMAs), floor,
//        block coordinates clamped to the always-empty border block (no bounds test, no branch),
//        one bit of a multi-resolution occupancy bitmap held in shared memory; survivors (~35 %)
//        are compacted, in model-point order, into a per-warp byte list;
//     2. survivors only, 32 at a time: one bit of the per-brick bitmap, then -- occupied bricks
//        only -- ONE 16-byte brick record of the eps-dilated voxel grid (64-bit occupancy mask +
//        rank base); lanes whose cell is occupied are compacted (ballot + popc rank) into a
//        per-warp shared-memory queue that runs ACROSS rounds, in model-point order.
//   Phase B (queue more than half full, or end of the model), 32 queued queries at a time, one
//     per "owner" lane: (1) owners fetch their candidate-list offsets and compute the exact
//     transformed point; (2) the 32 candidate lists are swept as ONE flat list, 32 consecutive
//     float4 records per trip (coalesced, no padding to the longest list); owners are found with a
//     redux.or + popc, their points arrive by shuffle, hits race with shared-memory atomicMin on the
//     bit pattern of d^2 and then on the scene index; an exact d^2 tie falls through to a walk of
//     the reference kd-tree so that it is broken the way kdtree.h:416-428 breaks it; (3) owners
//     fetch the matched scene point's normal + class probability, apply the 30-degree test, and
//     the class probabilities are accumulated in lane order == model-point order, which makes the
//     LCP bit-identical to the reference's sequential fp32 sum.
// All memory-dependent steps are batched 32 wide, so the latency chain
// brick -> offsets -> candidates -> attributes is paid once per drain, not once per model point.
#include "stocs_ctx.h"

using namespace stocsm;

namespace {

// queued queries per warp: a larger queue means fewer, fuller drains (S1: 64 -> 2.15 ms, 96 -> 2.09 ms,
// 128 -> 2.09 ms; shared memory per CTA grows by 8 KB per 32 entries)
#ifndef SCORE_QUEUE
#define SCORE_QUEUE 96
#endif
// shared-memory minimum over the candidates of a query: native 32-bit atomicMin on the d^2 bit
// pattern + index write-back (default; S1 2.28 -> 2.15 ms, S1-fit 4.06 -> 3.76 ms) or the original
// 64-bit (d^2, index) atomicMin, which compiles to a CAS loop (-DSCORE_ATOM64)
#ifndef SCORE_ATOM64
#define SCORE_ATOM32
#endif
#ifndef SCORE_MIN_BLOCKS
#define SCORE_MIN_BLOCKS 2
#endif
#ifndef SCORE_WARPS
#define SCORE_WARPS 32
#endif
constexpr int kWarps = SCORE_WARPS; // warps per CTA
constexpr int kQueue = SCORE_QUEUE; // queued queries per warp

struct ScoreArgs {
  const uint4* __restrict__ bricks;
  const uint32_t* __restrict__ coarse;   // 1 bit per block of 2^coarse_shift cells per axis: any occupied cell inside
  const uint32_t* __restrict__ brick_occ;  // 1 bit per brick: any occupied cell inside
  int coarse_words;                      // words of `coarse` (always staged in shared memory, <= 16 KB)
  int coarse_shift, coarse_nx, coarse_ny, coarse_nz;  // real blocks per axis; index n* is the empty border block
  int coarse_sx, coarse_sy;                           // row / slab strides of the map = blocks + 1
  const uint32_t* __restrict__ starts;
  const float4* __restrict__ cand;
  const float4* __restrict__ sattr;
  const float* __restrict__ model;   // 8*Mpad floats: float4 positions, then float4 normals
  const KdNodeDev* __restrict__ kd_nodes;
  const float4* __restrict__ kd_pts;
  const float* __restrict__ T;
  float* __restrict__ lcp;
  int* __restrict__ inl;
  unsigned long long* work_counter;
  unsigned long long* tie_counter;
  unsigned long long* cnt;   // counting variant only (stocs_b200_score_counters): see ScoreCounter
  const int* __restrict__ order;   // claim number -> hypothesis (heavy-first schedule, see probe_order_kernel) or NULL
  long long H;
  const long long* __restrict__ H_dev;   // optional: the count lives on the device (H is then an upper bound)
  GridDesc g;
  int M, Mpad;
  float sq_eps, dot_thr;
  float to_block, to_cell;   // 2^-coarse_shift and 2^coarse_shift (exact scalings, see loop 1)
};

// kdtree.h:394-459 on the device (tie path only).
__device__ __noinline__ int kd_query_dev(const KdNodeDev* __restrict__ nodes, const float4* __restrict__ pts,
                                         float qx, float qy, float qz, float sqdist) {
  uint32_t st_node[64];
  float st_sq[64];
  int cl_id = -1;
  float cl_dist = sqdist;
  st_node[0] = 0; st_sq[0] = 0.f;
  unsigned count = 1;
  while (count) {
    uint32_t nid = st_node[count - 1];
    float sq = st_sq[count - 1];
    KdNodeDev nd = nodes[nid];
    if (sq < cl_dist) {
      if (nd.leaf) {
        --count;
        uint32_t end = nd.first_or_start + nd.dim_or_size;
        for (uint32_t i = nd.first_or_start; i < end; ++i) {
          float4 p = pts[i];
          float dx = qx - p.x, dy = qy - p.y, dz = qz - p.z;
          float d = dx * dx + (dy * dy + dz * dz);
          if (d <= cl_dist) { cl_dist = d; cl_id = __float_as_int(p.w); }
        }
      } else {
        float qd = nd.dim_or_size == 0 ? qx : (nd.dim_or_size == 1 ? qy : qz);
        float new_off = qd - nd.split;
        if (new_off < 0.f) {
          st_node[count] = nd.first_or_start;
          st_node[count - 1] = nd.first_or_start + 1;
        } else {
          st_node[count] = nd.first_or_start + 1;
          st_node[count - 1] = nd.first_or_start;
        }
        st_sq[count] = sq;
        st_sq[count - 1] = new_off * new_off;
        ++count;
      }
    } else {
      --count;
    }
  }
  return cl_id;
}

// Streaming 16-byte load that does not allocate in L1 (candidate and attribute records are read
// once per query; keeping them out of L1 leaves it to the brick table and the spill slots).
__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// L2 policies (default on; -DSCORE_NO_L2_HINTS / -DSCORE_NO_BRICK_OCC are the A/B switches).  With
// 0.5-eps cells the launch moves 12 GB through DRAM at ~4.4 TB/s and waits on memory (long
// scoreboard is the top stall), so L2 residency matters: brick records (70 MB at S1, re-used by
// every hypothesis) are loaded evict_last, candidate records (1.4 GB, touched once per query)
// evict_first: 2.58 -> 2.50 ms; the one-bit-per-brick filter in front of the brick record:
// 2.50 -> 2.48 ms.  (sm_100a accepts .L2::evict_* directly only on 256-bit loads, hence
// createpolicy + .L2::cache_hint.)
#ifndef SCORE_NO_L2_HINTS
#define SCORE_L2_HINTS
#endif
#ifndef SCORE_NO_BRICK_OCC
#define SCORE_BRICK_OCC
#endif
#ifdef SCORE_L2_HINTS
// the policy words are pure values (plain asm, so the compiler may hoist or rematerialise them)
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
  unsigned long long p;
  asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint4 ld_brick_keep(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p), "l"(l2_policy_evict_last()));
  return v;
}
__device__ __forceinline__ float4 ld_cand_first(const float4* p) {
  float4 v;
  asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(l2_policy_evict_first()));
  return v;
}
#define LD_BRICK ld_brick_keep
#else
#define LD_BRICK __ldg
#endif

#if defined(SCORE_L2_HINTS)
#define LD_CAND ld_cand_first
#elif defined(SCORE_STREAM_CAND)
#define LD_CAND ld_stream
#else
#define LD_CAND __ldg
#endif

// %laneid / %lanemask_* as plain (non-volatile) asm: one instruction when the compiler has to
// rematerialise them inside the loops (registers are capped at 32 for 64 warps/SM)
__device__ __forceinline__ unsigned lane_id() { unsigned r; asm("mov.u32 %0, %%laneid;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned lanemask_lt() { unsigned r; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(r)); return r; }
__device__ __forceinline__ unsigned lanemask_le() { unsigned r; asm("mov.u32 %0, %%lanemask_le;" : "=r"(r)); return r; }

struct WarpQueue {   // one per warp, shared memory (single base register, constant offsets)
  // exact transform and the same map into grid-block coordinates (FMA-evaluated, phase A only), as
  // three rows {m_r0, m_r1, m_r2, t_r}: a row is ONE 16-byte broadcast load (12 scalar loads per use
  // were a third of the kernel's shared-memory load wavefronts)
  float4 T[3];
  float4 G[3];
  uint32_t a0[kQueue];   // occupied-cell rank of the queued query
  uint32_t pi[kQueue];   // model point index of the queued query
#ifdef SCORE_ATOM32
  uint32_t best_d[32];   // drain: min over hits of the d^2 bit pattern (native 32-bit shared atomicMin)
  uint32_t best_i[32];   //        scene index of a candidate that attains it
#else
  unsigned long long best[32];  // drain: min over hits of (d^2 bit pattern << 32 | scene index), atomicMin
#endif
  uint8_t plist[256];    // phase A: survivors of the coarse test within the current 256-point block
};

struct Acc { float acc; int inl; unsigned ties; };

// Model positions in shared memory as three float arrays x[Mpad] y[Mpad] z[Mpad]: a warp load is three
// LDS.32 = 3 wavefronts of the L1 data pipe -- the kernel's binding resource -- where a float4 record
// (LDS.128) takes 4, and a gather of 32 arbitrary points conflicts less (S1 2.074 -> 2.058 ms,
// S1-fit 3.582 -> 3.481 ms; -DSCORE_AOS restores the float4 records).
#ifndef SCORE_AOS
struct ModelPts {
  const float* x; int Mpad;
  __device__ __forceinline__ float4 operator[](int i) const { return make_float4(x[i], x[Mpad + i], x[2 * Mpad + i], 0.f); }
};
#else
typedef const float4* ModelPts;
#endif
#ifndef SCORE_AOS
constexpr int kModelFloats = 3;   // floats of shared memory per (padded) model point
#else
constexpr int kModelFloats = 4;
#endif

// Work counters of the counting variant (template parameter kCount; the timed kernel carries none
// of this).  One global atomic per warp-level event, issued by lane 0.
enum ScoreCounter { SC_SURVIVORS = 0, SC_BRICK_RECORDS, SC_QUEUED, SC_CANDIDATES, SC_HITS, SC_INLIERS, SC_DRAINS, SC_N };
template <bool kCount>
__device__ __forceinline__ void count(const ScoreArgs& a, int lane, int slot, unsigned v) {
  if (kCount) { if (lane == 0 && v) atomicAdd(a.cnt + slot, (unsigned long long)v); }
}

// Phase B.  The queued queries are processed 32 at a time, one per lane ("owner" lane e holds
// query e: exact transformed point, candidate offset, candidate count).
//   1. owners fetch their candidate-list offsets and compute the EXACT transformed point
//      (mat * p.homogeneous()).head<3>() in the reference's evaluation order (stocs_math.h
//      xform_point) -- phase A only located the cell;
//   2. the candidate lists of the 32 queries are swept as ONE flat list, 32 consecutive records
//      per trip (coalesced, no padding to the longest list).  The owner of flat position f is found
//      without any table: a redux.or of "my list starts at window offset j" bits + popc gives the
//      rank, the owner's point arrives by shuffle.  Candidates within eps race with a 64-bit
//      shared-memory atomicMin on (d^2 bit pattern, scene index); a hit that meets the SAME d^2
//      under a different index flags the query, and a flagged query is answered by a walk of the
//      reference kd-tree, which reproduces the reference's tie rule (kdtree.h:416-428);
//   3. owners fetch the matched scene point's normal + class probability, apply the 30-degree test
//      and the class probabilities are accumulated in lane order == model-point order, which makes
//      the LCP bit-identical to the reference's sequential fp32 sum.
template <bool kCount>
__device__ __forceinline__ void drain_queue(const ScoreArgs& a, WarpQueue& q, int qn, int lane,
                                            const ModelPts mp4, const float4* __restrict__ mn4, Acc& r) {
  __syncwarp();
  const float sq_eps = a.sq_eps;
  const unsigned le_mask = lanemask_le();
  for (int half = 0; half < qn; half += 32) {
    const int e = half + lane;
    const bool has = e < qn;
    // 1.
    uint32_t s = 0, cnt = 0, mi = 0;
    float qx = 0.f, qy = 0.f, qz = 0.f;
    if (has) {
      const uint32_t k = q.a0[e];
      s = __ldg(a.starts + k);
      cnt = __ldg(a.starts + k + 1) - s;
      mi = q.pi[e];
      const float4 mp = mp4[mi];
      const float4 t0 = q.T[0], t1 = q.T[1], t2 = q.T[2];
      qx = ((t0.x * mp.x + t0.y * mp.y) + t0.z * mp.z) + t0.w;
      qy = ((t1.x * mp.x + t1.y * mp.y) + t1.z * mp.z) + t1.w;
      qz = ((t2.x * mp.x + t2.y * mp.y) + t2.z * mp.z) + t2.w;
    }
#ifdef SCORE_ATOM32
    q.best_d[lane] = 0xffffffffu;
#else
    q.best[lane] = ~0ull;
#endif
    uint32_t pre = cnt;         // inclusive scan of the counts -> exclusive prefix
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t up = __shfl_up_sync(0xffffffffu, pre, o);
      if (lane >= o) pre += up;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, pre, 31);
    pre -= cnt;
    count<kCount>(a, lane, SC_CANDIDATES, total);
    count<kCount>(a, lane, SC_DRAINS, 1u);
    __syncwarp();
    // 2.
    unsigned flag = 0;          // queries this lane wants answered by the kd-tree (exact d^2 ties)
    for (uint32_t t0 = 0; t0 < total; t0 += 32) {
      const uint32_t f = t0 + lane;
      const int rel = (int)pre - (int)t0;
      const unsigned starts_here = __reduce_or_sync(0xffffffffu, (cnt > 0 && rel >= 0 && rel < 32) ? (1u << rel) : 0u);
      const int before = __popc(__ballot_sync(0xffffffffu, cnt > 0 && rel < 0));
      // owner = the last query whose list starts at or before f.  Lists are non-empty (every queued
      // cell is occupied), so ranks among non-empty lists == lane numbers among `has` lanes.
      const int own = before - 1 + __popc(starts_here & le_mask);
      const int src = own < 0 ? 0 : own;
      const float ox = __shfl_sync(0xffffffffu, qx, src), oy = __shfl_sync(0xffffffffu, qy, src),
                  oz = __shfl_sync(0xffffffffu, qz, src);
      const uint32_t os = __shfl_sync(0xffffffffu, s, src), op = __shfl_sync(0xffffffffu, pre, src);
#ifdef SCORE_ATOM32
      bool hit = false;
      uint32_t hb = 0, hi = 0;
#endif
      if (f < total) {
        const float4 c = LD_CAND(a.cand + (uint32_t)(os + (f - op)));
        const float dx = ox - c.x, dy = oy - c.y, dz = oz - c.z;
        const float d = dx * dx + (dy * dy + dz * dz);
#ifdef SCORE_ATOM32
        // (d^2 >= 0, so its bit pattern orders like the value.)  The minimum is taken with the
        // native 32-bit shared atomic; a candidate that meets its own d^2 already stored is an exact
        // tie (flag); after the trip's atomics, whoever equals the minimum records its index.
        hit = d <= sq_eps;
        hb = __float_as_uint(d); hi = __float_as_uint(c.w);
        if (hit && atomicMin(&q.best_d[src], hb) == hb) flag |= 1u << src;
      }
      __syncwarp();
      if (hit && q.best_d[src] == hb) q.best_i[src] = hi;
      __syncwarp();
    }
#else
        if (d <= sq_eps) {
          const uint32_t b = __float_as_uint(d), ci = __float_as_uint(c.w);
          const unsigned long long old = atomicMin(&q.best[src], ((unsigned long long)b << 32) | ci);
          if ((uint32_t)(old >> 32) == b && (uint32_t)old != ci) flag |= 1u << src;   // same d^2, another point
        }
      }
    }
#endif
    __syncwarp();
    flag = __reduce_or_sync(0xffffffffu, flag);
    // 3.
    bool match = false;
    float w = 0.f;
#ifdef SCORE_ATOM32
    const bool found = q.best_d[lane] != 0xffffffffu;
    if (has && found) {
      int res = (int)q.best_i[lane];
#else
    const unsigned long long mine = q.best[lane];
    const bool found = mine != ~0ull;
    if (has && found) {
      int res = (int)(uint32_t)mine;
#endif
      if ((flag >> lane) & 1u) {
        res = kd_query_dev(a.kd_nodes, a.kd_pts, qx, qy, qz, sq_eps);
        r.ties++;
      }
      if (res >= 0) {
        const float4 sa = ld_stream(a.sattr + res);
        const float4 mn = __ldg(mn4 + mi);
        // mat.block<3,3>(0,0) * n  -- see stocs_math.h xform_dir
        const float4 t0 = q.T[0], t1 = q.T[1], t2 = q.T[2];
        const float rx = t0.x * mn.x + (t0.y * mn.y + t0.z * mn.z);
        const float ry = t1.x * mn.x + (t1.y * mn.y + t1.z * mn.z);
        const float rz = t2.x * mn.x + (t2.y * mn.y + t2.z * mn.z);
        const float dt = sa.x * rx + (sa.y * ry + sa.z * rz);
        // acos(dt)*180/pi < 30  <=>  dot_thr <= dt <= 1   (threshold found by bisection on the host)
        match = (dt >= a.dot_thr) && (dt <= 1.0f);
        w = sa.w;
      }
    }
    unsigned mm = __ballot_sync(0xffffffffu, match);
    if (kCount) count<kCount>(a, lane, SC_HITS, __popc(__ballot_sync(0xffffffffu, has && found)));
    count<kCount>(a, lane, SC_INLIERS, __popc(mm));
    r.inl += __popc(mm);
    while (mm) {  // ordered fp32 accumulation == the reference's sequential loop
      const int b = __ffs(mm) - 1;
      r.acc += __shfl_sync(0xffffffffu, w, b);
      mm &= mm - 1;
    }
    __syncwarp();
  }
}

// next unit of work (lane 0 only): a claim number from the global counter, mapped through the
// heavy-first order when there is one; any value >= H ends the warp
__device__ __forceinline__ int claim_hypothesis(const ScoreArgs& a) {
  const unsigned long long c = atomicAdd(a.work_counter, 1ull);
  // (the online pipeline keeps its transform count on the device: H is then only an upper bound)
  const unsigned long long lim = a.H_dev ? (unsigned long long)__ldg(a.H_dev) : (unsigned long long)a.H;
  if (c >= lim) return 0x7fffffff;
  return a.order ? __ldg(a.order + c) : (int)c;
}

template <bool kCount>
__global__ void __launch_bounds__(kWarps * 32, SCORE_MIN_BLOCKS) score_lcp_kernel(ScoreArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_model = reinterpret_cast<float*>(smem_raw);
  const int Mpad = a.Mpad;
  const int lane = (int)lane_id();
  // warp index through redux: the result lives in a UNIFORM register, and so does everything
  // derived from it (the warp's queue address), outside the 32-register budget -- as a plain
  // tid >> 5 the queue address was spilled and re-loaded at 18 sites (37 local loads per hypothesis)
  const int warp = (int)__reduce_max_sync(0xffffffffu, threadIdx.x >> 5);
  // dynamic shared memory: [model positions float4 x Mpad][coarse bitmap][one WarpQueue per warp]
#ifndef SCORE_AOS
  for (int i = threadIdx.x; i < 4 * Mpad; i += blockDim.x) { const int c = i & 3; if (c < 3) s_model[c * Mpad + (i >> 2)] = a.model[i]; }
#else
  for (int i = threadIdx.x; i < 4 * Mpad; i += blockDim.x) s_model[i] = a.model[i];
#endif
  uint32_t* s_coarse = reinterpret_cast<uint32_t*>(s_model + kModelFloats * Mpad);   // Mpad % 64 == 0: 16-byte aligned
  for (int i = threadIdx.x; i < a.coarse_words; i += blockDim.x) s_coarse[i] = a.coarse[i];
  WarpQueue& q = reinterpret_cast<WarpQueue*>(s_coarse + ((a.coarse_words + 3) & ~3))[warp];
  __syncthreads();
#ifndef SCORE_AOS
  ModelPts mp4; mp4.x = s_model; mp4.Mpad = Mpad;                // positions as x[] y[] z[] (NaN beyond M)
#else
  const float4* mp4 = reinterpret_cast<const float4*>(s_model);  // positions as float4 (NaN beyond M)
#endif
  const float4* mn4 = reinterpret_cast<const float4*>(a.model) + Mpad;  // normals stay in global (hits only)

  const unsigned lt_mask = lanemask_lt();
  const int M = a.M;
  // shared-window address of the coarse bitmap, made warp-uniform through redux so that it lives in a
  // uniform register for the whole kernel
  const unsigned coarse_sa = __reduce_max_sync(0xffffffffu, (unsigned)__cvta_generic_to_shared(s_coarse));
  Acc r;
  r.ties = 0;

  // Work distribution: lane 0 claims the next hypothesis from the global counter when the warp has
  // finished one.  (Claiming ahead does not pay: the result cannot stay in a register across a
  // hypothesis at the 32-register cap, so the warp waits for the atomic either way -- 10 % of the
  // stall samples when it was a prefetch; claiming 4 or 8 at a time lengthens the tail by more
  // than it saves: 2.63 / 2.76 ms against 2.58 ms.)
  // (claimed index broadcast by redux for the same reason: h stays in a uniform register)
  int h = 0;
  if (lane == 0) h = claim_hypothesis(a);
  h = (int)__reduce_max_sync(0xffffffffu, (unsigned)h);
  while (h < a.H) {
    // lanes 0..11 fetch the 3x4 transform (column-major 4x4: element (r,c) at c*4+r) and publish
    // it, with its grid-coordinate version G = inv_cell * (T - origin), to the warp's shared slot
    if (lane < 12) {
      const int c = lane / 3, rr = lane - 3 * c;
      const float t = __ldg(a.T + 16 * (size_t)h + c * 4 + rr);
      const float o = (rr == 0) ? a.g.ox : (rr == 1 ? a.g.oy : a.g.oz);
      reinterpret_cast<float*>(q.T)[rr * 4 + c] = t;
      // G maps into BLOCK coordinates (cell coordinates * 2^-coarse_shift): loop 1 floors it straight
      // to the coarse-map index; loop 2 multiplies by 2^coarse_shift first.  Scaling by a power of two
      // commutes with every rounding of the FMA chain, so the cell found is bit-identical to mapping
      // with the unscaled G.
      reinterpret_cast<float*>(q.G)[rr * 4 + c] = ((c == 3) ? (t - o) * a.g.inv_cell : t * a.g.inv_cell) * a.to_block;
    }
    __syncwarp();
    // gK = element (row K % 3, column K / 3) of the grid map
    float4 gr0 = q.G[0], gr1 = q.G[1], gr2 = q.G[2];
    float g0 = gr0.x, g3 = gr0.y, g6 = gr0.z, g9 = gr0.w, g1 = gr1.x, g4 = gr1.y, g7 = gr1.z, g10 = gr1.w;
    float g2 = gr2.x, g5 = gr2.y, g8 = gr2.z, g11 = gr2.w;
    r.acc = 0.f;
    r.inl = 0;
    int qn = 0;
    // Phase A runs over blocks of 256 model points in two loops.  Loop 1 is the cheap test every
    // point takes: fused affine map -> cell -> one bit of the shared-memory coarse occupancy map;
    // the survivors' indices are compacted, in model-point order, into a byte list.  Loop 2 visits
    // only the survivors, 32 at a time: brick record, exact cell bit, enqueue.  (Explicit FMAs: the
    // map only selects a cell, the eps-dilation margin of the index absorbs its rounding; the exact
    // point is recomputed for queued queries.  Float->int floor saturates; NaN -- padding points,
    // rejected fits -- gives cell 0, whose queries find no candidate because every distance is NaN.)
    for (int blk = 0; blk < M; blk += 256) {
      int nl = 0;
      const int rounds = min(8, (M - blk + 31) >> 5);
      for (int rr = 0; rr < rounds; ++rr) {
        const float4 mp = mp4[blk + rr * 32 + lane];
        const float fx = __fmaf_rn(g0, mp.x, __fmaf_rn(g3, mp.y, __fmaf_rn(g6, mp.z, g9)));
        const float fy = __fmaf_rn(g1, mp.x, __fmaf_rn(g4, mp.y, __fmaf_rn(g7, mp.z, g10)));
        const float fz = __fmaf_rn(g2, mp.x, __fmaf_rn(g5, mp.y, __fmaf_rn(g8, mp.z, g11)));
        const unsigned ix = (unsigned)__float2int_rd(fx), iy = (unsigned)__float2int_rd(fy), iz = (unsigned)__float2int_rd(fz);
        // block coordinates clamped to the always-empty border block (index = number of real
        // blocks): no bounds test, no branch; blocks left of the grid are huge as unsigned
        const unsigned cx = min(ix, (unsigned)a.coarse_nx), cy = min(iy, (unsigned)a.coarse_ny),
                       cz = min(iz, (unsigned)a.coarse_nz);
        const uint32_t cidx = (cz * (unsigned)a.coarse_sy + cy) * (unsigned)a.coarse_sx + cx;
        // The bitmap word is loaded through its shared-window address held in a uniform register, and
        // the ballot is written in PTX on ONE predicate: as plain C++ the compiler re-derived the
        // address (S2UR + UMOV + 2 ULEA) and a second predicate (ISETP) in every iteration -- 40 -> 35
        // SASS instructions per 32 points (profiles/r02_score_loop1_sass.txt).
        unsigned cw;
        asm("ld.shared.u32 %0, [%1];" : "=r"(cw) : "r"(coarse_sa + ((cidx >> 5) << 2)));
        const unsigned bitv = (cw >> (cidx & 31)) & 1u;
        unsigned pm;
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\tvote.sync.ballot.b32 %0, p, 0xffffffff;\n\t}" : "=r"(pm) : "r"(bitv));
        const bool pass = bitv != 0u;
        if (pass) q.plist[nl + __popc(pm & lt_mask)] = (uint8_t)(rr * 32 + lane);
        nl += __popc(pm);
      }
      count<kCount>(a, lane, SC_SURVIVORS, (unsigned)nl);
      __syncwarp();
      for (int e0 = 0; e0 < nl; e0 += 32) {
        const bool valid = e0 + lane < nl;
        const int i = blk + (valid ? (int)q.plist[e0 + lane] : 0);
        const float4 mp = mp4[i];
        const float fx = __fmaf_rn(g0, mp.x, __fmaf_rn(g3, mp.y, __fmaf_rn(g6, mp.z, g9)));
        const float fy = __fmaf_rn(g1, mp.x, __fmaf_rn(g4, mp.y, __fmaf_rn(g7, mp.z, g10)));
        const float fz = __fmaf_rn(g2, mp.x, __fmaf_rn(g5, mp.y, __fmaf_rn(g8, mp.z, g11)));
        const unsigned ix = (unsigned)__float2int_rd(fx * a.to_cell), iy = (unsigned)__float2int_rd(fy * a.to_cell),
                       iz = (unsigned)__float2int_rd(fz * a.to_cell);
        uint4 br = make_uint4(0u, 0u, 0u, 0u);
        const unsigned bidx = ((iz >> 2) * (unsigned)a.g.nby + (iy >> 2)) * (unsigned)a.g.nbx + (ix >> 2);
#ifdef SCORE_BRICK_OCC
        // second-level filter: 1 bit per brick (L2-resident, 1/128 of the brick table) before the
        // 16 B brick record, which for an empty brick would be a wasted DRAM access
        if (valid && ((__ldg(a.brick_occ + (bidx >> 5)) >> (bidx & 31u)) & 1u)) br = LD_BRICK(a.bricks + bidx);
#else
        if (valid) br = LD_BRICK(a.bricks + bidx);
#endif
        if (kCount) {
#ifdef SCORE_BRICK_OCC
          const bool fetched = valid && ((__ldg(a.brick_occ + (bidx >> 5)) >> (bidx & 31u)) & 1u);
#else
          const bool fetched = valid;
#endif
          count<kCount>(a, lane, SC_BRICK_RECORDS, __popc(__ballot_sync(0xffffffffu, fetched)));
        }
        const unsigned bit = ((iz & 3u) << 4) | ((iy & 3u) << 2) | (ix & 3u);
        const unsigned half = (bit & 32u) ? br.y : br.x;      // 64-bit occupancy mask as two words
        const bool has = (half >> (bit & 31u)) & 1u;
        const unsigned hm = __ballot_sync(0xffffffffu, has);
        if (hm) {
          if (has) {
            const int slot = qn + __popc(hm & lt_mask);
            const unsigned below = __popc(half & ((1u << (bit & 31u)) - 1u)) + ((bit & 32u) ? __popc(br.x) : 0u);
            q.a0[slot] = br.z + below;
            q.pi[slot] = (uint32_t)i;
          }
          qn += __popc(hm);
          count<kCount>(a, lane, SC_QUEUED, __popc(hm));
        }
        if (qn > kQueue - 32) {
          // drain whole groups of 32 only; the (in-order) remainder stays at the front of the queue,
          // so a hypothesis pays for a partly filled group once, at its end (S1 2.098 -> 2.082 ms,
          // S1-fit 3.660 -> 3.611 ms)
          const int full = qn & ~31;
          drain_queue<kCount>(a, q, full, lane, mp4, mn4, r);
          const int rem = qn - full;
          uint32_t ca = 0, cp = 0;
          if (lane < rem) { ca = q.a0[full + lane]; cp = q.pi[full + lane]; }
          __syncwarp();
          if (lane < rem) { q.a0[lane] = ca; q.pi[lane] = cp; }
          __syncwarp();
          qn = rem;
          // the map is re-read after a drain so that it is not live (in registers) across it
          gr0 = q.G[0]; gr1 = q.G[1]; gr2 = q.G[2];
          g0 = gr0.x; g3 = gr0.y; g6 = gr0.z; g9 = gr0.w; g1 = gr1.x; g4 = gr1.y; g7 = gr1.z; g10 = gr1.w;
          g2 = gr2.x; g5 = gr2.y; g8 = gr2.z; g11 = gr2.w;
        }
      }
      __syncwarp();
    }
    if (qn > 0) drain_queue<kCount>(a, q, qn, lane, mp4, mn4, r);
    if (lane == 0) {
      a.lcp[h] = r.acc / (float)M;
      if (a.inl) a.inl[h] = r.inl;
    }
    h = 0;
    if (lane == 0) h = claim_hypothesis(a);
    h = (int)__reduce_max_sync(0xffffffffu, (unsigned)h);
    __syncwarp();
  }
  if (r.ties) atomicAdd(a.tie_counter, (unsigned long long)r.ties);
}

// Heavy-first schedule.  The cost of a hypothesis varies by an order of magnitude (a pose that lands
// the model on scene surface queues hundreds of queries, a pose in free space none), and with one
// warp per hypothesis the launch ends when the last HEAVY hypothesis claimed does: a straggler tail
// of ~0.09 ms, a quarter of the launch when 10^6 hypotheses are sharded over 8 GPUs.  This kernel
// estimates the cost of every hypothesis from 32 sample points of the model (coarse-map test only:
// 1/16 of loop 1, about 1 % of the scoring work) and writes a claim order with the predicted-heavy
// hypotheses first and the light ones last.  The order only changes WHO scores a hypothesis WHEN;
// every result is computed exactly as before.
__global__ void __launch_bounds__(256) probe_order_kernel(ScoreArgs a, int* __restrict__ order, unsigned* __restrict__ fill,
                                                          int heavy_threshold) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4* s_pts = reinterpret_cast<float4*>(smem_raw);                 // 32 sample points
  uint32_t* s_coarse = reinterpret_cast<uint32_t*>(s_pts + 32);
  if (threadIdx.x < 32) s_pts[threadIdx.x] = reinterpret_cast<const float4*>(a.model)[(int)(((long long)threadIdx.x * a.M) / 32)];
  for (int i = threadIdx.x; i < a.coarse_words; i += blockDim.x) s_coarse[i] = a.coarse[i];
  __syncthreads();
  const long long h = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned lane = lane_id();
  bool heavy = false;
  if (h < a.H) {
    float g[12];
    const float* t = a.T + 16 * (size_t)h;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int rr = 0; rr < 3; ++rr) {
        const float v = __ldg(t + c * 4 + rr);
        const float o = (rr == 0) ? a.g.ox : (rr == 1 ? a.g.oy : a.g.oz);
        g[rr * 4 + c] = ((c == 3) ? (v - o) * a.g.inv_cell : v * a.g.inv_cell) * a.to_block;
      }
    int cnt = 0;
#pragma unroll 4
    for (int s = 0; s < 32; ++s) {
      const float4 mp = s_pts[s];
      const float fx = __fmaf_rn(g[0], mp.x, __fmaf_rn(g[1], mp.y, __fmaf_rn(g[2], mp.z, g[3])));
      const float fy = __fmaf_rn(g[4], mp.x, __fmaf_rn(g[5], mp.y, __fmaf_rn(g[6], mp.z, g[7])));
      const float fz = __fmaf_rn(g[8], mp.x, __fmaf_rn(g[9], mp.y, __fmaf_rn(g[10], mp.z, g[11])));
      const unsigned cx = min((unsigned)__float2int_rd(fx), (unsigned)a.coarse_nx), cy = min((unsigned)__float2int_rd(fy), (unsigned)a.coarse_ny),
                     cz = min((unsigned)__float2int_rd(fz), (unsigned)a.coarse_nz);
      const uint32_t cidx = (cz * (unsigned)a.coarse_sy + cy) * (unsigned)a.coarse_sx + cx;
      cnt += (s_coarse[cidx >> 5] >> (cidx & 31)) & 1u;
    }
    heavy = cnt >= heavy_threshold;
  }
  // warp-aggregated append: heavy hypotheses fill the order from the front, light ones from the back
  const unsigned valid = __ballot_sync(0xffffffffu, h < a.H);
  const unsigned hm = __ballot_sync(0xffffffffu, heavy);
  const unsigned lm = valid & ~hm;
  unsigned base_h = 0, base_l = 0;
  if (lane == 0) {
    if (hm) base_h = atomicAdd(&fill[0], (unsigned)__popc(hm));
    if (lm) base_l = atomicAdd(&fill[1], (unsigned)__popc(lm));
  }
  base_h = __shfl_sync(0xffffffffu, base_h, 0);
  base_l = __shfl_sync(0xffffffffu, base_l, 0);
  if (h < a.H) {
    const unsigned below = lanemask_lt();
    if (heavy) order[base_h + __popc(hm & below)] = (int)h;
    else order[(unsigned)a.H - 1u - (base_l + __popc(lm & below))] = (int)h;
  }
}

__global__ void fmad_selftest_kernel(float a, float b, float c, float* out) { out[0] = a * b + c; }

}  // namespace

bool stocs_fmad_selftest(stocs_b200_ctx* ctx) {
  // a*b = 1 + 2^-11 + 2^-24 rounds to 1 + 2^-11 (ties-to-even); unfused a*b+c == 0, fused == 2^-24.
  if (ctx->d_small.ensure(256) != cudaSuccess) return false;
  float* d = ctx->d_small.as<float>() + 32;
  const float a = 1.0f + 1.0f / 4096.0f, c = -(1.0f + 1.0f / 2048.0f);
  fmad_selftest_kernel<<<1, 1, 0, ctx->stream>>>(a, a, c, d);
  float hres = 1.f;
  if (cudaMemcpyAsync(&hres, d, 4, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) return false;
  if (cudaStreamSynchronize(ctx->stream) != cudaSuccess) return false;
  return hres == 0.0f;
}

int stocs_launch_score(stocs_b200_ctx* ctx, const float* d_T, int64_t H, float* d_lcp, int32_t* d_inl,
                       cudaStream_t st, bool time_it, int slot, unsigned long long* d_counters, bool T_in_host_memory,
                       const long long* d_H, long long grid_hint) {
  if (H <= 0) return STOCS_OK;
  if (H >= (1ll << 31)) STOCS_FAIL(ctx, STOCS_E_ARG, "score: at most 2^31-1 hypotheses per call");
  // the kd-tree of a freshly uploaded scene may still be under construction on a host thread: collect
  // it and queue its upload in front of this launch (no-op on every later launch)
  if (ctx->kd_pending) { const int rc = stocs_kd_finish(ctx, st); if (rc) return rc; }
  ScoreArgs a;
  a.bricks = ctx->d_bricks.as<uint4>();
  a.coarse = ctx->d_coarse.as<uint32_t>();
  a.brick_occ = ctx->d_brick_occ.as<uint32_t>();
  a.coarse_words = ctx->coarse_words;
  a.coarse_shift = ctx->coarse_shift; a.coarse_nx = ctx->coarse_nx; a.coarse_ny = ctx->coarse_ny; a.coarse_nz = ctx->coarse_nz;
  a.coarse_sx = ctx->coarse_nx + 1; a.coarse_sy = ctx->coarse_ny + 1;
  a.starts = ctx->d_cell_start.as<uint32_t>();
  a.cand = ctx->d_cand.as<float4>();
  a.sattr = ctx->d_sattr.as<float4>();
  a.model = ctx->d_model.as<float>();
  a.kd_nodes = ctx->d_kd_nodes.as<KdNodeDev>();
  a.kd_pts = ctx->d_kd_pts.as<float4>();
  a.T = d_T;
  a.lcp = d_lcp;
  a.inl = d_inl;
  // slot 0 is the default work counter; launches that may run concurrently (score_lcp's chunks
  // on two streams) take distinct slots.  One tie counter accumulates over all launches.
  unsigned long long* ctr = (unsigned long long*)(ctx->d_small.as<char>() + 192);
  unsigned long long* wctr = slot == 0 ? ctr : (unsigned long long*)(ctx->d_small.as<char>() + 2048) + slot;
  a.work_counter = wctr;
  a.tie_counter = ctr + 1;
  a.cnt = d_counters;
  a.order = nullptr;
  a.H = H;
  a.H_dev = d_H;
  a.g = ctx->grid;
  a.M = ctx->M;
  a.Mpad = ctx->Mpad;
  a.sq_eps = ctx->eps * ctx->eps;
  a.dot_thr = ctx->dot_thr;
  a.to_block = 1.0f / (float)(1u << ctx->coarse_shift);
  a.to_cell = (float)(1u << ctx->coarse_shift);
  size_t smem = (size_t)kModelFloats * ctx->Mpad * 4 + (size_t)((a.coarse_words + 3) & ~3) * 4 + (size_t)kWarps * sizeof(WarpQueue);
  // static (per-warp queues) + dynamic (model, coarse bitmap) may exceed the 48 KB default
  // d_counters != NULL selects the counting variant (same code + one global atomic per warp event)
  void (*kernel)(ScoreArgs) = d_counters ? score_lcp_kernel<true> : score_lcp_kernel<false>;
  STOCS_CUDA(ctx, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  STOCS_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kWarps * 32, smem));
  if (per_sm < 1) STOCS_FAIL(ctx, STOCS_E_ARG, "score: model too large for shared memory");
  long long want = ((d_H && grid_hint > 0 && grid_hint < H ? grid_hint : H) + kWarps - 1) / kWarps;
  long long grid = (long long)ctx->num_sms * per_sm;
  if (grid > want) grid = want;
  STOCS_CUDA(ctx, cudaMemsetAsync(wctr, 0, 8, st));  // work counter only; tie counter accumulates
  // Heavy-first schedule for MID-SIZE launches only.  Measured on B200 (S1 hypotheses, kernel + probe):
  // 125 000 hypotheses (the per-GPU share of the named config on 8 GPUs) 0.347 -> 0.312 ms; 10^6
  // hypotheses 2.100 -> 2.139 ms (the probe costs 64 us per 10^6, the tail it removes is ~0.09 ms
  // whatever the size), and on the all-fitted S1-fit list the heavy-first order itself slows the
  // kernel by 2 % (3.622 -> 3.692 ms).  Small launches (the online pipeline's ~1 000 hypotheses)
  // cannot spare one more launch.  STOCS_NO_LPT=1 switches it off, STOCS_LPT_MAX overrides the bound.
  const bool no_lpt = getenv("STOCS_NO_LPT") != nullptr;
  const long long lpt_max = getenv("STOCS_LPT_MAX") ? atoll(getenv("STOCS_LPT_MAX")) : 300000;
  // (not when the transforms are read in place from page-locked host memory: the probe would pull
  // every transform over PCIe a second time -- measured 1.7e9 -> 0.9e9 hypotheses/s end to end on 8 GPUs)
  if (!no_lpt && !d_counters && !T_in_host_memory && !d_H && H >= 32768 && H <= lpt_max && slot == 0) {
    DevBuf& b_order = ctx->pool[POOL_SCORE_ORDER];
    STOCS_CUDA(ctx, b_order.ensure((size_t)H * 4));
    unsigned* fill = (unsigned*)(ctx->d_small.as<char>() + 3328);
    STOCS_CUDA(ctx, cudaMemsetAsync(fill, 0, 8, st));
    const size_t psm = 32 * 16 + (size_t)((a.coarse_words + 3) & ~3) * 4;
    probe_order_kernel<<<(unsigned)((H + 255) / 256), 256, psm, st>>>(a, b_order.as<int>(), fill, 16);
    a.order = b_order.as<int>();
  }
  if (time_it) {  // next pair of the ring (events are created on first use)
    const int k = (int)(ctx->ev_count % stocs_b200_ctx::kEvRing);
    for (int j = 0; j < 2; ++j)
      if (!ctx->ev_ring[2 * k + j]) STOCS_CUDA(ctx, cudaEventCreate(&ctx->ev_ring[2 * k + j]));
    ctx->ev0 = ctx->ev_ring[2 * k];
    ctx->ev1 = ctx->ev_ring[2 * k + 1];
    STOCS_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
  }
  kernel<<<(unsigned)grid, kWarps * 32, smem, st>>>(a);
  if (time_it) { STOCS_CUDA(ctx, cudaEventRecord(ctx->ev1, st)); ctx->ev_count++; }
  STOCS_CUDA(ctx, cudaGetLastError());
  ctx->counters[0] += 1;
  ctx->timing_valid = time_it;
  return STOCS_OK;
}
