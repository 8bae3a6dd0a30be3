// sample_common.cuh -- device helpers shared by the class-mode and instance-mode samplers.
#pragma once
#include "ppf_device.cuh"
#include "stocs_math.h"

namespace stocs_sample {
using namespace stocsm;

// segment_distance_and_invariants (reference src/stocs.cpp:155-222; Scalar = double)
__device__ inline double seg_dist_inv(V3 p1, V3 p2, V3 q1, V3 q2, double& inv1, double& inv2) {
  const double kSmall = 0.0001;
  const V3 u = sub(p2, p1), v = sub(q2, q1), w = sub(p1, q1);
  const double a = dot(u, u), b = dot(u, v), c = dot(v, v), d = dot(u, w), e = dot(v, w);
  const double f = a * c - b * b;
  double s1 = 0.0, s2 = f, t1 = 0.0, t2 = f;
  if (f < kSmall) {
    s1 = 0.0; s2 = 1.0; t1 = e; t2 = c;
  } else {
    s1 = (b * e - c * d);
    t1 = (a * e - b * d);
    if (s1 < 0.0) { s1 = 0.0; t1 = e; t2 = c; }
    else if (s1 > s2) { s1 = s2; t1 = e + b; t2 = c; }
  }
  if (t1 < 0.0) {
    t1 = 0.0;
    if (-d < 0.0) s1 = 0.0;
    else if (-d > a) s1 = s2;
    else { s1 = -d; s2 = a; }
  } else if (t1 > t2) {
    t1 = t2;
    if ((-d + b) < 0.0) s1 = 0;
    else if ((-d + b) > a) s1 = s2;
    else { s1 = (-d + b); s2 = a; }
  }
  inv1 = (fabs(s1) < kSmall ? 0.0 : s1 / s2);
  inv2 = (fabs(t1) < kSmall ? 0.0 : t1 / t2);
  const V3 r = sub(add(w, scale(u, (float)inv1)), scale(v, (float)inv2));
  return (double)norm(r);
}

__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}


// try_sampled_base (reference src/stocs.cpp:224-268): best of the 12 ordered segment pairings
__device__ inline bool order_base(const V3 b[4], int best[4], float& inv1, float& inv2) {
  float min_distance = 3.402823466e+38f;
  best[0] = best[1] = best[2] = best[3] = -1;
  inv1 = inv2 = 0.f;
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      if (i == j) continue;
      int k = 0; while (k == i || k == j) k++;
      int l = 0; while (l == i || l == j || l == k) l++;
      double li1, li2;
      const float sd = (float)seg_dist_inv(b[i], b[j], b[k], b[l], li1, li2);
      if (sd < min_distance) {
        min_distance = sd;
        best[0] = i; best[1] = j; best[2] = k; best[3] = l;
        inv1 = (float)li1; inv2 = (float)li2;
      }
    }
  return best[0] >= 0;
}

// plane through three points as Ax + By + Cz = 1 (reference src/stocs.cpp:458-479)
struct Plane { float A, B, C, denom; };
__device__ inline Plane fit_plane(V3 p1, V3 p2, V3 p3) {
  const double x1 = p1.x, y1 = p1.y, z1 = p1.z, x2 = p2.x, y2 = p2.y, z2 = p2.z, x3 = p3.x, y3 = p3.y, z3 = p3.z;
  Plane pl; pl.A = pl.B = pl.C = 0.f;
  pl.denom = (float)(-x3 * y2 * z1 + x2 * y3 * z1 + x3 * y1 * z2 - x1 * y3 * z2 - x2 * y1 * z3 + x1 * y2 * z3);
  if (pl.denom != 0) {
    pl.A = (float)((-y2 * z1 + y3 * z1 + y1 * z2 - y3 * z2 - y1 * z3 + y2 * z3) / pl.denom);
    pl.B = (float)((x2 * z1 - x3 * z1 - x1 * z2 + x3 * z2 + x1 * z3 - x2 * z3) / pl.denom);
    pl.C = (float)((-x2 * y1 + x3 * y1 + x1 * y2 - x3 * y2 - x1 * y3 + x2 * y3) / pl.denom);
  }
  return pl;
}

// predicates of the update passes (reference src/stocs.cpp:424-442 and :456-497); true = zero it
__device__ inline bool angle_too_small(V3 v_1, V3 p, V3 pb0) {
  const V3 v_2 = normalized(sub(p, pb0));
  return internal_angle_below_30(dot(v_1, v_2));   // ppf_device.cuh: the pinned predicate, decided by an fp32 estimate when safe
}
__device__ inline bool off_plane_or_too_close(const Plane& pl, V3 p, V3 pb0, V3 pb1, V3 pb2) {
  float planar = 10000.0f;
  if (pl.denom != 0) planar = (float)fabs((double)((pl.A * p.x + pl.B * p.y) + pl.C * p.z) - 1.0);
  return (planar > 0.015f) || (norm(sub(p, pb0)) < 0.01f) || (norm(sub(p, pb1)) < 0.01f) || (norm(sub(p, pb2)) < 0.01f);
}

}  // namespace stocs_sample
