// ppf_device.cuh -- device view of the compact PPF table.
//
// The reference stores every ordered model pair under up to 2*4*4*4 = 128 keys of a std::map
// (src/rgbd.cpp:123-154: keys {f1-tr, f1} x {f-2rot, f-rot, f, f+rot}^3, skipping p1 <= 5 and
// negative angles).  Here every ordered pair is stored ONCE, under its own bin
// (f1/tr, f2/rot, f3/rot, f4/rot), in a CSR sorted by (bin, id1, id2); a map lookup of key K
// becomes the union of the <= 128 own bins whose expansion contains K:
//   f1 in {K1, K1+tr},  f_j in {K_j - rot, K_j, K_j + rot, K_j + 2rot}.
// Key existence (map.find != end, src/stocs.cpp:403-405) is a bitmap over the expanded keys.
#pragma once
#include <stdint.h>

#include "stocs_math.h"

struct PpfView {
  const uint32_t* __restrict__ bin_start;  // [n1*na^3 + 1]
  const uint32_t* __restrict__ pairs;      // (id1 << 16) | id2, sorted by (bin, id1, id2)
  const uint32_t* __restrict__ keybits;    // [(n1+1)*(na+1)^3 bits]
  int n1, na, tr, rot;
};

__device__ __forceinline__ bool ppf_key_exists(const PpfView& v, const stocsm::Ppf4& f) {
  if (f.f[0] <= 5 || f.f[1] < 0 || f.f[2] < 0 || f.f[3] < 0) return false;
  const int k1 = f.f[0] / v.tr, k2 = f.f[1] / v.rot, k3 = f.f[2] / v.rot, k4 = f.f[3] / v.rot;
  const int nb = v.na + 1;
  if (k1 > v.n1 || k2 >= nb || k3 >= nb || k4 >= nb) return false;
  const uint32_t bit = (uint32_t)(((k1 * nb + k2) * nb + k3) * nb + k4);
  return (__ldg(v.keybits + (bit >> 5)) >> (bit & 31)) & 1u;
}

// Cheap necessary condition for ppf_key_exists(ppf_compute(p1, n1, p2, n2)): the distance component alone
// (first lines of ppf_compute, then the f[0] tests of ppf_key_exists).  False => the key cannot exist,
// whatever the three angles are, so callers may skip the atan2 evaluations; true decides nothing.
__device__ __forceinline__ bool ppf_distance_may_exist(const PpfView& v, stocsm::V3 p1, stocsm::V3 p2) {
  const int a1 = (int)(stocsm::norm(stocsm::sub(p1, p2)) * 1000.0f);
  const int f0 = stocsm::ppf_closest_bin(a1, v.tr);
  return !(f0 <= 5 || f0 / v.tr > v.n1);
}

// ---- fast exact evaluation of the PPF angles on the device ------------------------------------------
// The reference's angle features are int(atan2(y, x) * 180 / M_PI) with y = |n x u| >= 0: only the
// INTEGER part of the degree value enters the key.  stocs_math.h evaluates it the pinned way (binary64
// polynomial, correctly rounded to binary32, the reference's float * 180 / M_PI): ~150 binary64
// operations per angle, three angles per candidate, and the samplers evaluate thousands of candidates
// per base.  An fp32 estimate (CUDA atan2f, <= 2 ulp) lies within 1e-4 degree of the pinned value
// (2.2e-5 from the pinned value's own two fp32 roundings, 6.5e-5 from the estimate's), so whenever its
// fractional part is farther than 1e-3 from 0 and 1 -- 998 cases in 1000 -- its floor IS the pinned
// integer; the other cases, and anything that is not a finite value in [0, 180], take the pinned path.
__device__ __forceinline__ int deg_atan2_floor(float y, float x) {
  const float est = atan2f(y, x) * 57.29577951308232f;
  const float fl = floorf(est);
  const float fr = est - fl;
  if (fr > 1e-3f && fr < 0.999f && est > 0.f && est < 180.f) return (int)fl;   // (false for NaN)
  return (int)stocsm::deg_atan2_ref(y, x);
}

// ppf_compute (stocs_math.h, src/rgbd.cpp:85-121) with the three angles through deg_atan2_floor: same key
__device__ __forceinline__ stocsm::Ppf4 ppf_compute_dev(stocsm::V3 p1, stocsm::V3 n1, stocsm::V3 p2, stocsm::V3 n2,
                                                        int tr_disc, int rot_disc) {
  using namespace stocsm;
  const V3 u = sub(p1, p2);
  const int a1 = (int)(norm(u) * 1000.0f);
  const int a2 = deg_atan2_floor(norm(cross(n1, u)), dot(n1, u));
  const int a3 = deg_atan2_floor(norm(cross(n2, u)), dot(n2, u));
  const int a4 = deg_atan2_floor(norm(cross(n1, n2)), dot(n1, n2));
  Ppf4 r;
  r.f[0] = ppf_closest_bin(a1, tr_disc);
  r.f[1] = ppf_closest_bin(a2, rot_disc);
  r.f[2] = ppf_closest_bin(a3, rot_disc);
  r.f[3] = ppf_closest_bin(a4, rot_disc);
  return r;
}

// "internal angle < 30 degrees" (src/stocs.cpp:424-442): min(ang, 180 - ang) < 30 with ang the pinned
// float acos(d) * 180 / M_PI.  The estimate decides whenever it is more than 0.01 degree away from 30
// and 150 (its error is below 1e-4); NaN (|d| > 1) and the band around the thresholds take the pinned path.
__device__ __forceinline__ bool internal_angle_below_30(float d) {
  const float est = acosf(d) * 57.29577951308232f;
  if (est > 30.01f && est < 149.99f) return false;
  if (est < 29.99f || est > 150.01f) return true;
  float ang = stocsm::deg_acos_unqualified_ref(d);
  const float other = 180.0f - ang;
  ang = (other < ang) ? other : ang;
  return ang < 30.0f;
}

// Enumerates the own bins whose expansion contains key f; calls fn(bin_index) for each valid one
// (at most 128).  Returns false when the key cannot exist (p1 <= 5).
template <class Fn>
__device__ __forceinline__ bool ppf_for_each_source_bin(const PpfView& v, const stocsm::Ppf4& f, Fn fn) {
  if (f.f[0] <= 5 || f.f[1] < 0 || f.f[2] < 0 || f.f[3] < 0) return false;
  const int k1 = f.f[0] / v.tr, k2 = f.f[1] / v.rot, k3 = f.f[2] / v.rot, k4 = f.f[3] / v.rot;
  for (int a = 0; a < 2; ++a) {
    const int b1 = k1 + a;
    if (b1 >= v.n1) continue;
    for (int b = -1; b <= 2; ++b) {
      const int b2 = k2 + b;
      if (b2 < 0 || b2 >= v.na) continue;
      for (int c = -1; c <= 2; ++c) {
        const int b3 = k3 + c;
        if (b3 < 0 || b3 >= v.na) continue;
        for (int d = -1; d <= 2; ++d) {
          const int b4 = k4 + d;
          if (b4 < 0 || b4 >= v.na) continue;
          fn((uint32_t)(((b1 * v.na + b2) * v.na + b3) * v.na + b4));
        }
      }
    }
  }
  return true;
}

// The j-th (0..127) source bin of key (k1..k4), or 0xffffffff when out of range.
__device__ __forceinline__ uint32_t ppf_source_bin(const PpfView& v, int k1, int k2, int k3, int k4, int j) {
  const int b1 = k1 + (j >> 6), b2 = k2 + ((j >> 4) & 3) - 1, b3 = k3 + ((j >> 2) & 3) - 1, b4 = k4 + (j & 3) - 1;
  if (b1 >= v.n1 || b2 < 0 || b2 >= v.na || b3 < 0 || b3 >= v.na || b4 < 0 || b4 >= v.na) return 0xffffffffu;
  return (uint32_t)(((b1 * v.na + b2) * v.na + b3) * v.na + b4);
}

struct stocs_b200_ctx;
PpfView stocs_ppf_view(const stocs_b200_ctx* ctx);  // ppf_table.cu
