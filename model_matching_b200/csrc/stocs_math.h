// stocs_math.h -- deterministic fp32/fp64 leaf arithmetic shared by the sm_100a kernels (nvcc)
// and by host code (g++).  Every function here is built ONLY from IEEE-754 add/sub/mul/div/sqrt
// and integer bit operations, evaluated in one fixed order, so g++ (-ffp-contract=off) and nvcc
// (-fmad=false) produce bit-identical results.  That is what makes "congruent-set indices and
// inlier counts bit-exact" achievable: all integer decisions of the StoCS hot path sit behind
// acos/atan2/atan/sin/cos/log2 calls (reference: src/rgbd.cpp:112-115, src/stocs.cpp:428,1028,
// include/super4pcs/accelerators/normalset.h:117, normalset.hpp:178-195), and libm on the two
// toolchains differs in the last ulp.
//
// Arithmetic model that is pinned here (see DESIGN.md "Pinned arithmetic model"):
//   * binary32, round-to-nearest-even, NO fused multiply-add (reference CMakeLists.txt:6-7 builds
//     -O3 -std=c++11 without -march => SSE2, no FMA).
//   * 3-term reductions (dot, squaredNorm, 3x3 products) follow Eigen's redux_novec_unroller
//     split for Length=3:  a + (b + c).
//   * Matrix4f * homogeneous(Vector3f) follows Eigen's homogeneous product
//     (lhs.leftCols<3>() * v, packet order) + lhs.col(3):   ((m0*x + m1*y) + m2*z) + m3.
//   * transcendental functions are evaluated in binary64 by the series below and rounded once to
//     binary32 (float overloads, as libstdc++'s <math.h> wrapper resolves them).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define STOCS_HD __host__ __device__ __forceinline__
#else
#define STOCS_HD inline
#endif

namespace stocsm {

static constexpr double kPi      = 0x1.921fb54442d18p+1;
static constexpr double kPiO2Hi  = 0x1.921fb54442d18p+0;
static constexpr double kPiO2Lo  = 0x1.1a62633145c07p-54;
static constexpr double kTwoOPi  = 0x1.45f306dc9c883p-1;
static constexpr double kInvLn2  = 0x1.71547652b82fep+0;
static constexpr double kSqrt2   = 0x1.6a09e667f3bcdp+0;

// ---- arithmetic-model switches -------------------------------------------------------------------
// For the sensitivity study of the unpinned leaf arithmetic ONLY (oracle/sensitivity.py builds oracle
// variants with them; the product and the default oracle define none of them):
//   STOCS_MODEL_SUM3_LEFT     3-term sums as (a + b) + c  (Eigen 3.2 coefficient-based products, and
//                             SURVEY.md section 8c's reading) instead of a + (b + c)
//   STOCS_MODEL_TRIG_DOUBLE   the reference's UNQUALIFIED atan2()/acos() calls (src/rgbd.cpp:113-115,
//                             src/stocs.cpp:428) bound to the double overloads: the whole
//                             "f(..)*180/M_PI" expression is then evaluated in binary64
//   STOCS_MODEL_TRIG_ULP=+1/-1  every float atan2f/acosf result one ulp above / below the correctly
//                             rounded value (what a libm that is not correctly rounded may return)

struct V3 { float x, y, z; };

STOCS_HD V3 v3(float x, float y, float z) { V3 r; r.x = x; r.y = y; r.z = z; return r; }
#ifdef STOCS_MODEL_SUM3_LEFT
STOCS_HD float sum3(float a, float b, float c) { return (a + b) + c; }
#else
STOCS_HD float sum3(float a, float b, float c) { return a + (b + c); }
#endif
STOCS_HD V3 sub(V3 a, V3 b) { return v3(a.x - b.x, a.y - b.y, a.z - b.z); }
STOCS_HD V3 add(V3 a, V3 b) { return v3(a.x + b.x, a.y + b.y, a.z + b.z); }
STOCS_HD V3 scale(V3 a, float s) { return v3(a.x * s, a.y * s, a.z * s); }
STOCS_HD V3 divs(V3 a, float s) { return v3(a.x / s, a.y / s, a.z / s); }
STOCS_HD float dot(V3 a, V3 b) { return sum3(a.x * b.x, a.y * b.y, a.z * b.z); }
STOCS_HD float sqnorm(V3 a) { return sum3(a.x * a.x, a.y * a.y, a.z * a.z); }
STOCS_HD float norm(V3 a) { return sqrtf(sqnorm(a)); }
STOCS_HD V3 cross(V3 a, V3 b) {
  return v3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
// Eigen 3.3 MatrixBase::normalized(): divide by sqrt(squaredNorm) when squaredNorm > 0,
// otherwise return the vector unchanged.
STOCS_HD V3 normalized(V3 a) {
  float z = sqnorm(a);
  if (z > 0.0f) return divs(a, sqrtf(z));
  return a;
}

// Column-major 4x4 (Eigen::Matrix4f memory order): element (r,c) = m[c*4+r].
STOCS_HD V3 xform_point(const float* m, V3 p) {
  V3 q;
  q.x = ((m[0] * p.x + m[4] * p.y) + m[8] * p.z) + m[12];
  q.y = ((m[1] * p.x + m[5] * p.y) + m[9] * p.z) + m[13];
  q.z = ((m[2] * p.x + m[6] * p.y) + m[10] * p.z) + m[14];
  return q;
}
// mat.block<3,3>(0,0) * n  (coefficient-based 3x3 product => redux order a + (b + c)).
STOCS_HD V3 xform_dir(const float* m, V3 n) {
  V3 q;
  q.x = sum3(m[0] * n.x, m[4] * n.y, m[8] * n.z);
  q.y = sum3(m[1] * n.x, m[5] * n.y, m[9] * n.z);
  q.z = sum3(m[2] * n.x, m[6] * n.y, m[10] * n.z);
  return q;
}

STOCS_HD uint64_t dbits(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
STOCS_HD double bitsd(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
STOCS_HD uint32_t fbits(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
STOCS_HD float bitsf(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
STOCS_HD bool isnan_d(double d) { return d != d; }
STOCS_HD double nan_d() { return bitsd(0x7ff8000000000000ull); }

// atan(x) for 0 <= x <= 1: x = c + d with c = k/8, atan(x) = atan(c) + atan(t),
// t = (x-c)/(1+x*c), |t| <= 1/16; odd Taylor series to t^15 (remainder < 2e-22).
STOCS_HD double atan_unit_d(double x) {
  const double tab[9] = {0x0.0p+0,
                         0x1.fd5ba9aac2f6ep-4, 0x1.f5b75f92c80ddp-3, 0x1.6f61941e4def1p-2,
                         0x1.dac670561bb4fp-2, 0x1.1e00babdefeb4p-1, 0x1.4978fa3269ee1p-1,
                         0x1.700a7c5784634p-1, 0x1.921fb54442d18p-1};
  int k = (int)(x * 8.0 + 0.5);
  double c = (double)k * 0.125;
  double t = (x - c) / (1.0 + x * c);
  double t2 = t * t;
  double p = 1.0 / 15.0;
  p = 1.0 / 13.0 - t2 * p;
  p = 1.0 / 11.0 - t2 * p;
  p = 1.0 / 9.0 - t2 * p;
  p = 1.0 / 7.0 - t2 * p;
  p = 1.0 / 5.0 - t2 * p;
  p = 1.0 / 3.0 - t2 * p;
  p = 1.0 - t2 * p;
  return tab[k] + t * p;
}

STOCS_HD double atan_d(double x) {
  if (isnan_d(x)) return x;
  double ax = x < 0.0 ? -x : x;
  double a = (ax <= 1.0) ? atan_unit_d(ax) : (kPiO2Hi - atan_unit_d(1.0 / ax));
  return x < 0.0 ? -a : a;
}

STOCS_HD double atan2_d(double y, double x) {
  if (isnan_d(x) || isnan_d(y)) return nan_d();
  bool xneg = (dbits(x) >> 63) != 0;
  bool yneg = (dbits(y) >> 63) != 0;
  double ax = xneg ? -x : x;
  double ay = yneg ? -y : y;
  double a;
  if (ay == 0.0) {
    a = xneg ? kPi : 0.0;
    return yneg ? -a : a;
  }
  if (ax == 0.0) {
    a = kPiO2Hi;
    return yneg ? -a : a;
  }
  if (ay <= ax) a = atan_unit_d(ay / ax);
  else a = kPiO2Hi - atan_unit_d(ax / ay);
  if (xneg) a = kPi - a;
  return yneg ? -a : a;
}

#if defined(STOCS_MODEL_TRIG_ULP)
STOCS_HD float model_ulp(float r) {
  uint32_t u; memcpy(&u, &r, 4);
  if ((u & 0x7fffffffu) == 0u || (u & 0x7f800000u) == 0x7f800000u) return r;   // zero, inf, NaN unchanged
  u = (uint32_t)((int32_t)u + ((STOCS_MODEL_TRIG_ULP) > 0 ? 1 : -1));           // magnitude +- 1 ulp (r >= 0 here)
  memcpy(&r, &u, 4);
  return r;
}
#else
STOCS_HD float model_ulp(float r) { return r; }
#endif

// Float overloads (round the binary64 result once).
STOCS_HD float atan2_f(float y, float x) { return model_ulp((float)atan2_d((double)y, (double)x)); }
STOCS_HD float atan_f(float x) { return (float)atan_d((double)x); }

// acos(x) = atan2(sqrt((1-x)(1+x)), x); (1-x) and (1+x) are exact in binary64 for binary32 x.
STOCS_HD float acos_f(float x) {
  if (!(x >= -1.0f && x <= 1.0f)) return bitsf(0x7fc00000u);
  double xd = (double)x;
  double s = sqrt((1.0 - xd) * (1.0 + xd));
  return model_ulp((float)atan2_d(s, xd));
}

// sin/cos for finite |x| < ~1e5: quadrant reduction with a two-part pi/2, Taylor to r^17 / r^18.
STOCS_HD void sincos_d(double x, double* s, double* c) {
  double ax = x < 0.0 ? -x : x;
  long long k = (long long)(ax * kTwoOPi + 0.5);
  double kd = (double)k;
  double r = (ax - kd * kPiO2Hi) - kd * kPiO2Lo;
  double r2 = r * r;
  double ps = 1.0 / 355687428096000.0;           // 1/17!
  ps = 1.0 / 1307674368000.0 - r2 * ps;          // 1/15!
  ps = 1.0 / 6227020800.0 - r2 * ps;             // 1/13!
  ps = 1.0 / 39916800.0 - r2 * ps;               // 1/11!
  ps = 1.0 / 362880.0 - r2 * ps;                 // 1/9!
  ps = 1.0 / 5040.0 - r2 * ps;                   // 1/7!
  ps = 1.0 / 120.0 - r2 * ps;                    // 1/5!
  ps = 1.0 / 6.0 - r2 * ps;                      // 1/3!
  ps = 1.0 - r2 * ps;
  double sr = r * ps;
  double pc = 1.0 / 6402373705728000.0;          // 1/18!
  pc = 1.0 / 20922789888000.0 - r2 * pc;         // 1/16!
  pc = 1.0 / 87178291200.0 - r2 * pc;            // 1/14!
  pc = 1.0 / 479001600.0 - r2 * pc;              // 1/12!
  pc = 1.0 / 3628800.0 - r2 * pc;                // 1/10!
  pc = 1.0 / 40320.0 - r2 * pc;                  // 1/8!
  pc = 1.0 / 720.0 - r2 * pc;                    // 1/6!
  pc = 1.0 / 24.0 - r2 * pc;                     // 1/4!
  pc = 1.0 / 2.0 - r2 * pc;                      // 1/2!
  double cr = 1.0 - r2 * pc;
  double sv, cv;
  switch ((int)(k & 3)) {
    case 0: sv = sr; cv = cr; break;
    case 1: sv = cr; cv = -sr; break;
    case 2: sv = -sr; cv = -cr; break;
    default: sv = -cr; cv = sr; break;
  }
  *s = x < 0.0 ? -sv : sv;
  *c = cv;
}
STOCS_HD float sin_f(float x) { double s, c; sincos_d((double)x, &s, &c); return (float)s; }
STOCS_HD float cos_f(float x) { double s, c; sincos_d((double)x, &s, &c); return (float)c; }

// log2 for positive normal binary32 inputs: exponent split, ln(m) = 2 atanh((m-1)/(m+1)).
STOCS_HD float log2_f(float xf) {
  if (!(xf > 0.0f)) return bitsf(0x7fc00000u);
  double x = (double)xf;
  uint64_t b = dbits(x);
  int e = (int)((b >> 52) & 0x7ff) - 1023;
  double m = bitsd((b & 0x000fffffffffffffull) | 0x3ff0000000000000ull);
  if (m > kSqrt2) { m = m * 0.5; e += 1; }
  double s = (m - 1.0) / (m + 1.0);
  double s2 = s * s;
  double p = 1.0 / 23.0;
  p = 1.0 / 21.0 + s2 * p;
  p = 1.0 / 19.0 + s2 * p;
  p = 1.0 / 17.0 + s2 * p;
  p = 1.0 / 15.0 + s2 * p;
  p = 1.0 / 13.0 + s2 * p;
  p = 1.0 / 11.0 + s2 * p;
  p = 1.0 / 9.0 + s2 * p;
  p = 1.0 / 7.0 + s2 * p;
  p = 1.0 / 5.0 + s2 * p;
  p = 1.0 / 3.0 + s2 * p;
  p = 1.0 + s2 * p;
  double ln_m = 2.0 * s * p;
  return (float)((double)e + ln_m * kInvLn2);
}

// "float * 180 / M_PI" exactly as the reference writes it: the float is multiplied by int 180
// in binary32, then divided by the double M_PI (src/rgbd.cpp:113, src/stocs.cpp:428,1028).
STOCS_HD double rad_to_deg_ref(float rad) { return (double)(rad * 180.0f) / kPi; }

// The two UNQUALIFIED call shapes of the reference.  With <cmath> in scope (it is: Eigen includes it)
// float arguments select the float overloads, which is the model pinned here; see
// STOCS_MODEL_TRIG_DOUBLE above for the other reading.
//   int(atan2(y, x)*180/M_PI)              src/rgbd.cpp:113-115
STOCS_HD double deg_atan2_ref(float y, float x) {
#ifdef STOCS_MODEL_TRIG_DOUBLE
  return atan2_d((double)y, (double)x) * 180.0 / kPi;
#else
  return rad_to_deg_ref(atan2_f(y, x));
#endif
}
//   float int_angle = acos(d)*180/M_PI     src/stocs.cpp:428
STOCS_HD float deg_acos_unqualified_ref(float d) {
#ifdef STOCS_MODEL_TRIG_DOUBLE
  if (!(d >= -1.0f && d <= 1.0f)) return bitsf(0x7fc00000u);
  const double xd = (double)d;
  return (float)(atan2_d(sqrt((1.0 - xd) * (1.0 + xd)), xd) * 180.0 / kPi);
#else
  return (float)rad_to_deg_ref(acos_f(d));
#endif
}

// ---------------------------------------------------------------------------------------------
// Point-pair feature (reference src/rgbd.cpp:85-121).
STOCS_HD int ppf_closest_bin(int value, int disc) {
  int lower = value - (value % disc);
  int upper = lower + disc;
  return ((value - lower) < (upper - value)) ? lower : upper;
}

struct Ppf4 { int f[4]; };

STOCS_HD Ppf4 ppf_compute(V3 p1, V3 n1, V3 p2, V3 n2, int tr_disc, int rot_disc) {
  V3 u = sub(p1, p2);
  int a1 = (int)(norm(u) * 1000.0f);
  int a2 = (int)deg_atan2_ref(norm(cross(n1, u)), dot(n1, u));
  int a3 = (int)deg_atan2_ref(norm(cross(n2, u)), dot(n2, u));
  int a4 = (int)deg_atan2_ref(norm(cross(n1, n2)), dot(n1, n2));
  Ppf4 r;
  r.f[0] = ppf_closest_bin(a1, tr_disc);
  r.f[1] = ppf_closest_bin(a2, rot_disc);
  r.f[2] = ppf_closest_bin(a3, rot_disc);
  r.f[3] = ppf_closest_bin(a4, rot_disc);
  return r;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 counter-based generator (Salmon et al., SC'11).  Used for reproducible base
// sampling in place of the reference's wall-clock-seeded std::default_random_engine
// (src/stocs.cpp:135-145).
struct Philox4 { uint32_t v[4]; };
STOCS_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                               uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * (uint64_t)c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * (uint64_t)c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  Philox4 o; o.v[0] = c0; o.v[1] = c1; o.v[2] = c2; o.v[3] = c3;
  return o;
}

// One 64-bit draw keyed by (seed, base number, draw number).
STOCS_HD uint64_t draw_u64(uint64_t seed, uint32_t base_no, uint32_t draw_no) {
  Philox4 o = philox4x32_10(base_no, draw_no, 0x53744f43u /* "StOC" */, 0u,
                            (uint32_t)seed, (uint32_t)(seed >> 32));
  return ((uint64_t)o.v[1] << 32) | (uint64_t)o.v[0];
}

// floor(a*b / 2^64)
STOCS_HD uint64_t mulhi_u64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * (unsigned __int128)b) >> 64);
#endif
}

// Fixed-point weight of a sampling probability: floor(p * 2^40).  Sums of up to 2^23 weights fit
// in uint64 and are associative, so a parallel scan and a sequential loop agree exactly.
STOCS_HD uint64_t prob_weight(float p) {
  if (!(p > 0.0f)) return 0ull;
  double w = (double)p * 1099511627776.0;
  if (w >= 1099511627776.0 * 1024.0) w = 1099511627776.0 * 1024.0;
  return (uint64_t)w;
}

}  // namespace stocsm
