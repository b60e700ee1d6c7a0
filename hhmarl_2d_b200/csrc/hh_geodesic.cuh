// hh_geodesic.cuh -- WGS84 geodesic direct / inverse problems in FP64 for sm_100a.
//
// Replaces, on device, what the reference reaches through warsim/utils/geodesics.py:12-24
// (third-party geographiclib==2.0, Geodesic.WGS84.Direct / Inverse): Karney's order-6 series
// (J. Geodesy 87, 2013) specialised for this workload --
//   * ellipsoid constants and the n-polynomials A3x / C3x are folded at compile time;
//   * divisions by series denominators are multiplications by folded reciprocals;
//   * degree-argument trigonometry uses exact quadrant reduction + sincospi;
//   * the inverse keeps the short-line start, the Newton iteration, the meridian and the
//     coincident-point cases; the antipodal (astroid) start and the equatorial special case are
//     unreachable inside the <= 0.5 deg map box at latitude 5..5.5 N and are omitted.
// Both entry points are __noinline__: they are ~1-2.5 k instructions and are called from
// several sites of the fused step kernel (instruction-cache footprint matters more than the
// call overhead).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace hh {

// libdevice's FP64 sincos/atan2/acos/hypot/fmod expand to 60-250 instructions at EVERY call site.
// The step kernel has ~80 such sites; inlined it is ~200 KB of SASS streamed once per warp and the
// profile (profiles/r1b_*) shows instruction fetch as the top stall.  One shared copy of each keeps
// the hot math inside the 32 KB L1.5 instruction cache; the call costs ~10 cycles.
namespace m {
static __device__ __noinline__ void sincos_(double x, double* s, double* c) { ::sincos(x, s, c); }
static __device__ __noinline__ void sincospi_(double x, double* s, double* c) { ::sincospi(x, s, c); }
static __device__ __noinline__ double atan2_(double y, double x) { return ::atan2(y, x); }
static __device__ __noinline__ double acos_(double x) { return ::acos(x); }
__device__ __forceinline__ double hypot_(double x, double y) { return sqrt(x * x + y * y); }  // operands are O(1): no scaling needed
static __device__ __noinline__ double fmod_(double x, double y) { return ::fmod(x, y); }
}  // namespace m

namespace geo {

constexpr double kA = 6378137.0;
constexpr double kF = 1.0 / 298.257223563;
constexpr double kF1 = 1.0 - kF;
constexpr double kE2 = kF * (2.0 - kF);
constexpr double kEp2 = kE2 / (kF1 * kF1);
constexpr double kN = kF / (2.0 - kF);
constexpr double kB = kA * kF1;
constexpr double kPi = 3.14159265358979323846;
constexpr double kDeg = kPi / 180.0;
constexpr double kTiny = 1.4916681462400413e-154;  // sqrt(DBL_MIN)
constexpr double kTol0 = 2.220446049250313e-16;    // DBL_EPSILON
constexpr double kTol2 = 1.4901161193847656e-08;   // sqrt(tol0)

constexpr double cpoly(int n, const double* p, double x) {
  double y = p[0];
  for (int i = 1; i <= n; ++i) y = y * x + p[i];
  return y;
}
// A3x[k], k = 0..5 (coefficients of eps^5 .. eps^0), polynomials in n
constexpr double a3x(int k) {
  constexpr double c[] = {-3, 128, -2, -3, 64, -1, -3, -1, 16, 3, -1, -2, 8, 1, -1, 2, 1, 1};
  int o = 0, kk = 0;
  for (int j = 5; j >= 0; --j) {
    int m = (6 - j - 1 < j) ? 6 - j - 1 : j;
    if (kk == k) return cpoly(m, c + o, kN) / c[o + m + 1];
    o += m + 2;
    ++kk;
  }
  return 0;
}
constexpr double c3x(int k) {
  constexpr double c[] = {3, 128, 2, 5, 128, -1, 3, 3, 64, -1, 0, 1, 8, -1, 1, 4,
                          5, 256, 1, 3, 128, -3, -2, 3, 64, 1, -3, 2, 32,
                          7, 512, -10, 9, 384, 5, -9, 5, 192,
                          7, 512, -14, 7, 512,
                          21, 2560};
  int o = 0, kk = 0;
  for (int l = 1; l < 6; ++l)
    for (int j = 5; j >= l; --j) {
      int m = (6 - j - 1 < j) ? 6 - j - 1 : j;
      if (kk == k) return cpoly(m, c + o, kN) / c[o + m + 1];
      o += m + 2;
      ++kk;
    }
  return 0;
}
#define HH_A3X(k) constexpr double kA3x##k = a3x(k);
HH_A3X(0) HH_A3X(1) HH_A3X(2) HH_A3X(3) HH_A3X(4) HH_A3X(5)
#undef HH_A3X
#define HH_C3X(k) constexpr double kC3x##k = c3x(k);
HH_C3X(0) HH_C3X(1) HH_C3X(2) HH_C3X(3) HH_C3X(4) HH_C3X(5) HH_C3X(6) HH_C3X(7)
HH_C3X(8) HH_C3X(9) HH_C3X(10) HH_C3X(11) HH_C3X(12) HH_C3X(13) HH_C3X(14)
#undef HH_C3X
constexpr double kEtol2 = 3.6424611488788524e-08;  // 0.1*tol2/sqrt(max(.001,|f|)*min(1,1-f/2)/2)

__device__ __forceinline__ double sq(double x) { return x * x; }

__device__ __forceinline__ void norm2(double& s, double& c) {
  double r = 1.0 / m::hypot_(s, c);
  s *= r;
  c *= r;
}

__device__ __forceinline__ double ang_round(double x) {
  const double z = 1.0 / 16.0;
  double y = fabs(x);
  double w = z - y;
  y = w > 0 ? z - w : y;
  return copysign(y, x);
}

__device__ __forceinline__ double ang_normalize(double x) {
  if (fabs(x) < 180.0) return x;
  double y = remainder(x, 360.0);
  return fabs(y) == 180.0 ? copysign(180.0, x) : y;
}

// sin/cos of an angle in degrees with exact reduction to [-45, 45]
__device__ __forceinline__ void sincosd(double x, double& sx, double& cx) {
  double q = rint(x / 90.0);
  double r = fma(-90.0, q, x);
  double s, c;
  m::sincospi_(r * (1.0 / 180.0), &s, &c);
  int iq = (int)q & 3;
  double ss = (iq & 1) ? c : s;
  double cc = (iq & 1) ? s : c;
  sx = (iq == 2 || iq == 3) ? -ss : ss;
  cx = (iq == 1 || iq == 2) ? -cc : cc;
  cx += 0.0;
  if (sx == 0.0) sx = copysign(sx, x);
}

__device__ __forceinline__ double atan2d(double y, double x) {
  int q = 0;
  if (fabs(y) > fabs(x)) {
    double t = x;
    x = y;
    y = t;
    q = 2;
  }
  if (signbit(x)) {
    x = -x;
    ++q;
  }
  double ang = m::atan2_(y, x) * (180.0 / kPi);
  if (q == 1) ang = copysign(180.0, y) - ang;
  else if (q == 2) ang = 90.0 - ang;
  else if (q == 3) ang = -90.0 + ang;
  return ang;
}

// Clenshaw sum of c1*sin(2x) + ... + cn*sin(2nx); coefficients passed highest-first free of
// arrays so that everything stays in registers.
__device__ __forceinline__ double sin_series6(double sinx, double cosx, double c1, double c2,
                                              double c3, double c4, double c5, double c6) {
  double ar = 2.0 * (cosx - sinx) * (cosx + sinx);
  double y1 = c6;                 // n = 6 even: y0 = 0, y1 = 0 then unrolled pairs
  double y0 = ar * y1 + c5;       // y1 = ar*0 - 0 + c6 ; y0 = ar*y1 - 0 + c5
  y1 = ar * y0 - y1 + c4;
  y0 = ar * y1 - y0 + c3;
  y1 = ar * y0 - y1 + c2;
  y0 = ar * y1 - y0 + c1;
  return 2.0 * sinx * cosx * y0;
}
__device__ __forceinline__ double sin_series5(double sinx, double cosx, double c1, double c2,
                                              double c3, double c4, double c5) {
  double ar = 2.0 * (cosx - sinx) * (cosx + sinx);
  double y0 = c5, y1 = 0.0;       // n = 5 odd: y0 = c5
  y1 = ar * y0 - y1 + c4;
  y0 = ar * y1 - y0 + c3;
  y1 = ar * y0 - y1 + c2;
  y0 = ar * y1 - y0 + c1;
  return 2.0 * sinx * cosx * y0;
}

struct C6 {
  double c1, c2, c3, c4, c5, c6;
};
struct C5 {
  double c1, c2, c3, c4, c5;
};

__device__ __forceinline__ double A1m1f(double eps) {
  double e2 = eps * eps;
  double t = (((1.0 * e2 + 4.0) * e2 + 64.0) * e2 + 0.0) * (1.0 / 256.0);
  return (t + eps) / (1.0 - eps);
}
__device__ __forceinline__ double A2m1f(double eps) {
  double e2 = eps * eps;
  double t = (((-11.0 * e2 - 28.0) * e2 - 192.0) * e2 + 0.0) * (1.0 / 256.0);
  return (t - eps) / (1.0 + eps);
}
__device__ __forceinline__ C6 C1f(double eps) {
  double e2 = eps * eps, d = eps;
  C6 c;
  c.c1 = d * ((-1.0 * e2 + 6.0) * e2 - 16.0) * (1.0 / 32.0);
  d *= eps;
  c.c2 = d * ((-9.0 * e2 + 64.0) * e2 - 128.0) * (1.0 / 2048.0);
  d *= eps;
  c.c3 = d * (9.0 * e2 - 16.0) * (1.0 / 768.0);
  d *= eps;
  c.c4 = d * (3.0 * e2 - 5.0) * (1.0 / 512.0);
  d *= eps;
  c.c5 = d * -7.0 * (1.0 / 1280.0);
  d *= eps;
  c.c6 = d * -7.0 * (1.0 / 2048.0);
  return c;
}
__device__ __forceinline__ C6 C1pf(double eps) {
  double e2 = eps * eps, d = eps;
  C6 c;
  c.c1 = d * ((205.0 * e2 - 432.0) * e2 + 768.0) * (1.0 / 1536.0);
  d *= eps;
  c.c2 = d * ((4005.0 * e2 - 4736.0) * e2 + 3840.0) * (1.0 / 12288.0);
  d *= eps;
  c.c3 = d * (-225.0 * e2 + 116.0) * (1.0 / 384.0);
  d *= eps;
  c.c4 = d * (-7173.0 * e2 + 2695.0) * (1.0 / 7680.0);
  d *= eps;
  c.c5 = d * 3467.0 * (1.0 / 7680.0);
  d *= eps;
  c.c6 = d * 38081.0 * (1.0 / 61440.0);
  return c;
}
__device__ __forceinline__ C6 C2f(double eps) {
  double e2 = eps * eps, d = eps;
  C6 c;
  c.c1 = d * ((1.0 * e2 + 2.0) * e2 + 16.0) * (1.0 / 32.0);
  d *= eps;
  c.c2 = d * ((35.0 * e2 + 64.0) * e2 + 384.0) * (1.0 / 2048.0);
  d *= eps;
  c.c3 = d * (15.0 * e2 + 80.0) * (1.0 / 768.0);
  d *= eps;
  c.c4 = d * (7.0 * e2 + 35.0) * (1.0 / 512.0);
  d *= eps;
  c.c5 = d * 63.0 * (1.0 / 1280.0);
  d *= eps;
  c.c6 = d * 77.0 * (1.0 / 2048.0);
  return c;
}
__device__ __forceinline__ double A3f(double eps) {
  return ((((kA3x0 * eps + kA3x1) * eps + kA3x2) * eps + kA3x3) * eps + kA3x4) * eps + kA3x5;
}
__device__ __forceinline__ C5 C3f(double eps) {
  C5 c;
  double mult = eps;
  c.c1 = mult * ((((kC3x0 * eps + kC3x1) * eps + kC3x2) * eps + kC3x3) * eps + kC3x4);
  mult *= eps;
  c.c2 = mult * (((kC3x5 * eps + kC3x6) * eps + kC3x7) * eps + kC3x8);
  mult *= eps;
  c.c3 = mult * ((kC3x9 * eps + kC3x10) * eps + kC3x11);
  mult *= eps;
  c.c4 = mult * (kC3x12 * eps + kC3x13);
  mult *= eps;
  c.c5 = mult * kC3x14;
  return c;
}
__device__ __forceinline__ double S6(double s, double c, const C6& k) {
  return sin_series6(s, c, k.c1, k.c2, k.c3, k.c4, k.c5, k.c6);
}
__device__ __forceinline__ double S5(double s, double c, const C5& k) {
  return sin_series5(s, c, k.c1, k.c2, k.c3, k.c4, k.c5);
}

// ---------------------------------------------------------------------------------------------
// Direct problem. Returns (lat2, lon2) in degrees. azi1 in degrees (any range), s12 in metres.
// ---------------------------------------------------------------------------------------------
static __device__ __noinline__ double2 direct(double lat1, double lon1, double azi1, double s12) {
  double salp1, calp1, sbet1, cbet1;
  sincosd(ang_round(ang_normalize(azi1)), salp1, calp1);
  sincosd(ang_round(lat1), sbet1, cbet1);
  sbet1 *= kF1;
  norm2(sbet1, cbet1);
  cbet1 = fmax(kTiny, cbet1);
  double salp0 = salp1 * cbet1;
  double calp0 = m::hypot_(calp1, salp1 * sbet1);
  double ssig1 = sbet1, somg1 = salp0 * sbet1;
  double csig1 = (sbet1 != 0.0 || calp1 != 0.0) ? cbet1 * calp1 : 1.0;
  double comg1 = csig1;
  norm2(ssig1, csig1);
  double k2 = sq(calp0) * kEp2;
  double eps = k2 / (2.0 * (1.0 + sqrt(1.0 + k2)) + k2);

  double A1m1 = A1m1f(eps);
  double B11 = S6(ssig1, csig1, C1f(eps));
  double s, c;
  m::sincos_(B11, &s, &c);
  double stau1 = ssig1 * c + csig1 * s;
  double ctau1 = csig1 * c - ssig1 * s;
  C5 c3 = C3f(eps);
  double A3c = -kF * salp0 * A3f(eps);
  double B31 = S5(ssig1, csig1, c3);

  double tau12 = s12 / (kB * (1.0 + A1m1));
  m::sincos_(tau12, &s, &c);
  double B12 = -S6(stau1 * c + ctau1 * s, ctau1 * c - stau1 * s, C1pf(eps));
  double sig12 = tau12 - (B12 - B11);
  double ssig12, csig12;
  m::sincos_(sig12, &ssig12, &csig12);
  double ssig2 = ssig1 * csig12 + csig1 * ssig12;
  double csig2 = csig1 * csig12 - ssig1 * ssig12;
  double sbet2 = calp0 * ssig2;
  double cbet2 = m::hypot_(salp0, calp0 * csig2);
  if (cbet2 == 0.0) cbet2 = csig2 = kTiny;
  double somg2 = salp0 * ssig2, comg2 = csig2;
  double omg12 = m::atan2_(somg2 * comg1 - comg2 * somg1, comg2 * comg1 + somg2 * somg1);
  double lam12 = omg12 + A3c * (sig12 + (S5(ssig2, csig2, c3) - B31));
  double lon12 = lam12 * (180.0 / kPi);
  double2 out;
  out.x = atan2d(sbet2, kF1 * cbet2);
  out.y = ang_normalize(ang_normalize(lon1) + ang_normalize(lon12));
  return out;
}

// ---------------------------------------------------------------------------------------------
// Direct problem for one simulator tick: s12 <= 3 km (aircraft <= 463 m, rockets <= 1029 m per 1 s tick,
// cmano_simulator.py:65-72) from a latitude within 1.1 deg of the map centre.  Same series as direct()
// (Karney order 6: A1, C1, C1', A3, C3), but every angle that is small in this regime -- B11 (<= 8.4e-4 rad),
// tau12 / sig12 / omg12 / the latitude change (<= 4.8e-4 rad) -- goes through a Taylor polynomial whose
// truncation error is below 1e-20 instead of through sincos / atan2, the reduced latitude comes from an
// expansion about the map centre (as in inverse_local), and the sine / cosine of the azimuth are the
// caller's (the step kernel has them as the heading vector of env_base.py:428).  Agreement with direct():
// <= 2e-15 deg in both coordinates (tests/test_emu_v4.py, tests/test_gpu_geodesic.py), i.e. the last bit or
// two of a coordinate.  Outside the regime it hands over to direct().
// ---------------------------------------------------------------------------------------------
constexpr double kPhi0d = 5.25 * (3.14159265358979323846 / 180.0);
constexpr double kSinPhi0d = 0.09150161866340238;   // sin(5.25 deg)
constexpr double kCosPhi0d = 0.9958049275746618;    // cos(5.25 deg)

// 1 / x for the well-scaled denominators of direct_short (0.9 .. 2.1): one rounding in the reciprocal, one in the
// product that follows -- the quotient may differ from a correctly rounded division in the last bit
__device__ __forceinline__ double rcp_(double x) {
#if defined(__CUDA_ARCH__)
  return __drcp_rn(x);
#else
  return 1.0 / x;
#endif
}

static __device__ __noinline__ double2 direct_short(double lat1, double lon1, double azi1, double salp1, double calp1,
                                                    double s12) {
  const double dl = lat1 * kDeg - kPhi0d;
  if (!(fabs(dl) < 0.02 && fabs(s12) <= 3000.0 && fabs(lon1) < 170.0)) return direct(lat1, lon1, azi1, s12);
  double sbet1, cbet1;
  {
    const double d2 = dl * dl;
    const double sd = dl * (1.0 + d2 * (-1.0 / 6.0 + d2 * (1.0 / 120.0 - d2 * (1.0 / 5040.0))));
    const double cd = 1.0 + d2 * (-0.5 + d2 * (1.0 / 24.0 - d2 * (1.0 / 720.0)));
    sbet1 = kF1 * (kSinPhi0d * cd + kCosPhi0d * sd);
    cbet1 = kCosPhi0d * cd - kSinPhi0d * sd;
    const double r = rsqrt(sbet1 * sbet1 + cbet1 * cbet1);
    sbet1 *= r;
    cbet1 *= r;
  }
  const double salp0 = salp1 * cbet1;
  // (ssig1, csig1) = (sbet1, cbet1 calp1) / calp0  (hypot(sbet1, cbet1 calp1) = calp0 for unit vectors);
  // calp0 >= |sbet1| ~ 0.09 at these latitudes
  const double calp0sq = calp1 * calp1 + sq(salp1 * sbet1);
  const double icalp0 = rsqrt(calp0sq);
  const double calp0 = calp0sq * icalp0;
  const double ssig1 = sbet1 * icalp0, csig1 = cbet1 * calp1 * icalp0;
  // eps = k2 / (2 (1 + sqrt(1 + k2)) + k2) as its Taylor series: k2 <= e'^2 = 0.00674, degree 8 is exact to 3e-19
  const double k2 = calp0sq * kEp2;
  const double eps =
      k2 * (1.0 / 4.0 +
            k2 * (-1.0 / 8.0 +
                  k2 * (5.0 / 64.0 +
                        k2 * (-7.0 / 128.0 +
                              k2 * (21.0 / 512.0 + k2 * (-33.0 / 1024.0 + k2 * (429.0 / 16384.0 + k2 * (-715.0 / 32768.0))))))));
  const double B11 = S6(ssig1, csig1, C1f(eps));
  double s, c;
  {
    const double x2 = B11 * B11;
    s = B11 * (1.0 + x2 * (-1.0 / 6.0 + x2 * (1.0 / 120.0)));
    c = 1.0 + x2 * (-0.5 + x2 * (1.0 / 24.0 - x2 * (1.0 / 720.0)));
  }
  const double stau1 = ssig1 * c + csig1 * s;
  const double ctau1 = csig1 * c - ssig1 * s;
  const C5 c3 = C3f(eps);
  const double A3c = -kF * salp0 * A3f(eps);
  const double B31 = S5(ssig1, csig1, c3);
  // 1 + A1m1 = (1 + t) / (1 - eps),  t = eps^2 (64 + eps^2 (4 + eps^2)) / 256 <= 7.1e-7:  1 / (1 + t) = 1 - t + t^2
  const double e2 = eps * eps;
  const double t1 = ((e2 + 4.0) * e2 + 64.0) * e2 * (1.0 / 256.0);
  const double tau12 = s12 * (1.0 / kB) * ((1.0 - eps) * (1.0 - t1 * (1.0 - t1)));
  {
    const double x2 = tau12 * tau12;
    s = tau12 * (1.0 + x2 * (-1.0 / 6.0 + x2 * (1.0 / 120.0)));
    c = 1.0 + x2 * (-0.5 + x2 * (1.0 / 24.0));
  }
  const double B12 = -S6(stau1 * c + ctau1 * s, ctau1 * c - stau1 * s, C1pf(eps));
  const double sig12 = tau12 - (B12 - B11);
  double ssig12, csig12m1;   // sin(sig12), cos(sig12) - 1
  {
    const double x2 = sig12 * sig12;
    ssig12 = sig12 * (1.0 + x2 * (-1.0 / 6.0 + x2 * (1.0 / 120.0)));
    csig12m1 = x2 * (-0.5 + x2 * (1.0 / 24.0));
  }
  const double dss = ssig1 * csig12m1 + csig1 * ssig12;      // ssig2 - ssig1
  const double ssig2 = ssig1 + dss;
  const double csig2 = csig1 + (csig1 * csig12m1 - ssig1 * ssig12);
  const double sbet2 = calp0 * ssig2;
  const double cbet2 = sqrt(salp0 * salp0 + sq(calp0 * csig2));
  double omg12;
  {
    // tan(omg) = salp0 tan(sig)  =>  tan(omg2 - omg1) = salp0 sin(sig12) / (csig1 csig2 + salp0^2 ssig1 ssig2):
    // no cancellation (direct() takes the difference of two O(0.1) products, which costs ~20 ulp of the
    // longitude for azimuths near 90 / 270 deg); the denominator is ~cos^2(beta) ~ 0.99
    const double t = salp0 * ssig12 * rcp_(csig1 * csig2 + salp0 * salp0 * (ssig1 * ssig2));
    const double t2 = t * t;
    omg12 = t * (1.0 + t2 * (-1.0 / 3.0 + t2 * (1.0 / 5.0)));
  }
  const double lam12 = omg12 + A3c * (sig12 + (S5(ssig2, csig2, c3) - B31));
  double dphi;
  {
    // sin(bet2 - bet1) without cancellation: dsb (cbet1 + sbet1 (sbet1 + sbet2) / (cbet1 + cbet2)),
    // dsb = sbet2 - sbet1 = calp0 (ssig2 - ssig1);  tan(phi2 - phi1) = f1 sin(dbet) / (f1^2 cb1 cb2 + sb1 sb2)
    const double sdb = calp0 * dss * (cbet1 + sbet1 * (sbet1 + sbet2) * rcp_(cbet1 + cbet2));
    const double t = kF1 * sdb * rcp_(kF1 * kF1 * cbet1 * cbet2 + sbet1 * sbet2);
    const double t2 = t * t;
    dphi = t * (1.0 + t2 * (-1.0 / 3.0 + t2 * (1.0 / 5.0)));
  }
  double2 out;
  out.x = lat1 + dphi * (180.0 / kPi);
  out.y = lon1 + lam12 * (180.0 / kPi);
  return out;
}

// One-tick move with the azimuth given in degrees (callers that do not keep a heading vector)
__device__ __forceinline__ double2 direct_tick(double lat1, double lon1, double azi1, double s12) {
  double sa, ca;
  sincosd(ang_round(ang_normalize(azi1)), sa, ca);
  return direct_short(lat1, lon1, azi1, sa, ca, s12);
}

// ---------------------------------------------------------------------------------------------
// Inverse problem. Returns (s12 [m], azi1 [deg in (-180, 180]]).
// ---------------------------------------------------------------------------------------------
struct Lam12Out {
  double lam12, dlam12, salp2, calp2, sig12, ssig1, csig1, ssig2, csig2, eps;
};

// reduced length m12b / b (Lengths() of Karney with only m12b requested)
__device__ __forceinline__ double m12b_only(double eps, double sig12, double ssig1, double csig1,
                                            double dn1, double ssig2, double csig2, double dn2) {
  double A1 = A1m1f(eps), A2 = A2m1f(eps);
  C6 ca = C1f(eps), cb = C2f(eps);
  double m0 = A1 - A2;
  A2 = 1.0 + A2;
  A1 = 1.0 + A1;
  C6 cc;
  cc.c1 = A1 * ca.c1 - A2 * cb.c1;
  cc.c2 = A1 * ca.c2 - A2 * cb.c2;
  cc.c3 = A1 * ca.c3 - A2 * cb.c3;
  cc.c4 = A1 * ca.c4 - A2 * cb.c4;
  cc.c5 = A1 * ca.c5 - A2 * cb.c5;
  cc.c6 = A1 * ca.c6 - A2 * cb.c6;
  double J12 = m0 * sig12 + (S6(ssig2, csig2, cc) - S6(ssig1, csig1, cc));
  return dn2 * (csig1 * ssig2) - dn1 * (ssig1 * csig2) - csig1 * csig2 * J12;
}

__device__ __forceinline__ void lambda12(double sbet1, double cbet1, double dn1, double sbet2,
                                         double cbet2, double dn2, double salp1, double calp1,
                                         double slam120, double clam120, bool diffp, Lam12Out& o) {
  if (sbet1 == 0.0 && calp1 == 0.0) calp1 = -kTiny;
  double salp0 = salp1 * cbet1;
  double calp0 = m::hypot_(calp1, salp1 * sbet1);
  double ssig1 = sbet1, somg1 = salp0 * sbet1;
  double csig1 = calp1 * cbet1, comg1 = csig1;
  norm2(ssig1, csig1);
  double salp2 = cbet2 != cbet1 ? salp0 / cbet2 : salp1;
  double calp2 =
      (cbet2 != cbet1 || fabs(sbet2) != -sbet1)
          ? sqrt(sq(calp1 * cbet1) + (cbet1 < -sbet1 ? (cbet2 - cbet1) * (cbet1 + cbet2)
                                                      : (sbet1 - sbet2) * (sbet1 + sbet2))) /
                cbet2
          : fabs(calp1);
  double ssig2 = sbet2, somg2 = salp0 * sbet2;
  double csig2 = calp2 * cbet2, comg2 = csig2;
  norm2(ssig2, csig2);
  double sig12 = m::atan2_(fmax(0.0, csig1 * ssig2 - ssig1 * csig2) + 0.0, csig1 * csig2 + ssig1 * ssig2);
  double somg12 = fmax(0.0, comg1 * somg2 - somg1 * comg2) + 0.0;
  double comg12 = comg1 * comg2 + somg1 * somg2;
  double eta = m::atan2_(somg12 * clam120 - comg12 * slam120, comg12 * clam120 + somg12 * slam120);
  double k2 = sq(calp0) * kEp2;
  double eps = k2 / (2.0 * (1.0 + sqrt(1.0 + k2)) + k2);
  C5 c3 = C3f(eps);
  double B312 = S5(ssig2, csig2, c3) - S5(ssig1, csig1, c3);
  double domg12 = -kF * A3f(eps) * salp0 * (sig12 + B312);
  o.lam12 = eta + domg12;
  o.dlam12 = 0.0;
  if (diffp) {
    if (calp2 == 0.0)
      o.dlam12 = -2.0 * kF1 * dn1 / sbet1;
    else
      o.dlam12 = m12b_only(eps, sig12, ssig1, csig1, dn1, ssig2, csig2, dn2) * kF1 / (calp2 * cbet2);
  }
  o.salp2 = salp2;
  o.calp2 = calp2;
  o.sig12 = sig12;
  o.ssig1 = ssig1;
  o.csig1 = csig1;
  o.ssig2 = ssig2;
  o.csig2 = csig2;
  o.eps = eps;
}

static __device__ __noinline__ double2 inverse(double lat1, double lon1, double lat2, double lon2) {
  // lon12 = AngDiff(lon1, lon2) with its error term (both |lon| < 180 here, so the
  // remainder() calls of the general formulation are identities)
  double u = ang_normalize(-lon1), v = ang_normalize(lon2);
  double d = u + v;
  double up = d - v, vpp = d - up;
  up -= u;
  vpp -= v;
  double t = d != 0.0 ? 0.0 - (up + vpp) : d;
  d = ang_normalize(d);  // second sumx(remainder(d), t): |d| < 180 in the map box
  {
    double s2 = d + t;
    double up2 = s2 - t, vpp2 = s2 - up2;
    up2 -= d;
    vpp2 -= t;
    t = s2 != 0.0 ? 0.0 - (up2 + vpp2) : s2;
    d = s2;
  }
  if (d == 0.0 || fabs(d) == 180.0) d = copysign(d, t == 0.0 ? lon2 - lon1 : -t);
  double lon12 = d, lon12s = t;
  int lonsign = signbit(lon12) ? -1 : 1;
  lon12 *= lonsign;
  lon12s *= lonsign;
  double lam12 = lon12 * kDeg;
  double slam12, clam12;
  {  // sincosde(lon12, lon12s)
    double q = rint(lon12 / 90.0);
    double r = ang_round(fma(-90.0, q, lon12) + lon12s);
    double s, c;
    m::sincospi_(r * (1.0 / 180.0), &s, &c);
    int iq = (int)q & 3;
    double ss = (iq & 1) ? c : s;
    double cc = (iq & 1) ? s : c;
    slam12 = (iq == 2 || iq == 3) ? -ss : ss;
    clam12 = (iq == 1 || iq == 2) ? -cc : cc;
    clam12 += 0.0;
    if (slam12 == 0.0) slam12 = copysign(slam12, lon12);
  }

  lat1 = ang_round(lat1);
  lat2 = ang_round(lat2);
  int swapp = fabs(lat1) < fabs(lat2) ? -1 : 1;
  if (swapp < 0) {
    double tt = lat1;
    lat1 = lat2;
    lat2 = tt;
    lonsign = -lonsign;
  }
  int latsign = signbit(lat1) ? 1 : -1;
  lat1 *= latsign;
  lat2 *= latsign;

  double sbet1, cbet1, sbet2, cbet2;
  sincosd(lat1, sbet1, cbet1);
  sbet1 *= kF1;
  norm2(sbet1, cbet1);
  cbet1 = fmax(kTiny, cbet1);
  sincosd(lat2, sbet2, cbet2);
  sbet2 *= kF1;
  norm2(sbet2, cbet2);
  cbet2 = fmax(kTiny, cbet2);
  if (cbet1 < -sbet1) {
    if (cbet2 == cbet1) sbet2 = copysign(sbet1, sbet2);
  } else {
    if (fabs(sbet2) == -sbet1) cbet2 = cbet1;
  }
  double dn1 = sqrt(1.0 + kEp2 * sq(sbet1));
  double dn2 = sqrt(1.0 + kEp2 * sq(sbet2));

  double s12b = 0.0, salp1 = 0.0, calp1 = 0.0, salp2 = 0.0, calp2 = 0.0;
  bool meridian = slam12 == 0.0;  // lat1 == -90 unreachable
  if (meridian) {
    calp1 = clam12;
    salp1 = slam12;
    calp2 = 1.0;
    salp2 = 0.0;
    double ssig1 = sbet1, csig1 = calp1 * cbet1, ssig2 = sbet2, csig2 = calp2 * cbet2;
    double sig12 = m::atan2_(fmax(0.0, csig1 * ssig2 - ssig1 * csig2) + 0.0, csig1 * csig2 + ssig1 * ssig2);
    // Lengths(n, ...) for s12b (m12b is only used for the sig12 >= 1 test, never true here)
    double A1 = 1.0 + A1m1f(kN);
    C6 ca = C1f(kN);
    double B1 = S6(ssig2, csig2, ca) - S6(ssig1, csig1, ca);
    s12b = A1 * (sig12 + B1);
    if (sig12 < 3.0 * kTiny || (sig12 < kTol0 && s12b < 0.0)) s12b = 0.0;
  } else {
    // ---- InverseStart (short-line branch; astroid start unreachable: csig12 >= 0)
    double sbet12 = sbet2 * cbet1 - cbet2 * sbet1, cbet12 = cbet2 * cbet1 + sbet2 * sbet1;
    double sbet12a = sbet2 * cbet1 + cbet2 * sbet1;
    bool shortline = cbet12 >= 0.0 && sbet12 < 0.5 && cbet2 * lam12 < 0.5;
    double somg12, comg12, dnm = 1.0;
    if (shortline) {
      double sbetm2 = sq(sbet1 + sbet2);
      sbetm2 /= sbetm2 + sq(cbet1 + cbet2);
      dnm = sqrt(1.0 + kEp2 * sbetm2);
      double omg12 = lam12 / (kF1 * dnm);
      m::sincos_(omg12, &somg12, &comg12);
    } else {
      somg12 = slam12;
      comg12 = clam12;
    }
    salp1 = cbet2 * somg12;
    calp1 = comg12 >= 0.0 ? sbet12 + cbet2 * sbet1 * sq(somg12) / (1.0 + comg12)
                          : sbet12a - cbet2 * sbet1 * sq(somg12) / (1.0 - comg12);
    double ssig12 = m::hypot_(salp1, calp1);
    double csig12 = sbet1 * sbet2 + cbet1 * cbet2 * comg12;
    double sig12 = -1.0;
    if (shortline && ssig12 < kEtol2) {
      salp2 = cbet1 * somg12;
      calp2 = sbet12 - cbet1 * sbet2 * (comg12 >= 0.0 ? sq(somg12) / (1.0 + comg12) : 1.0 - comg12);
      norm2(salp2, calp2);
      sig12 = m::atan2_(ssig12, csig12);
    }
    if (!(salp1 <= 0.0)) {
      norm2(salp1, calp1);
    } else {
      salp1 = 1.0;
      calp1 = 0.0;
    }
    if (sig12 >= 0.0) {
      s12b = sig12 * dnm;
    } else {
      // ---- Newton on lambda12(alp1) = lam12
      double salp1a = kTiny, calp1a = 1.0, salp1b = kTiny, calp1b = -1.0;
      bool tripn = false, tripb = false;
      Lam12Out o;
      for (int numit = 0;; ++numit) {
        lambda12(sbet1, cbet1, dn1, sbet2, cbet2, dn2, salp1, calp1, slam12, clam12, numit < 20, o);
        double vv = o.lam12, dv = o.dlam12;
        if (tripb || !(fabs(vv) >= (tripn ? 8.0 : 1.0) * kTol0) || numit == 83) break;
        if (vv > 0.0 && (numit > 20 || calp1 / salp1 > calp1b / salp1b)) {
          salp1b = salp1;
          calp1b = calp1;
        } else if (vv < 0.0 && (numit > 20 || calp1 / salp1 < calp1a / salp1a)) {
          salp1a = salp1;
          calp1a = calp1;
        }
        if (numit < 20 && dv > 0.0) {
          double dalp1 = -vv / dv;
          if (fabs(dalp1) < kPi) {
            double sd, cd;
            m::sincos_(dalp1, &sd, &cd);
            double nsalp1 = salp1 * cd + calp1 * sd;
            if (nsalp1 > 0.0) {
              calp1 = calp1 * cd - salp1 * sd;
              salp1 = nsalp1;
              norm2(salp1, calp1);
              tripn = fabs(vv) <= 16.0 * kTol0;
              continue;
            }
          }
        }
        salp1 = (salp1a + salp1b) / 2.0;
        calp1 = (calp1a + calp1b) / 2.0;
        norm2(salp1, calp1);
        tripn = false;
        tripb = (fabs(salp1a - salp1) + (calp1a - calp1) < kTol0 ||
                 fabs(salp1 - salp1b) + (calp1 - calp1b) < kTol0);
      }
      salp2 = o.salp2;
      calp2 = o.calp2;
      double A1 = 1.0 + A1m1f(o.eps);
      C6 ca = C1f(o.eps);
      double B1 = S6(o.ssig2, o.csig2, ca) - S6(o.ssig1, o.csig1, ca);
      s12b = A1 * (o.sig12 + B1);
    }
  }
  if (swapp < 0) {
    salp1 = salp2;
    calp1 = calp2;
  }
  salp1 *= swapp * lonsign;
  calp1 *= swapp * latsign;
  double2 out;
  out.x = 0.0 + s12b * kB;
  out.y = atan2d(salp1, calp1);
  return out;
}


// ---------------------------------------------------------------------------------------------
// Local inverse: Delambre's analogies with the ellipsoid's principal radii at the mid-latitude.
// Returns (s12 [m], azi1 [deg, (-180, 180]]).  Measured against the exact solver over the map box
// (tests/test_gpu_geodesic.py): |ds| <= 10 um, |dazi| <= 3e-8 deg for s <= 7 km and <= 2 cm,
// 1e-5 deg for s <= 80 km.  Callers use it only to decide threshold tests and fall back to
// inverse() inside a margin >= 100x those errors, so decisions are identical to the exact ones.
// ---------------------------------------------------------------------------------------------
// sin/cos of the map-centre latitude (5.25 deg: the env's box is lat 5 .. 5+map_size, env_base.py:43);
// mid-latitudes inside the box are within 0.3 deg of it, so a degree-7 expansion is exact to 1e-17.
constexpr double kPhi0 = 5.25 * kDeg;
constexpr double kSinPhi0 = 0.09150161866340238;   // sin(5.25 deg)
constexpr double kCosPhi0 = 0.9958049275746618;   // cos(5.25 deg)

__device__ __forceinline__ double2 inverse_local(double lat1, double lon1, double lat2, double lon2) {
  const double phim = 0.5 * (lat1 + lat2) * kDeg;
  const double hp = 0.5 * (lat2 - lat1) * kDeg, hl = 0.5 * (lon2 - lon1) * kDeg;
  double sm, cm;
  const double dl = phim - kPhi0;
  if (fabs(dl) < 0.02) {   // |dl| < 1.15 deg: always true inside the map box
    const double d2 = dl * dl;
    const double sd = dl * (1.0 + d2 * (-1.0 / 6.0 + d2 * (1.0 / 120.0 - d2 * (1.0 / 5040.0))));
    const double cd = 1.0 + d2 * (-0.5 + d2 * (1.0 / 24.0 - d2 * (1.0 / 720.0)));
    sm = kSinPhi0 * cd + kCosPhi0 * sd;
    cm = kCosPhi0 * cd - kSinPhi0 * sd;
  } else {
    m::sincos_(phim, &sm, &cm);
  }
  const double W2 = 1.0 - kE2 * sm * sm;
  const double iW = 1.0 / sqrt(W2);
  const double N = kA * iW, M = kA * (1.0 - kE2) * iW * iW * iW;
  // |hp|, |hl| <= 0.01 rad: degree-7 / degree-6 Taylor polynomials are exact to < 1e-18
  const double hl2 = hl * hl, hp2 = hp * hp;
  const double shl = hl * (1.0 + hl2 * (-1.0 / 6.0 + hl2 * (1.0 / 120.0 - hl2 * (1.0 / 5040.0))));
  const double shp = hp * (1.0 + hp2 * (-1.0 / 6.0 + hp2 * (1.0 / 120.0 - hp2 * (1.0 / 5040.0))));
  const double chl = 1.0 + hl2 * (-0.5 + hl2 * (1.0 / 24.0 - hl2 * (1.0 / 720.0)));
  const double chp = 1.0 + hp2 * (-0.5 + hp2 * (1.0 / 24.0 - hp2 * (1.0 / 720.0)));
  const double u = shl * cm, v = chl * shp;
  const double X = N * u, Y = M * v;
  const double h2 = u * u + v * v;
  const double fac = 1.0 + h2 * (1.0 / 6.0 + h2 * (3.0 / 40.0 + h2 * (15.0 / 336.0)));  // asin(h)/h
  double2 out;
  out.x = 2.0 * sqrt(X * X + Y * Y) * fac;
  const double z = (shl * sm) / (chl * chp);                  // tan(dA/2), |z| < 1e-3
  const double z2 = z * z;
  const double dA_half = z * (1.0 + z2 * (-1.0 / 3.0 + z2 * (1.0 / 5.0)));
  out.y = (m::atan2_(X, Y) - dA_half) * (180.0 / kPi);
  if (out.y > 180.0) out.y -= 360.0;
  if (out.y <= -180.0) out.y += 360.0;
  return out;
}

}  // namespace geo
}  // namespace hh
