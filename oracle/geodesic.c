/*
 * oracle/geodesic.c -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * WGS84 geodesic Direct/Inverse following Karney (2013), order-6 series, restating the
 * arithmetic of the third-party package geographiclib==2.0 that the reference calls at
 * warsim/utils/geodesics.py:12-24.  See geodesic.h for the parity statement.
 *
 * Build: gcc -O2 -ffp-contract=off (CPython never contracts a*b+c into an FMA).
 */
#include "geodesic.h"

#include <float.h>
#include <math.h>

#define ORD 6 /* series order: nA1 = nC1 = nC1p = nA2 = nC2 = nA3 = nC3 = 6 */
#define NC3X 15

static const double WGS84_A = 6378137.0;
static const double WGS84_F = 1.0 / 298.257223563;

static const double QD = 90.0, HD = 180.0, TD = 360.0;

typedef struct {
  double a, f, f1, e2, ep2, n, b, etol2;
  double A3x[ORD], C3x[NC3X];
  double tiny, tol0, tol1, tol2, tolb, xthresh, pi, degree;
  int maxit1, maxit2;
  int ready;
} geod_t;

static geod_t G;
static __thread int g_last_numit = 0;

static double sq(double x) { return x * x; }

static double polyval(int N, const double* p, double x) {
  double y = N < 0 ? 0 : *p++;
  while (--N >= 0) y = y * x + *p++;
  return y;
}

static void norm2(double* s, double* c) {
  double r = hypot(*s, *c);
  *s /= r;
  *c /= r;
}

static double sumx(double u, double v, double* t) {
  volatile double s = u + v;
  volatile double up = s - v;
  volatile double vpp = s - up;
  up -= u;
  vpp -= v;
  if (t) *t = s != 0 ? 0 - (up + vpp) : s;
  return s;
}

static double AngRound(double x) {
  const double z = 1 / 16.0;
  volatile double y = fabs(x);
  volatile double w = z - y;
  y = w > 0 ? z - w : y;
  return copysign(y, x);
}

static double AngNormalize(double x) {
  double y = remainder(x, TD);
  return fabs(y) == HD ? copysign(HD, x) : y;
}

static double LatFix(double x) { return fabs(x) > QD ? NAN : x; }

static double AngDiff(double x, double y, double* e) {
  double t, d = sumx(remainder(-x, TD), remainder(y, TD), &t);
  d = sumx(remainder(d, TD), t, &t);
  if (d == 0 || fabs(d) == HD) d = copysign(d, t == 0 ? y - x : -t);
  if (e) *e = t;
  return d;
}

static void sincosdx(double x, double* sinx, double* cosx) {
  double r, s, c;
  int q = 0;
  r = remquo(x, QD, &q);
  r *= G.degree;
  s = sin(r);
  c = cos(r);
  switch ((unsigned)q & 3U) {
    case 0U: *sinx = s; *cosx = c; break;
    case 1U: *sinx = c; *cosx = -s; break;
    case 2U: *sinx = -s; *cosx = -c; break;
    default: *sinx = -c; *cosx = s; break;
  }
  *cosx += 0;
  if (*sinx == 0) *sinx = copysign(*sinx, x);
}

static void sincosde(double x, double t, double* sinx, double* cosx) {
  double r, s, c;
  int q = 0;
  r = AngRound(remquo(x, QD, &q) + t);
  r *= G.degree;
  s = sin(r);
  c = cos(r);
  switch ((unsigned)q & 3U) {
    case 0U: *sinx = s; *cosx = c; break;
    case 1U: *sinx = c; *cosx = -s; break;
    case 2U: *sinx = -s; *cosx = -c; break;
    default: *sinx = -c; *cosx = s; break;
  }
  *cosx += 0;
  if (*sinx == 0) *sinx = copysign(*sinx, x);
}

static double atan2dx(double y, double x) {
  int q = 0;
  double ang;
  if (fabs(y) > fabs(x)) { double t = x; x = y; y = t; q = 2; }
  if (signbit(x)) { x = -x; ++q; }
  ang = atan2(y, x) / G.degree;
  switch (q) {
    case 1: ang = copysign(HD, y) - ang; break;
    case 2: ang = QD - ang; break;
    case 3: ang = -QD + ang; break;
    default: break;
  }
  return ang;
}

/* Clenshaw summation of sum(c[i] * sin(2*i*x), i = 1..n); c[0] unused. */
static double SinSeries(double sinx, double cosx, const double c[], int n) {
  double ar, y0, y1;
  c += (n + 1);
  ar = 2 * (cosx - sinx) * (cosx + sinx);
  y0 = (n & 1) ? *--c : 0;
  y1 = 0;
  n /= 2;
  while (n--) {
    y1 = ar * y0 - y1 + *--c;
    y0 = ar * y1 - y0 + *--c;
  }
  return 2 * sinx * cosx * y0;
}

static double A1m1f(double eps) {
  static const double coeff[] = {1, 4, 64, 0, 256};
  int m = ORD / 2;
  double t = polyval(m, coeff, sq(eps)) / coeff[m + 1];
  return (t + eps) / (1 - eps);
}

static void C1f(double eps, double c[]) {
  static const double coeff[] = {
      -1, 6, -16, 32, -9, 64, -128, 2048, 9, -16, 768, 3, -5, 512, -7, 1280, -7, 2048,
  };
  double eps2 = sq(eps), d = eps;
  int o = 0, l;
  for (l = 1; l <= ORD; ++l) {
    int m = (ORD - l) / 2;
    c[l] = d * polyval(m, coeff + o, eps2) / coeff[o + m + 1];
    o += m + 2;
    d *= eps;
  }
}

static void C1pf(double eps, double c[]) {
  static const double coeff[] = {
      205, -432, 768, 1536, 4005, -4736, 3840, 12288, -225, 116, 384,
      -7173, 2695, 7680, 3467, 7680, 38081, 61440,
  };
  double eps2 = sq(eps), d = eps;
  int o = 0, l;
  for (l = 1; l <= ORD; ++l) {
    int m = (ORD - l) / 2;
    c[l] = d * polyval(m, coeff + o, eps2) / coeff[o + m + 1];
    o += m + 2;
    d *= eps;
  }
}

static double A2m1f(double eps) {
  static const double coeff[] = {-11, -28, -192, 0, 256};
  int m = ORD / 2;
  double t = polyval(m, coeff, sq(eps)) / coeff[m + 1];
  return (t - eps) / (1 + eps);
}

static void C2f(double eps, double c[]) {
  static const double coeff[] = {
      1, 2, 16, 32, 35, 64, 384, 2048, 15, 80, 768, 7, 35, 512, 63, 1280, 77, 2048,
  };
  double eps2 = sq(eps), d = eps;
  int o = 0, l;
  for (l = 1; l <= ORD; ++l) {
    int m = (ORD - l) / 2;
    c[l] = d * polyval(m, coeff + o, eps2) / coeff[o + m + 1];
    o += m + 2;
    d *= eps;
  }
}

static void A3coeff(geod_t* g) {
  static const double coeff[] = {
      -3, 128, -2, -3, 64, -1, -3, -1, 16, 3, -1, -2, 8, 1, -1, 2, 1, 1,
  };
  int o = 0, k = 0, j;
  for (j = ORD - 1; j >= 0; --j) {
    int m = ORD - j - 1 < j ? ORD - j - 1 : j;
    g->A3x[k++] = polyval(m, coeff + o, g->n) / coeff[o + m + 1];
    o += m + 2;
  }
}

static void C3coeff(geod_t* g) {
  static const double coeff[] = {
      3, 128, 2, 5, 128, -1, 3, 3, 64, -1, 0, 1, 8, -1, 1, 4,
      5, 256, 1, 3, 128, -3, -2, 3, 64, 1, -3, 2, 32,
      7, 512, -10, 9, 384, 5, -9, 5, 192,
      7, 512, -14, 7, 512,
      21, 2560,
  };
  int o = 0, k = 0, l, j;
  for (l = 1; l < ORD; ++l) {
    for (j = ORD - 1; j >= l; --j) {
      int m = ORD - j - 1 < j ? ORD - j - 1 : j;
      g->C3x[k++] = polyval(m, coeff + o, g->n) / coeff[o + m + 1];
      o += m + 2;
    }
  }
}

static double A3f(const geod_t* g, double eps) { return polyval(ORD - 1, g->A3x, eps); }

static void C3f(const geod_t* g, double eps, double c[]) {
  double mult = 1;
  int o = 0, l;
  for (l = 1; l < ORD; ++l) {
    int m = ORD - l - 1;
    mult *= eps;
    c[l] = mult * polyval(m, g->C3x + o, eps);
    o += m + 1;
  }
}

static void geod_init(void) {
  geod_t* g = &G;
  if (g->ready) return;
  g->pi = atan2(0.0, -1.0);
  g->degree = g->pi / HD;
  g->maxit1 = 20;
  g->maxit2 = g->maxit1 + DBL_MANT_DIG + 10;
  g->tiny = sqrt(DBL_MIN);
  g->tol0 = DBL_EPSILON;
  g->tol1 = 200 * g->tol0;
  g->tol2 = sqrt(g->tol0);
  g->tolb = g->tol0;
  g->xthresh = 1000 * g->tol2;
  g->a = WGS84_A;
  g->f = WGS84_F;
  g->f1 = 1 - g->f;
  g->e2 = g->f * (2 - g->f);
  g->ep2 = g->e2 / sq(g->f1);
  g->n = g->f / (2 - g->f);
  g->b = g->a * g->f1;
  g->etol2 = 0.1 * g->tol2 / sqrt(fmax(0.001, fabs(g->f)) * fmin(1.0, 1 - g->f / 2) / 2);
  A3coeff(g);
  C3coeff(g);
  g->ready = 1;
}

/* ---------------------------------------------------------------- Direct */
void orc_geod_direct(double lat1, double lon1, double azi1, double s12,
                     double* plat2, double* plon2, double* pazi2) {
  const geod_t* g;
  double salp1, calp1, sbet1, cbet1, salp0, calp0, ssig1, csig1, somg1, comg1, k2, eps;
  double A1m1, C1a[ORD + 1], C1pa[ORD + 1], C3a[ORD], B11, stau1, ctau1, A3c, B31, s, c;
  double tau12, B12, sig12, ssig12, csig12, ssig2, csig2, sbet2, cbet2, salp2, calp2;
  double somg2, comg2, omg12, lam12, lon12;
  geod_init();
  g = &G;

  azi1 = AngNormalize(azi1);
  sincosdx(AngRound(azi1), &salp1, &calp1);
  lat1 = LatFix(lat1);
  sincosdx(AngRound(lat1), &sbet1, &cbet1);
  sbet1 *= g->f1;
  norm2(&sbet1, &cbet1);
  cbet1 = fmax(g->tiny, cbet1);
  salp0 = salp1 * cbet1;
  calp0 = hypot(calp1, salp1 * sbet1);
  ssig1 = sbet1;
  somg1 = salp0 * sbet1;
  csig1 = comg1 = (sbet1 != 0 || calp1 != 0) ? cbet1 * calp1 : 1;
  norm2(&ssig1, &csig1);
  k2 = sq(calp0) * g->ep2;
  eps = k2 / (2 * (1 + sqrt(1 + k2)) + k2);

  A1m1 = A1m1f(eps);
  C1f(eps, C1a);
  B11 = SinSeries(ssig1, csig1, C1a, ORD);
  s = sin(B11);
  c = cos(B11);
  stau1 = ssig1 * c + csig1 * s;
  ctau1 = csig1 * c - ssig1 * s;
  C1pf(eps, C1pa);
  C3f(g, eps, C3a);
  A3c = -g->f * salp0 * A3f(g, eps);
  B31 = SinSeries(ssig1, csig1, C3a, ORD - 1);

  tau12 = s12 / (g->b * (1 + A1m1));
  s = sin(tau12);
  c = cos(tau12);
  B12 = -SinSeries(stau1 * c + ctau1 * s, ctau1 * c - stau1 * s, C1pa, ORD);
  sig12 = tau12 - (B12 - B11);
  ssig12 = sin(sig12);
  csig12 = cos(sig12);
  /* |f| <= 0.01: no Newton refinement */
  ssig2 = ssig1 * csig12 + csig1 * ssig12;
  csig2 = csig1 * csig12 - ssig1 * ssig12;
  sbet2 = calp0 * ssig2;
  cbet2 = hypot(salp0, calp0 * csig2);
  if (cbet2 == 0) cbet2 = csig2 = g->tiny;
  salp2 = salp0;
  calp2 = calp0 * csig2;

  somg2 = salp0 * ssig2;
  comg2 = csig2;
  omg12 = atan2(somg2 * comg1 - comg2 * somg1, comg2 * comg1 + somg2 * somg1);
  lam12 = omg12 + A3c * (sig12 + (SinSeries(ssig2, csig2, C3a, ORD - 1) - B31));
  lon12 = lam12 / g->degree;
  if (plon2) *plon2 = AngNormalize(AngNormalize(lon1) + AngNormalize(lon12));
  if (plat2) *plat2 = atan2dx(sbet2, g->f1 * cbet2);
  if (pazi2) *pazi2 = atan2dx(salp2, calp2);
}

/* ---------------------------------------------------------------- Inverse helpers */
static void Lengths(const geod_t* g, double eps, double sig12, double ssig1, double csig1,
                    double dn1, double ssig2, double csig2, double dn2, double* ps12b,
                    double* pm12b, double Ca[]) {
  double m0 = 0, J12 = 0, A1 = 0, A2 = 0;
  double Cb[ORD + 1];
  int redlp = pm12b != 0, l;
  A1 = A1m1f(eps);
  C1f(eps, Ca);
  if (redlp) {
    A2 = A2m1f(eps);
    C2f(eps, Cb);
    m0 = A1 - A2;
    A2 = 1 + A2;
  }
  A1 = 1 + A1;
  if (ps12b) {
    double B1 = SinSeries(ssig2, csig2, Ca, ORD) - SinSeries(ssig1, csig1, Ca, ORD);
    *ps12b = A1 * (sig12 + B1);
    if (redlp) {
      double B2 = SinSeries(ssig2, csig2, Cb, ORD) - SinSeries(ssig1, csig1, Cb, ORD);
      J12 = m0 * sig12 + (A1 * B1 - A2 * B2);
    }
  } else if (redlp) {
    for (l = 1; l <= ORD; ++l) Cb[l] = A1 * Ca[l] - A2 * Cb[l];
    J12 = m0 * sig12 + (SinSeries(ssig2, csig2, Cb, ORD) - SinSeries(ssig1, csig1, Cb, ORD));
  }
  if (pm12b) *pm12b = dn2 * (csig1 * ssig2) - dn1 * (ssig1 * csig2) - csig1 * csig2 * J12;
}

static double Astroid(double x, double y) {
  double k;
  double p = sq(x), q = sq(y), r = (p + q - 1) / 6;
  if (!(q == 0 && r <= 0)) {
    double S = p * q / 4, r2 = sq(r), r3 = r * r2, disc = S * (S + 2 * r3);
    double u = r, v, uv, w;
    if (disc >= 0) {
      double T3 = S + r3, T;
      T3 += T3 < 0 ? -sqrt(disc) : sqrt(disc);
      T = cbrt(T3);
      u += T + (T != 0 ? r2 / T : 0);
    } else {
      double ang = atan2(sqrt(-disc), -(S + r3));
      u += 2 * r * cos(ang / 3);
    }
    v = sqrt(sq(u) + q);
    uv = u < 0 ? q / (v - u) : u + v;
    w = (uv - q) / (2 * v);
    k = uv / (sqrt(uv + sq(w)) + w);
  } else {
    k = 0;
  }
  return k;
}

static double InverseStart(const geod_t* g, double sbet1, double cbet1, double dn1,
                           double sbet2, double cbet2, double dn2, double lam12, double slam12,
                           double clam12, double* psalp1, double* pcalp1, double* psalp2,
                           double* pcalp2, double* pdnm) {
  double salp1 = 0, calp1 = 0, salp2 = 0, calp2 = 0, dnm = 0;
  double sig12 = -1;
  double sbet12 = sbet2 * cbet1 - cbet2 * sbet1, cbet12 = cbet2 * cbet1 + sbet2 * sbet1;
  double sbet12a;
  int shortline = cbet12 >= 0 && sbet12 < 0.5 && cbet2 * lam12 < 0.5;
  double somg12, comg12, ssig12, csig12;
  (void)dn1;
  (void)dn2;
  sbet12a = sbet2 * cbet1 + cbet2 * sbet1;
  if (shortline) {
    double sbetm2 = sq(sbet1 + sbet2), omg12;
    sbetm2 /= sbetm2 + sq(cbet1 + cbet2);
    dnm = sqrt(1 + g->ep2 * sbetm2);
    omg12 = lam12 / (g->f1 * dnm);
    somg12 = sin(omg12);
    comg12 = cos(omg12);
  } else {
    somg12 = slam12;
    comg12 = clam12;
  }
  salp1 = cbet2 * somg12;
  calp1 = comg12 >= 0 ? sbet12 + cbet2 * sbet1 * sq(somg12) / (1 + comg12)
                      : sbet12a - cbet2 * sbet1 * sq(somg12) / (1 - comg12);
  ssig12 = hypot(salp1, calp1);
  csig12 = sbet1 * sbet2 + cbet1 * cbet2 * comg12;
  if (shortline && ssig12 < g->etol2) {
    salp2 = cbet1 * somg12;
    calp2 = sbet12 - cbet1 * sbet2 * (comg12 >= 0 ? sq(somg12) / (1 + comg12) : 1 - comg12);
    norm2(&salp2, &calp2);
    sig12 = atan2(ssig12, csig12);
  } else if (fabs(g->n) > 0.1 || csig12 >= 0 || ssig12 >= 6 * fabs(g->n) * g->pi * sq(cbet1)) {
    /* zeroth-order spherical approximation is good enough */
  } else {
    double x, y, lamscale, betscale;
    double lam12x = atan2(-slam12, -clam12);
    double k2 = sq(sbet1) * g->ep2, eps = k2 / (2 * (1 + sqrt(1 + k2)) + k2);
    lamscale = g->f * cbet1 * A3f(g, eps) * g->pi;
    betscale = lamscale * cbet1;
    x = lam12x / lamscale;
    y = sbet12a / betscale;
    if (y > -g->tol1 && x > -1 - g->xthresh) {
      salp1 = fmin(1.0, -x);
      calp1 = -sqrt(1 - sq(salp1));
    } else {
      double k = Astroid(x, y);
      double omg12a = lamscale * (-x * k / (1 + k));
      somg12 = sin(omg12a);
      comg12 = -cos(omg12a);
      salp1 = cbet2 * somg12;
      calp1 = sbet12a - cbet2 * sbet1 * sq(somg12) / (1 - comg12);
    }
  }
  if (!(salp1 <= 0)) {
    norm2(&salp1, &calp1);
  } else {
    salp1 = 1;
    calp1 = 0;
  }
  *psalp1 = salp1;
  *pcalp1 = calp1;
  if (shortline) *pdnm = dnm;
  if (sig12 >= 0) {
    *psalp2 = salp2;
    *pcalp2 = calp2;
  }
  return sig12;
}

static double Lambda12(const geod_t* g, double sbet1, double cbet1, double dn1, double sbet2,
                       double cbet2, double dn2, double salp1, double calp1, double slam120,
                       double clam120, double* psalp2, double* pcalp2, double* psig12,
                       double* pssig1, double* pcsig1, double* pssig2, double* pcsig2,
                       double* peps, int diffp, double* pdlam12, double Ca[]) {
  double salp2 = 0, calp2 = 0, sig12 = 0, ssig1 = 0, csig1 = 0, ssig2 = 0, csig2 = 0, eps = 0;
  double dlam12 = 0, salp0, calp0, somg1, comg1, somg2, comg2, somg12, comg12, lam12;
  double B312, eta, k2, domg12;
  if (sbet1 == 0 && calp1 == 0) calp1 = -g->tiny;
  salp0 = salp1 * cbet1;
  calp0 = hypot(calp1, salp1 * sbet1);
  ssig1 = sbet1;
  somg1 = salp0 * sbet1;
  csig1 = comg1 = calp1 * cbet1;
  norm2(&ssig1, &csig1);
  salp2 = cbet2 != cbet1 ? salp0 / cbet2 : salp1;
  calp2 = (cbet2 != cbet1 || fabs(sbet2) != -sbet1)
              ? sqrt(sq(calp1 * cbet1) + (cbet1 < -sbet1 ? (cbet2 - cbet1) * (cbet1 + cbet2)
                                                          : (sbet1 - sbet2) * (sbet1 + sbet2))) /
                    cbet2
              : fabs(calp1);
  ssig2 = sbet2;
  somg2 = salp0 * sbet2;
  csig2 = comg2 = calp2 * cbet2;
  norm2(&ssig2, &csig2);
  sig12 = atan2(fmax(0.0, csig1 * ssig2 - ssig1 * csig2) + 0, csig1 * csig2 + ssig1 * ssig2);
  somg12 = fmax(0.0, comg1 * somg2 - somg1 * comg2) + 0;
  comg12 = comg1 * comg2 + somg1 * somg2;
  eta = atan2(somg12 * clam120 - comg12 * slam120, comg12 * clam120 + somg12 * slam120);
  k2 = sq(calp0) * g->ep2;
  eps = k2 / (2 * (1 + sqrt(1 + k2)) + k2);
  C3f(g, eps, Ca);
  B312 = (SinSeries(ssig2, csig2, Ca, ORD - 1) - SinSeries(ssig1, csig1, Ca, ORD - 1));
  domg12 = -g->f * A3f(g, eps) * salp0 * (sig12 + B312);
  lam12 = eta + domg12;
  if (diffp) {
    if (calp2 == 0) {
      dlam12 = -2 * g->f1 * dn1 / sbet1;
    } else {
      Lengths(g, eps, sig12, ssig1, csig1, dn1, ssig2, csig2, dn2, 0, &dlam12, Ca);
      dlam12 *= g->f1 / (calp2 * cbet2);
    }
  }
  *psalp2 = salp2;
  *pcalp2 = calp2;
  *psig12 = sig12;
  *pssig1 = ssig1;
  *pcsig1 = csig1;
  *pssig2 = ssig2;
  *pcsig2 = csig2;
  *peps = eps;
  if (diffp) *pdlam12 = dlam12;
  return lam12;
}

/* ---------------------------------------------------------------- Inverse */
void orc_geod_inverse(double lat1, double lon1, double lat2, double lon2, double* ps12,
                      double* pazi1, double* pazi2) {
  const geod_t* g;
  double s12 = 0, lon12, lon12s, sbet1, cbet1, sbet2, cbet2, s12x = 0, m12x = 0;
  double dn1, dn2, lam12, slam12, clam12, sig12, calp1 = 0, salp1 = 0, calp2 = 0, salp2 = 0;
  double Ca[ORD + 1];
  int latsign, lonsign, swapp, meridian;
  geod_init();
  g = &G;
  g_last_numit = 0;

  lon12 = AngDiff(lon1, lon2, &lon12s);
  lonsign = signbit(lon12) ? -1 : 1;
  lon12 *= lonsign;
  lon12s *= lonsign;
  lam12 = lon12 * g->degree;
  sincosde(lon12, lon12s, &slam12, &clam12);
  lon12s = (HD - lon12) - lon12s;

  lat1 = AngRound(LatFix(lat1));
  lat2 = AngRound(LatFix(lat2));
  swapp = fabs(lat1) < fabs(lat2) || lat2 != lat2 ? -1 : 1;
  if (swapp < 0) {
    double t = lat1;
    lat1 = lat2;
    lat2 = t;
    lonsign *= -1;
  }
  latsign = signbit(lat1) ? 1 : -1;
  lat1 *= latsign;
  lat2 *= latsign;

  sincosdx(lat1, &sbet1, &cbet1);
  sbet1 *= g->f1;
  norm2(&sbet1, &cbet1);
  cbet1 = fmax(g->tiny, cbet1);
  sincosdx(lat2, &sbet2, &cbet2);
  sbet2 *= g->f1;
  norm2(&sbet2, &cbet2);
  cbet2 = fmax(g->tiny, cbet2);

  if (cbet1 < -sbet1) {
    if (cbet2 == cbet1) sbet2 = copysign(sbet1, sbet2);
  } else {
    if (fabs(sbet2) == -sbet1) cbet2 = cbet1;
  }

  dn1 = sqrt(1 + g->ep2 * sq(sbet1));
  dn2 = sqrt(1 + g->ep2 * sq(sbet2));

  meridian = lat1 == -QD || slam12 == 0;
  if (meridian) {
    double ssig1, csig1, ssig2, csig2;
    calp1 = clam12;
    salp1 = slam12;
    calp2 = 1;
    salp2 = 0;
    ssig1 = sbet1;
    csig1 = calp1 * cbet1;
    ssig2 = sbet2;
    csig2 = calp2 * cbet2;
    sig12 = atan2(fmax(0.0, csig1 * ssig2 - ssig1 * csig2) + 0, csig1 * csig2 + ssig1 * ssig2);
    Lengths(g, g->n, sig12, ssig1, csig1, dn1, ssig2, csig2, dn2, &s12x, &m12x, Ca);
    if (sig12 < 1 || m12x >= 0) {
      if (sig12 < 3 * g->tiny || (sig12 < g->tol0 && (s12x < 0 || m12x < 0)))
        sig12 = m12x = s12x = 0;
      m12x *= g->b;
      s12x *= g->b;
    } else {
      meridian = 0;
    }
  }

  if (!meridian && sbet1 == 0 && (g->f <= 0 || lon12s >= g->f * HD)) {
    calp1 = calp2 = 0;
    salp1 = salp2 = 1;
    s12x = g->a * lam12;
    sig12 = lam12 / g->f1;
    m12x = g->b * sin(sig12);
  } else if (!meridian) {
    double dnm = 0;
    sig12 = InverseStart(g, sbet1, cbet1, dn1, sbet2, cbet2, dn2, lam12, slam12, clam12, &salp1,
                         &calp1, &salp2, &calp2, &dnm);
    if (sig12 >= 0) {
      s12x = sig12 * g->b * dnm;
      m12x = sq(dnm) * g->b * sin(sig12 / dnm);
    } else {
      int numit = 0;
      double ssig1 = 0, csig1 = 0, ssig2 = 0, csig2 = 0, eps = 0;
      double salp1a = g->tiny, calp1a = 1, salp1b = g->tiny, calp1b = -1;
      int tripn = 0, tripb = 0;
      for (;; ++numit) {
        double dv = 0;
        double v = Lambda12(g, sbet1, cbet1, dn1, sbet2, cbet2, dn2, salp1, calp1, slam12,
                            clam12, &salp2, &calp2, &sig12, &ssig1, &csig1, &ssig2, &csig2,
                            &eps, numit < g->maxit1, &dv, Ca);
        if (tripb || !(fabs(v) >= (tripn ? 8 : 1) * g->tol0) || numit == g->maxit2) break;
        if (v > 0 && (numit > g->maxit1 || calp1 / salp1 > calp1b / salp1b)) {
          salp1b = salp1;
          calp1b = calp1;
        } else if (v < 0 && (numit > g->maxit1 || calp1 / salp1 < calp1a / salp1a)) {
          salp1a = salp1;
          calp1a = calp1;
        }
        if (numit < g->maxit1 && dv > 0) {
          double dalp1 = -v / dv;
          if (fabs(dalp1) < g->pi) {
            double sdalp1 = sin(dalp1), cdalp1 = cos(dalp1);
            double nsalp1 = salp1 * cdalp1 + calp1 * sdalp1;
            if (nsalp1 > 0) {
              calp1 = calp1 * cdalp1 - salp1 * sdalp1;
              salp1 = nsalp1;
              norm2(&salp1, &calp1);
              tripn = fabs(v) <= 16 * g->tol0;
              continue;
            }
          }
        }
        salp1 = (salp1a + salp1b) / 2;
        calp1 = (calp1a + calp1b) / 2;
        norm2(&salp1, &calp1);
        tripn = 0;
        tripb = (fabs(salp1a - salp1) + (calp1a - calp1) < g->tolb ||
                 fabs(salp1 - salp1b) + (calp1 - calp1b) < g->tolb);
      }
      g_last_numit = numit;
      Lengths(g, eps, sig12, ssig1, csig1, dn1, ssig2, csig2, dn2, &s12x, &m12x, Ca);
      m12x *= g->b;
      s12x *= g->b;
    }
  }

  s12 = 0 + s12x;
  if (swapp < 0) {
    double t = salp1;
    salp1 = salp2;
    salp2 = t;
    t = calp1;
    calp1 = calp2;
    calp2 = t;
  }
  salp1 *= swapp * lonsign;
  calp1 *= swapp * latsign;
  salp2 *= swapp * lonsign;
  calp2 *= swapp * latsign;
  if (ps12) *ps12 = s12;
  if (pazi1) *pazi1 = atan2dx(salp1, calp1);
  if (pazi2) *pazi2 = atan2dx(salp2, calp2);
}

int orc_geod_last_numit(void) { return g_last_numit; }

/* ---------------------------------------------------------------- reference wrappers */
/* warsim/utils/angles.py:10-15 */
static double normalize_angle(double a) {
  while (a >= 360.0) a -= 360;
  while (a < 0.0) a += 360;
  return a;
}

/* warsim/utils/geodesics.py:12-14 */
double orc_geodetic_distance_km(double lat1, double lon1, double lat2, double lon2) {
  double s12;
  orc_geod_inverse(lat1, lon1, lat2, lon2, &s12, 0, 0);
  return s12 / 1000.0;
}

/* warsim/utils/geodesics.py:17-19 */
double orc_geodetic_bearing_deg(double lat1, double lon1, double lat2, double lon2) {
  double azi1;
  orc_geod_inverse(lat1, lon1, lat2, lon2, 0, &azi1, 0);
  return normalize_angle(azi1);
}

/* warsim/utils/geodesics.py:22-24 */
void orc_geodetic_direct(double lat, double lon, double heading, double distance, double* lat2,
                         double* lon2) {
  orc_geod_direct(lat, lon, heading, distance, lat2, lon2, 0);
}
