/*
 * oracle/geodesic.h -- TEST INFRASTRUCTURE (CPU oracle). Not part of the product path.
 *
 * Restates the WGS84 geodesic arithmetic of the third-party package
 * `geographiclib==2.0` (pinned in the reference's README.md:22; NOT vendored under
 * /root/reference and not installable here), i.e. C. F. F. Karney, "Algorithms for
 * geodesics", J. Geodesy 87 (2013) 43-55, series order 6.  The reference reaches it at
 * warsim/utils/geodesics.py:12-24 (Geodesic.WGS84.Inverse / Direct).
 *
 * Parity status at this boundary: UNPINNED by the reference (it has no tests); pinned
 * by us against geographiclib's documented known answers and 40-digit mpmath
 * quadrature of the exact elliptic integrals (tests/test_oracle_geodesic.py).
 */
#ifndef HH_ORACLE_GEODESIC_H
#define HH_ORACLE_GEODESIC_H

#ifdef __cplusplus
extern "C" {
#endif

/* Direct problem: start (lat1, lon1) [deg], azimuth azi1 [deg], distance s12 [m]. */
void orc_geod_direct(double lat1, double lon1, double azi1, double s12,
                     double* lat2, double* lon2, double* azi2);

/* Inverse problem: distance s12 [m] and forward azimuths [deg, (-180,180]]. */
void orc_geod_inverse(double lat1, double lon1, double lat2, double lon2,
                      double* s12, double* azi1, double* azi2);

/* The three wrappers of warsim/utils/geodesics.py:12-24. */
double orc_geodetic_distance_km(double lat1, double lon1, double lat2, double lon2);
double orc_geodetic_bearing_deg(double lat1, double lon1, double lat2, double lon2);
void orc_geodetic_direct(double lat, double lon, double heading, double distance,
                         double* lat2, double* lon2);

/* Number of Newton iterations used by the last orc_geod_inverse call (diagnostics). */
int orc_geod_last_numit(void);

#ifdef __cplusplus
}
#endif
#endif
