// Quadrature rules on the reference triangle / tetrahedron / segment.
//
// The reference takes these from MFEM's IntegrationRules (deps/MFEM/Common.cpp:42-43,
// deps/MFEM3/Common.cpp:24-25, deps/MFEM/Common.cpp:329-341); MFEM is not vendored, so the tables
// are restated here from the published rules.  Point ORDER matters: coefficient arrays are indexed
// e*g + k in this order.
#pragma once
#include <cmath>

namespace adfem {

constexpr int MAX_QP = 12;

struct QuadRule {
  int n;
  double x[MAX_QP], y[MAX_QP], z[MAX_QP], w[MAX_QP];
};

namespace detail {
inline void qp(QuadRule& r, double x, double y, double z, double w) {
  r.x[r.n] = x; r.y[r.n] = y; r.z[r.n] = z; r.w[r.n] = w; r.n++;
}
inline void tri3(QuadRule& r, double a, double w) {
  double b = 1. - 2. * a;
  qp(r, a, a, 0, w); qp(r, a, b, 0, w); qp(r, b, a, 0, w);
}
inline void tri6(QuadRule& r, double a, double b, double w) {
  double c = 1. - a - b;
  qp(r, a, b, 0, w); qp(r, b, a, 0, w); qp(r, a, c, 0, w); qp(r, c, a, 0, w); qp(r, b, c, 0, w); qp(r, c, b, 0, w);
}
inline void tet4(QuadRule& r, double a, double b, double w) {
  qp(r, a, a, a, w); qp(r, a, a, b, w); qp(r, a, b, a, w); qp(r, b, a, a, w);
}
inline void tet6(QuadRule& r, double a, double w) {
  double b = 0.5 - a;
  qp(r, a, a, b, w); qp(r, a, b, a, w); qp(r, b, a, a, w); qp(r, a, b, b, w); qp(r, b, a, b, w); qp(r, b, b, a, w);
}
}  // namespace detail

// returns false for an unsupported order
inline bool triangle_rule(int order, QuadRule& r) {
  using namespace detail;
  r.n = 0;
  switch (order) {
    case 0: case 1: qp(r, 1. / 3., 1. / 3., 0, 0.5); return true;
    case 2: tri3(r, 1. / 6., 1. / 6.); return true;
    case 3: qp(r, 1. / 3., 1. / 3., 0, -0.28125); tri3(r, 0.2, 25. / 96.); return true;
    case 4:
      tri3(r, 0.091576213509770743460, 0.054975871827660933819);
      tri3(r, 0.44594849091596488632, 0.11169079483900573285);
      return true;
    case 5:
      qp(r, 1. / 3., 1. / 3., 0, 0.1125);
      tri3(r, 0.10128650732345633880, 0.062969590272413576298);
      tri3(r, 0.47014206410511508977, 0.066197076394253090369);
      return true;
    case 6:
      tri3(r, 0.063089014491502228340, 0.025422453185103408460);
      tri3(r, 0.24928674517091042129, 0.058393137863189683013);
      tri6(r, 0.053145049844816947353, 0.31035245103378440542, 0.041425537809186787597);
      return true;
    default: return false;
  }
}

inline bool tetrahedron_rule(int order, QuadRule& r) {
  using namespace detail;
  r.n = 0;
  switch (order) {
    case 0: case 1: qp(r, 0.25, 0.25, 0.25, 1. / 6.); return true;
    case 2: { double b = 0.58541019662496845446; tet4(r, (1. - b) / 3., b, 1. / 24.); return true; }
    case 3: qp(r, 0.25, 0.25, 0.25, -2. / 15.); tet4(r, (1. - 0.5) / 3., 0.5, 0.075); return true;
    case 4:
      tet4(r, 1. / 14., 1. - 3. * (1. / 14.), 343. / 45000.);
      qp(r, 0.25, 0.25, 0.25, -74. / 5625.);
      tet6(r, 0.10059642383320079500, 28. / 1125.);
      return true;
    default: return false;
  }
}

// Gauss-Legendre on [0,1], n = (order|1)/2 + 1 points in ascending order.
inline int segment_rule(int order, double* p, double* w) {
  int n = (order | 1) / 2 + 1;
  for (int i = 1; i <= (n + 1) / 2; i++) {
    double z = std::cos(M_PI * (i - 0.25) / (n + 0.5)), pp = 1, p1 = 1;
    for (int it = 0; it < 100; it++) {
      p1 = 1.0; double p2 = 0.0;
      for (int j = 1; j <= n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0);
      double dz = p1 / pp; z -= dz;
      if (std::fabs(dz) < 1e-16) break;
    }
    { p1 = 1.0; double p2 = 0.0;
      for (int j = 1; j <= n; j++) { double p3 = p2; p2 = p1; p1 = ((2.0 * j - 1.0) * z * p2 - (j - 1.0) * p3) / j; }
      pp = n * (z * p1 - p2) / (z * z - 1.0); }
    double xx = 0.5 * (1.0 - z), ww = 1.0 / ((1.0 - z * z) * pp * pp);
    p[i - 1] = xx; w[i - 1] = ww; p[n - i] = 1.0 - xx; w[n - i] = ww;
  }
  return n;
}

}  // namespace adfem
