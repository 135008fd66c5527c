// met_tables.hpp -- host-side preparation of the axis tables a MetView carries (reciprocal spacings, first-guess
// parameters for the interval searches).  Pure host code, shared by the engine and the test-only host emulation.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "physics.cuh"

namespace mpb {

struct AxisTables {
  std::vector<AxisCell> lonc, latc, pc;   // one record per axis interval
  std::vector<unsigned short> p_lut;
  unsigned p_lut_base = 0;
  int p_lut_shift = 0;
};

inline unsigned host_hi_word(double x) {
  unsigned long long b;
  std::memcpy(&b, &x, 8);
  return (unsigned)(b >> 32);
}

inline int host_bisect(const double *xx, int n, double x) {  // the reference bisection (src/mptrac.c:3495-3521)
  int lo = 0, hi = n - 1;
  const int m = (hi + lo) >> 1;
  if (xx[m] < xx[m + 1]) {
    while (hi > lo + 1) { const int i = (hi + lo) >> 1; if (xx[i] > x) hi = i; else lo = i; }
  } else {
    while (hi > lo + 1) { const int i = (hi + lo) >> 1; if (xx[i] <= x) hi = i; else lo = i; }
  }
  return lo;
}

inline AxisTables build_axis_tables(const double *lon, int nx, const double *lat, int ny, const double *p, int nz) {
  AxisTables t;
  auto cells = [](const double *x, int n) {
    std::vector<AxisCell> c((size_t)(n - 1));
    for (int i = 0; i < n - 1; i++) {
      c[i].lo = x[i]; c[i].hi = x[i + 1];
      c[i].d = x[i + 1] - x[i];
      c[i].rd = 1.0 / c[i].d;
    }
    return c;
  };
  t.lonc = cells(lon, nx); t.latc = cells(lat, ny); t.pc = cells(p, nz);
  // pressure first-guess table over the high word of the double (monotone in p for p > 0)
  double pmin = p[0], pmax = p[0];
  for (int i = 0; i < nz; i++) { pmin = std::min(pmin, p[i]); pmax = std::max(pmax, p[i]); }
  if (!(pmin > 0)) pmin = 1e-300;
  const unsigned h0 = host_hi_word(pmin), h1 = host_hi_word(std::max(pmax, pmin));
  int shift = 0;
  while (((h1 - h0) >> shift) + 1 > 2048) shift++;
  const int n = (int)((h1 - h0) >> shift) + 1;
  t.p_lut.resize(n);
  for (int k = 0; k < n; k++) {
    const unsigned long long bits = ((unsigned long long)(h0 + ((unsigned)k << shift))) << 32;
    double x;
    std::memcpy(&x, &bits, 8);
    t.p_lut[k] = (unsigned short)host_bisect(p, nz, x);
  }
  t.p_lut_base = h0;
  t.p_lut_shift = shift;
  return t;
}

// everything of a MetView that derives from the host axes (device / host pointers are set by the caller)
inline void fill_axis_scalars(MetView &g, const double *lon, int nx, const double *lat, int ny, const double *p, int nz,
                              int coord_type, double t0, double t1, const AxisTables &t) {
  g.nx = nx; g.ny = ny; g.nz = nz; g.coord_type = coord_type;
  g.t0 = t0; g.t1 = t1; g.dt01 = t1 - t0; g.r_dt01 = 1.0 / (t1 - t0);
  g.lon_first = lon[0]; g.lon_last = lon[nx - 1]; g.lon_d = lon[1] - lon[0]; g.r_lon_d = 1.0 / (lon[1] - lon[0]);
  g.lat_first = lat[0];
  g.lat_scale = (lat[ny - 1] != lat[0]) ? (ny - 1) / (lat[ny - 1] - lat[0]) : 0.0;
  g.lat_lo = *std::min_element(lat, lat + ny);
  g.lat_hi = *std::max_element(lat, lat + ny);
  g.lon_asc = lon[0] < lon[nx - 1];
  // direction test at the bisection midpoint, as the reference does (src/mptrac.c:3504)
  { const int m = (ny - 1) >> 1; g.lat_asc = lat[m] < lat[m + 1]; }
  { const int m = (nz - 1) >> 1; g.p_asc = p[m] < p[m + 1]; }
  g.local = std::fabs(lon[nx - 1] - lon[0] - 360.0) >= 0.01;   // src/mptrac.c:6011-6012
  g.p_lut_base = t.p_lut_base; g.p_lut_shift = t.p_lut_shift; g.p_lut_n = (int)t.p_lut.size();
}

}  // namespace mpb
