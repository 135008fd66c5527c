// physics.cuh -- per-parcel physics of the MPTRAC time-step path, written for one CUDA thread
// per parcel.  Everything here is __host__ __device__ so that tests/_hostemu can execute the very
// same source on the CPU of the (GPU-less) development container; the shipped library only ever
// calls it from kernels (see step_kernels.cu).
//
// Behavioural contract = the reference's module_* functions; the citations below are
// file:line in slcs-jsc/mptrac (src/mptrac.c unless stated otherwise).  Precision rules that are
// part of the contract: parcel state is fp64, met fields are fp32 and the FIRST subtraction of a
// lerp happens in fp32 (3023-3043), mesoscale statistics accumulate in fp32 in a fixed order
// (4288-4311), Box-Muller takes cosf/sinf of a float angle (5821-5826).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>
#include <cuda_runtime.h>

#define MPB_HD __host__ __device__ __forceinline__
#ifndef MPB_LEVEL_CACHE
#define MPB_LEVEL_CACHE 1   // model-level advection: keep the 8 level records across Runge-Kutta stages (see locate_on_levels);
#endif                      // measured on B200 (c2ml): 0.371 ms per step with, 0.425 ms without
// rarely executed paths are kept out of line so that they cost neither registers nor instruction-cache lines on the hot path
#define MPB_COLD __host__ __device__ __noinline__

namespace mpb {

// ----------------------------------------------------------------------------------------------
// constants (src/mptrac.h:265-340)
// ----------------------------------------------------------------------------------------------
constexpr double kPi = 3.14159265358979323846;
constexpr double kG0 = 9.80665;
constexpr double kH0 = 7.0;
constexpr double kKB = 1.3806504e-23;
constexpr double kMA = 28.9644;
constexpr double kP0 = 1013.25;
constexpr double kRI = 8.3144598;
constexpr double kRA = 1e3 * kRI / kMA;
constexpr double kRE = 6367.421;
constexpr double kMAirMolecule = 4.8096e-26;
constexpr double kYear = 365.25 * 86400.;

// module bits of one fused launch
enum : unsigned {
  MOD_TIMESTEPS = 1u << 0,  // compute dt in-kernel (otherwise read cache dt from memory)
  MOD_POS_PRE = 1u << 1,
  MOD_ADVECT = 1u << 2,
  MOD_TURB = 1u << 3,
  MOD_MESO = 1u << 4,
  MOD_SEDI = 1u << 5,
  MOD_POS_POST = 1u << 6,
  MOD_STORE_DT = 1u << 7,  // also write dt back (cache_t::dt for callers that need it)
};

// ----------------------------------------------------------------------------------------------
// device views
// ----------------------------------------------------------------------------------------------
// One met node holds BOTH bracketing time levels of u, v, w (omega) and T: a 32-byte, 32-byte-aligned record that one
// 256-bit load (LDG.E.256 on sm_100a) fetches as a single sector.  The two time levels of a field sit next to each other,
// so that a thread which follows ONE field (the lane-per-component step kernel, quad.cuh) reads it with one 8-byte load.
struct alignas(32) Node {
  float u0, u1;  // u at met0, met1
  float v0, v1;
  float w0, w1;
  float t0, t1;
};

// One interval of a grid axis: both end points, their difference and its correctly rounded reciprocal in one
// 32-byte record, so that index verification and the interpolation weight need a single 256-bit load.
struct alignas(32) AxisCell {
  double lo, hi;  // x[i], x[i+1]
  double d, rd;   // x[i+1] - x[i] and RN(1 / d): divisions by a grid constant become multiply + 2 FMA (Markstein)
};

struct alignas(32) LevelNode {
  float h0, a0, b0, c0;   // time level 0: search coordinate, three fields
  float h1, a1, b1, c1;   // time level 1
};

struct MetView {
  const Node *f;          // [nx][ny][nz], z fastest
  const float4 *s;        // [nx][ny] {ps0, pbl0, ps1, pbl1}
  const double *lon, *lat, *p;        // axes (slow paths, ptop)
  const AxisCell *lonc, *latc, *pc;   // [n-1] intervals of each axis
  const unsigned short *p_lut;        // first guess of the pressure interval from the high word of p
  unsigned p_lut_base;                // high 32 bits of the smallest tabulated pressure
  int p_lut_shift, p_lut_n;
  double t0, t1, dt01, r_dt01;        // met0->time, met1->time, t1 - t0 and its reciprocal
  double lon_first, lon_last, lon_d, r_lon_d;   // lon[0], lon[nx-1], lon[1]-lon[0] and its reciprocal
  double lat_first, lat_scale;        // first guess of the latitude interval: (lat - lat_first) * lat_scale
  double lat_lo, lat_hi;              // min / max of the latitude axis
  int nx, ny, nz;
  int coord_type;         // 0 lon/lat, 1 Cartesian
  int lon_asc, lat_asc, p_asc;
  int local;              // met domain is not global (5999-6012)
  // model levels (ADVECT_VERT_COORD 1, 2, 3), null when the met data has none.  A LevelNode holds the search coordinate
  // h and three fields of both time levels: lp {pl, ul, vl, wl} (omega on model levels), lz {zetal, ul, vl, zeta_dotl}
  // (zeta / eta); pz {pl0, zetal0, pl1, zetal1} serves the two conversions pressure <-> zeta
  const struct LevelNode *lp, *lz;
  const float4 *pz;
  int npl;
};

struct ClimView {
  const double *time;   // [ntime]
  const double *lat;    // [nlat]
  const double *tropo;  // [ntime][nlat]
  int ntime, nlat;
};

struct CtlView {
  double t, t_start, t_stop, dt_met;
  double utm_ref_lat;
  double dx_pbl, dx_trop, dx_strat, dz_pbl, dz_trop, dz_strat;
  double mesox, mesoz, pbl_trans;
  uint64_t ctr_turb, ctr_meso;  // Squares counters at the start of the two module_rng calls
  int direction;
  int pbl_scheme;
};

struct Parcel {
  double time, lon, lat, p;
};

// ----------------------------------------------------------------------------------------------
// small helpers
// ----------------------------------------------------------------------------------------------
template <typename T>
MPB_HD T ldg(const T *ptr) {
#ifdef __CUDA_ARCH__
  return __ldg(ptr);
#else
  return *ptr;
#endif
}

// fp32 arithmetic with exactly one rounding per operation (no FMA contraction)
MPB_HD float f_add(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fadd_rn(a, b);
#else
  return a + b;
#endif
}
MPB_HD float f_sub(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fsub_rn(a, b);
#else
  return a - b;
#endif
}
MPB_HD float f_mul(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}

// x / d for a divisor d whose correctly rounded reciprocal rd = RN(1/d) is known: one multiply and two FMAs
// (Markstein's correction step) instead of the ~30-instruction IEEE division sequence; the result is the
// correctly rounded quotient for finite, normal operands, i.e. the same bits `x / d` gives on the CPU.
MPB_HD double div_by(double x, double d, double rd) {
  const double q = x * rd;
  const double r = fma(-q, d, x);
  return fma(r, rd, q);
}

// Where a last-ulp difference is harmless -- interpolation weights, metre -> degree conversions, i.e. everything that
// enters a position continuously -- the production device build multiplies by the reciprocal and skips the correction
// (one fp64 op instead of a dependent chain of three; the RK stage is a latency-bound dependent chain).  The host build
// (tests/_hostemu) and the strict device build (-DMPB_STRICT) keep the exact sequence; indices that decide a cell for
// sorting or for the mesoscale statistics are always exact.  Measured deviations: see DESIGN.md.
#if defined(__CUDA_ARCH__) && !defined(MPB_STRICT)
#define MPB_FAST_QUOT 1
#else
#define MPB_FAST_QUOT 0
#endif
MPB_HD double quot(double x, double d, double rd) {
#if MPB_FAST_QUOT
  (void)d;
  return x * rd;
#else
  return div_by(x, d, rd);
#endif
}

// a / b for a divisor that is not a grid constant, again only where a last-ulp difference is harmless (climatological
// weights, Langevin coefficients, the settling velocity): a hardware reciprocal estimate refined by two Newton steps --
// 6 instructions and <= 1 ulp off, instead of the ~30-instruction IEEE sequence with its out-of-line slow path.
MPB_HD double fdiv(double a, double b) {
#if MPB_FAST_QUOT
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = fma(fma(-b, r, 1.0), r, r);
  r = fma(fma(-b, r, 1.0), r, r);
  return a * r;
#else
  return a / b;
#endif
}

// The quotient for an INDEX (a cell of a regular axis, the truncated quotient of FMOD).  div_by yields the correctly rounded
// quotient for all but rare operand pairs (division by a constant through its rounded reciprocal: Markstein's condition);
// a last-bit error only matters when the quotient sits within a few ulps of an integer, where it could move a parcel into the
// neighbouring cell.  There -- and only there -- a true division decides, so the index is the reference's for every input.
static MPB_COLD double div_true(double x, double d) { return x / d; }
MPB_HD double div_for_index(double x, double d, double rd) {
  const double q = div_by(x, d, rd);
  const double r = rint(q);
  if (fabs(q - r) <= 1e-15 * fabs(r)) return div_true(x, d);
  return q;
}

constexpr double kR360 = 1.0 / 360.0;
constexpr double kR1000 = 1.0 / 1000.0;
constexpr double kRH0 = 1.0 / kH0;
constexpr double kPiRE = kPi * kRE;
constexpr double kRPiRE = 1.0 / (kPi * kRE);

// x - trunc(x / 360) * 360 with the quotient truncated through int (FMOD, src/mptrac.h:1121-1122)
MPB_HD double mod360(double x) { return x - (int)div_for_index(x, 360., kR360) * 360.; }
// general form (cold paths)
MPB_HD double mod_trunc(double x, double y) { return x - (int)(x / y) * y; }

// y0 + (y1 - y0) / (x1 - x0) * (x - x0)  (src/mptrac.h:1351)
MPB_HD double lin(double x0, double y0, double x1, double y1, double x) {
  return y0 + fdiv(y1 - y0, x1 - x0) * (x - x0);
}

// log-pressure altitude [km], Z(p) (src/mptrac.h:2243)
MPB_HD double altitude(double p) { return kH0 * log(kP0 / p); }

// km -> hPa at pressure p (src/mptrac.h:941)
MPB_HD double dz2dp(double dz, double p) { return quot(-dz * p, kH0, kRH0); }

// metres east / north -> coordinate increment (src/mptrac.h:904-906, 922-923, 966, 989).
// DX2DEG(dx, lat) = dx * 180 / (pi * RE * cos(lat * pi / 180)), 0 within 0.001 deg of a pole.  The divisor only depends on
// the latitude, and every Runge-Kutta stage of a step uses the same one (3659-3673): LonScale holds it together with
// its correctly rounded reciprocal, so each further use costs a multiply and two FMAs instead of a cosine and a division
// and still yields the correctly rounded quotient.
struct LonScale {
  double d, rd;
  int mode;  // 0 = Cartesian (identity), 1 = polar cap (zero), 2 = divide by d
};
constexpr double kDegPerM = 180.0 / (1000.0 * kPi * kRE);   // fast path of DY2DEG(dy / 1000)
// cos(x) for |x| <= pi/2 (a latitude in radians): no argument reduction, no out-of-line slow path.  Polynomial kernels
// of the classic fdlibm k_cos / k_sin scheme (Sun Microsystems' published minimax coefficients, < 1 ulp on
// [-pi/4, pi/4]); beyond pi/4 the identity cos x = sin(pi/2 - |x|) with pi/2 split into two doubles.  Production device
// build only: the host and the strict build call the library cosine.
MPB_HD double cos_quarter(double x) {
#if MPB_FAST_QUOT
  const double ax = fabs(x);
  if (ax <= 0.78539816339744830962) {
    const double z = ax * ax;
    const double r = z * (4.16666666666666019037e-02 + z * (-1.38888888888741095749e-03 + z * (2.48015872894767294178e-05 +
                     z * (-2.75573143513906633035e-07 + z * (2.08757232129817482790e-09 + z * -1.13596475577881948265e-11)))));
    const double hz = 0.5 * z, w = 1.0 - hz;
    return w + (((1.0 - w) - hz) + z * r);
  }
  // y = pi/2 - |x| to twice the working precision, then sin(y + t) with the tail t folded into the kernel
  const double pio2_hi = 1.57079632679489655800e+00, pio2_lo = 6.12323399573676603587e-17;
  const double y = pio2_hi - ax, t = ((pio2_hi - y) - ax) + pio2_lo;
  const double z = y * y, v = z * y;
  const double r = 8.33333333332248946124e-03 + z * (-1.98412698298579493134e-04 + z * (2.75573137070700676789e-06 +
                   z * (-2.50507602534068634195e-08 + z * 1.58969099521155010221e-10)));
  return y - ((z * (0.5 * t - v * r) - t) - v * -1.66666666666666324348e-01);
#else
  return cos(x);
#endif
}

MPB_HD LonScale lon_scale(int coord_type, double lat) {
  LonScale k;
  k.d = 1.0; k.rd = 1.0;
  if (coord_type != 0) { k.mode = 0; return k; }
  if (lat < -89.999 || lat > 89.999) { k.mode = 1; return k; }
  k.mode = 2;
  k.d = kPiRE * cos_quarter(lat * (kPi / 180.0));   // |lat| <= 89.999 here
#if MPB_FAST_QUOT && !defined(MPB_LONSCALE_IEEE_RCP)
  k.rd = fdiv(0.18, k.d);   // metres -> degrees in one multiply: dx / 1000 * 180 / d (reciprocal estimate + two Newton steps)
#elif MPB_FAST_QUOT
  k.rd = 1.0 / k.d * 0.18;
#else
  k.rd = 1.0 / k.d;
#endif
  return k;
}
MPB_HD double dx2coord(const LonScale &k, double dx) {
  if (k.mode == 0) return dx;
  if (k.mode == 1) return 0.0;
#if MPB_FAST_QUOT
  return dx * k.rd;
#else
  return div_by(div_by(dx, 1000.0, kR1000) * 180., k.d, k.rd);
#endif
}
MPB_HD double dx2coord(int coord_type, double dx, double lat) { return dx2coord(lon_scale(coord_type, lat), dx); }
// the same conversions in the reference's own operation order (model-level advection: not tuned)
MPB_HD double dx2coord_exact(int coord_type, double dx, double lat) {
  if (coord_type != 0) return dx;
  if (lat < -89.999 || lat > 89.999) return 0.0;
  return dx / 1000. * 180. / (kPi * kRE * cos(lat * (kPi / 180.0)));
}
MPB_HD double dy2coord_exact(int coord_type, double dy) { return coord_type != 0 ? dy : dy / 1000. * 180. / (kPi * kRE); }
MPB_HD double dy2coord(int coord_type, double dy) {
  if (coord_type != 0) return dy;
#if MPB_FAST_QUOT
  return dy * kDegPerM;
#else
  return div_by(div_by(dy, 1000.0, kR1000) * 180., kPiRE, kRPiRE);
#endif
}

// Interval search on a monotone axis; same result as the reference bisection (3495-3521):
// ascending: largest i <= n-2 with xx[i] <= x (0 if none); descending: largest i with xx[i] > x.
MPB_HD int find_interval(const double *xx, int n, int ascending, double x) {
  int lo = 0, hi = n - 1;
  if (ascending) {
    while (hi > lo + 1) {
      const int mid = (hi + lo) >> 1;
      if (ldg(xx + mid) > x) hi = mid; else lo = mid;
    }
  } else {
    while (hi > lo + 1) {
      const int mid = (hi + lo) >> 1;
      if (ldg(xx + mid) <= x) hi = mid; else lo = mid;
    }
  }
  return lo;
}

// Same result as find_interval, starting from a first guess i (any value): walk to the interval.  For a
// monotone axis the answer of the reference bisection is unique, so the guess only affects the cost.
static MPB_COLD int refine_interval(const double *xx, int n, int ascending, double x, int i) {
  i = i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
  if (ascending) {
    while (i > 0 && ldg(xx + i) > x) i--;
    while (i < n - 2 && ldg(xx + i + 1) <= x) i++;
  } else {
    while (i > 0 && ldg(xx + i) <= x) i--;
    while (i < n - 2 && ldg(xx + i + 1) > x) i++;
  }
  return i;
}

MPB_HD unsigned hi_word(double x) {
#ifdef __CUDA_ARCH__
  return (unsigned)__double2hiint(x);
#else
  unsigned long long b;
  memcpy(&b, &x, 8);
  return (unsigned)(b >> 32);
#endif
}

// Regular-axis index by division, clamped to [0, n-2] (3559-3574)
MPB_HD int find_regular(double x0, double dx, double rdx, int n, double x) {
  const int i = (int)div_for_index(x - x0, dx, rdx);
  return i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
}
// the same with a true division (cold paths: climatology axis)
MPB_HD int find_regular_div(double x0, double dx, int n, double x) {
  const int i = (int)fdiv(x - x0, dx);   // (feeds a continuous interpolation: a flipped index at a node changes nothing)
  return i < 0 ? 0 : (i > n - 2 ? n - 2 : i);
}

// |x| < 360 from the high word alone (360.0 = 0x4076800000000000; its low word is zero)
MPB_HD bool below_360(double x) { return (hi_word(x) & 0x7fffffffu) < 0x40768000u; }

// FMOD(x, 360): for |x| < 360 the truncated quotient is 0 and x - 0 * 360 == x bit for bit, so the division is
// only executed on the (rare) other side
static MPB_COLD double mod360_cold(double x) { return mod360(x); }
MPB_HD double wrap360(double x) { return below_360(x) ? x : mod360_cold(x); }

// the reference's macros, as ternaries in their argument order (what they do with a NaN is part of the contract)
MPB_HD double max_of(double a, double b) { return a > b ? a : b; }    // MAX, src/mptrac.h:1378
MPB_HD double min_of(double a, double b) { return a < b ? a : b; }    // MIN :1479
MPB_HD double clamp_of(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }   // CLAMP :756
MPB_HD double clamp_to(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }  // MIN(MAX(x, lo), hi)

// horizontal range check before a lookup (2755-2803)
MPB_HD void clamp_horizontal(const MetView &g, double lon, double lat, double &lon2, double &lat2) {
  if (g.coord_type == 0) {
    lon2 = wrap360(lon);
    if (lon2 < g.lon_first) lon2 += 360;
    else if (lon2 > g.lon_last) lon2 -= 360;
    lat2 = clamp_to(lat, g.lat_lo, g.lat_hi);
  } else {
    const double xlo = g.lon_asc ? g.lon_first : g.lon_last;
    const double xhi = g.lon_asc ? g.lon_last : g.lon_first;
    lon2 = clamp_to(lon, xlo, xhi);
    lat2 = clamp_to(lat, g.lat_lo, g.lat_hi);
  }
}

// ----------------------------------------------------------------------------------------------
// met lookups
// ----------------------------------------------------------------------------------------------
MPB_HD AxisCell load_cell(const AxisCell *ptr) {
#ifdef __CUDA_ARCH__
  AxisCell c;
  asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(c.lo), "=d"(c.hi), "=d"(c.d), "=d"(c.rd) : "l"(ptr));
  return c;
#else
  return *ptr;
#endif
}

struct Stencil {
  int ix, iy, iz;
  double wx, wy, wz;  // weight of the LOWER index node along each axis (3015-3020)
};

// Is [c.lo, c.hi] the interval i the reference bisection (3495-3521) returns for x?  The answer of that bisection on a
// monotone axis is: ascending -- the largest i <= n-2 with xx[i] <= x (0 if none); descending -- the largest
// i <= n-2 with xx[i] > x (0 if none).
MPB_HD bool cell_holds(const AxisCell &c, int i, int n, int ascending, double x) {
  const bool asc = ascending != 0, lo_gt = c.lo > x, hi_gt = c.hi > x;
  const bool below_lo = (lo_gt == asc);    // ascending: lo > x, descending: lo <= x -- the answer lies at a smaller index
  const bool above_hi = (hi_gt != asc);    // ascending: hi <= x, descending: hi > x -- ... at a larger index
  return !(below_lo && i > 0) && !(above_hi && i < n - 2);
}

// Interval + cell of an irregular axis from a first guess: one 256-bit load when the guess is right (it nearly always
// is), one more when the answer is the neighbouring interval (a first-guess table bin that straddles a grid level), the
// exact search otherwise.
MPB_HD int locate_cell(const double *xx, const AxisCell *cells, int n, int ascending, double x, int guess, AxisCell &c) {
  int i = guess < 0 ? 0 : (guess > n - 2 ? n - 2 : guess);
  c = load_cell(cells + i);
  if (!cell_holds(c, i, n, ascending, x)) {
    const bool down = ((c.lo > x) == (ascending != 0));   // the answer lies at a smaller index
    const int j = down ? (i > 0 ? i - 1 : 0) : (i < n - 2 ? i + 1 : n - 2);
    c = load_cell(cells + j);
    i = j;
    if (!cell_holds(c, i, n, ascending, x)) {
      i = refine_interval(xx, n, ascending, x, i);
      c = load_cell(cells + i);
    }
  }
  return i;
}

MPB_HD int lat_guess(const MetView &g, double lat) { return (int)((lat - g.lat_first) * g.lat_scale); }

MPB_HD int p_guess(const MetView &g, double p) {
  const unsigned h = hi_word(p);
  int k = h > g.p_lut_base ? (int)((h - g.p_lut_base) >> g.p_lut_shift) : 0;
  k = k < g.p_lut_n ? k : g.p_lut_n - 1;
  return (int)ldg(g.p_lut + k);
}

MPB_HD int lon_interval(const MetView &g, double lon) { return find_regular(g.lon_first, g.lon_d, g.r_lon_d, g.nx, lon); }

// The three axis intervals of the cell a parcel was last looked up in.  Successive lookups of one parcel (Runge-Kutta
// stages, diffusion, sedimentation) nearly always stay inside them, so the interval search degenerates to two compares
// against registers: no first-guess table, no interval record load, no dependent-load latency.
struct CellAxes {
  AxisCell x, y, z;
  int ix, iy, iz;   // -1 = nothing yet
};
MPB_HD void axes_reset(CellAxes &a) { a.ix = -1; a.iy = -1; a.iz = -1; }

// (`exact` = the index decides something discontinuous -- the mesoscale statistics; an interpolation lookup is continuous
// across a cell face, so its index may come from the uncorrected quotient)
MPB_HD int lon_cell(const MetView &g, double lon, CellAxes &a, bool exact = false) {
#if MPB_FAST_QUOT
  int ix;
  if (exact) ix = lon_interval(g, lon);
  else { ix = (int)((lon - g.lon_first) * g.r_lon_d); ix = ix < 0 ? 0 : (ix > g.nx - 2 ? g.nx - 2 : ix); }
#else
  (void)exact;
  const int ix = lon_interval(g, lon);
#endif
  if (ix != a.ix) { a.x = load_cell(g.lonc + ix); a.ix = ix; }
  return ix;
}
MPB_HD int lat_cell(const MetView &g, double lat, CellAxes &a) {
  if (a.iy < 0 || !cell_holds(a.y, a.iy, g.ny, g.lat_asc, lat))
    a.iy = locate_cell(g.lat, g.latc, g.ny, g.lat_asc, lat, lat_guess(g, lat), a.y);
  return a.iy;
}
MPB_HD int p_cell(const MetView &g, double p, CellAxes &a) {
  if (a.iz < 0 || !cell_holds(a.z, a.iz, g.nz, g.p_asc, p))
    a.iz = locate_cell(g.p, g.pc, g.nz, g.p_asc, p, p_guess(g, p), a.z);
  return a.iz;
}

// uncached searches (sort keys)
MPB_HD int lat_interval(const MetView &g, double lat) {
  AxisCell c;
  return locate_cell(g.lat, g.latc, g.ny, g.lat_asc, lat, lat_guess(g, lat), c);
}
MPB_HD int p_interval(const MetView &g, double p) {
  AxisCell c;
  return locate_cell(g.p, g.pc, g.nz, g.p_asc, p, p_guess(g, p), c);
}

MPB_HD void stencil_2d(const MetView &g, double lon, double lat, CellAxes &a, Stencil &s) {
  double lon2, lat2;
  clamp_horizontal(g, lon, lat, lon2, lat2);
  s.ix = lon_cell(g, lon2, a);
  s.iy = lat_cell(g, lat2, a);
  s.wx = quot(a.x.hi - lon2, a.x.hi - a.x.lo, a.x.rd);
  s.wy = quot(a.y.hi - lat2, a.y.hi - a.y.lo, a.y.rd);
}

// (Measured alternative: promoting on the integer pipes -- IMAD.HI + shifts + LOP3, exact for normals and zero -- instead of
// F2F.F64.F32, which runs on the 16-lane XU pipe: 158 us vs 117 us per step; the extra issue slots cost more than XU.)
MPB_HD double lerp_f64(double w, double lo, double hi) { return w * (lo - hi) + hi; }

MPB_HD Node load_node(const Node *ptr) {
#ifdef __CUDA_ARCH__
  Node n;
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(n.u0), "=f"(n.u1), "=f"(n.v0), "=f"(n.v1), "=f"(n.w0), "=f"(n.w1), "=f"(n.t0), "=f"(n.t1)
      : "l"(ptr));
  return n;
#else
  return *ptr;
#endif
}

// The 8 corner nodes (both time levels each) of one grid cell.  A Cube outlives a single lookup: the Runge-Kutta stages
// of a step, the mesoscale statistics and the sedimentation lookup nearly always fall into the SAME cell (a stage moves
// a parcel by a few km, a cell is ~100 km x 1 km), so the cube is fetched once per step and only re-fetched by the
// threads whose cell changed.  That takes the scattered 32-byte gathers -- the kernel's first limiter, the L1 data pipe
// -- from 4-6 per parcel-step to little more than one.
// DIFF = true: the lower-level nodes are replaced, once per fetch, by the fp32 differences "lower - upper" every vertical
// lerp starts with (3023-3038), which takes those subtractions off the per-stage path; the mesoscale statistics need the
// raw corner values, so kernels that include them use DIFF = false.
// A window of the met grid staged in shared memory (the TMA-tile form of the step kernel, engine.cu tile_step_kernel): nodes
// [x0, x0 + nx) x [y0, y0 + ny) x [z0, z0 + nz), z fastest, as the bulk tensor copy lays them down.
struct TileRef {
  const Node *nodes;
  int x0, y0, z0, nx, ny, nz;
};
template <bool DIFF>
struct CubeT {
  Node n000, n001, n010, n011, n100, n101, n110, n111;  // index order: x, y, z
  int ix, iy, iz;                                        // the cell held; ix < 0 = nothing yet
  CellAxes ax;                                           // its axis intervals
  const TileRef *tile;                                   // cells inside this window are fetched from shared memory (or null)
};
using Cube = CubeT<false>;
template <bool DIFF>
MPB_HD void cube_reset(CubeT<DIFF> &c) { c.ix = -1; c.iy = -1; c.iz = -1; axes_reset(c.ax); c.tile = nullptr; }

MPB_HD void node_diff(Node &lo, const Node &hi) {
  lo.u0 = f_sub(lo.u0, hi.u0); lo.v0 = f_sub(lo.v0, hi.v0); lo.w0 = f_sub(lo.w0, hi.w0); lo.t0 = f_sub(lo.t0, hi.t0);
  lo.u1 = f_sub(lo.u1, hi.u1); lo.v1 = f_sub(lo.v1, hi.v1); lo.w1 = f_sub(lo.w1, hi.w1); lo.t1 = f_sub(lo.t1, hi.t1);
}

MPB_HD Node load_node_staged(const Node *ptr) {   // a node of the shared-memory window: two 16-byte loads (LDS.128)
  Node n;
#ifdef __CUDA_ARCH__
  const unsigned addr = (unsigned)__cvta_generic_to_shared(ptr);
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(n.u0), "=f"(n.u1), "=f"(n.v0), "=f"(n.v1) : "r"(addr));
  asm("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4+16];" : "=f"(n.w0), "=f"(n.w1), "=f"(n.t0), "=f"(n.t1) : "r"(addr));
#else
  n = *ptr;
#endif
  return n;
}

template <bool DIFF>
MPB_HD void load_cube(const MetView &g, const Stencil &s, CubeT<DIFF> &c) {
  if (c.tile) {
    const TileRef &t = *c.tile;
    const int lx = s.ix - t.x0, ly = s.iy - t.y0, lz = s.iz - t.z0;
    if (lx >= 0 && lx < t.nx - 1 && ly >= 0 && ly < t.ny - 1 && lz >= 0 && lz < t.nz - 1) {
      const int sy = t.nz, sx = t.ny * t.nz;
      const Node *b = t.nodes + (lx * sx + ly * sy + lz);
      c.n000 = load_node_staged(b);
      c.n001 = load_node_staged(b + 1);
      c.n010 = load_node_staged(b + sy);
      c.n011 = load_node_staged(b + sy + 1);
      c.n100 = load_node_staged(b + sx);
      c.n101 = load_node_staged(b + sx + 1);
      c.n110 = load_node_staged(b + sx + sy);
      c.n111 = load_node_staged(b + sx + sy + 1);
      if (DIFF) { node_diff(c.n000, c.n001); node_diff(c.n010, c.n011); node_diff(c.n100, c.n101); node_diff(c.n110, c.n111); }
      c.ix = s.ix; c.iy = s.iy; c.iz = s.iz;
      return;
    }
  }
  const size_t sy = (size_t)g.nz, sx = (size_t)g.ny * (size_t)g.nz;
  const Node *b = g.f + ((size_t)s.ix * sx + (size_t)s.iy * sy + (size_t)s.iz);
  c.n000 = load_node(b);
  c.n001 = load_node(b + 1);
  c.n010 = load_node(b + sy);
  c.n011 = load_node(b + sy + 1);
  c.n100 = load_node(b + sx);
  c.n101 = load_node(b + sx + 1);
  c.n110 = load_node(b + sx + sy);
  c.n111 = load_node(b + sx + sy + 1);
  if (DIFF) { node_diff(c.n000, c.n001); node_diff(c.n010, c.n011); node_diff(c.n100, c.n101); node_diff(c.n110, c.n111); }
  c.ix = s.ix; c.iy = s.iy; c.iz = s.iz;
}

// make c hold the cell of s
template <bool DIFF>
MPB_HD void fetch_cube(const MetView &g, const Stencil &s, CubeT<DIFF> &c) {
  if (c.ix != s.ix || c.iy != s.iy || c.iz != s.iz) load_cube(g, s, c);
}

// vertical lerp of one column: w * (lo - hi) + hi with the difference taken in fp32 first (3023-3038)
template <bool DIFF>
MPB_HD double lerp_col(double w, float lo_or_diff, float hi) {
  return w * (double)(DIFF ? lo_or_diff : f_sub(lo_or_diff, hi)) + (double)hi;
}

#define MPB_TRILERP(member)                                                    \
  lerp_f64(s.wx,                                                               \
           lerp_f64(s.wy, lerp_col<DIFF>(s.wz, c.n000.member, c.n001.member),  \
                    lerp_col<DIFF>(s.wz, c.n010.member, c.n011.member)),       \
           lerp_f64(s.wy, lerp_col<DIFF>(s.wz, c.n100.member, c.n101.member),  \
                    lerp_col<DIFF>(s.wz, c.n110.member, c.n111.member)))

// the same on a named cube and stencil (DIFF cube: CubeT<true>)
#define MPB_TRILERP_OF(c, s, member)                                                  \
  lerp_f64((s).wx,                                                                    \
           lerp_f64((s).wy, lerp_col<true>((s).wz, (c).n000.member, (c).n001.member), \
                    lerp_col<true>((s).wz, (c).n010.member, (c).n011.member)),        \
           lerp_f64((s).wy, lerp_col<true>((s).wz, (c).n100.member, (c).n101.member), \
                    lerp_col<true>((s).wz, (c).n110.member, (c).n111.member)))

// Stencil of (lon, lat, p) with the cube made to hold its cell: indices + weights of intpol_met_space_3d's init part
// (2997-3021).  INVARIANT: whenever c.ax names a cell, the cube holds that cell (every path that moves c.ax fetches).
// Production device build: ONE combined test "still inside the cell the cube holds" decides between the register-only
// fast path and the full search -- an RK stage moves a parcel by a few km, so ~85 % of all lookups take it.
template <class CubeT>
MPB_HD void locate(const MetView &g, double lon, double lat, double p, CubeT &c, Stencil &s) {
  double lon2, lat2;
  clamp_horizontal(g, lon, lat, lon2, lat2);
  CellAxes &a = c.ax;
#if MPB_FAST_QUOT
  const bool inside = a.ix >= 0 && cell_holds(a.x, a.ix, g.nx, g.lon_asc, lon2) && cell_holds(a.y, a.iy, g.ny, g.lat_asc, lat2) &&
                      cell_holds(a.z, a.iz, g.nz, g.p_asc, p);
  if (!inside) {
    lon_cell(g, lon2, a);
    lat_cell(g, lat2, a);
    p_cell(g, p, a);
    s.ix = a.ix; s.iy = a.iy; s.iz = a.iz;
    fetch_cube(g, s, c);
  }
#else
  lon_cell(g, lon2, a);
  lat_cell(g, lat2, a);
  p_cell(g, p, a);
  s.ix = a.ix; s.iy = a.iy; s.iz = a.iz;
  fetch_cube(g, s, c);
#endif
  s.ix = a.ix; s.iy = a.iy; s.iz = a.iz;
  s.wx = quot(a.x.hi - lon2, a.x.hi - a.x.lo, a.x.rd);
  s.wy = quot(a.y.hi - lat2, a.y.hi - a.y.lo, a.y.rd);
  s.wz = quot(a.z.hi - p, a.z.hi - a.z.lo, a.z.rd);
}

// time weight of met0 (3133)
MPB_HD double time_weight(const MetView &g, double ts) { return quot(g.t1 - ts, g.dt01, g.r_dt01); }

// u, v, w at (p, lon, lat) for the time weight wt: intpol_met_time_3d x3 sharing one stencil (3112-3137, 3638-3643)
template <bool DIFF>
MPB_HD void wind_at(const MetView &g, double wt, double lon, double lat, double p, CubeT<DIFF> &c,
                    double &u, double &v, double &w) {
  Stencil s;
  locate(g, lon, lat, p, c, s);
  u = lerp_f64(wt, MPB_TRILERP(u0), MPB_TRILERP(u1));
  v = lerp_f64(wt, MPB_TRILERP(v0), MPB_TRILERP(v1));
  w = lerp_f64(wt, MPB_TRILERP(w0), MPB_TRILERP(w1));
}

// Alternative cube for the wind lookups (build flag MPB_CUBE_F64): per (x, y) column and field the upper-level value
// and the fp32 difference "lower - upper", both already promoted to fp64.  A lookup in an unchanged cell then needs no
// F2F.F64.F32 conversions at all (they run at 16 lanes/clk/SM and are the kernel's second limiter), at the price of 96
// live registers.
struct WindCube {
  double hi[4][6], df[4][6];   // [column (x0y0, x0y1, x1y0, x1y1)][u0, v0, w0, u1, v1, w1]
  int ix, iy, iz;
  CellAxes ax;
};
MPB_HD void cube_reset(WindCube &c) { c.ix = -1; c.iy = -1; c.iz = -1; axes_reset(c.ax); }

MPB_HD void fetch_cube(const MetView &g, const Stencil &s, WindCube &c) {
  if (c.ix == s.ix && c.iy == s.iy && c.iz == s.iz) return;
  const size_t sy = (size_t)g.nz, sx = (size_t)g.ny * (size_t)g.nz;
  const Node *b = g.f + ((size_t)s.ix * sx + (size_t)s.iy * sy + (size_t)s.iz);
  const size_t off[4] = {0, sy, sx, sx + sy};
#pragma unroll
  for (int j = 0; j < 4; j++) {
    const Node lo = load_node(b + off[j]), hi = load_node(b + off[j] + 1);
    c.hi[j][0] = (double)hi.u0; c.df[j][0] = (double)f_sub(lo.u0, hi.u0);
    c.hi[j][1] = (double)hi.v0; c.df[j][1] = (double)f_sub(lo.v0, hi.v0);
    c.hi[j][2] = (double)hi.w0; c.df[j][2] = (double)f_sub(lo.w0, hi.w0);
    c.hi[j][3] = (double)hi.u1; c.df[j][3] = (double)f_sub(lo.u1, hi.u1);
    c.hi[j][4] = (double)hi.v1; c.df[j][4] = (double)f_sub(lo.v1, hi.v1);
    c.hi[j][5] = (double)hi.w1; c.df[j][5] = (double)f_sub(lo.w1, hi.w1);
  }
  c.ix = s.ix; c.iy = s.iy; c.iz = s.iz;
}

MPB_HD double trilerp(const Stencil &s, const WindCube &c, int k) {
  return lerp_f64(s.wx,
                  lerp_f64(s.wy, s.wz * c.df[0][k] + c.hi[0][k], s.wz * c.df[1][k] + c.hi[1][k]),
                  lerp_f64(s.wy, s.wz * c.df[2][k] + c.hi[2][k], s.wz * c.df[3][k] + c.hi[3][k]));
}

MPB_HD void wind_at(const MetView &g, double wt, double lon, double lat, double p, WindCube &c,
                    double &u, double &v, double &w) {
  Stencil s;
  locate(g, lon, lat, p, c, s);
  u = lerp_f64(wt, trilerp(s, c, 0), trilerp(s, c, 3));
  v = lerp_f64(wt, trilerp(s, c, 1), trilerp(s, c, 4));
  w = lerp_f64(wt, trilerp(s, c, 2), trilerp(s, c, 5));
}

// temperature at (ts, p, lon, lat)
template <bool DIFF>
MPB_HD double temperature_at(const MetView &g, double ts, double lon, double lat, double p, CubeT<DIFF> &c) {
  Stencil s;
  locate(g, lon, lat, p, c, s);
  return lerp_f64(time_weight(g, ts), MPB_TRILERP(t0), MPB_TRILERP(t1));
}

// bilinear value of 4 corners with the nearest-neighbour rule for non-finite data (3084-3107)
MPB_HD double bilerp_guarded(double wx, double wy, double a00, double a01, double a10, double a11) {
  if (isfinite(a00) && isfinite(a01) && isfinite(a10) && isfinite(a11))
    return lerp_f64(wx, lerp_f64(wy, a00, a01), lerp_f64(wy, a10, a11));
  if (wy < 0.5) return wx < 0.5 ? a11 : a01;
  return wx < 0.5 ? a10 : a00;
}

// time blend with the same guard (3163-3169)
MPB_HD double time_blend_guarded(double wt, double a0, double a1) {
  if (isfinite(a0) && isfinite(a1)) return lerp_f64(wt, a0, a1);
  return wt < 0.5 ? a1 : a0;
}

// ps and pbl at the parcel position (INTPOL_2D(pbl,1); INTPOL_2D(ps,0), 4606-4617)
MPB_HD void surface_at(const MetView &g, double ts, double lon, double lat, CellAxes &ax, double &ps, double &pbl) {
  Stencil s;
  stencil_2d(g, lon, lat, ax, s);
  const size_t b = (size_t)s.ix * (size_t)g.ny + (size_t)s.iy;
  const size_t sx = (size_t)g.ny;
  const double wt = time_weight(g, ts);
  const float4 a00 = ldg(g.s + b), a01 = ldg(g.s + b + 1), a10 = ldg(g.s + b + sx), a11 = ldg(g.s + b + sx + 1);
  const double ps0 = bilerp_guarded(s.wx, s.wy, a00.x, a01.x, a10.x, a11.x);
  const double pb0 = bilerp_guarded(s.wx, s.wy, a00.y, a01.y, a10.y, a11.y);
  const double ps1 = bilerp_guarded(s.wx, s.wy, a00.z, a01.z, a10.z, a11.z);
  const double pb1 = bilerp_guarded(s.wx, s.wy, a00.w, a01.w, a10.w, a11.w);
  ps = time_blend_guarded(wt, ps0, ps1);
  pbl = time_blend_guarded(wt, pb0, pb1);
}

// ----------------------------------------------------------------------------------------------
// module_timesteps (5999-6042)
// ----------------------------------------------------------------------------------------------
MPB_HD double parcel_dt(const MetView &g, const CtlView &c, const Parcel &a) {
  const double dir = (double)c.direction;
  double dt = 0.0;
  if (dir * (a.time - c.t_start) >= 0 && dir * (a.time - c.t_stop) <= 0 && dir * (a.time - c.t) < 0)
    dt = c.t - a.time;
  if (g.local && (a.lon <= g.lon_first || a.lon >= g.lon_last || a.lat <= g.lat_lo || a.lat >= g.lat_hi))
    dt = 0.0;
  return dt;
}

// ----------------------------------------------------------------------------------------------
// module_position (5435-5489)
// NB the reference looks up ps with a *fresh* (zeroed) stencil and init = 0 (5483): the value it
// gets is the time-blended surface pressure of grid node [1][1], not of the parcel's column.
// That is what the reference's outputs contain, so it is reproduced here.
// ----------------------------------------------------------------------------------------------
static MPB_COLD void reflect_at_poles(double &lon, double &lat) {   // 5444-5468
  // FMOD truncates its quotient through an int: beyond 2^31 * 360 (or for a non-finite value) it cannot bring a
  // coordinate back into range and the loops below would never end -- the reference does not return in that case;
  // a device kernel must, so such a parcel is left where it is (its lookups clamp to the grid edge).
  if (!(fabs(lon) < 7.7e11) || !(fabs(lat) < 7.7e11)) return;
  lon = mod360(lon);
  lat = mod360(lat);
  while (lat < -90 || lat > 90) {
    if (lat > 90) { lat = 180 - lat; lon += 180; }
    if (lat < -90) { lat = -180 - lat; lon += 180; }
  }
  while (lon < -180) lon += 360;
  while (lon >= 180) lon -= 360;
}
static MPB_COLD double reflect_p(double bound, double p) { return bound * bound / p; }

MPB_HD void fix_position(const MetView &g, Parcel &a) {
  if (g.coord_type == 0) {
    // already canonical (|lon|, |lat| < 360 make both FMODs the identity): nothing to do
    const bool canonical = below_360(a.lon) && a.lat >= -90 && a.lat <= 90 && a.lon >= -180 && a.lon < 180;
    if (!canonical) reflect_at_poles(a.lon, a.lat);
  } else {
    double x, y;
    clamp_horizontal(g, a.lon, a.lat, x, y);
    a.lon = x; a.lat = y;
  }
  const double ptop = ldg(g.p + g.nz - 1);
  if (a.p < ptop) {
    a.p = reflect_p(ptop, a.p);
  } else if (a.p > 300.) {
    const size_t n11 = (size_t)g.ny + 1;
    const float4 s11 = ldg(g.s + n11);
    const double ps = time_blend_guarded(time_weight(g, a.time), (double)s11.x, (double)s11.z);
    if (a.p > ps) a.p = reflect_p(ps, a.p);
  }
}

// ----------------------------------------------------------------------------------------------
// module_advect, pressure-level branch (3612-3677)
// ----------------------------------------------------------------------------------------------
// Where will the parcel be at the END of the step (one Euler step with the wind of the first stage)?  If that is another grid
// cell, its eight nodes are requested now (prefetch: no registers, no waiting), so that the stage which crosses into it a few
// hundred instructions later finds them in the cache instead of waiting for HBM in the middle of its dependent chain.  The
// indices are first guesses (one-off is harmless); results do not depend on this function.
// MEASURED on B200 and switched OFF: the ~40 extra instructions per parcel-step cost more than the shorter waits save -- C2
// 0.118 ms per step with the prefetch (into L1 or into L2 alike) against 0.111 ms without, C4 share 2.74 against 2.66 ms, C3
// 2.63 against 2.58 ms (profiles/r02l_sweep_prefetch_ahead.jsonl).  The kernel is as sensitive to issued instructions as
// it is to latency.
#ifndef MPB_PREFETCH_AHEAD
#define MPB_PREFETCH_AHEAD 0      // 0: off (default), 1: into L1, 2: into L2 only
#endif
#if MPB_PREFETCH_AHEAD == 2
#define MPB_PF "prefetch.global.L2 [%0];"
#else
#define MPB_PF "prefetch.global.L1 [%0];"
#endif
template <class CubeT>
MPB_HD void prefetch_ahead(const MetView &g, const Parcel &a, double dt, double u, double v, double w, const LonScale &ks, const CubeT &c) {
#if defined(__CUDA_ARCH__) && MPB_PREFETCH_AHEAD
  double x2, y2;
  clamp_horizontal(g, a.lon + dx2coord(ks, dt * u), a.lat + dy2coord(g.coord_type, dt * v), x2, y2);
  int ix = (int)((x2 - g.lon_first) * g.r_lon_d), iy = lat_guess(g, y2), iz = p_guess(g, a.p + dt * w);
  ix = ix < 0 ? 0 : (ix > g.nx - 2 ? g.nx - 2 : ix);
  iy = iy < 0 ? 0 : (iy > g.ny - 2 ? g.ny - 2 : iy);
  iz = iz < 0 ? 0 : (iz > g.nz - 2 ? g.nz - 2 : iz);
  if (ix != c.ix || iy != c.iy || iz != c.iz) {
    const size_t sy = (size_t)g.nz, sx = (size_t)g.ny * (size_t)g.nz;
    const Node *b = g.f + ((size_t)ix * sx + (size_t)iy * sy + (size_t)iz);
    asm volatile(MPB_PF ::"l"(b));           asm volatile(MPB_PF ::"l"(b + 1));
    asm volatile(MPB_PF ::"l"(b + sy));      asm volatile(MPB_PF ::"l"(b + sy + 1));
    asm volatile(MPB_PF ::"l"(b + sx));      asm volatile(MPB_PF ::"l"(b + sx + 1));
    asm volatile(MPB_PF ::"l"(b + sx + sy)); asm volatile(MPB_PF ::"l"(b + sx + sy + 1));
  }
#else
  (void)g; (void)a; (void)dt; (void)u; (void)v; (void)w; (void)ks; (void)c;
#endif
}

// ROLL: keep the stage loop rolled (same arithmetic, a quarter fewer instructions in the kernels that carry diffusion).  For
// advection alone both forms run at the same speed; which of the two is faster with the other modules compiled in was
// measured per module mix and grid (engine.cu MPB_ROLL_SEDI).
template <int ORDER, bool ROLL = false, class CubeT>
MPB_HD void advect(const MetView &g, double dt, Parcel &a, CubeT &c) {
  double um = 0, vm = 0, wm = 0;
  double u = 0, v = 0, w = 0;
  double lat_stage = a.lat;
  const LonScale ks = lon_scale(g.coord_type, a.lat);   // the stages and (Euler, RK4) the final update share it
  double wt = 0;
  auto stage = [&](int i) {
    double x, y, z, dts;
    if (i == 0) {
      dts = 0.0; x = a.lon; y = a.lat; z = a.p;
    } else {
      dts = (i == 3 ? 1.0 : 0.5) * dt;
      x = a.lon + dx2coord(ks, dts * u);
      y = a.lat + dy2coord(g.coord_type, dts * v);
      z = a.p + dts * w;
    }
    lat_stage = y;
    if (i != 2) wt = time_weight(g, a.time + dts);   // stages 1 and 2 are taken at the same time
    wind_at(g, wt, x, y, z, c, u, v, w);
    if (i == 0 && ORDER > 1) prefetch_ahead(g, a, dt, u, v, w, ks, c);
    double k = 1.0;
    if (ORDER == 2) k = (i == 0 ? 0.0 : 1.0);
    else if (ORDER == 4) k = (i == 0 || i == 3 ? 1.0 / 6.0 : 2.0 / 6.0);
    um += k * u; vm += k * v; wm += k * w;
  };
  if (ROLL) {
#pragma unroll 1
    for (int i = 0; i < ORDER; i++) stage(i);
  } else {
#pragma unroll
    for (int i = 0; i < ORDER; i++) stage(i);
  }
  a.time += dt;
  a.lon += (ORDER == 2) ? dx2coord(g.coord_type, dt * um, lat_stage) : dx2coord(ks, dt * um);
  a.lat += dy2coord(g.coord_type, dt * vm);
  a.p += dt * wm;
}

// ----------------------------------------------------------------------------------------------
// interpolation on model levels: intpol_met_4d_zeta (2808-2981), locate_vert (3578-3594), locate_irr_float
// (3525-3555).  Differences from the pressure-level routines that are part of the contract: weights are those of the
// UPPER-index node, time is interpolated first (fp32 subtraction, then fp64), and the level is searched on the
// time- and horizontally interpolated coordinate, walking up from the lowest bracketing level of the 4 columns x 2 times.
// The level accessors below say where a record keeps its search coordinate (h) and which fields are wanted.
// ----------------------------------------------------------------------------------------------
MPB_HD LevelNode load_level(const LevelNode *ptr) {
#ifdef __CUDA_ARCH__
  LevelNode n;
  asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(n.h0), "=f"(n.a0), "=f"(n.b0), "=f"(n.c0), "=f"(n.h1), "=f"(n.a1), "=f"(n.b1), "=f"(n.c1)
      : "l"(ptr));
  return n;
#else
  return *ptr;
#endif
}

// Level accessors: a record type that carries the search coordinate of both time levels next to the values it locates
// (so the 8 records the search ends on are the 8 corners of the interpolation: loaded once), plus a 4-byte load of one
// search coordinate for the column searches.
struct WindLevels {     // LevelNode: search on h, values a, b, c
  typedef LevelNode Rec;
  const LevelNode *f;
  MPB_HD Rec load(size_t i) const { return load_level(f + i); }
  MPB_HD float height(size_t i, int t) const { return t ? ldg(&f[i].h1) : ldg(&f[i].h0); }
  static MPB_HD float h(const Rec &r, int t) { return t ? r.h1 : r.h0; }
};
struct PressureOfZeta { // pz records searched on zeta (y, w), value pressure (x, z)
  typedef float4 Rec;
  const float4 *f;
  MPB_HD Rec load(size_t i) const { return ldg(f + i); }
  MPB_HD float height(size_t i, int t) const { return t ? ldg(&f[i].w) : ldg(&f[i].y); }
  static MPB_HD float h(const Rec &r, int t) { return t ? r.w : r.y; }
  static MPB_HD float value(const Rec &r, int t) { return t ? r.z : r.x; }
};
struct ZetaOfPressure { // pz records searched on pressure (x, z), value zeta (y, w)
  typedef float4 Rec;
  const float4 *f;
  MPB_HD Rec load(size_t i) const { return ldg(f + i); }
  MPB_HD float height(size_t i, int t) const { return t ? ldg(&f[i].z) : ldg(&f[i].x); }
  static MPB_HD float h(const Rec &r, int t) { return t ? r.z : r.x; }
  static MPB_HD float value(const Rec &r, int t) { return t ? r.w : r.y; }
};

template <class L>
struct LevelStencil {
  size_t col[4];                 // first record of the columns (ix,iy), (ix,iy+1), (ix+1,iy), (ix+1,iy+1)
  typename L::Rec lo[4], hi[4];  // their records at levels iz and iz + 1
  int iz;
  double wx, wy, wz, wt;         // weights of the upper-index node / of time level 1
};

// The four weights of the model-level lookup (2857-2861, 2940): quotients by grid constants and by the level spacing.  The
// production device build multiplies by the reciprocal (interval records, rcp + Newton), like the pressure-level path: a
// last-ulp difference in a weight enters the position continuously; the host and the strict build divide.
MPB_HD double level_weight(double num, double den, double rden) {
#if MPB_FAST_QUOT
  (void)den;
  return num * rden;
#else
  (void)rden;
  return num / den;
#endif
}
MPB_HD double level_ratio(double num, double den) {
#if MPB_FAST_QUOT
  return fdiv(num, den);
#else
  return num / den;
#endif
}

// t * (v1 - v0) + v0 with the difference taken in fp32 (2870-2873)
MPB_HD double time_lerp_f32(double wt, float v0, float v1) { return wt * (double)f_sub(v1, v0) + (double)v0; }
MPB_HD double up_lerp(double w, double lo, double hi) { return w * (hi - lo) + lo; }

// locate_irr_float (3525-3555): the guess test ...
MPB_HD bool level_guess_holds(float g0, float g1, double x) { return (g0 <= x && x < g1) || (g0 >= x && x > g1); }
// ... and the bisection that follows a failed guess
template <class L>
static MPB_COLD int bisect_level(const L &lv, size_t col, int n, int t, double x) {
  int lo = 0, hi = n - 1, i = (hi + lo) >> 1;
  if (lv.height(col + i, t) < lv.height(col + i + 1, t)) {
    while (hi > lo + 1) { i = (hi + lo) >> 1; if (lv.height(col + i, t) > x) hi = i; else lo = i; }
  } else {
    while (hi > lo + 1) { i = (hi + lo) >> 1; if (lv.height(col + i, t) <= x) hi = i; else lo = i; }
  }
  return lo;
}
// On a monotonic column (the reference's reader rejects non-monotonic pressure profiles, src/mptrac.c:10120-10128) the
// bisection has exactly one possible answer -- the interval that brackets x, or the end interval x lies beyond -- so a
// hint (the level this column gave at the previous Runge-Kutta stage) that satisfies the bisection's end condition IS
// its answer, at 2 loads instead of ~log2(npl) dependent ones.
MPB_HD bool level_hint_holds(float h0, float h1, int hint, int n, double x) {
  const bool first = hint == 0, last = hint == n - 2;
  return h0 < h1 ? ((first || h0 <= x) && (last || h1 > x)) : ((first || h0 > x) && (last || h1 <= x));
}
// A failed guess at level k (bounds a = h[k], b = h[k+1] already loaded): neighbouring columns and consecutive
// Runge-Kutta stages differ by one level far more often than by several, so the two adjacent intervals are tried against
// the bisection's end condition (one more load each, independent of one another) before the ~log2(npl) dependent loads
// of the bisection itself.  Same answer by the argument above.
#ifndef MPB_LEVEL_NEAR
#define MPB_LEVEL_NEAR 1
#endif
template <class L>
static MPB_COLD int locate_after_miss(const L &lv, size_t col, int n, int t, double x, int k, float a, float b) {
#if MPB_LEVEL_NEAR
  const bool has_below = k > 0, has_above = k + 2 <= n - 1;
  const float below = has_below ? lv.height(col + k - 1, t) : a;
  const float above = has_above ? lv.height(col + k + 2, t) : b;
  if (has_below && level_hint_holds(below, a, k - 1, n, x)) return k - 1;
  if (has_above && level_hint_holds(b, above, k + 1, n, x)) return k + 1;
#else
  (void)k; (void)a; (void)b;
#endif
  return bisect_level(lv, col, n, t, x);
}
template <class L>
static MPB_COLD int locate_level(const L &lv, size_t col, int n, int t, double x, int ig) {
  const float a = lv.height(col + ig, t), b = lv.height(col + ig + 1, t);
  if (level_guess_holds(a, b, x)) return ig;
  return locate_after_miss(lv, col, n, t, x, ig, a, b);
}

// search coordinate at one level of the four columns: time first, then latitude, then longitude (2866-2889)
template <class L>
MPB_HD double level_height(const LevelStencil<L> &s, const typename L::Rec (&r)[4]) {
  const double h00 = time_lerp_f32(s.wt, L::h(r[0], 0), L::h(r[0], 1)), h01 = time_lerp_f32(s.wt, L::h(r[1], 0), L::h(r[1], 1));
  const double h10 = time_lerp_f32(s.wt, L::h(r[2], 0), L::h(r[2], 1)), h11 = time_lerp_f32(s.wt, L::h(r[3], 0), L::h(r[3], 1));
  return up_lerp(s.wx, up_lerp(s.wy, h00, h01), up_lerp(s.wy, h10, h11));
}

// The init part of intpol_met_4d_zeta (2825-2938).  The reference searches the eight (column, time) pairs one after
// the other, each column starting from the previous column's answer (locate_vert, 3578-3594).  Here the loads are
// hoisted out of that chain -- the first column of both time levels together, then the guess tests of the other three
// columns at the first column's level -- and the chain itself is walked on values already in registers; a column whose
// guess fails takes the reference's path literally.  Answers are identical, the dependent load rounds drop from ~16 to 3.
template <class L>
MPB_HD void locate_on_levels(const MetView &g, const L &lv, double ts, double height, double lon, double lat, LevelStencil<L> &s,
                             int *hint = nullptr /* [2]: first-column levels of the previous lookup, -1 = none */,
                             bool cached = false /* s holds the stencil of the previous lookup (s.iz < 0: nothing) */) {
  double lon2, lat2;
  clamp_horizontal(g, lon, lat, lon2, lat2);
  const int ix = lon_interval(g, lon2);
  AxisCell cy;
  const int iy = locate_cell(g.lat, g.latc, g.ny, g.lat_asc, lat2, lat_guess(g, lat2), cy);
  const int n = g.npl;
  const size_t npl = (size_t)n, sx = (size_t)g.ny * npl;
  const AxisCell cx = load_cell(g.lonc + ix);
#if MPB_LEVEL_CACHE
  // Record cache: the stencil `s` of the previous Runge-Kutta
  // stage is still valid when the parcel is in the same four columns and every (column, time) pair still brackets
  // `height` at level s.iz -- the first column by the bisection's end condition (level_hint_holds), the others by the
  // reference's guess test (level_guess_holds) --, because then all eight column searches answer s.iz, the minimum and
  // the maximum coincide and the walk does not move.  Only the weights change: no load at all.
  if (cached && s.iz >= 0 && s.col[0] == (size_t)ix * sx + (size_t)iy * npl) {
    bool same = true;
#pragma unroll
    for (int t = 0; t < 2; t++) {
      same = same && level_hint_holds(L::h(s.lo[0], t), L::h(s.hi[0], t), s.iz, n, height);
#pragma unroll
      for (int j = 1; j < 4; j++) same = same && level_guess_holds(L::h(s.lo[j], t), L::h(s.hi[j], t), height);
    }
    if (same) {
      if (hint) { hint[0] = s.iz; hint[1] = s.iz; }
      s.wt = level_weight(ts - g.t0, g.t1 - g.t0, g.r_dt01);
      s.wx = level_weight(lon2 - cx.lo, cx.hi - cx.lo, cx.rd);
      s.wy = level_weight(lat2 - cy.lo, cy.hi - cy.lo, cy.rd);
      const double bot = level_height<L>(s, s.lo), top = level_height<L>(s, s.hi);
      s.wz = level_ratio(height - bot, top - bot);
      return;
    }
  }
#else
  (void)cached;
#endif
  s.col[0] = (size_t)ix * sx + (size_t)iy * npl;
  s.col[1] = s.col[0] + npl;
  s.col[2] = s.col[0] + sx;
  s.col[3] = s.col[2] + npl;
  const float f0 = lv.height(0, 0), f1 = lv.height(1, 0);          // heights0[0][0][0] vs [0][0][1], 2914-2921

  // round 1: first column, guess 0 and the hint, both time levels
  float g0[2], g1[2], q0[2], q1[2];
  int hk[2];
#pragma unroll
  for (int t = 0; t < 2; t++) {
    hk[t] = hint && hint[t] >= 0 ? hint[t] : 0;
    g0[t] = lv.height(s.col[0], t); g1[t] = lv.height(s.col[0] + 1, t);
    q0[t] = lv.height(s.col[0] + hk[t], t); q1[t] = lv.height(s.col[0] + hk[t] + 1, t);
  }
  int k0[2];
#pragma unroll
  for (int t = 0; t < 2; t++) {
    if (level_guess_holds(g0[t], g1[t], height)) k0[t] = 0;
    else if (hint && hint[t] >= 0 && level_hint_holds(q0[t], q1[t], hk[t], n, height)) k0[t] = hk[t];
    else if (t == 1 && level_hint_holds(lv.height(s.col[0] + k0[0], 1), lv.height(s.col[0] + k0[0] + 1, 1), k0[0], n, height))
      k0[t] = k0[0];   // the two time levels of a column nearly always agree
    else if (hint && hint[t] >= 0) k0[t] = locate_after_miss(lv, s.col[0], n, t, height, hk[t], q0[t], q1[t]);
    else k0[t] = bisect_level(lv, s.col[0], n, t, height);
    if (hint) hint[t] = k0[t];
  }
  // round 2: the other columns, in the reference's order (ix+1,iy), (ix,iy+1), (ix+1,iy+1), tested at the first column's level
  const int order[3] = {2, 1, 3};
  float a[2][3], b[2][3];
#pragma unroll
  for (int t = 0; t < 2; t++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      a[t][j] = lv.height(s.col[order[j]] + k0[t], t);
      b[t][j] = lv.height(s.col[order[j]] + k0[t] + 1, t);
    }
  int kmin = 0, kmax = 0;
#pragma unroll
  for (int t = 0; t < 2; t++) {
    int k = k0[t], lo = k, hi = k;
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (k == k0[t]) { if (!level_guess_holds(a[t][j], b[t][j], height)) k = locate_after_miss(lv, s.col[order[j]], n, t, height, k, a[t][j], b[t][j]); }
      else k = locate_level(lv, s.col[order[j]], n, t, height, k);
      lo = k < lo ? k : lo; hi = k > hi ? k : hi;
    }
    if (t == 0) { kmin = lo; kmax = hi; } else { kmin = lo < kmin ? lo : kmin; kmax = hi > kmax ? hi : kmax; }
  }
  // round 3: the eight records, then the walk up to the level that brackets `height` on the interpolated coordinate
  s.iz = kmin;
  s.wt = level_weight(ts - g.t0, g.t1 - g.t0, g.r_dt01);
  s.wx = level_weight(lon2 - cx.lo, cx.hi - cx.lo, cx.rd);
  s.wy = level_weight(lat2 - cy.lo, cy.hi - cy.lo, cy.rd);
#pragma unroll
  for (int j = 0; j < 4; j++) { s.lo[j] = lv.load(s.col[j] + s.iz); s.hi[j] = lv.load(s.col[j] + s.iz + 1); }
  double bot = level_height<L>(s, s.lo), top = level_height<L>(s, s.hi);
  const bool desc = f0 > f1, asc = f0 < f1;
  while ((desc && ((bot <= height) || (top > height)) && (bot >= height) && (s.iz < kmax)) ||
         (asc && ((bot >= height) || (top < height)) && (bot <= height) && (s.iz < kmax))) {
    s.iz++;
    bot = top;
#pragma unroll
    for (int j = 0; j < 4; j++) { s.lo[j] = s.hi[j]; s.hi[j] = lv.load(s.col[j] + s.iz + 1); }
    top = level_height<L>(s, s.hi);
  }
  s.wz = level_ratio(height - bot, top - bot);
}

// one field at the 8 corners: longitude, then latitude, then level (2941-2980); v[column][0 = iz, 1 = iz + 1]
template <class L>
MPB_HD double combine_levels(const LevelStencil<L> &s, const double (&v)[4][2]) {
  const double b00 = up_lerp(s.wx, v[0][0], v[2][0]), b10 = up_lerp(s.wx, v[1][0], v[3][0]);
  const double b01 = up_lerp(s.wx, v[0][1], v[2][1]), b11 = up_lerp(s.wx, v[1][1], v[3][1]);
  return up_lerp(s.wz, up_lerp(s.wy, b00, b10), up_lerp(s.wy, b01, b11));
}

// u, v and the vertical velocity on model levels: three intpol_met_4d_zeta calls sharing one stencil (3646-3657, 3714-3722)
MPB_HD void wind_on_levels(const MetView &g, const LevelNode *f, double ts, double height, double lon, double lat,
                           double &u, double &v, double &w, int *hint = nullptr, LevelStencil<WindLevels> *keep = nullptr) {
  const WindLevels lv = {f};
  LevelStencil<WindLevels> local;
  LevelStencil<WindLevels> &s = keep ? *keep : local;
  locate_on_levels(g, lv, ts, height, lon, lat, s, hint, keep != nullptr);
  double a[4][2], b[4][2], c[4][2];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    a[j][0] = time_lerp_f32(s.wt, s.lo[j].a0, s.lo[j].a1); a[j][1] = time_lerp_f32(s.wt, s.hi[j].a0, s.hi[j].a1);
    b[j][0] = time_lerp_f32(s.wt, s.lo[j].b0, s.lo[j].b1); b[j][1] = time_lerp_f32(s.wt, s.hi[j].b0, s.hi[j].b1);
    c[j][0] = time_lerp_f32(s.wt, s.lo[j].c0, s.lo[j].c1); c[j][1] = time_lerp_f32(s.wt, s.hi[j].c0, s.hi[j].c1);
  }
  u = combine_levels(s, a); v = combine_levels(s, b); w = combine_levels(s, c);
}

// the conversions pressure <-> zeta (3690-3695, 3750-3755, 3779-3782)
template <class L>
MPB_HD double convert_on_levels(const MetView &g, const L &lv, double ts, double height, double lon, double lat) {
  LevelStencil<L> s;
  locate_on_levels(g, lv, ts, height, lon, lat, s);
  double a[4][2];
#pragma unroll
  for (int j = 0; j < 4; j++) {
    a[j][0] = time_lerp_f32(s.wt, L::value(s.lo[j], 0), L::value(s.lo[j], 1));
    a[j][1] = time_lerp_f32(s.wt, L::value(s.hi[j], 0), L::value(s.hi[j], 1));
  }
  return combine_levels(s, a);
}
MPB_HD double zeta_of_pressure(const MetView &g, double ts, double p, double lon, double lat) {
  return convert_on_levels(g, ZetaOfPressure{g.pz}, ts, p, lon, lat);
}
MPB_HD double pressure_of_zeta(const MetView &g, double ts, double zeta, double lon, double lat) {
  return convert_on_levels(g, PressureOfZeta{g.pz}, ts, zeta, lon, lat);
}

// module_advect with a model-level vertical coordinate: VERT_COORD 2 (omega on model levels, the parcel's vertical
// coordinate is its pressure) and 1 / 3 (zeta / eta, kept in the quantity `zq`)
// level_hint: in / out, the level the parcel's first column gave last time (or -1); only ever used after verification
template <int ORDER>
MPB_HD void advect_on_levels(const MetView &g, int vert_coord, double dt, Parcel &a, double *zq, int *level_hint = nullptr) {
  const LevelNode *f = vert_coord == 2 ? g.lp : g.lz;
  if (zq) *zq = zeta_of_pressure(g, a.time, a.p, a.lon, a.lat);
  const double z0 = zq ? *zq : a.p;
  double um = 0, vm = 0, wm = 0, u = 0, v = 0, w = 0, lat_stage = a.lat;
  int hint[2] = {-1, -1};
  if (level_hint && *level_hint >= 0 && *level_hint <= g.npl - 2) hint[0] = *level_hint;
  const LonScale ks = lon_scale(g.coord_type, a.lat);   // metre -> degree divisor of the step's start latitude, shared by the stages
#if MPB_LEVEL_CACHE
  LevelStencil<WindLevels> kept;
  kept.iz = -1;
#endif
  // (a rolled stage loop: unrolled, the RK4 kernel is 9 400 instructions = 150 KB and stalls on instruction fetch --
  // ncu r01h: "no instruction" 5.8 stall cycles per issue)
#pragma unroll 1
  for (int i = 0; i < ORDER; i++) {
    double x, y, z, dts;
    if (i == 0) {
      dts = 0.0; x = a.lon; y = a.lat; z = z0;
    } else {
      dts = (i == 3 ? 1.0 : 0.5) * dt;
      x = a.lon + dx2coord(ks, dts * u);
      y = a.lat + dy2coord(g.coord_type, dts * v);
      z = z0 + dts * w;
    }
    lat_stage = y;
#if MPB_LEVEL_CACHE
    wind_on_levels(g, f, a.time + dts, z, x, y, u, v, w, hint, &kept);
#else
    wind_on_levels(g, f, a.time + dts, z, x, y, u, v, w, hint);
#endif
    double k = 1.0;
    if (ORDER == 2) k = (i == 0 ? 0.0 : 1.0);
    else if (ORDER == 4) k = (i == 0 || i == 3 ? 1.0 / 6.0 : 2.0 / 6.0);
    um += k * u; vm += k * v; wm += k * w;
  }
  if (level_hint) *level_hint = hint[0];
  a.time += dt;
  a.lon += (ORDER == 2) ? dx2coord(g.coord_type, dt * um, lat_stage) : dx2coord(ks, dt * um);
  a.lat += dy2coord(g.coord_type, dt * vm);
  if (zq) {
    *zq = z0 + dt * wm;
    a.p = pressure_of_zeta(g, a.time, *zq, a.lon, a.lat);
  } else {
    a.p = z0 + dt * wm;
  }
}

// ----------------------------------------------------------------------------------------------
// Squares counter RNG + Box-Muller (5784-5828)
// ----------------------------------------------------------------------------------------------
MPB_HD double squares_uniform(uint64_t ctr) {
  const uint64_t key = 0xc8e4fd154ce32f6dull;
  uint64_t x, y, z, t;
  y = x = ctr * key;
  z = y + key;
  x = x * x + y; x = (x >> 32) | (x << 32);
  x = x * x + z; x = (x >> 32) | (x << 32);
  x = x * x + y; x = (x >> 32) | (x << 32);
  t = x = x * x + z; x = (x >> 32) | (x << 32);
  const uint64_t r = t ^ ((x * x + y) >> 32);
  return (double)r / 18446744073709551616.0;   // (double) UINT64_MAX == 2^64
}

// The three normals rs[3*ig], rs[3*ig+1], rs[3*ig+2] of a module_rng(…, 3*np, 1) call whose
// uniform stream started at counter ctr0.  Normal j is made from the uniform pair (j & ~1, +1):
// even j takes the cosine, odd j the sine.
MPB_HD void sin_cos_f(float x, float &s, float &c) {
#ifdef __CUDA_ARCH__
  sincosf(x, &s, &c);   // one argument reduction for both; same values as sinf(x), cosf(x)
#else
  s = sinf(x); c = cosf(x);
#endif
}

MPB_HD void normals3(uint64_t ctr0, uint64_t ig, double &r0, double &r1, double &r2) {
  const uint64_t j0 = 3 * ig;
  const uint64_t pa = j0 & ~1ull;       // pair that holds j0
  const uint64_t pb = pa + 2;           // pair that holds j0+2 (and j0+1 when j0 is odd)
  const double ra = sqrt(-2.0 * log(squares_uniform(ctr0 + pa)));
  const float fa = (float)(2.0 * kPi * squares_uniform(ctr0 + pa + 1));
  const double rb = sqrt(-2.0 * log(squares_uniform(ctr0 + pb)));
  const float fb = (float)(2.0 * kPi * squares_uniform(ctr0 + pb + 1));
  float sa, ca, sb, cb;
  sin_cos_f(fa, sa, ca);
  sin_cos_f(fb, sb, cb);
  if ((j0 & 1ull) == 0) {
    r0 = ra * ca; r1 = ra * sa; r2 = rb * cb;
  } else {
    r0 = ra * sa; r1 = rb * cb; r2 = rb * sb;
  }
}

// ----------------------------------------------------------------------------------------------
// climatological tropopause + weights (213-237, 8358-8376, 12748-12770)
// ----------------------------------------------------------------------------------------------
struct TropoTime {   // the time part of the climatology lookup: second of the year and its month interval
  double sec;
  int it;
};
MPB_HD TropoTime tropopause_time(const ClimView &cl, double t) {
  TropoTime k;
  k.sec = mod_trunc(t, kYear);
  if (!(fabs(k.sec) <= kYear)) k.sec = 0;   // non-finite or beyond the int quotient of FMOD: the loop below would never end
  while (k.sec < 0) k.sec += kYear;
  k.it = find_interval(cl.time, cl.ntime, 1, k.sec);
  return k;
}
MPB_HD double tropopause_pressure(const ClimView &cl, const TropoTime &k, double lat) {
  const double la0 = ldg(cl.lat), la1 = ldg(cl.lat + 1);
  const int il = find_regular_div(la0, la1 - la0, cl.nlat, lat);
  const double y0 = ldg(cl.lat + il), y1 = ldg(cl.lat + il + 1);
  const double *row0 = cl.tropo + (size_t)k.it * cl.nlat + il;
  const double *row1 = row0 + cl.nlat;
  const double p0 = lin(y0, ldg(row0), y1, ldg(row0 + 1), lat);
  const double p1 = lin(y0, ldg(row1), y1, ldg(row1 + 1), lat);
  return lin(ldg(cl.time + k.it), p0, ldg(cl.time + k.it + 1), p1, k.sec);
}
MPB_HD double tropopause_pressure(const ClimView &cl, double t, double lat) {
  return tropopause_pressure(cl, tropopause_time(cl, t), lat);
}

MPB_HD double ramp_weight(double p_full, double p_none, double p) {
  if (p > p_full) return 1;
  if (p < p_none) return 0;
  return lin(p_full, 1.0, p_none, 0.0, p);
}

MPB_HD double weight_pbl(const CtlView &c, double p, double pbl, double ps) {
  return ramp_weight(pbl, pbl - c.pbl_trans * (ps - pbl), p);
}

MPB_HD double weight_tropo(double pt, double p) { return ramp_weight(fdiv(pt, 0.866877899), pt * 0.866877899, p); }

// ----------------------------------------------------------------------------------------------
// module_diff_turb (4588-4734)
// ----------------------------------------------------------------------------------------------
// The layer structure of the parcel's column (surface, boundary-layer top, tropopause) and the diffusivities it implies at a
// pressure: the control file's coefficients of the three layers blended with the two transition weights (4627-4640).  The
// blend is evaluated at the parcel and -- for the vertical gradient of Kz -- 10 m above and below it (4664-4690): one
// function, three pressures.  The ORDER of the floating-point operations is the reference's (it decides the last bit).
struct Column {
  double ps, pbl, tropopause;
};
struct Diffusivity {
  double horizontal, vertical;   // Kx, Kz [m2/s]
};
MPB_HD Diffusivity layer_blend(const CtlView &c, const Column &col, double p) {
  const double in_pbl = weight_pbl(c, p, col.pbl, col.ps);
  const double in_trop = weight_tropo(col.tropopause, p) * (1.0 - in_pbl);
  const double in_strat = 1.0 - in_pbl - in_trop;
  Diffusivity k;
  k.horizontal = in_pbl * c.dx_pbl + in_trop * c.dx_trop + in_strat * c.dx_strat;
  k.vertical = in_pbl * c.dz_pbl + in_trop * c.dz_trop + in_strat * c.dz_strat;
  return k;
}

MPB_HD void diffuse_turbulent(const MetView &g, const ClimView &cl, const CtlView &c, double dt,
                              uint64_t ig, Parcel &a) {
  Column col;
  CellAxes ax;            // a 2-D lookup of its own: the cube's axes must only move together with the cube (see locate)
  axes_reset(ax);
  surface_at(g, a.time, a.lon, a.lat, ax, col.ps, col.pbl);
  if (c.pbl_scheme > 0 && a.p >= col.pbl) return;      // a boundary-layer closure owns the parcels inside the PBL (4612)

  const double ptop = ldg(g.p + g.nz - 1);
  const bool on_sphere = (g.coord_type == 0);
  const TropoTime season = tropopause_time(cl, a.time);
  col.tropopause = tropopause_pressure(cl, season, on_sphere ? a.lat : c.utm_ref_lat);
  const Diffusivity here = layer_blend(c, col, a.p);
  const double span = fabs(dt);

  double g_lon, g_lat, g_p;      // the parcel's three normals of this module_rng call
  normals3(c.ctr_turb, ig, g_lon, g_lat, g_p);

  if (here.horizontal > 0) {
    const double spread = sqrt(2.0 * here.horizontal * span);          // [m]
    a.lon += dx2coord(g.coord_type, g_lon * spread, a.lat);
    a.lat += dy2coord(g.coord_type, g_lat * spread);
  }

  if (here.vertical > 0) {
    const double spread_km = sqrt(2.0 * here.vertical * span) * 1e-3;
    const double p0 = a.p;
    const double probe = 0.01;                                          // [km]: Kz is differenced over +-10 m
    // (MAX / MIN as the reference's ternary macros, src/mptrac.h:1378, 1479: a non-finite ps takes the same side)
    const double p_above = max_of(ptop, min_of(col.ps, p0 + dz2dp(probe, p0)));
    const double p_below = max_of(ptop, min_of(col.ps, p0 + dz2dp(-probe, p0)));
    // the latitude may just have moved: the tropopause is looked up again (12753-12757)
    if (on_sphere && here.horizontal > 0) col.tropopause = tropopause_pressure(cl, season, a.lat);
    const double kz_above = layer_blend(c, col, p_above).vertical, kz_below = layer_blend(c, col, p_below).vertical;
    // well-mixed criterion: the parcel drifts with dKz/dz + Kz dln(rho)/dz, rho ~ exp(-z / H0) (4692-4712)
    const double gradient = fdiv(kz_above - kz_below, 2.0 * probe * 1e3);
    const double drift = gradient + here.vertical * (-1.0 / (1e3 * kH0));
    const double dz = g_p * spread_km + drift * span * 1e-3;            // [km]

    double p_new = p0 + dz2dp(dz, p0);
    for (int bounce = 0; bounce < 10; bounce++) {                       // reflect at the surface and at the model top
      if (p_new > col.ps) p_new = col.ps * col.ps / p_new;
      else if (p_new < ptop) p_new = ptop * ptop / p_new;
      else break;
    }
    a.p = max_of(ptop, min_of(col.ps, p_new));
  }
}

// ----------------------------------------------------------------------------------------------
// module_diff_meso (4266-4339)
// ----------------------------------------------------------------------------------------------
struct Moments {
  float m, q;
  MPB_HD void add(float x) { m = f_add(m, x); q = f_add(q, f_mul(x, x)); }
  MPB_HD float sigma() const {
    const float mean = m / 16.f;
    const float var = f_sub(q / 16.f, f_mul(mean, mean));
    return var > 0 ? sqrtf(var) : 0.f;
  }
};

MPB_HD void diffuse_mesoscale(const MetView &g, const CtlView &k, double dt, uint64_t ig,
                              Parcel &a, float &up, float &vp, float &wp, Cube &c) {
  // raw index search at the parcel position: no wrap / clamp helper here (4283-4285)
  Stencil s;
  s.ix = lon_cell(g, a.lon, c.ax, true);
  s.iy = lat_cell(g, a.lat, c.ax);
  s.iz = p_cell(g, a.p, c.ax);

  fetch_cube(g, s, c);
  Moments mu = {0.f, 0.f}, mv = {0.f, 0.f}, mw = {0.f, 0.f};
#define MPB_ACC(node)                                        \
  mu.add(c.node.u0); mv.add(c.node.v0); mw.add(c.node.w0);   \
  mu.add(c.node.u1); mv.add(c.node.v1); mw.add(c.node.w1);
  MPB_ACC(n000) MPB_ACC(n001) MPB_ACC(n010) MPB_ACC(n011)
  MPB_ACC(n100) MPB_ACC(n101) MPB_ACC(n110) MPB_ACC(n111)
#undef MPB_ACC
  const float usig = mu.sigma(), vsig = mv.sigma(), wsig = mw.sigma();

  const double r = 1 - fdiv(2 * fabs(dt), k.dt_met);
  const double r2 = sqrt(1 - r * r);

  double n0, n1, n2;
  normals3(k.ctr_meso, ig, n0, n1, n2);

  if (k.mesox > 0) {
    up = (float)(r * up + r2 * n0 * k.mesox * usig);
    a.lon += dx2coord(g.coord_type, up * dt, a.lat);
    vp = (float)(r * vp + r2 * n1 * k.mesox * vsig);
    a.lat += dy2coord(g.coord_type, vp * dt);
  }
  if (k.mesoz > 0) {
    wp = (float)(r * wp + r2 * n2 * k.mesoz * wsig);
    a.p += wp * dt;
  }
}

// ----------------------------------------------------------------------------------------------
// module_sedi + sedi (5859-5883, 12506-12535)
// ----------------------------------------------------------------------------------------------
MPB_HD double settling_velocity(double p, double T, double rp, double rhop) {
  const double rp_m = rp * 1e-6;
  const double rho = fdiv(100. * p, kRA * T);
  const double eta = 1.8325e-5 * fdiv(416.16, T + 120.) * pow(fdiv(T, 296.16), 1.5);
  const double v = sqrt(fdiv(8. * kKB * T, kPi * kMAirMolecule));
  const double lambda = fdiv(2. * eta, rho * v);
  const double K = fdiv(lambda, rp_m);
  const double G = 1. + K * (1.249 + 0.42 * exp(fdiv(-0.87, K)));
  return fdiv(2. * (rp_m * rp_m) * (rhop - rho) * kG0, 9. * eta) * G;
}

template <bool DIFF>
MPB_HD void sediment(const MetView &g, double dt, double rp, double rhop, Parcel &a, CubeT<DIFF> &c) {
  const double T = temperature_at(g, a.time, a.lon, a.lat, a.p, c);
  const double vs = settling_velocity(a.p, T, rp, rhop);
  a.p += dz2dp(fdiv(vs * dt, 1000.), a.p);
}

// ----------------------------------------------------------------------------------------------
// module_meteo (5062-5165), the quantities that derive from the resident fields.  INTPOL_TIME_ALL (src/mptrac.h:1278)
// computes ONE stencil (wrapped longitude, clamped latitude) and applies it to every 3-D and 2-D field.
// ----------------------------------------------------------------------------------------------
struct MeteoValues {
  double ps, pbl, t, u, v, w;
};

template <bool DIFF>
MPB_HD void meteo_at(const MetView &g, const Parcel &a, CubeT<DIFF> &c, MeteoValues &m) {
  Stencil s;
  locate(g, a.lon, a.lat, a.p, c, s);
  const double wt = time_weight(g, a.time);
  m.t = lerp_f64(wt, MPB_TRILERP(t0), MPB_TRILERP(t1));
  m.u = lerp_f64(wt, MPB_TRILERP(u0), MPB_TRILERP(u1));
  m.v = lerp_f64(wt, MPB_TRILERP(v0), MPB_TRILERP(v1));
  m.w = lerp_f64(wt, MPB_TRILERP(w0), MPB_TRILERP(w1));
  const size_t b = (size_t)s.ix * (size_t)g.ny + (size_t)s.iy, sx = (size_t)g.ny;
  const float4 a00 = ldg(g.s + b), a01 = ldg(g.s + b + 1), a10 = ldg(g.s + b + sx), a11 = ldg(g.s + b + sx + 1);
  m.ps = time_blend_guarded(wt, bilerp_guarded(s.wx, s.wy, a00.x, a01.x, a10.x, a11.x),
                            bilerp_guarded(s.wx, s.wy, a00.z, a01.z, a10.z, a11.z));
  m.pbl = time_blend_guarded(wt, bilerp_guarded(s.wx, s.wy, a00.y, a01.y, a10.y, a11.y),
                             bilerp_guarded(s.wx, s.wy, a00.w, a01.w, a10.w, a11.w));
}

// One further field of INTPOL_TIME_ALL (src/mptrac.h:1278-1316) at a stencil that is already located; the two time
// levels of a node sit next to each other (.x = met0, .y = met1).  3-D: intpol_met_space_3d per level (3023-3043), plain
// time blend (3133-3136); 2-D: intpol_met_space_2d with its nearest-neighbour rule for non-finite data and the guarded
// time blend (3084-3107, 3163-3169).
MPB_HD double field3_at(const MetView &g, const float2 *f, const Stencil &s, double wt) {
  const size_t sy = (size_t)g.nz, sx = (size_t)g.ny * (size_t)g.nz;
  const float2 *b = f + ((size_t)s.ix * sx + (size_t)s.iy * sy + (size_t)s.iz);
  const float2 n000 = ldg(b), n001 = ldg(b + 1), n010 = ldg(b + sy), n011 = ldg(b + sy + 1);
  const float2 n100 = ldg(b + sx), n101 = ldg(b + sx + 1), n110 = ldg(b + sx + sy), n111 = ldg(b + sx + sy + 1);
  const double v0 = lerp_f64(s.wx, lerp_f64(s.wy, lerp_col<false>(s.wz, n000.x, n001.x), lerp_col<false>(s.wz, n010.x, n011.x)),
                             lerp_f64(s.wy, lerp_col<false>(s.wz, n100.x, n101.x), lerp_col<false>(s.wz, n110.x, n111.x)));
  const double v1 = lerp_f64(s.wx, lerp_f64(s.wy, lerp_col<false>(s.wz, n000.y, n001.y), lerp_col<false>(s.wz, n010.y, n011.y)),
                             lerp_f64(s.wy, lerp_col<false>(s.wz, n100.y, n101.y), lerp_col<false>(s.wz, n110.y, n111.y)));
  return lerp_f64(wt, v0, v1);
}
MPB_HD double field2_at(const MetView &g, const float2 *f, const Stencil &s, double wt) {
  const size_t b = (size_t)s.ix * (size_t)g.ny + (size_t)s.iy, sx = (size_t)g.ny;
  const float2 a00 = ldg(f + b), a01 = ldg(f + b + 1), a10 = ldg(f + b + sx), a11 = ldg(f + b + sx + 1);
  return time_blend_guarded(wt, bilerp_guarded(s.wx, s.wy, a00.x, a01.x, a10.x, a11.x),
                            bilerp_guarded(s.wx, s.wy, a00.y, a01.y, a10.y, a11.y));
}

constexpr double kT0 = 273.15, kKappa = 0.286;
constexpr double kEps = 18.01528 / 28.9644, kLV = 2501000., kCpd = 1003.5;   // EPS = MH2O / MA, LV, CPD (src/mptrac.h:260, 275, 255)
// the quantities module_meteo derives from temperature and water vapour (5136-5152)
struct MoistValues {
  double pw, sh, rh, rhice, tvirt, lapse, tdew, tice;
};
MPB_HD void moist_at(double p, double t, double h2o, MoistValues &m) {
  const double hh = h2o > 0.1e-6 ? h2o : 0.1e-6;                      // MAX((h2o), 0.1e-6)
  m.pw = p * hh / (1. + (1. - kEps) * hh);                             // PW, src/mptrac.h:1859
  m.sh = kEps * hh;                                                    // SH :2024
  m.rh = m.pw / (6.112 * exp(17.62 * (t - kT0) / (243.12 + t - kT0))) * 100.;    // RH :1906
  m.rhice = m.pw / (6.112 * exp(22.46 * (t - kT0) / (272.62 + t - kT0))) * 100.; // RHICE :1936
  m.tvirt = t * (1. + (1. - kEps) * hh);                               // TVIRT :2199
  const double a = kRA * (t * t), r = m.sh / (1. - m.sh);              // lapse_rate, src/mptrac.c:3324-3338
  m.lapse = 1e3 * kG0 * (a + kLV * r * t) / (kCpd * a + (kLV * kLV) * r * kEps);
  m.tdew = kT0 + 243.12 * log(m.pw / 6.112) / (17.62 - log(m.pw / 6.112));       // TDEW :2075
  m.tice = kT0 + 272.62 * log(m.pw / 6.112) / (22.46 - log(m.pw / 6.112));       // TICE :2100
}

MPB_HD double potential_temperature(double p, double t) { return t * pow(1000. / p, kKappa); }   // THETA, src/mptrac.h:2124
MPB_HD double saturation_pressure(double t) { return 6.112 * exp(17.62 * (t - kT0) / (243.12 + t - kT0)); }      // PSAT :1808
MPB_HD double saturation_pressure_ice(double t) { return 6.112 * exp(22.46 * (t - kT0) / (272.62 + t - kT0)); }  // PSICE :1832
MPB_HD double zeta_diagnosed(double ps, double p, double t) {                                                     // ZETA :2293
  return (p / ps <= 0.3 ? 1. : sin(kPi / 2. * (1. - p / ps) / (1. - 0.3))) * potential_temperature(p, t);
}

// ----------------------------------------------------------------------------------------------
// module_diff_pbl (4343-4584, TURB_PBL_SCHEME 1): Hanna / FLEXPART closure inside the boundary layer.  Velocity standard
// deviations, their vertical derivative and the Lagrangian time scales follow from the friction velocity (surface
// stresses ess, nss), the surface heat flux shf through the Monin-Obukhov length, and the height within the PBL; then
// a Langevin update of the three velocity perturbations, the horizontal displacement, and the vertical one in
// geometric height with reflection at the ground and the PBL top.  Parcels above the PBL are left to diff_turb.
// ----------------------------------------------------------------------------------------------
struct PblFields {
  const float2 *ess, *nss, *shf;   // 2-D (MPB_F2_ESS, _NSS, _SHF), both time levels
  const float2 *h2o;               // 3-D (MPB_F3_H2O)
};

// What the closure yields at a height: standard deviations of the three velocity components, the vertical derivative of
// sigma_w, and the Lagrangian time scales.  One function per stability regime (4421-4531); the operation order inside each
// formula is the reference's, everything else -- names, grouping, control flow -- is this file's.
struct PblTurbulence {
  double su, sv, sw;        // sigma_u, sigma_v, sigma_w [m/s]
  double dsw_dz;            // d sigma_w / dz [1/s]
  double tu, tv, tw;        // time scales [s]
};
struct PblState {
  double depth;             // PBL depth zi [m]
  double height;            // parcel height above ground, at least 1 m
  double rel;               // height / depth, kept inside (0, 1)
  double ustar;             // friction velocity, at least 1e-4 m/s
  double obukhov;           // Monin-Obukhov length [m] (1e12 = neutral)
};
MPB_HD PblTurbulence pbl_neutral(const PblState &b) {
  PblTurbulence t;
  const double ratio = b.height / b.ustar;
  const double sw0 = 1.3 * b.ustar * exp(-2e-4 * ratio);
  t.su = max_of(2.0 * b.ustar * exp(-3e-4 * ratio), 1e-5);
  t.sv = max_of(sw0, 1e-5);
  t.sw = max_of(sw0, 1e-5);
  t.dsw_dz = -2e-4 * sw0 / b.ustar;
  t.tu = 0.5 * b.height / t.sw / (1.0 + 1.5e-3 * ratio);
  t.tv = t.tu;
  t.tw = t.tu;
  return t;
}
MPB_HD PblTurbulence pbl_unstable(const PblState &b, double wstar) {
  PblTurbulence t;
  const double zi = b.depth, ol = b.obukhov, zeta = b.rel;
  double dsw2_dz = 0.0;      // d sigma_w^2 / dz
  t.su = max_of(b.ustar * pow(max_of(12.0 - 0.5 * zi / ol, 0.0), 1.0 / 3.0), 1e-6);
  t.sv = t.su;
  if (zeta < 0.03) {
    const double arg = max_of(3.0 * zeta - ol / zi, 1e-12);
    t.sw = 0.96 * wstar * pow(arg, 1.0 / 3.0);
    dsw2_dz = 1.8432 * (wstar * wstar) / zi * pow(arg, -1.0 / 3.0);
  } else if (zeta < 0.4) {
    const double arg = max_of(3.0 * zeta - ol / zi, 1e-12);
    const double surface_form = 0.96 * pow(arg, 1.0 / 3.0), mixed_form = 0.763 * pow(zeta, 0.175);
    if (surface_form < mixed_form) {
      t.sw = wstar * surface_form;
      dsw2_dz = 1.8432 * (wstar * wstar) / zi * pow(arg, -1.0 / 3.0);
    } else {
      t.sw = wstar * mixed_form;
      dsw2_dz = 0.203759 * (wstar * wstar) / zi * pow(zeta, -0.65);
    }
  } else if (zeta < 0.96) {
    t.sw = 0.722 * wstar * pow(1.0 - zeta, 0.207);
    dsw2_dz = -0.215812 * (wstar * wstar) / zi * pow(1.0 - zeta, -0.586);
  } else {
    t.sw = 0.37 * wstar;
    dsw2_dz = 0.0;
  }
  t.sw = max_of(t.sw, 1e-6);
  t.dsw_dz = t.sw > 1e-12 ? 0.5 * dsw2_dz / t.sw : 0.0;
  t.tu = 0.15 * zi / max_of(t.su, 1e-12);
  t.tv = t.tu;
  if (b.height < fabs(ol)) {
    const double denom = 0.55 - 0.38 * fabs(b.height / ol);
    t.tw = 0.1 * b.height / (t.sw * max_of(denom, 0.05));
  } else if (zeta < 0.1) {
    t.tw = 0.59 * b.height / t.sw;
  } else {
    t.tw = 0.15 * zi / t.sw * (1.0 - exp(-5.0 * zeta));
  }
  return t;
}
MPB_HD PblTurbulence pbl_stable(const PblState &b) {
  PblTurbulence t;
  const double fade = 1.0 - b.rel;
  t.su = max_of(2.0 * b.ustar * fade, 1e-6);
  t.sv = max_of(1.3 * b.ustar * fade, 1e-6);
  t.sw = max_of(1.3 * b.ustar * fade, 1e-6);
  t.dsw_dz = -1.3 * b.ustar / b.depth;
  t.tu = 0.15 * b.depth / t.su * sqrt(b.rel);
  t.tv = 0.467 * t.tu;
  t.tw = 0.1 * b.depth / t.sw * pow(b.rel, 0.8);
  return t;
}

MPB_HD void diffuse_pbl(const MetView &g, const PblFields &f, uint64_t ctr, double dt, uint64_t ig, Parcel &a,
                        float &up, float &vp, float &wp) {
  // where is the parcel inside the boundary layer?  (4367-4393)
  CellAxes ax;
  axes_reset(ax);
  double ps, pbl;
  surface_at(g, a.time, a.lon, a.lat, ax, ps, pbl);
  if (a.p < pbl) return;
  if (!(ps > 0.0 && pbl > 0.0 && ps > pbl)) return;
  const double p = a.p < ps ? a.p : ps;
  const double z_ground = altitude(ps);                              // [km]
  const double z_raw = 1e3 * (altitude(p) - z_ground);               // [m]
  PblState b;
  b.depth = 1e3 * (altitude(pbl) - z_ground);
  if (!(b.depth > 1.0)) return;
  const double z = clamp_of(z_raw, 0.0, b.depth);
  b.rel = clamp_of(z / b.depth, 1e-6, 1.0 - 1e-6);
  b.height = max_of(z, 1.0);
  // surface stress, air density and heat flux at the parcel -> friction velocity and Obukhov length (4395-4419)
  Stencil s2;
  stencil_2d(g, a.lon, a.lat, ax, s2);
  const double wt = time_weight(g, a.time);
  const double ess = field2_at(g, f.ess, s2, wt), nss = field2_at(g, f.nss, s2, wt);
  CubeT<true> cube;
  cube_reset(cube);
  Stencil s;
  locate(g, a.lon, a.lat, p, cube, s);
  const double t = lerp_f64(wt, MPB_TRILERP_OF(cube, s, t0), MPB_TRILERP_OF(cube, s, t1));
  const double h2o = field3_at(g, f.h2o, s, wt);
  const double hh = max_of(h2o, 0.1e-6);
  const double tv = t * (1. + (1. - kEps) * hh);                                    // TVIRT
  const double thetav = potential_temperature(p, t) * (1. + (1. - kEps) * max_of(hh, 0.1e-6));   // THETAVIRT :2153
  const double rho = 100. * p / (kRA * tv);                                         // RHO
  const double stress = sqrt(ess * ess + nss * nss);
  if (!(rho > 0.0)) return;
  b.ustar = max_of(1e-4, sqrt(max_of(stress / rho, 0.0)));
  const double shf = field2_at(g, f.shf, s2, wt);
  b.obukhov = 1e12;
  if (fabs(shf) > 1e-6) b.obukhov = thetav * rho * kCpd * (b.ustar * b.ustar) * b.ustar / (0.40 * kG0 * shf);   // KARMAN = 0.40
  // the closure of the stability regime
  PblTurbulence k;
  if (b.depth / fabs(b.obukhov) < 1.0) {
    k = pbl_neutral(b);
  } else if (b.obukhov < 0.0) {
    const double wstar_cubed = -kG0 / thetav * shf / (rho * kCpd) * b.depth;
    k = pbl_unstable(b, pow(max_of(wstar_cubed, 0.0), 1.0 / 3.0));
  } else {
    k = pbl_stable(b);
  }
  k.tu = max_of(k.tu, 10.0);
  k.tv = max_of(k.tv, 10.0);
  k.tw = max_of(k.tw, 30.0);
  if (!(k.su > 0.0 && k.sv > 0.0 && k.sw > 0.0 && k.tu > 0.0 && k.tv > 0.0 && k.tw > 0.0)) return;

  // Langevin update of the three velocity perturbations with the well-mixed drift term, then the displacement (4545-4583)
  double n0, n1, n2;
  normals3(ctr, ig, n0, n1, n2);
  const double span = fabs(dt);
  const double keep_u = exp(-span / k.tu), kick_u = sqrt(max_of(0.0, 1.0 - keep_u * keep_u));
  const double keep_v = exp(-span / k.tv), kick_v = sqrt(max_of(0.0, 1.0 - keep_v * keep_v));
  up = (float)(up * keep_u + k.su * kick_u * n0);
  vp = (float)(vp * keep_v + k.sv * kick_v * n1);
  const double keep_w = exp(-span / k.tw), kick_w = sqrt(max_of(0.0, 1.0 - keep_w * keep_w));
  const double dlnrho_dz = -1.0 / (1e3 * kH0);
  wp = (float)(wp * keep_w + k.sw * kick_w * n2 + k.tw * (1.0 - keep_w) * (2.0 * k.sw * k.dsw_dz + dlnrho_dz * (k.sw * k.sw)));
  a.lon += dx2coord(g.coord_type, up * dt, a.lat);
  a.lat += dy2coord(g.coord_type, vp * dt);
  double z_new = z + wp * dt;
  // reflect at the ground and at the PBL top (the reference's loop has no bound: an infinite displacement would never leave
  // it; a kernel must return, so the reflections are counted)
  for (int bounce = 0; bounce < 4096 && (z_new < 0.0 || z_new > b.depth); bounce++) {
    if (z_new < 0.0) { z_new = -z_new; wp = -wp; }
    if (z_new > b.depth) { z_new = 2.0 * b.depth - z_new; wp = -wp; }
  }
  a.p = kP0 * exp(-(z_ground + z_new / 1000.0) / kH0);    // P(z), src/mptrac.h:1784
  a.p = clamp_of(a.p, pbl, ps);
}

// ----------------------------------------------------------------------------------------------
// module_convection (4102-4171): the mixing range reaches from the surface to the PBL top (CONV_MIX_PBL) and / or to
// the equilibrium level where CAPE (and CIN) pass their thresholds; the parcel's new pressure is uniformly distributed
// in density over that range, `r` being the parcel's uniform random number
// ----------------------------------------------------------------------------------------------
struct ConvView {
  double cape, cin, pbl_trans;          // ctl->conv_cape, conv_cin, conv_pbl_trans
  int mix_pbl;                          // ctl->conv_mix_pbl
  const float2 *fcape, *fcin, *fpel;    // the 2-D fields (both time levels), needed when cape >= 0
};
MPB_HD void convect(const MetView &g, const ConvView &k, double r, Parcel &a) {
  CellAxes ax;
  axes_reset(ax);
  double ps, pbl;
  surface_at(g, a.time, a.lon, a.lat, ax, ps, pbl);
  const double pbot = ps;
  double ptop = ps;
  if (k.mix_pbl) ptop = pbl - k.pbl_trans * (ps - pbl);
  if (k.cape >= 0) {
    Stencil s;
    stencil_2d(g, a.lon, a.lat, ax, s);
    const double wt = time_weight(g, a.time);
    const double cape = field2_at(g, k.fcape, s, wt), cin = field2_at(g, k.fcin, s, wt), pel = field2_at(g, k.fpel, s, wt);
    if (isfinite(cape) && cape >= k.cape && (k.cin <= 0 || (isfinite(cin) && cin >= k.cin))) ptop = ptop < pel ? ptop : pel;   // GSL_MIN
  }
  if (ptop != pbot && a.p >= ptop) {
    CubeT<true> c;
    cube_reset(c);
    const double tbot = temperature_at(g, a.time, a.lon, a.lat, pbot, c);
    const double ttop = temperature_at(g, a.time, a.lon, a.lat, ptop, c);
    const double rhobot = pbot / tbot, rhotop = ptop / ttop;
    const double rho = rhobot + (rhotop - rhobot) * r;
    a.p = lin(rhobot, pbot, rhotop, ptop, rho);
  }
}

// module_isosurf_init / module_isosurf (4886-5004): the conserved variable of a parcel (pressure, density p / T,
// potential temperature) and the pressure that restores it; mode 4 follows a balloon's pressure time series
MPB_HD double isosurf_variable(const MetView &g, int mode, const Parcel &a) {
  if (mode == 1) return a.p;
  CubeT<true> c;
  cube_reset(c);
  const double t = temperature_at(g, a.time, a.lon, a.lat, a.p, c);
  return mode == 2 ? a.p / t : potential_temperature(a.p, t);
}
MPB_HD double isosurf_pressure(const MetView &g, int mode, double var, const Parcel &a, const double *ts, const double *ps, int n) {
  if (mode == 1) return var;
  if (mode == 2 || mode == 3) {
    CubeT<true> c;
    cube_reset(c);
    const double t = temperature_at(g, a.time, a.lon, a.lat, a.p, c);
    return mode == 2 ? var * t : 1000. * pow(var / t, -1. / kKappa);
  }
  if (a.time <= ts[0]) return ps[0];
  if (a.time >= ts[n - 1]) return ps[n - 1];
  const int mid = (n - 1) >> 1;
  const int i = find_interval(ts, n, ts[mid] < ts[mid + 1], a.time);   // locate_irr (3495-3521)
  return ps[i] + (ps[i + 1] - ps[i]) / (ts[i + 1] - ts[i]) * (a.time - ts[i]);   // LIN, src/mptrac.h:1351
}

// module_bound_cond (3789-3881): is the parcel inside the latitude / pressure window and -- where asked -- inside the
// surface layer (pressure depth, height, zeta, PBL)?  The quantities are then reset by the caller.
struct BoundView {
  double lat0, lat1, p0, p1, dps, dzs, zetas;   // ctl->bound_*
  int pbl;
};
MPB_HD bool bound_applies(const MetView &g, const BoundView &k, const Parcel &a) {
  if (a.lat < k.lat0 || a.lat > k.lat1 || a.p > k.p0 || a.p < k.p1) return false;
  if (k.dps > 0 || k.dzs > 0 || k.zetas > 0 || k.pbl) {
    CellAxes ax;
    axes_reset(ax);
    double ps, pbl;
    surface_at(g, a.time, a.lon, a.lat, ax, ps, pbl);
    if (k.dps > 0 && a.p < ps - k.dps) return false;
    if (k.dzs > 0 && altitude(a.p) > altitude(ps) + k.dzs) return false;
    if (k.zetas > 0) {
      CubeT<true> c;
      cube_reset(c);
      const double t = temperature_at(g, a.time, a.lon, a.lat, a.p, c);
      if (zeta_diagnosed(ps, a.p, t) > k.zetas) return false;
    }
    if (k.pbl && a.p < pbl) return false;
  }
  return true;
}
// clim_ts (396-410)
MPB_HD double series_at(const double *tm, const double *v, int n, double t) {
  if (t <= tm[0]) return v[0];
  if (t >= tm[n - 1]) return v[n - 1];
  const int mid = (n - 1) >> 1;
  const int i = find_interval(tm, n, tm[mid] < tm[mid + 1], t);
  return v[i] + (v[i + 1] - v[i]) / (tm[i + 1] - tm[i]) * (t - tm[i]);   // LIN
}

// module_chem_grid (3885-4054): box of the chemistry grid a parcel is in (or -1), and the volume mixing ratio of a box from
// its mass at the temperature of its centre
struct ChemGrid {
  double lon0, lon1, lat0, lat1, z0, z1, dlon, dlat, dz, t0, t1, tt, molmass;
  int nx, ny, nz;
};
MPB_HD int chem_box(const ChemGrid &k, double time, double lon, double lat, double p) {
  const double zpart = altitude(p);
  if (time < k.t0 || time > k.t1 || lon < k.lon0 || lon >= k.lon1 || lat < k.lat0 || lat >= k.lat1 || zpart < k.z0 || zpart >= k.z1)
    return -1;
  const int ix = (int)((lon - k.lon0) / k.dlon), iy = (int)((lat - k.lat0) / k.dlat), iz = (int)((zpart - k.z0) / k.dz);
  if (ix >= k.nx || iy >= k.ny || iz >= k.nz) return -1;
  return (ix * k.ny + iy) * k.nz + iz;   // ARRAY_3D, src/mptrac.h:709
}
MPB_HD double chem_vmr(const MetView &g, const ChemGrid &k, int box, double mass) {
  const int iz = box % k.nz, iy = (box / k.nz) % k.ny, ix = box / (k.nz * k.ny);
  const double z = k.z0 + k.dz * (iz + 0.5), press = kP0 * exp(-z / kH0);
  const double lon = k.lon0 + k.dlon * (ix + 0.5), lat = k.lat0 + k.dlat * (iy + 0.5);
  const double area = k.dlat * k.dlon * ((kRE * kPi / 180.) * (kRE * kPi / 180.)) * cos(lat * (kPi / 180.0));
  CubeT<true> c;
  cube_reset(c);
  const double temp = temperature_at(g, k.tt, lon, lat, press, c);
  return kMA / k.molmass * mass / ((100. * press / (kRA * temp)) * area * k.dz * 1e9);
}

// module_decay (4227-4263): the e-folding time blends the tropospheric and the stratospheric one with tropo_weight
// (12748-12770); returns exp(-dt / tdec)
MPB_HD double decay_factor(const ClimView &cl, int coord_type, double utm_ref_lat, double tdec_trop, double tdec_strat,
                           const Parcel &a, double dt, double &tdec) {
  const double w = weight_tropo(tropopause_pressure(cl, a.time, coord_type == 0 ? a.lat : utm_ref_lat), a.p);
  tdec = w * tdec_trop + (1 - w) * tdec_strat;
  return exp(-dt / tdec);
}

// ----------------------------------------------------------------------------------------------
// cell key of module_sort (5909-5919): raw index searches, no wrap
// ----------------------------------------------------------------------------------------------
MPB_HD int cell_key(const MetView &g, double lon, double lat, double p) {
  return (lon_interval(g, lon) * g.ny + lat_interval(g, lat)) * g.nz + p_interval(g, p);
}

// altitude of a pressure (src/mptrac.h:2243)

// box index of module_mixing / write_grid (5201-5218, 13844-13860); -1 = outside
// The regular lon x lat x log-pressure-altitude grid of module_mixing, module_chem_grid and write_grid
// (src/mptrac.c:5195-5217, 3951-3972, 13826-13860) with its cell sizes and their reciprocals, set once per launch on the host
// (box_cells): the three per-parcel quotients below are then quotients by grid constants (div_for_index: the reference's
// truncated index for every input, without the IEEE division sequence).
struct BoxGrid {
  double t0, t1, lon0, lon1, lat0, lat1, z0, z1;
  int nx, ny, nz;
  double dlon, dlat, dz, rdlon, rdlat, rdz;
};
MPB_HD void box_cells(BoxGrid &g) {
  g.dlon = (g.lon1 - g.lon0) / g.nx; g.dlat = (g.lat1 - g.lat0) / g.ny; g.dz = (g.z1 - g.z0) / g.nz;
  g.rdlon = 1.0 / g.dlon; g.rdlat = 1.0 / g.dlat; g.rdz = 1.0 / g.dz;
}
// The vertical index needs Z(p) = H0 log(P0 / p) (src/mptrac.h:2243) only to decide a cell.  Production device build: a
// single-precision altitude (|error| < 2e-4 km for |Z| < 1000 km: two roundings of the argument, 1 ulp of logf, one multiply)
// decides whenever it is further than kZMargin
// from the grid's bottom, top and cell faces -- all but ~2e-3 of the parcels on 1 km boxes -- and the double-precision
// sequence (a division, a logarithm, a quotient: ~130 instructions) runs only for the rest.  Same index for every input.
constexpr float kZMargin = 1e-3f;   // [km]
MPB_HD int box_level(const BoxGrid &g, double p) {   // -1: outside [z0, z1) or beyond the last cell
#if MPB_FAST_QUOT
  {
    const float zf = 7.0f * logf(1013.25f / (float)p);                   // (NaN for p <= 0 or non-finite p: falls through)
    const float lo = (float)g.z0, hi = (float)g.z1, m = kZMargin + 1e-6f * (fabsf(lo) + fabsf(hi));
    if (!(fabsf(zf) < 1e3f)) {
      // (outside the range the error bound was derived for, or not a number: the exact sequence decides)
    } else if (zf > lo + m && zf < hi - m) {
      const float t = (zf - lo) * (float)g.rdz;
      const float k = floorf(t), mt = m * (float)g.rdz + 4e-7f * t;       // (margin in cells, plus the rounding of t itself)
      if (t - k > mt && k + 1.0f - t > mt) return k < (float)g.nz ? (int)k : -1;
    } else if (zf < lo - m || zf > hi + m) {
      return -1;
    }
  }
#endif
  const double z = altitude(p);
  if (z < g.z0 || z >= g.z1) return -1;
  const int iz = (int)div_for_index(z - g.z0, g.dz, g.rdz);
  return iz < g.nz ? iz : -1;
}
MPB_HD int box_index(const BoxGrid &g, double time, double lon, double lat, double p) {
  if (time < g.t0 || time > g.t1 || lon < g.lon0 || lon >= g.lon1 || lat < g.lat0 || lat >= g.lat1) return -1;
  const int iz = box_level(g, p);
  if (iz < 0) return -1;
  const int ix = (int)div_for_index(lon - g.lon0, g.dlon, g.rdlon);
  const int iy = (int)div_for_index(lat - g.lat0, g.dlat, g.rdlat);
  if (ix >= g.nx || iy >= g.ny) return -1;
  return (ix * g.ny + iy) * g.nz + iz;
}
MPB_HD int box_index(double time, double lon, double lat, double p, double t0, double t1,
                     double lon0, double lon1, double lat0, double lat1, double z0, double z1,
                     int nx, int ny, int nz) {
  BoxGrid g;
  g.t0 = t0; g.t1 = t1; g.lon0 = lon0; g.lon1 = lon1; g.lat0 = lat0; g.lat1 = lat1; g.z0 = z0; g.z1 = z1;
  g.nx = nx; g.ny = ny; g.nz = nz;
  box_cells(g);
  return box_index(g, time, lon, lat, p);
}

}  // namespace mpb
