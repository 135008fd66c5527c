// quad.cuh -- the step kernel with one LANE per coordinate of a parcel.
//
// A group of G consecutive lanes of a warp owns one parcel: lane 0 follows the longitude and the zonal wind u, lane 1 the
// latitude and v, lane 2 the pressure and omega, and (G = 4, when sedimentation needs it) lane 3 the temperature.  Each lane
//   - searches ITS axis of the met grid and computes ITS interpolation weight; the three weights travel through shuffles;
//   - keeps the 8 corners x 2 time levels of ITS field of the current grid cell in registers, already in the form every
//     lookup needs -- the upper-level value and the fp32 difference "lower - upper", both promoted to fp64 -- so that the
//     f32 -> f64 conversions (F2F.F64.F32 runs on the 16-lane XU pipe: the one-thread-per-parcel kernel's hard floor,
//     192 per RK4 step) happen once per cell fetch instead of once per Runge-Kutta stage;
//   - advances ITS coordinate.
// The arithmetic of every component is, operation for operation, that of the one-thread-per-parcel kernel (physics.cuh,
// step_kernel in engine.cu) -- the strict builds of the two agree bit for bit, the production builds to the last FMA
// contraction (tests/test_gpu_quad.py) -- which in turn follows module_timesteps, module_position, module_advect
// (src/mptrac.c:5999-6042, 5435-5489, 3612-3677) and their intpol_met_time_3d / intpol_met_space_3d lookups (3112-3137,
// 2985-3044).  A third of the live state per thread buys more resident warps; the dependent chain of a stage shrinks
// from 66 interpolation operations to 22.
// MEASURED (B200, C2, RK4): slower than the one-thread-per-parcel kernel, 196 us against 114 us per step: issue
// utilisation rises (38 % -> 53 %) and the XU pipe falls to 15 %, but the warp executes 2.7x the instructions, because
// everything of a stage that is not interpolation -- range check, interval test, weight, position update, the
// bookkeeping around it -- is replicated in every lane (ncu: profiles/r02b_ncu_quad_step_kernel_c2.json).  It is kept as an
// opt-in variant (MPTRAC_B200_STEP=quad) with its tests; the default is step_kernel.
// Restrictions (the dispatcher falls back to step_kernel otherwise): longitude / latitude grids (no Cartesian met data),
// global met domain.
#pragma once

#include "physics.cuh"

namespace mpb {

// fp64 view of one field in one grid cell: per (x, y) column the upper-level value H and the difference D = lower - upper
// (taken in fp32 like the reference, 3023-3038), for both time levels
struct FieldCube {
  double h0[4], d0[4], h1[4], d1[4];   // columns (x0,y0), (x0,y1), (x1,y0), (x1,y1)
};

__device__ __forceinline__ float2 ldg_pair(const float2 *p) { return __ldg(p); }

// the cell (ix, iy, iz) of field `j` (0 u, 1 v, 2 w, 3 T) into the cube: 8 loads of 8 bytes, conversions once
__device__ __forceinline__ void fetch_field(const MetView &g, int j, int ix, int iy, int iz, FieldCube &c) {
  const size_t sy = (size_t)g.nz, sx = (size_t)g.ny * (size_t)g.nz;
  const float2 *b = reinterpret_cast<const float2 *>(g.f + ((size_t)ix * sx + (size_t)iy * sy + (size_t)iz)) + j;
  const size_t col[4] = {0, sy, sx, sx + sy};
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const float2 lo = ldg_pair(b + 4 * col[k]), hi = ldg_pair(b + 4 * (col[k] + 1));   // a node is four {met0, met1} pairs
    c.h0[k] = (double)hi.x; c.d0[k] = (double)f_sub(lo.x, hi.x);
    c.h1[k] = (double)hi.y; c.d1[k] = (double)f_sub(lo.y, hi.y);
  }
}

// trilinear value of one time level: vertical, then latitude, then longitude (3023-3043)
__device__ __forceinline__ double cube_level(double wx, double wy, double wz, const double (&h)[4], const double (&d)[4]) {
  return lerp_f64(wx, lerp_f64(wy, wz * d[0] + h[0], wz * d[1] + h[1]), lerp_f64(wy, wz * d[2] + h[2], wz * d[3] + h[3]));
}

// what a lane needs to follow its coordinate
struct LaneAxis {
  const AxisCell *cells;
  const double *xx;
  int n, asc;
};

template <int G>
struct QuadLane {
  int lane, j, base;
  unsigned gmask;
  bool idle;      // lanes beyond the last whole group of the warp (G = 3: lanes 30, 31)
  __device__ QuadLane() {
    lane = threadIdx.x & 31; j = lane % G; base = lane - j; gmask = (((1u << G) - 1u) << base);
    idle = lane >= (32 / G) * G;
  }
  __device__ __forceinline__ bool all(unsigned ballot) const { return (ballot & gmask) == gmask; }
};

// Stage lookup of a group: every lane brings the (stage) value of ITS coordinate; returns the field value of the lane
// (u, v, w or T at the group's position) for the time weight wt.
// held: the cell the cubes of the group hold (hx < 0: nothing yet); cell / ci: the lane's axis interval and its index.
template <int G>
__device__ __forceinline__ double quad_lookup(const MetView &g, const QuadLane<G> &q, const LaneAxis &ax, double x, double wt,
                                              AxisCell &cell, int &ci, int &hx, int &hy, int &hz, FieldCube &cube) {
  const unsigned full = 0xffffffffu;
  // range check of the lane's coordinate (2755-2803, latitude / longitude grids)
  double x2 = x;
  if (q.j == 0) {
    x2 = wrap360(x);
    if (x2 < g.lon_first) x2 += 360;
    else if (x2 > g.lon_last) x2 -= 360;
  } else if (q.j == 1) {
    x2 = clamp_to(x, g.lat_lo, g.lat_hi);
  }
  // still inside the interval the lane holds?
  bool ok;
#if MPB_FAST_QUOT
  ok = ci >= 0 && cell_holds(cell, ci, ax.n, ax.asc, x2);
#else
  int ireg = 0;
  if (q.j == 0) { ireg = lon_interval(g, x2); ok = ireg == ci; }
  else ok = ci >= 0 && cell_holds(cell, ci, ax.n, ax.asc, x2);
#endif
  if (q.j >= 3 || q.idle) ok = true;
  const bool group_ok = q.all(__ballot_sync(full, ok));
  const unsigned moving = __ballot_sync(full, !group_ok);
  if (!group_ok) {
    if (!ok) {
      if (q.j == 0) {
#if MPB_FAST_QUOT
        int i = (int)((x2 - g.lon_first) * g.r_lon_d);
        ci = i < 0 ? 0 : (i > g.nx - 2 ? g.nx - 2 : i);
#else
        ci = ireg;
#endif
        cell = load_cell(ax.cells + ci);
      } else {
        const int guess = q.j == 1 ? lat_guess(g, x2) : p_guess(g, x2);
        ci = locate_cell(ax.xx, ax.cells, ax.n, ax.asc, x2, guess, cell);
      }
    }
    const int ix = __shfl_sync(moving, ci, q.base), iy = __shfl_sync(moving, ci, q.base + 1), iz = __shfl_sync(moving, ci, q.base + 2);
    if (ix != hx || iy != hy || iz != hz) {
      fetch_field(g, q.j < 3 ? q.j : 3, ix, iy, iz, cube);
      hx = ix; hy = iy; hz = iz;
    }
  }
  const double w = quot(cell.hi - x2, cell.hi - cell.lo, cell.rd);
  const double wx = __shfl_sync(full, w, q.base), wy = __shfl_sync(full, w, q.base + 1), wz = __shfl_sync(full, w, q.base + 2);
  return lerp_f64(wt, cube_level(wx, wy, wz, cube.h0, cube.d0), cube_level(wx, wy, wz, cube.h1, cube.d1));
}

// metres -> coordinate increment of the lane: DX2DEG for the longitude lane (divisor from the latitude the step started at
// or, midpoint scheme, of the stage), DY2DEG for the latitude lane, identity for the pressure lane (src/mptrac.h:904-989)
__device__ __forceinline__ double quad_convert(int j, const LonScale &ks, double d) {
#if MPB_FAST_QUOT
  if (j == 2) return d;
  if (j == 1) return d * kDegPerM;
  return ks.mode == 2 ? d * ks.rd : 0.0;
#else
  if (j == 0) return dx2coord(ks, d);
  if (j == 1) return dy2coord(0, d);
  return d;
#endif
}

// module_position for a group (5435-5489): the canonical test is the lanes' own; the rare rest runs on the lane that owns
// the coordinate with the partner coordinate fetched by shuffle
template <int G>
__device__ __forceinline__ void quad_fix_position(const MetView &g, const QuadLane<G> &q, double time, double ptop, double &x) {
  const unsigned full = 0xffffffffu;
  bool canonical = true;
  if (q.j == 0) canonical = below_360(x) && x >= -180 && x < 180;
  else if (q.j == 1) canonical = x >= -90 && x <= 90;
  const bool group_ok = q.all(__ballot_sync(full, canonical));
  // (both shuffles are executed by every lane: the reflection needs longitude AND latitude in one thread)
  double lon = __shfl_sync(full, x, q.base), lat = __shfl_sync(full, x, q.base + 1);
  if (!group_ok) {
    reflect_at_poles(lon, lat);
    if (q.j == 0) x = lon;
    else if (q.j == 1) x = lat;
  }
  if (q.j == 2) {
    if (x < ptop) {
      x = reflect_p(ptop, x);
    } else if (x > 300.) {
      const float4 s11 = ldg(g.s + (size_t)g.ny + 1);
      const double ps = time_blend_guarded(time_weight(g, time), (double)s11.x, (double)s11.z);
      if (x > ps) x = reflect_p(ps, x);
    }
  }
}

}  // namespace mpb
