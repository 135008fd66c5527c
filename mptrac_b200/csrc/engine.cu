// engine.cu -- device context, kernels and the C ABI (include/mptrac_b200.h) of the B200-native
// MPTRAC time-step engine.  sm_100a only; no CPU fallback: every entry point needs a CUDA device.
//
// Data layout in HBM
//   parcels   SoA fp64: time | p | lon | lat | q[nq], each np_max long (np_max = capacity rounded up to whole 32-parcel
//             tiles, so every array starts 256-byte aligned); dt[np]; uvwp float[np][3]
//             (+ a second copy of the SoA used as the gather target of module_sort, then swapped)
//   met       ONE array of 32-byte nodes {u0,v0,w0,T0,u1,v1,w1,T1} [nx][ny][nz] (z fastest) holding both bracketing
//             time levels, so a stencil corner is one aligned 256-bit load (LDG.E.256) = one DRAM/L2 sector;
//             float4 {ps0,pbl0,ps1,pbl1} [nx][ny]; per axis one array of 32-byte interval records
//             {x[i], x[i+1], x[i+1]-x[i], 1/(x[i+1]-x[i])} + a first-guess table for the pressure search
//   clim      tropopause table fp64 [ntime][nlat] + its axes
//
// One fused, persistent kernel per model step does timesteps -> position -> advect -> diff_turb -> diff_meso ->
// sedi -> position for one parcel per thread at a time (the order of mptrac_run_timestep, src/mptrac.c:7877-7919);
// every warp walks the parcels 32 at a time with the next chunk's state staged asynchronously in shared memory.
// module_meteo (meteo_kernel), the cell sort and the box reductions are separate launches.
#include <cub/device/device_radix_sort.cuh>
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../../include/mptrac_b200.h"
#include "met_tables.hpp"
#include "physics.cuh"
#include "quad.cuh"

using namespace mpb;

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_err;

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess)                                                                        \
      throw std::runtime_error(std::string(#call) + ": " + cudaGetErrorString(e_));               \
  } while (0)

#define REQUIRE(cond, msg)                                                                        \
  do {                                                                                            \
    if (!(cond)) throw std::runtime_error(msg);                                                   \
  } while (0)

#define API_BEGIN try {
#define API_END                                                                                   \
  return 0;                                                                                       \
  }                                                                                               \
  catch (const std::exception &ex) {                                                              \
    g_err = ex.what();                                                                            \
    return 1;                                                                                     \
  }

// ------------------------------------------------------------------------------------------------
// kernels
// ------------------------------------------------------------------------------------------------
struct StepArgs {
  MetView met;
  ClimView clim;
  CtlView ctl;
  double *time, *lon, *lat, *p, *dt;
  float *uvwp;
  const double *rp, *rhop;
  // host-resident stepping without copy engines: parcels are read from and written back to mapped (pinned) host memory
  // by the kernel itself; the device arrays above still receive the result.  Null = parcels live on the device.
  const double *in_time, *in_lon, *in_lat, *in_p;
  double *host_time, *host_lon, *host_lat, *host_p;
  // host-resident parcels that all carry the SAME time (the common case): the time array does not cross the link at all --
  // every parcel starts at time_in, and the host fills its own array with the common result (mpb_run_timestep_host)
  double time_in;
  int uniform_time;
  long long np;
  long long ig0;  // global index of local parcel 0
  const int *slot; // array slot (the reference's index) of the parcel at each position, or null = the position itself:
                   // the random numbers of a parcel belong to its slot (src/mptrac.c:5797-5826), see do_sort
  unsigned modules;
};

constexpr unsigned PHYS_TURB = 1, PHYS_MESO = 2, PHYS_SEDI = 4;
// launch shape of the step kernel (compile-time so that variants can be swept: scripts/sweep_variants.sh)
#ifndef MPB_BLOCK
#define MPB_BLOCK 128
#endif
#ifndef MPB_MINBLOCKS
#define MPB_MINBLOCKS 4
#endif
#ifndef MPB_CUBE_F64
#define MPB_CUBE_F64 0
#endif
#ifndef MPB_TMA_STAGE    // 1: bulk copies (TMA, UBLKCP + mbarrier) stage the parcel stream of device-resident parcels;
#define MPB_TMA_STAGE 0  // 0: per-lane cp.async (LDGSTS) only.  Measured on B200 (C2): 117.5 us per step with cp.async, 120-123 us
#endif                   // with bulk copies (four copies serialised in one lane + mbarrier polling against four issued by all lanes at
                         // once), and 120 us with both paths compiled in (the unrolled RK4 loop sits at the instruction-cache edge:
                         // stall_no_instruction doubles).  The strict library is built with 1, so both paths are in the GPU test suite.
#ifndef MPB_ROLL_SEDI     // 1: the Runge-Kutta stage loop stays rolled in the kernels that also carry sedimentation.  Measured
#define MPB_ROLL_SEDI 1   // (same box, rolled against unrolled): configs[2] (0.5 deg grid, mesoscale diffusion + sedimentation)
#endif                    // 2.496 against 2.635 ms; the same modules on the 1 deg grid 2.690 against 2.670; rolled in the
                          // turbulent + mesoscale kernel: configs[3] 2.727 against 2.648, on the 0.5 deg grid 2.443 against 2.408;
                          // advection alone: no difference.  profiles/r02o_roll_stage_loop.txt
#ifndef MPB_PERSIST      // 1: grid = SMs x resident blocks, threads loop over parcels; 0: one block per kBlock parcels
#define MPB_PERSIST 1    // (measured: 126 us vs 135 us)
#endif
constexpr int kBlock = MPB_BLOCK;
constexpr int kLanes = 4;                       // concurrent chunk pipelines of mpb_run_timestep_host
constexpr long long kHostChunkMin = 16384;      // parcels per chunk: at least 128 KiB per array ...
constexpr long long kHostChunkMax = 1 << 20;    // ... at most 8 MiB

#ifndef MPB_LEVEL_MINBLOCKS
#define MPB_LEVEL_MINBLOCKS 16  // resident blocks per SM the model-level advection kernel is compiled for
#endif
#ifndef MPB_LEVEL_BLOCK
#define MPB_LEVEL_BLOCK 32      // ... and its block size: one warp, so that a finished warp is replaced at once (measured 128 / 64 / 32
                                // threads: 0.290 / 0.284 / 0.281 ms per step of c2ml, profiles/r02p_sweep_level_near.jsonl)
#endif
#ifdef MPB_NO_BOUNDS   // register budget given by -maxrregcount instead (variant sweeps)
#define MPB_BOUNDS
#else
#define MPB_BOUNDS __launch_bounds__(kBlock, MPB_MINBLOCKS)
#endif

// One parcel, one model step: timesteps -> position -> advect -> diff_turb -> diff_meso -> sedi -> position, in registers.
// First part: the time step of the parcel (false = nothing to do, PARTICLE_LOOP(check_dt = 1), src/mptrac.h:1754-1759) and the
// first position check.
__device__ __forceinline__ bool parcel_begin(const StepArgs &A, long long ip, Parcel &a, double &dt) {
  if (A.modules & MOD_TIMESTEPS) {
    dt = parcel_dt(A.met, A.ctl, a);
    if (A.modules & MOD_STORE_DT) A.dt[ip] = dt;
  } else {
    dt = A.dt[ip];
  }
  if (dt == 0) {
    if (A.in_time) { A.time[ip] = a.time; A.lon[ip] = a.lon; A.lat[ip] = a.lat; A.p[ip] = a.p; }   // keep the device mirror
    return false;
  }
  if (A.modules & MOD_POS_PRE) fix_position(A.met, a);
  return true;
}

// Second part: the modules that look the met data up, the final position check, the stores.  `cube` = the met cell the
// parcel sits in, shared by every lookup of the step.
template <int ADVECT, unsigned PHYS>
__device__ __forceinline__ void parcel_finish(const StepArgs &A, long long ip, Parcel &a, double dt, CubeT<!(PHYS & PHYS_MESO)> &cube) {
  const unsigned long long ig = (unsigned long long)(A.ig0 + (A.slot ? A.slot[ip] : ip));
#if MPB_CUBE_F64
  if (ADVECT > 0) {
    WindCube wc;
    cube_reset(wc);
    advect<(ADVECT > 0 ? ADVECT : 1)>(A.met, dt, a, wc);
  }
#else
  if (ADVECT > 0) advect<(ADVECT > 0 ? ADVECT : 1), (PHYS & PHYS_SEDI) != 0 && (MPB_ROLL_SEDI != 0)>(A.met, dt, a, cube);
#endif
  if (PHYS & PHYS_TURB) diffuse_turbulent(A.met, A.clim, A.ctl, dt, ig, a);
  if constexpr ((PHYS & PHYS_MESO) != 0) {
    float *s = A.uvwp + 3 * ip;
    float up = s[0], vp = s[1], wp = s[2];
    diffuse_mesoscale(A.met, A.ctl, dt, ig, a, up, vp, wp, cube);
    s[0] = up; s[1] = vp; s[2] = wp;
  }
  if (PHYS & PHYS_SEDI) sediment(A.met, dt, A.rp[ip], A.rhop[ip], a, cube);
  if (A.modules & MOD_POS_POST) fix_position(A.met, a);

  if (ADVECT > 0 || A.in_time) A.time[ip] = a.time;
  A.lon[ip] = a.lon;
  A.lat[ip] = a.lat;
  A.p[ip] = a.p;
  if (A.host_lon) {
    if (ADVECT > 0 && A.host_time) A.host_time[ip] = a.time;
    A.host_lon[ip] = a.lon;
    A.host_lat[ip] = a.lat;
    A.host_p[ip] = a.p;
  }
}

template <int ADVECT, unsigned PHYS>
__device__ __forceinline__ void step_parcel(const StepArgs &A, long long ip, Parcel a) {
  double dt;
  if (!parcel_begin(A, ip, a, dt)) return;
  CubeT<!(PHYS & PHYS_MESO)> cube;
  cube_reset(cube);
  parcel_finish<ADVECT, PHYS>(A, ip, a, dt, cube);
}

// ------------------------------------------------------------------------------------------------
// Staging of the parcel stream.  The four state arrays are contiguous SoA streams, and a warp consumes them 32 parcels
// (256 bytes per array) at a time: exactly what the bulk-copy engine (TMA, cp.async.bulk -> UBLKCP) moves with ONE
// instruction per array, into shared memory, with no destination registers and no per-lane address arithmetic.  That
// matters here because the kernel holds ~100 live fp64 registers per parcel: a register prefetch of the next parcel is
// spilled at once and the spill waits for the load (r01e profile: 7 % of all stall samples).  Every warp owns two
// 4 x 256-byte slots and one mbarrier per slot; lane 0 arms the barrier with the byte count and issues the four copies
// for chunk k+1 while the warp computes chunk k.  Warps never wait for each other.
// The alternative path is per-lane cp.async (LDGSTS, 8 bytes per lane and array): it is what host-mapped sources
// (zero-copy stepping of pinned host arrays) must use -- a bulk copy reads whole 256-byte tiles, which is only safe
// inside the context's own padded allocations -- and, measured, also the faster one on this kernel (MPB_TMA_STAGE).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  unsigned done = 0;
  while (!done)
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(done) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

constexpr int kWarps = kBlock / 32;
constexpr unsigned kTileBytes = 32 * sizeof(double);   // one array's share of a warp's chunk

// Persistent form: the grid is sized to what the GPU holds at once (SMs x resident blocks) and every warp walks the
// chunks w0, w0 + stride, ... of 32 parcels.  While chunk k computes, the state of chunk k+1 is already in flight, the
// axis tables stay hot in L1 for the whole launch and no block-launch gaps separate the chunks of a warp (only ~16
// warps are resident per SM, so every exposed latency counts).
template <int ADVECT, unsigned PHYS>
__global__ void MPB_BOUNDS step_kernel(const __grid_constant__ StepArgs A) {
  __shared__ alignas(128) double stage[2][4][kBlock];
  __shared__ alignas(8) unsigned long long full[kWarps][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long stride = (long long)gridDim.x * kBlock;
  long long w0 = (long long)blockIdx.x * kBlock + warp * 32;   // first parcel of this warp's chunk
  if (w0 >= A.np) return;                                        // (warp-uniform)
#if MPB_TMA_STAGE
  const bool tma = A.in_time == nullptr;   // device-resident parcels: bulk copies (warp-uniform: a launch parameter)
#else
  constexpr bool tma = false;
#endif

  auto issue = [&](long long first, int b) {   // start the copies of chunk [first, first + 32) into slot b
    if (tma) {
      if (lane == 0) {
        const unsigned bar = smem_u32(&full[warp][b]), dst = smem_u32(&stage[b][0][warp * 32]);
        mbar_expect(bar, 4 * kTileBytes);
        bulk_load(dst, A.time + first, kTileBytes, bar);
        bulk_load(dst + kBlock * 8, A.lon + first, kTileBytes, bar);
        bulk_load(dst + 2 * kBlock * 8, A.lat + first, kTileBytes, bar);
        bulk_load(dst + 3 * kBlock * 8, A.p + first, kTileBytes, bar);
      }
    } else {
      if (first + lane < A.np) {
        const unsigned dst = smem_u32(&stage[b][0][threadIdx.x]);
        const long long i = first + lane;
        const bool host = A.in_time != nullptr;
        if (!A.uniform_time)
          asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"((host ? A.in_time : A.time) + i) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + kBlock * 8), "l"((host ? A.in_lon : A.lon) + i) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 2 * kBlock * 8), "l"((host ? A.in_lat : A.lat) + i) : "memory");
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst + 3 * kBlock * 8), "l"((host ? A.in_p : A.p) + i) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };

  if (tma) {
    if (lane == 0) {
      mbar_init(smem_u32(&full[warp][0]));
      mbar_init(smem_u32(&full[warp][1]));
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
  }
  int buf = 0;
  unsigned phase = 0;   // bit b = parity the next completion of slot b's barrier will have
  issue(w0, 0);
  for (;;) {
    if (tma) { mbar_wait(smem_u32(&full[warp][buf]), (phase >> buf) & 1u); phase ^= 1u << buf; }
    else asm volatile("cp.async.wait_group 0;" ::: "memory");   // a lane only reads the slots it filled itself
    const long long ip = w0 + lane;
    Parcel a;
    a.time = A.uniform_time ? A.time_in : stage[buf][0][threadIdx.x]; a.lon = stage[buf][1][threadIdx.x];
    a.lat = stage[buf][2][threadIdx.x]; a.p = stage[buf][3][threadIdx.x];
    w0 += stride;
    const bool more = w0 < A.np;          // (warp-uniform)
    buf ^= 1;
    __syncwarp();                         // every lane has read its slot of the buffer the next-but-one copy reuses
    if (more) issue(w0, buf);
    if (ip < A.np) step_parcel<ADVECT, PHYS>(A, ip, a);
    if (!more) break;
    __syncwarp();                         // (reconverge before the next wait: measured 120 us with, 125 us without)
  }
}

// ------------------------------------------------------------------------------------------------
// The lane-per-coordinate step kernel (quad.cuh), a measured variant (MPTRAC_B200_STEP=quad): timesteps -> position ->
// advect -> position for groups of G lanes.  Persistent: every warp walks the parcels 32 / G at a time.  The diffusion
// modules and sedimentation stay in step_kernel (a second launch reading dt from memory, MPTRAC_B200_QUAD_SPLIT=1).
// ------------------------------------------------------------------------------------------------
#ifndef MPB_QUAD_MINBLOCKS
#define MPB_QUAD_MINBLOCKS 6      // 128-thread blocks per SM the kernel is compiled for (85 registers per thread)
#endif
template <int ORDER, int G>
__global__ void __launch_bounds__(128, MPB_QUAD_MINBLOCKS) quad_step_kernel(const __grid_constant__ StepArgs A) {
  const MetView &g = A.met;
  const QuadLane<G> q;
  constexpr int kPerWarp = 32 / G;
  const unsigned full = 0xffffffffu;
  const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
  const long long warp0 = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  LaneAxis ax;
  ax.cells = q.j == 0 ? g.lonc : (q.j == 1 ? g.latc : g.pc);
  ax.xx = q.j == 0 ? g.lon : (q.j == 1 ? g.lat : g.p);
  ax.n = q.j == 0 ? g.nx : (q.j == 1 ? g.ny : g.nz);
  ax.asc = q.j == 0 ? g.lon_asc : (q.j == 1 ? g.lat_asc : g.p_asc);
  // the lane's coordinate array: read from the host mapping when the parcels live there (zero-copy stepping)
  const double *src = A.in_time ? (q.j == 0 ? A.in_lon : (q.j == 1 ? A.in_lat : A.in_p)) : (q.j == 0 ? A.lon : (q.j == 1 ? A.lat : A.p));
  double *dst = q.j == 0 ? A.lon : (q.j == 1 ? A.lat : A.p);
  double *hdst = q.j == 0 ? A.host_lon : (q.j == 1 ? A.host_lat : A.host_p);
  const double *tsrc = A.in_time ? A.in_time : A.time;
  const double ptop = ldg(g.p + g.nz - 1);
  const bool owner = q.j < 3 && q.lane < kPerWarp * G;      // this lane owns a coordinate of a parcel

  for (long long first = warp0 * kPerWarp; first < A.np; first += nwarps * kPerWarp) {   // (warp-uniform)
    const long long ip = first + q.lane / G;
    const bool valid = owner && ip < A.np;
    const double x_in = valid ? src[ip] : (q.j == 2 ? 500.0 : 0.0);
    double x = x_in;
    double time = A.uniform_time ? A.time_in : ((ip < A.np) ? tsrc[ip] : 0.0);
    double dt;
    if (A.modules & MOD_TIMESTEPS) {
      Parcel a;
      a.time = time; a.lon = a.lat = a.p = 0;      // (global met domain: the position does not enter, 5999-6042)
      dt = parcel_dt(g, A.ctl, a);
      if ((A.modules & MOD_STORE_DT) && valid && q.j == 0) A.dt[ip] = dt;
    } else {
      dt = ip < A.np ? A.dt[ip] : 0.0;
    }
    const bool active = valid && dt != 0;             // PARTICLE_LOOP(check_dt = 1), src/mptrac.h:1754-1759
    if (A.modules & MOD_POS_PRE) quad_fix_position(g, q, time, ptop, x);

    if (ORDER > 0) {
      AxisCell cell;
      cell.lo = cell.hi = cell.d = cell.rd = 0;
      int ci = -1, hx = -1, hy = -1, hz = -1;
      FieldCube cube;
      // the metre -> degree divisor of the longitude lane belongs to the latitude the step starts at (3659-3673)
      const double lat0 = __shfl_sync(full, x, q.base + 1);
      const LonScale ks = lon_scale(0, lat0);
      double acc = 0, f = 0, wt = 0, lat_stage = lat0;
#pragma unroll
      for (int i = 0; i < ORDER; i++) {
        double xs = x, dts = 0.0;
        if (i > 0) {
          dts = (i == 3 ? 1.0 : 0.5) * dt;
          xs = x + quad_convert(q.j, ks, dts * f);
        }
        if (ORDER == 2) lat_stage = __shfl_sync(full, xs, q.base + 1);
        if (i != 2) wt = time_weight(g, time + dts);
        f = quad_lookup(g, q, ax, xs, wt, cell, ci, hx, hy, hz, cube);
        double k = 1.0;
        if (ORDER == 2) k = (i == 0 ? 0.0 : 1.0);
        else if (ORDER == 4) k = (i == 0 || i == 3 ? 1.0 / 6.0 : 2.0 / 6.0);
        acc += k * f;
      }
      time += dt;
      // (midpoint: the final longitude update takes the stage latitude, 3672-3673)
      x += quad_convert(q.j, ORDER == 2 ? lon_scale(0, lat_stage) : ks, dt * acc);
    }
    if (A.modules & MOD_POS_POST) quad_fix_position(g, q, time, ptop, x);

    if (active) {
      dst[ip] = x;
      if (hdst) hdst[ip] = x;
      if (q.j == 0 && (ORDER > 0 || A.in_time)) A.time[ip] = time;
      if (q.j == 0 && ORDER > 0 && A.host_time) A.host_time[ip] = time;
    } else if (valid && A.in_time) {
      dst[ip] = x_in;                                 // keep the device mirror of host-resident parcels complete
      if (q.j == 0) A.time[ip] = time;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// The TMA-tile form of the step kernel (MPTRAC_B200_STEP=tile), a measured variant.  One block owns 128 consecutive parcels.
// After a cell sort they sit in a few neighbouring grid columns, so the block (1) finds the cell of each of its parcels, (2)
// takes the minimum cell indices as the origin of a window of TX x TY x TZ met nodes, (3) has ONE thread fetch that window
// into shared memory with a single bulk tensor copy (cp.async.bulk.tensor.4d over the node array seen as [nx][ny][nz][8
// floats]; completion on an mbarrier), and (4) runs the step with every cube fetch inside the window served from shared
// memory (two 16-byte loads per node) instead of a 32-byte gather from L1 / L2 / HBM; cells outside the window -- parcels
// that drifted since the sort, very sparse blocks -- still take the global path.
// ------------------------------------------------------------------------------------------------
struct TileShape {
  int nx, ny, nz;
};
template <int ADVECT, unsigned PHYS>
__global__ void __launch_bounds__(128, 4) tile_step_kernel(const __grid_constant__ StepArgs A, const __grid_constant__ CUtensorMap tmap,
                                                           const TileShape shape) {
  extern __shared__ __align__(128) unsigned char tile_mem[];
  __shared__ TileRef tref;
  __shared__ alignas(8) unsigned long long bar;
  __shared__ int lo[3];
  const MetView &g = A.met;
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (threadIdx.x == 0) {
    lo[0] = lo[1] = lo[2] = 0x7fffffff;
    mbar_init(smem_u32(&bar));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  Parcel a;
  double dt = 0;
  bool active = false;
  CubeT<!(PHYS & PHYS_MESO)> cube;
  cube_reset(cube);
  if (ip < A.np) {
    a.time = A.time[ip]; a.lon = A.lon[ip]; a.lat = A.lat[ip]; a.p = A.p[ip];
    active = parcel_begin(A, ip, a, dt);
  }
  // the cell the parcel starts the step in (the searches leave their intervals in the cube's axis cache)
  int ix = 0x7fffffff, iy = 0x7fffffff, iz = 0x7fffffff;
  if (active) {
    double lon2, lat2;
    clamp_horizontal(g, a.lon, a.lat, lon2, lat2);
    ix = lon_cell(g, lon2, cube.ax);
    iy = lat_cell(g, lat2, cube.ax);
    iz = p_cell(g, a.p, cube.ax);
  }
  const unsigned full = 0xffffffffu;
  int mx = ix, my = iy, mz = iz;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    mx = min(mx, __shfl_xor_sync(full, mx, d)); my = min(my, __shfl_xor_sync(full, my, d)); mz = min(mz, __shfl_xor_sync(full, mz, d));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&lo[0], mx); atomicMin(&lo[1], my); atomicMin(&lo[2], mz); }
  __syncthreads();
  const bool any = lo[0] != 0x7fffffff;      // (block-uniform)
  if (any && threadIdx.x == 0) {
    const int x0 = lo[0], y0 = lo[1], z0 = max(0, min(lo[2], g.nz - shape.nz));
    tref.nodes = reinterpret_cast<const Node *>(tile_mem);
    tref.x0 = x0; tref.y0 = y0; tref.z0 = z0; tref.nx = shape.nx; tref.ny = shape.ny; tref.nz = shape.nz;
    const unsigned bytes = (unsigned)(shape.nx * shape.ny * shape.nz) * (unsigned)sizeof(Node);
    mbar_expect(smem_u32(&bar), bytes);
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(smem_u32(tile_mem)), "l"(&tmap), "r"(0), "r"(z0), "r"(y0), "r"(x0), "r"(smem_u32(&bar)) : "memory");
  }
  __syncthreads();                             // tref is visible
  if (!any) return;
  mbar_wait(smem_u32(&bar), 0);
  if (active) {
    cube.tile = &tref;
    Stencil s;
    s.ix = ix; s.iy = iy; s.iz = iz; s.wx = s.wy = s.wz = 0;
    fetch_cube(g, s, cube);                    // the axis cache names this cell: the cube must hold it (invariant of locate)
    parcel_finish<ADVECT, PHYS>(A, ip, a, dt, cube);
  }
}

typedef void (*tile_fn)(const StepArgs, const CUtensorMap, const TileShape);
template <int ADVECT>
static tile_fn pick_tile_phys(unsigned phys) {
  switch (phys) {
    case 0: return tile_step_kernel<ADVECT, 0>;
    case 1: return tile_step_kernel<ADVECT, 1>;
    case 2: return tile_step_kernel<ADVECT, 2>;
    case 3: return tile_step_kernel<ADVECT, 3>;
    case 4: return tile_step_kernel<ADVECT, 4>;
    case 5: return tile_step_kernel<ADVECT, 5>;
    case 6: return tile_step_kernel<ADVECT, 6>;
    default: return tile_step_kernel<ADVECT, 7>;
  }
}
static tile_fn pick_tile(int advect, unsigned phys) {
  switch (advect) {
    case 1: return pick_tile_phys<1>(phys);
    case 2: return pick_tile_phys<2>(phys);
    case 4: return pick_tile_phys<4>(phys);
    default: throw std::runtime_error("the tile form needs ADVECT 1, 2 or 4");
  }
}

typedef void (*step_fn)(const StepArgs);

template <int ADVECT>
static step_fn pick_phys(unsigned phys) {
  switch (phys) {
    case 0: return step_kernel<ADVECT, 0>;
    case 1: return step_kernel<ADVECT, 1>;
    case 2: return step_kernel<ADVECT, 2>;
    case 3: return step_kernel<ADVECT, 3>;
    case 4: return step_kernel<ADVECT, 4>;
    case 5: return step_kernel<ADVECT, 5>;
    case 6: return step_kernel<ADVECT, 6>;
    default: return step_kernel<ADVECT, 7>;
  }
}

static step_fn pick_step(int advect, unsigned phys) {
  switch (advect) {
    case 0: return pick_phys<0>(phys);
    case 1: return pick_phys<1>(phys);
    case 2: return pick_phys<2>(phys);
    case 4: return pick_phys<4>(phys);
    default: throw std::runtime_error("ADVECT must be 0, 1, 2 or 4");
  }
}

// bounds of read_met_bin_3d (src/mptrac.c:9168-9178): values below / above are set to the bound (NaN stays NaN)
__global__ void clamp_field_kernel(float *f, size_t n, float lo, float hi) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float v = f[i];
  if (v < lo) f[i] = lo;
  else if (v > hi) f[i] = hi;
}

// write four dense fields [n] into time-level `slot` (0 / 1) of the met nodes {u0,u1, v0,v1, w0,w1, t0,t1}
__global__ void pack_met_nodes_kernel(const float *u, const float *v, const float *w, const float *t, float *nodes, int slot, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float *o = nodes + 8 * i + slot;
  o[0] = u[i]; o[2] = v[i]; o[4] = w[i]; o[6] = t ? t[i] : 0.f;
}
// the same for records that keep the four values of a time level together (model-level records {h,a,b,c}0 {h,a,b,c}1)
__global__ void pack_nodes_kernel(const float *u, const float *v, const float *w, const float *t,
                                  float4 *nodes, int slot, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) nodes[2 * i + slot] = make_float4(u[i], v[i], w[i], t ? t[i] : 0.f);
}
__global__ void pack_surface_kernel(const float *ps, const float *pbl, float2 *surf, int slot, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) surf[2 * i + slot] = make_float2(ps ? ps[i] : 0.f, pbl ? pbl[i] : 0.f);
}
// exchange the two time levels in place (mptrac_get_met's pointer swap, src/mptrac.c:6489-6491)
__global__ void swap_levels_kernel(float4 *nodes, size_t n, float2 *surf, size_t ncol) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float4 a = nodes[2 * i], b = nodes[2 * i + 1]; nodes[2 * i] = b; nodes[2 * i + 1] = a; }
  if (i < ncol) { const float2 a = surf[2 * i], b = surf[2 * i + 1]; surf[2 * i] = b; surf[2 * i + 1] = a; }
}

// (dt_out: module_timesteps folded into this pass -- the reference computes dt BEFORE it permutes the parcels and leaves
// cache->dt in slot order, src/mptrac.c:7877-7881, so on a sort step dt cannot come from the fused launch after the sort)
__global__ void sort_keys_kernel(MetView met, CtlView ctl, const double *time, const double *lon, const double *lat, const double *p,
                                 int *keys, int *idx, double *dt_out, long long np) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  Parcel a;
  a.lon = lon[ip]; a.lat = lat[ip]; a.p = p[ip];
  keys[ip] = cell_key(met, a.lon, a.lat, a.p);
  idx[ip] = (int)ip;
  if (dt_out) {
    a.time = time[ip];
    dt_out[ip] = parcel_dt(met, ctl, a);
  }
}

// dst[a][ip] = src[a][perm[ip]] for narr arrays spaced `stride` doubles apart
__global__ void gather_kernel(const double *__restrict__ src, double *__restrict__ dst,
                              const int *__restrict__ perm, long long np, long long stride, int narr) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  const int j = perm[ip];
  for (int a = 0; a < narr; a++) dst[a * stride + ip] = src[a * stride + j];
}

// ---- the engine's own parcel order (do_sort) ----
// the cell a parcel's lookups fall into: longitude wrapped and clamped as the interpolation does it (module_sort's own key
// takes the raw longitude, src/mptrac.c:5905)
__device__ __forceinline__ int lookup_cell_key(const MetView &g, double lon, double lat, double p) {
  double lon2, lat2;
  clamp_horizontal(g, lon, lat, lon2, lat2);
  return (lon_interval(g, lon2) * g.ny + lat_interval(g, lat2)) * g.nz + p_interval(g, p);
}
__global__ void invert_kernel(const int *__restrict__ slot, int *__restrict__ inv, long long np) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < np) inv[slot[i]] = (int)i;
}
// both keys of the parcel at every position, one streaming pass: module_sort's (raw longitude) and the engine's (the cell the
// lookups fall into); dt when module_timesteps rides along (the reference computes it before it permutes, src/mptrac.c:7877-7881)
__global__ void two_keys_kernel(MetView met, CtlView ctl, const double *time, const double *lon, const double *lat, const double *p,
                                int *key_ref, int *key_own, double *dt_out, long long np) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  Parcel a;
  a.lon = lon[i]; a.lat = lat[i]; a.p = p[i];
  key_ref[i] = cell_key(met, a.lon, a.lat, a.p);
  key_own[i] = lookup_cell_key(met, a.lon, a.lat, a.p);
  if (dt_out) {
    a.time = time[i];
    dt_out[i] = parcel_dt(met, ctl, a);
  }
}
// keys[s] = key_at[at[s]] (at = null: the identity), vals[s] = at[s] or s
__global__ void pick_keys_kernel(const int *key_at, const int *at, int *keys, int *vals, bool vals_are_positions, long long np) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= np) return;
  const int a = at ? at[s] : (int)s;
  keys[s] = key_at[a];
  vals[s] = vals_are_positions ? a : (int)s;
}
// position i of the new order holds the parcel of (new) slot order[i], which sits at position from[order[i]] now
__global__ void compose_kernel(const int *order, const int *from, int *slot_new, int *src, long long np) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= np) return;
  const int s = order[i];
  slot_new[i] = s;
  src[i] = from[s];
}
// What belongs to the SLOT in the reference (cache_t::uvwp, ::dt, ::iso_var: module_sort leaves them where they are,
// src/mptrac.c:5944-5949) travels with the parcel here and is handed over at a sort: the parcel that takes slot s gets what
// the previous holder of slot s had.  held_at[s] = position of that previous holder (null: position s).  The level hint
// belongs to the parcel (src[i] = its old position).
struct StateArgs {
  const int *slot_new, *held_at, *src;
  const float *uvwp_old; float *uvwp_new;
  const double *dt_old, *dt_of_slot; double *dt_new;
  const double *iso_old; double *iso_new;
  const unsigned short *hint_old; unsigned short *hint_new;
  long long np;
};
__global__ void hand_over_kernel(const StateArgs A) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.np) return;
  const int s = A.slot_new ? A.slot_new[i] : (int)i;
  const int a = A.held_at ? A.held_at[s] : s;
  A.uvwp_new[3 * i] = A.uvwp_old[3 * (long long)a]; A.uvwp_new[3 * i + 1] = A.uvwp_old[3 * (long long)a + 1];
  A.uvwp_new[3 * i + 2] = A.uvwp_old[3 * (long long)a + 2];
  A.dt_new[i] = A.dt_of_slot ? A.dt_of_slot[s] : A.dt_old[a];
  if (A.iso_old) A.iso_new[i] = A.iso_old[a];
  if (A.hint_old) A.hint_new[i] = A.hint_old[A.src ? A.src[i] : a];
}

// module_meteo for the quantities the resident fields give (src/mptrac.c:5062-5165): every parcel, dt or not
struct MeteoArgs {
  MetView met;
  const double *time, *lon, *lat, *p;
  double *q;            // [nq][q_stride]
  long long q_stride, np;
  int qnt[MPB_METEO_SLOTS];
};

__global__ void __launch_bounds__(128) meteo_kernel(const __grid_constant__ MeteoArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  Parcel a;
  a.time = A.time[ip]; a.lon = A.lon[ip]; a.lat = A.lat[ip]; a.p = A.p[ip];
  CubeT<true> cube;
  cube_reset(cube);
  MeteoValues m;
  meteo_at(A.met, a, cube, m);
  auto set = [&](int slot, double v) { if (A.qnt[slot] >= 0) A.q[(long long)A.qnt[slot] * A.q_stride + ip] = v; };
  set(MPB_Q_PS, m.ps);
  set(MPB_Q_PBL, m.pbl);
  set(MPB_Q_P, a.p);
  set(MPB_Q_T, m.t);
  set(MPB_Q_RHO, 100. * a.p / (kRA * m.t));                 // RHO, src/mptrac.h:1961
  set(MPB_Q_U, m.u);
  set(MPB_Q_V, m.v);
  set(MPB_Q_W, m.w);
  set(MPB_Q_VH, sqrt(m.u * m.u + m.v * m.v));
  set(MPB_Q_VZ, -1e3 * kH0 / a.p * m.w);
  if (A.qnt[MPB_Q_PSAT] >= 0) set(MPB_Q_PSAT, saturation_pressure(m.t));
  if (A.qnt[MPB_Q_PSICE] >= 0) set(MPB_Q_PSICE, saturation_pressure_ice(m.t));
  if (A.qnt[MPB_Q_THETA] >= 0) set(MPB_Q_THETA, potential_temperature(a.p, m.t));
  if (A.qnt[MPB_Q_ZETA_D] >= 0) set(MPB_Q_ZETA_D, zeta_diagnosed(m.ps, a.p, m.t));
}

// module_meteo, the quantities that need further met fields (INTPOL_TIME_ALL's other 9 3-D and 22 2-D fields) and the
// ones derived from temperature and water vapour: same stencil and time weight as meteo_kernel, fields that were not
// asked for are not touched
struct MeteoFieldArgs {
  MetView met;
  const double *time, *lon, *lat, *p;
  double *q;
  long long q_stride, np;
  const float2 *x2[MPB_NX2], *x3[MPB_NX3];   // null = not wanted
  int qnt[MPB_METEO_SLOTS];
  int moist;                                 // any of PW .. TICE wanted
};

__global__ void __launch_bounds__(128) meteo_fields_kernel(const __grid_constant__ MeteoFieldArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  Parcel a;
  a.time = A.time[ip]; a.lon = A.lon[ip]; a.lat = A.lat[ip]; a.p = A.p[ip];
  CubeT<true> cube;
  cube_reset(cube);
  Stencil s;
  locate(A.met, a.lon, a.lat, a.p, cube, s);
  const double wt = time_weight(A.met, a.time);
  auto set = [&](int slot, double v) { A.q[(long long)A.qnt[slot] * A.q_stride + ip] = v; };
#pragma unroll 1
  for (int f = 0; f < MPB_NX2; f++)
    if (A.x2[f] && A.qnt[MPB_Q_TS + f] >= 0) set(MPB_Q_TS + f, field2_at(A.met, A.x2[f], s, wt));
#pragma unroll 1
  for (int f = 0; f < MPB_NX3; f++)
    if (A.x3[f] && A.qnt[MPB_Q_ZG + f] >= 0) set(MPB_Q_ZG + f, field3_at(A.met, A.x3[f], s, wt));
  if (A.moist) {
    const double t = lerp_f64(wt, MPB_TRILERP_OF(cube, s, t0), MPB_TRILERP_OF(cube, s, t1));
    MoistValues w;
    moist_at(a.p, t, field3_at(A.met, A.x3[MPB_F3_H2O], s, wt), w);
    if (A.qnt[MPB_Q_PW] >= 0) set(MPB_Q_PW, w.pw);
    if (A.qnt[MPB_Q_SH] >= 0) set(MPB_Q_SH, w.sh);
    if (A.qnt[MPB_Q_RH] >= 0) set(MPB_Q_RH, w.rh);
    if (A.qnt[MPB_Q_RHICE] >= 0) set(MPB_Q_RHICE, w.rhice);
    if (A.qnt[MPB_Q_TVIRT] >= 0) set(MPB_Q_TVIRT, w.tvirt);
    if (A.qnt[MPB_Q_LAPSE] >= 0) set(MPB_Q_LAPSE, w.lapse);
    if (A.qnt[MPB_Q_TDEW] >= 0) set(MPB_Q_TDEW, w.tdew);
    if (A.qnt[MPB_Q_TICE] >= 0) set(MPB_Q_TICE, w.tice);
  }
}

// dst[2 i + slot] = src[i]: one time level of a further field into its interleaved array
__global__ void pack_scalar_kernel(const float *src, float *dst, int slot, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[2 * i + slot] = src[i];
}
__global__ void swap_pairs_kernel(float2 *a, size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) { const float2 v = a[i]; a[i] = make_float2(v.y, v.x); }
}

// module_diff_pbl (src/mptrac.c:4343-4584): parcels with dt != 0, three normals per parcel from the shared counter stream
struct PblArgs {
  MetView met;
  PblFields f;
  double *time, *lon, *lat, *p;
  const double *dt;
  float *uvwp;
  unsigned long long ctr;
  long long ig0, np;
  const int *slot;      // as StepArgs::slot
};
__global__ void __launch_bounds__(128) diff_pbl_kernel(const __grid_constant__ PblArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  const double dt = A.dt[ip];
  if (dt == 0) return;
  Parcel a;
  a.time = A.time[ip]; a.lon = A.lon[ip]; a.lat = A.lat[ip]; a.p = A.p[ip];
  float *s = A.uvwp + 3 * ip;
  float up = s[0], vp = s[1], wp = s[2];
  diffuse_pbl(A.met, A.f, A.ctr, dt, (unsigned long long)(A.ig0 + (A.slot ? A.slot[ip] : ip)), a, up, vp, wp);
  s[0] = up; s[1] = vp; s[2] = wp;
  A.lon[ip] = a.lon; A.lat[ip] = a.lat; A.p[ip] = a.p;
}

// module_convection (src/mptrac.c:4102-4171): parcels with dt != 0; the uniform random number of parcel ip is number
// ig0 + ip of the module's module_rng call (method 0), addressed by global index like the normals of the diffusion modules
struct ConvArgs {
  MetView met;
  ConvView conv;
  const double *time, *lon, *lat, *dt;
  double *p;
  unsigned long long ctr;
  long long ig0, np;
  const int *slot;      // as StepArgs::slot
};
__global__ void __launch_bounds__(128) convection_kernel(const __grid_constant__ ConvArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np || A.dt[ip] == 0) return;
  Parcel a;
  a.time = A.time[ip]; a.lon = A.lon[ip]; a.lat = A.lat[ip]; a.p = A.p[ip];
  convect(A.met, A.conv, squares_uniform(A.ctr + (unsigned long long)(A.ig0 + (A.slot ? A.slot[ip] : ip))), a);
  A.p[ip] = a.p;
}

// module_chem_grid (src/mptrac.c:3885-4054): box per parcel + mass per box, then the box's volume mixing ratio per parcel
struct ChemArgs {
  MetView met;
  ChemGrid k;
  const double *time, *lon, *lat, *p, *m, *ens;
  double *cx, *mass;
  int *box;
  long long np;
  int ngrid;
};
__global__ void chem_mass_kernel(const __grid_constant__ ChemArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  const int b = chem_box(A.k, A.time[ip], A.lon[ip], A.lat[ip], A.p[ip]);
  A.box[ip] = b;
  if (b >= 0) atomicAdd(A.mass + b + (A.ens ? (int)A.ens[ip] * A.ngrid : 0), A.m[ip]);
}
__global__ void __launch_bounds__(128) chem_apply_kernel(const __grid_constant__ ChemArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  const int b = A.box[ip];
  if (b >= 0) A.cx[ip] = chem_vmr(A.met, A.k, b, A.mass[b + (A.ens ? (int)A.ens[ip] * A.ngrid : 0)]);
}

// module_bound_cond (src/mptrac.c:3789-3881): parcels with dt != 0
struct BoundArgs {
  MetView met;
  BoundView k;
  const double *time, *lon, *lat, *p, *dt;
  double *m, *vmr, *aoa, *cts[5];          // quantity rows or null
  const double *cts_time[5], *cts_vmr[5];
  int cts_n[5];
  double mass, mass_trend, vmr0, vmr_trend;
  long long np;
};
__global__ void __launch_bounds__(128) bound_cond_kernel(const __grid_constant__ BoundArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np || A.dt[ip] == 0) return;
  Parcel a;
  a.time = A.time[ip]; a.lon = A.lon[ip]; a.lat = A.lat[ip]; a.p = A.p[ip];
  if (!bound_applies(A.met, A.k, a)) return;
  if (A.m && A.mass >= 0) A.m[ip] = A.mass + A.mass_trend * a.time;
  if (A.vmr && A.vmr0 >= 0) A.vmr[ip] = A.vmr0 + A.vmr_trend * a.time;
#pragma unroll
  for (int k = 0; k < 5; k++)
    if (A.cts[k]) A.cts[k][ip] = series_at(A.cts_time[k], A.cts_vmr[k], A.cts_n[k], a.time);
  if (A.aoa) A.aoa[ip] = a.time;
}

// module_isosurf_init (init = 1) and module_isosurf (src/mptrac.c:4886-5004): every parcel, dt or not
struct IsoArgs {
  MetView met;
  const double *time, *lon, *lat;
  double *p, *iso_var;
  const double *ts, *ps;
  long long np;
  int mode, n, init;
};
__global__ void __launch_bounds__(128) isosurf_kernel(const __grid_constant__ IsoArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  Parcel a;
  a.time = A.time[ip]; a.lon = A.lon[ip]; a.lat = A.lat[ip]; a.p = A.p[ip];
  if (A.init) A.iso_var[ip] = isosurf_variable(A.met, A.mode, a);
  else A.p[ip] = isosurf_pressure(A.met, A.mode, A.mode == 4 ? 0.0 : A.iso_var[ip], a, A.ts, A.ps, A.n);
}

// the reset of the total loss rate (src/mptrac.c:7931-7936) and module_decay (4227-4263): parcels with dt != 0
struct DecayArgs {
  ClimView clim;
  const double *time, *lat, *p, *dt;
  double *m, *vmr, *mloss, *loss_rate;   // quantity rows or null
  double tdec_trop, tdec_strat, utm_ref_lat;
  long long np;
  int coord_type, decay;                 // decay = 0: only the reset
};
__global__ void __launch_bounds__(128) decay_kernel(const __grid_constant__ DecayArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  const double dt = A.dt[ip];
  if (dt == 0) return;
  double rate = 0;
  if (A.decay) {
    Parcel a;
    a.time = A.time[ip]; a.lon = 0; a.lat = A.lat[ip]; a.p = A.p[ip];
    double tdec;
    const double aux = decay_factor(A.clim, A.coord_type, A.utm_ref_lat, A.tdec_trop, A.tdec_strat, a, dt, tdec);
    if (A.m) {
      const double m = A.m[ip];
      if (A.mloss) A.mloss[ip] += m * (1 - aux);
      A.m[ip] = m * aux;
      rate = 1. / tdec;
    }
    if (A.vmr) A.vmr[ip] *= aux;
  }
  if (A.loss_rate) A.loss_rate[ip] = 0 + rate;
}

// module_advect on model levels (src/mptrac.c:3646-3657, 3680-3757) and module_advect_init (3762-3785): one parcel per
// thread.  `modules`: the per-parcel modules around the advection that the plan folded into this launch -- timesteps and
// the first position check before it, the second position check after it, when nothing else runs in between (a plain
// model-level configuration is then ONE launch per step instead of three; c2ml 0.281 -> see DESIGN.md 7) -- otherwise dt
// comes from the cache (a step segment ran before).
struct LevelArgs {
  MetView met;
  CtlView ctl;
  double *time, *lon, *lat, *p;
  double *dt;
  double *zq;           // the parcel's zeta / eta (null for ADVECT_VERT_COORD 2)
  unsigned short *hint; // per array slot: level + 1 found by the previous step's first lookup (0 = none); a search hint
                        // that is verified before use, so stale values (after a cell sort) only cost the bisection
  long long np;
  int vert_coord;
  unsigned modules;
};

template <int ORDER>
__global__ void __launch_bounds__(MPB_LEVEL_BLOCK, MPB_LEVEL_MINBLOCKS) advect_levels_kernel(const __grid_constant__ LevelArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  Parcel a;
  a.time = A.time[ip]; a.lon = A.lon[ip]; a.lat = A.lat[ip]; a.p = A.p[ip];
  double dt;
  if (A.modules & MOD_TIMESTEPS) {
    dt = parcel_dt(A.met, A.ctl, a);
    if (A.modules & MOD_STORE_DT) A.dt[ip] = dt;
  } else {
    dt = A.dt[ip];
  }
  if (dt == 0) return;
  if (A.modules & MOD_POS_PRE) fix_position(A.met, a);
  double z = 0;
  int hint = (int)A.hint[ip] - 1;
  advect_on_levels<ORDER>(A.met, A.vert_coord, dt, a, A.zq ? &z : nullptr, A.zq ? nullptr : &hint);
  if (A.modules & MOD_POS_POST) fix_position(A.met, a);
  A.time[ip] = a.time; A.lon[ip] = a.lon; A.lat[ip] = a.lat; A.p[ip] = a.p;
  if (A.zq) A.zq[ip] = z;
  else A.hint[ip] = (unsigned short)(hint + 1);
}

__global__ void __launch_bounds__(128) advect_init_kernel(const __grid_constant__ LevelArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  A.p[ip] = pressure_of_zeta(A.met, A.time[ip], A.zq[ip], A.lon[ip], A.lat[ip]);
}

typedef BoxGrid BoxArgs;   // (physics.cuh; the host fills the bounds and calls box_cells)

__global__ void box_index_kernel(BoxArgs b, const double *time, const double *lon, const double *lat,
                                 const double *p, int *box, long long np) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= np) return;
  box[ip] = box_index(b, time[ip], lon[ip], lat[ip], p[ip]);
}

// ------------------------------------------------------------------------------------------------
// Inter-parcel mixing (src/mptrac.c:5169-5347) -- the one step of the path in which parcels exchange information.
// Box records: per box (and ensemble member) one record {count, sum_0 .. sum_(nmix-1)} of doubles (counts stay exact:
// integers below 2^53), ALL mixed quantities at once -- the reference recomputes the same counts for every quantity
// (5287-5303) -- so a box costs one contiguous read per parcel.  With several ranks the box space is cut into contiguous
// slices, one per rank; every rank adds its parcels' contributions straight into the OWNER's slice through peer memory
// (NVLink atomics) and, after one barrier, reads its parcels' records back from there: the exchange moves only the
// records of occupied boxes (tens of bytes per parcel) instead of all-reducing the dense 5.8 M-box arrays.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = MPB_MAX_RANKS;

struct MixArgs {
  double *rec[kMaxRanks];     // every rank's slice of the box records (this rank's own: local memory)
  long long slice;            // boxes per rank
  int nmix, ngrid;            // mixed quantities; boxes per ensemble member
  int stride;                 // doubles per record: nmix + 1 rounded up to an even number (records are 16-byte aligned)
  const int *box;             // box of each parcel (-1 = outside)
  const double *ens;          // ensemble index per parcel (or null)
  double *q0;                 // quantity 0; quantity k starts k * q_stride further
  long long q_stride, np;
  int iq[MPB_MIX_MAXQ];
};

// Runs of equal records among consecutive lanes are summed inside the warp (parcels arrive cell-sorted, so the lanes of a
// warp mostly share a few boxes); only the last lane of a run touches memory.  src/mptrac.c:5287-5303
__global__ void mix_accumulate_kernel(const __grid_constant__ MixArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (the grid covers whole warps: no early return)
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  long long idx = -1;
  if (ip < A.np) {
    const int b = A.box[ip];
    if (b >= 0) idx = (long long)(A.ens ? (int)A.ens[ip] : 0) * A.ngrid + b;
  }
  const long long prev = __shfl_up_sync(full, idx, 1);
  const unsigned heads = __ballot_sync(full, lane == 0 || prev != idx);
  const int start = 31 - __clz(heads & (full >> (31 - lane)));             // first lane of this lane's run
  const bool tail = lane == 31 || ((heads >> (lane + 1)) & 1u);
  double *rec = nullptr;
  if (tail && idx >= 0) {
    const long long owner = idx / A.slice;
    rec = A.rec[owner] + (idx - owner * A.slice) * A.stride;
  }
  int n = 1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int tn = __shfl_up_sync(full, n, d);
    if (lane - d >= start) n += tn;
  }
  if (rec) atomicAdd(rec, (double)n);
  for (int k = 0; k < A.nmix; k++) {
    double x = idx >= 0 ? A.q0[(long long)A.iq[k] * A.q_stride + ip] : 0.0;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const double tx = __shfl_up_sync(full, x, d);
      if (lane - d >= start) x += tx;
    }
    if (rec) atomicAdd(rec + 1 + k, x);
  }
}

// src/mptrac.c:5307-5335: box mean, then relaxation towards it with the tropopause-weighted mixing parameter
__global__ void mix_apply_kernel(const __grid_constant__ MixArgs A, ClimView clim, const double *time, const double *lat,
                                 const double *p, double mix_trop, double mix_strat, int latlon, double utm_ref_lat) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  const int b = A.box[ip];
  if (b < 0) return;
  const long long idx = (long long)(A.ens ? (int)A.ens[ip] : 0) * A.ngrid + b;
  const long long owner = idx / A.slice;
  const double2 *rec = reinterpret_cast<const double2 *>(A.rec[owner] + (idx - owner * A.slice) * A.stride);
  double mixparam = 1.0;
  if (mix_trop < 1 || mix_strat < 1) {
    const double pt = tropopause_pressure(clim, time[ip], latlon ? lat[ip] : utm_ref_lat);
    const double w = weight_tropo(pt, p[ip]);
    mixparam = w * mix_trop + (1.0 - w) * mix_strat;
  }
  // (a record may live on another GPU: 16-byte loads, {count, sum_0} {sum_1, sum_2} ..., halve the requests over NVLink)
  double2 v = __ldcg(rec);
  const int n = (int)v.x;
  for (int k = 0; k < A.nmix; k++) {
    if (k > 0 && (k & 1)) v = __ldcg(rec + (k + 1) / 2);
    double mean = (k & 1) ? v.x : v.y;
    if (n > 0) mean /= n;
    double *q = A.q0 + (long long)A.iq[k] * A.q_stride + ip;
    const double qq = *q;
    *q = qq + (mean - qq) * mixparam;
  }
}

__global__ void mix_clear_kernel(const MixArgs A) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  const int b = A.box[ip];
  if (b < 0) return;
  const long long idx = (long long)(A.ens ? (int)A.ens[ip] : 0) * A.ngrid + b;
  double2 *rec = reinterpret_cast<double2 *>(A.rec[0] + idx * A.stride);
  for (int k = 0; k < A.stride / 2; k++) rec[k] = make_double2(0.0, 0.0);
}

// ---- several ranks: the routed exchange ----
// Remote atomics (one NVLink transaction per contribution) measured no better than NCCL's dense all-reduce (8 GPUs: 0.39 ms
// against 0.42 ms per step for configs[4]).  Instead every contribution travels ONCE as part of a coalesced store stream:
//   route   each parcel's {box, q_0 ..} goes into the inbox its box's OWNER keeps for this sender (peer memory), at a slot the
//           sender allocates with one warp-aggregated atomic per (warp, owner); the parcel remembers (owner, slot);
//   -- barrier --
//   serve   the owner folds all its inboxes into its slice of the box records (LOCAL atomics), then answers every entry with
//           the finished box means into the sender's outbox, same slot (coalesced stores to peer memory);
//   -- barrier --
//   apply   every parcel reads its answer from its own outbox (local memory) and relaxes.
// Per parcel 8 (1 + quantities) bytes cross NVLink on the way out and 8 x quantities on the way back, all of it as streaming stores.
struct RouteArgs {
  BoxArgs grid;                        // the mixing grid (the box of a parcel is computed here: no box array in memory)
  const double *time, *lon, *lat, *p;
  double *inbox[kMaxRanks];            // rank r's inbox region for THIS sender: [cap][E]
  unsigned long long *counts_at[kMaxRanks];   // rank r's table of entry counts, this sender's cell
  unsigned int *alloc;                 // [nranks] slots handed out so far (local)
  int2 *route;                         // [np] (owner, slot) per parcel (local)
  long long slice, np, q_stride;
  int nmix, ngrid, nranks, E;
  const double *ens;
  const double *q0;
  int iq[MPB_MIX_MAXQ];
};

constexpr int kRouteBlock = 1024;
__global__ void __launch_bounds__(kRouteBlock) mix_route_kernel(const __grid_constant__ RouteArgs A) {
  // Slot allocation in two levels: the lanes of a warp that share an owner reserve consecutive places in the BLOCK's count
  // for that owner (shared-memory atomic), one thread per owner then reserves the block's range with ONE global atomic.
  // (One global atomic per (warp, owner) put ~60 000 atomics per launch on at most `nranks` addresses, which serialise in L2:
  // the kernel took 69 us for 1.25 M parcels on 2 ranks, three times the single-rank accumulation.)
  __shared__ unsigned s_cnt[kMaxRanks], s_base[kMaxRanks];
  if (threadIdx.x < kMaxRanks) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // (the grid covers whole warps)
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int owner = -1;
  long long local = 0;
  if (ip < A.np) {
    const BoxArgs &g = A.grid;
    const int b = box_index(g, A.time[ip], A.lon[ip], A.lat[ip], A.p[ip]);
    if (b >= 0) {
      const long long idx = (long long)(A.ens ? (int)A.ens[ip] : 0) * A.ngrid + b;
      owner = (int)(idx / A.slice);
      local = idx - (long long)owner * A.slice;
    }
  }
  const unsigned peers = __match_any_sync(full, owner);
  unsigned place = 0;
  if (owner >= 0) {
    const int leader = __ffs(peers) - 1;
    if (lane == leader) place = atomicAdd(&s_cnt[owner], (unsigned)__popc(peers));
    place = __shfl_sync(peers, place, leader) + (unsigned)__popc(peers & ((1u << lane) - 1u));
  }
  __syncthreads();
  if (threadIdx.x < A.nranks && s_cnt[threadIdx.x] > 0) s_base[threadIdx.x] = atomicAdd(A.alloc + threadIdx.x, s_cnt[threadIdx.x]);
  __syncthreads();
  int slot = -1;
  if (owner >= 0) {
    slot = (int)(s_base[owner] + place);
    double *e = A.inbox[owner] + (size_t)slot * A.E;
    if (A.E == 2) {     // one mixed quantity: the entry is one 16-byte store
      *reinterpret_cast<double2 *>(e) = make_double2((double)local, A.q0[(long long)A.iq[0] * A.q_stride + ip]);
    } else {
      e[0] = (double)local;
      for (int k = 0; k < A.nmix; k++) e[1 + k] = A.q0[(long long)A.iq[k] * A.q_stride + ip];
    }
  }
  if (ip < A.np) A.route[ip] = make_int2(owner, slot);
}

// tell every owner how many entries this sender left in its inbox
__global__ void mix_publish_kernel(const __grid_constant__ RouteArgs A) {
  const int r = threadIdx.x;
  if (r < A.nranks) *A.counts_at[r] = (unsigned long long)A.alloc[r];
}

struct ServeArgs {
  const double *inbox;                 // this rank's inboxes [nranks][cap][E]
  const unsigned long long *counts;    // [nranks] entries per sender
  double *outbox_at[kMaxRanks];        // sender s's outbox region for THIS owner: [cap][E]
  double *rec;                         // this rank's slice of the box records [slice][E] (local)
  long long cap;
  int E;
};
// (the number of entries a sender left is only known on the device: a fixed grid strides over them)
__global__ void mix_fold_kernel(const __grid_constant__ ServeArgs A) {     // src/mptrac.c:5287-5303 for the boxes this rank owns
  const int s = blockIdx.y;
  const long long n = (long long)A.counts[s], stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const double *in = A.inbox + ((size_t)s * A.cap + (size_t)e) * A.E;
    double *rec = A.rec + (size_t)in[0] * A.E;
    atomicAdd(rec, 1.0);
    for (int k = 1; k < A.E; k++) atomicAdd(rec + k, in[k]);
  }
}
__global__ void mix_answer_kernel(const __grid_constant__ ServeArgs A) {
  const int s = blockIdx.y;
  const long long n = (long long)A.counts[s], stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const double *in = A.inbox + ((size_t)s * A.cap + (size_t)e) * A.E;
    const double *rec = A.rec + (size_t)in[0] * A.E;
    // the answer is the box MEAN per quantity (src/mptrac.c:5314-5316; the count of a box that holds an entry is >= 1): one
    // double per quantity crosses NVLink instead of {count, sums}
    double *out = A.outbox_at[s] + (size_t)e * (A.E - 1);
    const int n = (int)rec[0];
    for (int k = 1; k < A.E; k++) out[k - 1] = rec[k] / n;
  }
}

// src/mptrac.c:5307-5335 with the box record taken from this rank's outbox
__global__ void mix_apply_routed_kernel(const double *outbox, long long cap, int E, const int2 *route, int nmix, MixArgs A, ClimView clim,
                                        const double *time, const double *lat, const double *p, double mix_trop, double mix_strat,
                                        int latlon, double utm_ref_lat) {
  const long long ip = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ip >= A.np) return;
  const int2 r = route[ip];
  if (r.x < 0) return;
  const double *mean_of = outbox + (size_t)r.x * cap * E + (size_t)r.y * nmix;    // (owner r.x's region, packed means)
  double mixparam = 1.0;
  if (mix_trop < 1 || mix_strat < 1) {
    const double pt = tropopause_pressure(clim, time[ip], latlon ? lat[ip] : utm_ref_lat);
    const double w = weight_tropo(pt, p[ip]);
    mixparam = w * mix_trop + (1.0 - w) * mix_strat;
  }
  for (int k = 0; k < nmix; k++) {
    const double mean = mean_of[k];
    double *q = A.q0 + (long long)A.iq[k] * A.q_stride + ip;
    const double qq = *q;
    *q = qq + (mean - qq) * mixparam;
  }
}

// Barrier over the ranks of a multi-process run, in stream order: every rank writes the barrier's number into its slot
// of every peer's flag array (peer memory) and waits until all peers' numbers have arrived in its own.  The wait is
// bounded (~5 s): a missing peer raises *err instead of hanging the device.
struct FlagArgs {
  unsigned long long *flags[kMaxRanks];
};
__global__ void peer_barrier_kernel(const __grid_constant__ FlagArgs F, int rank, int nranks, unsigned long long epoch, int *err) {
  const int t = threadIdx.x;
  if (t >= nranks) return;
  __threadfence_system();
  unsigned long long *mine = F.flags[t] + rank, *theirs = F.flags[rank] + t;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(mine), "l"(epoch) : "memory");
  unsigned long long v = 0;
  for (long long spin = 0; spin < (1ll << 25); spin++) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(theirs) : "memory");
    if (v >= epoch) break;
    __nanosleep(128);
  }
  if (v < epoch) *err = 1;
  __threadfence_system();
}

// Gridded output over ranks: rank 0 adds the peers' partial arrays to its own (read through peer memory)
struct GridPeers {
  const double *sum[kMaxRanks], *sq[kMaxRanks];
  const int *cnt[kMaxRanks];
};
__global__ void grid_pull_kernel(const __grid_constant__ GridPeers G, int nranks, long long nbox, long long nval,
                                 double *sum, double *sq, int *cnt) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nval) {
    double a = sum[i], b = sq[i];
    for (int r = 1; r < nranks; r++) { a += __ldcg(G.sum[r] + i); b += __ldcg(G.sq[r] + i); }
    sum[i] = a; sq[i] = b;
  }
  if (i < nbox) {
    int n = cnt[i];
    for (int r = 1; r < nranks; r++) n += __ldcg(G.cnt[r] + i);
    cnt[i] = n;
  }
}

// src/mptrac.c:13862-13872 with kernel weight 1 (no GRID_KERNEL file)
// Gridded output: count, sum and sum of squares per box (src/mptrac.c:13840-13872).  Parcels arrive cell-sorted, so the
// lanes of a warp mostly fall into one or two boxes: each run of equal box indices is summed inside the warp (segmented
// scan over shuffles) and only the last lane of a run touches memory.  Runs need not be maximal (an unsorted stream just
// issues more atomics).
// ONE pass over the parcels (no box array in memory), few atomics and few shuffles: a warp owns 8 x 32 consecutive parcels
// and computes their boxes once.  A round of 32 whose lanes all fall into ONE box (the common case: ~200 parcels per output
// box in configs[3], parcels cell-sorted) only adds to lane-private partial sums of the open run -- no shuffle, no atomic;
// the partials are reduced over the warp and added to memory when the box changes.  A round that holds several runs closes
// the open run and sums each of its runs by a segmented scan.  Measured on the configs[3] share (12.5 M parcels, 360x180x1
// boxes; ncu): box_index_kernel 126 us + grid_accumulate_kernel 286 us (round 1), one kernel with a scan in every round 340 us,
// this form: see profiles/r02r_*.
constexpr int kBinRounds = 8;
#ifndef MPB_BIN_SCATTER
#define MPB_BIN_SCATTER 1
#endif
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  return v;
}
// (Measured and dropped: a shared-memory table of partial sums per block, to take the contention off the few boxes that the
// blocks in flight share when the parcels are sorted -- the reference's sort key clamps longitudes below the met grid's first
// column into that column (src/mptrac.c:5905, 3565-3573), so on a 0..360 grid the western half of the parcels is ordered by
// latitude and level only and every such parcel is a run of its own.  335 against 355 us sorted, 470 against 302 us unsorted:
// the fp64 shared-memory atomics are compare-and-swap loops.  profiles/r02r_binning.txt)
__global__ void __launch_bounds__(256) grid_bin_kernel(BoxArgs g, const double *time, const double *lon, const double *lat, const double *p,
                                                       const double *q, long long q_stride, int nq, long long nbox, int *cnt, double *sum,
                                                       double *sq, long long np) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  // Blocks take the chunks of the parcel array in a scattered order (a bijection of the block index: 2^31 - 1 is prime):
  // cell-sorted parcels put the blocks that run at the same time onto the same few output boxes, and atomics to one address
  // serialise (measured: the kernel was SLOWER on sorted parcels than on unsorted ones, 355 against 301 us).
  const unsigned long long chunk = MPB_BIN_SCATTER ? ((unsigned long long)blockIdx.x * 2147483647ull) % gridDim.x : blockIdx.x;
  const long long warp = ((long long)chunk * blockDim.x + threadIdx.x) >> 5;
  const long long first = warp * 32 * kBinRounds;
  if (first >= np) return;                                  // (warp-uniform)
  // the runs of equal boxes of every round, found once: box, first lane of the lane's run, "last lane of its run"
  int bx[kBinRounds];
  unsigned char start[kBinRounds];
  unsigned tails = 0, single = 0, crowded = 0;              // bit r: this lane ends a run in round r / round r is one run /
#pragma unroll                                              //        round r has (nearly) as many runs as lanes
  for (int r = 0; r < kBinRounds; r++) {
    const long long ip = first + r * 32 + lane;
    bx[r] = ip < np ? box_index(g, time[ip], lon[ip], lat[ip], p[ip]) : -1;
    const int prev = __shfl_up_sync(full, bx[r], 1);
    const unsigned heads = __ballot_sync(full, lane == 0 || prev != bx[r]);
    start[r] = (unsigned char)(31 - __clz(heads & (full >> (31 - lane))));
    if (lane == 31 || ((heads >> (lane + 1)) & 1u)) tails |= 1u << r;
    if (heads == 1u) single |= 1u << r;
    if (__popc(heads) >= 24) crowded |= 1u << r;
  }
  {  // counts: run lengths (no arithmetic on values)
    int open_box = -1, open_n = 0;                          // warp-uniform
#pragma unroll
    for (int r = 0; r < kBinRounds; r++) {
      const int b = bx[r];
      if ((single >> r) & 1u) {
        if (b != open_box) {
          if (open_box >= 0 && lane == 0) atomicAdd(cnt + open_box, open_n);
          open_box = b; open_n = 0;
        }
        open_n += 32;
      } else {
        if (open_box >= 0 && lane == 0) atomicAdd(cnt + open_box, open_n);
        open_box = -1; open_n = 0;
        if (((tails >> r) & 1u) && b >= 0) atomicAdd(cnt + b, lane - start[r] + 1);
      }
    }
    if (open_box >= 0 && lane == 0) atomicAdd(cnt + open_box, open_n);
  }
  for (int iq = 0; iq < nq; iq++) {      // sum and sum of squares of quantity iq
    double *const sum_q = sum + iq * nbox, *const sq_q = sq + iq * nbox;
    int open_box = -1;                   // warp-uniform
    double part_a = 0, part_b = 0;       // this lane's share of the open run
    auto close = [&]() {
      if (open_box >= 0) {
        const double a = warp_sum(part_a), bb = warp_sum(part_b);
        if (lane == 0) { atomicAdd(sum_q + open_box, a); atomicAdd(sq_q + open_box, bb); }
      }
      open_box = -1; part_a = 0; part_b = 0;
    };
#pragma unroll
    for (int r = 0; r < kBinRounds; r++) {
      const int b = bx[r];
      const long long ip = first + r * 32 + lane;
      const double x = b >= 0 ? q[(long long)iq * q_stride + ip] : 0.0;
      if ((single >> r) & 1u) {
        if (b != open_box) { close(); open_box = b; }
        part_a += x; part_b += x * x;
      } else {
        close();
        double a = x, bb = x * x;
        if ((crowded >> r) & 1u) {       // (nearly) every lane its own box: no point in a scan
          if (b >= 0) { atomicAdd(sum_q + b, a); atomicAdd(sq_q + b, bb); }
        } else {
          const int st = start[r];
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const double ta = __shfl_up_sync(full, a, d), tb = __shfl_up_sync(full, bb, d);
            if (lane - d >= st) { a += ta; bb += tb; }
          }
          if (((tails >> r) & 1u) && b >= 0) { atomicAdd(sum_q + b, a); atomicAdd(sq_q + b, bb); }
        }
      }
    }
    close();
  }
}

// uniform / normal stream of module_rng into an array (src/mptrac.c:5784-5828); test + shim support
__global__ void rng_fill_kernel(unsigned long long ctr0, double *rs, long long n, int method) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  if (method == 0) {
    rs[i] = squares_uniform(ctr0 + (unsigned long long)i);
  } else {
    const unsigned long long pa = (unsigned long long)i & ~1ull;
    const double r = sqrt(-2.0 * log(squares_uniform(ctr0 + pa)));
    const double u1 = squares_uniform(ctr0 + pa + 1);
    const float phi = (float)(2.0 * kPi * u1);
    if (i == n) {
      // slot n only ever holds a uniform, or the sine of the last pair when n is odd
      rs[i] = (n & 1) ? r * sinf(phi) : squares_uniform(ctr0 + (unsigned long long)i);
    } else {
      rs[i] = (i & 1) ? r * sinf(phi) : r * cosf(phi);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct mpb_team;

struct MetLevel {
  double time = 0;
  bool valid = false;
};

struct mpb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr, own_stream = nullptr;
  long long np_max = 0, np = 0;
  int nq = 0;
  long long ig0 = 0, global_np = -1;
  unsigned long long rng_ctr = 0;
  long long launches = 0;
  long long mix_clean = 0;   // the first mix_clean doubles of mix_rec are known to be zero (single rank: the records are cleared
                             // behind the parcels that touched them instead of zeroing all 5.8 M boxes every step)
  double mix_trace_ms[6] = {0, 0, 0, 0, 0, 0};   // MPTRAC_B200_TRACE_MIXING: summed phase times of mixing_inline
  long long mix_trace_n = 0;

  // parcels: [time | p | lon | lat | q0 .. q(nq-1)] each np_max doubles, two copies
  double *soa[2] = {nullptr, nullptr};
  int cur = 0;
  double *dt = nullptr;
  float *uvwp = nullptr;

  // sort scratch
  int *keys[2] = {nullptr, nullptr}, *perm[2] = {nullptr, nullptr};
  void *cub_tmp = nullptr;
  size_t cub_tmp_bytes = 0;
  // The order of the parcels in memory is the engine's own between two cell sorts (do_sort): slot[i] = the reference's array
  // slot of the parcel at position i.  Everything that leaves or enters by slot (mpb_get_atm, mpb_set_uvwp ...) first
  // restores the reference's order (unscramble).
  bool scrambled = false, leaving_soon = false;
  int *slot = nullptr, *slot2 = nullptr, *inv = nullptr, *perm2 = nullptr, *src = nullptr;
  float *uvwp2 = nullptr;
  double *dt2 = nullptr, *iso_var2 = nullptr;
  unsigned short *lev_hint2 = nullptr;

  // met
  MetLevel lev[2];
  Node *nodes = nullptr;     // both levels interleaved
  float4 *surf = nullptr;
  double *ax_lon = nullptr, *ax_lat = nullptr, *ax_p = nullptr;
  AxisCell *ax_lonc = nullptr, *ax_latc = nullptr, *ax_pc = nullptr;
  unsigned short *p_lut = nullptr;
  AxisTables tables;
  std::vector<double> h_lon, h_lat, h_p;
  int nx = 0, ny = 0, nz = 0, coord_type = 0;
  size_t node_cap = 0, col_cap = 0;
  float *stage_h = nullptr;  // pinned, 4 dense fields (6 model-level fields)
  char *grid_stage_h = nullptr;      // pinned landing area of mpb_grid_fetch (a copy into pageable memory is staged by the
  size_t grid_stage_bytes = 0;       // driver in synchronous pieces: 0.15 ms for the 1.3 MB of the reference's default grid)
  float *stage_d = nullptr;
  size_t stage_cap = 0;
  // model levels (ADVECT_VERT_COORD 1 / 2 / 3): both time levels interleaved like the nodes
  LevelNode *lev_p = nullptr, *lev_z = nullptr;   // {pl, ul, vl, wl}, {zetal, ul, vl, zeta_dotl}
  float4 *lev_pz = nullptr;                        // {pl0, zetal0, pl1, zetal1}
  int npl = 0;
  size_t lev_cap = 0;
  bool lev_valid[2] = {false, false};
  unsigned short *lev_hint = nullptr;              // LevelArgs::hint
  // module_chem_grid: mass per box (and ensemble member)
  double *chem_mass = nullptr;
  long long chem_cap = 0;
  // module_bound_cond: the five trace-gas time series of clim_t (ccl4, ccl3f, ccl2f2, n2o, sf6)
  double *cts_time[5] = {}, *cts_vmr[5] = {};
  int cts_n[5] = {};
  // module_isosurf: cache_t::iso_var (attached to the array slot, like uvwp) and the balloon series of ISOSURF 4
  double *iso_var = nullptr, *iso_ts = nullptr, *iso_ps = nullptr;
  int iso_n = 0;
  // further fields for module_meteo (mpb_met_view_t::x2 / x3): per field one array, both time levels of a node adjacent
  float2 *x2[MPB_NX2] = {}, *x3[MPB_NX3] = {};
  bool x2_valid[2][MPB_NX2] = {}, x3_valid[2][MPB_NX3] = {};

  // clim
  double *cl_time = nullptr, *cl_lat = nullptr, *cl_tropo = nullptr;
  int cl_ntime = 0, cl_nlat = 0;

  // mixing / grid boxes
  int *box = nullptr;
  double *mix_rec = nullptr;            // box records {count, sums} of a single-rank run (dense, zeroed every step)
  long long mix_cap = 0;                // ... its capacity in doubles
  int nmix = 0, mix_iq[MPB_MIX_MAXQ] = {};   // quantities mixed by the current step
  long long mix_total = 0, mix_slice = 0;    // boxes (x ensemble members); boxes per rank
  double *grid_sum = nullptr, *grid_sq = nullptr;
  int *grid_cnt = nullptr;
  long long grid_cap = 0, grid_nbox = 0;
  bool grid_in_area = false;

  // ranks that exchange box records through peer memory (mpb_peer_*): this rank's exchange area holds the barrier flags, the
  // inboxes (contributions of every rank to the boxes this rank owns) and outboxes (the owners' answers to this rank's
  // parcels) of the routed mixing exchange, and its partial arrays of the gridded output; area[r] is rank r's area as
  // mapped into this process
  int rank = 0, nranks = 1;
  char *area[kMaxRanks] = {};
  bool area_ipc[kMaxRanks] = {};
  size_t area_bytes = 0, area_mix_bytes = 0, area_grid_bytes = 0;
  BoxArgs mix_grid;                     // the mixing grid of the current step
  unsigned int *mix_alloc = nullptr;    // routed exchange: slots handed out per owner this step [kMaxRanks]
  int2 *mix_route = nullptr;            // ... (owner, slot) of every parcel [np_max]
  unsigned long long epoch = 0;
  int *peer_err = nullptr;              // device flag raised by a barrier that timed out
  struct mpb_team *team = nullptr;      // set when a team drives this context (barriers are then stream events)

  mpb_ctl_t ctl;
  bool have_ctl = false;

  std::vector<std::pair<step_fn, unsigned>> resident;   // blocks the device holds at once, per step kernel
  long long host_h2d = 0, host_d2h = 0;   // bytes the last mpb_run_timestep_host moved across the host link
  bool q_stale = false;   // mpb_run_timestep_host moved only the quantities the path reads: the device copy of q[] is not current
  bool quad = false, quad_split = false;                // form of the step kernel (MPTRAC_B200_STEP, MPTRAC_B200_QUAD_SPLIT)
  bool tile = false;                                    // MPTRAC_B200_STEP=tile: met window staged in shared memory by TMA
  CUtensorMap tmap;                                     // the node array as a 4-D tensor [nx][ny][nz][8 floats] (tile form)
  bool tmap_ok = false;
  TileShape tshape = {0, 0, 0};

  // host-resident stepping (mpb_run_timestep_host): streams that each carry whole chunks
  cudaStream_t lane[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t lane_done[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t lane_go = nullptr;

  double *arr(int which) const { return soa[cur] + (size_t)which * np_max; }  // 0 time 1 p 2 lon 3 lat 4+ q
  double *time() const { return arr(0); }
  double *p() const { return arr(1); }
  double *lon() const { return arr(2); }
  double *lat() const { return arr(3); }
  double *q(int iq) const { return arr(4 + iq); }
};

static inline unsigned nblocks(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

static void use(mpb_ctx *c) {
  REQUIRE(c != nullptr, "null context");
  CK(cudaSetDevice(c->device));
}

static MetView met_view(const mpb_ctx *c) {
  REQUIRE(c->lev[0].valid && c->lev[1].valid, "both met levels must be set before stepping");
  REQUIRE(c->lev[0].time != c->lev[1].time, "met0 and met1 carry the same time");
  MetView g;
  g.f = c->nodes; g.s = c->surf;
  g.lon = c->ax_lon; g.lat = c->ax_lat; g.p = c->ax_p;
  g.lonc = c->ax_lonc; g.latc = c->ax_latc; g.pc = c->ax_pc; g.p_lut = c->p_lut;
  fill_axis_scalars(g, c->h_lon.data(), c->nx, c->h_lat.data(), c->ny, c->h_p.data(), c->nz, c->coord_type,
                    c->lev[0].time, c->lev[1].time, c->tables);
  const bool levels = c->npl >= 2 && c->lev_valid[0] && c->lev_valid[1];
  g.lp = levels ? c->lev_p : nullptr; g.lz = levels ? c->lev_z : nullptr; g.pz = levels ? c->lev_pz : nullptr;
  g.npl = levels ? c->npl : 0;
  return g;
}

static ClimView clim_view(const mpb_ctx *c) {
  ClimView v;
  v.time = c->cl_time; v.lat = c->cl_lat; v.tropo = c->cl_tropo;
  v.ntime = c->cl_ntime; v.nlat = c->cl_nlat;
  return v;
}

static CtlView ctl_view(const mpb_ctx *c, double t) {
  const mpb_ctl_t &k = c->ctl;
  CtlView v;
  v.t = t; v.t_start = k.t_start; v.t_stop = k.t_stop; v.dt_met = k.dt_met;
  v.utm_ref_lat = k.met_utm_ref_lat;
  v.dx_pbl = k.turb_dx_pbl; v.dx_trop = k.turb_dx_trop; v.dx_strat = k.turb_dx_strat;
  v.dz_pbl = k.turb_dz_pbl; v.dz_trop = k.turb_dz_trop; v.dz_strat = k.turb_dz_strat;
  v.mesox = k.turb_mesox; v.mesoz = k.turb_mesoz; v.pbl_trans = k.turb_pbl_trans;
  v.ctr_turb = 0; v.ctr_meso = 0;
  v.direction = k.direction; v.pbl_scheme = k.turb_pbl_scheme;
  return v;
}

static bool turb_enabled(const mpb_ctl_t &k) {  // src/mptrac.c:7891-7895
  return k.diffusion && (k.turb_dx_pbl > 0 || k.turb_dz_pbl > 0 || k.turb_dx_trop > 0 ||
                         k.turb_dz_trop > 0 || k.turb_dx_strat > 0 || k.turb_dz_strat > 0);
}
static bool meso_enabled(const mpb_ctl_t &k) {  // src/mptrac.c:7902
  return k.diffusion && (k.turb_mesox > 0 || k.turb_mesoz > 0);
}
static bool sedi_enabled(const mpb_ctl_t &k) { return k.qnt_rp >= 0 && k.qnt_rhop >= 0; }  // :7911

static unsigned long long rng_draw(mpb_ctx *c) {  // one module_rng(…, 3*np, …) call
  const unsigned long long start = c->rng_ctr;
  const long long n = c->global_np >= 0 ? c->global_np : c->np;
  c->rng_ctr += 3ull * (unsigned long long)n + 1ull;
  return start;
}

// kernel arguments of one (possibly restricted) step over all parcels; draws the module random-number counters
static StepArgs step_args(mpb_ctx *c, double t, int advect, unsigned phys, unsigned modules) {
  REQUIRE(c->have_ctl, "mpb_set_ctl has not been called");
  StepArgs A;
  A.met = met_view(c);
  A.clim = clim_view(c);
  A.ctl = ctl_view(c, t);
  if (phys & PHYS_TURB) {
    REQUIRE(c->cl_tropo != nullptr, "diff_turb needs the tropopause climatology (mpb_set_clim_tropo)");
    A.ctl.ctr_turb = rng_draw(c);
  }
  if (phys & PHYS_MESO) A.ctl.ctr_meso = rng_draw(c);
  if (phys & (PHYS_TURB | PHYS_MESO)) REQUIRE(c->ctl.rng_type == 1, "only RNG_TYPE 1 (Squares) runs on the device");
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p();
  A.dt = c->dt; A.uvwp = c->uvwp;
  A.rp = A.rhop = nullptr;
  A.in_time = A.in_lon = A.in_lat = A.in_p = nullptr;
  A.host_time = A.host_lon = A.host_lat = A.host_p = nullptr;
  A.time_in = 0; A.uniform_time = 0;
  if (phys & PHYS_SEDI) {
    REQUIRE(c->ctl.qnt_rp >= 0 && c->ctl.qnt_rp < c->nq && c->ctl.qnt_rhop >= 0 && c->ctl.qnt_rhop < c->nq,
            "sedimentation needs quantities rp and rhop");
    A.rp = c->q(c->ctl.qnt_rp); A.rhop = c->q(c->ctl.qnt_rhop);
  }
  A.np = c->np; A.ig0 = c->ig0; A.modules = modules;
  A.slot = c->scrambled ? c->slot : nullptr;
  if (advect > 0) REQUIRE(c->ctl.advect_vert_coord == 0, "the fused step advects on pressure levels only");
  return A;
}

// blocks of `fn` the device holds at once (cached per kernel instantiation)
// (measurement aid: MPTRAC_B200_PAD_SMEM=<bytes> of unused dynamic shared memory per block lowers the number of resident
// blocks of the step kernel -- how sensitive is it to occupancy? -- DESIGN.md 3.1)
static size_t pad_smem() {
  static long long v = -1;
  if (v < 0) { const char *e = std::getenv("MPTRAC_B200_PAD_SMEM"); v = e ? std::atoll(e) : 0; }
  return (size_t)v;
}
static unsigned resident_blocks(mpb_ctx *c, step_fn fn) {
  for (auto &e : c->resident) if (e.first == fn) return e.second;
  int per_sm = 0, sms = 0;
  if (pad_smem() > 40 * 1024) CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad_smem()));
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, kBlock, pad_smem()));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
  const unsigned n = (unsigned)std::max(1, per_sm * sms);
  c->resident.emplace_back(fn, n);
  return n;
}

// Which form of the step kernel?  The default keeps one thread per parcel (step_kernel); MPTRAC_B200_STEP=quad gives every
// coordinate of a parcel its own lane (quad.cuh) wherever that form exists: advection on a global longitude / latitude grid.
// Measured on B200 (C2, RK4, 1 M parcels): 196 us per step against 114 us -- the lane split removes three quarters of the
// f32 -> f64 conversions and raises issue utilisation from 38 % to 53 %, but it executes 2.7x the warp instructions (the
// search / range check / position bookkeeping of a stage is replicated in every lane while only the 66 interpolation
// operations are divided), profiles/r02b_*.  Kept as a tested variant, not the default.
// (read when a context is created; MPTRAC_B200_QUAD_SPLIT=1: with diffusion / sedimentation on, advect in the quad kernel and
// run the other modules in a second, classic launch)
static bool quad_enabled(const mpb_ctx *c) { return c->quad; }
static bool quad_split_enabled(const mpb_ctx *c) { return c->quad_split; }

static step_fn pick_quad(int advect) {
  switch (advect) {
    case 1: return quad_step_kernel<1, 3>;
    case 2: return quad_step_kernel<2, 3>;
    case 4: return quad_step_kernel<4, 3>;
    default: throw std::runtime_error("ADVECT must be 1, 2 or 4");
  }
}

// launch the step kernel for parcels [off, off + cnt) on `stream`
static void launch_range(mpb_ctx *c, StepArgs A, int advect, unsigned phys, long long off, long long cnt, cudaStream_t stream) {
  if (cnt <= 0) return;
  A.time += off; A.lon += off; A.lat += off; A.p += off; A.dt += off; A.uvwp += 3 * off;
  if (A.in_time) { A.in_time += off; A.in_lon += off; A.in_lat += off; A.in_p += off; }
  if (A.host_lon) { if (A.host_time) A.host_time += off; A.host_lon += off; A.host_lat += off; A.host_p += off; }
  if (A.rp) { A.rp += off; A.rhop += off; }
  A.np = cnt; A.ig0 += off;
  if (A.slot) A.slot += off;
  if (c->tile && c->tmap_ok && advect > 0 && !A.in_time && !A.host_lon) {
    tile_fn tf = pick_tile(advect, phys);
    const size_t bytes = (size_t)c->tshape.nx * c->tshape.ny * c->tshape.nz * sizeof(Node);
    if (bytes > 48 * 1024) CK(cudaFuncSetAttribute(tf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    tf<<<nblocks(cnt, 128), 128, bytes, stream>>>(A, c->tmap, c->tshape);
    CK(cudaGetLastError());
    c->launches++;
    return;
  }
  const bool quad_ok = quad_enabled(c) && advect > 0 && A.met.coord_type == 0 && !A.met.local;
  if (quad_ok && (phys == 0 || (quad_split_enabled(c) && !A.in_time))) {
    step_fn qf = pick_quad(advect);
    StepArgs Q = A;
    if (phys != 0) Q.modules = (A.modules & (MOD_TIMESTEPS | MOD_POS_PRE)) | MOD_STORE_DT;   // the second launch reads dt from memory
    const long long warps = (cnt + 9) / 10;            // 10 parcels per warp (groups of 3 lanes)
    unsigned qgrid = (unsigned)std::min<long long>((warps + 3) / 4, (long long)resident_blocks(c, qf));
    qf<<<qgrid, 128, 0, stream>>>(Q);
    CK(cudaGetLastError());
    c->launches++;
    if (phys == 0) return;
    advect = 0;
    A.modules &= ~(MOD_TIMESTEPS | MOD_STORE_DT | MOD_POS_PRE);
  }
  step_fn fn = pick_step(advect, phys);
  unsigned grid = nblocks(cnt, kBlock);
#if MPB_PERSIST
  grid = std::min(grid, resident_blocks(c, fn));
#endif
  fn<<<grid, kBlock, pad_smem(), stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

static void launch_step(mpb_ctx *c, double t, int advect, unsigned phys, unsigned modules) {
  const StepArgs A = step_args(c, t, advect, phys, modules);
  launch_range(c, A, advect, phys, 0, c->np, c->stream);
}

static LevelArgs level_args(mpb_ctx *c, double t = 0, unsigned modules = 0) {
  const mpb_ctl_t &k = c->ctl;
  LevelArgs A;
  A.met = met_view(c);
  A.ctl = ctl_view(c, t);
  A.modules = modules;
  REQUIRE(A.met.npl >= 2, "ADVECT_VERT_COORD 1, 2 and 3 need the model-level fields of both met levels (mpb_met_view_t::pl ...)");
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p(); A.dt = c->dt;
  A.np = c->np; A.vert_coord = k.advect_vert_coord;
  if (!c->lev_hint) {
    CK(cudaMalloc(&c->lev_hint, sizeof(unsigned short) * (size_t)c->np_max));
    CK(cudaMemsetAsync(c->lev_hint, 0, sizeof(unsigned short) * (size_t)c->np_max, c->stream));
  }
  A.hint = c->lev_hint;
  A.zq = nullptr;
  if (k.advect_vert_coord == 1 || k.advect_vert_coord == 3) {
    const int iq = k.advect_vert_coord == 1 ? k.qnt_zeta : k.qnt_eta;
    REQUIRE(iq >= 0 && iq < c->nq, "ADVECT_VERT_COORD 1 / 3 need the quantity zeta / eta");
    A.zq = c->q(iq);
  }
  return A;
}

static void launch_advect_levels(mpb_ctx *c, double t, unsigned modules) {
  if (c->np == 0) return;
  const LevelArgs A = level_args(c, t, modules);
  const unsigned grid = nblocks(c->np, MPB_LEVEL_BLOCK);
  switch (c->ctl.advect) {
    case 1: advect_levels_kernel<1><<<grid, MPB_LEVEL_BLOCK, 0, c->stream>>>(A); break;
    case 2: advect_levels_kernel<2><<<grid, MPB_LEVEL_BLOCK, 0, c->stream>>>(A); break;
    case 4: advect_levels_kernel<4><<<grid, MPB_LEVEL_BLOCK, 0, c->stream>>>(A); break;
    default: REQUIRE(false, "ADVECT must be 0, 1, 2 or 4");
  }
  CK(cudaGetLastError());
  c->launches++;
}

static void launch_advect_init(mpb_ctx *c) {   // src/mptrac.c:3762-3785: ADVECT_VERT_COORD 1 only
  if (c->np == 0 || c->ctl.advect_vert_coord != 1) return;
  const LevelArgs A = level_args(c);
  advect_init_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

static void ensure_boxes(mpb_ctx *c) {
  if (!c->box) CK(cudaMalloc(&c->box, sizeof(int) * (size_t)std::max<long long>(c->np_max, 1)));
}

static void radix_sort_pairs(mpb_ctx *c, const int *keys_in, int *keys_out, const int *vals_in, int *vals_out, long long np, int bits) {
  size_t need = 0;
  CK(cub::DeviceRadixSort::SortPairs(nullptr, need, keys_in, keys_out, vals_in, vals_out, (int)np, 0, bits, c->stream));
  if (need > c->cub_tmp_bytes) {
    if (c->cub_tmp) CK(cudaFree(c->cub_tmp));
    CK(cudaMalloc(&c->cub_tmp, need));
    c->cub_tmp_bytes = need;
  }
  CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp, need, keys_in, keys_out, vals_in, vals_out, (int)np, 0, bits, c->stream));
  c->launches += 3;  // cub: histogram + onesweep passes (>= 3 launches)
}

static void ensure_sort_buffers(mpb_ctx *c) {
  if (!c->keys[0]) {
    for (int i = 0; i < 2; i++) {
      CK(cudaMalloc(&c->keys[i], sizeof(int) * (size_t)c->np_max));
      CK(cudaMalloc(&c->perm[i], sizeof(int) * (size_t)c->np_max));
    }
  }
  if (!c->soa[1]) CK(cudaMalloc(&c->soa[1], sizeof(double) * (size_t)c->np_max * (size_t)(4 + c->nq)));
}
static void ensure_order_buffers(mpb_ctx *c) {
  ensure_sort_buffers(c);
  if (c->slot) return;
  const size_t n = (size_t)std::max<long long>(c->np_max, 1);
  for (int **b : {&c->slot, &c->slot2, &c->inv, &c->perm2, &c->src}) CK(cudaMalloc(b, sizeof(int) * n));
  CK(cudaMalloc(&c->uvwp2, sizeof(float) * 3 * n));
  CK(cudaMalloc(&c->dt2, sizeof(double) * n));
}

// hand the slot-bound state over and swap the buffers
static void hand_over(mpb_ctx *c, const int *slot_new, const int *held_at, const int *src, const double *dt_of_slot) {
  const size_t n = (size_t)std::max<long long>(c->np_max, 1);
  if (c->iso_var && !c->iso_var2) CK(cudaMalloc(&c->iso_var2, sizeof(double) * n));
  if (c->lev_hint && !c->lev_hint2) CK(cudaMalloc(&c->lev_hint2, sizeof(unsigned short) * n));
  StateArgs A;
  A.slot_new = slot_new; A.held_at = held_at; A.src = src;
  A.uvwp_old = c->uvwp; A.uvwp_new = c->uvwp2;
  // (dt_of_slot may live in dt2: then the new dt goes into the old array, whose values are dead once every dt is recomputed)
  A.dt_old = c->dt; A.dt_of_slot = dt_of_slot; A.dt_new = dt_of_slot ? c->dt : c->dt2;
  A.iso_old = c->iso_var; A.iso_new = c->iso_var2;
  A.hint_old = c->lev_hint; A.hint_new = c->lev_hint2;
  A.np = c->np;
  hand_over_kernel<<<nblocks(c->np, 256), 256, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
  std::swap(c->uvwp, c->uvwp2);
  if (!dt_of_slot) std::swap(c->dt, c->dt2);
  if (c->iso_var) std::swap(c->iso_var, c->iso_var2);
  if (c->lev_hint) std::swap(c->lev_hint, c->lev_hint2);
}

// back to the reference's order: position = slot
static void unscramble(mpb_ctx *c) {
  if (!c->scrambled) return;
  c->scrambled = false;
  if (c->np == 0) return;
  const long long np = c->np;
  invert_kernel<<<nblocks(np, 256), 256, 0, c->stream>>>(c->slot, c->inv, np);
  gather_kernel<<<nblocks(np, 256), 256, 0, c->stream>>>(c->soa[c->cur], c->soa[c->cur ^ 1], c->inv, np, c->np_max, 4 + c->nq);
  CK(cudaGetLastError());
  c->launches += 2;
  c->cur ^= 1;
  hand_over(c, nullptr, c->inv, c->inv, nullptr);      // slot s takes what sits at position inv[s], hint included
}

// Does this configuration sort rarely enough for a second sort per event to pay?  (MPTRAC_B200_PRIVATE_ORDER=0: never)
static bool own_order_wanted(const mpb_ctx *c) {
  if (c->leaving_soon) return false;
  const char *e = std::getenv("MPTRAC_B200_PRIVATE_ORDER");
  if (e && std::atoi(e) == 0) return false;
  if (e && std::atoi(e) == 2) return true;             // (tests: whatever the cadence and the size)
  // Measured (profiles/r02v_*): a sort event costs ~1 ms more per 10 M parcels; the step gains 6-15 % where the met grid does not
  // sit in L2 or the parcels are dense, 3 % on configs[1] (1 M parcels on a grid that fits L2), where the event costs more
  // than 24 steps gain.
  return c->ctl.sort_dt >= 6 * std::fabs(c->ctl.dt_mod) && c->np >= 4000000;
}

// module_sort (src/mptrac.c:5887-5995).  The reference orders the parcels by the met cell of their RAW longitude: on a grid
// that runs 0..360 every parcel west of Greenwich (module_position keeps longitudes in [-180, 180)) falls into column 0 of the
// key and is ordered by latitude and level only -- half a sort (measured: DESIGN.md 9).  The engine therefore keeps TWO
// orders.  The reference's sort is carried out on the SLOTS: which parcel sits in which slot of atm_t afterwards (a stable
// sort in slot order), with dt computed per slot before the permutation (src/mptrac.c:7877-7881).  The parcels themselves are
// then laid out by the cell their lookups fall into (wrapped longitude); slot[i] remembers the slot of the parcel at position
// i for its random numbers, and what the reference leaves in the slot (uvwp, dt, iso_var) is handed to the slot's new parcel.
// Everything that addresses parcels by slot from outside restores the reference's order first (unscramble).
static void do_sort(mpb_ctx *c, double t = 0, unsigned modules = 0, bool own_order = false) {
  if (c->np == 0) return;
  const long long np = c->np;
  MetView g = met_view(c);
  const long long ncell = (long long)g.nx * g.ny * g.nz;
  int bits = 1;
  while ((1ll << bits) < ncell && bits < 31) bits++;
  const unsigned grid = nblocks(np, 256);
  if (!own_order) {
    unscramble(c);
    ensure_sort_buffers(c);
    sort_keys_kernel<<<grid, 256, 0, c->stream>>>(g, ctl_view(c, t), c->time(), c->lon(), c->lat(), c->p(), c->keys[0], c->perm[0],
                                                  (modules & MOD_TIMESTEPS) ? c->dt : nullptr, np);
    CK(cudaGetLastError());
    c->launches++;
    radix_sort_pairs(c, c->keys[0], c->keys[1], c->perm[0], c->perm[1], np, bits);
    gather_kernel<<<grid, 256, 0, c->stream>>>(c->soa[c->cur], c->soa[c->cur ^ 1], c->perm[1], np, c->np_max, 4 + c->nq);
    CK(cudaGetLastError());
    c->launches++;
    c->cur ^= 1;
    return;
  }
  ensure_order_buffers(c);
  const int *held_at = nullptr;                 // position of the parcel that holds slot s now (null: s)
  if (c->scrambled) {
    invert_kernel<<<grid, 256, 0, c->stream>>>(c->slot, c->inv, np);
    held_at = c->inv;
    c->launches++;
  }
  // both keys per position (and dt: per position = per parcel = per slot before the permutation, handed over below like uvwp)
  two_keys_kernel<<<grid, 256, 0, c->stream>>>(g, ctl_view(c, t), c->time(), c->lon(), c->lat(), c->p(), c->perm2, c->src,
                                               (modules & MOD_TIMESTEPS) ? c->dt : nullptr, np);
  // 1. the reference's sort, on the slots: perm[1][s] = position (now) of the parcel that takes slot s
  pick_keys_kernel<<<grid, 256, 0, c->stream>>>(c->perm2, held_at, c->keys[0], c->perm[0], true, np);
  CK(cudaGetLastError());
  radix_sort_pairs(c, c->keys[0], c->keys[1], c->perm[0], c->perm[1], np, bits);
  // 2. the engine's order: perm2[i] = slot of the parcel that goes to position i
  pick_keys_kernel<<<grid, 256, 0, c->stream>>>(c->src, c->perm[1], c->keys[0], c->perm[0], false, np);
  CK(cudaGetLastError());
  radix_sort_pairs(c, c->keys[0], c->keys[1], c->perm[0], c->perm2, np, bits);
  compose_kernel<<<grid, 256, 0, c->stream>>>(c->perm2, c->perm[1], c->slot2, c->src, np);
  gather_kernel<<<grid, 256, 0, c->stream>>>(c->soa[c->cur], c->soa[c->cur ^ 1], c->src, np, c->np_max, 4 + c->nq);
  CK(cudaGetLastError());
  c->launches += 5;
  c->cur ^= 1;
  hand_over(c, c->slot2, held_at, c->src, nullptr);
  std::swap(c->slot, c->slot2);
  c->scrambled = true;
}

static long long mixing_total(const mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  return (long long)k.mixing_nx * k.mixing_ny * k.mixing_nz * (k.nens > 0 ? k.nens : 1);
}

// ---- exchange area of a rank (mpb_peer_init): [flags + error word | three sets of its slice of the box records | its partial
// arrays of the gridded output] ----
constexpr size_t kAreaHeader = 4096, kAreaErrOffset = 2048;
// mixing region of an area: [entry counts per sender, 256 B | inboxes [nranks][cap][E] | outboxes [nranks][cap][E]], E = mixed
// quantities + 1 doubles per entry, cap = parcels one rank may send (the same on every rank: equal area sizes)
constexpr size_t kRouteHeader = 256;
static long long route_cap(const mpb_ctx *c) {
  if (c->area_mix_bytes <= kRouteHeader) return 0;
  return (long long)((c->area_mix_bytes - kRouteHeader) / (2 * (size_t)c->nranks * (size_t)(c->nmix + 1) * sizeof(double)));
}
static unsigned long long *route_counts(const mpb_ctx *c, int owner) { return (unsigned long long *)(c->area[owner] + kAreaHeader); }
static double *route_inbox(const mpb_ctx *c, int owner, int sender) {
  return (double *)(c->area[owner] + kAreaHeader + kRouteHeader) + (size_t)sender * (size_t)route_cap(c) * (size_t)(c->nmix + 1);
}
static double *route_outbox(const mpb_ctx *c, int holder, int owner) {
  return (double *)(c->area[holder] + kAreaHeader + kRouteHeader) + ((size_t)c->nranks + (size_t)owner) * (size_t)route_cap(c) * (size_t)(c->nmix + 1);
}
static char *grid_region(const mpb_ctx *c, int r) { return c->area[r] + kAreaHeader + c->area_mix_bytes; }

// stream-ordered barrier over the ranks of a multi-process run (a team of contexts in ONE process uses stream events instead)
static void peer_barrier(mpb_ctx *c) {
  if (c->nranks <= 1) return;
  REQUIRE(c->team == nullptr, "internal: a team context must not use the flag barrier");
  for (int r = 0; r < c->nranks; r++) REQUIRE(c->area[r] != nullptr, "mpb_peer_attach has not been called");
  FlagArgs F;
  for (int r = 0; r < kMaxRanks; r++) F.flags[r] = r < c->nranks ? (unsigned long long *)c->area[r] : nullptr;
  c->epoch++;
  peer_barrier_kernel<<<1, 32, 0, c->stream>>>(F, c->rank, c->nranks, c->epoch, c->peer_err);
  CK(cudaGetLastError());
  c->launches++;
}

// Box index of every parcel, the list of mixed quantities, the (zeroed) box records this rank keeps.
static void mixing_prepare(mpb_ctx *c, double t) {
  const mpb_ctl_t &k = c->ctl;
  REQUIRE(!c->q_stale, "mixing: the device copy of the quantities is not current (mpb_run_timestep_host): call mpb_set_atm");
  ensure_boxes(c);
  BoxArgs b;
  b.t0 = t - 0.5 * k.dt_mod; b.t1 = t + 0.5 * k.dt_mod;
  b.lon0 = k.mixing_lon0; b.lon1 = k.mixing_lon1; b.lat0 = k.mixing_lat0; b.lat1 = k.mixing_lat1;
  b.z0 = k.mixing_z0; b.z1 = k.mixing_z1; b.nx = k.mixing_nx; b.ny = k.mixing_ny; b.nz = k.mixing_nz;
  box_cells(b);
  const long long total = mixing_total(c);
  REQUIRE(total > 0 && total < (1ll << 31), "mixing grid size out of range");
  if (k.mixing_trop < 1 || k.mixing_strat < 1)
    REQUIRE(c->cl_tropo != nullptr, "mixing needs the tropopause climatology (mpb_set_clim_tropo)");
  c->nmix = 0;
  for (int i = 0; i < k.n_mix_qnt; i++)
    if (k.mix_qnt[i] >= 0) {
      REQUIRE(k.mix_qnt[i] < c->nq, "mixing quantity index out of range");
      c->mix_iq[c->nmix++] = k.mix_qnt[i];
    }
  c->mix_total = total;
  const long long stride = (c->nmix + 2) / 2 * 2;      // {count, sums} padded to 16-byte records
  if (c->nranks == 1) {
    c->mix_slice = total;
    const long long need = stride * total;
    if (need > c->mix_cap) {
      if (c->mix_rec) CK(cudaFree(c->mix_rec));
      CK(cudaMalloc(&c->mix_rec, sizeof(double) * (size_t)need));
      c->mix_cap = need;
      c->mix_clean = 0;
    }
    if (c->mix_clean != need) CK(cudaMemsetAsync(c->mix_rec, 0, sizeof(double) * (size_t)need, c->stream));
    c->mix_clean = 0;     // dirty until mixing_inline has cleared behind its parcels
  } else {
    // routed exchange: this rank keeps the records of its slice of the box space locally and zeroes them every step
    c->mix_slice = (total + c->nranks - 1) / c->nranks;
    const long long E = c->nmix + 1, need = E * c->mix_slice;
    for (int r = 0; r < c->nranks; r++) REQUIRE(c->area[r] != nullptr, "mpb_peer_attach has not been called");
    REQUIRE(c->np <= route_cap(c), "the exchange area is too small for this rank's parcels (mpb_peer_init: mix_bytes >= 256 + 16 x nranks x "
                                   "parcels per rank x (mixed quantities + 1))");
    if (need > c->mix_cap) {
      if (c->mix_rec) CK(cudaFree(c->mix_rec));
      CK(cudaMalloc(&c->mix_rec, sizeof(double) * (size_t)need));
      c->mix_cap = need;
    }
    c->mix_clean = 0;
    if (!c->mix_alloc) {
      CK(cudaMalloc(&c->mix_alloc, sizeof(unsigned int) * kMaxRanks));
      CK(cudaMalloc(&c->mix_route, sizeof(int2) * (size_t)c->np_max));
    }
    CK(cudaMemsetAsync(c->mix_rec, 0, sizeof(double) * (size_t)need, c->stream));
    CK(cudaMemsetAsync(c->mix_alloc, 0, sizeof(unsigned int) * kMaxRanks, c->stream));
  }
  c->mix_grid = b;
  if (c->np > 0 && c->nranks == 1) {      // (the routed exchange computes the box while it routes)
    box_index_kernel<<<nblocks(c->np, 256), 256, 0, c->stream>>>(b, c->time(), c->lon(), c->lat(), c->p(), c->box, c->np);
    CK(cudaGetLastError());
    c->launches++;
  }
}

static MixArgs mix_args(mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  MixArgs A;
  for (int r = 0; r < kMaxRanks; r++) A.rec[r] = nullptr;
  A.rec[0] = c->mix_rec;      // (single rank: the dense records; several ranks use the routed exchange below)
  A.slice = c->mix_slice; A.nmix = c->nmix; A.ngrid = k.mixing_nx * k.mixing_ny * k.mixing_nz;
  A.stride = (c->nmix + 2) / 2 * 2;
  A.box = c->box;
  A.ens = (k.nens > 0 && k.qnt_ens >= 0) ? c->q(k.qnt_ens) : nullptr;
  A.q0 = c->nq ? c->q(0) : nullptr; A.q_stride = c->np_max; A.np = c->np;
  for (int i = 0; i < MPB_MIX_MAXQ; i++) A.iq[i] = i < c->nmix ? c->mix_iq[i] : 0;
  return A;
}

static void mixing_accumulate_all(mpb_ctx *c) {
  REQUIRE(c->mix_total > 0, "mixing: the box records have not been prepared");
  REQUIRE(c->nranks == 1, "internal: attached ranks use the routed exchange");
  if (c->np == 0 || c->nmix == 0) return;
  mix_accumulate_kernel<<<nblocks(c->np, 256), 256, 0, c->stream>>>(mix_args(c));
  CK(cudaGetLastError());
  c->launches++;
}

static void mixing_apply_all(mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  if (c->np > 0 && c->nmix > 0) {
    mix_apply_kernel<<<nblocks(c->np, 256), 256, 0, c->stream>>>(mix_args(c), clim_view(c), c->time(), c->lat(), c->p(), k.mixing_trop,
                                                                 k.mixing_strat, k.met_coord_type == 0, k.met_utm_ref_lat);
    CK(cudaGetLastError());
    c->launches++;
  }
}

// Single rank: zero the records of the boxes this call's parcels fell into, which are all the records that are not zero,
// so that the next call finds clean records without a memset of the whole grid (reference default: 5.8 M boxes = 93 MB
// for one quantity, against 1.25 M parcels in configs[4]'s share).
static void mixing_clear_touched(mpb_ctx *c) {
  if (c->np > 0 && c->nmix > 0) {
    mix_clear_kernel<<<nblocks(c->np, 256), 256, 0, c->stream>>>(mix_args(c));
    CK(cudaGetLastError());
    c->launches++;
  }
  c->mix_clean = (c->nmix + 2) / 2 * 2 * c->mix_total;
}

// the three phases of the routed exchange (several ranks); a barrier over the ranks separates them
static void mixing_route(mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  RouteArgs A;
  for (int r = 0; r < kMaxRanks; r++) {
    A.inbox[r] = r < c->nranks ? route_inbox(c, r, c->rank) : nullptr;
    A.counts_at[r] = r < c->nranks ? route_counts(c, r) + c->rank : nullptr;
  }
  A.grid = c->mix_grid;
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p();
  A.alloc = c->mix_alloc; A.route = c->mix_route;
  A.slice = c->mix_slice; A.np = c->np; A.q_stride = c->np_max;
  A.nmix = c->nmix; A.ngrid = k.mixing_nx * k.mixing_ny * k.mixing_nz; A.nranks = c->nranks; A.E = c->nmix + 1;
  A.ens = (k.nens > 0 && k.qnt_ens >= 0) ? c->q(k.qnt_ens) : nullptr;
  A.q0 = c->nq ? c->q(0) : nullptr;
  for (int i = 0; i < MPB_MIX_MAXQ; i++) A.iq[i] = i < c->nmix ? c->mix_iq[i] : 0;
  if (c->np > 0 && c->nmix > 0) {
    mix_route_kernel<<<nblocks(c->np, kRouteBlock), kRouteBlock, 0, c->stream>>>(A);
    CK(cudaGetLastError());
    c->launches++;
  }
  mix_publish_kernel<<<1, 32, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}
static void mixing_serve(mpb_ctx *c) {
  ServeArgs A;
  A.inbox = route_inbox(c, c->rank, 0);
  A.counts = route_counts(c, c->rank);
  for (int r = 0; r < kMaxRanks; r++) A.outbox_at[r] = r < c->nranks ? route_outbox(c, r, c->rank) : nullptr;
  A.rec = c->mix_rec; A.cap = route_cap(c); A.E = c->nmix + 1;
  if (A.cap <= 0 || c->nmix == 0) return;
  // every sender holds about 1 / nranks of its parcels' boxes in this rank's slice: size the grid for twice that share
  const dim3 grid(std::max(1u, std::min(nblocks(A.cap, 256), nblocks(2 * A.cap / c->nranks + 1, 256))), (unsigned)c->nranks);
  mix_fold_kernel<<<grid, 256, 0, c->stream>>>(A);
  mix_answer_kernel<<<grid, 256, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches += 2;
}
static void mixing_apply_routed(mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  if (c->np == 0 || c->nmix == 0) return;
  MixArgs A = mix_args(c);
  mix_apply_routed_kernel<<<nblocks(c->np, 256), 256, 0, c->stream>>>(
      route_outbox(c, c->rank, 0), route_cap(c), c->nmix + 1, c->mix_route, c->nmix, A, clim_view(c), c->time(), c->lat(), c->p(),
      k.mixing_trop, k.mixing_strat, k.met_coord_type == 0, k.met_utm_ref_lat);
  CK(cudaGetLastError());
  c->launches++;
}

// module_mixing on one context: alone, or as one rank of a multi-process run (barriers in stream order between the phases)
static void mixing_inline(mpb_ctx *c, double t) {
  // MPTRAC_B200_TRACE_MIXING=1 (diagnostics): events around the phases, read back after a stream synchronisation per call;
  // the averages are printed by mpb_destroy.  Phases: prepare | route (accumulate) | barrier | serve | barrier | apply
  static const bool trace = std::getenv("MPTRAC_B200_TRACE_MIXING") != nullptr;
  cudaEvent_t ev[7];
  int ne = 0;
  auto mark = [&]() {
    if (!trace) return;
    CK(cudaEventCreate(&ev[ne]));
    CK(cudaEventRecord(ev[ne], c->stream));
    ne++;
  };
  mark();
  mixing_prepare(c, t);
  mark();
  if (c->nranks == 1) {
    mixing_accumulate_all(c);
    mark(); mark(); mark(); mark();
    mixing_apply_all(c);
    mixing_clear_touched(c);
  } else {
    mixing_route(c);
    mark();
    peer_barrier(c);
    mark();
    mixing_serve(c);
    mark();
    peer_barrier(c);
    mark();
    mixing_apply_routed(c);
  }
  mark();
  if (trace) {
    CK(cudaStreamSynchronize(c->stream));
    for (int i = 0; i < 6; i++) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
      c->mix_trace_ms[i] += ms;
    }
    for (int i = 0; i < 7; i++) CK(cudaEventDestroy(ev[i]));
    c->mix_trace_n++;
  }
}

static bool hits(double t, double every) { return std::fmod(t, every) == 0; }

static bool meteo_wanted(const mpb_ctl_t &k) {
  for (int i = 0; i < MPB_NMETEO; i++) if (k.qnt_meteo[i] >= 0) return true;
  return false;
}
static bool meteo_wanted(const mpb_ctl_t &k, int first, int last) {
  for (int i = first; i <= last; i++) if (k.qnt_meteo[i] >= 0) return true;
  return false;
}

static bool convection_enabled(const mpb_ctl_t &k) { return k.conv_mix_pbl || k.conv_cape >= 0; }   // src/mptrac.c:7905
static bool decay_enabled(const mpb_ctl_t &k) { return k.tdec_trop > 0 && k.tdec_strat > 0; }        // :7938

static void launch_diff_pbl(mpb_ctx *c) {
  REQUIRE(c->ctl.rng_type == 1, "only RNG_TYPE 1 (Squares) runs on the device");
  const unsigned long long ctr = rng_draw(c);   // module_rng(3 np, normal)
  if (c->np == 0) return;
  PblArgs A;
  A.met = met_view(c);
  for (int f : {MPB_F2_ESS, MPB_F2_NSS, MPB_F2_SHF})
    REQUIRE(c->x2[f] && c->x2_valid[0][f] && c->x2_valid[1][f],
            "module_diff_pbl needs the met fields ess, nss and shf of both levels (mpb_met_view_t::x2)");
  REQUIRE(c->x3[MPB_F3_H2O] && c->x3_valid[0][MPB_F3_H2O] && c->x3_valid[1][MPB_F3_H2O],
          "module_diff_pbl needs the met field h2o of both levels (mpb_met_view_t::x3)");
  A.f.ess = c->x2[MPB_F2_ESS]; A.f.nss = c->x2[MPB_F2_NSS]; A.f.shf = c->x2[MPB_F2_SHF]; A.f.h2o = c->x3[MPB_F3_H2O];
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p(); A.dt = c->dt; A.uvwp = c->uvwp;
  A.ctr = ctr; A.ig0 = c->ig0; A.np = c->np; A.slot = c->scrambled ? c->slot : nullptr;
  diff_pbl_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

static void launch_convection(mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  REQUIRE(k.rng_type == 1, "only RNG_TYPE 1 (Squares) runs on the device");
  // one module_rng(np, uniform) call: np + 1 counters, whether or not any parcel is active
  const unsigned long long ctr = c->rng_ctr;
  c->rng_ctr += (unsigned long long)(c->global_np >= 0 ? c->global_np : c->np) + 1ull;
  if (c->np == 0) return;
  ConvArgs A;
  A.met = met_view(c);
  A.conv.cape = k.conv_cape; A.conv.cin = k.conv_cin; A.conv.pbl_trans = k.conv_pbl_trans; A.conv.mix_pbl = k.conv_mix_pbl;
  A.conv.fcape = A.conv.fcin = A.conv.fpel = nullptr;
  if (k.conv_cape >= 0) {
    for (int f : {MPB_F2_CAPE, MPB_F2_CIN, MPB_F2_PEL})
      REQUIRE(c->x2[f] && c->x2_valid[0][f] && c->x2_valid[1][f],
              "module_convection with CONV_CAPE >= 0 needs the met fields cape, cin and pel of both levels (mpb_met_view_t::x2)");
    A.conv.fcape = c->x2[MPB_F2_CAPE]; A.conv.fcin = c->x2[MPB_F2_CIN]; A.conv.fpel = c->x2[MPB_F2_PEL];
  }
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p(); A.dt = c->dt;
  A.ctr = ctr; A.ig0 = c->ig0; A.np = c->np; A.slot = c->scrambled ? c->slot : nullptr;
  convection_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

static bool isosurf_enabled(const mpb_ctl_t &k) { return k.isosurf >= 1 && k.isosurf <= 4; }   // src/mptrac.c:7866, 7914

static void launch_isosurf(mpb_ctx *c, bool init) {
  const mpb_ctl_t &k = c->ctl;
  if (c->np == 0 || (init && k.isosurf == 4)) return;   // (the balloon series arrives through mpb_set_balloon)
  if (!c->iso_var) {
    CK(cudaMalloc(&c->iso_var, sizeof(double) * (size_t)std::max<long long>(c->np_max, 1)));
    CK(cudaMemsetAsync(c->iso_var, 0, sizeof(double) * (size_t)std::max<long long>(c->np_max, 1), c->stream));
  }
  if (k.isosurf == 4) REQUIRE(c->iso_n >= 1, "ISOSURF 4 needs the balloon pressure series (mpb_set_balloon)");
  IsoArgs A;
  A.met = met_view(c);
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p(); A.iso_var = c->iso_var;
  A.ts = c->iso_ts; A.ps = c->iso_ps; A.n = c->iso_n;
  A.np = c->np; A.mode = k.isosurf; A.init = init ? 1 : 0;
  isosurf_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

static void launch_chem_grid(mpb_ctx *c, double t) {
  const mpb_ctl_t &k = c->ctl;
  if (c->np == 0 || k.qnt_m < 0 || k.qnt_Cx < 0) return;   // src/mptrac.c:3896-3897
  REQUIRE(k.met_coord_type == 0, "module_chem_grid supports the lat/lon grid only");
  REQUIRE(k.molmass > 0, "module_chem_grid: molar mass is not defined");
  REQUIRE(k.qnt_m < c->nq && k.qnt_Cx < c->nq && (k.nens <= 0 || (k.qnt_ens >= 0 && k.qnt_ens < c->nq)), "chemistry-grid quantity index out of range");
  REQUIRE(k.chemgrid_nx > 0 && k.chemgrid_ny > 0 && k.chemgrid_nz > 0, "bad chemistry grid");
  const long long ngrid = (long long)k.chemgrid_nx * k.chemgrid_ny * k.chemgrid_nz, total = ngrid * (k.nens > 0 ? k.nens : 1);
  REQUIRE(total < (1ll << 31), "chemistry grid too large");
  ensure_boxes(c);
  if (total > c->chem_cap) {
    if (c->chem_mass) CK(cudaFree(c->chem_mass));
    CK(cudaMalloc(&c->chem_mass, sizeof(double) * (size_t)total));
    c->chem_cap = total;
  }
  CK(cudaMemsetAsync(c->chem_mass, 0, sizeof(double) * (size_t)total, c->stream));
  ChemArgs A;
  A.met = met_view(c);
  A.k.lon0 = k.chemgrid_lon0; A.k.lon1 = k.chemgrid_lon1; A.k.lat0 = k.chemgrid_lat0; A.k.lat1 = k.chemgrid_lat1;
  A.k.z0 = k.chemgrid_z0; A.k.z1 = k.chemgrid_z1;
  A.k.dlon = (k.chemgrid_lon1 - k.chemgrid_lon0) / k.chemgrid_nx; A.k.dlat = (k.chemgrid_lat1 - k.chemgrid_lat0) / k.chemgrid_ny;
  A.k.dz = (k.chemgrid_z1 - k.chemgrid_z0) / k.chemgrid_nz;
  A.k.t0 = t - 0.5 * k.dt_mod; A.k.t1 = t + 0.5 * k.dt_mod; A.k.tt = t; A.k.molmass = k.molmass;
  A.k.nx = k.chemgrid_nx; A.k.ny = k.chemgrid_ny; A.k.nz = k.chemgrid_nz;
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p();
  A.m = c->q(k.qnt_m); A.cx = c->q(k.qnt_Cx); A.ens = k.nens > 0 ? c->q(k.qnt_ens) : nullptr;
  A.mass = c->chem_mass; A.box = c->box; A.np = c->np; A.ngrid = (int)ngrid;
  chem_mass_kernel<<<nblocks(c->np, 256), 256, 0, c->stream>>>(A);
  chem_apply_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches += 2;
}

static bool bound_enabled(const mpb_ctl_t &k) { return k.bound_lat0 < k.bound_lat1 && k.bound_p0 > k.bound_p1; }   // src/mptrac.c:7926

static void launch_bound_cond(mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  if (c->np == 0) return;
  // (the reference tests qnt_Cccl4 for truth, not for >= 0: src/mptrac.c:3802)
  if (k.qnt_m < 0 && k.qnt_vmr < 0 && k.qnt_cts[0] && k.qnt_cts[1] < 0 && k.qnt_cts[2] < 0 && k.qnt_cts[3] < 0 && k.qnt_cts[4] < 0 &&
      k.qnt_aoa < 0)
    return;
  auto row = [&](int iq) -> double * {
    if (iq < 0) return nullptr;
    REQUIRE(iq < c->nq, "boundary-condition quantity index out of range");
    return c->q(iq);
  };
  BoundArgs A;
  A.met = met_view(c);
  A.k.lat0 = k.bound_lat0; A.k.lat1 = k.bound_lat1; A.k.p0 = k.bound_p0; A.k.p1 = k.bound_p1;
  A.k.dps = k.bound_dps; A.k.dzs = k.bound_dzs; A.k.zetas = k.bound_zetas; A.k.pbl = k.bound_pbl;
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p(); A.dt = c->dt;
  A.m = row(k.qnt_m); A.vmr = row(k.qnt_vmr); A.aoa = row(k.qnt_aoa);
  for (int i = 0; i < 5; i++) {
    const bool on = k.qnt_cts[i] >= 0 && ((k.cts_on >> i) & 1);
    REQUIRE(!on || c->cts_n[i] >= 1, "module_bound_cond needs the time series of a trace gas (mpb_set_clim_ts)");
    A.cts[i] = on ? row(k.qnt_cts[i]) : nullptr;
    A.cts_time[i] = c->cts_time[i]; A.cts_vmr[i] = c->cts_vmr[i]; A.cts_n[i] = c->cts_n[i];
  }
  A.mass = k.bound_mass; A.mass_trend = k.bound_mass_trend; A.vmr0 = k.bound_vmr; A.vmr_trend = k.bound_vmr_trend;
  A.np = c->np;
  bound_cond_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

static void launch_decay(mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  const bool decay = decay_enabled(k);
  if (c->np == 0 || (!decay && k.qnt_loss_rate < 0)) return;
  auto row = [&](int iq) -> double * {
    if (iq < 0) return nullptr;
    REQUIRE(iq < c->nq, "decay quantity index out of range");
    return c->q(iq);
  };
  if (decay) {
    REQUIRE(k.qnt_m >= 0 || k.qnt_vmr >= 0, "module_decay needs quantity mass or volume mixing ratio");   // src/mptrac.c:4237
    REQUIRE(c->cl_tropo != nullptr, "module_decay needs the tropopause climatology (mpb_set_clim_tropo)");
  }
  DecayArgs A;
  A.clim = clim_view(c);
  A.time = c->time(); A.lat = c->lat(); A.p = c->p(); A.dt = c->dt;
  A.m = row(k.qnt_m); A.vmr = row(k.qnt_vmr); A.mloss = row(k.qnt_mloss_decay); A.loss_rate = row(k.qnt_loss_rate);
  A.tdec_trop = k.tdec_trop; A.tdec_strat = k.tdec_strat; A.utm_ref_lat = k.met_utm_ref_lat;
  A.np = c->np; A.coord_type = k.met_coord_type; A.decay = decay ? 1 : 0;
  decay_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

static void launch_meteo_fields(mpb_ctx *c) {
  const mpb_ctl_t &k = c->ctl;
  MeteoFieldArgs A;
  A.met = met_view(c);
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p();
  A.q = c->nq ? c->q(0) : nullptr; A.q_stride = c->np_max; A.np = c->np;
  for (int i = 0; i < MPB_METEO_SLOTS; i++) {
    A.qnt[i] = i < MPB_NMETEO ? k.qnt_meteo[i] : -1;
    REQUIRE(A.qnt[i] < c->nq, "meteo quantity index out of range");
  }
  A.moist = meteo_wanted(k, MPB_Q_PW, MPB_Q_TICE) ? 1 : 0;
  for (int f = 0; f < MPB_NX2; f++) {
    const bool want = k.qnt_meteo[MPB_Q_TS + f] >= 0;
    REQUIRE(!want || (c->x2[f] && c->x2_valid[0][f] && c->x2_valid[1][f]),
            "a module_meteo quantity needs a 2-D met field that was not given for both levels (mpb_met_view_t::x2)");
    A.x2[f] = want ? c->x2[f] : nullptr;
  }
  for (int f = 0; f < MPB_NX3; f++) {
    const bool want = k.qnt_meteo[MPB_Q_ZG + f] >= 0 || (f == MPB_F3_H2O && A.moist);
    REQUIRE(!want || (c->x3[f] && c->x3_valid[0][f] && c->x3_valid[1][f]),
            "a module_meteo quantity needs a 3-D met field that was not given for both levels (mpb_met_view_t::x3)");
    A.x3[f] = want ? c->x3[f] : nullptr;
  }
  meteo_fields_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

static void launch_meteo(mpb_ctx *c) {
  REQUIRE(c->have_ctl, "mpb_set_ctl has not been called");
  if (c->np == 0 || !meteo_wanted(c->ctl)) return;
  if (meteo_wanted(c->ctl, MPB_Q_TS, MPB_NMETEO - 1)) launch_meteo_fields(c);
  if (!meteo_wanted(c->ctl, MPB_Q_PS, MPB_Q_ZETA_D)) return;
  MeteoArgs A;
  A.met = met_view(c);
  A.time = c->time(); A.lon = c->lon(); A.lat = c->lat(); A.p = c->p();
  A.q = c->nq ? c->q(0) : nullptr; A.q_stride = c->np_max; A.np = c->np;
  for (int i = 0; i < MPB_METEO_SLOTS; i++) {
    A.qnt[i] = i < MPB_NMETEO ? c->ctl.qnt_meteo[i] : -1;
    REQUIRE(A.qnt[i] < c->nq, "meteo quantity index out of range");
  }
  meteo_kernel<<<nblocks(c->np, 128), 128, 0, c->stream>>>(A);
  CK(cudaGetLastError());
  c->launches++;
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
// The node array as a tensor for bulk tensor copies (tile form of the step kernel): float32 [nx][ny][nz][8], box = a window of
// TX x TY x TZ whole nodes.  MPTRAC_B200_TILE="tx,ty,tz" overrides the window (default 3 x 4 x min(nz, 32): 12 KB).
static void build_tensor_map(mpb_ctx *c) {
  c->tmap_ok = false;
  if (!c->tile || !c->nodes) return;
  int tx = 3, ty = 4, tz = std::min(c->nz, 32);
  if (const char *e = std::getenv("MPTRAC_B200_TILE")) REQUIRE(std::sscanf(e, "%d,%d,%d", &tx, &ty, &tz) == 3, "MPTRAC_B200_TILE must be tx,ty,tz");
  REQUIRE(tx >= 2 && ty >= 2 && tz >= 2 && tx <= 256 && ty <= 256 && tz <= 256, "tile window out of range");
  tz = std::min(tz, c->nz);
  REQUIRE((size_t)tx * ty * tz * sizeof(Node) <= 200 * 1024, "tile window does not fit shared memory");
  typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr));
  REQUIRE(fn != nullptr && qr == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled is not available in this driver");
  const cuuint64_t dims[4] = {8, (cuuint64_t)c->nz, (cuuint64_t)c->ny, (cuuint64_t)c->nx};
  const cuuint64_t strides[3] = {sizeof(Node), sizeof(Node) * (cuuint64_t)c->nz, sizeof(Node) * (cuuint64_t)c->nz * (cuuint64_t)c->ny};
  const cuuint32_t box[4] = {8, (cuuint32_t)tz, (cuuint32_t)ty, (cuuint32_t)tx};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = ((encode_fn)fn)(&c->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, c->nodes, dims, strides, box, estr,
                                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  c->tshape = TileShape{tx, ty, tz};
  c->tmap_ok = true;
}

// Make the context's met arrays, axis tables and (when `staging`) upload buffers fit a grid; a grid that differs from the
// current one invalidates both met levels (src/mptrac.c:6545-6558 demands identical grids).
static void ensure_grid(mpb_ctx *c, int nx, int ny, int np, int coord_type, const double *lon, const double *lat, const double *p,
                        bool staging) {
  const size_t nnode = (size_t)nx * ny * np, ncol = (size_t)nx * ny;
  bool same_axes = (nx == c->nx && ny == c->ny && np == c->nz && coord_type == c->coord_type);
  if (same_axes)
    same_axes = std::equal(lon, lon + nx, c->h_lon.begin()) && std::equal(lat, lat + ny, c->h_lat.begin()) &&
                std::equal(p, p + np, c->h_p.begin());
  if (staging && 4 * nnode > c->stage_cap) {
    if (c->stage_h) CK(cudaFreeHost(c->stage_h));
    if (c->stage_d) CK(cudaFree(c->stage_d));
    CK(cudaMallocHost(&c->stage_h, sizeof(float) * 4 * nnode));
    CK(cudaMalloc(&c->stage_d, sizeof(float) * 4 * nnode));
    c->stage_cap = 4 * nnode;
  }
  if (same_axes) return;
  c->lev[0].valid = c->lev[1].valid = false;
  c->lev_valid[0] = c->lev_valid[1] = false;
  for (int f = 0; f < MPB_NX2; f++) { if (c->x2[f]) CK(cudaFree(c->x2[f])); c->x2[f] = nullptr; c->x2_valid[0][f] = c->x2_valid[1][f] = false; }
  for (int f = 0; f < MPB_NX3; f++) { if (c->x3[f]) CK(cudaFree(c->x3[f])); c->x3[f] = nullptr; c->x3_valid[0][f] = c->x3_valid[1][f] = false; }
  c->nx = nx; c->ny = ny; c->nz = np; c->coord_type = coord_type;
  if (nnode > c->node_cap) {
    if (c->nodes) CK(cudaFree(c->nodes));
    CK(cudaMalloc(&c->nodes, sizeof(Node) * nnode));
    c->node_cap = nnode;
  }
  if (ncol > c->col_cap) {
    if (c->surf) CK(cudaFree(c->surf));
    CK(cudaMalloc(&c->surf, sizeof(float4) * ncol));
    c->col_cap = ncol;
  }
  c->h_lon.assign(lon, lon + nx);
  c->h_lat.assign(lat, lat + ny);
  c->h_p.assign(p, p + np);
  c->tables = build_axis_tables(c->h_lon.data(), nx, c->h_lat.data(), ny, c->h_p.data(), np);
  void **olds[] = {(void **)&c->ax_lon, (void **)&c->ax_lat, (void **)&c->ax_p, (void **)&c->ax_lonc,
                   (void **)&c->ax_latc, (void **)&c->ax_pc, (void **)&c->p_lut};
  for (void **o : olds) if (*o) { CK(cudaFree(*o)); *o = nullptr; }
  auto up = [&](auto **dst, const auto &vec) {
    using T = typename std::remove_reference<decltype(vec)>::type::value_type;
    CK(cudaMalloc((void **)dst, sizeof(T) * std::max<size_t>(vec.size(), 1)));
    CK(cudaMemcpyAsync(*dst, vec.data(), sizeof(T) * vec.size(), cudaMemcpyHostToDevice, c->stream));
  };
  up(&c->ax_lon, c->h_lon); up(&c->ax_lat, c->h_lat); up(&c->ax_p, c->h_p);
  up(&c->ax_lonc, c->tables.lonc); up(&c->ax_latc, c->tables.latc); up(&c->ax_pc, c->tables.pc);
  up(&c->p_lut, c->tables.p_lut);
  CK(cudaStreamSynchronize(c->stream));
  build_tensor_map(c);
}

extern "C" {

const char *mpb_last_error(void) { return g_err.c_str(); }
int mpb_abi_version(void) { return MPB_ABI_VERSION; }

int mpb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n;
}

// Create the CUDA context of `device` ahead of time (a few hundred milliseconds on a cold process).  Thread-safe; a host
// driver can call it from a helper thread while it is still reading its input files.
int mpb_warmup(int device) {
  API_BEGIN
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) { cudaGetLastError(); throw std::runtime_error("no CUDA device: mptrac_b200 has no CPU fallback"); }
  REQUIRE(device >= 0 && device < ndev, "device index out of range");
  CK(cudaSetDevice(device));
  CK(cudaFree(nullptr));
  API_END
}

int mpb_create(mpb_ctx **out, int device, int64_t np_max, int nq) {
  API_BEGIN
  REQUIRE(out != nullptr, "null output pointer");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    throw std::runtime_error("no CUDA device: mptrac_b200 has no CPU fallback");
  }
  REQUIRE(device >= 0 && device < ndev, "device index out of range");
  REQUIRE(np_max >= 0 && np_max < (1ll << 31), "np_max out of range (parcel indices are int)");
  REQUIRE(nq >= 0, "nq must be >= 0");
  CK(cudaSetDevice(device));
  mpb_ctx *c = new mpb_ctx();
  // capacity in whole 32-parcel tiles: every array then starts 256-byte aligned and a warp's bulk copy of its (possibly
  // partial) last tile stays inside the allocation
  c->device = device; c->np_max = (std::max<long long>(np_max, 1) + 31) / 32 * 32; c->nq = nq;
  std::memset(&c->ctl, 0, sizeof(c->ctl));
  if (const char *e = std::getenv("MPTRAC_B200_STEP")) { c->quad = std::strcmp(e, "quad") == 0; c->tile = std::strcmp(e, "tile") == 0; }
  if (const char *e = std::getenv("MPTRAC_B200_QUAD_SPLIT")) c->quad_split = std::atoi(e) != 0;
  CK(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
  c->stream = c->own_stream;
  CK(cudaMalloc(&c->soa[0], sizeof(double) * (size_t)c->np_max * (size_t)(4 + nq)));
  CK(cudaMalloc(&c->dt, sizeof(double) * (size_t)c->np_max));
  CK(cudaMalloc(&c->uvwp, sizeof(float) * 3 * (size_t)c->np_max));
  CK(cudaMemsetAsync(c->dt, 0, sizeof(double) * (size_t)c->np_max, c->stream));
  CK(cudaMemsetAsync(c->uvwp, 0, sizeof(float) * 3 * (size_t)c->np_max, c->stream));
  *out = c;
  API_END
}

int mpb_destroy(mpb_ctx *c) {
  API_BEGIN
  if (!c) return 0;
  use(c);
  CK(cudaStreamSynchronize(c->stream));
  if (c->mix_trace_n > 0) {
    const double n = (double)c->mix_trace_n;
    const double *m = c->mix_trace_ms;
    char line[512];
    std::snprintf(line, sizeof(line), "[mptrac_b200] mixing trace, rank %d of %d, %lld calls, ms per call: prepare %.4f  route %.4f  barrier %.4f  "
                  "serve %.4f  barrier %.4f  apply %.4f  (sum %.4f)\n", c->rank, c->nranks, c->mix_trace_n, m[0] / n, m[1] / n, m[2] / n, m[3] / n,
                  m[4] / n, m[5] / n, (m[0] + m[1] + m[2] + m[3] + m[4] + m[5]) / n);
    std::fputs(line, stderr);     // (one write: the ranks of a run share the terminal)
  }
  void *ptrs[] = {c->soa[0], c->soa[1], c->dt, c->uvwp, c->keys[0], c->keys[1], c->perm[0], c->perm[1],
                  c->cub_tmp, c->nodes, c->surf, c->ax_lon, c->ax_lat, c->ax_p, c->ax_lonc, c->ax_latc, c->ax_pc,
                  c->p_lut, c->stage_d, c->cl_time, c->cl_lat, c->cl_tropo, c->box, c->mix_rec, c->mix_alloc, c->mix_route,
                  c->grid_in_area ? nullptr : c->grid_sum, c->grid_in_area ? nullptr : c->grid_sq, c->grid_in_area ? nullptr : c->grid_cnt, c->lev_p, c->lev_z, c->lev_pz, c->lev_hint, c->iso_var, c->iso_ts, c->iso_ps, c->chem_mass,
                  c->slot, c->slot2, c->inv, c->perm2, c->src, c->uvwp2, c->dt2, c->iso_var2, c->lev_hint2};
  for (void *p : ptrs) if (p) cudaFree(p);
  for (int r = 0; r < kMaxRanks; r++)
    if (c->area[r]) { if (r == c->rank) cudaFree(c->area[r]); else if (c->area_ipc[r]) cudaIpcCloseMemHandle(c->area[r]); }
  for (double *p : c->cts_time) if (p) cudaFree(p);
  for (double *p : c->cts_vmr) if (p) cudaFree(p);
  for (float2 *p : c->x2) if (p) cudaFree(p);
  for (float2 *p : c->x3) if (p) cudaFree(p);
  if (c->stage_h) cudaFreeHost(c->stage_h);
  if (c->grid_stage_h) cudaFreeHost(c->grid_stage_h);
  for (int i = 0; i < kLanes; i++) {
    if (c->lane[i]) cudaStreamDestroy(c->lane[i]);
    if (c->lane_done[i]) cudaEventDestroy(c->lane_done[i]);
  }
  if (c->lane_go) cudaEventDestroy(c->lane_go);
  cudaStreamDestroy(c->own_stream);
  delete c;
  API_END
}

int mpb_set_stream(mpb_ctx *c, void *s) {
  API_BEGIN
  use(c);
  CK(cudaStreamSynchronize(c->stream));
  c->stream = s ? (cudaStream_t)s : c->own_stream;
  API_END
}

int mpb_sync(mpb_ctx *c) {
  API_BEGIN
  use(c);
  CK(cudaStreamSynchronize(c->stream));
  if (c->peer_err && c->epoch > 0) {
    int err = 0;
    CK(cudaMemcpy(&err, c->peer_err, sizeof(int), cudaMemcpyDeviceToHost));
    REQUIRE(err == 0, "a barrier over the ranks timed out: a peer did not reach the exchange step");
  }
  API_END
}

int mpb_set_ctl(mpb_ctx *c, const mpb_ctl_t *ctl) {
  API_BEGIN
  use(c);
  REQUIRE(ctl != nullptr, "null ctl");
  REQUIRE(ctl->nq == c->nq, "ctl.nq differs from the context's nq");
  REQUIRE(ctl->advect == 0 || ctl->advect == 1 || ctl->advect == 2 || ctl->advect == 4, "ADVECT must be 0, 1, 2 or 4");
  REQUIRE(ctl->n_mix_qnt >= 0 && ctl->n_mix_qnt <= MPB_MIX_MAXQ, "n_mix_qnt out of range");
  for (int i = 0; i < MPB_NMETEO; i++) REQUIRE(ctl->qnt_meteo[i] < ctl->nq, "meteo quantity index out of range");
  c->ctl = *ctl;
  c->have_ctl = true;
  API_END
}

int mpb_set_clim_tropo(mpb_ctx *c, int ntime, int nlat, const double *time, const double *lat, const double *tropo) {
  API_BEGIN
  use(c);
  REQUIRE(ntime >= 2 && nlat >= 2 && time && lat && tropo, "bad tropopause table");
  if (c->cl_time) { CK(cudaFree(c->cl_time)); CK(cudaFree(c->cl_lat)); CK(cudaFree(c->cl_tropo)); }
  CK(cudaMalloc(&c->cl_time, sizeof(double) * ntime));
  CK(cudaMalloc(&c->cl_lat, sizeof(double) * nlat));
  CK(cudaMalloc(&c->cl_tropo, sizeof(double) * (size_t)ntime * nlat));
  CK(cudaMemcpyAsync(c->cl_time, time, sizeof(double) * ntime, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->cl_lat, lat, sizeof(double) * nlat, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->cl_tropo, tropo, sizeof(double) * (size_t)ntime * nlat, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->cl_ntime = ntime; c->cl_nlat = nlat;
  API_END
}

int mpb_set_met(mpb_ctx *c, int slot, const mpb_met_view_t *m) {
  API_BEGIN
  use(c);
  REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  REQUIRE(m && m->lon && m->lat && m->p && m->u && m->v && m->w, "met view lacks axes or wind fields");
  REQUIRE(m->nx >= 2 && m->ny >= 2 && m->np >= 2, "met grid needs at least 2 nodes per axis");
  const size_t nnode = (size_t)m->nx * m->ny * m->np, ncol = (size_t)m->nx * m->ny;
  CK(cudaStreamSynchronize(c->stream));  // staging buffers are reused
  ensure_grid(c, m->nx, m->ny, m->np, m->coord_type, m->lon, m->lat, m->p, true);

  // compact the strided host fields into the pinned staging area (columns are contiguous runs of np floats)
  const float *src3[4] = {m->u, m->v, m->w, m->t};
  const size_t row = sizeof(float) * (size_t)m->np;
  for (int f = 0; f < 4; f++) {
    if (!src3[f]) continue;
    float *dst = c->stage_h + (size_t)f * nnode;
#pragma omp parallel for collapse(2) schedule(static)
    for (int ix = 0; ix < m->nx; ix++)
      for (int iy = 0; iy < m->ny; iy++)
        std::memcpy(dst + ((size_t)ix * m->ny + iy) * m->np, src3[f] + (size_t)ix * m->sx + (size_t)iy * m->sy, row);
    CK(cudaMemcpyAsync(c->stage_d + (size_t)f * nnode, dst, sizeof(float) * nnode, cudaMemcpyHostToDevice, c->stream));
  }
  pack_met_nodes_kernel<<<nblocks((long long)nnode, 256), 256, 0, c->stream>>>(
      c->stage_d, c->stage_d + nnode, c->stage_d + 2 * nnode, m->t ? c->stage_d + 3 * nnode : nullptr, (float *)c->nodes, slot, nnode);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));

  // surface fields reuse the (now idle) staging area
  const float *src2[2] = {m->ps, m->pbl};
  for (int f = 0; f < 2; f++) {
    if (!src2[f]) continue;
    float *dst = c->stage_h + (size_t)f * ncol;
    for (int ix = 0; ix < m->nx; ix++)
      std::memcpy(dst + (size_t)ix * m->ny, src2[f] + (size_t)ix * m->sx2, sizeof(float) * (size_t)m->ny);
    CK(cudaMemcpyAsync(c->stage_d + (size_t)f * ncol, dst, sizeof(float) * ncol, cudaMemcpyHostToDevice, c->stream));
  }
  pack_surface_kernel<<<nblocks((long long)ncol, 256), 256, 0, c->stream>>>(
      m->ps ? c->stage_d : nullptr, m->pbl ? c->stage_d + ncol : nullptr, (float2 *)c->surf, slot, ncol);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->launches += 2;

  // model levels: {pl, ul, vl, wl} and {zetal, ul, vl, zeta_dotl} records plus the {pl, zetal} pairs of the conversions
  c->lev_valid[slot] = false;
  if (m->npl >= 2 && m->pl && m->ul && m->vl) {
    REQUIRE(m->wl || m->zeta_dotl, "model levels need wl or zeta_dotl");
    const size_t nlev = ncol * (size_t)m->npl;
    if (m->npl != c->npl || nlev > c->lev_cap) {
      c->lev_valid[slot ^ 1] = false;
      for (void *o : {(void *)c->lev_p, (void *)c->lev_z, (void *)c->lev_pz}) if (o) CK(cudaFree(o));
      c->lev_p = c->lev_z = nullptr; c->lev_pz = nullptr;
      CK(cudaMalloc(&c->lev_p, sizeof(LevelNode) * nlev));
      CK(cudaMalloc(&c->lev_z, sizeof(LevelNode) * nlev));
      CK(cudaMalloc(&c->lev_pz, sizeof(float4) * nlev));
      CK(cudaMemsetAsync(c->lev_p, 0, sizeof(LevelNode) * nlev, c->stream));
      CK(cudaMemsetAsync(c->lev_z, 0, sizeof(LevelNode) * nlev, c->stream));
      CK(cudaMemsetAsync(c->lev_pz, 0, sizeof(float4) * nlev, c->stream));
      c->lev_cap = nlev; c->npl = m->npl;
    }
    if (6 * nlev > c->stage_cap) {
      CK(cudaFreeHost(c->stage_h)); CK(cudaFree(c->stage_d));
      c->stage_h = nullptr; c->stage_d = nullptr;
      CK(cudaMallocHost(&c->stage_h, sizeof(float) * 6 * nlev));
      CK(cudaMalloc(&c->stage_d, sizeof(float) * 6 * nlev));
      c->stage_cap = 6 * nlev;
    }
    const float *srcl[6] = {m->pl, m->ul, m->vl, m->wl, m->zetal, m->zeta_dotl};
    const size_t lrow = sizeof(float) * (size_t)m->npl;
    for (int f = 0; f < 6; f++) {
      float *dst = c->stage_h + (size_t)f * nlev;
      if (srcl[f]) {
#pragma omp parallel for collapse(2) schedule(static)
        for (int ix = 0; ix < m->nx; ix++)
          for (int iy = 0; iy < m->ny; iy++)
            std::memcpy(dst + ((size_t)ix * m->ny + iy) * m->npl, srcl[f] + (size_t)ix * m->sxl + (size_t)iy * m->syl, lrow);
      } else {
        std::memset(dst, 0, sizeof(float) * nlev);
      }
      CK(cudaMemcpyAsync(c->stage_d + (size_t)f * nlev, dst, sizeof(float) * nlev, cudaMemcpyHostToDevice, c->stream));
    }
    float *d = c->stage_d;
    const unsigned gl = nblocks((long long)nlev, 256);
    pack_nodes_kernel<<<gl, 256, 0, c->stream>>>(d, d + nlev, d + 2 * nlev, d + 3 * nlev, (float4 *)c->lev_p, slot, nlev);
    pack_nodes_kernel<<<gl, 256, 0, c->stream>>>(d + 4 * nlev, d + nlev, d + 2 * nlev, d + 5 * nlev, (float4 *)c->lev_z, slot, nlev);
    pack_surface_kernel<<<gl, 256, 0, c->stream>>>(d, d + 4 * nlev, (float2 *)c->lev_pz, slot, nlev);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->launches += 3;
    c->lev_valid[slot] = true;
  }
  // further fields for module_meteo: one interleaved array per field, filled level by level
  auto put_field = [&](const float *src, float2 **dst, size_t n, bool three_d) {
    if (!*dst) {
      CK(cudaMalloc(dst, sizeof(float2) * n));
      CK(cudaMemsetAsync(*dst, 0, sizeof(float2) * n, c->stream));
    }
    float *h = c->stage_h;
    if (three_d) {
#pragma omp parallel for collapse(2) schedule(static)
      for (int ix = 0; ix < m->nx; ix++)
        for (int iy = 0; iy < m->ny; iy++)
          std::memcpy(h + ((size_t)ix * m->ny + iy) * m->np, src + (size_t)ix * m->sx + (size_t)iy * m->sy, sizeof(float) * (size_t)m->np);
    } else {
      for (int ix = 0; ix < m->nx; ix++)
        std::memcpy(h + (size_t)ix * m->ny, src + (size_t)ix * m->sx2, sizeof(float) * (size_t)m->ny);
    }
    CK(cudaMemcpyAsync(c->stage_d, h, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
    pack_scalar_kernel<<<nblocks((long long)n, 256), 256, 0, c->stream>>>(c->stage_d, (float *)*dst, slot, n);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));   // the staging area is reused by the next field
    c->launches++;
  };
  for (int f = 0; f < MPB_NX2; f++) {
    c->x2_valid[slot][f] = false;
    if (m->x2[f]) { put_field(m->x2[f], &c->x2[f], ncol, false); c->x2_valid[slot][f] = true; }
  }
  for (int f = 0; f < MPB_NX3; f++) {
    c->x3_valid[slot][f] = false;
    if (m->x3[f]) { put_field(m->x3[f], &c->x3[f], nnode, true); c->x3_valid[slot][f] = true; }
  }
  c->lev[slot].time = m->time;
  c->lev[slot].valid = true;
  API_END
}

// A met level straight from one of the reference's binary files (MET_TYPE 1, uncompressed; written by write_met_bin,
// src/mptrac.c:14204 ff., read by read_met_bin, :8887-9181) into the packed device layout.  The file holds dense arrays --
// [nx][ny] and [nx][ny][np], the level index fastest -- i.e. exactly what the pack kernels take: every field goes from the
// file into the pinned staging buffer and from there to the device, without the detour through a 10.7 GB met_t and without
// any strided compaction.  Like the reference, nothing is pre-processed (binary files hold finished fields, :7769-7773), and
// 3-D fields are clamped to read_met_bin_3d's bounds.  all_fields != 0 also uploads the further 2-D / 3-D fields of
// module_meteo and of the boundary-layer / convection modules.
int mpb_set_met_bin(mpb_ctx *c, int slot, const char *path, int all_fields) {
  API_BEGIN
  use(c);
  REQUIRE(slot == 0 || slot == 1, "slot must be 0 or 1");
  REQUIRE(path != nullptr, "null path");
  REQUIRE(c->have_ctl, "mpb_set_ctl comes first (MET_COORD_TYPE)");
  FILE *in = std::fopen(path, "rb");
  REQUIRE(in != nullptr, std::string("cannot open ") + path);
  struct Closer { FILE *f; ~Closer() { if (f) std::fclose(f); } } closer{in};
  auto rd = [&](void *dst, size_t size, size_t count, const char *what) {
    REQUIRE(std::fread(dst, size, count, in) == count, std::string("binary met file ends early (") + what + ")");
  };
  int met_type = 0, version = 0, nx = 0, ny = 0, np = 0;
  double time = 0;
  rd(&met_type, sizeof(int), 1, "type");
  REQUIRE(met_type == 1, "only uncompressed binary met files (MET_TYPE 1) can be read directly");
  rd(&version, sizeof(int), 1, "version");
  REQUIRE(version == 104, "wrong version of binary met data (104 expected)");
  rd(&time, sizeof(double), 1, "time");
  rd(&nx, sizeof(int), 1, "nx"); rd(&ny, sizeof(int), 1, "ny"); rd(&np, sizeof(int), 1, "np");
  REQUIRE(nx >= 2 && ny >= 2 && np >= 2 && nx < (1 << 20) && ny < (1 << 20) && np < (1 << 16), "binary met file: dimensions out of range");
  std::vector<double> lon(nx), lat(ny), p(np);
  rd(lon.data(), sizeof(double), (size_t)nx, "lon"); rd(lat.data(), sizeof(double), (size_t)ny, "lat"); rd(p.data(), sizeof(double), (size_t)np, "p");
  const size_t nnode = (size_t)nx * ny * np, ncol = (size_t)nx * ny;
  CK(cudaStreamSynchronize(c->stream));
  ensure_grid(c, nx, ny, np, c->ctl.met_coord_type, lon.data(), lat.data(), p.data(), true);

  auto further = [&](float2 **dst, size_t n) {          // stage_h[0 .. n) -> one further field (mpb_met_view_t::x2 / x3)
    if (!*dst) {
      CK(cudaMalloc(dst, sizeof(float2) * n));
      CK(cudaMemsetAsync(*dst, 0, sizeof(float2) * n, c->stream));
    }
    CK(cudaMemcpyAsync(c->stage_d, c->stage_h, sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
  };
  // ---- the 24 surface fields, file order of read_met_bin (:8994-9017) ----
  // PS TS ZS US VS ESS NSS SHF LSM SST PBL PT TT ZT H2OT PCT PCB CL PLCL PLFC PEL CAPE CIN O3C
  static const int f2_of_file[24] = {-1, MPB_F2_TS, MPB_F2_ZS, MPB_F2_US, MPB_F2_VS, MPB_F2_ESS, MPB_F2_NSS, MPB_F2_SHF, MPB_F2_LSM, MPB_F2_SST,
                                     -2, MPB_F2_PT, MPB_F2_TT, MPB_F2_ZT, MPB_F2_H2OT, MPB_F2_PCT, MPB_F2_PCB, MPB_F2_CL, MPB_F2_PLCL,
                                     MPB_F2_PLFC, MPB_F2_PEL, MPB_F2_CAPE, MPB_F2_CIN, MPB_F2_O3C};
  float *ps_pbl = c->stage_h + 2 * nnode;                 // kept aside until both have been read
  for (int k = 0; k < 24; k++) {
    const int f = f2_of_file[k];
    if (f < 0) { rd(ps_pbl + (f == -1 ? 0 : ncol), sizeof(float), ncol, "surface field"); continue; }
    if (!all_fields) { REQUIRE(std::fseek(in, (long)(sizeof(float) * ncol), SEEK_CUR) == 0, "binary met file ends early (surface field)"); continue; }
    rd(c->stage_h, sizeof(float), ncol, "surface field");
    further(&c->x2[f], ncol);
    pack_scalar_kernel<<<nblocks((long long)ncol, 256), 256, 0, c->stream>>>(c->stage_d, (float *)c->x2[f], slot, ncol);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->launches++;
    c->x2_valid[slot][f] = true;
  }
  if (!all_fields) for (int f = 0; f < MPB_NX2; f++) c->x2_valid[slot][f] = false;
  CK(cudaMemcpyAsync(c->stage_d, ps_pbl, sizeof(float) * 2 * ncol, cudaMemcpyHostToDevice, c->stream));
  pack_surface_kernel<<<nblocks((long long)ncol, 256), 256, 0, c->stream>>>(c->stage_d, c->stage_d + ncol, (float2 *)c->surf, slot, ncol);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));
  c->launches++;
  // ---- the 13 level fields (:9020-9032): Z T U V W PV H2O O3 LWC RWC IWC SWC CC, with read_met_bin_3d's bounds ----
  static const int f3_of_file[13] = {MPB_F3_Z, -1, -1, -1, -1, MPB_F3_PV, MPB_F3_H2O, MPB_F3_O3, MPB_F3_LWC, MPB_F3_RWC, MPB_F3_IWC, MPB_F3_SWC, MPB_F3_CC};
  static const float lo3[13] = {-1e34f, 0.f, -1e34f, -1e34f, -1e34f, -1e34f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  static const float hi3[13] = {1e34f, 1e34f, 1e34f, 1e34f, 1e34f, 1e34f, 1e34f, 1e34f, 1e34f, 1e34f, 1e34f, 1e34f, 1.f};
  static const int node_slot[13] = {-1, 3, 0, 1, 2, -1, -1, -1, -1, -1, -1, -1, -1};   // T, U, V, W -> staging order u, v, w, t
  const unsigned gn = nblocks((long long)nnode, 256);
  for (int k = 0; k < 13; k++) {
    const int f = f3_of_file[k];
    if (node_slot[k] >= 0) {
      float *h = c->stage_h + (size_t)node_slot[k] * nnode, *d = c->stage_d + (size_t)node_slot[k] * nnode;
      rd(h, sizeof(float), nnode, "level field");
      CK(cudaMemcpyAsync(d, h, sizeof(float) * nnode, cudaMemcpyHostToDevice, c->stream));
      clamp_field_kernel<<<gn, 256, 0, c->stream>>>(d, nnode, lo3[k], hi3[k]);
      CK(cudaGetLastError());
      c->launches++;
      if (k == 4) {      // W was the last of the four: pack the nodes, then the staging area is free again
        pack_met_nodes_kernel<<<gn, 256, 0, c->stream>>>(c->stage_d, c->stage_d + nnode, c->stage_d + 2 * nnode, c->stage_d + 3 * nnode,
                                                         (float *)c->nodes, slot, nnode);
        CK(cudaGetLastError());
        CK(cudaStreamSynchronize(c->stream));
        c->launches++;
      }
      continue;
    }
    if (!all_fields) { REQUIRE(std::fseek(in, (long)(sizeof(float) * nnode), SEEK_CUR) == 0, "binary met file ends early (level field)"); continue; }
    rd(c->stage_h, sizeof(float), nnode, "level field");
    further(&c->x3[f], nnode);
    clamp_field_kernel<<<gn, 256, 0, c->stream>>>(c->stage_d, nnode, lo3[k], hi3[k]);
    pack_scalar_kernel<<<gn, 256, 0, c->stream>>>(c->stage_d, (float *)c->x3[f], slot, nnode);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));
    c->launches += 2;
    c->x3_valid[slot][f] = true;
  }
  if (!all_fields) for (int f = 0; f < MPB_NX3; f++) c->x3_valid[slot][f] = false;
  int final_flag = 0;
  rd(&final_flag, sizeof(int), 1, "final flag");
  REQUIRE(final_flag == 999, "binary met file: final flag missing");
  c->lev_valid[slot] = false;       // (binary files carry no model-level fields)
  c->lev[slot].time = time;
  c->lev[slot].valid = true;
  API_END
}

int mpb_swap_met(mpb_ctx *c) {
  API_BEGIN
  use(c);
  std::swap(c->lev[0], c->lev[1]);
  const size_t nnode = (size_t)c->nx * c->ny * c->nz, ncol = (size_t)c->nx * c->ny;
  if (nnode > 0 && c->nodes) {
    swap_pairs_kernel<<<nblocks((long long)(4 * nnode), 256), 256, 0, c->stream>>>((float2 *)c->nodes, 4 * nnode);   // {x0, x1} -> {x1, x0}
    swap_levels_kernel<<<nblocks((long long)ncol, 256), 256, 0, c->stream>>>(nullptr, 0, (float2 *)c->surf, ncol);
    CK(cudaGetLastError());
    c->launches += 2;
  }
  for (int f = 0; f < MPB_NX2; f++) {
    std::swap(c->x2_valid[0][f], c->x2_valid[1][f]);
    if (c->x2[f]) { swap_pairs_kernel<<<nblocks((long long)ncol, 256), 256, 0, c->stream>>>(c->x2[f], ncol); c->launches++; }
  }
  for (int f = 0; f < MPB_NX3; f++) {
    std::swap(c->x3_valid[0][f], c->x3_valid[1][f]);
    if (c->x3[f]) { swap_pairs_kernel<<<nblocks((long long)nnode, 256), 256, 0, c->stream>>>(c->x3[f], nnode); c->launches++; }
  }
  CK(cudaGetLastError());
  std::swap(c->lev_valid[0], c->lev_valid[1]);
  const size_t nlev = ncol * (size_t)c->npl;
  if (nlev > 0 && c->lev_p) {
    const unsigned gl = nblocks((long long)nlev, 256);
    swap_levels_kernel<<<gl, 256, 0, c->stream>>>((float4 *)c->lev_p, nlev, (float2 *)c->lev_pz, nlev);
    swap_levels_kernel<<<gl, 256, 0, c->stream>>>((float4 *)c->lev_z, nlev, nullptr, 0);
    CK(cudaGetLastError());
    c->launches += 2;
  }
  API_END
}

int mpb_set_atm(mpb_ctx *c, int64_t np, const double *time, const double *p, const double *lon,
                const double *lat, const double *q, int64_t q_stride) {
  API_BEGIN
  use(c);
  unscramble(c);
  REQUIRE(np >= 0 && np <= c->np_max, "np exceeds the context capacity");
  REQUIRE(np == 0 || (time && p && lon && lat), "null parcel array");
  REQUIRE(c->nq == 0 || np == 0 || q != nullptr, "null quantity array");
  c->np = np;
  c->q_stale = false;
  if (c->have_ctl && c->ctl.sort_dt > 0 && np > 0) {      // the sort's buffers now, not at the first sort in the middle of the run
    if (own_order_wanted(c)) ensure_order_buffers(c); else ensure_sort_buffers(c);
  }
  const size_t bytes = sizeof(double) * (size_t)np;
  if (np > 0) {
    CK(cudaMemcpyAsync(c->time(), time, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->p(), p, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->lon(), lon, bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemcpyAsync(c->lat(), lat, bytes, cudaMemcpyHostToDevice, c->stream));
    for (int iq = 0; iq < c->nq; iq++)
      CK(cudaMemcpyAsync(c->q(iq), q + (size_t)iq * q_stride, bytes, cudaMemcpyHostToDevice, c->stream));
  }
  API_END
}

int mpb_set_uvwp(mpb_ctx *c, const float *uvwp) {
  API_BEGIN
  use(c);
  unscramble(c);
  REQUIRE(uvwp != nullptr, "null uvwp");
  if (c->np > 0) CK(cudaMemcpyAsync(c->uvwp, uvwp, sizeof(float) * 3 * (size_t)c->np, cudaMemcpyHostToDevice, c->stream));
  API_END
}

int mpb_set_iso_var(mpb_ctx *c, const double *iso_var) {
  API_BEGIN
  use(c);
  unscramble(c);
  REQUIRE(iso_var != nullptr, "null iso_var");
  if (!c->iso_var) CK(cudaMalloc(&c->iso_var, sizeof(double) * (size_t)std::max<long long>(c->np_max, 1)));
  if (c->np > 0) CK(cudaMemcpyAsync(c->iso_var, iso_var, sizeof(double) * (size_t)c->np, cudaMemcpyHostToDevice, c->stream));
  API_END
}

int mpb_get_iso_var(mpb_ctx *c, double *iso_var) {
  API_BEGIN
  use(c);
  unscramble(c);
  REQUIRE(iso_var != nullptr && c->iso_var != nullptr, "no iso_var on the device");
  if (c->np > 0) CK(cudaMemcpyAsync(iso_var, c->iso_var, sizeof(double) * (size_t)c->np, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END
}

int mpb_set_clim_ts(mpb_ctx *c, int species, int n, const double *time, const double *vmr) {
  API_BEGIN
  use(c);
  REQUIRE(species >= 0 && species < 5 && n >= 1 && time && vmr, "bad time series");
  CK(cudaStreamSynchronize(c->stream));
  if (c->cts_time[species]) { CK(cudaFree(c->cts_time[species])); CK(cudaFree(c->cts_vmr[species])); c->cts_time[species] = c->cts_vmr[species] = nullptr; }
  CK(cudaMalloc(&c->cts_time[species], sizeof(double) * (size_t)n));
  CK(cudaMalloc(&c->cts_vmr[species], sizeof(double) * (size_t)n));
  CK(cudaMemcpyAsync(c->cts_time[species], time, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->cts_vmr[species], vmr, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->cts_n[species] = n;
  API_END
}

int mpb_set_balloon(mpb_ctx *c, int n, const double *ts, const double *ps) {
  API_BEGIN
  use(c);
  REQUIRE(n >= 1 && ts && ps, "bad balloon series");
  CK(cudaStreamSynchronize(c->stream));
  if (c->iso_ts) { CK(cudaFree(c->iso_ts)); CK(cudaFree(c->iso_ps)); c->iso_ts = c->iso_ps = nullptr; }
  CK(cudaMalloc(&c->iso_ts, sizeof(double) * (size_t)n));
  CK(cudaMalloc(&c->iso_ps, sizeof(double) * (size_t)n));
  CK(cudaMemcpyAsync(c->iso_ts, ts, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->iso_ps, ps, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->iso_n = n;
  API_END
}

int mpb_get_atm(mpb_ctx *c, double *time, double *p, double *lon, double *lat, double *q, int64_t q_stride) {
  API_BEGIN
  use(c);
  unscramble(c);
  const size_t bytes = sizeof(double) * (size_t)c->np;
  if (c->np > 0) {
    if (time) CK(cudaMemcpyAsync(time, c->time(), bytes, cudaMemcpyDeviceToHost, c->stream));
    if (p) CK(cudaMemcpyAsync(p, c->p(), bytes, cudaMemcpyDeviceToHost, c->stream));
    if (lon) CK(cudaMemcpyAsync(lon, c->lon(), bytes, cudaMemcpyDeviceToHost, c->stream));
    if (lat) CK(cudaMemcpyAsync(lat, c->lat(), bytes, cudaMemcpyDeviceToHost, c->stream));
    if (q && c->nq > 0) {
      REQUIRE(!c->q_stale, "the device copy of the quantities is not current after mpb_run_timestep_host (it moves only rp / rhop): "
                           "the caller's arrays hold them; mpb_set_atm makes the device copy current again");
      for (int iq = 0; iq < c->nq; iq++)
        CK(cudaMemcpyAsync(q + (size_t)iq * q_stride, c->q(iq), bytes, cudaMemcpyDeviceToHost, c->stream));
    }
  }
  CK(cudaStreamSynchronize(c->stream));
  API_END
}

int mpb_get_uvwp(mpb_ctx *c, float *uvwp) {
  API_BEGIN
  use(c);
  unscramble(c);
  if (c->np > 0) CK(cudaMemcpyAsync(uvwp, c->uvwp, sizeof(float) * 3 * (size_t)c->np, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END
}

int mpb_get_dt(mpb_ctx *c, double *dt) {
  API_BEGIN
  use(c);
  unscramble(c);
  if (c->np > 0) CK(cudaMemcpyAsync(dt, c->dt, sizeof(double) * (size_t)c->np, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  API_END
}

int64_t mpb_get_np(mpb_ctx *c) { return c ? c->np : -1; }

int mpb_set_shard(mpb_ctx *c, int64_t off, int64_t global_np) {
  API_BEGIN
  use(c);
  REQUIRE(off >= 0 && global_np >= 0, "bad shard");
  c->ig0 = off; c->global_np = global_np;
  API_END
}

int mpb_set_rng_ctr(mpb_ctx *c, uint64_t ctr) {
  API_BEGIN
  use(c);
  c->rng_ctr = ctr;
  API_END
}
uint64_t mpb_get_rng_ctr(mpb_ctx *c) { return c ? c->rng_ctr : 0; }

// The launches of one mpb_run_modules call, in order.  Planning is pure host logic on the control structure (no device
// state), so that the dispatch can be checked without a GPU (mpb_plan_modules, tests/test_dispatch_plan.py).
struct Op {
  enum Kind { STEP, SORT, ISOSURF_INIT, ADVECT_INIT, ADVECT_LEVELS, DIFF_PBL, CONVECTION, ISOSURF, METEO, BOUND_COND, DECAY, MIXING,
              CHEM_GRID } kind;
  int advect;
  unsigned phys, modules;
};

static std::vector<Op> plan_modules(const mpb_ctl_t &k, double t, unsigned mask) {
  std::vector<Op> ops;
  auto op = [&](Op::Kind kind) { ops.push_back(Op{kind, 0, 0u, 0u}); };
  unsigned phys = 0;
  if ((mask & MPB_MOD_DIFF_TURB) && turb_enabled(k)) phys |= PHYS_TURB;
  if ((mask & MPB_MOD_DIFF_MESO) && meso_enabled(k)) phys |= PHYS_MESO;
  if ((mask & MPB_MOD_SEDI) && sedi_enabled(k)) phys |= PHYS_SEDI;
  const int advect = (mask & MPB_MOD_ADVECT) ? k.advect : 0;
  unsigned modules = 0;
  if (mask & MPB_MOD_POSITION0) modules |= MOD_POS_PRE;
  if (mask & MPB_MOD_POSITION1) modules |= MOD_POS_POST;
  const bool on_levels = advect > 0 && k.advect_vert_coord != 0;   // advection runs as its own launch between two segments
  const bool pbl_now = (mask & MPB_MOD_DIFF_PBL) && k.diffusion && k.turb_pbl_scheme == 1;   // src/mptrac.c:7897-7899
  const bool conv_now = (mask & MPB_MOD_CONVECTION) && convection_enabled(k) && (k.conv_dt <= 0 || hits(t, k.conv_dt));
  const bool iso_now = (mask & MPB_MOD_ISOSURF) && isosurf_enabled(k);
  const bool bound0 = (mask & MPB_MOD_BOUND0) && bound_enabled(k), bound1 = (mask & MPB_MOD_BOUND1) && bound_enabled(k);
  const bool decay_now = (mask & MPB_MOD_DECAY) && (decay_enabled(k) || k.qnt_loss_rate >= 0);
  // timesteps ... position1 in one launch: dt stays in registers (modules that run as their own launch read it from memory)
  const bool whole = (mask & 0xff) == 0xff && !on_levels && !pbl_now && !conv_now && !decay_now && !iso_now && !bound0 && !bound1;
  if ((mask & MPB_MOD_TIMESTEPS) && t == k.t_start) {   // src/mptrac.c:7863-7873
    // (whichever segment carries MPB_MOD_ISOSURF later; the balloon series of ISOSURF 4 arrives through mpb_set_balloon)
    if (isosurf_enabled(k) && k.isosurf != 4) op(Op::ISOSURF_INIT);
    if (k.advect_vert_coord == 1) op(Op::ADVECT_INIT);
  }
  const bool sort_now = (mask & MPB_MOD_SORT) && k.sort_dt > 0 && hits(t, k.sort_dt);
  if (mask & MPB_MOD_TIMESTEPS) {
    if (sort_now) {
      // the reference computes dt BEFORE it permutes the parcels and leaves cache->dt in slot order
      // (src/mptrac.c:7877-7881): dt is computed and stored first, then the parcels are sorted, then the fused launch
      // reads dt from memory
      // -- folded into the sort's key pass
      ops.push_back(Op{Op::SORT, 0, 0u, MOD_TIMESTEPS | MOD_STORE_DT});
    } else {
      modules |= MOD_TIMESTEPS;
      if (!whole) modules |= MOD_STORE_DT;   // later segments read cache->dt from memory
    }
  } else if (sort_now) {
    op(Op::SORT);
  }
  // The per-parcel modules in the reference's order (src/mptrac.c:7876-7919).  Everything the fused kernel covers
  // accumulates in one segment; a module that runs as its own launch -- model-level advection, diff_pbl, convection,
  // isosurf -- flushes the segment before it, so a plain configuration is ONE launch and each such module adds two.
  Op seg{Op::STEP, 0, 0u, modules & (MOD_TIMESTEPS | MOD_STORE_DT | MOD_POS_PRE)};
  auto flush = [&]() {
    if (seg.advect || seg.phys || seg.modules) ops.push_back(seg);
    seg = Op{Op::STEP, 0, 0u, 0u};
  };
  if (on_levels) { flush(); op(Op::ADVECT_LEVELS); }
  else seg.advect = advect;
  seg.phys |= phys & PHYS_TURB;
  if (pbl_now) { flush(); op(Op::DIFF_PBL); }
  seg.phys |= phys & PHYS_MESO;
  if (conv_now) { flush(); op(Op::CONVECTION); }
  seg.phys |= phys & PHYS_SEDI;
  if (iso_now) { flush(); op(Op::ISOSURF); }
  seg.modules |= modules & MOD_POS_POST;
  flush();
  // The model-level advection takes its neighbours in: a segment right before it that only carries timesteps / the first
  // position check, and one right after it that only carries the second position check.
  static const bool fold = std::getenv("MPTRAC_B200_NO_LEVEL_FOLD") == nullptr;   // (measurement aid: three launches as before)
  for (size_t i = 0; fold && i < ops.size(); i++) {
    if (ops[i].kind != Op::ADVECT_LEVELS) continue;
    if (i + 1 < ops.size() && ops[i + 1].kind == Op::STEP && !ops[i + 1].advect && !ops[i + 1].phys && ops[i + 1].modules == MOD_POS_POST) {
      ops[i].modules |= MOD_POS_POST;
      ops.erase(ops.begin() + (long)i + 1);
    }
    if (i > 0 && ops[i - 1].kind == Op::STEP && !ops[i - 1].advect && !ops[i - 1].phys &&
        !(ops[i - 1].modules & ~(MOD_TIMESTEPS | MOD_STORE_DT | MOD_POS_PRE))) {
      ops[i].modules |= ops[i - 1].modules;
      ops.erase(ops.begin() + (long)i - 1);
    }
    break;
  }
  if ((mask & MPB_MOD_METEO) && meteo_wanted(k) && k.met_dt_out > 0 && (k.met_dt_out < k.dt_mod || hits(t, k.met_dt_out)))   // :7921-7924
    op(Op::METEO);
  if (bound0) op(Op::BOUND_COND);     // :7926-7929
  if (decay_now) op(Op::DECAY);       // :7931-7940
  if ((mask & MPB_MOD_MIXING) && k.mixing_trop >= 0 && k.mixing_strat >= 0 && (k.mixing_dt <= 0 || hits(t, k.mixing_dt)))   // :7943-7945
    op(Op::MIXING);
  if ((mask & MPB_MOD_CHEMGRID) && k.chemgrid) op(Op::CHEM_GRID);   // :7947-7950
  if (bound1) op(Op::BOUND_COND);     // :7997-8000
  return ops;
}

static void run_op(mpb_ctx *c, double t, const Op &o) {
    switch (o.kind) {
      case Op::STEP: launch_step(c, t, o.advect, o.phys, o.modules); break;
      case Op::SORT: do_sort(c, t, o.modules, own_order_wanted(c)); break;
      case Op::ISOSURF_INIT: launch_isosurf(c, true); break;
      case Op::ADVECT_INIT: launch_advect_init(c); break;
      case Op::ADVECT_LEVELS: launch_advect_levels(c, t, o.modules); break;
      case Op::DIFF_PBL: launch_diff_pbl(c); break;
      case Op::CONVECTION: launch_convection(c); break;
      case Op::ISOSURF: launch_isosurf(c, false); break;
      case Op::METEO: launch_meteo(c); break;
      case Op::BOUND_COND: launch_bound_cond(c); break;
      case Op::DECAY: launch_decay(c); break;
      case Op::CHEM_GRID: launch_chem_grid(c, t); break;
      case Op::MIXING: mixing_inline(c, t); break;
    }
}

static void run_modules(mpb_ctx *c, double t, unsigned mask) {
  REQUIRE(c->have_ctl, "mpb_set_ctl has not been called");
  REQUIRE(c->team == nullptr, "this context belongs to a team: step it through mpb_team_run_modules");
  for (const Op &o : plan_modules(c->ctl, t, mask)) run_op(c, t, o);
}

// The plan of mpb_run_modules(ctx, t, mask) for a control structure, as text: one launch per token, e.g.
// "step(advect=4,phys=0x0,mod=0x43) meteo".  Needs neither a context nor a device.
int mpb_plan_modules(const mpb_ctl_t *ctl, double t, unsigned mask, char *buf, int len) {
  API_BEGIN
  REQUIRE(ctl && buf && len > 0, "bad arguments");
  static const char *names[] = {"step", "sort", "isosurf_init", "advect_init", "advect_levels", "diff_pbl", "convection", "isosurf",
                                "meteo", "bound_cond", "decay", "mixing", "chem_grid"};
  std::string out;
  for (const Op &o : plan_modules(*ctl, t, mask)) {
    if (!out.empty()) out += ' ';
    out += names[o.kind];
    if (o.kind == Op::STEP) {
      char tmp[64];
      std::snprintf(tmp, sizeof(tmp), "(advect=%d,phys=0x%x,mod=0x%x)", o.advect, o.phys, o.modules);
      out += tmp;
    } else if ((o.kind == Op::ADVECT_LEVELS || o.kind == Op::SORT) && o.modules) {
      char tmp[32];
      std::snprintf(tmp, sizeof(tmp), "(mod=0x%x)", o.modules);
      out += tmp;
    }
  }
  REQUIRE((int)out.size() < len, "plan buffer too small");
  std::memcpy(buf, out.c_str(), out.size() + 1);
  API_END
}

int mpb_run_timestep(mpb_ctx *c, double t) {
  API_BEGIN
  use(c);
  run_modules(c, t, MPB_MOD_ALL);
  API_END
}

// One model step for parcels that live in HOST memory: chunk k+1 uploads while chunk k computes and chunk k-1 downloads.
// Each of kLanes streams carries whole chunks (H2D -> step kernel -> D2H in stream order); the copy engines of the two
// directions and the SMs overlap across lanes.  The result equals mpb_set_atm + mpb_run_timestep + mpb_get_atm.
int mpb_run_timestep_host(mpb_ctx *c, double t, int64_t np, double *time, double *p, double *lon, double *lat,
                          double *q, int64_t q_stride) {
  API_BEGIN
  use(c);
  REQUIRE(c->have_ctl, "mpb_set_ctl has not been called");
  REQUIRE(np >= 0 && np <= c->np_max, "np exceeds the context capacity");
  REQUIRE(np == 0 || (time && p && lon && lat), "null parcel array");
  const mpb_ctl_t &k = c->ctl;
  const bool sort_now = k.sort_dt > 0 && hits(t, k.sort_dt);
  const bool mix_now = k.mixing_trop >= 0 && k.mixing_strat >= 0 && (k.mixing_dt <= 0 || hits(t, k.mixing_dt)) && k.n_mix_qnt > 0;
  const bool meteo_now = meteo_wanted(k) && k.met_dt_out > 0 && (k.met_dt_out < k.dt_mod || hits(t, k.met_dt_out));
  if (sort_now || mix_now || meteo_now || k.advect_vert_coord != 0 || convection_enabled(k) || decay_enabled(k) || k.qnt_loss_rate >= 0 || isosurf_enabled(k) ||
      (k.diffusion && k.turb_pbl_scheme == 1) || bound_enabled(k) || k.chemgrid ||
      np < 4 * kHostChunkMin) {
    // steps with a global phase (cell sort, box means), steps that write quantities (meteo) and tiny problems take the
    // plain sequence
    REQUIRE(mpb_set_atm(c, np, time, p, lon, lat, q, q_stride) == 0, g_err);
    c->leaving_soon = true;      // (the parcels go back to the host right away: no point in an order of the engine's own)
    try { run_modules(c, t, MPB_MOD_ALL); } catch (...) { c->leaving_soon = false; throw; }
    c->leaving_soon = false;
    REQUIRE(mpb_get_atm(c, time, p, lon, lat, q, q_stride) == 0, g_err);
    c->host_h2d = c->host_d2h = (32 + 8 * (long long)c->nq) * np;
    return 0;
  }
  unscramble(c);
  c->np = np;
  c->q_stale = c->nq > 0;      // only rp / rhop cross the link below; everything that reads q[] on the device wants mpb_set_atm first
  unsigned phys = 0;
  if (turb_enabled(k)) phys |= PHYS_TURB;
  if (meso_enabled(k)) phys |= PHYS_MESO;
  if (sedi_enabled(k)) phys |= PHYS_SEDI;
  REQUIRE(!(phys & PHYS_SEDI) || q != nullptr, "null quantity array");
  StepArgs A = step_args(c, t, k.advect, phys, MOD_TIMESTEPS | MOD_POS_PRE | MOD_POS_POST);

  // MPTRAC_B200_HOST_MODE picks how the parcels cross the host link (read per call):
  //   zerocopy (default)  kernel reads and writes the mapped host arrays, no copy engine
  //   dma_in              copy engine uploads chunk by chunk, the kernel writes its results straight to the host arrays
  //   dma_out             the kernel reads the host arrays, the copy engine downloads chunk by chunk
  //   copy                copy engine both ways (also what unmapped = pageable host arrays get)
  // MPTRAC_B200_HOST_ZEROCOPY=0 is the older spelling of "copy".
  enum { HM_ZEROCOPY, HM_DMA_IN, HM_DMA_OUT, HM_COPY } mode = HM_ZEROCOPY;
  if (const char *e = std::getenv("MPTRAC_B200_HOST_MODE")) {
    if (!std::strcmp(e, "dma_in")) mode = HM_DMA_IN;
    else if (!std::strcmp(e, "dma_out")) mode = HM_DMA_OUT;
    else if (!std::strcmp(e, "copy")) mode = HM_COPY;
    else REQUIRE(!std::strcmp(e, "zerocopy"), "MPTRAC_B200_HOST_MODE must be zerocopy, dma_in, dma_out or copy");
  }
  if (std::getenv("MPTRAC_B200_HOST_ZEROCOPY") && std::atoi(std::getenv("MPTRAC_B200_HOST_ZEROCOPY")) == 0) mode = HM_COPY;
  void *d[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (mode != HM_COPY) {
    double *h[6] = {time, p, lon, lat, nullptr, nullptr};
    int nh = 4;
    if (phys & PHYS_SEDI) { h[4] = q + (size_t)k.qnt_rp * q_stride; h[5] = q + (size_t)k.qnt_rhop * q_stride; nh = 6; }
    bool mapped = true;
    for (int i = 0; i < nh && mapped; i++) {
      cudaPointerAttributes at;
      mapped = cudaPointerGetAttributes(&at, h[i]) == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer != nullptr;
      d[i] = mapped ? at.devicePointer : nullptr;
    }
    cudaGetLastError();
    if (!mapped) mode = HM_COPY;
  }
  // Pinned host arrays are mapped into the device address space: the kernel then reads the parcels straight from host
  // memory and writes them straight back -- both PCIe directions run concurrently for the whole launch, with no copy
  // engine, no chunking and no staging (the persistent kernel keeps the next parcel's loads in flight a whole parcel
  // ahead, which is what hides the link latency).
  if (mode == HM_ZEROCOPY) {
    A.in_time = (const double *)d[0]; A.in_p = (const double *)d[1]; A.in_lon = (const double *)d[2]; A.in_lat = (const double *)d[3];
    A.host_time = (double *)d[0]; A.host_p = (double *)d[1]; A.host_lon = (double *)d[2]; A.host_lat = (double *)d[3];
    if (phys & PHYS_SEDI) { A.rp = (const double *)d[4]; A.rhop = (const double *)d[5]; }
    // Do all parcels carry the same time (they do unless they were released at different times)?  Then module_timesteps
    // takes the same decision for every parcel of a global met domain (src/mptrac.c:6019-6024 depends on the time only) and
    // all of them end at the same time: neither direction of time[] needs the link -- the kernel starts every parcel at
    // that value and the host writes the common result into its own array while the kernel runs.  A quarter of the traffic.
    // (MPTRAC_B200_HOST_TIME=explicit always moves the array.)
    bool uniform = np > 0 && !A.met.local && !(std::getenv("MPTRAC_B200_HOST_TIME") && !std::strcmp(std::getenv("MPTRAC_B200_HOST_TIME"), "explicit"));
    if (uniform) {
      long long differ = 0;
      unsigned long long first_bits;
      std::memcpy(&first_bits, time, sizeof(first_bits));
      const unsigned long long *bits = reinterpret_cast<const unsigned long long *>(time);
#pragma omp parallel for reduction(+ : differ) schedule(static)
      for (long long i = 0; i < np; i++) differ += bits[i] != first_bits;
      uniform = differ == 0;
    }
    double t_out = 0;
    bool moves = false;
    if (uniform) {
      A.uniform_time = 1; A.time_in = time[0]; A.host_time = nullptr;
      Parcel a0;
      a0.time = time[0]; a0.lon = a0.lat = a0.p = 0;
      const double dt0 = parcel_dt(A.met, A.ctl, a0);
      moves = dt0 != 0 && k.advect > 0;          // (module_advect is what advances a parcel's time, src/mptrac.c:3671)
      t_out = a0.time + dt0;
    }
    c->host_h2d = (uniform ? 24 : 32) * np + ((phys & PHYS_SEDI) ? 16 * np : 0);
    c->host_d2h = ((uniform || k.advect == 0) ? 24 : 32) * np;
    launch_range(c, A, k.advect, phys, 0, np, c->stream);
    if (moves) {
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < np; i++) time[i] = t_out;
    }
    CK(cudaStreamSynchronize(c->stream));   // the host arrays are valid on return
    return 0;
  }
  c->host_h2d = 32 * np + ((phys & PHYS_SEDI) ? 16 * np : 0);
  c->host_d2h = 32 * np;
  const bool dma_in = mode != HM_DMA_OUT, dma_out = mode != HM_DMA_IN;
  if (!dma_in) {
    A.in_time = (const double *)d[0]; A.in_p = (const double *)d[1]; A.in_lon = (const double *)d[2]; A.in_lat = (const double *)d[3];
    if (phys & PHYS_SEDI) { A.rp = (const double *)d[4]; A.rhop = (const double *)d[5]; }
  }
  if (!dma_out) { A.host_time = (double *)d[0]; A.host_p = (double *)d[1]; A.host_lon = (double *)d[2]; A.host_lat = (double *)d[3]; }
  if (!c->lane[0]) {
    for (int i = 0; i < kLanes; i++) {
      CK(cudaStreamCreateWithFlags(&c->lane[i], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&c->lane_done[i], cudaEventDisableTiming));
    }
    CK(cudaEventCreateWithFlags(&c->lane_go, cudaEventDisableTiming));
  }
  // everything already queued on the context's stream (met uploads, earlier steps) comes first
  CK(cudaEventRecord(c->lane_go, c->stream));
  for (int i = 0; i < kLanes; i++) CK(cudaStreamWaitEvent(c->lane[i], c->lane_go, 0));
  long long chunk = (np + 2 * kLanes - 1) / (2 * kLanes);
  chunk = std::max<long long>(kHostChunkMin, std::min<long long>(chunk, kHostChunkMax));
  if (const char *e = std::getenv("MPTRAC_B200_HOST_CHUNK")) chunk = std::max<long long>(kHostChunkMin, std::atoll(e));   // tuning aid
  chunk = (chunk + kBlock - 1) / kBlock * kBlock;
  // MPTRAC_B200_TRACE=1: print the device timeline of this call's chunks (diagnostics; adds events, nothing else)
  static const bool trace = std::getenv("MPTRAC_B200_TRACE") != nullptr;
  std::vector<cudaEvent_t> tev;
  auto mark = [&](cudaStream_t st) {
    if (!trace) return;
    cudaEvent_t e;
    CK(cudaEventCreate(&e));
    CK(cudaEventRecord(e, st));
    tev.push_back(e);
  };
  mark(c->stream);
  static const bool allow_2d = !(std::getenv("MPTRAC_B200_HOST_2D") && std::atoi(std::getenv("MPTRAC_B200_HOST_2D")) == 0);
  const bool strided = allow_2d && np > 0 && (p - time) >= np && (lon - p) == (p - time) && (lat - lon) == (p - time);
  const size_t spitch = strided ? sizeof(double) * (size_t)(p - time) : 0, dpitch = sizeof(double) * (size_t)c->np_max;
  int lane = 0;
  for (long long off = 0; off < np; off += chunk, lane = (lane + 1) % kLanes) {
    const long long cnt = std::min<long long>(chunk, np - off);
    const size_t bytes = sizeof(double) * (size_t)cnt;
    cudaStream_t st = c->lane[lane];
    if (!dma_in) {
      // the kernel reads this chunk from the mapped host arrays
    } else if (strided) {
      // time, p, lon, lat sit at one constant stride in the caller's memory (they are consecutive members of atm_t):
      // one 2-D copy per direction and chunk instead of four
      CK(cudaMemcpy2DAsync(c->time() + off, dpitch, time + off, spitch, bytes, 4, cudaMemcpyHostToDevice, st));
    } else {
      CK(cudaMemcpyAsync(c->time() + off, time + off, bytes, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(c->p() + off, p + off, bytes, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(c->lon() + off, lon + off, bytes, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(c->lat() + off, lat + off, bytes, cudaMemcpyHostToDevice, st));
    }
    if (dma_in && (phys & PHYS_SEDI)) {   // the only quantities the path reads; none is modified
      CK(cudaMemcpyAsync(c->q(k.qnt_rp) + off, q + (size_t)k.qnt_rp * q_stride + off, bytes, cudaMemcpyHostToDevice, st));
      CK(cudaMemcpyAsync(c->q(k.qnt_rhop) + off, q + (size_t)k.qnt_rhop * q_stride + off, bytes, cudaMemcpyHostToDevice, st));
    }
    mark(st);
    launch_range(c, A, k.advect, phys, off, cnt, st);
    mark(st);
    if (!dma_out) {
      // the kernel has written this chunk to the mapped host arrays
    } else if (strided) {
      CK(cudaMemcpy2DAsync(time + off, spitch, c->time() + off, dpitch, bytes, 4, cudaMemcpyDeviceToHost, st));
    } else {
      CK(cudaMemcpyAsync(time + off, c->time() + off, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(p + off, c->p() + off, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(lon + off, c->lon() + off, bytes, cudaMemcpyDeviceToHost, st));
      CK(cudaMemcpyAsync(lat + off, c->lat() + off, bytes, cudaMemcpyDeviceToHost, st));
    }
    mark(st);
  }
  for (int i = 0; i < kLanes; i++) {
    CK(cudaEventRecord(c->lane_done[i], c->lane[i]));
    CK(cudaStreamWaitEvent(c->stream, c->lane_done[i], 0));   // later work on the context's stream sees the step
  }
  for (int i = 0; i < kLanes; i++) CK(cudaStreamSynchronize(c->lane[i]));   // the host arrays are valid on return
  if (trace && !tev.empty()) {
    std::fprintf(stderr, "[mpb trace] host step, %lld parcels, chunk %lld%s:", (long long)np, chunk, strided ? ", 2-D copies" : "");
    for (size_t i = 1; i < tev.size(); i++) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, tev[0], tev[i]));
      std::fprintf(stderr, "%s%.0f", (i % 3 == 1) ? " | " : " ", ms * 1e3);
    }
    std::fprintf(stderr, "  (us since call: after H2D, after kernel, after D2H per chunk)\n");
    for (cudaEvent_t e : tev) cudaEventDestroy(e);
  }
  API_END
}

int mpb_host_step_bytes(mpb_ctx *c, int64_t *h2d, int64_t *d2h) {
  API_BEGIN
  REQUIRE(c && h2d && d2h, "null argument");
  *h2d = c->host_h2d; *d2h = c->host_d2h;
  API_END
}

int mpb_run_modules(mpb_ctx *c, double t, unsigned mask) {
  API_BEGIN
  use(c);
  run_modules(c, t, mask);
  API_END
}

int mpb_module_timesteps(mpb_ctx *c, double t) {
  API_BEGIN
  use(c);
  // dt-only pass; position arrays of active parcels are rewritten with identical values
  launch_step(c, t, 0, 0, MOD_TIMESTEPS | MOD_STORE_DT);
  API_END
}
int mpb_module_position(mpb_ctx *c) {
  API_BEGIN
  use(c);
  launch_step(c, 0.0, 0, 0, MOD_POS_PRE);
  API_END
}
int mpb_module_advect(mpb_ctx *c) {
  API_BEGIN
  use(c);
  REQUIRE(c->ctl.advect > 0, "ADVECT is 0");
  launch_step(c, 0.0, c->ctl.advect, 0, 0);
  API_END
}
int mpb_module_diff_turb(mpb_ctx *c) {
  API_BEGIN
  use(c);
  launch_step(c, 0.0, 0, PHYS_TURB, 0);
  API_END
}
int mpb_module_diff_meso(mpb_ctx *c) {
  API_BEGIN
  use(c);
  launch_step(c, 0.0, 0, PHYS_MESO, 0);
  API_END
}
int mpb_module_sedi(mpb_ctx *c) {
  API_BEGIN
  use(c);
  launch_step(c, 0.0, 0, PHYS_SEDI, 0);
  API_END
}
int mpb_module_meteo(mpb_ctx *c) {
  API_BEGIN
  use(c);
  REQUIRE(c->have_ctl && meteo_wanted(c->ctl), "module_meteo: the control structure names no quantity the device computes (qnt_meteo)");
  launch_meteo(c);
  API_END
}
int mpb_module_sort(mpb_ctx *c) {
  API_BEGIN
  use(c);
  do_sort(c);
  API_END
}

// --- module_mixing split in two, for callers that sum the box records over ranks themselves (e.g. one NCCL all-reduce of
//     mpb_device_ptr("mix_rec"), mpb_mixing_rec_len() doubles, between the two calls) ---
int mpb_mixing_accumulate_all(mpb_ctx *c, double t) {
  API_BEGIN
  use(c);
  REQUIRE(c->have_ctl, "mpb_set_ctl has not been called");
  REQUIRE(c->nranks == 1, "ranks attached through mpb_peer_attach exchange inside mpb_module_mixing / mpb_run_timestep");
  mixing_prepare(c, t);
  mixing_accumulate_all(c);
  API_END
}
int mpb_mixing_apply_all(mpb_ctx *c) {
  API_BEGIN
  use(c);
  REQUIRE(c->nranks == 1 && c->mix_total > 0, "mpb_mixing_accumulate_all has not been called");
  mixing_apply_all(c);
  API_END
}
int64_t mpb_mixing_nbox(mpb_ctx *c) { return c && c->have_ctl ? mixing_total(c) : -1; }
int64_t mpb_mixing_rec_len(mpb_ctx *c) { return c && c->mix_total > 0 ? (long long)((c->nmix + 2) / 2 * 2) * c->mix_total : -1; }

int mpb_module_mixing(mpb_ctx *c, double t) {
  API_BEGIN
  use(c);
  REQUIRE(c->have_ctl, "mpb_set_ctl has not been called");
  REQUIRE(c->team == nullptr, "this context belongs to a team");
  mixing_inline(c, t);
  API_END
}

int mpb_module_rng(mpb_ctx *c, double *rs_host, int64_t n, int method) {
  API_BEGIN
  use(c);
  REQUIRE(n >= 0 && (method == 0 || method == 1), "bad rng request");
  const unsigned long long start = c->rng_ctr;
  c->rng_ctr += (unsigned long long)n + 1ull;  // src/mptrac.c:5812
  if (rs_host) {
    double *d = nullptr;
    CK(cudaMalloc(&d, sizeof(double) * (size_t)(n + 1)));
    rng_fill_kernel<<<nblocks(n + 1, 256), 256, 0, c->stream>>>(start, d, n, method);
    CK(cudaGetLastError());
    c->launches++;
    CK(cudaMemcpyAsync(rs_host, d, sizeof(double) * (size_t)(n + 1), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaFree(d));
  }
  API_END
}

int mpb_grid_accumulate(mpb_ctx *c, const mpb_grid_t *g) {
  API_BEGIN
  use(c);
  REQUIRE(g && g->nx > 0 && g->ny > 0 && g->nz > 0, "bad grid");
  REQUIRE(!c->q_stale, "gridded output: the device copy of the quantities is not current (mpb_run_timestep_host): call mpb_set_atm");
  const long long nbox = (long long)g->nx * g->ny * g->nz;
  REQUIRE(nbox < (1ll << 31), "grid too large");
  ensure_boxes(c);
  const long long need = nbox * std::max(c->nq, 1);
  if (c->nranks > 1) {
    // partial arrays of a multi-rank run live in the exchange area, where rank 0 can read them (mpb_grid_reduce)
    REQUIRE(c->area[c->rank] != nullptr && (size_t)need * 16 + (size_t)nbox * 4 <= c->area_grid_bytes,
            "the exchange area is too small for this output grid (mpb_peer_init: boxes x (16 x quantities + 4) bytes)");
    if (!c->grid_in_area && c->grid_sum) { CK(cudaFree(c->grid_sum)); CK(cudaFree(c->grid_sq)); CK(cudaFree(c->grid_cnt)); c->grid_cap = 0; }
    char *g0 = grid_region(c, c->rank);
    c->grid_sum = (double *)g0; c->grid_sq = c->grid_sum + need; c->grid_cnt = (int *)(c->grid_sq + need);
    c->grid_in_area = true;
  } else if (need > c->grid_cap || c->grid_in_area) {
    if (c->grid_sum && !c->grid_in_area) { CK(cudaFree(c->grid_sum)); CK(cudaFree(c->grid_sq)); CK(cudaFree(c->grid_cnt)); }
    CK(cudaMalloc(&c->grid_sum, sizeof(double) * (size_t)need));
    CK(cudaMalloc(&c->grid_sq, sizeof(double) * (size_t)need));
    CK(cudaMalloc(&c->grid_cnt, sizeof(int) * (size_t)need));
    c->grid_cap = need; c->grid_in_area = false;
  }
  c->grid_nbox = nbox;
  CK(cudaMemsetAsync(c->grid_sum, 0, sizeof(double) * (size_t)need, c->stream));
  CK(cudaMemsetAsync(c->grid_sq, 0, sizeof(double) * (size_t)need, c->stream));
  CK(cudaMemsetAsync(c->grid_cnt, 0, sizeof(int) * (size_t)nbox, c->stream));
  if (c->np == 0) return 0;
  BoxArgs b;
  b.t0 = g->t0; b.t1 = g->t1; b.lon0 = g->lon0; b.lon1 = g->lon1; b.lat0 = g->lat0; b.lat1 = g->lat1;
  b.z0 = g->z0; b.z1 = g->z1; b.nx = g->nx; b.ny = g->ny; b.nz = g->nz;
  box_cells(b);
  grid_bin_kernel<<<nblocks(c->np, 256 * kBinRounds), 256, 0, c->stream>>>(
      b, c->time(), c->lon(), c->lat(), c->p(), c->nq ? c->q(0) : nullptr, c->np_max, c->nq, nbox, c->grid_cnt, c->grid_sum, c->grid_sq, c->np);
  CK(cudaGetLastError());
  c->launches++;
  API_END
}

// rank 0 adds the partial arrays of all ranks to its own (peer reads), between two barriers; the other ranks only pass the barriers
static void grid_pull(mpb_ctx *c) {
  const long long nbox = c->grid_nbox, nval = nbox * std::max(c->nq, 1);
  GridPeers G;
  for (int r = 0; r < kMaxRanks; r++) {
    G.sum[r] = G.sq[r] = nullptr; G.cnt[r] = nullptr;
    if (r < c->nranks) {
      REQUIRE(c->area[r] != nullptr, "mpb_peer_attach has not been called");
      G.sum[r] = (const double *)grid_region(c, r); G.sq[r] = G.sum[r] + nval; G.cnt[r] = (const int *)(G.sq[r] + nval);
    }
  }
  grid_pull_kernel<<<nblocks(std::max(nval, nbox), 256), 256, 0, c->stream>>>(G, c->nranks, nbox, nval, c->grid_sum, c->grid_sq, c->grid_cnt);
  CK(cudaGetLastError());
  c->launches++;
}

int mpb_grid_reduce(mpb_ctx *c) {
  API_BEGIN
  use(c);
  REQUIRE(c->grid_nbox > 0, "mpb_grid_accumulate has not been called");
  REQUIRE(c->team == nullptr, "this context belongs to a team");
  if (c->nranks > 1) {
    peer_barrier(c);
    if (c->rank == 0) grid_pull(c);
    peer_barrier(c);
  }
  API_END
}

int mpb_grid_fetch(mpb_ctx *c, int *count, double *sum, double *sq) {
  API_BEGIN
  use(c);
  REQUIRE(c->grid_nbox > 0, "mpb_grid_accumulate has not been called");
  const size_t nb = (size_t)c->grid_nbox, nv = nb * (size_t)c->nq;
  const size_t need = 2 * sizeof(double) * nv + sizeof(int) * nb;
  if (need > c->grid_stage_bytes) {
    if (c->grid_stage_h) CK(cudaFreeHost(c->grid_stage_h));
    c->grid_stage_h = nullptr; c->grid_stage_bytes = 0;
    CK(cudaMallocHost(&c->grid_stage_h, need));
    c->grid_stage_bytes = need;
  }
  double *h_sum = reinterpret_cast<double *>(c->grid_stage_h), *h_sq = h_sum + nv;
  int *h_cnt = reinterpret_cast<int *>(h_sq + nv);
  if (count) CK(cudaMemcpyAsync(h_cnt, c->grid_cnt, sizeof(int) * nb, cudaMemcpyDeviceToHost, c->stream));
  if (sum && nv) CK(cudaMemcpyAsync(h_sum, c->grid_sum, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
  if (sq && nv) CK(cudaMemcpyAsync(h_sq, c->grid_sq, sizeof(double) * nv, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (count) std::memcpy(count, h_cnt, sizeof(int) * nb);
  if (sum && nv) std::memcpy(sum, h_sum, sizeof(double) * nv);
  if (sq && nv) std::memcpy(sq, h_sq, sizeof(double) * nv);
  API_END
}

void *mpb_device_ptr(mpb_ctx *c, const char *name) {
  if (!c || !name) return nullptr;
  const std::string n(name);
  if (c->scrambled && (n == "time" || n == "p" || n == "lon" || n == "lat" || n == "q" || n == "dt" || n == "uvwp")) {
    try { use(c); unscramble(c); } catch (const std::exception &ex) { g_err = ex.what(); return nullptr; }
  }
  if (n == "time") return c->time();
  if (n == "p") return c->p();
  if (n == "lon") return c->lon();
  if (n == "lat") return c->lat();
  if (n == "q") return c->nq ? c->q(0) : nullptr;
  if (n == "dt") return c->dt;
  if (n == "uvwp") return c->uvwp;
  if (n == "mix_rec") return c->mix_rec;
  if (n == "grid_sum") return c->grid_sum;
  if (n == "grid_sq") return c->grid_sq;
  if (n == "grid_cnt") return c->grid_cnt;
  return nullptr;
}

int64_t mpb_launch_count(mpb_ctx *c) { return c ? c->launches : -1; }

int mpb_met_bytes(mpb_ctx *c, int64_t *bytes) {
  API_BEGIN
  REQUIRE(c && bytes, "null argument");
  *bytes = (long long)c->nx * c->ny * c->nz * (long long)sizeof(Node) + (long long)c->nx * c->ny * (long long)sizeof(float4);
  API_END
}

// ------------------------------------------------------------------------------------------------
// ranks that exchange through peer memory
// ------------------------------------------------------------------------------------------------
int mpb_peer_init(mpb_ctx *c, int rank, int nranks, int64_t mix_bytes, int64_t grid_bytes, void *ipc_handle_out) {
  API_BEGIN
  use(c);
  REQUIRE(nranks >= 1 && nranks <= kMaxRanks && rank >= 0 && rank < nranks, "bad rank / number of ranks");
  REQUIRE(mix_bytes >= 0 && grid_bytes >= 0, "bad exchange area size");
  CK(cudaStreamSynchronize(c->stream));
  for (int r = 0; r < kMaxRanks; r++) {
    if (c->area[r]) { if (r == c->rank) CK(cudaFree(c->area[r])); else if (c->area_ipc[r]) cudaIpcCloseMemHandle(c->area[r]); }
    c->area[r] = nullptr; c->area_ipc[r] = false;
  }
  if (c->grid_in_area) { c->grid_sum = c->grid_sq = nullptr; c->grid_cnt = nullptr; c->grid_in_area = false; c->grid_cap = 0; c->grid_nbox = 0; }
  c->rank = rank; c->nranks = nranks;
  c->epoch = 0; c->peer_err = nullptr;
  c->area_bytes = c->area_mix_bytes = c->area_grid_bytes = 0;
  if (nranks == 1) return 0;
  c->area_mix_bytes = ((size_t)mix_bytes + 255) / 256 * 256;
  c->area_grid_bytes = ((size_t)grid_bytes + 255) / 256 * 256;
  c->area_bytes = kAreaHeader + c->area_mix_bytes + c->area_grid_bytes;
  CK(cudaMalloc(&c->area[rank], c->area_bytes));
  CK(cudaMemset(c->area[rank], 0, c->area_bytes));
  c->peer_err = (int *)(c->area[rank] + kAreaErrOffset);
  if (ipc_handle_out) {
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->area[rank]));
    static_assert(sizeof(h) == MPB_IPC_HANDLE_BYTES, "IPC handle size");
    std::memcpy(ipc_handle_out, &h, sizeof(h));
  }
  API_END
}

int mpb_peer_attach(mpb_ctx *c, const void *ipc_handles) {
  API_BEGIN
  use(c);
  REQUIRE(c->nranks > 1 && c->area[c->rank] != nullptr && ipc_handles != nullptr, "mpb_peer_init comes first");
  for (int r = 0; r < c->nranks; r++) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const char *)ipc_handles + (size_t)r * MPB_IPC_HANDLE_BYTES, sizeof(h));
    void *ptr = nullptr;
    CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    c->area[r] = (char *)ptr; c->area_ipc[r] = true;
  }
  API_END
}

int mpb_peer_attach_local(mpb_ctx *c, void *const *areas, const int *devices) {
  API_BEGIN
  use(c);
  REQUIRE(c->nranks > 1 && c->area[c->rank] != nullptr && areas != nullptr && devices != nullptr, "mpb_peer_init comes first");
  for (int r = 0; r < c->nranks; r++) {
    if (r == c->rank) continue;
    REQUIRE(areas[r] != nullptr, "null peer area");
    if (devices[r] != c->device) {
      int can = 0;
      CK(cudaDeviceCanAccessPeer(&can, c->device, devices[r]));
      REQUIRE(can, "the devices of the team cannot access each other's memory");
      cudaError_t e = cudaDeviceEnablePeerAccess(devices[r], 0);
      if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError(); else CK(e);
    }
    c->area[r] = (char *)areas[r]; c->area_ipc[r] = false;
  }
  API_END
}

void *mpb_peer_area(mpb_ctx *c) { return c && c->nranks > 1 ? c->area[c->rank] : nullptr; }

int mpb_peer_barrier(mpb_ctx *c) {
  API_BEGIN
  use(c);
  peer_barrier(c);
  API_END
}

// ------------------------------------------------------------------------------------------------
// A team: several devices behind ONE host thread (what the reference's single-process driver needs).  Parcels are cut
// into contiguous index ranges, one per device; every device holds both met levels (packed once, then copied from device
// to device); steps are launched on all devices before anything is waited for; the phases of the box exchange are
// ordered with stream events.
// ------------------------------------------------------------------------------------------------
struct mpb_team {
  std::vector<mpb_ctx *> ctx;
  std::vector<cudaEvent_t> ev;
  std::vector<long long> off;     // parcel ranges: context i owns [off[i], off[i+1])
  long long np = 0, np_max = 0;
  int nq = 0;
  size_t mix_bytes = 0, grid_bytes = 0;
};

static void team_barrier(mpb_team *T) {
  const size_t n = T->ctx.size();
  if (n < 2) return;
  for (size_t i = 0; i < n; i++) { use(T->ctx[i]); CK(cudaEventRecord(T->ev[i], T->ctx[i]->stream)); }
  for (size_t i = 0; i < n; i++) {
    use(T->ctx[i]);
    for (size_t j = 0; j < n; j++) if (j != i) CK(cudaStreamWaitEvent(T->ctx[i]->stream, T->ev[j], 0));
  }
}

// exchange areas of all members large enough for `mix_bytes` of box records and `grid_bytes` of gridded output
static void team_ensure_area(mpb_team *T, size_t mix_bytes, size_t grid_bytes) {
  const int n = (int)T->ctx.size();
  if (n < 2) return;
  if (T->ctx[0]->area[0] && mix_bytes <= T->mix_bytes && grid_bytes <= T->grid_bytes) return;
  mix_bytes = std::max(mix_bytes, T->mix_bytes); grid_bytes = std::max(grid_bytes, T->grid_bytes);
  for (mpb_ctx *c : T->ctx) { use(c); CK(cudaStreamSynchronize(c->stream)); }
  std::vector<void *> areas(n);
  std::vector<int> devs(n);
  for (int i = 0; i < n; i++) {
    REQUIRE(mpb_peer_init(T->ctx[i], i, n, (int64_t)mix_bytes, (int64_t)grid_bytes, nullptr) == 0, g_err);
    areas[i] = T->ctx[i]->area[i]; devs[i] = T->ctx[i]->device;
  }
  for (int i = 0; i < n; i++) REQUIRE(mpb_peer_attach_local(T->ctx[i], areas.data(), devs.data()) == 0, g_err);
  T->mix_bytes = mix_bytes; T->grid_bytes = grid_bytes;
}

// copy the packed met data (both levels) and its bookkeeping from one context to another device
static void met_clone(mpb_ctx *d, mpb_ctx *s) {
  use(d);
  ensure_grid(d, s->nx, s->ny, s->nz, s->coord_type, s->h_lon.data(), s->h_lat.data(), s->h_p.data(), false);
  const size_t nnode = (size_t)s->nx * s->ny * s->nz, ncol = (size_t)s->nx * s->ny;
  auto copy = [&](void *dst, const void *src, size_t bytes) {
    CK(cudaMemcpyPeerAsync(dst, d->device, src, s->device, bytes, d->stream));
  };
  copy(d->nodes, s->nodes, sizeof(Node) * nnode);
  copy(d->surf, s->surf, sizeof(float4) * ncol);
  d->lev[0] = s->lev[0]; d->lev[1] = s->lev[1];
  const size_t nlev = ncol * (size_t)s->npl;
  if (s->lev_p && nlev > 0) {
    if (d->npl != s->npl || nlev > d->lev_cap) {
      for (void *o : {(void *)d->lev_p, (void *)d->lev_z, (void *)d->lev_pz}) if (o) CK(cudaFree(o));
      CK(cudaMalloc(&d->lev_p, sizeof(LevelNode) * nlev));
      CK(cudaMalloc(&d->lev_z, sizeof(LevelNode) * nlev));
      CK(cudaMalloc(&d->lev_pz, sizeof(float4) * nlev));
      d->lev_cap = nlev; d->npl = s->npl;
    }
    copy(d->lev_p, s->lev_p, sizeof(LevelNode) * nlev);
    copy(d->lev_z, s->lev_z, sizeof(LevelNode) * nlev);
    copy(d->lev_pz, s->lev_pz, sizeof(float4) * nlev);
  }
  d->lev_valid[0] = s->lev_valid[0]; d->lev_valid[1] = s->lev_valid[1];
  for (int f = 0; f < MPB_NX2; f++) {
    if (s->x2[f]) {
      if (!d->x2[f]) CK(cudaMalloc(&d->x2[f], sizeof(float2) * ncol));
      copy(d->x2[f], s->x2[f], sizeof(float2) * ncol);
    }
    d->x2_valid[0][f] = s->x2_valid[0][f]; d->x2_valid[1][f] = s->x2_valid[1][f];
  }
  for (int f = 0; f < MPB_NX3; f++) {
    if (s->x3[f]) {
      if (!d->x3[f]) CK(cudaMalloc(&d->x3[f], sizeof(float2) * nnode));
      copy(d->x3[f], s->x3[f], sizeof(float2) * nnode);
    }
    d->x3_valid[0][f] = s->x3_valid[0][f]; d->x3_valid[1][f] = s->x3_valid[1][f];
  }
}

#define TEAM_EACH(call)                                                      \
  do {                                                                       \
    REQUIRE(T != nullptr && !T->ctx.empty(), "null team");                   \
    for (mpb_ctx *c : T->ctx) REQUIRE((call) == 0, g_err);                   \
  } while (0)

int mpb_team_create(mpb_team **out, int ndev, const int *devices, int64_t np_max, int nq) {
  API_BEGIN
  REQUIRE(out != nullptr && devices != nullptr, "null argument");
  *out = nullptr;
  REQUIRE(ndev >= 1 && ndev <= kMaxRanks, "a team has 1 .. MPB_MAX_RANKS members");
  mpb_team *T = new mpb_team();
  T->np_max = np_max; T->nq = nq;
  const int64_t share = np_max / ndev + 1;
  for (int i = 0; i < ndev; i++) {
    mpb_ctx *c = nullptr;
    if (mpb_create(&c, devices[i], share, nq) != 0) {
      const std::string why = g_err;
      for (mpb_ctx *x : T->ctx) { x->team = nullptr; mpb_destroy(x); }
      delete T;
      throw std::runtime_error(why);
    }
    c->team = ndev > 1 ? T : nullptr;
    T->ctx.push_back(c);
    cudaEvent_t e;
    CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    T->ev.push_back(e);
  }
  T->off.assign(ndev + 1, 0);
  *out = T;
  API_END
}

int mpb_team_destroy(mpb_team *T) {
  API_BEGIN
  if (!T) return 0;
  for (mpb_ctx *c : T->ctx) { use(c); CK(cudaStreamSynchronize(c->stream)); }
  for (size_t i = 0; i < T->ctx.size(); i++) {
    use(T->ctx[i]);
    cudaEventDestroy(T->ev[i]);
    T->ctx[i]->team = nullptr;
    mpb_destroy(T->ctx[i]);
  }
  delete T;
  API_END
}

int mpb_team_size(mpb_team *T) { return T ? (int)T->ctx.size() : -1; }
mpb_ctx *mpb_team_member(mpb_team *T, int i) { return T && i >= 0 && i < (int)T->ctx.size() ? T->ctx[i] : nullptr; }

int mpb_team_set_ctl(mpb_team *T, const mpb_ctl_t *ctl) {
  API_BEGIN
  TEAM_EACH(mpb_set_ctl(c, ctl));
  if (T->ctx.size() > 1 && ctl->mixing_trop >= 0 && ctl->mixing_strat >= 0) {
    int nmix = 0;
    for (int i = 0; i < ctl->n_mix_qnt; i++) nmix += ctl->mix_qnt[i] >= 0;
    const long long total = mixing_total(T->ctx[0]), n = (long long)T->ctx.size();
    if (nmix > 0 && total > 0)
      team_ensure_area(T, kRouteHeader + 2 * sizeof(double) * (size_t)n * (size_t)T->ctx[0]->np_max * (size_t)(nmix + 1), T->grid_bytes);
  }
  API_END
}

int mpb_team_set_clim_tropo(mpb_team *T, int ntime, int nlat, const double *time, const double *lat, const double *tropo) {
  API_BEGIN
  TEAM_EACH(mpb_set_clim_tropo(c, ntime, nlat, time, lat, tropo));
  API_END
}
int mpb_team_set_clim_ts(mpb_team *T, int species, int n, const double *time, const double *vmr) {
  API_BEGIN
  TEAM_EACH(mpb_set_clim_ts(c, species, n, time, vmr));
  API_END
}
int mpb_team_set_balloon(mpb_team *T, int n, const double *ts, const double *ps) {
  API_BEGIN
  TEAM_EACH(mpb_set_balloon(c, n, ts, ps));
  API_END
}

// one upload + pack on the first device, then device-to-device copies of the packed arrays (NVLink) instead of one host
// upload per device: the host link is the scarce resource (the reference broadcasts its met data likewise, src/mptrac.c:45-69)
int mpb_team_set_met(mpb_team *T, int slot, const mpb_met_view_t *met) {
  API_BEGIN
  REQUIRE(T != nullptr && !T->ctx.empty(), "null team");
  REQUIRE(mpb_set_met(T->ctx[0], slot, met) == 0, g_err);
  for (size_t i = 1; i < T->ctx.size(); i++) {
    met_clone(T->ctx[i], T->ctx[0]);
    CK(cudaEventRecord(T->ev[i], T->ctx[i]->stream));          // the source arrays must not change (mpb_team_swap_met, the next
    use(T->ctx[0]);                                             // upload) before the copies have read them
    CK(cudaStreamWaitEvent(T->ctx[0]->stream, T->ev[i], 0));
  }
  API_END
}
int mpb_team_swap_met(mpb_team *T) {
  API_BEGIN
  TEAM_EACH(mpb_swap_met(c));
  API_END
}

int mpb_team_set_atm(mpb_team *T, int64_t np, const double *time, const double *p, const double *lon, const double *lat,
                     const double *q, int64_t q_stride) {
  API_BEGIN
  REQUIRE(T != nullptr && !T->ctx.empty(), "null team");
  REQUIRE(np >= 0 && np <= T->np_max, "np exceeds the team's capacity");
  const long long n = (long long)T->ctx.size();
  T->np = np;
  for (long long i = 0; i <= n; i++) T->off[i] = np / n * i + std::min<long long>(i, np % n);   // sizes differ by at most one
  for (long long i = 0; i < n; i++) {
    const long long lo = T->off[i], cnt = T->off[i + 1] - lo;
    REQUIRE(mpb_set_atm(T->ctx[i], cnt, time + lo, p + lo, lon + lo, lat + lo, q ? q + lo : nullptr, q_stride) == 0, g_err);
    REQUIRE(mpb_set_shard(T->ctx[i], lo, np) == 0, g_err);
  }
  API_END
}

int mpb_team_get_atm(mpb_team *T, double *time, double *p, double *lon, double *lat, double *q, int64_t q_stride) {
  API_BEGIN
  REQUIRE(T != nullptr && !T->ctx.empty(), "null team");
  for (size_t i = 0; i < T->ctx.size(); i++) {
    const long long lo = T->off[i];
    REQUIRE(mpb_get_atm(T->ctx[i], time ? time + lo : nullptr, p ? p + lo : nullptr, lon ? lon + lo : nullptr, lat ? lat + lo : nullptr,
                        q ? q + lo : nullptr, q_stride) == 0, g_err);
  }
  API_END
}

int mpb_team_set_uvwp(mpb_team *T, const float *uvwp) {
  API_BEGIN
  REQUIRE(T != nullptr && uvwp != nullptr, "null argument");
  for (size_t i = 0; i < T->ctx.size(); i++) REQUIRE(mpb_set_uvwp(T->ctx[i], uvwp + 3 * T->off[i]) == 0, g_err);
  API_END
}
int mpb_team_get_uvwp(mpb_team *T, float *uvwp) {
  API_BEGIN
  REQUIRE(T != nullptr && uvwp != nullptr, "null argument");
  for (size_t i = 0; i < T->ctx.size(); i++) REQUIRE(mpb_get_uvwp(T->ctx[i], uvwp + 3 * T->off[i]) == 0, g_err);
  API_END
}
int mpb_team_get_dt(mpb_team *T, double *dt) {
  API_BEGIN
  REQUIRE(T != nullptr && dt != nullptr, "null argument");
  for (size_t i = 0; i < T->ctx.size(); i++) REQUIRE(mpb_get_dt(T->ctx[i], dt + T->off[i]) == 0, g_err);
  API_END
}
int mpb_team_set_iso_var(mpb_team *T, const double *v) {
  API_BEGIN
  REQUIRE(T != nullptr && v != nullptr, "null argument");
  for (size_t i = 0; i < T->ctx.size(); i++) REQUIRE(mpb_set_iso_var(T->ctx[i], v + T->off[i]) == 0, g_err);
  API_END
}
int mpb_team_get_iso_var(mpb_team *T, double *v) {
  API_BEGIN
  REQUIRE(T != nullptr && v != nullptr, "null argument");
  for (size_t i = 0; i < T->ctx.size(); i++) REQUIRE(mpb_get_iso_var(T->ctx[i], v + T->off[i]) == 0, g_err);
  API_END
}
int64_t mpb_team_get_np(mpb_team *T) { return T ? T->np : -1; }

int mpb_team_set_rng_ctr(mpb_team *T, uint64_t ctr) {
  API_BEGIN
  TEAM_EACH(mpb_set_rng_ctr(c, ctr));
  API_END
}
uint64_t mpb_team_get_rng_ctr(mpb_team *T) { return T && !T->ctx.empty() ? T->ctx[0]->rng_ctr : 0; }

// every launch of the step goes to all devices before the next one is issued; module_mixing is cut into its phases
int mpb_team_run_modules(mpb_team *T, double t, unsigned mask) {
  API_BEGIN
  REQUIRE(T != nullptr && !T->ctx.empty(), "null team");
  REQUIRE(T->ctx[0]->have_ctl, "mpb_team_set_ctl has not been called");
  const bool many = T->ctx.size() > 1;
  for (const Op &o : plan_modules(T->ctx[0]->ctl, t, mask)) {
    if (o.kind == Op::MIXING && many) {
      for (mpb_ctx *c : T->ctx) { use(c); mixing_prepare(c, t); mixing_route(c); }
      team_barrier(T);
      for (mpb_ctx *c : T->ctx) { use(c); mixing_serve(c); }
      team_barrier(T);
      for (mpb_ctx *c : T->ctx) { use(c); mixing_apply_routed(c); }
    } else {
      for (mpb_ctx *c : T->ctx) { use(c); run_op(c, t, o); }
    }
  }
  API_END
}
int mpb_team_run_timestep(mpb_team *T, double t) { return mpb_team_run_modules(T, t, MPB_MOD_ALL); }

int mpb_team_sync(mpb_team *T) {
  API_BEGIN
  TEAM_EACH(mpb_sync(c));
  API_END
}
int64_t mpb_team_launch_count(mpb_team *T) {
  long long n = 0;
  if (T) for (mpb_ctx *c : T->ctx) n += c->launches;
  return n;
}

// gridded output: partial arrays on every device, added up on the first one (mpb_team_grid_fetch reads them there)
int mpb_team_grid_accumulate(mpb_team *T, const mpb_grid_t *g) {
  API_BEGIN
  REQUIRE(T != nullptr && !T->ctx.empty() && g != nullptr, "null argument");
  const size_t nbox = (size_t)g->nx * g->ny * g->nz;
  team_ensure_area(T, T->mix_bytes, nbox * (16 * (size_t)std::max(T->nq, 1) + 4));
  TEAM_EACH(mpb_grid_accumulate(c, g));
  if (T->ctx.size() > 1) {
    team_barrier(T);
    use(T->ctx[0]);
    grid_pull(T->ctx[0]);
    team_barrier(T);
  }
  API_END
}
int mpb_team_grid_fetch(mpb_team *T, int *count, double *sum, double *sumsq) {
  API_BEGIN
  REQUIRE(T != nullptr && !T->ctx.empty(), "null team");
  REQUIRE(mpb_grid_fetch(T->ctx[0], count, sum, sumsq) == 0, g_err);
  API_END
}

}  // extern "C"
